"""
The VI warm start as a device path (irsgmcmc_b200/vi.py, csrc/irs_vi.cu; reference Trainer._run_VI, trainer/trainer.py:119-171)
against the same iteration through the drop-in modules and autograd (Trainer._run_VI(fused=False)), which is itself pinned to the
oracle / reference in tests/test_gpu_dropin.py::test_vi_sample_loss_parity.
"""
import math

import pytest
import torch

from tests.util import rel
from tests.test_gpu_dropin import _reference_style_config

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _trainer(n, reg_type, learnable, data='GMM', cps=None, lr=0.01):
    from irsgmcmc_b200.trainer import Trainer
    from irsgmcmc_b200.data_loader.synthetic import make_pair
    import irsgmcmc_b200.utils as U
    fixed, moving, vp0 = make_pair(n)
    cfg = _reference_style_config(2)
    cfg['reg_loss'] = {'type': reg_type, 'args': {'diff_op': 'GradientOperator', 'w_reg': 1.6 if reg_type == 'RegLoss_LogNormal' else 1.4,
                                                  'learnable': learnable}}
    if data == 'SSD':
        cfg['data_loss'] = {'type': 'SSD', 'args': {}}
        cfg['data_loss_proportion_prior'] = {'type': 'DirichletPrior', 'args': {'no_classes': 1, 'alpha': 0.5}}
    cfg['optimizer_q_v'] = {'type': 'Adam', 'args': {'lr_mu': lr, 'lr_log_var': lr, 'lr_u': lr, 'lr_decay': 0.001}}
    if cps:
        cfg['transformation_module'] = {'type': 'SVFFD_3D', 'args': {'cps': cps}}
        g = (1, 3, *U.get_control_grid_size((n,) * 3, cps))
        vp0 = {'mu': 0.3 * torch.randn(g), 'log_var': torch.full(g, math.log(0.5 ** 2)) + 0.1 * torch.randn(g),
               'u': 0.1 + 0.05 * torch.randn(g)}
    else:
        s = (1, 3, n, n, n)
        vp0 = {'mu': 0.3 * torch.randn(s), 'log_var': vp0['log_var'] + 0.1 * torch.randn(s), 'u': vp0['u'] + 0.05 * torch.randn(s)}
    return Trainer(cfg, fixed, moving, vp0, device=torch.device(DEV)), vp0


@pytest.mark.parametrize('reg_type,learnable,data,cps', [('RegLoss_LogNormal', True, 'GMM', None), ('RegLoss_L2', True, 'GMM', None),
                                                         ('RegLoss_L2', False, 'SSD', None), ('RegLoss_LogNormal', True, 'GMM', [4, 4, 4])])
def test_fused_vi_iteration_matches_the_autograd_path(built, reg_type, learnable, data, cps):
    n, iters = 16, 3
    torch.manual_seed(11)
    outs = {}
    for fused in (True, False):
        torch.manual_seed(12)
        t, vp0 = _trainer(n, reg_type, learnable, data, cps)
        m = t._build_VI_modules()
        m['data_loss'].init_parameters(0.7)
        gen = torch.Generator().manual_seed(5)
        shp = tuple(vp0['mu'].shape)
        noise = [(torch.randn(shp, generator=gen), torch.randn(1, generator=gen), torch.rand(1, 3, n, n, n, generator=gen),
                  torch.rand(1, 3, n, n, n, generator=gen)) for _ in range(iters)]
        # gradients: after ONE iteration Adam's first moment is (1 - beta1) * gradient
        vp1, _, hist1 = t._run_VI(vp0, no_iters=1, modules=m, noise=iter([tuple(x.to(DEV) for x in noise[0])]), fused=fused)
        if fused:
            grads = {k: t._vi._m[i].clone() / 0.1 for i, k in enumerate(('mu', 'log_var', 'u'))}
        else:
            grads = {k: t._optimizer_q_v.state[t._vp_leaves[k]]['exp_avg'].clone() / 0.1 for k in ('mu', 'log_var', 'u')}
        outs[fused] = {'grads': grads, 'hist1': hist1[0], 'vp1': {k: v.clone() for k, v in vp1.items()},
                       'log_std': m['data_loss'].log_std.detach().clone(), 'reg_p': [p.detach().clone() for p in t._reg_params(m['reg_loss'])],
                       'gmm_step': m['optimizer_GMM'].state[m['data_loss'].log_std]['step']}
    a, b = outs[True], outs[False]
    for k in ('data_samples', 'reg_samples'):
        assert rel(a['hist1'][k], b['hist1'][k]) < 1e-5, (k, a['hist1'][k], b['hist1'][k])
    assert abs(float(a['hist1']['entropy']) - float(b['hist1']['entropy'])) < 1e-5 * abs(float(b['hist1']['entropy']))
    for k in ('mu', 'log_var', 'u'):
        e = rel(a['grads'][k], b['grads'][k])
        print(reg_type, data, cps, 'grad', k, e)
        assert e < 2e-3, (k, e)            # both sides are fp32 with kink flips (tests/util.py::grad_ok); typically 1e-5
    assert rel(a['log_std'], b['log_std']) < 1e-5 and a['gmm_step'] == b['gmm_step'] == 2
    for pa, pb in zip(a['reg_p'], b['reg_p']):
        assert abs(float(pa) - float(pb)) < 1e-6 * max(1.0, abs(float(pb)))
    # the first Adam step moves every element by ~lr * sign(gradient): identical wherever the gradient is not at rounding level
    for k in ('mu', 'log_var', 'u'):
        same = ((a['vp1'][k] - b['vp1'][k]).abs() < 1e-4).float().mean()
        assert same > 0.995, (k, float(same))


def test_fused_vi_runs_in_a_graph_and_is_deterministic(built):
    n = 16
    res = []
    for use_graph in (False, True):
        torch.manual_seed(2)
        t, vp0 = _trainer(n, 'RegLoss_LogNormal', True)
        from irsgmcmc_b200.vi import VIWarmStart
        vi = VIWarmStart(t.fixed, t.moving, vp0, t.sampler.cfg, device=DEV)
        vi.sampler.init_gmm(sigma_hat=0.7)
        vi.step(5, use_graph=use_graph)
        torch.cuda.synchronize()
        res.append((vi.mu.clone(), vi.log_var.clone(), vi.u.clone(), vi.sampler.hyper.clone(), vi.vi_state.clone()))
    for x, y in zip(*res):
        assert torch.equal(x, y)
    assert float(res[0][4][0]) == 5 and float(res[0][3][56]) == 5     # Adam step counter, Philox offset of the jitter
    assert not torch.equal(res[0][0].cpu(), vp0['mu'])


def test_fused_vi_checkpoint_resume_is_bit_identical(built):
    """5 iterations in one go == 3 iterations, state_dict -> a fresh VIWarmStart -> 2 more (own Philox noise, graph replays)"""
    from irsgmcmc_b200.vi import VIWarmStart
    n = 16
    torch.manual_seed(2)
    t, vp0 = _trainer(n, 'RegLoss_LogNormal', True)

    def fresh():
        vi = VIWarmStart(t.fixed, t.moving, vp0, t.sampler.cfg, device=DEV)
        vi.sampler.init_gmm(sigma_hat=0.7)
        return vi

    a = fresh()
    a.step(5)
    b = fresh()
    b.step(3)
    sd = b.state_dict()
    assert sd['iteration'] == 3 and all(not v.is_cuda for v in (sd['mu'], sd['hyper'], sd['vi_state'], *sd['adam_m']))
    c = fresh()
    c.step(1)                      # its own history (and a captured graph) must not matter
    c.load_state_dict(sd)
    c.step(2)
    torch.cuda.synchronize()
    for x, y in ((a.mu, c.mu), (a.log_var, c.log_var), (a.u, c.u), (a.vi_state, c.vi_state), (a.sampler.hyper, c.sampler.hyper),
                 (a._m[1], c._m[1]), (a._v[2], c._v[2])):
        assert torch.equal(x, y)
    assert c.iteration == 5
    sd['seed'] += 1
    with pytest.raises(ValueError):
        c.load_state_dict(sd)


@pytest.mark.parametrize('reg_key', ['lognormal', 'l2'])
def test_fused_vi_iteration_vs_oracle(built, reg_key):
    """one fused VI iteration against the oracle's restatement of reference trainer/trainer.py:130-171 (oracle.vi_iteration,
    fp32 and fp64): loss terms, gradients w.r.t. (mu, log_var, u) read back from Adam's first moments, the mixture after its
    two steps and the regulariser's hyper-parameters after theirs"""
    from oracle import sgld_oracle as O
    from tests.util import grad_ok
    from irsgmcmc_b200.sampler import SGLDConfig
    from irsgmcmc_b200.vi import VIWarmStart
    from irsgmcmc_b200.data_loader.synthetic import make_pair
    n = 16
    torch.manual_seed(31)
    fixed, moving, vp0 = make_pair(n)
    s = (1, 3, n, n, n)
    vp0 = {'mu': 0.3 * torch.randn(s), 'log_var': vp0['log_var'] + 0.1 * torch.randn(s), 'u': vp0['u'] + 0.05 * torch.randn(s)}
    eps, x = torch.randn(s), torch.randn(1)
    j1, j2 = torch.rand(s), torch.rand(s)
    cfg = SGLDConfig(reg_loss='RegLoss_LogNormal' if reg_key == 'lognormal' else 'RegLoss_L2', w_reg=1.6, reg_learnable=True)
    vi = VIWarmStart(fixed, moving, vp0, cfg, device=DEV)
    vi.sampler.init_gmm(sigma_hat=0.7)
    vi.set_noise(eps, x, torch.cat((j1, j2), 0))
    vi.step(1, use_graph=False)
    torch.cuda.synchronize()
    lt = vi.loss_terms()
    res = {}
    for dtype in (torch.float32, torch.float64):
        st = O.State(O.Config(reg=reg_key, w_reg=1.6, exact_grid=dtype == torch.float64), torch.zeros(1, 3, n, n, n, dtype=dtype),
                     torch.ones(1, 3, n, n, n, dtype=dtype), (n, n, n), dtype)
        st.init_gmm(0.7)
        cast = lambda d_: {k: (v.to(dtype) if v.dtype == torch.float32 else v) for k, v in d_.items()}
        terms, grads = O.vi_iteration(st, cast(fixed), cast(moving), {k: v.to(dtype) for k, v in vp0.items()}, eps.to(dtype),
                                      x.to(dtype), j1.to(dtype), j2.to(dtype))
        res[dtype] = (terms, grads, st)
    t64, g64, st64 = res[torch.float64]
    _, g32, _ = res[torch.float32]
    assert abs(lt['entropy_sample'] + lt['entropy_log_det'] - float(t64['entropy'])) < 1e-5 * abs(float(t64['entropy']))
    assert abs(float(lt['alpha'][0]) - float(t64['alpha'][0])) < 1e-4 and abs(float(lt['alpha'][1]) - float(t64['alpha'][1])) < 1e-4
    for i, k in enumerate(('mu', 'log_var', 'u')):
        assert grad_ok(vi._m[i] / 0.1, g32[k], g64[k], f'fused VI grad {k} ({reg_key})')
    ls, lg = vi.sampler.gmm_parameters()
    assert rel(ls, st64.log_std) < 1e-4 and (lg.double() - st64.logits.double()).abs().max() < 1e-4
    want = torch.stack((st64.loc, st64.log_scale)) if reg_key == 'lognormal' else st64.log_w_reg.view(1)
    assert rel(vi.sampler.reg_parameters()[:want.numel()], want) < 1e-6
