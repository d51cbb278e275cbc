"""
Pins the oracle to the UNMODIFIED reference imported from /root/reference (build container only; skipped elsewhere --
tests/test_golden.py carries the same comparison through committed vectors): three consecutive
Trainer._SGLD_transition calls with injected noise, fp32 and fp64, both regularisers.
"""
import math
import warnings

import pytest
import torch
import torch.nn.functional as F

from oracle import ref_import, sgld_oracle as O
from tests.util import rel

pytestmark = pytest.mark.skipif(not ref_import.available(), reason='reference checkout not present')


@pytest.fixture(scope='module')
def ref():
    warnings.filterwarnings('ignore')
    r = ref_import.load()
    # several tests replace the reference's two noise functions to inject numbers; the loop tests need the real ones back
    if not hasattr(r, 'real_noise'):
        r.real_noise = {name: getattr(r.util, name) for name in ('get_noise_Langevin', 'get_noise_uniform')}
    return r


@pytest.mark.parametrize('dtype', [torch.float32, torch.float64])
@pytest.mark.parametrize('reg_name,reg_type', [('lognormal', 'RegLoss_LogNormal'), ('l2', 'RegLoss_L2')])
def test_transition_matches_reference(ref, dtype, reg_name, reg_type):
    from irsgmcmc_b200.data_loader.synthetic import make_pair
    n, C = 14, 2
    torch.manual_seed(123)
    fixed, moving, vp = make_pair(n)
    t = ref_import.make_trainer(ref, (n, n, n), C, reg_type=reg_type, w_reg=1.6, uniform_noise=0.1,
                                dtype=dtype if dtype == torch.float64 else None)
    if dtype == torch.float64:  # RegistrationModule rejects fp64 (reference utils/registration.py:13-15,32)
        t.registration_module = lambda im, T: F.grid_sample(im, T.permute(0, 2, 3, 4, 1), mode='bilinear',
                                                            padding_mode='border', align_corners=True)
    gmm, reg = t.losses['data']['loss'], t.losses['reg']['loss']
    gmm.init_parameters(torch.tensor(1.0))
    cast = lambda d: {k: (v.to(dtype) if v.dtype == torch.float32 else v).expand(C, *v.shape[1:]) for k, v in d.items()}
    fx, mv = cast(fixed), cast(moving)
    v0 = (1.0 * torch.randn(C, 3, n, n, n)).to(dtype)
    sigma = (0.5 + torch.rand(1, 3, n, n, n)).to(dtype).expand(C, -1, -1, -1, -1)
    ref_import.attach_state(t, v0, sigma, 0.4)
    st = O.State(O.Config(reg=reg_name, w_reg=1.6), v0, sigma, (n, n, n), dtype)
    st.init_gmm(1.0)
    f_tol, g_tol = (1e-5, 5e-4) if dtype == torch.float32 else (1e-12, 1e-6)
    for it in range(3):
        eps, ju = torch.randn(C, 3, n, n, n).to(dtype), torch.rand(C, 3, n, n, n).to(dtype)
        ref.util.get_noise_Langevin = lambda s, tau, e=eps: math.sqrt(2.0 * tau) * s * e
        ref.util.get_noise_uniform = lambda shape, device, alpha, j=ju: -2.0 * alpha * j + alpha
        st.v = t.v_curr_state.detach().clone()     # same starting point every iteration: errors do not compound
        v_before = st.v.clone()
        lt, out, aux = t._SGLD_transition(fx, mv, gmm, reg)
        lt2, out2, aux2, grad_v = O.sgld_transition(st, {k: v[:1] for k, v in fx.items()}, {k: v[:1] for k, v in mv.items()},
                                                    eps, ju)
        for key in ('curr_state', 'transformation', 'displacement', 'im_moving_warped'):
            assert rel(out2[key], out[key]) < f_tol, (it, key)
        mask = fixed['mask'].expand(C, -1, -1, -1, -1)
        assert rel(aux2['residuals'][mask].view(C, -1), aux['residuals']) < 30 * f_tol
        assert rel(torch.stack(aux2['alpha']), torch.stack([a.detach() for a in aux['alpha']])) < 1e-4
        assert rel(torch.stack(lt2['data']), torch.stack([a.detach() for a in lt['data']])) < 1e-4
        assert rel(torch.stack(lt2['reg']), torch.stack([a.detach() for a in lt['reg']])) < 1e-6
        assert rel(grad_v, (v_before - t.v_curr_state.detach()) / 0.4) < g_tol
        # the shared hyper-parameters follow the reference from here on
        assert rel(st.log_std, gmm.log_std.detach()) < 1e-4 and rel(st.logits, gmm.logits.detach()) < 1e-3
        st.log_std.copy_(gmm.log_std.detach())
        st.logits.copy_(gmm.logits.detach())
        for dst, p in zip(st.adam_gmm.m + st.adam_gmm.v, [t.optimizer_GMM.state[q][k] for k in ('exp_avg', 'exp_avg_sq')
                                                            for q in (gmm.log_std, gmm.logits)]):
            dst.copy_(p)
        if reg_name == 'lognormal':
            assert abs(float(st.loc) - float(reg.loc)) < 1e-6 and abs(float(st.log_scale) - float(reg.log_scale)) < 1e-6
        else:
            assert abs(float(st.log_w_reg) - float(reg.log_w_reg)) < 1e-6


@pytest.mark.parametrize('dtype', [torch.float32, torch.float64])
def test_svffd_transition_matches_reference(ref, dtype):
    """the same with SVFFD_3D as the transformation module (configs/experiment5): the chain state on the control grid"""
    from irsgmcmc_b200.data_loader.synthetic import make_pair
    n, C, cps = 16, 2, (4, 4, 4)
    grid = O.control_grid_size((n,) * 3, cps)
    torch.manual_seed(321)
    fixed, moving, _ = make_pair(n)
    t = ref_import.make_trainer(ref, (n, n, n), C, reg_type='RegLoss_LogNormal', w_reg=1.6, uniform_noise=0.1,
                                dtype=dtype if dtype == torch.float64 else None, cps=cps)
    if dtype == torch.float64:
        t.registration_module = lambda im, T: F.grid_sample(im, T.permute(0, 2, 3, 4, 1), mode='bilinear',
                                                            padding_mode='border', align_corners=True)
    gmm, reg = t.losses['data']['loss'], t.losses['reg']['loss']
    gmm.init_parameters(torch.tensor(1.0))
    cast = lambda d: {k: (v.to(dtype) if v.dtype == torch.float32 else v).expand(C, *v.shape[1:]) for k, v in d.items()}
    fx, mv = cast(fixed), cast(moving)
    v0 = (1.5 * torch.randn(C, 3, *grid)).to(dtype)
    sigma = (0.5 + torch.rand(1, 3, *grid)).to(dtype).expand(C, -1, -1, -1, -1)
    ref_import.attach_state(t, v0, sigma, 0.4)
    st = O.State(O.Config(reg='lognormal', w_reg=1.6, cps=cps), v0, sigma, (n, n, n), dtype)
    st.init_gmm(1.0)
    f_tol, g_tol = (1e-5, 5e-4) if dtype == torch.float32 else (1e-12, 1e-6)
    for it in range(2):
        eps, ju = torch.randn(C, 3, *grid).to(dtype), torch.rand(C, 3, n, n, n).to(dtype)
        ref.util.get_noise_Langevin = lambda s, tau, e=eps: math.sqrt(2.0 * tau) * s * e
        ref.util.get_noise_uniform = lambda shape, device, alpha, j=ju: -2.0 * alpha * j + alpha
        st.v = t.v_curr_state.detach().clone()
        v_before = st.v.clone()
        lt, out, aux = t._SGLD_transition(fx, mv, gmm, reg)
        lt2, out2, aux2, grad_v = O.sgld_transition(st, {k: v[:1] for k, v in fx.items()}, {k: v[:1] for k, v in mv.items()},
                                                    eps, ju)
        assert out['curr_state'].shape == (C, 3, *grid) and out['displacement'].shape == (C, 3, n, n, n)
        for key in ('curr_state', 'transformation', 'displacement', 'im_moving_warped'):
            assert rel(out2[key], out[key]) < f_tol, (it, key)
        assert rel(torch.stack(lt2['data']), torch.stack([a.detach() for a in lt['data']])) < 1e-4
        assert rel(torch.stack(lt2['reg']), torch.stack([a.detach() for a in lt['reg']])) < 1e-6
        assert rel(torch.stack(aux2['reg_energy']), torch.stack([a.detach() for a in aux['reg_energy']])) < 1e-6
        assert rel(grad_v, (v_before - t.v_curr_state.detach()) / 0.4) < g_tol
        assert rel(st.log_std, gmm.log_std.detach()) < 1e-4 and rel(st.logits, gmm.logits.detach()) < 1e-3
        st.log_std.copy_(gmm.log_std.detach())
        st.logits.copy_(gmm.logits.detach())
        for dst, p in zip(st.adam_gmm.m + st.adam_gmm.v, [t.optimizer_GMM.state[q][k] for k in ('exp_avg', 'exp_avg_sq')
                                                            for q in (gmm.log_std, gmm.logits)]):
            dst.copy_(p)
        assert abs(float(st.loc) - float(reg.loc)) < 1e-6 and abs(float(st.log_scale) - float(reg.log_scale)) < 1e-6


def test_reference_adam_matches_oracle_adam(ref):
    torch.manual_seed(0)
    p_ref = [torch.nn.Parameter(torch.randn(4)), torch.nn.Parameter(torch.randn(4))]
    p_or = [p.detach().clone() for p in p_ref]
    opt = ref.optim.Adam([{'params': [p_ref[0]], 'lr': 0.2}, {'params': [p_ref[1]], 'lr': 0.05}], lr_decay=1e-3)
    mine = O.AdamState(p_or, [0.2, 0.05], 1e-3)
    for _ in range(5):
        g = [torch.randn(4), torch.randn(4)]
        for p, gg in zip(p_ref, g):
            p.grad = gg.clone()
        opt.step()
        mine.step(g)
        for a, b in zip(p_ref, p_or):
            assert torch.allclose(a.detach(), b, atol=1e-7)


def test_dropin_adam_matches_reference_adam(ref):
    from irsgmcmc_b200.optimizers import Adam
    torch.manual_seed(1)
    a = [torch.nn.Parameter(torch.randn(3)), torch.nn.Parameter(torch.randn(()))]
    b = [torch.nn.Parameter(p.detach().clone()) for p in a]
    oa = ref.optim.Adam([{'params': [a[0]], 'lr': 0.2}, {'params': [a[1]], 'lr': 0.01}], lr_decay=1e-3)
    ob = Adam([{'params': [b[0]], 'lr': 0.2}, {'params': [b[1]], 'lr': 0.01}], lr_decay=1e-3)
    for _ in range(6):
        for pa, pb in zip(a, b):
            g = torch.randn_like(pa)
            pa.grad, pb.grad = g.clone(), g.clone()
        oa.step()
        ob.step()
    for pa, pb in zip(a, b):
        assert torch.equal(pa.detach(), pb.detach())


def test_dropin_distributions_match_reference(ref):
    import irsgmcmc_b200.model.distributions as D
    R = ref.distr
    x = torch.tensor([-1.3, 0.2, 2.5])
    assert torch.allclose(D.LogScaleNormalPrior(0.0, 2.3)(x), R.LogScaleNormalPrior(0.0, 2.3)(x))
    lp = torch.log_softmax(torch.randn(4), 0)
    assert torch.allclose(D.DirichletPrior(4, 0.5)(lp), R.DirichletPrior(4, 0.5)(lp))
    dof = 3.0 * 32 ** 3
    ly = torch.tensor([10.2, 11.0])
    assert torch.allclose(D.LogEnergyExpGammaPrior(1.6, dof)(ly), R.LogEnergyExpGammaPrior(1.6, dof)(ly))
    assert torch.allclose(D.LogEnergyExpGammaPrior(1.6, dof).expectation(), R.LogEnergyExpGammaPrior(1.6, dof).expectation())
    lw = torch.tensor(0.4)
    assert torch.allclose(D.LogPrecisionExpGammaPrior(shape=0.5 * dof, rate=2.0 / dof)(lw),
                          R.LogPrecisionExpGammaPrior(shape=0.5 * dof, rate=2.0 / dof)(lw))


@pytest.mark.parametrize('reg_name,reg_type', [('lognormal', 'RegLoss_LogNormal'), ('l2', 'RegLoss_L2')])
def test_vi_sample_loss_matches_reference(ref, reg_name, reg_type):
    """the VI per-sample loss (reference Trainer.__calc_sample_loss_VI, trainer/trainer.py:79-117) and its gradients"""
    from irsgmcmc_b200.data_loader.synthetic import make_pair
    n = 14
    torch.manual_seed(123)
    fixed, moving, vp0 = make_pair(n)
    t = ref_import.make_trainer(ref, (n, n, n), 1, reg_type=reg_type, w_reg=1.6, uniform_noise=0.1)
    t.losses['entropy'] = ref.loss.EntropyMultivariateNormal()
    gmm, reg = t.losses['data']['loss'], t.losses['reg']['loss']
    gmm.init_parameters(torch.tensor(1.0))
    st = O.State(O.Config(reg=reg_name, w_reg=1.6), torch.zeros(1, 3, n, n, n), torch.ones(1, 3, n, n, n), (n, n, n))
    st.init_gmm(1.0)
    vp_ref = {k: v.clone().requires_grad_(True) for k, v in vp0.items()}
    vp_or = {k: v.clone().requires_grad_(True) for k, v in vp0.items()}
    eps, x, ju = torch.randn(1, 3, n, n, n), torch.randn(1), torch.rand(1, 3, n, n, n)
    ref.util.get_noise_uniform = lambda shape, device, alpha, j=ju: -2.0 * alpha * j + alpha
    sample_ref = vp_ref['mu'] + eps * torch.exp(0.5 * vp_ref['log_var']) + x * vp_ref['u']
    lt, out, aux = t._Trainer__calc_sample_loss_VI(gmm, reg, t.losses['entropy'], fixed, moving, vp_ref, sample_ref)
    sample_or = vp_or['mu'] + eps * torch.exp(0.5 * vp_or['log_var']) + x * vp_or['u']
    if reg_name == 'lognormal':
        leaves = (st.loc.clone().requires_grad_(True), st.log_scale.clone().requires_grad_(True))
    else:
        leaves = (st.log_w_reg.clone().requires_grad_(True),)
    terms = O.vi_sample_loss(st, fixed, moving, vp_or, sample_or, ju, leaves)
    assert rel(terms['im_w'], out['im_moving_warped']) < 1e-5 and rel(terms['disp'], out['displacement']) < 1e-5
    assert abs(float(terms['alpha']) - float(aux['alpha'])) < 1e-5
    for key in ('data', 'reg', 'entropy'):
        assert rel(terms[key], lt[key]) < 1e-4, key
    prior_key = 'reg_loc_prior' if reg_name == 'lognormal' else 'w_reg_prior'
    assert rel(terms[prior_key], lt[prior_key]) < 1e-6
    loss_ref = lt['data'] + lt['reg'] - lt['entropy']
    loss_or = terms['data'] + terms['reg'] - terms['entropy']
    g_ref = torch.autograd.grad(loss_ref, [vp_ref[k] for k in ('mu', 'log_var', 'u')])
    g_or = torch.autograd.grad(loss_or, [vp_or[k] for k in ('mu', 'log_var', 'u')])
    for a, b in zip(g_or, g_ref):
        assert rel(a, b) < 5e-4
    assert rel(st.log_std, gmm.log_std.detach()) < 1e-4


@pytest.mark.parametrize('init', ['VI', 'identity', 'noise'])
def test_chain_initialisation_matches_reference(ref, init):
    """A16: Trainer.__SGLD_init / sample_q_v of the unmodified reference (trainer/trainer.py:585-611, utils/sampler.py:4-21)
    against draw_chain_states (what SGLDSampler.init_chains runs) for the same seeded default generator: bit-identical
    states and preconditioner; a shard of the chains reproduces the same states for its global chain ids"""
    from irsgmcmc_b200.utils.sampler import draw_chain_states, sample_q_v
    n, C = 10, 5
    g = torch.Generator().manual_seed(3)
    vp = {'mu': torch.randn(1, 3, n, n, n, generator=g), 'log_var': torch.randn(1, 3, n, n, n, generator=g) - 1.0,
          'u': 0.1 * torch.randn(1, 3, n, n, n, generator=g)}
    t = ref.trainer.Trainer.__new__(ref.trainer.Trainer)
    t.device, t.no_chains, t.MCMC_init = 'cpu', C, init
    t.config = {'optimizer_SG_MCMC': {'args': {'lr': 0.4}}}
    t._Trainer__init_optimizer_SG_MCMC = lambda: None
    torch.manual_seed(77)
    t._Trainer__SGLD_init({k: v.clone() for k, v in vp.items()})
    torch.manual_seed(77)
    v, sigma = draw_chain_states(init, vp, C)
    assert torch.equal(v, t.v_curr_state.detach())
    ref_sigma = t.SGLD_params['sigma']
    assert torch.equal(ref_sigma, torch.ones_like(ref_sigma) if sigma is None else sigma.expand_as(ref_sigma))
    assert t.SGLD_params['tau'] == 0.4
    # sharded: chains 2..4 of 5 on "rank 1" draw what the single process gave those global ids
    torch.manual_seed(77)
    v_shard, _ = draw_chain_states(init, vp, 3, chain_offset=2, no_chains_total=C)
    assert torch.equal(v_shard, v[2:])
    # the mirror of utils/sampler.py itself, single draw and antithetic pair
    for k in (1, 2):
        torch.manual_seed(5)
        a = ref.sampler.sample_q_v(vp, no_samples=k)
        torch.manual_seed(5)
        b = sample_q_v(vp, no_samples=k)
        assert all(torch.equal(x, y) for x, y in zip(a if k == 2 else (a,), b if k == 2 else (b,)))


@pytest.mark.parametrize('reg_name,reg_type', [('lognormal', 'RegLoss_LogNormal'), ('l2', 'RegLoss_L2')])
def test_vi_iterations_match_reference_run_VI(ref, reg_name, reg_type, monkeypatch):
    """
    The unmodified Trainer._run_VI of the reference (trainer/trainer.py:119-223) for three iterations -- its own sample_q_v
    draws, both antithetic sample losses, hyper-priors, the entropy terms, loss.backward(), the reference's Adam on
    (mu, log_var, u) and on the regulariser's hyper-parameters, the mixture stepped twice per iteration -- against the oracle's
    vi_iteration + AdamState fed the same random numbers.  Only I/O is stubbed (file writers, TensorBoard, SimpleITK metrics).
    """
    import types
    import numpy as np
    import trainer.trainer as ref_trainer_module
    from irsgmcmc_b200.data_loader.synthetic import make_pair
    n, iters, lr = 12, 3, 0.01
    torch.manual_seed(123)
    fixed, moving, vp0 = make_pair(n)
    shape = (1, 3, n, n, n)
    vp0 = {'mu': 0.2 * torch.randn(shape), 'log_var': vp0['log_var'] + 0.1 * torch.randn(shape), 'u': vp0['u'] + 0.05 * torch.randn(shape)}
    jitters = [torch.rand(shape) for _ in range(2 * iters + 2)]

    t = ref_import.make_trainer(ref, (n, n, n), 1, reg_type=reg_type, w_reg=1.6, uniform_noise=0.1)
    t.losses['entropy'] = ref.loss.EntropyMultivariateNormal()
    gmm, reg = t.losses['data']['loss'], t.losses['reg']['loss']
    gmm.init_parameters(torch.tensor(1.0))
    Adam = ref.optim.Adam
    t.config = types.SimpleNamespace(init_optimizer_q_v=lambda vp: Adam(
        [{'params': [vp['mu']], 'lr': lr}, {'params': [vp['log_var']], 'lr': lr}, {'params': [vp['u']], 'lr': lr}], lr_decay=1e-3))
    t.start_iter_VI, t.no_iters_VI, t.log_period_VI = 1, iters, 10 ** 9
    t.save_dirs, t.im_spacing, t.structures_dict = None, (1.0, 1.0, 1.0), {}
    t.writer = types.SimpleNamespace(set_step=lambda *a, **k: None)
    t.metrics = types.SimpleNamespace(update=lambda *a, **k: None)
    for name in ('save_fixed_im', 'save_fixed_mask', 'save_moving_im', 'save_moving_mask', 'log_hist_res', 'log_images', 'log_fields'):
        monkeypatch.setattr(ref_trainer_module, name, lambda *a, **k: None)
    monkeypatch.setattr(ref_trainer_module, 'calc_metrics', lambda *a, **k: (np.zeros((1, 0)), np.zeros((1, 0))))
    it_j = iter(jitters)
    monkeypatch.setattr(ref.util, 'get_noise_uniform', lambda shape, device, alpha: -2.0 * alpha * next(it_j) + alpha)

    vp_ref = {k: v.clone() for k, v in vp0.items()}
    torch.manual_seed(77)
    t._run_VI(fixed, moving, vp_ref)

    # the oracle with the same numbers: per iteration randn_like(sigma), randn(1) (utils/sampler.py:14-15), two jitter fields
    torch.manual_seed(77)
    st = O.State(O.Config(reg=reg_name, w_reg=1.6), torch.zeros(shape), torch.ones(shape), (n, n, n))
    st.init_gmm(1.0)
    vp = {k: v.clone() for k, v in vp0.items()}
    adam = O.AdamState([vp['mu'], vp['log_var'], vp['u']], [lr, lr, lr], 1e-3)
    for i in range(iters):
        eps, x = torch.randn(shape), torch.randn(1)
        terms, grads = O.vi_iteration(st, fixed, moving, vp, eps, x, jitters[2 * i], jitters[2 * i + 1])
        adam.step([grads['mu'], grads['log_var'], grads['u']])
    # Adam's first steps have size lr whatever the gradient: compare the UPDATES, not the parameters
    for k in ('mu', 'log_var', 'u'):
        d_ref, d_or = vp_ref[k].detach() - vp0[k], vp[k] - vp0[k]
        assert float(d_ref.abs().max()) > 0.5 * lr
        assert rel(d_or, d_ref) < 1e-4, (k, rel(d_or, d_ref))   # measured 1.0e-5 ... 1.2e-5 (fp32, three iterations)
    assert rel(st.log_std, gmm.log_std.detach()) < 1e-4 and rel(st.logits, gmm.logits.detach()) < 1e-3
    if reg_name == 'lognormal':
        assert abs(float(st.loc) - float(reg.loc)) < 1e-6 * abs(float(reg.loc)) and abs(float(st.log_scale) - float(reg.log_scale)) < 1e-6
    else:
        assert abs(float(st.log_w_reg) - float(reg.log_w_reg)) < 1e-6
    assert t.optimizer_q_v.state[vp_ref['mu']]['step'] == iters == adam.step_no


@pytest.mark.parametrize('dtype', [torch.float64, torch.float32])
def test_run_MCMC_loop_matches_reference(ref, monkeypatch, dtype):
    """
    The unmodified Trainer._run_MCMC of the reference (trainer/trainer.py:358-476): chain initialisation from q(v), burn-in, the
    kept-sample rule, the host buffer of displacement samples and calc_posterior_statistics -- with its OWN random draws
    (sample_q_v per chain, randn_like for the Langevin noise, rand for the jitter) -- against the oracle loop fed the same
    generator: sgld_transition per iteration, kept samples by the same rule, posterior_statistics.  Only I/O is stubbed.
    fp64 (the reference run under torch.set_default_dtype(float64)): the two loops must agree to rounding.  fp32: eight chained
    transitions compound the isolated cell-face flips of the interpolation gradients (SURVEY surprise 9), so the
    bound there only says "same random numbers, same loop" (a wrong draw order or kept-sample rule gives O(1)).
    """
    import types
    import numpy as np
    import trainer.trainer as ref_trainer_module
    from irsgmcmc_b200.data_loader.synthetic import make_pair
    n, C, burn_in, no_samples, period, tau = 12, 2, 2, 6, 2, 0.4
    torch.manual_seed(123)
    fixed, moving, vp0 = make_pair(n)
    shape = (1, 3, n, n, n)
    vp0 = {'mu': 0.2 * torch.randn(shape), 'log_var': vp0['log_var'] + 0.1 * torch.randn(shape), 'u': vp0['u'] + 0.05 * torch.randn(shape)}
    cast = lambda d: {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in d.items()}
    fixed, moving, vp0 = cast(fixed), cast(moving), cast(vp0)
    old_default = torch.get_default_dtype()
    torch.set_default_dtype(dtype)      # the reference allocates its state and draws its noise in the default dtype
    try:
        t = ref_import.make_trainer(ref, (n, n, n), C, reg_type='RegLoss_LogNormal', w_reg=1.6, uniform_noise=0.1, tau=tau,
                                    dtype=dtype if dtype == torch.float64 else None)
        if dtype == torch.float64:  # RegistrationModule rejects fp64 (reference utils/registration.py:13-15,32)
            t.registration_module = lambda im, T: F.grid_sample(
                im if im.is_floating_point() else im.to(dtype), T.permute(0, 2, 3, 4, 1), mode='bilinear' if im.is_floating_point() else 'nearest',
                padding_mode='border', align_corners=True)
        gmm, reg = t.losses['data']['loss'], t.losses['reg']['loss']
        gmm.init_parameters(torch.tensor(1.0))

        class _Config(dict):
            def init_obj(self, name, module, *args, **kwargs):      # parse_config.py:251-266
                return getattr(module, self[name]['type'])(*args, **self[name]['args'], **kwargs)

        for name, fn in ref.real_noise.items():     # the reference's own randn_like / rand draws
            monkeypatch.setattr(ref.util, name, fn)
        t.config = _Config(optimizer_SG_MCMC={'type': 'SGD', 'args': {'lr': tau}})
        t.MCMC_init, t.no_iters_burn_in, t.no_samples_MCMC, t.log_period_MCMC = 'VI', burn_in, no_samples, period
        t.dims, t.no_voxels = (n, n, n), n ** 3
        t.save_dirs, t.im_spacing, t.structures_dict = None, (1.0, 1.0, 1.0), {}
        t.writer = types.SimpleNamespace(set_step=lambda *a, **k: None)
        t.metrics = types.SimpleNamespace(update=lambda *a, **k: None)
        t.logger = types.SimpleNamespace(info=lambda *a, **k: None)
        for name in ('log_sample', 'log_hist_res', 'save_sample', 'log_displacement_mean_and_std_dev'):
            monkeypatch.setattr(ref_trainer_module, name, lambda *a, **k: None)
        monkeypatch.setattr(ref_trainer_module, 'calc_metrics', lambda *a, **k: (np.zeros((C, 0)), np.zeros((C, 0))))
        # utils/util.py:114-120 defaults to device='cuda:0': the reference's own function, called with its own `device` argument
        real_stats = ref_trainer_module.calc_posterior_statistics
        monkeypatch.setattr(ref_trainer_module, 'calc_posterior_statistics', lambda samples: real_stats(samples, device='cpu'))
        got = {}
        monkeypatch.setattr(ref_trainer_module, 'save_displacement_mean_and_std_dev',
                            lambda logger, dirs, spacing, mean, std, mask, model: got.update(mean=mean.clone(), std=std.clone()))
        # the built-in speed test (100 more transitions, :467-476) runs after the statistics: stop it at its first transition
        calls = {'n': 0}
        real_transition = t._SGLD_transition

        class _Stop(Exception):
            pass

        def counted(*a, **k):
            calls['n'] += 1
            if calls['n'] > burn_in + no_samples:
                raise _Stop()
            return real_transition(*a, **k)

        t._SGLD_transition = counted
        torch.manual_seed(77)
        with pytest.raises(_Stop):
            t._run_MCMC(fixed, moving, {k: v.clone() for k, v in vp0.items()})
        assert calls['n'] == burn_in + no_samples + 1 and 'mean' in got

        # the oracle loop with the same generator: per chain randn_like(sigma), randn(1); per transition randn, rand of (C,3,...)
        torch.manual_seed(77)
        sigma1 = torch.exp(0.5 * vp0['log_var'])
        v0 = torch.cat([vp0['mu'] + torch.randn(shape) * sigma1 + torch.randn(1) * vp0['u'] for _ in range(C)], 0)
        st = O.State(O.Config(reg='lognormal', w_reg=1.6, tau=tau, exact_grid=dtype == torch.float64), v0,
                     sigma1.expand(C, -1, -1, -1, -1), (n, n, n), dtype)
        st.init_gmm(1.0)
        kept = []
        for sample_no in range(1, burn_in + no_samples + 1):
            eps, ju = torch.randn(C, 3, n, n, n), torch.rand(C, 3, n, n, n)
            _, out, _, _ = O.sgld_transition(st, fixed, moving, eps, ju)
            if sample_no > burn_in and (sample_no % period == 0 or sample_no == no_samples):     # trainer.py:414-415
                kept.extend(out['displacement'][c] for c in range(C))
    finally:
        torch.set_default_dtype(old_default)
    assert len(kept) == C * no_samples // period       # the reference's buffer size (:365-366): every row is filled
    mean, std = O.posterior_statistics(torch.stack(kept))
    e_v, e_mean, e_std = rel(st.v, t.v_curr_state.detach()), rel(mean, got['mean']), rel(std, got['std'])
    e_gmm = rel(st.log_std, gmm.log_std.detach())
    # measured: fp64 1.4e-13 / 7.5e-14 / 5.0e-14 / 1.5e-13; fp32 8.6e-3 / 7.4e-4 / 2.6e-4 / 3.5e-4
    tol_v, tol = (1e-10, 1e-10) if dtype == torch.float64 else (5e-2, 1e-2)
    assert e_v < tol_v and e_mean < tol and e_std < tol and e_gmm < tol, (e_v, e_mean, e_std, e_gmm)


def test_GMM_init_matches_reference(ref):
    """Trainer.__GMM_init of the reference (trainer/trainer.py:529-547: one q(v) draw -> residuals -> sigma_hat -> linspace
    initialisation -> VD factor -> 25 Adam steps on the mixture) against the oracle's gmm_init on the same draw"""
    from irsgmcmc_b200.data_loader.synthetic import make_pair
    n = 14
    torch.manual_seed(123)
    fixed, moving, vp0 = make_pair(n)
    t = ref_import.make_trainer(ref, (n, n, n), 1, reg_type='RegLoss_LogNormal', w_reg=1.6)
    gmm = t.losses['data']['loss']
    torch.manual_seed(5)
    t._Trainer__GMM_init(fixed, moving, {k: v.clone() for k, v in vp0.items()})
    torch.manual_seed(5)
    v_sample = vp0['mu'] + torch.randn_like(vp0['log_var']) * torch.exp(0.5 * vp0['log_var']) + torch.randn(1) * vp0['u']
    st = O.State(O.Config(reg='lognormal', w_reg=1.6), torch.zeros(1, 3, n, n, n), torch.ones(1, 3, n, n, n), (n, n, n))
    O.gmm_init(st, fixed, moving, v_sample)
    # fp32, 25 chained Adam steps: measured 1.7e-5 / 5.1e-7
    assert rel(st.log_std, gmm.log_std.detach()) < 1e-4 and rel(st.logits, gmm.logits.detach()) < 1e-4, \
        (rel(st.log_std, gmm.log_std.detach()), rel(st.logits, gmm.logits.detach()))
    assert t.optimizer_GMM.state[gmm.log_std]['step'] == 25 == st.adam_gmm.step_no
    for mine, theirs in zip(st.adam_gmm.m + st.adam_gmm.v, [t.optimizer_GMM.state[q][k] for k in ('exp_avg', 'exp_avg_sq')
                                                            for q in (gmm.log_std, gmm.logits)]):
        assert rel(mine, theirs) < 1e-3
