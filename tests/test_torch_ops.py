"""
torch.ops.irsgmcmc.* -- the operators as PyTorch custom ops (irsgmcmc_b200/torch_ops.py).
 - CPU: every op is registered with a schema; the dispatcher refuses CPU tensors (CUDA kernels only, no fallback); fake
   (meta) implementations propagate shapes for fake CUDA tensors; the autograd wiring is checked in a separate process in
   which the ops get oracle-backed CPU kernels (tests/helpers/torch_ops_cpu_wiring.py).
 - GPU: every op against the drop-in module / launcher wrapper it shares its kernels with, forward and backward.
"""
import os
import subprocess
import sys

import pytest
import torch

from oracle import sgld_oracle as O
from tests.util import rel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV = 'cuda:0'


def test_ops_are_registered_with_schemas():
    import irsgmcmc_b200  # noqa: F401
    from irsgmcmc_b200 import torch_ops
    for name in torch_ops.OPS:
        op = getattr(torch.ops.irsgmcmc, name)
        assert str(op.default._schema).startswith(f'irsgmcmc::{name}(')
    assert 'Tensor? jitter_unit' in str(torch.ops.irsgmcmc.warp3d.default._schema)
    assert str(torch.ops.irsgmcmc.svf_exp.default._schema).endswith('-> (Tensor, Tensor, Tensor)')


def test_dispatcher_refuses_cpu_tensors():
    """CUDA dispatch key only: there is no CPU kernel behind any op"""
    import irsgmcmc_b200  # noqa: F401
    n = 6
    v, im = torch.zeros(1, 3, n, n, n), torch.zeros(1, 1, n, n, n)
    calls = [lambda: torch.ops.irsgmcmc.warp3d(im, v, None, 0.0), lambda: torch.ops.irsgmcmc.svf_exp(v, 12),
             lambda: torch.ops.irsgmcmc.sobolev_smooth(v, [0.25, 0.5, 0.25]), lambda: torch.ops.irsgmcmc.lcc_normalise(im, 2),
             lambda: torch.ops.irsgmcmc.reg_energy(v), lambda: torch.ops.irsgmcmc.warp3d_nearest(im.short(), v),
             lambda: torch.ops.irsgmcmc.ffd(torch.zeros(1, 3, 5, 5, 5), [0.1] * 7, [0.1] * 7, [0.1] * 7, [2, 2, 2], [n, n, n])]
    for call in calls:
        with pytest.raises(NotImplementedError, match="'CPU' backend"):
            call()


def test_fake_implementations_propagate_shapes(built):
    """shape / dtype / device propagation without running a kernel (what torch.compile and export trace through)"""
    import irsgmcmc_b200  # noqa: F401
    from torch._subclasses.fake_tensor import FakeTensorMode
    n, C = 16, 3
    with FakeTensorMode():
        v = torch.empty(C, 3, n, n, n, device='cuda')
        im = torch.empty(1, 1, n, n, n, device='cuda')
        ops = torch.ops.irsgmcmc
        out = ops.warp3d(im, v, None, 0.1)
        assert out.shape == (C, 1, n, n, n) and out.device.type == 'cuda'
        assert ops.warp3d_bwd_grid(im, v, out, None, 0.1).shape == v.shape
        seg = torch.empty(1, 1, n, n, n, device='cuda', dtype=torch.int16)
        assert ops.warp3d_nearest(seg, v).dtype == torch.int16 and ops.warp3d_nearest(seg, v).shape == (C, 1, n, n, n)
        disp, hist, maxabs = ops.svf_exp(v, 12)
        assert disp.shape == v.shape and hist.shape == (12, C, 3, n, n, n) and maxabs.dim() == 1 and maxabs.numel() >= 12
        assert ops.svf_exp_bwd(v, hist, maxabs, disp, 2).shape == v.shape
        assert ops.sobolev_smooth(v, [1 / 6, 4 / 6, 1 / 6]).shape == v.shape
        zn, a, rs = ops.lcc_normalise(out, 2)
        assert zn.shape == a.shape == rs.shape == out.shape and ops.lcc_normalise_bwd(zn, a, rs, 2).shape == out.shape
        e = ops.reg_energy(v)
        assert e.shape == (C,) and e.dtype == torch.float64 and ops.reg_energy_grad(v, e).shape == v.shape
        cps, grid = [4, 4, 4], list(O.control_grid_size((n,) * 3, (4,) * 3))
        k = [0.1] * 15
        dense = ops.ffd(torch.empty(C, 3, *grid, device='cuda'), k, k, k, cps, [n, n, n])
        assert dense.shape == (C, 3, n, n, n) and ops.ffd_adjoint(dense, k, k, k, cps, grid).shape == (C, 3, *grid)


def test_autograd_wiring_with_oracle_backed_cpu_kernels():
    """in its own process: the ops get CPU kernels made of the oracle, forward / backward through the dispatcher must
    reproduce the oracle's autograd; plus torch.library.opcheck (schema, fake tensor, autograd registration)"""
    res = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'helpers', 'torch_ops_cpu_wiring.py')],
                         capture_output=True, text=True, timeout=600, env={**os.environ, 'CUDA_VISIBLE_DEVICES': ''})
    assert res.returncode == 0 and 'torch_ops wiring OK' in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]


# ---------------------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope='module')
def pkg(built):
    import irsgmcmc_b200.utils as U
    from irsgmcmc_b200 import ops
    return U, ops


def _field(C, n, amp, seed):
    from tests.util import smooth_field
    return smooth_field((C, 3, n, n, n), amp, seed).to(DEV)


@pytest.mark.gpu
def test_gpu_warp_ops_match_the_module(pkg):
    U, ops = pkg
    n, C = 20, 2
    torch.manual_seed(0)
    svf = U.SVF_3D((n, n, n)).to(DEV)
    T0, _ = svf(_field(C, n, 2.0, 1))
    img = torch.rand(1, 1, n, n, n, device=DEV)
    G = torch.randn(C, 1, n, n, n, device=DEV)
    T = T0.detach().clone().requires_grad_(True)
    out = torch.ops.irsgmcmc.warp3d(img, T, None, 0.0)
    (out * G).sum().backward()
    Tm = T0.detach().clone().requires_grad_(True)
    ref = U.RegistrationModule()(img.expand(C, -1, -1, -1, -1), Tm)
    (ref * G).sum().backward()
    assert torch.equal(out, ref) and torch.equal(T.grad, Tm.grad)
    ju = torch.rand(C, 3, n, n, n, device=DEV)
    assert torch.equal(torch.ops.irsgmcmc.warp3d(img, T0.detach(), ju, 0.1), ops.warp3d(img, T0.detach().contiguous(), ju, 0.1))
    seg = (torch.rand(1, 1, n, n, n, device=DEV) * 40).short()
    assert torch.equal(torch.ops.irsgmcmc.warp3d_nearest(seg, T0.detach()), ops.warp3d_nearest(seg, T0.detach().contiguous()))
    with pytest.raises(NotImplementedError):
        torch.ops.irsgmcmc.warp3d(img.cpu(), T0.detach().cpu(), None, 0.0)


@pytest.mark.gpu
def test_gpu_svf_and_smoothing_ops_match_the_modules(pkg):
    U, ops = pkg
    n, C = 20, 2
    v0 = _field(C, n, 1.0, 2)    # inside the deterministic gather regime of the adjoint
    G = torch.randn(C, 3, n, n, n, device=DEV, generator=torch.Generator(device=DEV).manual_seed(3))
    v = v0.clone().requires_grad_(True)
    disp, hist, maxabs = torch.ops.irsgmcmc.svf_exp(v, 12)
    (disp * G).sum().backward()
    vm = v0.clone().requires_grad_(True)
    _, dm = U.SVF_3D((n, n, n)).to(DEV)(vm)
    (dm * G).sum().backward()
    assert torch.equal(disp, dm) and torch.equal(hist[-1], dm) and rel(v.grad, vm.grad) < 1e-6
    taps = [float(t) for t in O.sobolev_taps(3, 0.5).astype('float32')]
    x = torch.randn(C, 3, n, n, n, device=DEV, requires_grad=True)
    y = torch.ops.irsgmcmc.sobolev_smooth(x, taps)
    (y * G).sum().backward()
    assert rel(y, O.sobolev_smooth(x.detach().cpu(), O.sobolev_taps(3, 0.5).astype("float32"))) < 1e-5
    assert torch.equal(x.grad, G)     # SobolevGrad.backward is the identity (reference utils/functions.py:107-109)


@pytest.mark.gpu
def test_gpu_lcc_energy_and_ffd_ops_match_the_launchers(pkg):
    U, ops = pkg
    n, C = 16, 2
    gen = torch.Generator(device=DEV).manual_seed(5)
    im = torch.rand(C, 1, n, n, n, device=DEV, generator=gen)
    G = torch.randn(C, 1, n, n, n, device=DEV, generator=gen)
    x = im.clone().requires_grad_(True)
    zn, a, rs = torch.ops.irsgmcmc.lcc_normalise(x, 2)
    (zn * G).sum().backward()
    zn_r, a_r, rs_r = ops.lcc_normalise(im, 2)
    assert torch.equal(zn, zn_r) and torch.equal(x.grad, ops.lcc_normalise_bwd(G, a_r, rs_r, 2))
    assert rel(zn, O.lcc_normalise(im.cpu(), 2)) < 1e-4   # wiring check; parity proper is tests/test_gpu_ops.py
    v = torch.randn(C, 3, n, n, n, device=DEV, generator=gen).requires_grad_(True)
    e = torch.ops.irsgmcmc.reg_energy(v)
    w = torch.tensor([0.5, -2.0], device=DEV, dtype=torch.float64)
    (e * w).sum().backward()
    assert e.dtype == torch.float64 and torch.equal(e, ops.reg_energy(v.detach()))
    assert torch.equal(v.grad, ops.reg_energy_grad(v.detach(), w)) and rel(e, O.reg_energy(v.detach().cpu().double())) < 1e-6
    cps, dims = (4, 4, 4), (n, n, n)
    m = U.Cubic_B_spline_FFD_3D(dims, cps).to(DEV)
    ks = [[float(t) for t in k] for k in m.kernels]
    cp0 = torch.randn(C, 3, *U.get_control_grid_size(dims, cps), device=DEV, generator=gen)
    Gd = torch.randn(C, 3, n, n, n, device=DEV, generator=gen)
    cp = cp0.clone().requires_grad_(True)
    dense = torch.ops.irsgmcmc.ffd(cp, ks[0], ks[1], ks[2], list(cps), list(dims))
    (dense * Gd).sum().backward()
    cm = cp0.clone().requires_grad_(True)
    dm = m(cm)
    (dm * Gd).sum().backward()
    assert torch.equal(dense, dm) and torch.equal(cp.grad, cm.grad)


@pytest.mark.gpu
def test_gpu_mixture_proposal_and_evaluation_ops(pkg):
    """the ops added in round 2 against the oracle and the wrappers they share kernels with"""
    U, ops = pkg
    import irsgmcmc_b200.model.loss as M
    n, C, K = 16, 2, 4
    gen = torch.Generator(device=DEV).manual_seed(7)
    # -- mixture log-density with autograd to z, log_std, logits: against torch autograd of the oracle's expression
    z0 = torch.randn(C, 1, n, n, n, device=DEV, generator=gen)
    ls0, lg0 = torch.linspace(-2.0, 0.5, K, device=DEV), torch.tensor([0.1, -0.2, 0.3, 0.0], device=DEV)
    Gp = torch.randn(C, 1, n, n, n, device=DEV, generator=gen)
    z, ls, lg = (t.clone().requires_grad_(True) for t in (z0, ls0, lg0))
    logp, dz = torch.ops.irsgmcmc.gmm_log_pdf(z, ls, lg)
    (logp * Gp).sum().backward()
    zr, lsr, lgr = (t.double().cpu().clone().requires_grad_(True) for t in (z0, ls0, lg0))
    ref = O.gmm_log_pdf(zr.reshape(-1), lsr, lgr).view(z0.shape)
    (ref * Gp.double().cpu()).sum().backward()
    assert rel(logp, ref) < 1e-5 and rel(z.grad, zr.grad) < 1e-5
    assert rel(ls.grad, lsr.grad) < 1e-4 and rel(lg.grad, lgr.grad) < 1e-4
    # -- virtual decimation factor
    mask = torch.rand(1, 1, n, n, n, device=DEV, generator=gen) > 0.3
    alpha = torch.ops.irsgmcmc.vd_factor(z0[:1], mask, ls0, lg0)
    assert alpha.dtype == torch.float64 and abs(float(alpha) - float(ops.vd_factor(z0[:1], mask, ls0, lg0))) == 0.0
    # -- Langevin proposal: explicit noise against the oracle, Philox reproducible, backward = sigma^2 g
    taps = [float(t) for t in O.sobolev_taps(3, 0.5).astype('float32')]
    v0 = torch.randn(C, 3, n, n, n, device=DEV, generator=gen)
    sigma = 0.5 + torch.rand(1, 3, n, n, n, device=DEV, generator=gen)
    eps = torch.randn(C, 3, n, n, n, device=DEV, generator=gen)
    G = torch.randn(C, 3, n, n, n, device=DEV, generator=gen)
    v = v0.clone().requires_grad_(True)
    out = torch.ops.irsgmcmc.langevin_proposal(v, sigma, eps, 0.9, taps, 0, 0, 0)
    (out * G).sum().backward()
    ref = O.sobolev_smooth((v0 + 0.9 * sigma * eps).cpu(), O.sobolev_taps(3, 0.5).astype('float32'))
    assert rel(out, ref) < 1e-5 and torch.equal(v.grad, G * sigma ** 2)
    a = torch.ops.irsgmcmc.langevin_proposal(v0, sigma, None, 0.9, taps, 11, 3, 5)
    b = torch.ops.irsgmcmc.langevin_proposal(v0, sigma, None, 0.9, taps, 11, 3, 5)
    c = torch.ops.irsgmcmc.langevin_proposal(v0, sigma, None, 0.9, taps, 11, 4, 5)
    assert torch.equal(a, b) and not torch.equal(a, c)
    # -- Welford moments (in place) against torch.mean / torch.std
    samples = torch.randn(7, 3, n, n, n, device=DEV, generator=gen) * 2 + 1
    mean, m2 = torch.zeros(3, n, n, n, device=DEV), torch.zeros(3, n, n, n, device=DEV)
    torch.ops.irsgmcmc.welford_update(samples[:4], 0, mean, m2)
    torch.ops.irsgmcmc.welford_update(samples[4:], 4, mean, m2)
    assert rel(mean, samples.mean(0)) < 1e-6 and rel(torch.ops.irsgmcmc.welford_std(m2, 7), samples.std(0)) < 1e-5
    # -- det J and Dice counts
    T, _ = U.SVF_3D((n, n, n)).to(DEV)(_field(C, n, 2.0, 9))
    counts, log_det = torch.ops.irsgmcmc.log_det_jacobian(T.detach())
    c2, l2 = ops.log_det_jacobian(T.detach().contiguous())
    assert torch.equal(counts, c2) and torch.equal(log_det, l2) and log_det.shape == (C, n, n, n)
    sa = (torch.rand(1, 1, n, n, n, device=DEV, generator=gen) * 4).short()
    sb = (torch.rand(C, 1, n, n, n, device=DEV, generator=gen) * 4).short()
    dc = torch.ops.irsgmcmc.dice_counts(sa, sb, [1, 2, 3])
    for i, lab in enumerate([1, 2, 3]):
        for c_ in range(C):
            assert tuple(dc[c_, i].tolist()) == (int((sa == lab).sum()), int((sb[c_] == lab).sum()),
                                                  int(((sa[0] == lab) & (sb[c_] == lab)).sum()))


@pytest.mark.gpu
def test_gpu_fused_step_op_is_the_sampler_step(built):
    from irsgmcmc_b200 import torch_ops
    from irsgmcmc_b200.sampler import SGLDSampler, SGLDConfig
    from irsgmcmc_b200.data_loader.synthetic import make_pair
    n, C = 16, 2
    fixed, moving, vp = make_pair(n)
    outs = []
    for through_op in (False, True):
        torch.manual_seed(0)
        s = SGLDSampler(fixed, moving, C, SGLDConfig(), device=DEV)
        s.set_state(0.5 * torch.randn(C, 3, n, n, n), torch.exp(0.5 * vp['log_var']))
        s.init_gmm(sigma_hat=0.7)
        if through_op:
            h = torch_ops.register_sampler(s)
            torch.ops.irsgmcmc.sgld_step(s.v, s.hyper, s.stats, h, 3)
            with pytest.raises(RuntimeError):
                torch.ops.irsgmcmc.sgld_step(s.v.clone(), s.hyper, s.stats, h, 1)
        else:
            s.step(3, use_graph=False)
        torch.cuda.synchronize()
        outs.append((s.v.clone(), s.hyper.clone(), s.stats.clone()))
    assert all(torch.equal(a, b) for a, b in zip(*outs))
