"""TEST INFRASTRUCTURE -- run as a script in its own process by tests/test_torch_ops.py.

Checks the wiring of torch.ops.irsgmcmc.* (schemas, fake implementations, autograd formulas) without a GPU: for this
process only, every op gets a CPU kernel made of the ORACLE (oracle/sgld_oracle.py), then forward and backward through the
dispatcher are compared with the oracle's own autograd.  The product registers CUDA kernels only; nothing here ships.
"""
import concurrent.futures
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import irsgmcmc_b200  # noqa: E402,F401  (registers torch.ops.irsgmcmc)
from irsgmcmc_b200 import torch_ops as TO  # noqa: E402
from oracle import sgld_oracle as O  # noqa: E402

_pool = concurrent.futures.ThreadPoolExecutor(1)


def with_autograd(fn):
    """kernels run below the autograd dispatch key; a fresh thread has fresh dispatch state, so the oracle's autograd works"""
    return _pool.submit(fn).result()


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300))


def jitter(T, ju, alpha):
    return T if ju is None else T + O.uniform_jitter_normalised(ju, alpha, T.shape)


# ---- oracle-backed CPU kernels (this process only) ----------------------------------------------------------------------
@TO.warp3d.register_kernel('cpu')
def _(img, T, jitter_unit, alpha):
    return O.warp_aten(img.expand(T.shape[0], -1, -1, -1, -1), jitter(T, jitter_unit, alpha))


@TO.warp3d_bwd_grid.register_kernel('cpu')
def _(img, T, g_out, jitter_unit, alpha):
    def run():
        Tt = T.detach().clone().requires_grad_(True)
        out = O.warp_aten(img.expand(T.shape[0], -1, -1, -1, -1), jitter(Tt, jitter_unit, alpha))
        return torch.autograd.grad(out, Tt, g_out)[0]
    return with_autograd(run)


@TO.warp3d_nearest.register_kernel('cpu')
def _(seg, T):
    return O.warp_nearest(seg.expand(T.shape[0], -1, -1, -1, -1), T)


@TO.svf_exp.register_kernel('cpu')
def _(v, n_steps):
    _, disp = O.svf_exp_aten(v, n_steps)
    return disp.clone(), disp.unsqueeze(0).repeat(n_steps, 1, 1, 1, 1, 1), v.new_zeros(n_steps)


@TO.svf_exp_bwd.register_kernel('cpu')
def _(v, hist, maxabs, g_u, gather_radius_max):
    def run():
        vv = v.detach().clone().requires_grad_(True)
        _, disp = O.svf_exp_aten(vv, hist.shape[0])
        return torch.autograd.grad(disp, vv, g_u)[0]
    return with_autograd(run)


@TO.sobolev_smooth.register_kernel('cpu')
def _(v, taps):
    return O.sobolev_smooth(v, torch.tensor(list(taps), dtype=torch.float32).numpy())


@TO.lcc_normalise.register_kernel('cpu')
def _(im, s):
    k = 2 * s + 1
    u = O.box_sum(im, s) / k ** 3
    a = im - u
    rs = 1.0 / torch.sqrt(O.box_sum(a * a, s) / k ** 3 + 1e-10)
    return a * rs, a, rs


@TO.lcc_normalise_bwd.register_kernel('cpu')
def _(g_zn, a, rs, s):
    raise NotImplementedError('filled in by main(): needs the image')


@TO.reg_energy.register_kernel('cpu')
def _(v):
    return O.reg_energy(v.double())


@TO.reg_energy_grad.register_kernel('cpu')
def _(v, coef):
    def run():
        vv = v.detach().double().requires_grad_(True)
        return torch.autograd.grad(O.reg_energy(vv), vv, coef.double())[0].float()
    return with_autograd(run)


@TO.ffd.register_kernel('cpu')
def _(cp, kernel_d, kernel_h, kernel_w, cps, dims):
    return O.ffd_dense(cp, tuple(dims), tuple(cps))


@TO.ffd_adjoint.register_kernel('cpu')
def _(g_dense, kernel_d, kernel_h, kernel_w, cps, grid):
    def run():
        cp = torch.zeros(g_dense.shape[0], 3, *grid, requires_grad=True)
        return torch.autograd.grad(O.ffd_dense(cp, tuple(g_dense.shape[2:]), tuple(cps)), cp, g_dense)[0]
    return with_autograd(run)


def main():
    torch.manual_seed(0)
    n, C = 10, 2
    ops = torch.ops.irsgmcmc

    # warp: forward, gradient w.r.t. the grid only, jitter
    img = torch.rand(1, 1, n, n, n)
    v = 1.5 * torch.randn(C, 3, n, n, n)
    T0, _ = O.svf_exp_aten(F.avg_pool3d(F.pad(v, (1,) * 6, mode='replicate'), 3, 1))
    ju = torch.rand(C, 3, n, n, n)
    for jit in (None, ju):
        T = T0.clone().requires_grad_(True)
        img_g = img.clone().requires_grad_(True)
        out = ops.warp3d(img_g, T, jit, 0.1)
        G = torch.randn_like(out)
        (out * G).sum().backward()
        Tr = T0.clone().requires_grad_(True)
        ref = O.warp_aten(img.expand(C, -1, -1, -1, -1), jitter(Tr, jit, 0.1))
        (ref * G).sum().backward()
        assert rel(out, ref) < 1e-6 and rel(T.grad, Tr.grad) < 1e-6 and img_g.grad is None
    seg = (torch.rand(1, 1, n, n, n) * 40).short()
    assert torch.equal(ops.warp3d_nearest(seg, T0), O.warp_nearest(seg.expand(C, -1, -1, -1, -1), T0))

    # SVF: three outputs, gradient through the displacement only
    vs = F.avg_pool3d(F.pad(v, (1,) * 6, mode='replicate'), 3, 1).requires_grad_(True)
    disp, hist, maxabs = ops.svf_exp(vs, 12)
    G = torch.randn_like(disp)
    (disp * G).sum().backward()
    vr = vs.detach().clone().requires_grad_(True)
    _, dr = O.svf_exp_aten(vr, 12)
    (dr * G).sum().backward()
    assert hist.shape == (12, C, 3, n, n, n) and rel(disp, dr) < 1e-6 and rel(vs.grad, vr.grad) < 1e-6

    # Sobolev smoothing: forward = separable smoothing, backward = identity (the reference's quirk)
    taps = [float(t) for t in O.sobolev_taps(3, 0.5)]
    x = torch.randn(C, 3, n, n, n, requires_grad=True)
    y = ops.sobolev_smooth(x, taps)
    G = torch.randn_like(y)
    (y * G).sum().backward()
    assert rel(y, O.sobolev_smooth(x.detach(), O.sobolev_taps(3, 0.5).astype('float32'))) < 1e-6 and torch.equal(x.grad, G)

    # LCC normalisation: backward through zn only
    im = torch.rand(C, 1, n, n, n)

    def lcc_bwd_cpu(g_zn, a, rs, s):
        def run():
            ii = im.detach().clone().requires_grad_(True)
            return torch.autograd.grad(O.lcc_normalise(ii, s), ii, g_zn)[0]
        return with_autograd(run)

    TO.lcc_normalise_bwd.register_kernel('cpu')(lcc_bwd_cpu)
    imr = im.clone().requires_grad_(True)
    zn, a, rs = ops.lcc_normalise(imr, 2)
    G = torch.randn_like(zn)
    (zn * G).sum().backward()
    ir = im.clone().requires_grad_(True)
    (O.lcc_normalise(ir, 2) * G).sum().backward()
    assert rel(zn, O.lcc_normalise(im, 2)) < 1e-6 and rel(imr.grad, ir.grad) < 1e-6

    # regulariser energy: float64 (C,), float32 gradient
    x = torch.randn(C, 3, n, n, n, requires_grad=True)
    e = ops.reg_energy(x)
    w = torch.tensor([0.7, -1.3], dtype=torch.float64)
    (e * w).sum().backward()
    xr = x.detach().double().requires_grad_(True)
    (O.reg_energy(xr) * w).sum().backward()
    assert e.dtype == torch.float64 and e.shape == (C,) and x.grad.dtype == torch.float32 and rel(x.grad, xr.grad) < 1e-6

    # FFD
    cps, dims = (3, 2, 4), (n, n, n)
    grid = O.control_grid_size(dims, cps)
    ks = [[float(t) for t in O.bspline_taps(s)] for s in cps]
    cp = torch.randn(C, 3, *grid, requires_grad=True)
    dense = ops.ffd(cp, ks[0], ks[1], ks[2], list(cps), list(dims))
    G = torch.randn_like(dense)
    (dense * G).sum().backward()
    cr = cp.detach().clone().requires_grad_(True)
    (O.ffd_dense(cr, dims, cps) * G).sum().backward()
    assert dense.shape == (C, 3, *dims) and rel(dense, O.ffd_dense(cp.detach(), dims, cps)) < 1e-6
    assert cp.grad.shape == cp.shape and rel(cp.grad, cr.grad) < 1e-6

    # torch.library's own consistency checks (schema vs. implementation, fake vs. real, autograd registration)
    checks = ('test_schema', 'test_faketensor', 'test_autograd_registration')
    torch.library.opcheck(TO.warp3d, (img, T0.clone().requires_grad_(True), ju, 0.1), test_utils=checks)
    torch.library.opcheck(TO.reg_energy, (torch.randn(C, 3, n, n, n, requires_grad=True),), test_utils=checks)
    torch.library.opcheck(TO.ffd, (cp.detach().clone().requires_grad_(True), ks[0], ks[1], ks[2], list(cps), list(dims)),
                          test_utils=checks)
    torch.library.opcheck(TO.sobolev_smooth, (torch.randn(C, 3, n, n, n, requires_grad=True), taps), test_utils=checks)
    torch.library.opcheck(TO.lcc_normalise, (im.clone().requires_grad_(True), 2), test_utils=checks)
    print('torch_ops wiring OK')


if __name__ == '__main__':
    main()
