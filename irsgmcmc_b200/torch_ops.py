"""
The operators of the SGLD registration step as PyTorch custom ops: ``torch.ops.irsgmcmc.*``.

Each op is a thin ``torch.library`` registration over the C ABI of libirsgmcmc.so (through ``ops.py``): a CUDA kernel
only -- the dispatcher itself refuses CPU tensors ("Could not run 'irsgmcmc::...' with arguments from the 'CPU' backend") --
a fake (meta) implementation for shape propagation under FakeTensorMode / torch.compile tracing, and the autograd formula,
which calls the adjoint op (the closed forms of SURVEY Appendix A, no scatter atomics).

    reference call site                                                   op
    RegistrationModule.forward, float branch (utils/registration.py:29)   irsgmcmc::warp3d, ::warp3d_bwd_grid
    RegistrationModule.forward, seg / mask  (utils/registration.py:20)    irsgmcmc::warp3d_nearest
    SVF_3D.forward + autograd (utils/transformation.py:63-76)             irsgmcmc::svf_exp, ::svf_exp_bwd
    SobolevGrad.apply (utils/functions.py:98-109)                         irsgmcmc::sobolev_smooth (backward = identity)
    GMM.map, one image side (model/loss.py:102-111)                       irsgmcmc::lcc_normalise, ::lcc_normalise_bwd
    RegLoss.forward energy (model/loss.py:152-161)                        irsgmcmc::reg_energy, ::reg_energy_grad
    Cubic_B_spline_FFD_3D.forward (utils/transformation.py:132-152)       irsgmcmc::ffd, ::ffd_adjoint

The drop-in modules (utils/, model/) and the fused sampler use the same launchers; this module adds the dispatcher-visible
surface the north star asks for and nothing else.
"""
from typing import Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib, ops

_NS = 'irsgmcmc'


def _op(name, **kw):
    return torch.library.custom_op(f'{_NS}::{name}', mutates_args=(), device_types='cuda', **kw)


# ---- warps ------------------------------------------------------------------------------------------------------------
@_op('warp3d')
def warp3d(img: Tensor, T: Tensor, jitter_unit: Optional[Tensor], alpha: float) -> Tensor:
    return ops.warp3d(img.contiguous(), T.contiguous(), None if jitter_unit is None else jitter_unit.contiguous(), alpha)


@warp3d.register_fake
def _(img, T, jitter_unit, alpha):
    return T.new_empty(T.shape[0], 1, T.shape[2], T.shape[3], T.shape[4])


@_op('warp3d_bwd_grid')
def warp3d_bwd_grid(img: Tensor, T: Tensor, g_out: Tensor, jitter_unit: Optional[Tensor], alpha: float) -> Tensor:
    return ops.warp3d_bwd_grid(img.contiguous(), T.contiguous(), g_out.contiguous(),
                               None if jitter_unit is None else jitter_unit.contiguous(), alpha)


@warp3d_bwd_grid.register_fake
def _(img, T, g_out, jitter_unit, alpha):
    return torch.empty_like(T)


def _warp3d_setup(ctx, inputs, output):
    img, T, jitter_unit, alpha = inputs
    ctx.has_jitter, ctx.alpha = jitter_unit is not None, alpha
    ctx.save_for_backward(*((img, T, jitter_unit) if ctx.has_jitter else (img, T)))


def _warp3d_backward(ctx, g):
    saved = ctx.saved_tensors
    jitter_unit = saved[2] if ctx.has_jitter else None
    # like the reference, only the grid receives a gradient (the moving image is data: utils/registration.py:29-30)
    return None, torch.ops.irsgmcmc.warp3d_bwd_grid(saved[0], saved[1], g, jitter_unit, ctx.alpha), None, None


warp3d.register_autograd(_warp3d_backward, setup_context=_warp3d_setup)


@_op('warp3d_nearest')
def warp3d_nearest(seg: Tensor, T: Tensor) -> Tensor:
    return ops.warp3d_nearest(seg.contiguous(), T.contiguous())


@warp3d_nearest.register_fake
def _(seg, T):
    return seg.new_empty(T.shape[0], 1, T.shape[2], T.shape[3], T.shape[4])


# ---- stationary velocity field ------------------------------------------------------------------------------------------
@_op('svf_exp')
def svf_exp(v: Tensor, n_steps: int) -> Tuple[Tensor, Tensor, Tensor]:
    """(displacement in voxels, history u_1..u_n, max |u_k| workspace) of scaling and squaring"""
    hist, maxabs = ops.svf_exp_fwd(v.contiguous(), n_steps)
    workspace = maxabs._base if maxabs._base is not None else maxabs   # per-step maxima followed by the per-cell maps
    return hist[-1].clone(), hist, workspace


@svf_exp.register_fake
def _(v, n_steps):
    C, D, H, W = v.shape[0], v.shape[2], v.shape[3], v.shape[4]
    n_max = int(_lib.load().irs_svf_maxabs_floats(int(C), int(D), int(H), int(W), int(n_steps)))
    return torch.empty_like(v), v.new_empty(n_steps, *v.shape), v.new_empty(n_max)


@_op('svf_exp_bwd')
def svf_exp_bwd(v: Tensor, hist: Tensor, maxabs: Tensor, g_u: Tensor, gather_radius_max: int) -> Tensor:
    return ops.svf_exp_bwd(v.contiguous(), hist, maxabs, g_u.contiguous(), gather_radius_max)


@svf_exp_bwd.register_fake
def _(v, hist, maxabs, g_u, gather_radius_max):
    return torch.empty_like(v)


def _svf_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0], output[1], output[2])
    ctx.set_materialize_grads(False)


def _svf_backward(ctx, g_disp, g_hist, g_maxabs):
    v, hist, maxabs = ctx.saved_tensors
    if g_disp is None:
        return None, None
    return torch.ops.irsgmcmc.svf_exp_bwd(v, hist, maxabs, g_disp, 3), None


svf_exp.register_autograd(_svf_backward, setup_context=_svf_setup)


# ---- Sobolev smoothing --------------------------------------------------------------------------------------------------
@_op('sobolev_smooth')
def sobolev_smooth(v: Tensor, taps: Sequence[float]) -> Tensor:
    from .utils.functions import langevin_sobolev
    return langevin_sobolev(v.contiguous(), None, 0.0, [float(t) for t in taps])


@sobolev_smooth.register_fake
def _(v, taps):
    return torch.empty_like(v)


# SobolevGrad.backward passes the gradient through unchanged (reference utils/functions.py:107-109)
sobolev_smooth.register_autograd(lambda ctx, g: (g, None), setup_context=lambda ctx, inputs, output: None)


# ---- LCC normalisation ----------------------------------------------------------------------------------------------------
@_op('lcc_normalise')
def lcc_normalise(im: Tensor, s: int) -> Tuple[Tensor, Tensor, Tensor]:
    """(zn, a, rs): a = I - Box(I)/k^3, rs = 1/sqrt(Box(a^2)/k^3 + 1e-10), zn = a rs, k = 2 s + 1"""
    return ops.lcc_normalise(im.contiguous(), s)


@lcc_normalise.register_fake
def _(im, s):
    return torch.empty_like(im), torch.empty_like(im), torch.empty_like(im)


@_op('lcc_normalise_bwd')
def lcc_normalise_bwd(g_zn: Tensor, a: Tensor, rs: Tensor, s: int) -> Tensor:
    return ops.lcc_normalise_bwd(g_zn.contiguous(), a, rs, s)


@lcc_normalise_bwd.register_fake
def _(g_zn, a, rs, s):
    return torch.empty_like(a)


def _lcc_setup(ctx, inputs, output):
    ctx.s = inputs[1]
    ctx.save_for_backward(output[1], output[2])
    ctx.set_materialize_grads(False)


def _lcc_backward(ctx, g_zn, g_a, g_rs):
    if g_a is not None or g_rs is not None:
        raise NotImplementedError('irsgmcmc::lcc_normalise: only zn is differentiable')
    if g_zn is None:
        return None, None
    a, rs = ctx.saved_tensors
    return torch.ops.irsgmcmc.lcc_normalise_bwd(g_zn, a, rs, ctx.s), None


lcc_normalise.register_autograd(_lcc_backward, setup_context=_lcc_setup)


# ---- regulariser energy -----------------------------------------------------------------------------------------------------
@_op('reg_energy')
def reg_energy(v: Tensor) -> Tensor:
    """sum of squared forward differences per batch entry, float64 (C,)"""
    return ops.reg_energy(v.contiguous())


@reg_energy.register_fake
def _(v):
    return v.new_empty(v.shape[0], dtype=torch.float64)


@_op('reg_energy_grad')
def reg_energy_grad(v: Tensor, coef: Tensor) -> Tensor:
    """coef[c] * d energy_c / d v"""
    return ops.reg_energy_grad(v.contiguous(), coef.contiguous())


@reg_energy_grad.register_fake
def _(v, coef):
    return torch.empty_like(v)


def _energy_backward(ctx, g_y):
    v, = ctx.saved_tensors
    return torch.ops.irsgmcmc.reg_energy_grad(v, g_y.double())


reg_energy.register_autograd(_energy_backward, setup_context=lambda ctx, inputs, output: ctx.save_for_backward(inputs[0]))


# ---- cubic B-spline FFD -----------------------------------------------------------------------------------------------------
@_op('ffd')
def ffd(cp: Tensor, kernel_d: Sequence[float], kernel_h: Sequence[float], kernel_w: Sequence[float], cps: Sequence[int],
        dims: Sequence[int]) -> Tensor:
    return ops.ffd_fwd(cp.contiguous(), [list(kernel_d), list(kernel_h), list(kernel_w)], list(cps), list(dims))


@ffd.register_fake
def _(cp, kernel_d, kernel_h, kernel_w, cps, dims):
    return cp.new_empty(cp.shape[0], 3, dims[0], dims[1], dims[2])


@_op('ffd_adjoint')
def ffd_adjoint(g_dense: Tensor, kernel_d: Sequence[float], kernel_h: Sequence[float], kernel_w: Sequence[float],
                cps: Sequence[int], grid: Sequence[int]) -> Tensor:
    return ops.ffd_bwd(g_dense.contiguous(), [list(kernel_d), list(kernel_h), list(kernel_w)], list(cps), list(grid))


@ffd_adjoint.register_fake
def _(g_dense, kernel_d, kernel_h, kernel_w, cps, grid):
    return g_dense.new_empty(g_dense.shape[0], 3, grid[0], grid[1], grid[2])


def _ffd_setup(ctx, inputs, output):
    cp, ctx.kd, ctx.kh, ctx.kw, ctx.cps, _ = inputs
    ctx.grid = [int(n) for n in cp.shape[2:]]


def _ffd_backward(ctx, g):
    return torch.ops.irsgmcmc.ffd_adjoint(g, ctx.kd, ctx.kh, ctx.kw, ctx.cps, ctx.grid), None, None, None, None, None


ffd.register_autograd(_ffd_backward, setup_context=_ffd_setup)

OPS = ('warp3d', 'warp3d_bwd_grid', 'warp3d_nearest', 'svf_exp', 'svf_exp_bwd', 'sobolev_smooth', 'lcc_normalise',
       'lcc_normalise_bwd', 'reg_energy', 'reg_energy_grad', 'ffd', 'ffd_adjoint')
