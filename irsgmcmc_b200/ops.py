"""
Functional wrappers over the C ABI (include/irsgmcmc.h): allocate outputs with torch, pass raw pointers, launch on
torch's current stream.  CUDA tensors only -- every function raises on CPU tensors (no fallback).
"""
import torch

from . import _lib


def _dims(t):
    if t.dim() != 5:
        raise ValueError('expected a (C, channels, D, H, W) tensor')
    return t.shape[0], t.shape[2], t.shape[3], t.shape[4]


def _f32(*tensors):
    for t in tensors:
        if t is not None and t.dtype != torch.float32:
            raise NotImplementedError(f'fp32 only, got {t.dtype}')


def warp3d(img, T, jitter_unit=None, alpha=0.0):
    """trilinear warp, border padding, align_corners (reference utils/registration.py:29-30)"""
    lib = _lib.load()
    _lib.require_cuda(img, T, jitter_unit)
    _f32(img, T, jitter_unit)
    C, D, H, W = _dims(T)
    out = torch.empty(C, 1, D, H, W, device=T.device, dtype=torch.float32)
    stride = 0 if img.shape[0] == 1 else D * H * W
    _lib.check(lib.irs_warp3d_fwd(_lib.ptr(img), stride, _lib.ptr(T), _lib.ptr(jitter_unit), float(alpha), _lib.ptr(out),
                                  C, D, H, W, _lib.stream()))
    return out


def warp3d_bwd_grid(img, T, g_out, jitter_unit=None, alpha=0.0):
    lib = _lib.load()
    _lib.require_cuda(img, T, g_out, jitter_unit)
    _f32(img, T, g_out, jitter_unit)
    C, D, H, W = _dims(T)
    g_T = torch.empty_like(T)
    stride = 0 if img.shape[0] == 1 else D * H * W
    _lib.check(lib.irs_warp3d_bwd_grid(_lib.ptr(img), stride, _lib.ptr(T), _lib.ptr(jitter_unit), float(alpha),
                                       _lib.ptr(g_out), _lib.ptr(g_T), C, D, H, W, _lib.stream()))
    return g_T


def warp3d_nearest(seg, T):
    """nearest-neighbour warp of int16 / bool volumes, bit-exact with the reference (utils/registration.py:20-27)"""
    lib = _lib.load()
    _lib.require_cuda(seg, T)
    _f32(T)
    C, D, H, W = _dims(T)
    stride = 0 if seg.shape[0] == 1 else D * H * W
    out = torch.empty(C, 1, D, H, W, device=T.device, dtype=seg.dtype)
    if seg.dtype == torch.int16:
        _lib.check(lib.irs_warp3d_nearest_i16(_lib.ptr(seg), stride, _lib.ptr(T), _lib.ptr(out), C, D, H, W, _lib.stream()))
    elif seg.dtype in (torch.bool, torch.uint8):
        _lib.check(lib.irs_warp3d_nearest_u8(_lib.ptr(seg), stride, _lib.ptr(T), _lib.ptr(out), C, D, H, W, _lib.stream()))
    else:
        raise NotImplementedError
    return out


def svf_exp_fwd(v, n_steps=12):
    """returns (hist (n_steps,C,3,D,H,W) with hist[-1] = displacement in voxels, maxabs (n_steps,))"""
    lib = _lib.load()
    _lib.require_cuda(v)
    _f32(v)
    C, D, H, W = _dims(v)
    hist = torch.empty(n_steps, C, 3, D, H, W, device=v.device, dtype=torch.float32)
    # per-step max |u_k| followed by the per-cell maxima the adjoint reads (irs_svf_maxabs_floats); the returned tensor is
    # the first n_steps entries, a view that keeps the whole block alive
    mbuf = torch.zeros(int(lib.irs_svf_maxabs_floats(C, D, H, W, n_steps)), device=v.device, dtype=torch.float32)
    maxabs = mbuf[:n_steps]
    _lib.check(lib.irs_svf_exp_fwd(_lib.ptr(v), _lib.ptr(hist), _lib.ptr(maxabs), n_steps, C, D, H, W, _lib.stream()))
    return hist, maxabs


def svf_outputs(u, lin, want_T=True):
    lib = _lib.load()
    C, D, H, W = _dims(u)
    T = torch.empty_like(u) if want_T else None
    _lib.check(lib.irs_svf_outputs(_lib.ptr(u), _lib.ptr(lin[0]), _lib.ptr(lin[1]), _lib.ptr(lin[2]), _lib.ptr(T), None,
                                   C, D, H, W, _lib.stream()))
    return T


def svf_exp_bwd(v, hist, maxabs, g_u, gather_radius_max=2):
    """dL/dv from dL/du_n (g_u is left untouched: a copy is consumed)"""
    lib = _lib.load()
    _lib.require_cuda(v, hist, maxabs, g_u)
    C, D, H, W = _dims(v)
    g_in = g_u.clone()
    work, g_v = torch.empty_like(v), torch.empty_like(v)
    _lib.check(lib.irs_svf_exp_bwd(_lib.ptr(v), _lib.ptr(hist), _lib.ptr(maxabs), _lib.ptr(g_in), _lib.ptr(work),
                                   _lib.ptr(g_v), hist.shape[0], int(gather_radius_max), C, D, H, W, _lib.stream()))
    return g_v


def _ffd_args(cp, kernels, cps, dims):
    lib = _lib.load()
    _lib.require_cuda(cp)
    _f32(cp)
    if cp.dim() != 5 or cp.shape[1] != 3 or len(kernels) != 3 or len(cps) != 3 or len(dims) != 3:
        raise RuntimeError('cubic B-spline FFD: expected a (C,3,gD,gH,gW) field, three kernels, spacings and sizes')
    for k, s in zip(kernels, cps):
        if len(k) != 4 * int(s) - 1:
            raise RuntimeError('cubic B-spline FFD: a kernel of 4 * cps - 1 taps per axis is expected')
    return lib, [_lib.host_floats([float(x) for x in k]) for k in kernels], [int(s) for s in cps]


def ffd_fwd(cp, kernels, cps, dims):
    """dense (C,3,D,H,W) velocity field of control-point velocities cp (C,3,gD,gH,gW); kernels: the three
    B_spline_1D_kernel(s) as host sequences (reference utils/transformation.py:132-152)"""
    lib, hk, st = _ffd_args(cp, kernels, cps, dims)
    C, (gD, gH, gW), (D, H, W) = cp.shape[0], cp.shape[2:], [int(n) for n in dims]
    work = torch.empty(int(lib.irs_ffd_work_floats(C, gD, gH, gW, D, H, W)), device=cp.device, dtype=torch.float32)
    dense = torch.empty(C, 3, D, H, W, device=cp.device, dtype=torch.float32)
    _lib.check(lib.irs_ffd_fwd(_lib.ptr(cp), hk[0], hk[1], hk[2], st[0], st[1], st[2], _lib.ptr(work), _lib.ptr(dense),
                               C, gD, gH, gW, D, H, W, _lib.stream()))
    return dense


def ffd_bwd(g_dense, kernels, cps, grid_size):
    """adjoint of ffd_fwd: (C,3,D,H,W) -> (C,3,gD,gH,gW)"""
    lib, hk, st = _ffd_args(g_dense, kernels, cps, grid_size)
    C, (D, H, W), (gD, gH, gW) = g_dense.shape[0], g_dense.shape[2:], [int(n) for n in grid_size]
    work = torch.empty(int(lib.irs_ffd_work_floats(C, gD, gH, gW, D, H, W)), device=g_dense.device, dtype=torch.float32)
    g_cp = torch.empty(C, 3, gD, gH, gW, device=g_dense.device, dtype=torch.float32)
    _lib.check(lib.irs_ffd_bwd(_lib.ptr(g_dense), hk[0], hk[1], hk[2], st[0], st[1], st[2], _lib.ptr(work),
                               _lib.ptr(g_cp), C, gD, gH, gW, D, H, W, _lib.stream()))
    return g_cp


def bspline_axis(x, kernel, dim, stride, adjoint=False, crop_start=0, out_len=None):
    """one axis of the FFD on an arbitrary contiguous fp32 tensor: conv1D(x, kernel, dim, stride, padding=2 stride - 1,
    transpose=True) (reference utils/transformation.py:106-129) when adjoint is False, its transpose otherwise"""
    lib = _lib.load()
    _lib.require_cuda(x)
    _f32(x)
    stride = int(stride)
    if len(kernel) != 4 * stride - 1:
        raise RuntimeError('bspline_axis: a kernel of 4 * stride - 1 taps is expected')
    dim = dim % x.dim()
    outer = 1
    for n in x.shape[:dim]:
        outer *= int(n)
    inner = 1
    for n in x.shape[dim + 1:]:
        inner *= int(n)
    if adjoint:
        if out_len is None:
            raise ValueError('bspline_axis(adjoint=True) needs out_len = the number of control points along the axis')
        n, g = int(x.shape[dim]), int(out_len)
    else:
        g = int(x.shape[dim])
        n = (g - 1) * stride + 1 - int(crop_start) if out_len is None else int(out_len)
    shape = list(x.shape)
    shape[dim] = g if adjoint else n
    out = torch.empty(shape, device=x.device, dtype=torch.float32)
    _lib.check(lib.irs_bspline_axis(_lib.ptr(x), _lib.ptr(out), int(bool(adjoint)), outer, g, n, inner,
                                    _lib.host_floats([float(v) for v in kernel]), stride, int(crop_start), _lib.stream()))
    return out


def diff_fwd(v, transformation=False):
    lib = _lib.load()
    _lib.require_cuda(v)
    _f32(v)
    C, D, H, W = _dims(v)
    nabla = torch.empty(C, 3, D, H, W, 3, device=v.device, dtype=torch.float32)
    _lib.check(lib.irs_diff_fwd(_lib.ptr(v), _lib.ptr(nabla), int(bool(transformation)), C, D, H, W, _lib.stream()))
    return nabla


def diff_bwd(g_nabla, transformation=False):
    lib = _lib.load()
    _lib.require_cuda(g_nabla)
    C, _, D, H, W, _ = g_nabla.shape
    g_v = torch.empty(C, 3, D, H, W, device=g_nabla.device, dtype=torch.float32)
    _lib.check(lib.irs_diff_bwd(_lib.ptr(g_nabla), _lib.ptr(g_v), int(bool(transformation)), C, D, H, W, _lib.stream()))
    return g_v


def _scratch(C, D, H, W, device):
    n = _lib.load().irs_reduce_scratch_doubles(C, D, H, W)
    return torch.zeros(n, device=device, dtype=torch.float64), torch.zeros(C + 1, device=device, dtype=torch.int32)


def reg_energy(v):
    """sum of squared forward differences per chain, float64 (C,)"""
    lib = _lib.load()
    _lib.require_cuda(v)
    _f32(v)
    C, D, H, W = _dims(v)
    partials, counters = _scratch(C, D, H, W, v.device)
    energy = torch.empty(C, device=v.device, dtype=torch.float64)
    _lib.check(lib.irs_reg_energy(_lib.ptr(v), _lib.ptr(energy), _lib.ptr(partials), _lib.ptr(counters), C, D, H, W,
                                  _lib.stream()))
    return energy


def reg_energy_grad(v, coef):
    """coef[c] * d energy_c / d v"""
    lib = _lib.load()
    _lib.require_cuda(v, coef)
    C, D, H, W = _dims(v)
    g = torch.zeros_like(v)
    _lib.check(lib.irs_reg_energy_grad(_lib.ptr(v), _lib.ptr(coef.double().contiguous()), _lib.ptr(g), C, D, H, W,
                                       _lib.stream()))
    return g


def lcc_normalise(im, s):
    """returns (zn, a, rs): zn = (I - u)/sigma of reference model/loss.py:103-105"""
    lib = _lib.load()
    _lib.require_cuda(im)
    _f32(im)
    C, D, H, W = _dims(im)
    a, rs, zn = torch.empty_like(im), torch.empty_like(im), torch.empty_like(im)
    _lib.check(lib.irs_lcc_normalise(_lib.ptr(im), int(s), _lib.ptr(a), _lib.ptr(rs), _lib.ptr(zn), C, D, H, W,
                                     _lib.stream()))
    return zn, a, rs


def lcc_normalise_bwd(g_zn, a, rs, s):
    lib = _lib.load()
    _lib.require_cuda(g_zn, a, rs)
    C, D, H, W = _dims(a)
    work, g_im = torch.empty_like(a), torch.empty_like(a)
    _lib.check(lib.irs_lcc_normalise_bwd(_lib.ptr(g_zn), _lib.ptr(a), _lib.ptr(rs), int(s), _lib.ptr(work), _lib.ptr(g_im),
                                         C, D, H, W, _lib.stream()))
    return g_im


def gmm_log_pdf(z, log_std, logits, want_dz=False, weights=None, want_param_grads=False):
    """per-element mixture log-density of a flat residual tensor; optional d/dz and weighted parameter gradient sums"""
    lib = _lib.load()
    _lib.require_cuda(z, weights)
    _f32(z, weights)
    K = int(log_std.numel())
    n = z.numel()
    host = _lib.host_floats(list(log_std.detach().cpu().float().tolist()) + list(logits.detach().cpu().float().tolist()))
    logp = torch.empty(n, device=z.device, dtype=torch.float32)
    dz = torch.empty(n, device=z.device, dtype=torch.float32) if want_dz else None
    gp = partials = counter = None
    if want_param_grads:
        gp = torch.zeros(2 * _lib.MAX_K, device=z.device, dtype=torch.float64)
        partials = torch.zeros(1184 * 2 * _lib.MAX_K, device=z.device, dtype=torch.float64)
        counter = torch.zeros(1, device=z.device, dtype=torch.int32)
    _lib.check(lib.irs_gmm_log_pdf(_lib.ptr(z), n, host, K, _lib.ptr(logp), _lib.ptr(dz), _lib.ptr(weights), _lib.ptr(gp),
                                   _lib.ptr(partials), _lib.ptr(counter), _lib.stream()))
    return logp, dz, gp


def vd_factor(z, mask, log_std, logits):
    """virtual decimation factor of one chain (reference utils/util.py:330-347,446-485); float64 scalar tensor"""
    lib = _lib.load()
    _lib.require_cuda(z, mask)
    _f32(z)
    D, H, W = z.shape[-3:]
    K = int(log_std.numel())
    host = _lib.host_floats(list(log_std.detach().cpu().float().tolist()) + list(logits.detach().cpu().float().tolist()))
    alpha = torch.empty(1, device=z.device, dtype=torch.float64)
    partials = torch.zeros(1184 * 32, device=z.device, dtype=torch.float64)
    counter = torch.zeros(1, device=z.device, dtype=torch.int32)
    m8 = mask.view(torch.uint8) if mask.dtype == torch.bool else mask
    _lib.check(lib.irs_vd_factor(_lib.ptr(z), _lib.ptr(m8), host, K, _lib.ptr(alpha), _lib.ptr(partials),
                                 _lib.ptr(counter), D, H, W, _lib.stream()))
    return alpha[0]


def vd_factor_from_residual(r, mask):
    """calc_VD_factor of the reference: alpha from a rescaled residual field (1,1,D,H,W)"""
    lib = _lib.load()
    _lib.require_cuda(r, mask)
    _f32(r)
    D, H, W = r.shape[-3:]
    alpha = torch.empty(1, device=r.device, dtype=torch.float64)
    partials = torch.zeros(1184 * 5, device=r.device, dtype=torch.float64)
    counter = torch.zeros(1, device=r.device, dtype=torch.int32)
    m8 = mask.view(torch.uint8) if mask.dtype == torch.bool else mask
    _lib.check(lib.irs_vd_factor_residual(_lib.ptr(r), _lib.ptr(m8), _lib.ptr(alpha), _lib.ptr(partials), _lib.ptr(counter),
                                          D, H, W, _lib.stream()))
    return alpha[0]


def masked_mean_std(z, mask):
    lib = _lib.load()
    _lib.require_cuda(z, mask)
    out = torch.empty(3, device=z.device, dtype=torch.float64)
    partials = torch.zeros(1184 * 3, device=z.device, dtype=torch.float64)
    counter = torch.zeros(1, device=z.device, dtype=torch.int32)
    m8 = mask.view(torch.uint8) if mask.dtype == torch.bool else mask
    _lib.check(lib.irs_masked_mean_std(_lib.ptr(z), _lib.ptr(m8), z.numel(), _lib.ptr(out), _lib.ptr(partials),
                                       _lib.ptr(counter), _lib.stream()))
    return out


def welford_update(sample, count_before, mean, m2):
    lib = _lib.load()
    _lib.require_cuda(sample, mean, m2)
    n_new, n = sample.shape[0], mean.numel()
    _lib.check(lib.irs_welford_update(_lib.ptr(sample), n_new, n, float(count_before), _lib.ptr(mean), _lib.ptr(m2),
                                      _lib.stream()))
    return count_before + n_new


def welford_std(m2, count):
    lib = _lib.load()
    out = torch.empty_like(m2)
    _lib.check(lib.irs_welford_std(_lib.ptr(m2), float(count), _lib.ptr(out), m2.numel(), _lib.stream()))
    return out


def log_det_jacobian(T, want_log_det=True):
    """log det J of a transformation (C,3,D,H,W) and the per-sample count of folded (NaN) voxels"""
    lib = _lib.load()
    _lib.require_cuda(T)
    _f32(T)
    C, D, H, W = _dims(T)
    log_det = torch.empty(C, D, H, W, device=T.device, dtype=torch.float32) if want_log_det else None
    counts = torch.empty(C, device=T.device, dtype=torch.int32)
    _lib.check(lib.irs_log_det_jacobian(_lib.ptr(T), _lib.ptr(log_det), _lib.ptr(counts), C, D, H, W, _lib.stream()))
    return counts, log_det


def dice_counts(seg_a, seg_b, labels):
    """(C, n_labels, 3) counts |A = l|, |B = l|, |A = l and B = l| for int16 label volumes (seg_a may be shared by chains)"""
    import ctypes
    lib = _lib.load()
    _lib.require_cuda(seg_a, seg_b)
    if seg_a.dtype != torch.int16 or seg_b.dtype != torch.int16:
        raise NotImplementedError('int16 segmentations only')
    C = seg_b.shape[0]
    V = seg_b[0].numel()
    stride = 0 if seg_a.shape[0] == 1 else V
    lab = (ctypes.c_int * len(labels))(*[int(x) for x in labels])
    counts = torch.empty(C, len(labels), 3, device=seg_b.device, dtype=torch.int32)
    _lib.check(lib.irs_dice_counts(_lib.ptr(seg_a), stride, _lib.ptr(seg_b), lab, len(labels), _lib.ptr(counts), C, V,
                                   _lib.stream()))
    return counts
