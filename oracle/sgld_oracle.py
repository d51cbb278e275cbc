"""TEST INFRASTRUCTURE — the parity oracle.  Not product code; nothing under ``irsgmcmc_b200/`` imports it.

CPU restatement (torch CPU tensors, fp32 or fp64) of the reference's per-iteration SGLD registration step,
``Trainer._SGLD_transition`` (reference ``trainer/trainer.py:291-356``) and of every operator on that path.  Written
from the maths, stage by stage, each function citing the reference lines it follows.

Where the arithmetic lives: the reference has no arithmetic of its own for warping -- it calls ``F.grid_sample`` of an
un-pinned PyTorch (``utils/registration.py:22,30``, ``utils/transformation.py:72-73``).  The oracle anchors on the
installed torch 2.11.0 for exactly those calls (``*_aten`` functions) and *additionally* restates the published
grid-sampler algorithm by hand (``trilinear_sample_voxel``/``nearest_index``; ATen ``GridSampler.cuh:21-31,53-81``) so
that the two can be checked against each other.

Pinning (tests/test_oracle_vs_reference.py, tests/test_golden.py): every function here is compared
 (a) with the unmodified reference imported from /root/reference (build container only), and
 (b) with golden vectors generated from the unmodified reference (tests/golden/*.npz, script tests/golden/make_golden.py),
 (c) with the reference's own known-answer tests (tests/test_diff.py, tests/test_utils.py of the reference), restated.

The same functions, run in fp32 with all host threads, are the "port" CPU baseline of bench.py.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

LOG_SQRT_2PI = 0.5 * math.log(2.0 * math.pi)


# ----------------------------------------------------------------------------------------------------------------------
# Sobolev kernel, Langevin proposal, separable smoothing
# ----------------------------------------------------------------------------------------------------------------------

def sobolev_taps(s, lam):
    """
    reference utils/functions.py:24-49 (Sobolev_kernel_1D): the smoothing kernel is the middle column of
    (I - lam * L)^-1, L the (2s+1)x(2s+1) 1-D Laplacian with zero boundary, normalised to unit sum.
    s=3, lam=.5 -> [1,4,15,56,15,4,1]/96
    """
    k = 2 * s + 1
    lap = -2.0 * np.eye(k) + np.eye(k, k=1) + np.eye(k, k=-1)
    e = np.zeros(k)
    e[s] = 1.0
    col = np.linalg.solve(np.eye(k) - lam * lap, e)
    return col / col.sum()


def langevin_proposal(v, sigma, tau, eps):
    """reference utils/util.py:48-58: v + sqrt(2 tau) * sigma * eps, eps ~ N(0,1)"""
    return v + math.sqrt(2.0 * tau) * sigma * eps


def _shift_replicate(x, axis, offset):
    """x sampled at index+offset along axis with the index clamped to the valid range (replicate padding)"""
    n = x.shape[axis]
    idx = torch.clamp(torch.arange(n) + offset, 0, n - 1)
    return x.index_select(axis, idx)


def sobolev_smooth(v, taps):
    """
    reference utils/functions.py:98-105 + utils/util.py:394-404: replicate-pad by s on all six sides, then depthwise
    correlation along z (dim 2), y (dim 3), x (dim 4) in that order with the same symmetric taps.
    Clamped indexing == replicate padding followed by a 'valid' correlation.
    """
    s = (len(taps) - 1) // 2
    out = v
    for axis in (2, 3, 4):
        acc = torch.zeros_like(out)
        for t, w in enumerate(taps):
            acc = acc + float(w) * _shift_replicate(out, axis, t - s)
        out = acc
    return out


# ----------------------------------------------------------------------------------------------------------------------
# coordinates, SVF, warps
# ----------------------------------------------------------------------------------------------------------------------

def identity_grid(dims, dtype=torch.float32, exact=False):
    """
    reference utils/util.py:263-278: (1,nz,ny,nx,3) with last-dim order (x,y,z); built from fp32 torch.linspace.
    exact=False with fp64 reproduces `SVF_3D(...).double()`: the fp32-rounded grid cast to double (SURVEY surprise 9).
    """
    nx, ny, nz = dims
    lin = (lambda n: torch.linspace(-1, 1, steps=n, dtype=dtype)) if exact else \
          (lambda n: torch.linspace(-1, 1, steps=n).to(dtype))
    x, y, z = lin(nx), lin(ny), lin(nz)
    gz, gy, gx = torch.meshgrid(z, y, x, indexing='ij')
    return torch.stack((gx, gy, gz), -1).unsqueeze(0)


def to_normalised(field):
    """reference utils/util.py:418-429: channel i times 2/(shape[2+i]-1) (sic: channel 0 uses dim D)"""
    scale = torch.tensor([2.0 / (n - 1) for n in field.shape[2:]], dtype=field.dtype).view(1, -1, 1, 1, 1)
    return field * scale


def to_voxels(field):
    """reference utils/util.py:432-443"""
    scale = torch.tensor([(n - 1) / 2.0 for n in field.shape[2:]], dtype=field.dtype).view(1, -1, 1, 1, 1)
    return field * scale


def svf_exp_aten(v, no_steps=12, exact_grid=False):
    """
    reference utils/transformation.py:63-76: d0 = normalised(v) / 2^steps; steps x  d <- d + grid_sample(d, id + d);
    returns (T = id + d, displacement in voxels)
    """
    D, H, W = v.shape[2:]
    grid = identity_grid((W, H, D), v.dtype, exact_grid)
    d = to_normalised(v) / float(2 ** no_steps)
    for _ in range(no_steps):
        d = d + F.grid_sample(d, grid + d.permute(0, 2, 3, 4, 1), padding_mode='border', align_corners=True)
    return grid.permute(0, 4, 1, 2, 3) + d, to_voxels(d)


# ----------------------------------------------------------------------------------------------------------------------
# cubic B-spline FFD
# ----------------------------------------------------------------------------------------------------------------------

def control_grid_size(dims, cps):
    """reference utils/util.py:61-69 (get_control_grid_size): ceil((n - 1) / cps) + 1 control points + 2 outside"""
    return tuple(int(math.ceil((n - 1) / c) + 1 + 2) for n, c in zip(dims, cps))


def bspline_taps(stride, dtype=torch.float32):
    """
    reference utils/transformation.py:79-103 (cubic_B_spline_1D_value, B_spline_1D_kernel): the cubic B-spline
    B(t) = 2/3 + (t/2 - 1) t^2 for t < 1, -(t - 2)^3 / 6 for 1 <= t < 2, sampled at (i - (2 s - 1)) / s, i = 0 .. 4 s - 2;
    evaluated in Python doubles and stored in an fp32 tensor like the reference does
    """
    taps = torch.ones(4 * stride - 1)
    r = taps.shape[0] // 2
    for i in range(taps.shape[0]):
        t = abs((i - r) / stride)
        taps[i] = 0.0 if t >= 2 else (2.0 / 3.0 + (0.5 * t - 1.0) * t ** 2 if t < 1 else -1.0 * ((t - 2.0) ** 3) / 6.0)
    return taps.to(dtype)


def bspline_axis(x, axis, s, n=None, crop_start=0):
    """
    one axis of the FFD = reference conv1D(x, B_spline_1D_kernel(s), dim=axis, stride=s, padding=2 s - 1, transpose=True)
    (utils/transformation.py:106-129), restated without a convolution: element p = j + crop_start of the transposed
    convolution is sum_i x[i] B(p / s - i), i.e. with p + 2 s - 1 = q s + r the four inputs i = q - m (m = 0..3) weighted
    by taps[r + m s]; n elements from crop_start on (default: all (g - 1) s + 1 of them)
    """
    taps = bspline_taps(s, x.dtype)
    g = x.shape[axis]
    n = (g - 1) * s + 1 - crop_start if n is None else n
    t = torch.arange(n) + crop_start + 2 * s - 1
    q, r = t // s, t % s
    shape = [1] * x.dim()
    shape[axis] = n
    acc = 0
    for m in range(3, -1, -1):
        i, j = q - m, r + m * s
        ok = (i >= 0) & (i < g) & (j <= 4 * s - 2)
        w = torch.where(ok, taps[j.clamp(max=4 * s - 2)], torch.zeros((), dtype=x.dtype))
        acc = acc + x.index_select(axis, i.clamp(0, g - 1)) * w.view(shape)
    return acc


def ffd_dense(cp, dims, cps):
    """
    reference utils/transformation.py:132-152 (Cubic_B_spline_FFD_3D.forward): the transposed convolution per axis (D, H,
    W in that order), then the crop [s, s + n): dense[x] = sum_i cp[i] B((x + s) / s - i) along each axis
    """
    out = cp
    for a, (n, s) in enumerate(zip(dims, cps)):
        out = bspline_axis(out, a + 2, s, n, crop_start=s)
    return out


def svffd_exp_aten(cp, dims, cps, no_steps=12):
    """reference utils/transformation.py:155-164 (SVFFD_3D): SVF_3D of the dense B-spline velocity field"""
    return svf_exp_aten(ffd_dense(cp, dims, cps), no_steps)


def warp_aten(im, T):
    """reference utils/registration.py:29-30 (float images)"""
    return F.grid_sample(im, T.permute(0, 2, 3, 4, 1), mode='bilinear', padding_mode='border', align_corners=True)


def warp_nearest_aten(seg, T):
    """reference utils/registration.py:20-27 (bool masks / int16 segmentations through a float round trip)"""
    out = F.grid_sample(seg.float(), T.permute(0, 2, 3, 4, 1).float(), mode='nearest', padding_mode='border',
                        align_corners=True)
    return out.to(seg.dtype)


def unnormalise(T):
    """ATen GridSampler.cuh:21-31 (align_corners) + :53-57 (border clip): voxel coordinates (x,y,z) of a grid T (N,3,D,H,W)"""
    D, H, W = T.shape[2:]
    out = []
    for ch, n in enumerate((W, H, D)):
        c = ((T[:, ch] + 1.0) / 2) * (n - 1)
        out.append(torch.clamp(c, 0, n - 1))
    return out


def trilinear_sample_voxel(vol, px, py, pz):
    """
    hand-written restatement of ATen's trilinear border sampler (GridSampler.cuh:149-218 of torch 2.11): `vol` (N,C,D,H,W)
    sampled at clamped voxel coordinates px,py,pz (N,D,H,W); corners outside the volume contribute zero.
    Differentiable w.r.t. vol and (almost everywhere) w.r.t. the coordinates; like ATen, the derivative w.r.t. a
    coordinate that sits on/outside the border is zero -- callers apply that mask (see svf_exp_voxel).
    """
    N, C, D, H, W = vol.shape
    x0, y0, z0 = torch.floor(px), torch.floor(py), torch.floor(pz)
    fx, fy, fz = px - x0, py - y0, pz - z0
    x0, y0, z0 = x0.long(), y0.long(), z0.long()
    flat = vol.reshape(N, C, -1)
    out = 0
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                xi, yi, zi = x0 + dx, y0 + dy, z0 + dz
                w = (fx if dx else 1 - fx) * (fy if dy else 1 - fy) * (fz if dz else 1 - fz)
                ok = (xi < W) & (yi < H) & (zi < D)
                idx = (zi.clamp(max=D - 1) * H + yi.clamp(max=H - 1)) * W + xi.clamp(max=W - 1)
                val = torch.gather(flat, 2, idx.view(N, 1, -1).expand(N, C, -1)).view(N, C, D, H, W)
                out = out + val * (w * ok).unsqueeze(1)
    return out


class _BorderClamp(torch.autograd.Function):
    """clip to [0, n-1] with zero gradient on and outside the border (GridSampler.cuh:62-81)"""

    @staticmethod
    def forward(ctx, c, n):
        ctx.save_for_backward((c > 0) & (c < n - 1))
        return torch.clamp(c, 0, n - 1)

    @staticmethod
    def backward(ctx, g):
        inside, = ctx.saved_tensors
        return g * inside, None


def svf_exp_voxel(v, no_steps=12):
    """
    voxel-unit form of scaling and squaring, independent of F.grid_sample:
    u0 = v / 2^steps; u <- u + trilinear(u)(clamp(index + u)); returns displacement in voxels.
    Identical to svf_exp_aten in exact arithmetic (cubic volumes).
    """
    N, _, D, H, W = v.shape
    iz, iy, ix = torch.meshgrid(torch.arange(D, dtype=v.dtype), torch.arange(H, dtype=v.dtype),
                                torch.arange(W, dtype=v.dtype), indexing='ij')
    u = v / float(2 ** no_steps)
    for _ in range(no_steps):
        px = _BorderClamp.apply(ix + u[:, 0], W)
        py = _BorderClamp.apply(iy + u[:, 1], H)
        pz = _BorderClamp.apply(iz + u[:, 2], D)
        u = u + trilinear_sample_voxel(u, px, py, pz)
    return u


def nearest_index(T):
    """
    nearest-neighbour source index of every output voxel: unnormalise -> clip -> round half to even
    (GridSampler.cuh:21-31,53-57 + nearbyint), computed in fp32 exactly in ATen's operation order
    """
    T = T.float()
    D, H, W = T.shape[2:]
    px, py, pz = unnormalise(T)
    ix, iy, iz = torch.round(px).long(), torch.round(py).long(), torch.round(pz).long()  # torch.round = half to even
    return (iz * H + iy) * W + ix


def warp_nearest(seg, T):
    idx = nearest_index(T)
    N = T.shape[0]
    flat = seg.expand(N, *seg.shape[1:]).reshape(N, -1)
    return torch.gather(flat, 1, idx.view(N, -1)).view(N, 1, *T.shape[2:])


def uniform_jitter_normalised(jitter_unit, alpha, shape):
    """
    reference utils/util.py:44-45,52-53: T + normalised(-2 alpha U + alpha), U ~ U[0,1) supplied by the caller
    """
    return to_normalised(-2.0 * alpha * jitter_unit + alpha)


# ----------------------------------------------------------------------------------------------------------------------
# data term: LCC map, GMM, virtual decimation
# ----------------------------------------------------------------------------------------------------------------------

def box_sum(x, s):
    """sum over the (2s+1)^3 window with replicate padding (the all-ones nn.Conv3d of reference model/loss.py:54-59)"""
    out = x
    for axis in (2, 3, 4):
        acc = torch.zeros_like(out)
        for o in range(-s, s + 1):
            acc = acc + _shift_replicate(out, axis, o)
        out = acc
    return out


def lcc_normalise(im, s):
    """reference model/loss.py:103-105: (I - u) / sqrt(Box((I-u)^2)/k^3 + 1e-10),  u = Box(I)/k^3"""
    sz = float((2 * s + 1) ** 3)
    u = box_sum(im, s) / sz
    var = box_sum((im - u) ** 2, s) / sz
    return (im - u) / torch.sqrt(var + 1e-10)


def lcc_map(im_fixed, im_moving, s):
    """reference model/loss.py:102-111"""
    return lcc_normalise(im_fixed, s) - lcc_normalise(im_moving, s)


def ssd_map(im_fixed, im_moving):
    """the SSD data term (SURVEY surprise 1: build-defined DataLoss with map = F - M, K = 1)"""
    return im_fixed - im_moving


def gmm_log_proportions(logits):
    """reference model/loss.py:67-69"""
    return torch.log_softmax(logits + 1e-2, dim=0)


def gmm_log_pdf(z, log_std, logits):
    """reference model/loss.py:87-93; z any shape -> (1, V)"""
    E = 0.5 * (z.reshape(1, -1, 1) * torch.exp(-log_std)) ** 2
    return torch.logsumexp((gmm_log_proportions(logits) - log_std - LOG_SQRT_2PI) - E, dim=-1)


def gmm_nll(z, log_std, logits):
    """reference model/loss.py:113-114"""
    return -gmm_log_pdf(z, log_std, logits).sum()


def gmm_responsibilities(z, log_std, logits):
    E = 0.5 * (z.unsqueeze(-1) * torch.exp(-log_std)) ** 2
    return torch.softmax((gmm_log_proportions(logits) - log_std) - E, dim=-1)


def vd_residual(z, mask, log_std, logits):
    """
    reference utils/util.py:330-347 in closed form: r = z^2 sum_k rho_k(z)/sigma_k^2 on the mask, 0 elsewhere
    """
    zm = torch.where(mask, z, torch.zeros_like(z))
    rho = gmm_responsibilities(zm, log_std, logits)
    return zm ** 2 * (rho * torch.exp(-2.0 * log_std)).sum(-1)


def vd_factor(r, mask):
    """reference utils/util.py:446-485; r, mask of shape (1,1,D,H,W)"""
    n = mask.sum()
    var = (r[mask] ** 2).mean()
    rm = torch.where(mask, r, torch.zeros_like(r))
    cov = [(rm[:, :, :-1] * rm[:, :, 1:]).sum() / n, (rm[:, :, :, :-1] * rm[:, :, :, 1:]).sum() / n,
           (rm[:, :, :, :, :-1] * rm[:, :, :, :, 1:]).sum() / n]
    sq = [torch.clamp(-2.0 / math.pi * torch.log(c / var), max=1.0) for c in cov]
    return torch.sqrt(sq[0] * sq[1] * sq[2])


def normal_log_pdf(x, loc, scale):
    """reference model/distributions.py:56-58"""
    return -0.5 * ((x - loc) / scale) ** 2 - math.log(scale) - LOG_SQRT_2PI


def dirichlet_log_pdf(log_p, alpha):
    """reference model/distributions.py:209-211, symmetric concentration alpha"""
    K = log_p.shape[-1]
    return (log_p * (alpha - 1.0)).sum(-1) + math.lgamma(K * alpha) - K * math.lgamma(alpha)


# ----------------------------------------------------------------------------------------------------------------------
# regulariser
# ----------------------------------------------------------------------------------------------------------------------

def forward_differences(v, transformation=False):
    """
    reference utils/diff_op.py:78-96: forward differences with the last one replicated; output (N,3,D,H,W,3) with
    [n, j, ..., i] = d v_i / d x_j ; with transformation=True divided by the spacing 2/(dims-1) (x by dims[2], ...)
    """
    def diff(axis):
        n = v.shape[axis]
        d = v.narrow(axis, 1, n - 1) - v.narrow(axis, 0, n - 1)
        return torch.cat((d, d.narrow(axis, n - 2, 1)), axis)

    dx, dy, dz = diff(4), diff(3), diff(2)
    if transformation:
        D, H, W = v.shape[2:]
        dx, dy, dz = dx / (2.0 / (W - 1)), dy / (2.0 / (H - 1)), dz / (2.0 / (D - 1))
    per_component = [torch.stack((dx[:, i], dy[:, i], dz[:, i]), 1) for i in range(3)]
    return torch.stack(per_component, -1)


def reg_energy(v):
    """reference model/loss.py:158-159: y_c = sum |D v|^2"""
    return (forward_differences(v) ** 2).sum(dim=(1, 2, 3, 4, 5))


def det_jacobian(nabla):
    """reference utils/util.py:72-91"""
    a, b, c = nabla[..., 0], nabla[..., 1], nabla[..., 2]
    return a[:, 0] * b[:, 1] * c[:, 2] + b[:, 0] * c[:, 1] * a[:, 2] + c[:, 0] * a[:, 1] * b[:, 2] \
        - a[:, 2] * b[:, 1] * c[:, 0] - b[:, 2] * c[:, 1] * a[:, 0] - c[:, 2] * a[:, 1] * b[:, 0]


def no_non_diffeomorphic_voxels(T):
    """reference utils/util.py:209-212: log det J of the forward-difference Jacobian of a transformation; a voxel counts as
    folded when the logarithm is NaN (det J < 0).  Returns (counts per sample, log det J)"""
    log_det = det_jacobian(forward_differences(T, transformation=True)).log()
    return torch.isnan(log_det).sum(dim=(1, 2, 3)), log_det


def dice_scores(seg_fixed, seg_moving, labels):
    """reference utils/util.py:123-148 (calc_DSC_GPU): per sample and structure 2 |A & B| / (|A| + |B|); 0 / 0 is NaN there
    too (tensor division does not raise, so the `except` branch never runs)"""
    out = torch.zeros(seg_moving.shape[0], len(labels))
    for i in range(seg_moving.shape[0]):
        a = seg_fixed[min(i, seg_fixed.shape[0] - 1)]
        for j, label in enumerate(labels):
            fa, mb = a == label, seg_moving[i] == label
            out[i, j] = 2.0 * (fa & mb).sum() / (fa.sum() + mb.sum())
    return out


def field_norm(field):
    """reference utils/util.py:215-225 (calc_norm): voxel-wise Euclidean norm, (N,1,D,H,W)"""
    return torch.linalg.vector_norm(field, ord=2, dim=1, keepdim=True)


def expgamma_log_pdf(x, shape, rate):
    """reference model/distributions.py:111-112,167-168"""
    return shape * math.log(rate) + (shape - 1) * x - rate * torch.exp(x) - math.lgamma(shape) + x


def lognormal_init(w_reg, dof):
    """reference model/loss.py:300-305 + model/distributions.py:171-172,241-242: (loc, log_scale)"""
    # nu and w_reg are fp32 tensors in the reference (model/distributions.py:234-236): the rate and its log are fp32
    log_rate = float(torch.log(0.5 * torch.tensor(1.0) * torch.tensor(w_reg, dtype=torch.float32)))
    loc = float(torch.digamma(torch.tensor(0.5 * dof, dtype=torch.float64))) - log_rate
    return loc, math.log(4.0) + math.log(loc)


# ----------------------------------------------------------------------------------------------------------------------
# the reference's Adam with rate decay
# ----------------------------------------------------------------------------------------------------------------------

class AdamState:
    """reference optimizers/adam_rate_decay.py:32-99 (no amsgrad, no weight decay, no re-init after the first step)"""

    def __init__(self, params, lrs, lr_decay, betas=(0.9, 0.999), eps=1e-8):
        self.params, self.lrs, self.lr_decay, self.betas, self.eps = params, lrs, lr_decay, betas, eps
        self.step_no = 0
        self.m = [torch.zeros_like(p) for p in params]
        self.v = [torch.zeros_like(p) for p in params]

    def step(self, grads):
        b1, b2 = self.betas
        decay = 1 + self.step_no * self.lr_decay
        self.step_no += 1
        bc1, bc2 = 1 - b1 ** self.step_no, 1 - b2 ** self.step_no
        for p, g, m, v, lr in zip(self.params, grads, self.m, self.v, self.lrs):
            m.mul_(b1).add_(g, alpha=1 - b1)
            v.mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = (v.sqrt() / math.sqrt(bc2)).add_(self.eps)
            p.addcdiv_(m, denom, value=-(lr / decay) / bc1)


# ----------------------------------------------------------------------------------------------------------------------
# the SGLD transition
# ----------------------------------------------------------------------------------------------------------------------

class Config:
    """hyper-parameters of the hot path; defaults = reference configs/experiment3/config.json"""

    def __init__(self, data='lcc', K=4, s=2, reg='lognormal', w_reg=1.6, reg_learnable=True, sobolev_s=3,
                 sobolev_lambda=0.5, svf_steps=12, tau=0.4, jitter_alpha=0.1, virtual_decimation=True, lr_gmm=0.2,
                 lr_reg=0.01, lr_decay=1e-3, exact_grid=False, cps=None):
        # cps: control point spacing of SVFFD_3D as the transformation module (configs/experiment5); None = SVF_3D
        self.__dict__.update(locals())
        del self.__dict__['self']


class State:
    """everything a chain group carries between iterations (SURVEY §5 'checkpoint/resume' row)"""

    def __init__(self, cfg, v, sigma, dims, dtype=torch.float32):
        # dims: the IMAGE size (dof of the regulariser, model/loss.py:134); with cfg.cps the state v lives on the control grid
        self.cfg, self.v, self.sigma, self.dims = cfg, v.clone(), sigma, tuple(dims)
        K = cfg.K if cfg.data == 'lcc' else 1
        self.log_std, self.logits = torch.zeros(K, dtype=dtype), torch.zeros(K, dtype=dtype)
        self.adam_gmm = AdamState([self.log_std, self.logits], [cfg.lr_gmm, cfg.lr_gmm], cfg.lr_decay)
        self.dof = float(np.prod(dims) * 3.0)
        if cfg.reg == 'lognormal':
            loc, log_scale = lognormal_init(cfg.w_reg, self.dof)
            self.loc = torch.tensor(loc, dtype=torch.float64)            # fp64 in the reference too (dof is np.float64)
            self.log_scale = torch.tensor(log_scale, dtype=torch.float64)
            self.adam_reg = AdamState([self.loc, self.log_scale], [cfg.lr_reg, cfg.lr_reg], cfg.lr_decay)
        else:
            self.log_w_reg = torch.tensor(math.log(cfg.w_reg), dtype=dtype)
            self.adam_reg = AdamState([self.log_w_reg], [cfg.lr_reg], cfg.lr_decay)
        self.taps = sobolev_taps(cfg.sobolev_s, cfg.sobolev_lambda).astype(np.float32)  # trainer.py:573 `.float()`

    def init_gmm(self, sigma_hat):
        """reference model/loss.py:61-65.  SSD (build-defined, SURVEY surprise 1: the reference ships no SSD class): ONE Gaussian at
        the residuals' own scale, log_std = log sigma_hat -- the mixture's linspace with a single point would start at
        sigma_hat / 100, a precision of 1e4 / sigma_hat^2 that throws a chain across the volume within a few transitions."""
        K = self.log_std.numel()
        if self.cfg.data == 'ssd':
            self.log_std.fill_(math.log(sigma_hat))
            return
        self.log_std.copy_(torch.linspace(math.log(sigma_hat / 100.0), math.log(sigma_hat * 5.0), steps=K))


def gmm_step(st, z_masked, alpha):
    """reference trainer/trainer.py:68-77: one Adam step on (log_std, logits) with detached residuals"""
    ls = st.log_std.clone().requires_grad_(True)
    lg = st.logits.clone().requires_grad_(True)
    loss = gmm_nll(z_masked.detach(), ls, lg) * alpha
    loss = loss - normal_log_pdf(ls, 0.0, 2.3).sum() - dirichlet_log_pdf(gmm_log_proportions(lg), 0.5)
    g_ls, g_lg = torch.autograd.grad(loss, (ls, lg))
    st.adam_gmm.step([g_ls, g_lg])


def reg_term_fn(st, y):
    """
    per-chain regularisation loss and the hyper-prior total, reference model/loss.py:197-198,264-270,311-312 and
    trainer/trainer.py:334-339.  Returns (per-chain loss (C,), total to differentiate, hyper-parameter leaves)
    """
    cfg = st.cfg
    log_y = y.log()
    if cfg.reg == 'lognormal':
        loc, log_scale = st.loc.clone().requires_grad_(True), st.log_scale.clone().requires_grad_(True)
        per_chain = log_y + log_scale + 0.5 * ((log_y - loc) / log_scale.exp()) ** 2 + (0.5 * st.dof - 1.0) * log_y
        per_chain = per_chain.to(y.dtype)
        total = per_chain.sum()
        if cfg.reg_learnable:
            total = total - expgamma_log_pdf(log_y, 0.5 * st.dof, 0.5 * cfg.w_reg).sum()
            total = total - normal_log_pdf(log_scale, 2.8, 5.0)
        return per_chain, total, (loc, log_scale)

    lw = st.log_w_reg.clone().requires_grad_(True)
    per_chain = 0.5 * lw.exp() * y - 0.5 * st.dof * lw
    total = per_chain.sum()
    if cfg.reg_learnable:
        shape = 0.5 * st.dof
        total = total - expgamma_log_pdf(lw, shape, 1.0 / shape)
    return per_chain, total, (lw,)


def sgld_transition(st, fixed, moving, eps=None, jitter_unit=None):
    """
    reference trainer/trainer.py:291-356.  `fixed`/`moving` are the (1,1,D,H,W) dicts; chains are dim 0 of st.v.
    eps (C,3,D,H,W) ~ N(0,1) and jitter_unit (C,3,D,H,W) ~ U[0,1) are drawn here unless supplied.
    Mutates st (v, GMM and regulariser hyper-parameters, Adam states) and returns (loss_terms, output, aux, grad_v).
    """
    cfg = st.cfg
    C = st.v.shape[0]
    dtype = st.v.dtype
    if eps is None:
        eps = torch.randn_like(st.v)
    if cfg.jitter_alpha is not None and jitter_unit is None:
        jitter_unit = torch.rand(C, 3, *st.dims, dtype=dtype)

    # trainer.py:292-293.  SGLD.backward = sigma^2 * g and SobolevGrad.backward = identity (utils/functions.py:82-84,
    # 107-109), i.e. dL/dv := sigma^2 * dL/d(css): differentiate w.r.t. css and scale afterwards.
    tau_noise = math.sqrt(2.0 * cfg.tau) * st.sigma * eps if cfg.tau > 0 else 0.0
    css = sobolev_smooth(st.v + tau_noise, st.taps).detach().requires_grad_(True)

    velocity = css if cfg.cps is None else ffd_dense(css, st.dims, cfg.cps)         # utils/transformation.py:163-164
    T, disp = svf_exp_aten(velocity, cfg.svf_steps, cfg.exact_grid)                 # trainer.py:294
    T_s = T + uniform_jitter_normalised(jitter_unit, cfg.jitter_alpha, T.shape) if cfg.jitter_alpha is not None else T
    im_w = warp_aten(moving['im'].expand(C, -1, -1, -1, -1).to(dtype), T_s)          # trainer.py:296-300

    F_im = fixed['im'].to(dtype)
    mask = fixed['mask']
    z = lcc_map(F_im, im_w, cfg.s) if cfg.data == 'lcc' else ssd_map(F_im, im_w)     # trainer.py:307
    y = reg_energy(css)                                                             # trainer.py:311

    data_terms, alphas = [], []
    data_total = 0.0
    for c in range(C):                                                              # trainer.py:316-327
        zc = z[c:c + 1]
        if cfg.virtual_decimation:
            alpha = vd_factor(vd_residual(zc.detach(), mask, st.log_std, st.logits), mask)
        else:
            alpha = torch.tensor(1.0, dtype=dtype)
        gmm_step(st, zc[mask], alpha)
        term = gmm_nll(zc[mask], st.log_std, st.logits) * alpha
        data_total = data_total + term
        data_terms.append(term.detach())
        alphas.append(alpha)

    reg_per_chain, reg_total, hyper = reg_term_fn(st, y)
    loss = data_total + reg_total                                                   # trainer.py:342

    grads = torch.autograd.grad(loss, (css,) + (hyper if cfg.reg_learnable else ()))
    grad_v = st.sigma ** 2 * grads[0]
    st.v = st.v - cfg.tau * grad_v                                                  # trainer.py:351 (plain SGD)
    if cfg.reg_learnable:                                                           # trainer.py:353-354
        st.adam_reg.step([g.to(p.dtype) for g, p in zip(grads[1:], st.adam_reg.params)])

    loss_terms = {'data': data_terms, 'reg': [r.detach() for r in reg_per_chain]}
    output = {'im_moving_warped': im_w.detach(), 'displacement': disp.detach(), 'transformation': T.detach(),
              'curr_state': css.detach()}
    aux = {'residuals': z.detach(), 'alpha': alphas, 'reg_energy': [yy.detach() for yy in y]}
    return loss_terms, output, aux, grad_v.detach()


def gmm_init(st, fixed, moving, v_sample, warm_up=25):
    """reference trainer/trainer.py:529-547"""
    cfg = st.cfg
    css = sobolev_smooth(v_sample, st.taps)
    T, _ = svf_exp_aten(css if cfg.cps is None else ffd_dense(css, st.dims, cfg.cps), cfg.svf_steps, cfg.exact_grid)
    im_w = warp_aten(moving['im'].to(css.dtype), T)
    z = lcc_map(fixed['im'].to(css.dtype), im_w, cfg.s) if cfg.data == 'lcc' else ssd_map(fixed['im'], im_w)
    mask = fixed['mask']
    st.init_gmm(float(torch.std(z[mask])))
    alpha = vd_factor(vd_residual(z, mask, st.log_std, st.logits), mask) if cfg.virtual_decimation else 1.0
    for _ in range(warm_up):
        gmm_step(st, z[mask], alpha)
    return z, alpha


def posterior_statistics(samples):
    """reference utils/util.py:114-120: mean and unbiased std over dim 0"""
    return samples.mean(0), samples.std(0)


def welford_merge(parts):
    """Chan et al. pairwise merge of (n, mean, M2) triples -- the reduction the multi-GPU path implements with NCCL"""
    n, mean, m2 = parts[0]
    for nb, mb, m2b in parts[1:]:
        tot = n + nb
        delta = mb - mean
        mean = mean + delta * (nb / tot)
        m2 = m2 + m2b + delta ** 2 * (n * nb / tot)
        n = tot
    return n, mean, m2


# ----------------------------------------------------------------------------------------------------------------------
# VI warm start: one iteration of Trainer._run_VI (reference trainer/trainer.py:79-223)
# ----------------------------------------------------------------------------------------------------------------------

class _SobolevIdentityBackward(torch.autograd.Function):
    """SobolevGrad: smoothing in the forward pass, identity in the backward pass (reference utils/functions.py:98-109)"""

    @staticmethod
    def forward(ctx, x, taps):
        return sobolev_smooth(x, taps)

    @staticmethod
    def backward(ctx, g):
        return g, None


def entropy_terms(log_var, u, sample=None, mu=None):
    """reference model/loss.py:350-372"""
    sigma = torch.exp(0.5 * log_var)
    dims = (1, 2, 3, 4)
    if sample is None:
        return 0.5 * (torch.log1p(torch.sum((u / sigma) ** 2, dim=dims)) + torch.sum(log_var, dim=dims))
    sn, un = (sample - mu) / sigma, u / sigma
    return 0.5 * (torch.sum(sn ** 2, dim=dims) - torch.sum(sn * un, dim=dims) ** 2 / (1.0 + torch.sum(un ** 2, dim=dims)))


def vi_sample_loss(st, fixed, moving, vp, v_unsmoothed, jitter_unit, reg_leaves):
    """reference trainer/trainer.py:79-117; steps the shared mixture like the reference does (detached residuals)"""
    cfg = st.cfg
    dtype = v_unsmoothed.dtype
    css = _SobolevIdentityBackward.apply(v_unsmoothed, st.taps)
    T, disp = svf_exp_aten(css, cfg.svf_steps, cfg.exact_grid)
    T_s = T + uniform_jitter_normalised(jitter_unit, cfg.jitter_alpha, T.shape) if cfg.jitter_alpha is not None else T
    im_w = warp_aten(moving['im'].to(dtype), T_s)
    F_im, mask = fixed['im'].to(dtype), fixed['mask']
    z = lcc_map(F_im, im_w, cfg.s) if cfg.data == 'lcc' else ssd_map(F_im, im_w)
    if cfg.virtual_decimation:
        alpha = vd_factor(vd_residual(z.detach(), mask, st.log_std, st.logits), mask)
    else:
        alpha = torch.tensor(1.0, dtype=dtype)
    gmm_step(st, z[mask], alpha)
    data = gmm_nll(z[mask], st.log_std, st.logits) * alpha
    y = reg_energy(css)
    log_y = y.log()
    terms = {'data': data, 'alpha': alpha, 'energy': y.detach(), 'im_w': im_w.detach(), 'disp': disp.detach()}
    if cfg.reg == 'lognormal':
        loc, log_scale = reg_leaves
        reg = (log_y + log_scale + 0.5 * ((log_y - loc) / log_scale.exp()) ** 2 + (0.5 * st.dof - 1.0) * log_y).sum()
        if cfg.reg_learnable:
            terms['reg_loc_prior'] = expgamma_log_pdf(log_y, 0.5 * st.dof, 0.5 * float(np.float32(cfg.w_reg))).sum()
    else:
        lw, = reg_leaves
        reg = (0.5 * lw.exp() * y - 0.5 * st.dof * lw).sum()
        if cfg.reg_learnable:
            shape = 0.5 * st.dof
            terms['w_reg_prior'] = expgamma_log_pdf(lw, shape, 1.0 / shape)
    terms['reg'] = reg
    terms['entropy'] = entropy_terms(vp['log_var'], vp['u'], v_unsmoothed, vp['mu']).sum()
    return terms


def vi_iteration(st, fixed, moving, vp, eps, x, jitter1, jitter2):
    """
    one iteration of reference trainer/trainer.py:130-171 without the q(v) optimiser step: returns the loss terms and the
    gradients w.r.t. (mu, log_var, u); steps the mixture (twice) and the regulariser hyper-parameters.
    vp: dict of (1,3,D,H,W) tensors; eps ~ N(0,1) (1,3,D,H,W); x ~ N(0,1) scalar tensor.
    """
    cfg = st.cfg
    vp = {k: v.detach().clone().requires_grad_(True) for k, v in vp.items()}
    sigma = torch.exp(0.5 * vp['log_var'])
    delta = eps * sigma + x * vp['u']
    if cfg.reg == 'lognormal':
        leaves = (st.loc.clone().requires_grad_(True), st.log_scale.clone().requires_grad_(True))
    else:
        leaves = (st.log_w_reg.clone().requires_grad_(True),)
    t1 = vi_sample_loss(st, fixed, moving, vp, vp['mu'] + delta, jitter1, leaves)
    t2 = vi_sample_loss(st, fixed, moving, vp, vp['mu'] - delta, jitter2, leaves)
    data = (t1['data'] + t2['data']) / 2.0
    reg = (t1['reg'] + t2['reg']) / 2.0
    if cfg.reg_learnable:
        if cfg.reg == 'lognormal':
            reg = reg - (t1['reg_loc_prior'] + t2['reg_loc_prior']) / 2.0 - normal_log_pdf(leaves[1], 2.8, 5.0)
        else:
            reg = reg - (t1['w_reg_prior'] + t2['w_reg_prior']) / 2.0
    entropy = (t1['entropy'] + t2['entropy']) / 2.0 + entropy_terms(vp['log_var'], vp['u']).sum()
    loss = data + reg - entropy
    wanted = (vp['mu'], vp['log_var'], vp['u']) + (leaves if cfg.reg_learnable else ())
    grads = torch.autograd.grad(loss, wanted)
    if cfg.reg_learnable:
        st.adam_reg.step([g.to(p.dtype) for g, p in zip(grads[3:], st.adam_reg.params)])
    return {'data': data.detach(), 'reg': reg.detach(), 'entropy': entropy.detach(), 'loss': loss.detach(),
            'alpha': (t1['alpha'], t2['alpha']), 'im_w': t1['im_w'], 'disp': t1['disp']}, \
        {'mu': grads[0], 'log_var': grads[1], 'u': grads[2]}
