"""
Helpers of the SGLD hot path with the reference's names (reference utils/util.py).  Volume-sized work goes to the CUDA
kernels of libirsgmcmc.so; what is left in torch is scalar or book-keeping glue.  File / VTK / SimpleITK helpers of the
reference are out of scope.
"""
import math

import numpy as np
import torch

from .. import ops


def get_noise_uniform(shape, device, alpha):
    """reference :52-53"""
    return -2.0 * alpha * torch.rand(shape, device=device) + alpha


def get_noise_Langevin(sigma, tau):
    """reference :56-58"""
    return math.sqrt(2.0 * tau) * sigma * torch.randn_like(sigma)


def add_noise_uniform_field(field, alpha):
    """reference :44-45"""
    return field + transform_coordinates(get_noise_uniform(field.shape, field.device, alpha))


def add_noise_Langevin(field, sigma, tau):
    """reference :48-49"""
    return field + get_noise_Langevin(sigma, tau)


def get_control_grid_size(dims, cps):
    """control grid size of a B-spline FFD with control point spacing cps (reference utils/util.py:61-69)"""
    return tuple([int(math.ceil((sz - 1) / c) + 1 + 2) for sz, c in zip(dims, cps)])


def transform_coordinates(field):
    """voxel units -> normalised units: channel i times 2 / (shape[2 + i] - 1)  (reference :418-429)"""
    scale = torch.tensor([2.0 / float(n - 1) for n in field.shape[2:]], device=field.device, dtype=field.dtype)
    return field * scale.view(1, -1, *([1] * (field.dim() - 2)))


def transform_coordinates_inv(field):
    """reference :432-443"""
    scale = torch.tensor([float(n - 1) / 2.0 for n in field.shape[2:]], device=field.device, dtype=field.dtype)
    return field * scale.view(1, -1, *([1] * (field.dim() - 2)))


def init_identity_grid_3D(dims):
    """(1, nz, ny, nx, 3) identity sampling grid, last dim (x, y, z), from fp32 torch.linspace (reference :263-278)"""
    nx, ny, nz = dims[0], dims[1], dims[2]
    x, y, z = (torch.linspace(-1, 1, steps=n) for n in (nx, ny, nz))
    gz, gy, gx = torch.meshgrid(z, y, x, indexing='ij')
    return torch.stack((gx, gy, gz), 3).unsqueeze(0)


def separable_conv_3D(field, *args):
    """
    separable smoothing of a vector field with replicate padding.  Two call forms like the reference (:350-406):
    (field, kernel_1d, padding_sz) with a (3,1,k) conv1d weight, or (field, S_x, S_y, S_z, padding6) with the three
    depthwise conv3d weights.  Same taps on every axis and channel are assumed (all the reference ever builds).
    """
    from .functions import langevin_sobolev
    kernel = args[0]
    taps = [float(t) for t in kernel[0].reshape(-1).tolist()]
    return langevin_sobolev(field.contiguous(), None, 0.0, taps)


def rescale_residuals(res, mask, data_loss):
    """
    precision-weighted squared residuals r = z^2 sum_k rho_k(z) / sigma_k^2 on the mask, 0 elsewhere -- the closed form
    of the reference's inner autograd pass (:330-347)
    """
    z = res.detach().contiguous()
    _, dz, _ = ops.gmm_log_pdf(z.reshape(-1), data_loss.log_std, data_loss.logits, want_dz=True)
    r = (-z.reshape(-1) * dz).view(res.shape)
    return torch.where(mask, r, torch.zeros_like(r))


@torch.no_grad()
def calc_VD_factor(residual, mask):
    """virtual decimation factor from the rescaled residual field (reference :446-485)"""
    return ops.vd_factor_from_residual(residual.contiguous(), mask.contiguous()).float()


def calc_det_J(nabla):
    """Jacobian determinant of field gradients (reference :72-91)"""
    a, b, c = nabla[..., 0], nabla[..., 1], nabla[..., 2]
    return a[:, 0] * b[:, 1] * c[:, 2] + b[:, 0] * c[:, 1] * a[:, 2] + c[:, 0] * a[:, 1] * b[:, 2] \
        - a[:, 2] * b[:, 1] * c[:, 0] - b[:, 2] * c[:, 1] * a[:, 0] - c[:, 2] * a[:, 1] * b[:, 0]


def calc_no_non_diffeomorphic_voxels(transformation, diff_op):
    """number of voxels with a NaN log det J per sample, and log det J (reference :209-212): one fused kernel over the
    forward-difference GradientOperator (no (N,3,D,H,W,3) gradient tensor is materialised).  There is no PyTorch path:
    another operator, a CPU tensor or another dtype raises."""
    from .diff_op import GradientOperator
    if type(diff_op) is not GradientOperator:
        raise NotImplementedError('calc_no_non_diffeomorphic_voxels: only GradientOperator has a CUDA kernel '
                                  '(the reference ships no other working operator, utils/diff_op.py)')
    if not transformation.is_cuda or transformation.dtype != torch.float32:
        raise NotImplementedError('calc_no_non_diffeomorphic_voxels needs a float32 CUDA tensor: irsgmcmc_b200 has no CPU path')
    counts, log_det_J = ops.log_det_jacobian(transformation.detach().contiguous())
    return counts.cpu().numpy(), log_det_J


def calc_norm(field):
    """voxel-wise Euclidean norm, (N,1,D,H,W) (reference :215-225)"""
    return torch.linalg.vector_norm(field, ord=2, dim=1, keepdim=True)


@torch.no_grad()
def calc_posterior_statistics(samples, device='cuda:0'):
    """mean and unbiased std over dim 0 through the Welford kernels (reference :114-120)"""
    samples = samples.to(device).contiguous()
    mean, m2 = torch.zeros_like(samples[0]), torch.zeros_like(samples[0])
    count = ops.welford_update(samples, 0, mean, m2)
    return mean, ops.welford_std(m2, count)


@torch.no_grad()
def calc_DSC_GPU(no_samples, seg_fixed, seg_moving, structures_dict):
    """Dice score per sample and structure (reference :123-148): one pass over the volumes for all samples and up to 32
    structures at a time (the reference loops samples x structures).  int16 CUDA segmentations only -- no PyTorch path."""
    labels = [int(x) for x in structures_dict.values()]
    a, b = seg_fixed[:no_samples], seg_moving[:no_samples].contiguous()
    if not (a.is_cuda and b.is_cuda) or a.dtype != torch.int16 or b.dtype != torch.int16:
        raise NotImplementedError('calc_DSC_GPU needs int16 CUDA segmentations: irsgmcmc_b200 has no CPU path')
    if 0 in labels:
        raise NotImplementedError('calc_DSC_GPU: label 0 is the background of the counting kernel')
    a = a[:1].contiguous() if a.stride(0) == 0 else a.contiguous()
    cnt = torch.cat([ops.dice_counts(a, b, labels[i:i + 32]) for i in range(0, len(labels), 32)], dim=1).double().cpu()
    return (2.0 * cnt[..., 2] / (cnt[..., 0] + cnt[..., 1])).float().numpy()


def _label_contour(mask):
    """ITK LabelContourImageFilter with its defaults on a binary image: a foreground voxel belongs to the contour when one of its
    six face neighbours is background (FullyConnectedOff); voxels on the image border count as touching background."""
    from scipy import ndimage
    return mask & ~ndimage.binary_erosion(mask, structure=ndimage.generate_binary_structure(3, 1), border_value=0)


def calc_ASD_host(seg_fixed_structure, seg_moving_structure, spacing=(1.0, 1.0, 1.0)):
    """
    Average surface distance of two binary structures the way the reference computes it (utils/util.py:171-176):
    sitk.LabelContour of both, then HausdorffDistanceImageFilter.GetAverageHausdorffDistance() = the mean of the two directed
    average distances, each the mean over the contour voxels of one image of the Euclidean distance (image spacing applied) to
    the nearest contour voxel of the other.  Host-side like the reference's (SimpleITK there, scipy's exact Euclidean distance
    transform here).  PARITY UNPINNED: SimpleITK is absent from the build image, so this restates ITK's documented algorithm and is
    checked against a brute-force evaluation of the same definition (tests/test_writers.py), not against SimpleITK itself.
    Returns inf when a structure is empty (the reference's `except: ASD = inf`).
    """
    from scipy import ndimage
    a, b = np.asarray(seg_fixed_structure).astype(bool), np.asarray(seg_moving_structure).astype(bool)
    if not a.any() or not b.any():
        return float('inf')
    # crop to the bounding box of both structures (+ 1 voxel so that the contour sees their background): every contour voxel of
    # either image lies inside it, so the distances are unchanged and the distance transforms stay small
    idx = np.argwhere(a | b)
    lo, hi = np.maximum(idx.min(0) - 1, 0), np.minimum(idx.max(0) + 2, a.shape)
    box = tuple(slice(int(l), int(h)) for l, h in zip(lo, hi))
    a, b = a[box], b[box]
    ca, cb = _label_contour(a), _label_contour(b)
    if not ca.any() or not cb.any():
        return float('inf')
    sp = [float(x) for x in np.asarray(spacing).reshape(-1)[:3]]
    # array axes are (z, y, x) of an image whose spacing is given as (x, y, z): sitk.GetImageFromArray reverses the axes
    sampling = sp[::-1]
    d_to_b = ndimage.distance_transform_edt(~cb, sampling=sampling)
    d_to_a = ndimage.distance_transform_edt(~ca, sampling=sampling)
    return 0.5 * (float(d_to_b[ca].mean()) + float(d_to_a[cb].mean()))


@torch.no_grad()
def calc_metrics(seg_fixed, seg_moving, structures_dict, spacing, GPU=True, no_samples=1):
    """(ASD, DSC) per sample and structure like the reference's calc_metrics (:151-206).  DSC: the counting kernel
    (calc_DSC_GPU; `GPU=False` is not offered -- there is no host Dice).  ASD: calc_ASD_host per sample and structure on the
    host, as in the reference (parity unpinned, see there); needs scipy."""
    if not GPU:
        raise NotImplementedError('calc_metrics(GPU=False): the Dice scores come from the CUDA counting kernel only')
    DSC = calc_DSC_GPU(no_samples, seg_fixed, seg_moving, structures_dict)
    a, b = seg_fixed[:no_samples].cpu().numpy(), seg_moving[:no_samples].cpu().numpy()
    sp = spacing.numpy().tolist() if hasattr(spacing, 'numpy') else list(spacing)
    ASD = np.zeros([no_samples, len(structures_dict)])
    for idx in range(no_samples):
        fa, mb = a[min(idx, a.shape[0] - 1)].squeeze(), b[idx].squeeze()
        for j, label in enumerate(structures_dict.values()):
            ASD[idx, j] = calc_ASD_host(fa == label, mb == label, sp)
    return ASD, DSC
