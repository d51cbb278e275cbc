"""
Transformation models with the reference's interface (reference utils/transformation.py).  SVF_3D integrates a
stationary velocity field by scaling and squaring on the CUDA kernels of libirsgmcmc.so; its backward pass is the
gather-form adjoint (no scatter atomics).  Cubic_B_spline_FFD_3D / SVFFD_3D put the cubic B-spline control-point
parametrisation in front of it (irs_ffd.cu: three axis passes, adjoint as gathers).
"""
from abc import ABC, abstractmethod

import torch
from torch import nn

from .. import ops
from .util import init_identity_grid_3D


class TransformationModule(nn.Module, ABC):
    """abstract transformation model"""

    @abstractmethod
    def forward(self, v):
        pass


class _ScalingAndSquaring(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v, no_steps, lin, spacing):
        v_c = v.contiguous()
        hist, maxabs = ops.svf_exp_fwd(v_c, no_steps)
        ctx.save_for_backward(v_c, hist, maxabs, spacing)
        displacement = hist[-1]
        transformation = ops.svf_outputs(displacement, lin)
        return transformation, displacement.clone()

    @staticmethod
    def backward(ctx, g_transformation, g_displacement):
        v, hist, maxabs, spacing = ctx.saved_tensors
        # T = id + spacing * u_n and displacement = u_n
        g_u = torch.zeros_like(v)
        if g_transformation is not None:
            g_u = g_u + g_transformation * spacing
        if g_displacement is not None:
            g_u = g_u + g_displacement
        return ops.svf_exp_bwd(v, hist, maxabs, g_u.contiguous(), gather_radius_max=3), None, None, None


class SVF_3D(TransformationModule):
    """
    stationary velocity field: forward(v) -> (transformation in normalised [-1,1] units, displacement in voxels)
    (reference utils/transformation.py:51-76)
    """

    def __init__(self, dims, no_steps=12):
        super().__init__()
        self.identity_grid = nn.Parameter(init_identity_grid_3D(dims), requires_grad=False)
        self.no_steps = no_steps
        self.dims = tuple(dims)

    def forward(self, v):
        if v.dtype != torch.float32:
            raise NotImplementedError('SVF_3D: fp32 only')
        D, H, W = v.shape[2:]
        grid = self.identity_grid
        # the fp32 linspace tables the reference's identity grid is made of (utils/util.py:270-272)
        lin = [grid[0, 0, 0, :, 0].contiguous(), grid[0, 0, :, 0, 1].contiguous(), grid[0, :, 0, 0, 2].contiguous()]
        # transform_coordinates: channel i times 2 / (shape[2 + i] - 1)  (sic, utils/util.py:418-429)
        spacing = torch.tensor([2.0 / (D - 1), 2.0 / (H - 1), 2.0 / (W - 1)], device=v.device).view(1, 3, 1, 1, 1)
        return _ScalingAndSquaring.apply(v, self.no_steps, lin, spacing)


def cubic_B_spline_1D_value(x):
    """evaluate a 1D cubic B-spline (reference utils/transformation.py:79-92)"""
    t = abs(x)
    if t >= 2:
        return 0
    if t < 1:
        return 2.0 / 3.0 + (0.5 * t - 1.0) * t ** 2
    return -1.0 * ((t - 2.0) ** 3) / 6.0


def B_spline_1D_kernel(stride):
    """the 4 * stride - 1 taps of the cubic B-spline sampled every 1 / stride (reference utils/transformation.py:95-103)"""
    kernel = torch.ones(4 * stride - 1)
    radius = kernel.shape[0] // 2
    for i in range(kernel.shape[0]):
        kernel[i] = cubic_B_spline_1D_value((i - radius) / stride)
    return kernel


class _BSplineAxis(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, taps, dim, stride):
        ctx.taps, ctx.dim, ctx.stride, ctx.g = taps, dim, stride, x.shape[dim]
        return ops.bspline_axis(x.contiguous(), taps, dim, stride)

    @staticmethod
    def backward(ctx, g):
        return ops.bspline_axis(g.contiguous(), ctx.taps, ctx.dim, ctx.stride, adjoint=True, out_len=ctx.g), None, None, None


def conv1D(x, kernel, dim=-1, stride=1, dilation=1, padding=0, transpose=False):
    """
    convolve data with a 1-dimensional kernel along the specified dimension (reference utils/transformation.py:106-129).
    The CUDA path covers what the reference uses it for: the transposed convolution of the B-spline FFD, i.e.
    transpose=True, dilation=1, a kernel of 4 * stride - 1 taps and padding = (len(kernel) - 1) // 2.
    """
    if not transpose or dilation != 1 or kernel.numel() != 4 * stride - 1 or padding != (kernel.numel() - 1) // 2:
        raise NotImplementedError('conv1D: only the transposed B-spline convolution of Cubic_B_spline_FFD_3D is implemented')
    taps = tuple(float(k) for k in kernel.detach().cpu())
    return _BSplineAxis.apply(x.type(kernel.dtype), taps, dim % x.dim(), int(stride))


class _FFD(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v, taps, cps, dims):
        ctx.taps, ctx.cps, ctx.grid = taps, cps, tuple(v.shape[2:])
        return ops.ffd_fwd(v.contiguous(), taps, cps, dims)

    @staticmethod
    def backward(ctx, g):
        return ops.ffd_bwd(g.contiguous(), ctx.taps, ctx.cps, ctx.grid), None, None, None


class Cubic_B_spline_FFD_3D(TransformationModule):
    """
    dense velocity field of the cubic B-spline FFD model from control-point parameters: forward(v (N,3,gD,gH,gW)) ->
    (N,3,*dims); cps = control point spacing per axis (reference utils/transformation.py:132-152)
    """

    def __init__(self, dims, cps):
        super().__init__()
        self.dims = dims
        self.stride = cps
        self.kernels, self.padding = nn.ParameterList(), list()
        for s in self.stride:
            kernel = B_spline_1D_kernel(s)
            self.kernels.append(nn.Parameter(kernel, requires_grad=False))
            self.padding.append((len(kernel) - 1) // 2)
        self._taps = tuple(tuple(float(k) for k in kernel) for kernel in self.kernels)

    def forward(self, v):
        if v.dtype != torch.float32:
            raise NotImplementedError('Cubic_B_spline_FFD_3D: fp32 only')
        return _FFD.apply(v, self._taps, tuple(int(s) for s in self.stride), tuple(int(n) for n in self.dims))


class SVFFD_3D(TransformationModule):
    """stationary velocity field parametrised by cubic B-spline control points (reference utils/transformation.py:155-164)"""

    def __init__(self, dims, cps):
        super().__init__()
        self.cubic_B_spline_FFD = Cubic_B_spline_FFD_3D(dims, cps)
        self.SVF_3D = SVF_3D(dims)

    def forward(self, v):
        return self.SVF_3D(self.cubic_B_spline_FFD(v))
