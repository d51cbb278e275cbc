"""
Transformation models with the reference's interface (reference utils/transformation.py).  SVF_3D integrates a
stationary velocity field by scaling and squaring on the CUDA kernels of libirsgmcmc.so; its backward pass is the
gather-form adjoint (no scatter atomics).
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
