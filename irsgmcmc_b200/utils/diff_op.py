"""
Differential operators with the reference's interface (reference utils/diff_op.py): forward differences with the last
difference replicated, output (N, 3, D, H, W, 3) with [n, j, ..., i] = d v_i / d x_j.
"""
from abc import ABC

import torch
from torch import nn

from .. import ops


def _all_subclasses(cls):
    for sub in cls.__subclasses__():
        yield sub
        yield from _all_subclasses(sub)


class DifferentialOperator(nn.Module, ABC):
    """identity by default; `from_string('GradientOperator')` builds a subclass by name"""

    @staticmethod
    def from_string(s, *args, **kwargs):
        for cls in _all_subclasses(DifferentialOperator):
            if cls.__name__ in s:
                return cls(*args, **kwargs)
        raise ValueError('Unknown differential operator: {}'.format(s))

    def forward(self, input):
        return input


class _ForwardDifferences(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v, transformation):
        ctx.transformation = transformation
        return ops.diff_fwd(v.contiguous(), transformation)

    @staticmethod
    def backward(ctx, g_nabla):
        return ops.diff_bwd(g_nabla.contiguous(), ctx.transformation), None


class GradientOperator(DifferentialOperator):
    def __init__(self):
        super().__init__()
        self.pixel_spacing = None

    def forward(self, v, transformation=False):
        if transformation and self.pixel_spacing is None:
            self.pixel_spacing = [2.0 / (n - 1) for n in v.shape[2:]]
        return _ForwardDifferences.apply(v, bool(transformation))
