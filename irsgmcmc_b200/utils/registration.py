"""
RegistrationModule with the reference's interface (reference utils/registration.py:5-41) on the CUDA warps of
libirsgmcmc.so: trilinear for float images (gradient w.r.t. the transformation as a pure gather), nearest neighbour
for bool masks / int16 segmentations, bit-exact with the reference.  border padding, align_corners=True.
"""
import torch
from torch import nn

from .. import ops


class _Warp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, im, transformation):
        im_c, T_c = im.contiguous(), transformation.contiguous()
        ctx.save_for_backward(im_c, T_c)
        return ops.warp3d(im_c, T_c)

    @staticmethod
    def backward(ctx, grad_output):
        im, T = ctx.saved_tensors
        if ctx.needs_input_grad[0]:
            # the reference never differentiates w.r.t. the image (moving['im'] is data, trainer/trainer.py:296-300)
            raise NotImplementedError('gradient w.r.t. the warped image is not part of the SGLD step')
        return None, ops.warp3d_bwd_grid(im, T, grad_output.contiguous())


class RegistrationModule(nn.Module):
    """module for warping images, masks and segmentations"""

    def __init__(self):
        super().__init__()
        self.im_types = ['torch.cuda.FloatTensor']
        self.mask_types = ['torch.cuda.BoolTensor']
        self.seg_types = ['torch.cuda.ShortTensor']

    def forward(self, im_or_seg_moving, transformation):
        kind = im_or_seg_moving.type()
        if kind in self.mask_types or kind in self.seg_types:
            return ops.warp3d_nearest(im_or_seg_moving.contiguous(), transformation.detach().contiguous())
        if kind in self.im_types:
            return _Warp.apply(im_or_seg_moving, transformation)
        # same error as the reference for an unsupported dtype -- and for CPU tensors: there is no CPU path
        raise NotImplementedError

    def is_im(self, input):
        return input.type() in self.im_types

    def is_mask(self, input):
        return input.type() in self.mask_types

    def is_seg(self, input):
        return input.type() in self.seg_types
