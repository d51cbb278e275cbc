"""
Langevin proposal and Sobolev-gradient operators with the reference's names and signatures
(reference utils/functions.py: Sobolev_kernel_1D :24-49, SGLD :76-84, SobolevGrad :98-109).
"""
import math

import numpy as np
import torch

from .. import _lib


def Sobolev_kernel_1D(_s, _lambda):
    """
    approximate Sobolev smoothing kernel and its square root; host-side numpy like the reference.
    The kernel is the middle column of (I - lambda L)^-1 and of its principal square root, L the 1-D Laplacian of size
    2 s + 1, both normalised to unit sum: s=3, lambda=.5 -> [1, 4, 15, 56, 15, 4, 1] / 96
    """
    k = 2 * _s + 1
    lap = -2.0 * np.eye(k) + np.eye(k, k=1) + np.eye(k, k=-1)
    w, q = np.linalg.eigh(np.eye(k) - _lambda * lap)
    ok = np.abs(w) > 1e-10
    inv, inv_sqrt = np.zeros(k), np.zeros(k)
    inv[ok], inv_sqrt[ok] = 1.0 / w[ok], 1.0 / np.sqrt(w[ok])
    kernel = (q * inv) @ q[_s]
    kernel_sqrt = (q * inv_sqrt) @ q[_s]
    return kernel / kernel.sum(), kernel_sqrt / kernel_sqrt.sum()


def _taps_from_S(S):
    """the 1-D taps inside the reference's S dict of depthwise conv3d weights (trainer/trainer.py:572-583)"""
    return [float(x) for x in S['x'][0].reshape(-1).tolist()]


def langevin_sobolev(v, sigma, coef, taps, eps=None, seed=0, iteration=0, chain0=0):
    """out = S_x * S_y * S_z * pad_replicate(v + coef sigma eps); taps empty = no smoothing; coef 0 = no noise"""
    lib = _lib.load()
    _lib.require_cuda(v, sigma, eps)
    C, _, D, H, W = v.shape
    out, work = torch.empty_like(v), torch.empty_like(v) if len(taps) else None
    stride = 0 if sigma is None or sigma.shape[0] == 1 else 3 * D * H * W
    _lib.check(lib.irs_langevin_sobolev(_lib.ptr(v), _lib.ptr(sigma), stride, float(coef), _lib.ptr(eps), int(seed),
                                        int(iteration), int(chain0), _lib.host_floats(taps) if len(taps) else None,
                                        len(taps), _lib.ptr(work), _lib.ptr(out), C, D, H, W, _lib.stream()))
    return out


class SGLD(torch.autograd.Function):
    """forward v + sqrt(2 tau) sigma eps, eps ~ N(0,1) (Philox on the device); backward sigma^2 g"""
    seed, calls = 123, 0

    @staticmethod
    def forward(ctx, v_curr_state, sigma, tau):
        ctx.sigma = sigma
        SGLD.calls += 1
        sg = sigma.contiguous() if sigma.shape == v_curr_state.shape else sigma.expand_as(v_curr_state).contiguous()
        return langevin_sobolev(v_curr_state.detach().contiguous(), sg, math.sqrt(2.0 * tau), [], seed=SGLD.seed,
                                iteration=SGLD.calls)

    @staticmethod
    def backward(ctx, grad_output):
        return ctx.sigma ** 2 * grad_output, None, None


class SobolevGrad(torch.autograd.Function):
    """forward: separable smoothing with replicate padding; backward: identity (the reference's quirk, :107-109)"""

    @staticmethod
    def forward(ctx, input, S, padding):
        return langevin_sobolev(input.detach().contiguous(), None, 0.0, _taps_from_S(S))

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output, None, None
