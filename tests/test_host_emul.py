"""
CPU checks of the library's device arithmetic: the __host__ __device__ per-voxel functions that the CUDA kernels call
(csrc/irs_common.cuh, irs_bodies.cuh, irs_hyper.cuh) run in plain loops (tests/host_emul) against the oracle.
"""
import ctypes
import math

import numpy as np
import pytest
import torch

from oracle import sgld_oracle as O
from tests.util import grad_ok, rel, smooth_field


@pytest.fixture(scope='module')
def emul(built):
    return ctypes.CDLL(built['emul'])


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_philox_known_answers(emul):
    """Random123 known-answer vectors for Philox4x32-10"""
    out = (ctypes.c_uint * 4)()
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, expect in kat:
        emul.emul_philox(*[ctypes.c_uint(x) for x in ctr], *[ctypes.c_uint(x) for x in key], out)
        assert tuple(out) == expect


def test_noise_statistics(emul):
    n = 200000
    e = np.zeros((n, 3), np.float32)
    emul.emul_normal3(ctypes.c_ulonglong(123), n, 0, ctypes.c_ulonglong(5), P(e))
    x = e.astype(np.float64).ravel()
    N = x.size
    assert abs(x.mean()) < 5 / math.sqrt(N) and abs(x.var() - 1) < 5 * math.sqrt(2 / N)
    assert abs((x ** 4).mean() - 3) < 5 * math.sqrt(96 / N)
    assert abs(np.mean(e[:, 0] * e[:, 1])) < 5 / math.sqrt(n) and abs(np.mean(e[:-1, 2] * e[1:, 2])) < 5 / math.sqrt(n)
    u = np.zeros((n, 3), np.float32)
    emul.emul_uniform3(ctypes.c_ulonglong(123), n, 1, ctypes.c_ulonglong(5), P(u))
    assert u.min() >= 0.0 and u.max() < 1.0 and abs(u.mean() - 0.5) < 5 / math.sqrt(12 * 3 * n)


@pytest.mark.parametrize('amp', [0.5, 3.0, 8.0])
def test_svf_forward_and_adjoint(emul, amp):
    n, C, steps = 12, 2, 12
    v = smooth_field((C, 3, n, n, n), amp, 0)
    vn = np.ascontiguousarray(v.numpy())
    hist, maxabs = np.zeros((steps, C, 3, n, n, n), np.float32), np.zeros(steps, np.float32)
    emul.emul_svf_fwd(P(vn), P(hist), P(maxabs), steps, C, n, n, n)
    v64 = v.double().requires_grad_(True)
    _, d64 = O.svf_exp_aten(v64, steps, exact_grid=True)
    assert rel(hist[-1], d64) < 1e-6
    G = torch.randn(C, 3, n, n, n, generator=torch.Generator().manual_seed(1))
    g64, = torch.autograd.grad((d64 * G.double()).sum(), v64)
    v32 = v.clone().requires_grad_(True)
    _, d32 = O.svf_exp_aten(v32, steps)
    g32, = torch.autograd.grad((d32 * G).sum(), v32)
    for mode in (0, 1):  # gather with radius floor(maxabs)+1 / scatter
        gv = np.zeros((C, 3, n, n, n), np.float32)
        emul.emul_svf_bwd(P(vn), P(hist), P(maxabs), P(np.ascontiguousarray(G.numpy())), P(gv), steps, mode, C, n, n, n)
        assert grad_ok(gv, g32, g64, f'svf adjoint amp={amp} mode={mode}')


def test_warps(emul):
    n, C = 14, 2
    torch.manual_seed(3)
    im = torch.rand(1, 1, n, n, n)
    T, _ = O.svf_exp_aten(smooth_field((C, 3, n, n, n), 2.0, 2), 6)
    T = T.contiguous()
    T[0, :, 0, 0, :] = -1.2
    T[1, :, 1, :, 0] = 1.0
    out = np.zeros((C, 1, n, n, n), np.float32)
    emul.emul_warp_fwd(P(im.numpy()), P(T.numpy()), P(out), C, n, n, n)
    assert rel(out, O.warp_aten(im.double().expand(C, -1, -1, -1, -1), T.double())) < 1e-6
    g_out = torch.randn(C, 1, n, n, n)
    T64 = T.double().requires_grad_(True)
    (O.warp_aten(im.double().expand(C, -1, -1, -1, -1), T64) * g_out.double()).sum().backward()
    gT = np.zeros((C, 3, n, n, n), np.float32)
    emul.emul_warp_bwd_grid(P(im.numpy()), P(T.numpy()), P(np.ascontiguousarray(g_out.numpy())), P(gT), C, n, n, n)
    assert rel(gT, T64.grad) < 1e-5
    seg = (torch.rand(1, 1, n, n, n) * 60).to(torch.int16)
    k = torch.arange(n, dtype=torch.float32)
    T[0, 0, 0, 0, :] = 2.0 * (k + 0.5) / (n - 1) - 1.0
    out_seg = np.zeros((C, 1, n, n, n), np.int16)
    emul.emul_warp_nearest_i16(P(seg.numpy()), P(T.numpy()), P(out_seg), C, n, n, n)
    assert np.array_equal(out_seg, O.warp_nearest_aten(seg.expand(C, -1, -1, -1, -1), T).numpy())


def test_regulariser_energy_and_gradient(emul):
    D, H, W = 9, 7, 11
    v = torch.randn(1, 3, D, H, W)
    grad = np.zeros((3, D, H, W), np.float32)
    emul.emul_reg_energy.restype = ctypes.c_double
    e = emul.emul_reg_energy(P(np.ascontiguousarray(v.numpy())), P(grad), D, H, W)
    v64 = v.double().requires_grad_(True)
    y = O.reg_energy(v64)
    y.sum().backward()
    assert abs(e - float(y)) < 1e-5 * float(y) and rel(grad, v64.grad[0]) < 1e-6


@pytest.mark.parametrize('K', [1, 4])
def test_mixture_step(emul, K):
    """statistics + VD factor + Adam step on the mixture (what the last block of gmm_stats_kernel does) vs the oracle"""
    n = 16
    torch.manual_seed(4)
    z = smooth_field((1, 1, n, n, n), 2.0, 4, passes=1) + 0.3 * torch.randn(1, 1, n, n, n)
    mask = torch.rand(1, 1, n, n, n) > 0.3
    cfg = O.Config(K=K)
    st = O.State(cfg, torch.zeros(1, 3, n, n, n), torch.ones(1, 3, n, n, n), (n, n, n))
    if K > 1:
        st.init_gmm(0.7)
    else:
        st.log_std.fill_(0.2)
    hyper = np.zeros(64, np.float64)
    hyper[1:1 + K], hyper[9:9 + K] = st.log_std.numpy(), st.logits.numpy()
    cfg_d = np.array([0.2, 0.2, 1e-3, 0.9, 0.999, 1e-8, 0.0, 2.3, 0.5, float(mask.sum())], np.float64)
    table, dz = np.zeros(16, np.float32), np.zeros(n ** 3, np.float32)
    emul.emul_gmm_stats_step.restype = ctypes.c_double
    for it in range(3):
        alpha_ref = O.vd_factor(O.vd_residual(z, mask, st.log_std, st.logits), mask)
        O.gmm_step(st, z[mask], alpha_ref)
        alpha = emul.emul_gmm_stats_step(P(np.ascontiguousarray(z.numpy())), P(mask.numpy().view(np.uint8)), P(hyper), K, 1,
                                         P(cfg_d), P(table), None, P(dz), n, n, n)
        assert abs(alpha - float(alpha_ref)) < 2e-5 * float(alpha_ref)
        assert np.abs(hyper[1:1 + K] - st.log_std.numpy()).max() < 2e-5
        assert np.abs(hyper[9:9 + K] - st.logits.numpy()).max() < 5e-5
        zz = z.clone().requires_grad_(True)
        (O.gmm_nll(zz[mask], st.log_std, st.logits) * alpha_ref).backward()
        assert rel(dz, zz.grad.flatten()) < 1e-4


def test_vd_factor_non_positive_correlations(emul):
    """reference utils/util.py:481-485: alpha = sqrt(prod_a clamp(-(2/pi) log(corr_a), max=1)); torch.clamp propagates the
    NaN of a NEGATIVE lag-1 correlation, a correlation of exactly zero gives +inf -> 1 (fminf would hide the NaN)"""
    emul.emul_vd_alpha.restype = ctypes.c_double
    n_mask = 1000.0
    for corr in ([0.5, 0.4, 0.3], [0.5, -0.1, 0.3], [0.0, 0.4, 0.3], [0.9, 0.95, 0.99], [-0.2, -0.1, 0.0]):
        sums = np.zeros(32, np.float64)
        sums[1] = 2.0 * n_mask                      # sum r^2 -> var = 2
        sums[2:5] = np.array(corr) * 2.0 * n_mask   # lag sums / n_mask / var = corr
        got = emul.emul_vd_alpha(P(sums), ctypes.c_double(n_mask))
        c = torch.tensor(corr, dtype=torch.float32)
        want = float(torch.sqrt(torch.prod(torch.clamp(-2.0 / math.pi * torch.log(c), max=1.0))))
        assert (math.isnan(got) and math.isnan(want)) or abs(got - want) <= 1e-6 * max(abs(want), 1e-30), (corr, got, want)


@pytest.mark.parametrize('reg,learnable', [('lognormal', True), ('lognormal', False), ('l2', True), ('l2', False)])
def test_regulariser_hyper_step(emul, reg, learnable):
    n, C, w_reg = 32, 3, 1.6
    dof = 3.0 * n ** 3
    y = torch.tensor([4.1e4, 5.3e4, 6.0e4], dtype=torch.float64)
    cfg = O.Config(reg=reg, w_reg=w_reg, reg_learnable=learnable)
    st = O.State(cfg, torch.zeros(C, 3, 4, 4, 4), torch.ones(C, 3, 4, 4, 4), (n, n, n), torch.float64)
    hyper = np.zeros(64, np.float64)
    if reg == 'lognormal':
        hyper[50], hyper[51] = float(st.loc), float(st.log_scale)
    else:
        hyper[50] = float(st.log_w_reg)
    shape = 0.5 * dof
    cfg_d = np.array([1 if reg == 'lognormal' else 0, int(learnable), 0.01, 0.01, 1e-3, 0.9, 0.999, 1e-8, 2.8, 5.0, w_reg,
                      dof, shape, 1.0 / shape], np.float64)
    for it in range(3):
        yy = y.clone().requires_grad_(True)
        per_chain, total, leaves = O.reg_term_fn(st, yy)
        grads = torch.autograd.grad(total, (yy,) + (leaves if learnable else ()))
        stats = np.zeros((C, 8), np.float64)
        stats[:, 3] = y.numpy()
        emul.emul_reg_hyper_step(P(hyper), P(cfg_d), C, P(stats))
        assert rel(stats[:, 2], per_chain) < 1e-7
        assert rel(stats[:, 5], grads[0]) < 1e-6     # the coefficient multiplying dE/dv in the field gradient
        if learnable:
            st.adam_reg.step([g.to(p.dtype) for g, p in zip(grads[1:], st.adam_reg.params)])
            ref = [float(st.loc), float(st.log_scale)] if reg == 'lognormal' else [float(st.log_w_reg)]
            assert np.allclose(hyper[50:50 + len(ref)], ref, rtol=1e-6, atol=0)
        y = y * 1.01
