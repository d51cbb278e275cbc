"""
The oracle (oracle/sgld_oracle.py) against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py), on the CPU; and the CUDA path against the same vectors (gpu-marked).
"""
import math
import os

import numpy as np
import pytest
import torch

from oracle import sgld_oracle as O
from tests.util import grad_ok, rel

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(name):
    return {k: torch.from_numpy(np.asarray(v)) for k, v in np.load(os.path.join(GOLD, name)).items()}


@pytest.fixture(scope='module')
def ops_gold():
    return load('ops.npz')


# ---------------------------------------------------------------------------------------------------------------------
# oracle vs reference golden vectors (CPU)
# ---------------------------------------------------------------------------------------------------------------------
def test_oracle_sobolev_kernel(ops_gold):
    for s in (1, 2, 3):
        assert np.abs(O.sobolev_taps(s, 0.5) - ops_gold[f'sobolev_s{s}'].numpy()).max() < 1e-14
    assert np.abs(O.sobolev_taps(3, 0.5) * 96 - np.array([1, 4, 15, 56, 15, 4, 1])).max() < 1e-12


def test_oracle_smoothing(ops_gold):
    taps = O.sobolev_taps(3, 0.5).astype(np.float32)
    assert rel(O.sobolev_smooth(ops_gold['smooth_in'], taps), ops_gold['smooth_out']) < 1e-6


def test_oracle_svf(ops_gold):
    v, G = ops_gold['svf_v'], ops_gold['svf_G']
    v32 = v.clone().requires_grad_(True)
    T, disp = O.svf_exp_aten(v32)
    assert rel(T, ops_gold['svf_T']) < 1e-6 and rel(disp, ops_gold['svf_disp']) < 2e-6
    g, = torch.autograd.grad((disp * G).sum(), v32)
    assert rel(g, ops_gold['svf_grad']) < 1e-6
    # fp64, exact identity grid; and the independent voxel-unit restatement of the same function
    v64 = v.double().requires_grad_(True)
    _, d64 = O.svf_exp_aten(v64, exact_grid=True)
    g64, = torch.autograd.grad((d64 * G.double()).sum(), v64)
    assert rel(d64, ops_gold['svf_disp_f64']) < 1e-6 and rel(g64, ops_gold['svf_grad_f64']) < 1e-6
    v64b = v.double().requires_grad_(True)
    dv = O.svf_exp_voxel(v64b)
    gv, = torch.autograd.grad((dv * G.double()).sum(), v64b)
    assert rel(dv, d64) < 1e-12 and rel(gv, g64) < 1e-10


def test_oracle_warps(ops_gold):
    T, im, seg, mask = ops_gold['warp_T'], ops_gold['warp_im'], ops_gold['warp_seg'], ops_gold['warp_mask']
    assert torch.equal(O.warp_aten(im, T), ops_gold['warp_im_out'])
    assert torch.equal(O.warp_nearest_aten(seg, T), ops_gold['warp_seg_out'])
    assert torch.equal(O.warp_nearest(seg, T), ops_gold['warp_seg_out'])
    assert torch.equal(O.warp_nearest(mask, T), ops_gold['warp_mask_out'])
    # hand-written trilinear restatement vs ATen
    px, py, pz = O.unnormalise(T.double())
    assert rel(O.trilinear_sample_voxel(im.double(), px, py, pz), ops_gold['warp_im_out']) < 1e-6


def test_oracle_c_nearest(ops_gold, built):
    import ctypes
    lib = ctypes.CDLL(built['oracle_c'])
    T, seg, mask = ops_gold['warp_T'].contiguous(), ops_gold['warp_seg'].contiguous(), ops_gold['warp_mask'].contiguous()
    C, _, D, H, W = T.shape
    out = torch.empty_like(seg)
    lib.oracle_warp_nearest_i16(ctypes.c_void_p(seg.data_ptr()), ctypes.c_longlong(D * H * W), ctypes.c_void_p(T.data_ptr()),
                                ctypes.c_void_p(out.data_ptr()), C, D, H, W)
    assert torch.equal(out, ops_gold['warp_seg_out'])
    m8, out8 = mask.view(torch.uint8), torch.empty_like(mask.view(torch.uint8))
    lib.oracle_warp_nearest_u8(ctypes.c_void_p(m8.data_ptr()), ctypes.c_longlong(D * H * W), ctypes.c_void_p(T.data_ptr()),
                               ctypes.c_void_p(out8.data_ptr()), C, D, H, W)
    assert torch.equal(out8.view(torch.bool), ops_gold['warp_mask_out'])


def test_oracle_diff_op_and_det_j(ops_gold):
    v = ops_gold['smooth_in']
    assert torch.equal(O.forward_differences(v), ops_gold['nabla_v'])
    nab = O.forward_differences(ops_gold['svf_T'], transformation=True)
    assert rel(nab, ops_gold['nabla_T']) < 1e-6
    assert rel(O.det_jacobian(nab), ops_gold['det_J']) < 1e-5


def test_oracle_evaluation_functions():
    """per-sample evaluation: Dice, folded-voxel count / log det J on a transformation that folds, norms --
    vectors from the reference's calc_DSC_GPU, calc_no_non_diffeomorphic_voxels, calc_norm (make_golden.py eval)"""
    g = load('eval.npz')
    labels = [int(x) for x in g['labels']]
    assert torch.equal(O.dice_scores(g['seg_fixed'], g['seg_moving_warped'], labels), g['DSC'])
    counts, log_det = O.no_non_diffeomorphic_voxels(g['T'])
    assert counts.tolist() == g['no_non_diffeomorphic_voxels'].tolist() and counts[2] > 0
    assert torch.equal(torch.isnan(log_det), torch.isnan(g['log_det_J']))
    ok = ~torch.isnan(log_det)
    assert rel(log_det[ok].exp(), g['log_det_J'][ok].exp()) < 1e-6
    assert rel(O.field_norm(g['disp']), g['disp_norm']) < 1e-7


def test_oracle_data_term(ops_gold):
    for s in (1, 2):
        z = O.lcc_map(ops_gold[f'lcc_s{s}_F'], ops_gold[f'lcc_s{s}_M'], s)
        assert rel(z, ops_gold[f'lcc_s{s}_z']) < 2e-5     # 125-tap dense conv vs separable sums, amplified by 1/sigma
        z64 = O.lcc_map(ops_gold[f'lcc_s{s}_F'].double(), ops_gold[f'lcc_s{s}_M'].double(), s)
        assert rel(ops_gold[f'lcc_s{s}_z'], z64) < 2e-5
        zr, mask = ops_gold[f'lcc_s{s}_z'], ops_gold[f'lcc_s{s}_mask']
        ls, lg = ops_gold[f'gmm_s{s}_log_std'], ops_gold[f'gmm_s{s}_logits']
        assert rel(O.gmm_log_pdf(zr[mask], ls, lg), ops_gold[f'gmm_s{s}_log_pdf']) < 1e-6
        r = O.vd_residual(zr, mask, ls, lg)
        assert rel(r, ops_gold[f'vd_s{s}_rescaled']) < 1e-5
        assert abs(float(O.vd_factor(r, mask)) - float(ops_gold[f'vd_s{s}_alpha'])) < 1e-5


def test_oracle_regulariser(ops_gold):
    v = ops_gold['smooth_in']
    n = v.shape[-1]
    y = O.reg_energy(v)
    assert rel(y.log(), ops_gold['reg_l2_log_y']) < 1e-6
    dof = 3.0 * n ** 3
    l2 = 0.5 * 1.4 * y - 0.5 * dof * math.log(1.4)
    assert rel(l2, ops_gold['reg_l2_loss']) < 1e-6
    loc, log_scale = O.lognormal_init(1.4, dof)
    ly = y.log()
    ln = ly + log_scale + 0.5 * ((ly - loc) / math.exp(log_scale)) ** 2 + (0.5 * dof - 1) * ly
    assert rel(ln, ops_gold['reg_lognormal_loss']) < 1e-6
    loc128, ls128 = O.lognormal_init(1.6, 3.0 * 128 ** 3)
    assert abs(loc128 - float(ops_gold['lognormal_init_128'][0])) < 1e-9
    assert abs(ls128 - float(ops_gold['lognormal_init_128'][1])) < 1e-9


def test_oracle_posterior_statistics(ops_gold):
    mean, std = O.posterior_statistics(ops_gold['post_samples'])
    assert rel(mean, ops_gold['post_mean']) < 1e-6 and rel(std, ops_gold['post_std']) < 1e-6
    parts = []
    for chunk in ops_gold['post_samples'].double().split(3):
        parts.append((chunk.shape[0], chunk.mean(0), ((chunk - chunk.mean(0)) ** 2).sum(0)))
    n, m, m2 = O.welford_merge(parts)
    assert n == 9 and rel(m, ops_gold['post_mean']) < 1e-6 and rel((m2 / (n - 1)).sqrt(), ops_gold['post_std']) < 1e-6


def _oracle_state(g, reg, w_reg, dtype, it):
    n, C = int(g['n']), int(g['C'])
    cfg = O.Config(reg=reg, w_reg=w_reg, exact_grid=False)
    sfx = '' if dtype == torch.float32 else '_f64'
    st = O.State(cfg, g[f'it{it}{sfx}_v_before'].to(dtype), g['sigma'].to(dtype).expand(C, -1, -1, -1, -1), (n, n, n), dtype)
    return st


@pytest.mark.parametrize('tag,reg,w_reg', [('lcc_lognormal', 'lognormal', 1.6), ('lcc_l2', 'l2', 1.4)])
def test_oracle_transition_vs_reference_golden(tag, reg, w_reg):
    """first transition from the reference's initial state, same injected noise: every output of _SGLD_transition"""
    from irsgmcmc_b200.data_loader.synthetic import make_pair
    g = load(f'transition_{tag}.npz')
    n = int(g['n'])
    torch.manual_seed(123)
    fixed, moving, _ = make_pair(n)
    for dtype, sfx, tol_f, tol_g in ((torch.float32, '', 1e-5, 2e-3), (torch.float64, '_f64', 1e-6, 1e-5)):
        st = _oracle_state(g, reg, w_reg, dtype, 0)
        st.init_gmm(0.7)
        cast = lambda d: {k: (v.to(dtype) if v.dtype == torch.float32 else v) for k, v in d.items()}
        lt, out, aux, grad_v = O.sgld_transition(st, cast(fixed), cast(moving), g['it0_eps'].to(dtype),
                                                 g['it0_jitter'].to(dtype))
        p = f'it0{sfx}_'
        for key in ('curr_state', 'transformation', 'displacement', 'im_moving_warped'):
            assert rel(out[key], g[p + key]) < tol_f, (key, dtype)
        assert rel(aux['residuals'], g[p + 'residuals']) < 10 * tol_f
        assert rel(torch.stack(aux['alpha']), g[p + 'alpha']) < 1e-4
        assert rel(torch.stack(lt['data']), g[p + 'data']) < 1e-4
        assert rel(torch.stack(lt['reg']), g[p + 'reg']) < 1e-6
        assert rel(torch.stack(aux['reg_energy']), g[p + 'reg_energy']) < 1e-6
        assert rel(grad_v, g[p + 'grad_v']) < tol_g, dtype
        assert rel(st.log_std, g[p + 'log_std']) < 1e-5 and rel(st.logits, g[p + 'logits']) < 1e-4
        regp = torch.stack((st.loc, st.log_scale)) if reg == 'lognormal' else st.log_w_reg.view(1)
        assert rel(regp, g[p + 'reg_params']) < 1e-7


def test_oracle_svffd_transition_vs_reference_golden():
    """SVFFD_3D as the transformation module (configs/experiment5): chain state on the control grid"""
    from irsgmcmc_b200.data_loader.synthetic import make_pair
    g = load('transition_svffd_lognormal.npz')
    n, C, cps = int(g['n']), int(g['C']), tuple(int(x) for x in g['cps'])
    assert O.control_grid_size((n,) * 3, cps) == tuple(int(x) for x in g['grid'])
    torch.manual_seed(123)
    fixed, moving, _ = make_pair(n)
    for dtype, sfx, tol_f, tol_g in ((torch.float32, '', 1e-5, 2e-3), (torch.float64, '_f64', 1e-6, 1e-5)):
        cfg = O.Config(reg='lognormal', w_reg=1.6, cps=cps)
        st = O.State(cfg, g[f'it0{sfx}_v_before'].to(dtype), g['sigma'].to(dtype).expand(C, -1, -1, -1, -1), (n, n, n), dtype)
        st.init_gmm(0.7)
        cast = lambda d: {k: (v.to(dtype) if v.dtype == torch.float32 else v) for k, v in d.items()}
        lt, out, aux, grad_v = O.sgld_transition(st, cast(fixed), cast(moving), g['it0_eps'].to(dtype),
                                                 g['it0_jitter'].to(dtype))
        p = f'it0{sfx}_'
        assert out['curr_state'].shape == g[p + 'curr_state'].shape == (C, 3, *g['grid'].tolist())
        for key in ('curr_state', 'transformation', 'displacement', 'im_moving_warped'):
            assert rel(out[key], g[p + key]) < tol_f, (key, dtype)
        assert rel(O.ffd_dense(out['curr_state'], (n,) * 3, cps), g[p + 'velocity']) < tol_f
        assert rel(torch.stack(aux['alpha']), g[p + 'alpha']) < 1e-4
        assert rel(torch.stack(lt['data']), g[p + 'data']) < 1e-4
        assert rel(torch.stack(lt['reg']), g[p + 'reg']) < 1e-6
        assert rel(torch.stack(aux['reg_energy']), g[p + 'reg_energy']) < 1e-6
        assert rel(grad_v, g[p + 'grad_v']) < tol_g, dtype
        assert rel(torch.stack((st.loc, st.log_scale)), g[p + 'reg_params']) < 1e-7


# ---------------------------------------------------------------------------------------------------------------------
# CUDA path vs reference golden vectors (GPU)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_cuda_ops_vs_reference_golden(ops_gold, built):
    from irsgmcmc_b200 import ops
    from irsgmcmc_b200.utils.functions import langevin_sobolev
    dev = 'cuda:0'
    taps = list(O.sobolev_taps(3, 0.5).astype(np.float32))
    assert rel(langevin_sobolev(ops_gold['smooth_in'].to(dev), None, 0.0, taps), ops_gold['smooth_out']) < 1e-6
    v, G = ops_gold['svf_v'].to(dev), ops_gold['svf_G'].to(dev)
    hist, maxabs = ops.svf_exp_fwd(v)
    assert rel(hist[-1], ops_gold['svf_disp_f64']) < 1e-5
    n = v.shape[-1]
    lin = [torch.linspace(-1, 1, steps=n).to(dev)] * 3
    assert rel(ops.svf_outputs(hist[-1], lin), ops_gold['svf_T']) < 1e-6
    assert grad_ok(ops.svf_exp_bwd(v, hist, maxabs, G), ops_gold['svf_grad'], ops_gold['svf_grad_f64'], 'golden svf adjoint')
    T = ops_gold['warp_T'].to(dev)
    assert rel(ops.warp3d(ops_gold['warp_im'].to(dev), T), ops_gold['warp_im_out']) < 1e-5
    assert torch.equal(ops.warp3d_nearest(ops_gold['warp_seg'].to(dev), T).cpu(), ops_gold['warp_seg_out'])
    assert torch.equal(ops.warp3d_nearest(ops_gold['warp_mask'].to(dev), T).cpu(), ops_gold['warp_mask_out'])
    assert rel(ops.diff_fwd(ops_gold['smooth_in'].to(dev)), ops_gold['nabla_v']) < 1e-6
    assert rel(ops.diff_fwd(ops_gold['svf_T'].to(dev), True), ops_gold['nabla_T']) < 1e-6
    for s in (1, 2):
        zF, _, _ = ops.lcc_normalise(ops_gold[f'lcc_s{s}_F'].to(dev), s)
        zM, _, _ = ops.lcc_normalise(ops_gold[f'lcc_s{s}_M'].to(dev), s)
        assert rel(zF - zM, ops_gold[f'lcc_s{s}_z']) < 2e-5
        zr, mask = ops_gold[f'lcc_s{s}_z'].to(dev), ops_gold[f'lcc_s{s}_mask'].to(dev)
        ls, lg = ops_gold[f'gmm_s{s}_log_std'], ops_gold[f'gmm_s{s}_logits']
        logp, _, _ = ops.gmm_log_pdf(zr[mask].contiguous(), ls, lg)
        assert rel(logp, ops_gold[f'gmm_s{s}_log_pdf']) < 1e-5
        assert abs(float(ops.vd_factor(zr, mask, ls, lg)) - float(ops_gold[f'vd_s{s}_alpha'])) < 1e-5
    assert rel(ops.reg_energy(ops_gold['smooth_in'].to(dev)).log(), ops_gold['reg_l2_log_y']) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize('tag,reg,w_reg', [('lcc_lognormal', 'RegLoss_LogNormal', 1.6), ('lcc_l2', 'RegLoss_L2', 1.4)])
def test_cuda_transition_vs_reference_golden(tag, reg, w_reg, built):
    """the fused CUDA step on the reference's own inputs and injected noise, both golden iterations"""
    from irsgmcmc_b200.sampler import SGLDSampler, SGLDConfig
    from irsgmcmc_b200.data_loader.synthetic import make_pair
    g = load(f'transition_{tag}.npz')
    n, C = int(g['n']), int(g['C'])
    torch.manual_seed(123)
    fixed, moving, _ = make_pair(n)
    s = SGLDSampler(fixed, moving, C, SGLDConfig(reg_loss=reg, w_reg=w_reg), device='cuda:0')
    s.set_state(g['v0'], g['sigma'])
    s.init_gmm(sigma_hat=0.7)
    for it in range(2):
        if it == 1:
            # iteration 1 starts from the reference's own state after iteration 0 (chain state, mixture and regulariser
            # parameters), so that it is held to the same 1e-5 as iteration 0 instead of inheriting the kink noise of the
            # first gradient (the Adam moments stay this run's own: they agree to rounding)
            from irsgmcmc_b200 import _lib as L
            K = s.cfg.no_components
            s.v.copy_(torch.as_tensor(g['it1_v_before']))
            s.hyper[L.HYPER_LOG_STD:L.HYPER_LOG_STD + K] = torch.as_tensor(g['it0_log_std']).double().to(s.device)
            s.hyper[L.HYPER_LOGITS:L.HYPER_LOGITS + K] = torch.as_tensor(g['it0_logits']).double().to(s.device)
            rp = torch.as_tensor(g['it0_reg_params']).double().reshape(-1)
            s.hyper[L.HYPER_REG_P:L.HYPER_REG_P + rp.numel()] = rp.to(s.device)
        s.set_noise(g[f'it{it}_eps'], g[f'it{it}_jitter'])
        s.step(1, use_graph=False)
        torch.cuda.synchronize()
        tol = 1e-5
        p, p64 = f'it{it}_', f'it{it}_f64_'
        out = s.output()
        for key in ('curr_state', 'transformation', 'displacement', 'im_moving_warped'):
            assert rel(out[key], g[p + key]) < tol, (it, key)
        terms = s.loss_terms()
        assert rel(terms['reg_energy'], g[p + 'reg_energy']) < tol and rel(terms['reg'], g[p + 'reg']) < tol
        assert rel(terms['alpha'], g[p + 'alpha']) < max(tol, 1e-4) and rel(terms['data'], g[p + 'data']) < max(tol, 1e-4)
        if it == 0:
            assert grad_ok(s.grad_v, g[p + 'grad_v'], g[p64 + 'grad_v'], 'golden grad_v')
            assert rel(s.gmm_parameters()[0], g[p + 'log_std']) < 1e-5


@pytest.mark.gpu
def test_cuda_svffd_transition_vs_reference_golden(built):
    """the fused CUDA step with SVFFD_3D as the transformation model on the reference's inputs and injected noise"""
    from irsgmcmc_b200.sampler import SGLDSampler, SGLDConfig
    from irsgmcmc_b200.data_loader.synthetic import make_pair
    g = load('transition_svffd_lognormal.npz')
    n, C, cps = int(g['n']), int(g['C']), tuple(int(x) for x in g['cps'])
    torch.manual_seed(123)
    fixed, moving, _ = make_pair(n)
    s = SGLDSampler(fixed, moving, C, SGLDConfig(transformation='SVFFD_3D', cps=cps), device='cuda:0')
    assert s.v.shape == (C, 3, *g['grid'].tolist())
    # + 6 FFD axis passes + the stand-alone energy kernel, + 2 because the control grid (W = 7) cannot take the fused Langevin / x / y kernel
    assert s.launches_per_step() == SGLDSampler(fixed, moving, C, SGLDConfig(), device='cuda:0').launches_per_step() + 7 + 2
    s.set_state(g['v0'], g['sigma'])
    s.init_gmm(sigma_hat=0.7)
    for it in range(2):
        if it == 1:
            # iteration 1 starts from the reference's own state after iteration 0 (chain state, mixture and regulariser
            # parameters), so that it is held to the same 1e-5 as iteration 0 instead of inheriting the kink noise of the
            # first gradient (the Adam moments stay this run's own: they agree to rounding)
            from irsgmcmc_b200 import _lib as L
            K = s.cfg.no_components
            s.v.copy_(torch.as_tensor(g['it1_v_before']))
            s.hyper[L.HYPER_LOG_STD:L.HYPER_LOG_STD + K] = torch.as_tensor(g['it0_log_std']).double().to(s.device)
            s.hyper[L.HYPER_LOGITS:L.HYPER_LOGITS + K] = torch.as_tensor(g['it0_logits']).double().to(s.device)
            rp = torch.as_tensor(g['it0_reg_params']).double().reshape(-1)
            s.hyper[L.HYPER_REG_P:L.HYPER_REG_P + rp.numel()] = rp.to(s.device)
        s.set_noise(g[f'it{it}_eps'], g[f'it{it}_jitter'])
        s.step(1, use_graph=False)
        torch.cuda.synchronize()
        tol = 1e-5
        p, p64 = f'it{it}_', f'it{it}_f64_'
        out = s.output()
        for key in ('curr_state', 'transformation', 'displacement', 'im_moving_warped'):
            assert out[key].shape == g[p + key].shape and rel(out[key], g[p + key]) < tol, (it, key)
        assert rel(s._ffd_dense, g[p + 'velocity']) < tol
        terms = s.loss_terms()
        assert rel(terms['reg_energy'], g[p + 'reg_energy']) < tol and rel(terms['reg'], g[p + 'reg']) < tol
        assert rel(terms['alpha'], g[p + 'alpha']) < max(tol, 1e-4) and rel(terms['data'], g[p + 'data']) < max(tol, 1e-4)
        if it == 0:
            # kink flips (SURVEY surprise 9) are isolated voxels of the DENSE gradient: the acceptance rule applies there;
            # a control point sums 15^3 of them, so its gradient is compared in the L2 norm
            assert grad_ok(s._ffd_grad, g[p + 'grad_dense'], g[p64 + 'grad_dense'], 'golden SVFFD dense gradient')
            assert rel(s.grad_v, g[p64 + 'grad_v']) < 2e-3 and s.grad_v.shape == g[p + 'grad_v'].shape
            assert rel(s.gmm_parameters()[0], g[p + 'log_std']) < 1e-5
        assert rel(s.v, g[p + 'v_after']) < 2e-3
    # graph replay == eager, Philox noise on the control grid (the FFD launches are captured like the others)
    outs = []
    for use_graph in (False, True):
        r = SGLDSampler(fixed, moving, C, SGLDConfig(transformation='SVFFD_3D', cps=cps), device='cuda:0')
        r.set_state(g['v0'], g['sigma'])
        r.init_gmm(sigma_hat=0.7)
        r.step(4, use_graph=use_graph)
        torch.cuda.synchronize()
        outs.append((r.v.clone(), r.hyper.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert bool(torch.isfinite(outs[0][0]).all()) and not torch.equal(outs[0][0].cpu(), g['v0'])
    # Trainer.__GMM_init through the FFD (irs_sgld_gmm_init) against the oracle
    torch.manual_seed(1)
    v_sample = 1.5 * torch.randn(1, 3, *g['grid'].tolist())
    r = SGLDSampler(fixed, moving, 1, SGLDConfig(transformation='SVFFD_3D', cps=cps), device='cuda:0')
    r.init_gmm(v_sample)
    st = O.State(O.Config(cps=cps), v_sample, torch.ones_like(v_sample), (n, n, n))
    O.gmm_init(st, fixed, moving, v_sample)
    ls, lg = r.gmm_parameters()
    assert rel(ls, st.log_std) < 1e-4 and rel(lg, st.logits) < 1e-3


@pytest.mark.gpu
def test_cuda_evaluation_kernels_vs_reference_golden(built):
    """calc_DSC_GPU / calc_no_non_diffeomorphic_voxels / calc_norm of the drop-in package (CUDA kernels) against the vectors of
    the reference's functions of the same names (utils/util.py:123-148,209-225)"""
    import irsgmcmc_b200.utils as U
    dev = 'cuda:0'
    g = load('eval.npz')
    labels = [int(x) for x in g['labels']]
    C = g['seg_moving_warped'].shape[0]
    structures = {f's{l}': l for l in labels}
    dsc = U.calc_DSC_GPU(C, g['seg_fixed'].to(dev).expand(C, -1, -1, -1, -1), g['seg_moving_warped'].to(dev), structures)
    assert np.allclose(dsc, g['DSC'].numpy(), rtol=1e-6, atol=0, equal_nan=True)
    counts, log_det = U.calc_no_non_diffeomorphic_voxels(g['T'].to(dev), U.GradientOperator())
    want = g['no_non_diffeomorphic_voxels'].numpy()
    assert counts[0] == 0 and counts[1] == 0 and abs(int(counts[2]) - int(want[2])) <= 1, (counts, want)
    nan_new, nan_ref = torch.isnan(log_det.cpu()), torch.isnan(g['log_det_J'])
    assert int((nan_new != nan_ref).sum()) <= 1        # a determinant within rounding of zero may change sign
    ok = ~(nan_new | nan_ref)
    # fp32 determinants of a folding deformation: six products of O(10) cancel to O(1) (measured 3.5e-5; the oracle on the CPU: < 1e-6)
    assert rel(log_det.cpu()[ok].exp(), g['log_det_J'][ok].exp()) < 2e-4
    assert rel(U.calc_norm(g['disp'].to(dev)), g['disp_norm']) < 1e-5
