"""
The drop-in classes (reference names and signatures on the CUDA ops) on the GPU:
 - the reference's own known-answer tests, restated (reference tests/test_diff.py, tests/test_utils.py);
 - one SGLD transition composed exactly like the reference's Trainer._SGLD_transition (trainer/trainer.py:291-356) but
   out of irsgmcmc_b200's drop-in modules + autograd, against the oracle;
 - the Trainer facade: return structure of _SGLD_transition, _run_MCMC.
"""
import json
import math

import numpy as np
import pytest
import torch

from oracle import sgld_oracle as O
from tests.util import rel

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
ATOL = 1e-4  # reference tests/test_setup.py:46


@pytest.fixture(scope='module')
def pkg(built):
    import irsgmcmc_b200.utils as U
    import irsgmcmc_b200.model as M
    import irsgmcmc_b200.optimizers as Opt
    return U, M, Opt


# ---- reference tests/test_diff.py ------------------------------------------------------------------------------------
def test_kat_uniform_and_linear_fields(pkg):
    U = pkg[0]
    n = 64
    op = U.GradientOperator()
    v = torch.ones(1, 3, n, n, n, device=DEV) * 5.0
    assert torch.allclose(op(v), torch.zeros(1, 3, n, n, n, 3, device=DEV), atol=ATOL)   # test_diff.py:9-23
    ax = torch.arange(n, dtype=torch.float32, device=DEV)
    z, y, x = torch.meshgrid(ax, ax, ax, indexing='ij')
    v = torch.zeros(1, 3, n, n, n, device=DEV)
    v[0, 0], v[0, 1] = x, 1.5 * y + 3.0 * z + 1.0                                         # test_diff.py:25-49
    nab = op(v)
    assert torch.allclose(nab[0, 0, ..., 0], torch.ones_like(x), atol=ATOL)       # d v_x / d x
    assert torch.allclose(nab[0, 1, ..., 1], 1.5 * torch.ones_like(x), atol=ATOL)  # d v_y / d y
    assert torch.allclose(nab[0, 2, ..., 1], 3.0 * torch.ones_like(x), atol=ATOL)  # d v_y / d z
    assert torch.allclose(nab[0, 1, ..., 0], torch.zeros_like(x), atol=ATOL)


def test_kat_log_det_J(pkg):
    U = pkg[0]
    n = 64
    op = U.GradientOperator()
    ident = U.init_identity_grid_3D((n, n, n)).permute(0, 4, 1, 2, 3).contiguous().to(DEV)
    counts, log_det = U.calc_no_non_diffeomorphic_voxels(ident, op)                        # test_diff.py:51-57
    assert counts.sum() == 0 and torch.allclose(log_det, torch.zeros_like(log_det), atol=ATOL)
    _, log_det2 = U.calc_no_non_diffeomorphic_voxels(2.0 * ident, U.GradientOperator())     # test_diff.py:92-113
    assert torch.allclose(log_det2, math.log(8.0) * torch.ones_like(log_det2), atol=ATOL)


def test_kat_det_J_polynomial(pkg):
    """hand-built Jacobian on a 4^3 grid -> x^4 - x^2 y^3 - x^2 z + x y^2 - x y z^2 + y^2 z^3 (test_diff.py:59-90)"""
    U = pkg[0]
    n = 4
    ax = torch.arange(n, dtype=torch.float32, device=DEV)
    z, y, x = torch.meshgrid(ax, ax, ax, indexing='ij')
    nabla = torch.zeros(1, 3, n, n, n, 3, device=DEV)
    # rows j = d/dx_j, last dim i = component: J = [[x^2, y, z^2], [z, x^2, x], [y^2, y z... ]] built to give the polynomial
    nabla[0, 0, ..., 0], nabla[0, 1, ..., 0], nabla[0, 2, ..., 0] = x ** 2, y ** 2, z
    nabla[0, 0, ..., 1], nabla[0, 1, ..., 1], nabla[0, 2, ..., 1] = z ** 2, x ** 2, y
    nabla[0, 0, ..., 2], nabla[0, 1, ..., 2], nabla[0, 2, ..., 2] = y, x, x ** 0 * 1.0 + 0 * x
    det = U.calc_det_J(nabla)[0]
    J = torch.stack([torch.stack([nabla[0, j, ..., i] for i in range(3)], -1) for j in range(3)], -2)
    assert torch.allclose(det, torch.linalg.det(J.double()).float(), atol=1e-3)


# ---- reference tests/test_utils.py -----------------------------------------------------------------------------------
def test_kat_calc_norm(pkg):
    U = pkg[0]
    n = 8
    v = torch.ones(1, 3, n, n, n, device=DEV)
    assert torch.allclose(U.calc_norm(v), math.sqrt(3) * torch.ones(1, 1, n, n, n, device=DEV), atol=ATOL)
    assert torch.allclose(U.calc_norm(2 * v), math.sqrt(12) * torch.ones(1, 1, n, n, n, device=DEV), atol=ATOL)


def test_kat_separable_conv_all_ones(pkg):
    """all-ones 3-tap kernel on an all-ones field -> 27 everywhere, borders included (test_utils.py:101-151)"""
    U = pkg[0]
    n = 16
    v = torch.ones(1, 3, n, n, n, device=DEV)
    k1 = torch.ones(3, 1, 3, device=DEV)
    out = U.separable_conv_3D(v, k1, 1)
    assert torch.allclose(out, 27.0 * torch.ones_like(v), atol=ATOL)
    S = torch.ones(3, 1, 3, device=DEV)
    out = U.separable_conv_3D(v, S.unsqueeze(2).unsqueeze(2), S.unsqueeze(2).unsqueeze(4), S.unsqueeze(3).unsqueeze(4),
                              (1,) * 6)
    assert torch.allclose(out, 27.0 * torch.ones_like(v), atol=ATOL)


def test_registration_module_dtypes_and_errors(pkg):
    U = pkg[0]
    n = 12
    reg = U.RegistrationModule()
    T = U.init_identity_grid_3D((n, n, n)).permute(0, 4, 1, 2, 3).contiguous().to(DEV)
    im = torch.rand(1, 1, n, n, n, device=DEV)
    seg = (torch.rand(1, 1, n, n, n, device=DEV) * 50).short()
    mask = torch.rand(1, 1, n, n, n, device=DEV) > 0.5
    assert rel(reg(im, T), im) < 1e-6                       # identity transformation
    assert torch.equal(reg(seg, T), seg) and reg(seg, T).dtype == torch.int16
    assert torch.equal(reg(mask, T), mask) and reg(mask, T).dtype == torch.bool
    with pytest.raises(NotImplementedError):                # reference utils/registration.py:32
        reg(im.double(), T)
    with pytest.raises(NotImplementedError):                # no CPU path
        reg(im.cpu(), T.cpu())
    with pytest.raises(ValueError):                         # reference utils/diff_op.py:31
        U.DifferentialOperator.from_string('NoSuchOperator')


# ---- one transition composed like the reference's Trainer, out of the drop-in modules ---------------------------------
@pytest.mark.parametrize('reg_name', ['RegLoss_LogNormal', 'RegLoss_L2'])
def test_reference_style_transition_with_dropin_modules(pkg, reg_name):
    U, M, Opt = pkg
    from irsgmcmc_b200.data_loader.synthetic import make_pair
    n, C, tau, alpha_j = 16, 2, 0.4, 0.1
    torch.manual_seed(123)
    fixed, moving, vp = make_pair(n)
    dof = 3.0 * n ** 3
    gmm = M.GMM(4, 2).to(DEV)
    gmm.init_parameters(0.7)
    w_reg = 1.6 if reg_name == 'RegLoss_LogNormal' else 1.4
    reg = getattr(M, reg_name)(w_reg=w_reg, diff_op='GradientOperator', dims=(n, n, n), learnable=True).to(DEV)
    scale_prior, prop_prior = M.LogScaleNormalPrior(0.0, 2.3).to(DEV), M.DirichletPrior(4, 0.5).to(DEV)
    if reg_name == 'RegLoss_LogNormal':
        loc_prior, reg_scale_prior = M.LogEnergyExpGammaPrior(w_reg, dof).to(DEV), M.LogScaleNormalPrior(2.8, 5.0).to(DEV)
        opt_reg = Opt.Adam([{'params': [reg.loc], 'lr': 0.01}, {'params': [reg.log_scale], 'lr': 0.01}], lr_decay=1e-3)
    else:
        w_prior = M.LogPrecisionExpGammaPrior(shape=0.5 * dof, rate=1.0 / (0.5 * dof)).to(DEV)
        opt_reg = Opt.Adam(reg.parameters(), lr=0.01, lr_decay=1e-3)
    opt_gmm = Opt.Adam([{'params': [gmm.log_std], 'lr': 0.2}, {'params': [gmm.logits], 'lr': 0.2}], lr_decay=1e-3)
    svf, regm = U.SVF_3D((n, n, n)).to(DEV), U.RegistrationModule()
    taps = torch.from_numpy(U.Sobolev_kernel_1D(3, 0.5)[0]).float().unsqueeze(0)
    S3 = torch.stack((taps, taps, taps), 0).to(DEV)
    S = {'x': S3.unsqueeze(2).unsqueeze(2), 'y': S3.unsqueeze(2).unsqueeze(4), 'z': S3.unsqueeze(3).unsqueeze(4)}

    v0 = 0.8 * torch.randn(C, 3, n, n, n)
    sigma = torch.exp(0.5 * vp['log_var']).expand(C, -1, -1, -1, -1).contiguous()
    eps, ju = torch.randn(C, 3, n, n, n), torch.rand(C, 3, n, n, n)
    v = v0.clone().to(DEV).requires_grad_(True)
    opt_v = torch.optim.SGD([v], lr=tau)
    fx = {k: t.to(DEV).expand(C, *t.shape[1:]) for k, t in fixed.items()}
    mv = {k: t.to(DEV).expand(C, *t.shape[1:]) for k, t in moving.items()}

    # --- trainer.py:292-356, line by line, with injected noise instead of torch.randn / torch.rand ---
    class _SGLDExplicit(torch.autograd.Function):       # SGLD.apply with a given eps (utils/functions.py:76-84)
        @staticmethod
        def forward(ctx, state, sg, tau_):
            ctx.sg = sg
            return U.langevin_sobolev(state.detach().contiguous(), sg, math.sqrt(2 * tau_), [], eps=eps.to(DEV))

        @staticmethod
        def backward(ctx, g):
            return ctx.sg ** 2 * g, None, None

    curr_state = _SGLDExplicit.apply(v, sigma.to(DEV), tau)
    curr_state_smoothed = U.SobolevGrad.apply(curr_state, S, (3,) * 6)
    transformation, displacement = svf(curr_state_smoothed)
    T_noise = transformation + U.transform_coordinates(-2.0 * alpha_j * ju.to(DEV) + alpha_j)
    im_w = regm(mv['im'].contiguous(), T_noise)
    residuals = gmm.map(fx['im'].contiguous(), im_w)
    residuals_masked = residuals[fx['mask']].view(C, -1)
    reg_term, log_y = reg(curr_state_smoothed)
    data_term, alphas, data_terms = 0.0, [], []
    for idx in range(C):
        r = U.rescale_residuals(residuals[idx].unsqueeze(0).detach(), fx['mask'][idx].unsqueeze(0), gmm)
        a = U.calc_VD_factor(r, fx['mask'][idx].unsqueeze(0))
        step_loss = gmm(residuals_masked[idx].unsqueeze(0).detach()).sum() * a
        step_loss = step_loss - scale_prior(gmm.log_scales).sum() - prop_prior(gmm.log_proportions).sum()
        opt_gmm.zero_grad()
        step_loss.backward()
        opt_gmm.step()
        term = gmm(residuals_masked[idx]).sum() * a
        data_term = data_term + term
        alphas.append(a)
        data_terms.append(term.detach())
    data_term = data_term - scale_prior(gmm.log_scales).sum() - prop_prior(gmm.log_proportions).sum()
    reg_total = reg_term.sum()
    if reg_name == 'RegLoss_LogNormal':
        reg_total = reg_total - loc_prior(log_y).sum() - reg_scale_prior(reg.log_scale).sum()
    else:
        reg_total = reg_total - w_prior(reg.log_w_reg)
    loss = data_term + reg_total
    opt_v.zero_grad()
    opt_reg.zero_grad()
    loss.backward()
    grad_v = v.grad.detach().clone()
    opt_v.step()
    opt_reg.step()

    # --- the oracle on the same inputs ---
    reg_key = 'lognormal' if reg_name == 'RegLoss_LogNormal' else 'l2'
    for dtype in (torch.float32, torch.float64):
        st = O.State(O.Config(reg=reg_key, w_reg=w_reg, exact_grid=dtype == torch.float64), v0.to(dtype), sigma.to(dtype),
                     (n, n, n), dtype)
        st.init_gmm(0.7)
        cast = lambda d: {k: (t.to(dtype) if t.dtype == torch.float32 else t) for k, t in d.items()}
        lt, out, aux, g_or = O.sgld_transition(st, cast(fixed), cast(moving), eps.to(dtype), ju.to(dtype))
        if dtype == torch.float32:
            g32, st32 = g_or, st
        else:
            g64, out64, aux64, lt64, st64 = g_or, out, aux, lt, st
    assert rel(curr_state_smoothed, out64['curr_state']) < 1e-5
    assert rel(transformation, out64['transformation']) < 1e-5 and rel(displacement, out64['displacement']) < 1e-5
    assert rel(im_w, out64['im_moving_warped']) < 1e-5
    assert rel(residuals, aux64['residuals']) < 2e-5
    assert rel(torch.stack(alphas), torch.stack(aux64['alpha'])) < 1e-4
    assert rel(torch.stack(data_terms), torch.stack(lt64['data'])) < 1e-4
    assert rel(reg_term, torch.stack(lt64['reg'])) < 1e-6
    from tests.util import grad_ok
    assert grad_ok(grad_v, g32, g64, 'drop-in composed gradient')   # the three-number protocol, no extra floor
    assert rel(gmm.log_std, st64.log_std) < 1e-5 and rel(gmm.logits, st64.logits) < 1e-4
    if reg_name == 'RegLoss_LogNormal':
        assert rel(torch.stack((reg.loc, reg.log_scale)), torch.stack((st64.loc, st64.log_scale))) < 1e-7
    else:
        assert rel(reg.log_w_reg, st64.log_w_reg) < 1e-6


# ---- Trainer facade ---------------------------------------------------------------------------------------------------
def _reference_style_config(no_chains=2, burn_in=4, samples=12, period=4):
    return {'data_loss': {'type': 'GMM', 'args': {'no_components': 4, 's': 2}},
            'data_loss_scale_prior': {'type': 'LogScaleNormalPrior', 'args': {'loc': 0.0, 'scale': 2.3}},
            'data_loss_proportion_prior': {'type': 'DirichletPrior', 'args': {'no_classes': 4, 'alpha': 0.5}},
            'reg_loss': {'type': 'RegLoss_LogNormal', 'args': {'diff_op': 'GradientOperator', 'w_reg': 1.6, 'learnable': True}},
            'reg_loss_scale_prior': {'type': 'LogScaleNormalPrior', 'args': {'loc': 2.8, 'scale': 5.0}},
            'optimizer_GMM': {'type': 'Adam', 'args': {'lr_log_std': 0.2, 'lr_logits': 0.2, 'lr_decay': 0.001}},
            'optimizer_reg': {'type': 'Adam', 'args': {'lr_loc': 0.01, 'lr_log_scale': 0.01, 'lr_decay': 0.001}},
            'optimizer_SG_MCMC': {'type': 'SGD', 'args': {'lr': 0.4}},
            'Sobolev_grad': {'enabled': True, 's': 3, 'lambda': 0.5}, 'virtual_decimation': True,
            'trainer': {'MCMC_init': 'VI', 'no_chains': no_chains, 'no_iters_burn_in': burn_in, 'no_samples_MCMC': samples,
                        'log_period_MCMC': period, 'uniform_noise': {'enabled': True, 'magnitude': 0.1}}}


def test_trainer_facade(pkg):
    U, M, _ = pkg
    from irsgmcmc_b200.trainer import Trainer
    from irsgmcmc_b200.data_loader.synthetic import make_pair, STRUCTURE_LABELS
    n, C = 16, 2
    torch.manual_seed(5)
    fixed, moving, vp = make_pair(n)
    structures = {f's{l}': l for l in STRUCTURE_LABELS}
    import tempfile
    save_dir = tempfile.mkdtemp()
    t = Trainer(_reference_style_config(C), fixed, moving, vp, structures_dict=structures, device=torch.device(DEV),
                save_dir=save_dir, im_spacing=(2.0, 2.0, 2.0))
    gmm = M.GMM(4, 2).to(DEV)
    gmm.init_parameters(0.7)
    reg = M.RegLoss_LogNormal(w_reg=1.6, diff_op='GradientOperator', dims=(n, n, n), learnable=True).to(DEV)
    t._SGLD_init()
    before = gmm.log_std.detach().clone()
    loss_terms, output, aux = t._SGLD_transition(fixed, moving, gmm, reg)
    assert set(loss_terms) == {'data', 'reg'} and len(loss_terms['data']) == C and len(loss_terms['reg']) == C
    assert set(output) == {'im_moving_warped', 'displacement', 'transformation', 'curr_state'}
    assert set(aux) == {'residuals', 'alpha', 'reg_energy'} and len(aux['alpha']) == C
    assert output['displacement'].shape == (C, 3, n, n, n) and output['im_moving_warped'].shape == (C, 1, n, n, n)
    assert not torch.equal(gmm.log_std.detach(), before)     # the shared mixture was stepped and mirrored back
    assert all(torch.isfinite(x) for x in loss_terms['data'] + loss_terms['reg'] + aux['alpha'])
    res = t._run_MCMC(gmm, reg, speed_test_iters=3)
    # kept: iterations 8, 12 (> burn-in 4, multiple of 4 or == no_samples) and 16 -> 3 kept x 2 chains
    assert res['n'] == 3 * C and res['mean'].shape == (3, n, n, n) and torch.isfinite(res['std_dev']).all()
    assert len(res['DSC']) == 3 and res['DSC'][0].shape == (C, len(structures))
    assert res['samples_per_sec'] > 0
    assert isinstance(res['ASD'], str) and res['ASD'].startswith('unavailable')
    # kept samples and posterior statistics on disk with the reference's names (logger/logger.py:215-238)
    import os
    from irsgmcmc_b200.logger import load_field_from_disk, load_im_from_disk
    names = sorted(os.listdir(save_dir))
    assert 'chain_0_sample_0000008_displacement.vtk' in names and 'chain_1_sample_0000016_im_moving_warped.nii.gz' in names
    assert 'chain_1_sample_0000012_log_det_J.nii.gz' in names and len(names) == 3 * C * 3 + 4
    mean, sp = load_field_from_disk(os.path.join(save_dir, 'MCMC_sample_mean.vtk'))   # reference logger/logger.py:110-131
    assert np.allclose(mean, 2.0 * res['mean'].cpu().numpy(), atol=1e-6) and sp == [2.0, 2.0, 2.0]
    std_masked, _ = load_field_from_disk(os.path.join(save_dir, 'MCMC_sample_std_dev_masked.vtk'))
    assert np.allclose(std_masked, 2.0 * (res['std_dev'] * moving['mask'][0].to(res['std_dev'].device)).cpu().numpy(), atol=1e-6)
    im, _ = load_im_from_disk(os.path.join(save_dir, 'chain_0_sample_0000016_im_moving_warped.nii.gz'))
    assert im.shape == (n, n, n) and np.isfinite(im).all()


# ---- section 8f "next" rows: evaluation kernels and the VI warm start ---------------------------------------------------
def test_log_det_jacobian_and_dice_kernels(pkg):
    from irsgmcmc_b200 import ops
    from tests.util import smooth_field
    n, C = 20, 2
    T, _ = O.svf_exp_aten(smooth_field((C, 3, n, n, n), 6.0, 3), 6)
    T = T.contiguous()
    T[0, :, 5:8, 5:8, 5:8] = T[0, :, 5:8, 5:8, 5:8].flip(-1)        # fold a few voxels
    counts, log_det = ops.log_det_jacobian(T.to(DEV))
    ref = O.det_jacobian(O.forward_differences(T.double(), transformation=True)).log()
    ok = torch.isfinite(ref)
    assert counts.cpu().tolist() == torch.isnan(ref).sum(dim=(1, 2, 3)).tolist() and counts.sum() > 0
    assert rel(log_det.cpu()[ok], ref[ok]) < 1e-4
    labels = [10, 11, 12, 13, 16]
    a = torch.tensor(labels + [0, 0, 7])[torch.randint(0, 8, (1, 1, n, n, n))].short()
    b = torch.tensor(labels + [0, 0, 9])[torch.randint(0, 8, (C, 1, n, n, n))].short()
    cnt = ops.dice_counts(a.to(DEV), b.to(DEV), labels).cpu()
    for c in range(C):
        for j, l in enumerate(labels):
            assert cnt[c, j].tolist() == [int((a == l).sum()), int((b[c] == l).sum()), int(((a[0] == l) & (b[c] == l)).sum())]
    U = pkg[0]
    dsc = U.calc_DSC_GPU(C, a.to(DEV).expand(C, -1, -1, -1, -1), b.to(DEV), {f's{l}': l for l in labels})
    want = [[2.0 * float(((a[0] == l) & (b[c] == l)).sum()) / float((a == l).sum() + (b[c] == l).sum()) for l in labels]
            for c in range(C)]
    assert np.allclose(dsc, np.array(want), rtol=1e-6)
    # calc_metrics = (ASD on the host, DSC from the kernel) with the reference's signature (utils/util.py:151-206)
    structures = {f's{l}': l for l in labels}
    ASD, DSC = U.calc_metrics(a.to(DEV).expand(C, -1, -1, -1, -1), b.to(DEV), structures, torch.tensor([1.0, 1.0, 1.0]), no_samples=C)
    assert ASD.shape == DSC.shape == (C, len(labels)) and np.allclose(DSC, dsc) and np.isfinite(ASD).all() and (ASD > 0).all()
    same, _ = U.calc_metrics(b.to(DEV), b.to(DEV), structures, (1.0, 1.0, 1.0), no_samples=C)
    assert (same == 0).all()


@pytest.mark.parametrize('reg_name', ['RegLoss_LogNormal', 'RegLoss_L2'])
def test_vi_sample_loss_parity(pkg, reg_name):
    """Trainer._calc_sample_loss_VI (= reference Trainer.__calc_sample_loss_VI, trainer.py:79-117) on the CUDA operators,
    loss terms and gradients w.r.t. (mu, log_var, u) against the oracle"""
    from irsgmcmc_b200.trainer import Trainer
    from irsgmcmc_b200.data_loader.synthetic import make_pair
    n = 16
    torch.manual_seed(9)
    fixed, moving, vp0 = make_pair(n)
    cfg = _reference_style_config(1)
    cfg['reg_loss']['type'] = reg_name
    cfg['optimizer_reg'] = {'type': 'Adam', 'args': {'lr_loc': 0.01, 'lr_log_scale': 0.01, 'lr_log_w_reg': 0.01, 'lr_decay': 0.001}}
    t = Trainer(cfg, fixed, moving, vp0, device=torch.device(DEV))
    m = t._build_VI_modules()
    m['data_loss'].init_parameters(0.7)
    reg_key = 'lognormal' if reg_name == 'RegLoss_LogNormal' else 'l2'
    eps, x, ju = torch.randn(1, 3, n, n, n), torch.randn(1), torch.rand(1, 3, n, n, n)
    vp = {k: v.clone().to(DEV).requires_grad_(True) for k, v in vp0.items()}
    sample = vp['mu'] + eps.to(DEV) * torch.exp(0.5 * vp['log_var']) + x.to(DEV) * vp['u']
    fx, mv = {k: v.to(DEV) for k, v in fixed.items()}, {k: v.to(DEV) for k, v in moving.items()}
    lt, out, aux = t._calc_sample_loss_VI(m, fx, mv, vp, sample, ju.to(DEV))
    g_new = torch.autograd.grad(lt['data'] + lt['reg'] - lt['entropy'], [vp[k] for k in ('mu', 'log_var', 'u')])
    results = {}
    for dtype in (torch.float32, torch.float64):
        st = O.State(O.Config(reg=reg_key, w_reg=1.6, exact_grid=dtype == torch.float64), torch.zeros(1, 3, n, n, n, dtype=dtype),
                     torch.ones(1, 3, n, n, n, dtype=dtype), (n, n, n), dtype)
        st.init_gmm(0.7)
        vpo = {k: v.clone().to(dtype).requires_grad_(True) for k, v in vp0.items()}
        so = vpo['mu'] + eps.to(dtype) * torch.exp(0.5 * vpo['log_var']) + x.to(dtype) * vpo['u']
        leaves = (st.loc.clone().requires_grad_(True), st.log_scale.clone().requires_grad_(True)) if reg_key == 'lognormal' \
            else (st.log_w_reg.clone().requires_grad_(True),)
        cast = lambda d_: {k: (v.to(dtype) if v.dtype == torch.float32 else v) for k, v in d_.items()}
        terms = O.vi_sample_loss(st, cast(fixed), cast(moving), vpo, so, ju.to(dtype), leaves)
        g = torch.autograd.grad(terms['data'] + terms['reg'] - terms['entropy'], [vpo[k] for k in ('mu', 'log_var', 'u')])
        results[dtype] = (terms, g, st)
    t64, g64, st64 = results[torch.float64]
    _, g32, _ = results[torch.float32]
    assert rel(out['im_moving_warped'], t64['im_w']) < 1e-5 and rel(out['displacement'], t64['disp']) < 1e-5
    assert abs(float(aux['alpha']) - float(t64['alpha'])) < 1e-4 * float(t64['alpha'])
    for key in ('data', 'reg', 'entropy'):
        assert rel(lt[key], t64[key]) < 1e-4, key
    from tests.util import grad_ok
    for name, a, b32, b64 in zip(('mu', 'log_var', 'u'), g_new, g32, g64):
        assert grad_ok(a, b32, b64, f'VI grad {name}')
    assert rel(m['data_loss'].log_std, st64.log_std) < 1e-4


def test_run_VI_then_MCMC(pkg):
    """config 1 of BASELINE.json in miniature: VI warm start, then SGLD chains initialised from q(v)"""
    from irsgmcmc_b200.trainer import Trainer
    from irsgmcmc_b200.data_loader.synthetic import make_pair
    n, C = 16, 2
    torch.manual_seed(3)
    fixed, moving, vp0 = make_pair(n)
    cfg = _reference_style_config(C, burn_in=2, samples=6, period=2)
    cfg['trainer']['no_iters_VI'] = 4
    t = Trainer(cfg, fixed, moving, vp0, device=torch.device(DEV))
    vp, m, hist = t._run_VI()
    assert len(hist) == 4 and all(torch.isfinite(h['loss']) for h in hist)
    assert not torch.equal(vp['mu'].cpu(), vp0['mu']) and vp['log_var'].shape == vp0['log_var'].shape
    t._SGLD_init(vp)
    res = t._run_MCMC(m['data_loss'], m['reg_loss'], speed_test_iters=0)
    assert res['n'] == 3 * C and torch.isfinite(res['mean']).all()


def test_test_VI(pkg):
    """Trainer._test_VI (reference trainer.py:226-289): samples of q(v) -> folding count, Dice / ASD, files, sample statistics --
    against the same quantities composed by hand from the drop-in modules with the same random numbers"""
    import os
    import tempfile
    U, M, _ = pkg
    from irsgmcmc_b200.trainer import Trainer
    from irsgmcmc_b200.data_loader.synthetic import make_pair, STRUCTURE_LABELS
    from irsgmcmc_b200.logger import load_field_from_disk, load_im_from_disk
    n, S = 16, 3
    torch.manual_seed(11)
    fixed, moving, vp0 = make_pair(n)
    structures = {f's{l}': l for l in STRUCTURE_LABELS}
    cfg = _reference_style_config(1)
    cfg['trainer'].update(no_iters_VI=2, no_samples_VI_test=S)
    save_dir = tempfile.mkdtemp()
    t = Trainer(cfg, fixed, moving, vp0, structures_dict=structures, device=torch.device(DEV), save_dir=save_dir,
                im_spacing=(1.5, 1.5, 1.5))
    vp, m, _ = t._run_VI()
    torch.manual_seed(77)
    res = t._test_VI(speed_test_samples=2)
    assert res['n'] == S and len(res['no_non_diffeomorphic_voxels']) == S and res['samples_per_sec'] > 0
    assert res['DSC'].shape == (S, len(structures)) and res['ASD'].shape == (S, len(structures))
    ok = np.isfinite(res['DSC'])   # a structure that is empty in both segmentations at this size gives 0 / 0
    assert ok.any() and (res['DSC'][ok] >= 0).all() and (res['DSC'][ok] <= 1).all()

    # the same draws by hand (same generator state, same order of random numbers: randn_like(sigma), randn(1) per sample)
    torch.manual_seed(77)
    vpd = {k: v.to(DEV) for k, v in vp.items()}
    disp, dsc = [], []
    for _ in range(S):
        v = U.SobolevGrad.apply(U.sample_q_v(vpd), m['S'], m['padding'])
        T, d = m['transformation_module'](v)
        disp.append(d[0])
        seg_w = m['registration_module'](moving['seg'].to(DEV), T)
        dsc.append(U.calc_DSC_GPU(1, fixed['seg'].to(DEV), seg_w, structures)[0])
    disp = torch.stack(disp)
    assert rel(res['mean'], disp.mean(0)) < 1e-6 and rel(res['std_dev'], disp.std(0)) < 1e-5   # unbiased, like torch.std
    assert np.array_equal(res['DSC'], np.asarray(dsc), equal_nan=True)
    T_mu, d_mu = m['transformation_module'](U.SobolevGrad.apply(vpd['mu'], m['S'], m['padding']))
    assert torch.equal(res['displacement_mu'], d_mu)

    names = sorted(os.listdir(save_dir))
    expect = [f'sample_{i:07}_{k}' for i in range(1, S + 1) for k in ('displacement.vtk', 'im_moving_warped.nii.gz', 'log_det_J.nii.gz')]
    expect += ['VI_sample_mean.vtk', 'VI_sample_mean_masked.vtk', 'VI_sample_std_dev.vtk', 'VI_sample_std_dev_masked.vtk',
               'displacement_mu.vtk', 'im_moving_warped_mu.nii.gz']
    assert names == sorted(expect)
    f, sp = load_field_from_disk(os.path.join(save_dir, 'displacement_mu.vtk'))
    assert np.allclose(f, 1.5 * d_mu[0].cpu().numpy(), atol=1e-6) and sp == [1.5, 1.5, 1.5]
    f, _ = load_field_from_disk(os.path.join(save_dir, 'sample_0000002_displacement.vtk'))
    assert np.allclose(f, 1.5 * disp[1].cpu().numpy(), atol=1e-6)
    im, _ = load_im_from_disk(os.path.join(save_dir, 'im_moving_warped_mu.nii.gz'))
    assert np.allclose(im, res['im_moving_warped_mu'][0, 0].cpu().numpy(), atol=1e-7)

    # _run_model runs it after VI when the config asks for test samples (reference trainer.py:488-498)
    t2 = Trainer(cfg, fixed, moving, vp0, structures_dict=structures, device=torch.device(DEV))
    out = t2._run_model(VI=True, MCMC=False)
    assert out['VI_test']['n'] == S and 'samples_per_sec' not in out['VI_test']


def test_run_VI_then_MCMC_svffd(pkg):
    """the same with SVFFD_3D as the transformation model (reference configs/experiment5/config_SVFFD_4.json): variational
    parameters, chain state and preconditioner on the control grid; VI in drop-in mode, sampling with the fused step"""
    from irsgmcmc_b200.trainer import Trainer
    from irsgmcmc_b200.data_loader.synthetic import make_pair
    U = pkg[0]
    n, C, cps = 16, 2, [4, 4, 4]
    torch.manual_seed(4)
    fixed, moving, _ = make_pair(n)
    gdims = (1, 3, *U.get_control_grid_size((n,) * 3, cps))
    vp0 = {'mu': torch.zeros(gdims), 'log_var': torch.full(gdims, math.log(0.5 ** 2)), 'u': torch.full(gdims, 0.1)}
    cfg = _reference_style_config(C, burn_in=2, samples=6, period=2)
    cfg['transformation_module'] = {'type': 'SVFFD_3D', 'args': {'cps': cps}}
    cfg['trainer']['no_iters_VI'] = 3
    t = Trainer(cfg, fixed, moving, vp0, device=torch.device(DEV))
    assert t.sampler.v.shape == (C, *gdims[1:]) and t.sampler.displacement.shape == (C, 3, n, n, n)
    vp, m, hist = t._run_VI()
    assert isinstance(m['transformation_module'], U.SVFFD_3D)
    assert len(hist) == 3 and all(torch.isfinite(h['loss']) for h in hist)
    assert vp['mu'].shape == gdims and not torch.equal(vp['mu'].cpu(), vp0['mu'])
    t._SGLD_init(vp)
    lt, out, aux = t._SGLD_transition(None, None, m['data_loss'], m['reg_loss'])
    assert out['curr_state'].shape == (C, *gdims[1:]) and out['transformation'].shape == (C, 3, n, n, n)
    assert all(torch.isfinite(x) for x in lt['data'] + lt['reg'])
    res = t._run_MCMC(m['data_loss'], m['reg_loss'], speed_test_iters=0)
    assert res['n'] == 3 * C and torch.isfinite(res['mean']).all() and res['mean'].shape == (3, n, n, n)


def test_adam_state_travels_from_VI_to_MCMC(pkg):
    """the reference creates optimizer_GMM / optimizer_reg once and they persist through __GMM_init, VI and
    MCMC (trainer/trainer.py:62-66).  After two VI iterations the first SGLD transition must step the mixture and the
    regulariser hyper-parameters with the step counters, decayed rates and moments VI left behind -- compared with the
    oracle carrying ONE AdamState through the same sequence (a restart from step 0 moves log_std ~3x further)."""
    from irsgmcmc_b200.trainer import Trainer
    from irsgmcmc_b200.data_loader.synthetic import make_pair
    n, C, n_vi = 16, 2, 2
    torch.manual_seed(21)
    fixed, moving, vp0 = make_pair(n)
    cfg = _reference_style_config(C)
    cfg['optimizer_reg'] = {'type': 'Adam', 'args': {'lr_loc': 0.01, 'lr_log_scale': 0.01, 'lr_decay': 0.001}}
    cfg['optimizer_q_v'] = {'type': 'Adam', 'args': {'lr_mu': 0.0, 'lr_log_var': 0.0, 'lr_u': 0.0, 'lr_decay': 0.0}}
    t = Trainer(cfg, fixed, moving, vp0, device=torch.device(DEV))
    m = t._build_VI_modules()
    m['data_loss'].init_parameters(0.7)
    noise = [(torch.randn(1, 3, n, n, n), torch.randn(1), torch.rand(1, 3, n, n, n), torch.rand(1, 3, n, n, n))
             for _ in range(n_vi)]
    t._run_VI(vp0, no_iters=n_vi, modules=m, noise=iter([tuple(x.to(DEV) for x in nz) for nz in noise]))
    step_after_vi = m['optimizer_GMM'].state[m['data_loss'].log_std]['step']
    assert step_after_vi == 2 * n_vi

    st = O.State(O.Config(reg='lognormal', w_reg=1.6), torch.zeros(C, 3, n, n, n), torch.ones(1), (n, n, n))
    st.init_gmm(0.7)
    for eps, x, j1, j2 in noise:   # q(v) is frozen (lr 0): both sides see the same samples
        O.vi_iteration(st, fixed, moving, vp0, eps, x, j1, j2)
    assert rel(m['data_loss'].log_std, st.log_std) < 1e-4 and st.adam_gmm.step_no == step_after_vi

    v0 = 0.5 * torch.randn(C, 3, n, n, n)
    sigma = torch.exp(0.5 * vp0['log_var'])
    eps, ju = torch.randn(C, 3, n, n, n), torch.rand(C, 3, n, n, n)
    t.sampler.set_state(v0, sigma)
    t.sampler.set_noise(eps, ju)
    t.SGLD_params = {'sigma': t.sampler.sigma, 'tau': 0.4}
    ls_before = m['data_loss'].log_std.detach().clone()
    t._SGLD_transition(None, None, m['data_loss'], m['reg_loss'])     # pushes parameters + optimiser state, steps, pulls
    st.v, st.sigma = v0.clone(), sigma.expand(C, -1, -1, -1, -1)
    O.sgld_transition(st, fixed, moving, eps, ju)
    moved = float((m['data_loss'].log_std.detach().cpu() - ls_before.cpu()).norm())
    err = float((m['data_loss'].log_std.detach().cpu() - st.log_std).norm())
    print('log_std moved by', moved, 'distance from the oracle', err)
    assert err < 2e-3 * moved + 1e-6
    assert rel(m['data_loss'].logits, st.logits) < 1e-3
    assert rel(torch.stack((m['reg_loss'].loc.detach().cpu(), m['reg_loss'].log_scale.detach().cpu())).double(),
               torch.stack((st.loc, st.log_scale))) < 1e-7
    t._pull_hyper(m['data_loss'], m['reg_loss'])                      # optimiser state mirrored back on request
    assert m['optimizer_GMM'].state[m['data_loss'].log_std]['step'] == step_after_vi + C
    assert rel(m['optimizer_GMM'].state[m['data_loss'].log_std]['exp_avg'], st.adam_gmm.m[0]) < 1e-3
    assert m['optimizer_reg'].state[m['reg_loss'].loc]['step'] == n_vi + 1


def test_run_model_GMM_init_then_VI_then_MCMC(pkg):
    """Trainer._run_model = the reference's orchestration (trainer/trainer.py:478-504): __GMM_init, _run_VI, _run_MCMC with
    the optimiser state handed over between the stages"""
    from irsgmcmc_b200.trainer import Trainer
    from irsgmcmc_b200.data_loader.synthetic import make_pair
    n, C = 16, 2
    torch.manual_seed(3)
    fixed, moving, vp0 = make_pair(n)
    cfg = _reference_style_config(C, burn_in=2, samples=6, period=2)
    cfg['trainer'].update({'no_iters_VI': 3, 'VI': True, 'MCMC': True})
    t = Trainer(cfg, fixed, moving, vp0, device=torch.device(DEV))
    res = t._run_model()
    m = res['modules']
    assert len(res['VI_history']) == 3 and res['n'] == 3 * C and torch.isfinite(res['mean']).all()
    # 25 warm-up steps + 2 per VI iteration + C per transition (8 transitions)
    assert m['optimizer_GMM'].state[m['data_loss'].log_std]['step'] == 25 + 2 * 3 + C * 8
