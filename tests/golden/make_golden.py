"""
Generates the golden vectors under tests/golden/ by running the UNMODIFIED reference (dgrzech/ir-sgmcmc, imported from
/root/reference with I/O modules stubbed: oracle/ref_import.py) on seeded inputs.  Run in the build container only:

    python tests/golden/make_golden.py

Everything the reference computes on the SGLD hot path is stored for two consecutive `Trainer._SGLD_transition` calls
(fp32, as the reference runs) together with the injected noise, plus op-level vectors and the fp64 gradient of the
reference modules for the three-number gradient protocol.
"""
import math
import os
import sys
import warnings

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
warnings.filterwarnings('ignore')

from oracle import ref_import  # noqa: E402
from irsgmcmc_b200.data_loader.synthetic import make_pair  # noqa: E402
from tests.util import smooth_field  # noqa: E402


def np32(t):
    return t.detach().cpu().numpy().copy()   # copy: parameters are updated in place by the optimisers later on


def transition_vectors(ref, n, C, reg_type, learnable, w_reg, tag, iters=2, cps=None):
    """cps: SVFFD_3D as the transformation module (configs/experiment5): state, sigma and eps on the control grid"""
    torch.manual_seed(123)
    fixed, moving, vp = make_pair(n)
    out = {'n': n, 'C': C}
    sdims = (n, n, n) if cps is None else ref.util.get_control_grid_size((n, n, n), cps)
    if cps is not None:
        out['cps'], out['grid'] = np.array(cps), np.array(sdims)
    for dtype, sfx in ((torch.float32, ''), (torch.float64, '_f64')):
        torch.manual_seed(7)
        t = ref_import.make_trainer(ref, (n, n, n), C, reg_type=reg_type, w_reg=w_reg, learnable=learnable,
                                    uniform_noise=0.1, dtype=dtype if dtype == torch.float64 else None, cps=cps)
        dense = []
        if cps is not None:   # keep the gradient w.r.t. the dense velocity field (see ffd_vectors)
            def keep(mod, inp, res):
                res.retain_grad()
                dense.append(res)

            t.transformation_module.cubic_B_spline_FFD.register_forward_hook(keep)
        if dtype == torch.float64:  # RegistrationModule rejects fp64 (utils/registration.py:13-15,32)
            t.registration_module = lambda im, T: F.grid_sample(im, T.permute(0, 2, 3, 4, 1), mode='bilinear',
                                                                padding_mode='border', align_corners=True)
        gmm, reg = t.losses['data']['loss'], t.losses['reg']['loss']
        gmm.init_parameters(torch.tensor(0.7))
        cast = lambda d: {k: (v.to(dtype) if v.dtype == torch.float32 else v).expand(C, *v.shape[1:]) for k, v in d.items()}
        fx, mv = cast(fixed), cast(moving)
        v0 = ((0.8 if cps is None else 1.5) * torch.randn(C, 3, *sdims)).to(dtype)
        sigma = torch.exp(0.5 * vp['log_var']).to(dtype)
        if cps is not None:
            sigma = (0.5 + torch.rand(1, 3, *sdims)).to(dtype)
        sigma = sigma.expand(C, -1, -1, -1, -1)
        ref_import.attach_state(t, v0, sigma, 0.4)
        if sfx == '':
            out['v0'], out['sigma'] = np32(v0), np32(sigma[:1])
        for it in range(iters):
            eps = torch.randn(C, 3, *sdims)
            ju = torch.rand(C, 3, n, n, n)
            dense.clear()
            ref.util.get_noise_Langevin = lambda s, tau, e=eps.to(dtype): math.sqrt(2.0 * tau) * s * e
            ref.util.get_noise_uniform = lambda shape, device, alpha, j=ju.to(dtype): -2.0 * alpha * j + alpha
            v_before = t.v_curr_state.detach().clone()
            lt, output, aux = t._SGLD_transition(fx, mv, gmm, reg)
            residual_full = gmm.map(fx['im'], output['im_moving_warped'])
            p = f'it{it}{sfx}_'
            if sfx == '':
                out[f'it{it}_eps'], out[f'it{it}_jitter'] = np32(eps), np32(ju)
            out[p + 'v_before'] = np32(v_before)
            out[p + 'curr_state'] = np32(output['curr_state'])
            out[p + 'transformation'] = np32(output['transformation'])
            out[p + 'displacement'] = np32(output['displacement'])
            out[p + 'im_moving_warped'] = np32(output['im_moving_warped'])
            out[p + 'residuals'] = np32(residual_full)
            out[p + 'data'] = np.array([float(x) for x in lt['data']])
            out[p + 'reg'] = np.array([float(x) for x in lt['reg']])
            out[p + 'alpha'] = np.array([float(x) for x in aux['alpha']])
            out[p + 'reg_energy'] = np.array([float(x) for x in aux['reg_energy']])
            out[p + 'grad_v'] = np32((v_before - t.v_curr_state.detach()) / 0.4)
            out[p + 'v_after'] = np32(t.v_curr_state)
            if cps is not None:
                out[p + 'velocity'], out[p + 'grad_dense'] = np32(dense[0]), np32(dense[0].grad)
            out[p + 'log_std'], out[p + 'logits'] = np32(gmm.log_std), np32(gmm.logits)
            out[p + 'reg_params'] = np.array([float(reg.loc), float(reg.log_scale)] if reg_type == 'RegLoss_LogNormal'
                                             else [float(reg.log_w_reg)])
    np.savez_compressed(os.path.join(HERE, f'transition_{tag}.npz'), **out)
    print('wrote', tag)


def op_vectors(ref):
    out = {}
    torch.manual_seed(11)
    n, C = 14, 2
    # Sobolev kernels (utils/functions.py:24-49)
    for s in (1, 2, 3):
        k, ks = ref.functions.Sobolev_kernel_1D(s, 0.5)
        out[f'sobolev_s{s}'], out[f'sobolev_sqrt_s{s}'] = k, ks
    # separable_conv_3D with the trainer's S dict (trainer/trainer.py:568-583)
    v = torch.randn(C, 3, n, n, n)
    S = torch.from_numpy(ref.functions.Sobolev_kernel_1D(3, 0.5)[0]).float().unsqueeze(0)
    S = torch.stack((S, S, S), 0)
    Sd = {'x': S.unsqueeze(2).unsqueeze(2), 'y': S.unsqueeze(2).unsqueeze(4), 'z': S.unsqueeze(3).unsqueeze(4)}
    out['smooth_in'] = np32(v)
    out['smooth_out'] = np32(ref.util.separable_conv_3D(v, Sd['x'], Sd['y'], Sd['z'], (3,) * 6))
    # SVF_3D (utils/transformation.py:51-76), fp32 and fp64 with gradient of sum(T * G)
    vs = smooth_field((C, 3, n, n, n), 3.0, 5)
    G = torch.randn(C, 3, n, n, n)
    svf = ref.transformation.SVF_3D((n, n, n))
    v32 = vs.clone().requires_grad_(True)
    T, disp = svf(v32)
    (disp * G).sum().backward()
    out['svf_v'], out['svf_G'] = np32(vs), np32(G)
    out['svf_T'], out['svf_disp'], out['svf_grad'] = np32(T), np32(disp), np32(v32.grad)
    svf64 = ref.transformation.SVF_3D((n, n, n)).double()
    svf64.identity_grid.data = ref.util.init_identity_grid_3D((n, n, n)).double() * 0 + \
        torch.stack(torch.meshgrid(*[torch.linspace(-1, 1, n, dtype=torch.float64)] * 3, indexing='ij')[::-1], -1).unsqueeze(0)
    v64 = vs.double().requires_grad_(True)
    T64, disp64 = svf64(v64)
    (disp64 * G.double()).sum().backward()
    out['svf_disp_f64'], out['svf_grad_f64'] = np32(disp64), np32(v64.grad)
    # RegistrationModule (utils/registration.py): trilinear + nearest int16 / bool
    regm = ref.registration.RegistrationModule()
    im = torch.rand(C, 1, n, n, n)
    seg = (torch.rand(C, 1, n, n, n) * 60).short()
    msk = torch.rand(C, 1, n, n, n) > 0.5
    Tw = T.detach().clone()
    k = torch.arange(n, dtype=torch.float32)
    Tw[0, 0, 0, 0, :] = 2.0 * (k + 0.5) / (n - 1) - 1.0   # exact half-voxel positions: round half to even
    Tw[1, :, 1, 1, :4] = torch.tensor([-1.0, 1.0, -7.0, 9.0])
    out['warp_T'], out['warp_im'], out['warp_seg'], out['warp_mask'] = np32(Tw), np32(im), np32(seg), np32(msk)
    out['warp_im_out'] = np32(regm(im, Tw))
    out['warp_seg_out'] = np32(regm(seg, Tw))
    out['warp_mask_out'] = np32(regm(msk, Tw))
    # GradientOperator (utils/diff_op.py:78-96) and det J (utils/util.py:72-91)
    op = ref.diff_op.GradientOperator()
    out['nabla_v'] = np32(op(v))
    op2 = ref.diff_op.GradientOperator()
    nabT = op2(T.detach(), transformation=True)
    out['nabla_T'] = np32(nabT)
    out['det_J'] = np32(ref.util.calc_det_J(nabT))
    # GMM.map / log_pdf (model/loss.py:87-111), VD (utils/util.py:330-347,446-485)
    for s in (1, 2):
        gmm = ref.loss.GMM(4, s)
        gmm.init_parameters(torch.tensor(0.7))
        with torch.no_grad():
            gmm.logits.copy_(torch.tensor([0.1, -0.2, 0.3, 0.0]))
        a, b = torch.rand(1, 1, n, n, n), torch.rand(1, 1, n, n, n)
        a = F.avg_pool3d(F.pad(a, (1,) * 6, mode='replicate'), 3, 1) + 0.05 * a
        b = F.avg_pool3d(F.pad(b, (1,) * 6, mode='replicate'), 3, 1) + 0.05 * b
        z = gmm.map(a, b).detach()
        mask = torch.rand(1, 1, n, n, n) > 0.3
        out[f'lcc_s{s}_F'], out[f'lcc_s{s}_M'], out[f'lcc_s{s}_z'], out[f'lcc_s{s}_mask'] = np32(a), np32(b), np32(z), np32(mask)
        out[f'gmm_s{s}_log_std'], out[f'gmm_s{s}_logits'] = np32(gmm.log_std), np32(gmm.logits)
        out[f'gmm_s{s}_log_pdf'] = np32(gmm.log_pdf(z[mask]))
        r = ref.util.rescale_residuals(z, mask, gmm)
        out[f'vd_s{s}_rescaled'] = np32(r)
        out[f'vd_s{s}_alpha'] = np.array(float(ref.util.calc_VD_factor(r, mask)))
    # RegLoss (model/loss.py:152-312)
    for name, cls in (('l2', ref.loss.RegLoss_L2), ('lognormal', ref.loss.RegLoss_LogNormal)):
        reg = cls(w_reg=1.4, diff_op='GradientOperator', dims=(n, n, n), learnable=False)
        loss, log_y = reg(v)
        out[f'reg_{name}_loss'], out[f'reg_{name}_log_y'] = np32(loss), np32(log_y)
    lg = ref.loss.RegLoss_LogNormal(w_reg=1.6, diff_op='GradientOperator', dims=(128, 128, 128), learnable=True)
    out['lognormal_init_128'] = np.array([float(lg.loc), float(lg.log_scale)])
    # posterior statistics (utils/util.py:114-120)
    samples = torch.randn(9, 3, 6, 6, 6)
    mean, std = ref.util.calc_posterior_statistics(samples, device='cpu')
    out['post_samples'], out['post_mean'], out['post_std'] = np32(samples), np32(mean), np32(std)
    np.savez_compressed(os.path.join(HERE, 'ops.npz'), **out)
    print('wrote ops')


def ffd_vectors(ref):
    """Cubic_B_spline_FFD_3D / SVFFD_3D (utils/transformation.py:79-164), get_control_grid_size (utils/util.py:61-69)"""
    out = {}
    torch.manual_seed(17)
    T = ref.transformation
    for tag, dims, cps in (('a', (16, 16, 16), (4, 4, 4)), ('b', (13, 13, 13), (2, 3, 5))):
        g = ref.util.get_control_grid_size(dims, cps)
        cp = torch.randn(2, 3, *g)
        G = torch.randn(2, 3, *dims)
        ffd = T.Cubic_B_spline_FFD_3D(dims, cps)
        cp32 = cp.clone().requires_grad_(True)
        dense = ffd(cp32)
        (dense * G).sum().backward()
        out[f'{tag}_dims'], out[f'{tag}_cps'], out[f'{tag}_grid'] = np.array(dims), np.array(cps), np.array(g)
        out[f'{tag}_cp'], out[f'{tag}_G'] = np32(cp), np32(G)
        out[f'{tag}_dense'], out[f'{tag}_grad'] = np32(dense), np32(cp32.grad)
        for i, s in enumerate(cps):
            out[f'{tag}_kernel{i}'] = np32(T.B_spline_1D_kernel(s))
        # one un-cropped axis of it: conv1D(transpose=True) along H
        out[f'{tag}_conv1d_dim3'] = np32(T.conv1D(cp, ffd.kernels[1], dim=3, stride=cps[1], padding=ffd.padding[1],
                                                  transpose=True))
    # SVFFD_3D: smooth control-point velocities of a few voxels, fp32 and fp64 with the gradient of sum(disp * G)
    dims, cps = (16, 16, 16), (4, 4, 4)
    g = ref.util.get_control_grid_size(dims, cps)
    cp = 2.0 * torch.randn(2, 3, *g)
    G = torch.randn(2, 3, *dims)
    m = T.SVFFD_3D(dims, cps)
    # the gradient w.r.t. the dense velocity field is kept too: trilinear kink flips (SURVEY surprise 9) are isolated
    # voxels there, whereas one flipped voxel reaches the 64 control points around it
    def run(module, cp_in, G_in):
        dense = []
        def keep(mod, inp, res):
            res.retain_grad()
            dense.append(res)

        hook = module.cubic_B_spline_FFD.register_forward_hook(keep)
        Tr, disp = module(cp_in)
        (disp * G_in).sum().backward()
        hook.remove()
        return Tr, disp, dense[0].grad

    cp32 = cp.clone().requires_grad_(True)
    Tr, disp, g_dense = run(m, cp32, G)
    out['svffd_cp'], out['svffd_G'] = np32(cp), np32(G)
    out['svffd_T'], out['svffd_disp'], out['svffd_grad'] = np32(Tr), np32(disp), np32(cp32.grad)
    out['svffd_grad_dense'] = np32(g_dense)
    m64 = T.SVFFD_3D(dims, cps).double()
    cp64 = cp.double().requires_grad_(True)
    _, disp64, g_dense64 = run(m64, cp64, G.double())
    out['svffd_disp_f64'], out['svffd_grad_f64'] = np32(disp64), np32(cp64.grad)
    out['svffd_grad_dense_f64'] = np32(g_dense64)
    np.savez_compressed(os.path.join(HERE, 'ffd.npz'), **out)
    print('wrote ffd')


def eval_vectors(ref):
    """per-sample evaluation functions of the reference: calc_DSC_GPU (utils/util.py:123-148, runs on
    CPU tensors as well), calc_no_non_diffeomorphic_voxels (:209-212) on a transformation that folds, calc_norm (:215-225)"""
    out = {}
    torch.manual_seed(21)
    n, C = 24, 3
    fixed, moving, _ = make_pair(n)
    svf = ref.transformation.SVF_3D((n, n, n))
    regm = ref.registration.RegistrationModule()
    v = smooth_field((C, 3, n, n, n), 2.5, 9)
    v[2] *= 6.0                                   # a deformation large enough to fold
    T, disp = svf(v)
    seg_w = regm(moving['seg'].expand(C, -1, -1, -1, -1).contiguous(), T)
    labels = sorted(int(x) for x in torch.unique(fixed['seg']) if int(x) != 0)
    structures = {f's{l}': l for l in labels}
    out['labels'] = np.array(labels, dtype=np.int64)
    out['seg_fixed'], out['seg_moving_warped'] = np32(fixed['seg']), np32(seg_w)
    out['DSC'] = ref.util.calc_DSC_GPU(C, fixed['seg'].expand(C, -1, -1, -1, -1), seg_w, structures)
    counts, log_det_J = ref.util.calc_no_non_diffeomorphic_voxels(T, ref.diff_op.GradientOperator())
    assert counts[2] > 0 and counts[0] == 0, counts
    out['T'], out['no_non_diffeomorphic_voxels'], out['log_det_J'] = np32(T), np.asarray(counts), np32(log_det_J)
    out['disp'], out['disp_norm'] = np32(disp), np32(ref.util.calc_norm(disp))
    np.savez_compressed(os.path.join(HERE, 'eval.npz'), **out)
    print('wrote eval', counts, out['DSC'].shape)


if __name__ == '__main__':
    ref = ref_import.load()
    which = sys.argv[1:] or ['ops', 'transitions', 'ffd', 'eval']
    if 'ops' in which:
        op_vectors(ref)
    if 'transitions' in which:
        transition_vectors(ref, 12, 2, 'RegLoss_LogNormal', True, 1.6, 'lcc_lognormal')
        transition_vectors(ref, 12, 2, 'RegLoss_L2', True, 1.4, 'lcc_l2')
    if 'eval' in which:
        eval_vectors(ref)
    if 'ffd' in which:
        ffd_vectors(ref)
        transition_vectors(ref, 16, 2, 'RegLoss_LogNormal', True, 1.6, 'svffd_lognormal', cps=(4, 4, 4))
