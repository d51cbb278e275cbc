"""the C-ABI shared library loads and exports every symbol include/irsgmcmc.h declares; host logic without a GPU"""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'irsgmcmc.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(irs_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(built['lib'])
    names = declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), name


def test_python_binding_covers_the_header(built):
    from irsgmcmc_b200 import _lib
    _lib.load()
    assert set(declared_symbols()) == set(_lib.SYMBOLS)


def test_version_and_error_strings(built):
    from irsgmcmc_b200 import _lib
    lib = _lib.load()
    assert lib.irs_abi_version() == 4
    assert lib.irs_error_string(0) == b'ok'
    assert b'bad argument' in lib.irs_error_string(-1) and b'unsupported' in lib.irs_error_string(-2)


def test_argument_validation_without_gpu(built):
    """bad arguments are rejected before any CUDA call"""
    from irsgmcmc_b200 import _lib
    lib = _lib.load()
    assert lib.irs_warp3d_fwd(None, 0, None, None, 0.0, None, 1, 8, 8, 8, None) == -1
    assert lib.irs_svf_exp_fwd(None, None, None, 12, 1, 8, 8, 8, None) == -1
    assert lib.irs_lcc_normalise(None, 9, None, None, None, 1, 8, 8, 8, None) == -1
    cfg = _lib.SgldConfig()
    assert lib.irs_sgld_step(ctypes.byref(cfg), None, None) == -1
    assert lib.irs_sgld_launches_per_step(ctypes.byref(cfg)) == -1
    assert lib.irs_svf_hist_floats(2, 4, 5, 6, 12) == 12 * 2 * 3 * 4 * 5 * 6
    # 12 global maxima + 12 cell maps of 2 chains x (1 x 1 x 1) cells of 32 x 8 x 8 voxels
    assert lib.irs_svf_maxabs_floats(2, 4, 5, 6, 12) == 12 * (1 + 2 * 1)
    assert lib.irs_svf_maxabs_floats(1, 128, 128, 128, 12) == 12 * (1 + 4 * 16 * 16)
    # cubic B-spline FFD: workspace = the two intermediates of the axis passes; null pointers / spacings out of range
    assert lib.irs_ffd_work_floats(2, 35, 35, 35, 128, 128, 128) == 2 * 3 * (128 * 35 * 35 + 128 * 128 * 35)
    assert lib.irs_ffd_work_floats(0, 35, 35, 35, 128, 128, 128) == 0
    k = _lib.host_floats([0.0] * 35)
    assert lib.irs_ffd_fwd(None, k, k, k, 4, 4, 4, None, None, 1, 7, 7, 7, 16, 16, 16, None) == -1
    assert lib.irs_ffd_bwd(None, k, k, k, 4, 4, 4, None, None, 1, 7, 7, 7, 16, 16, 16, None) == -1
    assert lib.irs_bspline_axis(None, None, 0, 6, 7, 16, 49, k, 4, 4, None) == -1
    cfg = _lib.SgldConfig()
    cfg.C, cfg.D, cfg.H, cfg.W, cfg.K, cfg.lcc_s, cfg.svf_steps, cfg.n_mask = 1, 16, 16, 16, 4, 2, 12, 100.0
    assert lib.irs_sgld_launches_per_step(ctypes.byref(cfg)) > 0
    plain = lib.irs_sgld_launches_per_step(ctypes.byref(cfg))
    for a in range(3):
        cfg.ffd_cps[a], cfg.ffd_grid[a] = 4, 7
    assert lib.irs_sgld_launches_per_step(ctypes.byref(cfg)) == plain + 7   # 6 FFD passes + the stand-alone energy kernel
    cfg.ffd_grid[1] = 5   # (5 - 1) * 4 + 1 = 17 elements cannot hold the crop [4, 20)
    assert lib.irs_sgld_launches_per_step(ctypes.byref(cfg)) == -1
    cfg.ffd_grid[1], cfg.ffd_cps[2] = 7, 9
    assert lib.irs_sgld_launches_per_step(ctypes.byref(cfg)) == -1


def test_config_struct_layout_matches_c(built):
    """sizeof(irs_sgld_config / irs_sgld_buffers) as the C compiler sees them"""
    import subprocess
    import tempfile
    src = ('#include <stdio.h>\n#include <stddef.h>\n#include "irsgmcmc.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu", '
           'sizeof(irs_sgld_config), sizeof(irs_sgld_buffers), offsetof(irs_sgld_config, seed), '
           'offsetof(irs_sgld_config, ffd_cps), offsetof(irs_sgld_config, ffd_grid), offsetof(irs_sgld_config, ffd_kernel), '
           'offsetof(irs_sgld_buffers, ffd_dense));return 0;}')
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, 't.c'), 'w') as f:
            f.write(src)
        subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), os.path.join(d, 't.c'), '-o', os.path.join(d, 't')])
        a, b, *offs = [int(x) for x in subprocess.check_output([os.path.join(d, 't')]).decode().split()]
    from irsgmcmc_b200 import _lib
    assert a == ctypes.sizeof(_lib.SgldConfig) and b == ctypes.sizeof(_lib.SgldBuffers)
    C, B = _lib.SgldConfig, _lib.SgldBuffers
    assert offs == [C.seed.offset, C.ffd_cps.offset, C.ffd_grid.offset, C.ffd_kernel.offset, B.ffd_dense.offset]


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the behaviour without a GPU')
def test_product_fails_loudly_without_cuda(built):
    from irsgmcmc_b200 import ops
    from irsgmcmc_b200.sampler import SGLDSampler
    from irsgmcmc_b200.data_loader.synthetic import make_pair
    with pytest.raises(RuntimeError):
        ops.warp3d(torch.rand(1, 1, 8, 8, 8), torch.rand(1, 3, 8, 8, 8))
    fixed, moving, _ = make_pair(8)
    with pytest.raises(RuntimeError):
        SGLDSampler(fixed, moving, 1)


def test_no_eager_torch_fallbacks_in_the_product():
    """the volume-sized helpers either launch kernels or raise -- no silent PyTorch path"""
    import irsgmcmc_b200.model.loss as M
    import irsgmcmc_b200.utils.util as U
    from irsgmcmc_b200.utils.diff_op import DifferentialOperator, GradientOperator
    seg = torch.zeros(2, 1, 4, 4, 4, dtype=torch.int16)
    with pytest.raises(NotImplementedError):
        U.calc_DSC_GPU(2, seg, seg, {'a': 10})                     # CPU tensors
    with pytest.raises(NotImplementedError):
        U.calc_no_non_diffeomorphic_voxels(torch.zeros(1, 3, 4, 4, 4), GradientOperator())   # CPU tensor
    with pytest.raises(NotImplementedError):
        U.calc_no_non_diffeomorphic_voxels(torch.zeros(1, 3, 4, 4, 4), DifferentialOperator())
    with pytest.raises(NotImplementedError):
        U.calc_metrics(seg, seg, {'a': 10}, (1, 1, 1), GPU=False)
    with pytest.raises(NotImplementedError):
        M.GMM(4, 2).log_pdf_VD(torch.zeros(3, 4))
    with pytest.raises(NotImplementedError):
        M.RegLoss_L2(1.4, diff_op=None, dims=(4, 4, 4))(torch.zeros(1, 3, 4, 4, 4))   # identity operator: no kernel


def test_product_does_not_import_the_oracle():
    import subprocess
    import sys
    code = ('import sys; sys.path.insert(0, %r); import irsgmcmc_b200.sampler, irsgmcmc_b200.trainer, irsgmcmc_b200.utils, '
            'irsgmcmc_b200.model, irsgmcmc_b200.optimizers; '
            'bad = [m for m in sys.modules if m == "oracle" or m.startswith("oracle.")]; assert not bad, bad') % ROOT
    subprocess.check_call([sys.executable, '-c', code])
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'irsgmcmc_b200')):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, fn)).read()
                assert 'from oracle' not in text and 'import oracle' not in text, fn


def test_synthetic_pair_layout():
    """the reference's data dict layout (data_loader/datasets.py:117,128,135)"""
    from irsgmcmc_b200.data_loader.synthetic import make_pair
    n = 16
    fixed, moving, vp = make_pair(n)
    for d in (fixed, moving):
        assert d['im'].shape == (1, 1, n, n, n) and d['im'].dtype == torch.float32
        assert d['mask'].dtype == torch.bool and d['seg'].dtype == torch.int16
        assert 0.0 <= float(d['im'].min()) and float(d['im'].max()) <= 1.0
    assert set(vp) == {'mu', 'log_var', 'u'} and vp['mu'].shape == (1, 3, n, n, n)
    f2, m2, _ = make_pair(n)
    assert torch.equal(fixed['im'], f2['im']) and torch.equal(moving['im'], m2['im'])


def test_sobolev_kernel():
    from irsgmcmc_b200.utils.functions import Sobolev_kernel_1D
    import numpy as np
    k, ks = Sobolev_kernel_1D(3, 0.5)
    assert np.allclose(k * 96, [1, 4, 15, 56, 15, 4, 1]) and abs(ks.sum() - 1) < 1e-12
    assert np.allclose(Sobolev_kernel_1D(1, 0.5)[0] * 6, [1, 4, 1])


def test_config_from_reference_json():
    from irsgmcmc_b200.trainer import sampler_config_from_json
    cfg = {'data_loss': {'type': 'GMM', 'args': {'no_components': 4, 's': 2}},
           'reg_loss': {'type': 'RegLoss_LogNormal', 'args': {'diff_op': 'GradientOperator', 'w_reg': 1.6, 'learnable': True}},
           'optimizer_SG_MCMC': {'type': 'SGD', 'args': {'lr': 0.4}}, 'Sobolev_grad': {'enabled': True, 's': 3, 'lambda': 0.5},
           'virtual_decimation': True,
           'trainer': {'MCMC_init': 'VI', 'no_chains': 2, 'no_iters_burn_in': 1, 'no_samples_MCMC': 2, 'log_period_MCMC': 1,
                       'uniform_noise': {'enabled': True, 'magnitude': 0.1}}}
    c = sampler_config_from_json(cfg)
    assert (c.data_loss, c.no_components, c.s, c.reg_loss, c.w_reg, c.tau) == ('lcc', 4, 2, 'RegLoss_LogNormal', 1.6, 0.4)
    assert c.transformation == 'SVF_3D' and c.cps is None and c.hyper_mode == 'reference'
    cfg['trainer']['hyper_mode'] = 'frozen'
    assert sampler_config_from_json(cfg).hyper_mode == 'frozen'
    cfg['trainer']['hyper_mode'] = 'per_chain'
    with pytest.raises(NotImplementedError):
        sampler_config_from_json(cfg)
    del cfg['trainer']['hyper_mode']
    cfg['transformation_module'] = {'type': 'SVFFD_3D', 'args': {'cps': [4, 4, 4]}}   # configs/experiment5/config_SVFFD_4.json
    c = sampler_config_from_json(cfg)
    assert c.transformation == 'SVFFD_3D' and c.cps == (4, 4, 4)
    cfg['transformation_module'] = {'type': 'SVFFD_3D', 'args': {}}
    with pytest.raises(ValueError):
        sampler_config_from_json(cfg)
    cfg['transformation_module'] = {'type': 'SVF_2D', 'args': {}}
    with pytest.raises(NotImplementedError):
        sampler_config_from_json(cfg)
    del cfg['transformation_module']
    cfg['reg_loss']['type'] = 'RegLoss_Student'
    with pytest.raises(NotImplementedError):
        sampler_config_from_json(cfg)
