"""
Cubic B-spline FFD / SVFFD (reference utils/transformation.py:79-164, utils/util.py:61-69):
 - CPU: the oracle's restatement and the library's device arithmetic (tests/host_emul) against golden vectors produced by
   the UNMODIFIED reference (tests/golden/ffd.npz, script tests/golden/make_golden.py), and the reference's own shape
   tests restated (reference tests/test_utils.py:75-99);
 - GPU: the CUDA kernels through the C ABI (ops) and through the drop-in modules against the same vectors and the oracle,
   and at 128^3 through the adjoint identity <A x, y> = <x, A^T y>.
"""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import sgld_oracle as O
from tests.util import grad_ok, rel

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ffd.npz')
DEV = 'cuda:0'


@pytest.fixture(scope='module')
def gold():
    return {k: torch.from_numpy(np.asarray(v)) for k, v in np.load(GOLD).items()}


def case(gold, tag):
    dims, cps = tuple(int(x) for x in gold[f'{tag}_dims']), tuple(int(x) for x in gold[f'{tag}_cps'])
    return dims, cps, tuple(int(x) for x in gold[f'{tag}_grid'])


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# ---------------------------------------------------------------------------------------------------------------------
# CPU: oracle and device arithmetic vs the reference's vectors
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('tag', ['a', 'b'])
def test_oracle_ffd(gold, tag):
    dims, cps, grid = case(gold, tag)
    assert O.control_grid_size(dims, cps) == grid
    for i, s in enumerate(cps):
        assert torch.equal(O.bspline_taps(s), gold[f'{tag}_kernel{i}'])   # bit-exact taps
    cp = gold[f'{tag}_cp'].clone().requires_grad_(True)
    dense = O.ffd_dense(cp, dims, cps)
    assert dense.shape == gold[f'{tag}_dense'].shape and rel(dense, gold[f'{tag}_dense']) < 1e-6
    g, = torch.autograd.grad((dense * gold[f'{tag}_G']).sum(), cp)
    assert rel(g, gold[f'{tag}_grad']) < 1e-6


def test_oracle_svffd(gold):
    dims, cps = (16, 16, 16), (4, 4, 4)
    cp = gold['svffd_cp'].clone().requires_grad_(True)
    T, disp = O.svffd_exp_aten(cp, dims, cps)
    assert rel(T, gold['svffd_T']) < 1e-6 and rel(disp, gold['svffd_disp']) < 2e-6
    g, = torch.autograd.grad((disp * gold['svffd_G']).sum(), cp)
    assert rel(g, gold['svffd_grad']) < 1e-5
    cp64 = gold['svffd_cp'].double().requires_grad_(True)
    _, d64 = O.svffd_exp_aten(cp64, dims, cps)
    g64, = torch.autograd.grad((d64 * gold['svffd_G'].double()).sum(), cp64)
    assert rel(d64, gold['svffd_disp_f64']) < 1e-6 and rel(g64, gold['svffd_grad_f64']) < 1e-4   # fp32 taps in the reference


def test_reference_shape_tests(gold):
    """reference tests/test_utils.py:75-99 restated on the oracle: dims 64^3, cps 4 -> dense (1,3,64,64,64)"""
    dims, cps = (64,) * 3, (4,) * 3
    grid = O.control_grid_size(dims, cps)
    assert grid == (19, 19, 19)
    v = torch.randn(1, 3, *grid)
    assert O.ffd_dense(v, dims, cps).shape == (1, 3, *dims)


@pytest.mark.parametrize('tag', ['a', 'b'])
def test_device_arithmetic_on_host(built, gold, tag):
    emul = ctypes.CDLL(built['emul'])
    dims, cps, grid = case(gold, tag)
    ks = [np.ascontiguousarray(gold[f'{tag}_kernel{i}'].numpy()) for i in range(3)]
    cp = np.ascontiguousarray(gold[f'{tag}_cp'].numpy())
    C = cp.shape[0]
    dense = np.zeros((C, 3, *dims), np.float32)
    emul.emul_ffd(P(cp), P(dense), 0, P(ks[0]), P(ks[1]), P(ks[2]), *cps, C, *grid, *dims)
    assert rel(dense, gold[f'{tag}_dense']) < 1e-6
    G = np.ascontiguousarray(gold[f'{tag}_G'].numpy())
    g_cp = np.zeros_like(cp)
    emul.emul_ffd(P(G), P(g_cp), 1, P(ks[0]), P(ks[1]), P(ks[2]), *cps, C, *grid, *dims)
    assert rel(g_cp, gold[f'{tag}_grad']) < 1e-6
    # one un-cropped axis = conv1D(transpose=True) along H
    full = gold[f'{tag}_conv1d_dim3']
    out = np.zeros(tuple(full.shape), np.float32)
    emul.emul_bspline_axis(P(cp), P(out), 0, ctypes.c_longlong(C * 3 * grid[0]), grid[1], out.shape[3],
                           ctypes.c_longlong(grid[2]), P(ks[1]), cps[1], 0)
    assert rel(out, full) < 1e-6


@pytest.mark.parametrize('n,cps,C', [(10, (3, 2, 4), 2), (6, (1, 1, 1), 2), (18, (8, 5, 7), 4)])
def test_device_arithmetic_groups_crossing_rows(built, n, cps, C):
    """sizes where a thread's four consecutive outputs straddle row and plane boundaries (rows not a multiple of four
    while the total is), against the oracle"""
    emul = ctypes.CDLL(built['emul'])
    dims = (n,) * 3
    grid = O.control_grid_size(dims, cps)
    assert (C * 3 * n ** 3) % 4 == 0 and n % 4 != 0
    ks = [np.ascontiguousarray(O.bspline_taps(s).numpy()) for s in cps]
    gen = torch.Generator().manual_seed(5)
    cp = torch.randn(C, 3, *grid, generator=gen)
    G = torch.randn(C, 3, *dims, generator=gen)
    cp64 = cp.double().requires_grad_(True)
    ref = O.ffd_dense(cp64, dims, cps)
    g64, = torch.autograd.grad((ref * G.double()).sum(), cp64)
    dense, g_cp = np.zeros((C, 3, *dims), np.float32), np.zeros((C, 3, *grid), np.float32)
    emul.emul_ffd(P(cp.numpy()), P(dense), 0, P(ks[0]), P(ks[1]), P(ks[2]), *cps, C, *grid, *dims)
    emul.emul_ffd(P(G.numpy()), P(g_cp), 1, P(ks[0]), P(ks[1]), P(ks[2]), *cps, C, *grid, *dims)
    assert rel(dense, ref) < 1e-6 and rel(g_cp, g64) < 1e-6


def test_device_arithmetic_random_shapes(built):
    """seeded sweep over volume sizes, spacings 1..8 per axis and chain counts: forward against the oracle, and the
    adjoint identity <A x, y> = <x, A^T y> between the two directions of the device arithmetic"""
    emul = ctypes.CDLL(built['emul'])
    rng = np.random.default_rng(2024)
    for _ in range(16):
        n = int(rng.integers(4, 23))
        cps = tuple(int(c) for c in rng.integers(1, 9, size=3))
        C = int(rng.integers(1, 4))
        dims = (n,) * 3
        grid = O.control_grid_size(dims, cps)
        ks = [np.ascontiguousarray(O.bspline_taps(s).numpy()) for s in cps]
        x = rng.standard_normal((C, 3, *grid)).astype(np.float32)
        y = rng.standard_normal((C, 3, *dims)).astype(np.float32)
        Ax, Aty = np.zeros_like(y), np.zeros_like(x)
        emul.emul_ffd(P(x), P(Ax), 0, P(ks[0]), P(ks[1]), P(ks[2]), *cps, C, *grid, *dims)
        emul.emul_ffd(P(y), P(Aty), 1, P(ks[0]), P(ks[1]), P(ks[2]), *cps, C, *grid, *dims)
        assert rel(Ax, O.ffd_dense(torch.from_numpy(x).double(), dims, cps)) < 1e-6, (n, cps, C)
        lhs, rhs = float((Ax.astype(np.float64) * y).sum()), float((x.astype(np.float64) * Aty).sum())
        assert abs(lhs - rhs) <= 1e-5 * float(np.linalg.norm(Ax.astype(np.float64))) + 1e-9, (n, cps, C)


def test_oracle_conv1d_axis(gold):
    for tag in 'ab':
        _, cps, _ = case(gold, tag)
        assert rel(O.bspline_axis(gold[f'{tag}_cp'], 3, cps[1]), gold[f'{tag}_conv1d_dim3']) < 1e-6


def test_device_arithmetic_svffd_on_host(built, gold):
    """FFD -> scaling and squaring -> adjoints, all in the library's device arithmetic, against the reference's SVFFD_3D"""
    emul = ctypes.CDLL(built['emul'])
    dims, cps, steps = (16, 16, 16), (4, 4, 4), 12
    grid = O.control_grid_size(dims, cps)
    k = np.ascontiguousarray(O.bspline_taps(4).numpy())
    cp = np.ascontiguousarray(gold['svffd_cp'].numpy())
    C, n = cp.shape[0], dims[0]
    v = np.zeros((C, 3, *dims), np.float32)
    emul.emul_ffd(P(cp), P(v), 0, P(k), P(k), P(k), *cps, C, *grid, *dims)
    hist, maxabs = np.zeros((steps, C, 3, *dims), np.float32), np.zeros(steps, np.float32)
    emul.emul_svf_fwd(P(v), P(hist), P(maxabs), steps, C, n, n, n)
    assert rel(hist[-1], gold['svffd_disp']) < 1e-5 and rel(hist[-1], gold['svffd_disp_f64']) < 1e-5
    # gradient in two stages: trilinear kink flips (SURVEY surprise 9) are isolated voxels of the DENSE gradient -- one
    # flipped voxel reaches the 64 control points around it -- so the acceptance rule is applied there, and the (linear)
    # FFD adjoint is checked exactly on the reference's own dense gradient
    g_v, g_cp = np.zeros_like(v), np.zeros_like(cp)
    emul.emul_svf_bwd(P(v), P(hist), P(maxabs), P(np.ascontiguousarray(gold['svffd_G'].numpy())), P(g_v), steps, 0,
                      C, n, n, n)
    assert grad_ok(g_v, gold['svffd_grad_dense'], gold['svffd_grad_dense_f64'], 'SVFFD dense gradient (host emulation)')
    emul.emul_ffd(P(np.ascontiguousarray(gold['svffd_grad_dense'].numpy())), P(g_cp), 1, P(k), P(k), P(k), *cps, C,
                  *grid, *dims)
    assert rel(g_cp, gold['svffd_grad']) < 1e-6


def test_module_interface_without_gpu():
    """constructor attributes of the reference (utils/transformation.py:132-147) and the loud failure without CUDA"""
    import irsgmcmc_b200.utils as U
    assert U.get_control_grid_size((128,) * 3, (4,) * 3) == (35, 35, 35)
    m = U.SVFFD_3D((16,) * 3, (4,) * 3)
    ffd = m.cubic_B_spline_FFD
    assert [tuple(k.shape) for k in ffd.kernels] == [(15,)] * 3 and ffd.padding == [7, 7, 7]
    assert all(not k.requires_grad for k in ffd.kernels) and m.SVF_3D.no_steps == 12
    assert torch.equal(ffd.kernels[0].data, O.bspline_taps(4))
    assert U.cubic_B_spline_1D_value(0) == 2.0 / 3.0 and U.cubic_B_spline_1D_value(-2.5) == 0
    with pytest.raises(RuntimeError, match='no CPU implementation'):
        m(torch.zeros(1, 3, 7, 7, 7))


# ---------------------------------------------------------------------------------------------------------------------
# GPU: CUDA kernels through the C ABI and the drop-in modules
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize('tag', ['a', 'b'])
def test_gpu_ffd_ops(built, gold, tag):
    from irsgmcmc_b200 import ops
    dims, cps, grid = case(gold, tag)
    ks = [tuple(float(x) for x in gold[f'{tag}_kernel{i}']) for i in range(3)]
    cp = gold[f'{tag}_cp'].to(DEV)
    dense = ops.ffd_fwd(cp, ks, cps, dims)
    assert rel(dense, gold[f'{tag}_dense']) < 1e-6
    g_cp = ops.ffd_bwd(gold[f'{tag}_G'].to(DEV), ks, cps, grid)
    assert rel(g_cp, gold[f'{tag}_grad']) < 1e-6
    full = ops.bspline_axis(cp, ks[1], 3, cps[1])
    assert full.shape == gold[f'{tag}_conv1d_dim3'].shape and rel(full, gold[f'{tag}_conv1d_dim3']) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize('tag', ['a', 'b'])
def test_gpu_ffd_module_autograd(built, gold, tag):
    import irsgmcmc_b200.utils as U
    dims, cps, grid = case(gold, tag)
    ffd = U.Cubic_B_spline_FFD_3D(dims, cps).to(DEV)
    cp = gold[f'{tag}_cp'].to(DEV).requires_grad_(True)
    dense = ffd(cp)
    (dense * gold[f'{tag}_G'].to(DEV)).sum().backward()
    assert rel(dense, gold[f'{tag}_dense']) < 1e-6 and rel(cp.grad, gold[f'{tag}_grad']) < 1e-6
    # conv1D with the reference's arguments (utils/transformation.py:149), and what it does not cover
    x = gold[f'{tag}_cp'].to(DEV).requires_grad_(True)
    y = U.conv1D(x, ffd.kernels[1], dim=3, stride=cps[1], padding=ffd.padding[1], transpose=True)
    assert rel(y, gold[f'{tag}_conv1d_dim3']) < 1e-6
    w = torch.randn_like(y)
    (y * w).sum().backward()
    x64 = gold[f'{tag}_cp'].double().requires_grad_(True)
    full = O.bspline_axis(x64, 3, cps[1])
    g64, = torch.autograd.grad((full * w.double().cpu()).sum(), x64)
    assert rel(y, full) < 1e-6 and rel(x.grad, g64) < 1e-6
    with pytest.raises(NotImplementedError):
        U.conv1D(x, ffd.kernels[1], dim=3, stride=cps[1], padding=0, transpose=False)


@pytest.mark.gpu
def test_gpu_svffd_module(built, gold):
    import irsgmcmc_b200.utils as U
    dims, cps = (16, 16, 16), (4, 4, 4)
    m = U.SVFFD_3D(dims, cps).to(DEV)
    cp = gold['svffd_cp'].to(DEV).requires_grad_(True)
    T, disp = m(cp)
    assert T.shape == (2, 3, *dims) and disp.shape == (2, 3, *dims)
    assert rel(T, gold['svffd_T']) < 1e-5 and rel(disp, gold['svffd_disp']) < 1e-5
    assert rel(disp, gold['svffd_disp_f64']) < 1e-5
    # gradient in two stages (see test_device_arithmetic_svffd_on_host): kink-flip acceptance on the dense gradient, the
    # linear FFD adjoint exactly; then the module's end-to-end gradient = FFD adjoint of its own dense gradient
    dense = m.cubic_B_spline_FFD(cp.detach()).requires_grad_(True)
    _, disp2 = m.SVF_3D(dense)
    (disp2 * gold['svffd_G'].to(DEV)).sum().backward()
    assert grad_ok(dense.grad, gold['svffd_grad_dense'], gold['svffd_grad_dense_f64'], 'SVFFD_3D dense gradient')
    ffd = m.cubic_B_spline_FFD
    x = gold['svffd_cp'].to(DEV).requires_grad_(True)
    (ffd(x) * gold['svffd_grad_dense'].to(DEV)).sum().backward()
    assert rel(x.grad, gold['svffd_grad']) < 1e-6
    (disp * gold['svffd_G'].to(DEV)).sum().backward()
    y = gold['svffd_cp'].to(DEV).requires_grad_(True)
    (ffd(y) * dense.grad).sum().backward()
    assert rel(cp.grad, y.grad) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize('n,cps', [(128, (4, 4, 4)), (96, (2, 3, 8))])
def test_gpu_ffd_fullsize_adjoint_identity(built, n, cps):
    """size-independent properties at the benchmark volume: <A x, y> = <x, A^T y>, partition of unity, linearity"""
    from irsgmcmc_b200 import ops
    dims = (n,) * 3
    grid = O.control_grid_size(dims, cps)
    ks = [tuple(float(x) for x in O.bspline_taps(s)) for s in cps]
    g = torch.Generator(device=DEV).manual_seed(3)
    x = torch.randn(2, 3, *grid, device=DEV, generator=g)
    y = torch.randn(2, 3, *dims, device=DEV, generator=g)
    Ax, Aty = ops.ffd_fwd(x, ks, cps, dims), ops.ffd_bwd(y, ks, cps, grid)
    lhs, rhs = float((Ax.double() * y.double()).sum()), float((x.double() * Aty.double()).sum())
    # y has unit variance: rounding of A x (1e-7 relative per element) moves <A x, y> by about 1e-7 |A x|
    assert abs(lhs - rhs) <= 1e-5 * float(Ax.double().norm()) and abs(lhs) > 1e-3 * float(Ax.double().norm())
    ones = ops.ffd_fwd(torch.ones(1, 3, *grid, device=DEV), ks, cps, dims)
    assert float((ones - 1).abs().max()) < 1e-6          # B-splines sum to one
    x2 = torch.randn(2, 3, *grid, device=DEV, generator=g)
    assert rel(ops.ffd_fwd(x + 2 * x2, ks, cps, dims), Ax + 2 * ops.ffd_fwd(x2, ks, cps, dims)) < 1e-6
    # a 16^3 corner of the volume against the oracle (the support of a voxel is local)
    sub = O.ffd_dense(x.cpu(), dims, cps)[..., :16, :16, :16]
    assert rel(Ax[..., :16, :16, :16], sub) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize('n,cps,C', [(128, (4, 4, 4), 1), (40, (2, 6, 2), 3), (33, (3, 1, 6), 2), (20, (8, 8, 8), 1)])
def test_gpu_ffd_fast_kernels_match_generic(built, n, cps, C, monkeypatch):
    """the kernels of the contiguous axis (rows marched with the table entry in registers; rows staged in shared memory
    with the bank-conflict skew) against the generic axis kernel: same terms in the same order"""
    from irsgmcmc_b200 import ops
    dims = (n,) * 3
    grid = O.control_grid_size(dims, cps)
    ks = [tuple(float(x) for x in O.bspline_taps(s)) for s in cps]
    g = torch.Generator(device=DEV).manual_seed(11)
    x = torch.randn(C, 3, *grid, device=DEV, generator=g)
    y = torch.randn(C, 3, *dims, device=DEV, generator=g)
    fast = ops.ffd_fwd(x, ks, cps, dims), ops.ffd_bwd(y, ks, cps, grid)
    monkeypatch.setenv('IRS_FFD_GENERIC', '1')
    slow = ops.ffd_fwd(x, ks, cps, dims), ops.ffd_bwd(y, ks, cps, grid)
    torch.cuda.synchronize()
    for a, b in zip(fast, slow):
        assert bool(torch.isfinite(a).all()) and float((a - b).abs().max()) <= 1e-6 * float(b.abs().max())
