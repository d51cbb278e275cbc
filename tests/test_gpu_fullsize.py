"""
Parity at BASELINE.json's full sizes (128^3, 256^3) through size-independent properties of the path, plus edge cases
(non-cubic and tiny volumes, volumes that are not a multiple of any tile, many chains).
"""
import math

import pytest
import torch

from oracle import sgld_oracle as O
from tests.util import grad_ok, rel, smooth_field

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.fixture(scope='module')
def ops(built):
    from irsgmcmc_b200 import ops
    return ops


@pytest.mark.parametrize('n', [128, 256])
def test_svf_constant_and_zero_velocity(ops, n):
    """exp of a constant velocity is that translation away from the clamped border; the adjoint at v = 0 doubles the
    gradient at every step, so dL/dv = 2^12 / 2^12 * g = g exactly"""
    c = torch.tensor([0.75, -1.25, 0.5], device=DEV).view(1, 3, 1, 1, 1)
    v = c.expand(1, 3, n, n, n).contiguous()
    hist, maxabs = ops.svf_exp_fwd(v, 12)
    m = 4
    inner = hist[-1][..., m:-m, m:-m, m:-m]
    assert torch.equal(inner, c.expand_as(inner).contiguous())
    assert abs(float(maxabs[-1]) - 1.25 / 2) < 1e-6
    v0 = torch.zeros(1, 3, n, n, n, device=DEV)
    hist0, maxabs0 = ops.svf_exp_fwd(v0, 12)
    assert float(hist0.abs().max()) == 0.0
    g = torch.randn(1, 3, n, n, n, device=DEV)
    g_v = ops.svf_exp_bwd(v0, hist0, maxabs0, g)
    assert torch.equal(g_v, g)


@pytest.mark.parametrize('n', [128, 256])
def test_interpolation_transpose_conserves_mass(ops, n):
    """the hat weights of one source sum to one: with a gradient that is constant in space the adjoint's transpose term
    adds exactly that constant wherever no source is clamped, for any velocity -> checked through <g_v, 1> bookkeeping:
    sum over voxels of dL/dv for L = sum(u_n . G0), G0 constant, equals the analytic value of a pure translation field"""
    G0 = torch.tensor([1.0, -2.0, 0.5], device=DEV).view(1, 3, 1, 1, 1)
    v = torch.zeros(1, 3, n, n, n, device=DEV)
    v[:, 0] = 0.3   # uniform translation: every step doubles, nothing is clamped in the interior
    hist, maxabs = ops.svf_exp_fwd(v, 12)
    g_v = ops.svf_exp_bwd(v, hist, maxabs, G0.expand(1, 3, n, n, n).contiguous())
    m = 8
    inner = g_v[..., m:-m, m:-m, m:-m]
    assert rel(inner, G0.expand_as(inner)) < 1e-6


@pytest.mark.parametrize('n', [128, 256])
def test_warps_identity_and_nearest_idempotent(ops, n):
    lin = torch.linspace(-1, 1, steps=n, device=DEV)
    gz, gy, gx = torch.meshgrid(lin, lin, lin, indexing='ij')
    T = torch.stack((gx, gy, gz), 0).unsqueeze(0).contiguous()
    im = torch.rand(1, 1, n, n, n, device=DEV)
    w = ops.warp3d(im, T)
    assert rel(w, im) < 2e-7 * n   # the fp32 identity grid is ~n * 2^-24 voxels off the nodes, in the reference too
    aten = torch.nn.functional.grid_sample(im, T.permute(0, 2, 3, 4, 1), padding_mode='border', align_corners=True)
    assert rel(w, aten) < 1e-6
    seg = (torch.rand(1, 1, n, n, n, device=DEV) * 60).short()
    assert torch.equal(ops.warp3d_nearest(seg, T), seg)
    # a whole-voxel shift in x: nearest warp == index shift with the border replicated
    Ts = T.clone()
    Ts[:, 0] += 2.0 * 3 / (n - 1)
    shifted = ops.warp3d_nearest(seg, Ts)
    ref = torch.cat((seg[..., 3:], seg[..., -1:].expand(-1, -1, -1, -1, 3)), -1)
    frac = float((shifted != ref).float().mean())
    assert frac < 1e-3   # fp32 unnormalisation can land a hair off the .0 of a few columns; never more than a column's worth
    counts, log_det = ops.log_det_jacobian(T)
    assert int(counts.sum()) == 0 and float(log_det.abs().max()) < 1e-4
    _, log_det2 = ops.log_det_jacobian((2 * T).contiguous())
    assert abs(float(log_det2.mean()) - math.log(8.0)) < 1e-4


@pytest.mark.parametrize('n', [128])
def test_smoothing_and_lcc_invariants(ops, n):
    from irsgmcmc_b200.utils.functions import langevin_sobolev, Sobolev_kernel_1D
    taps = list(Sobolev_kernel_1D(3, 0.5)[0].astype('float32'))
    v = torch.full((1, 3, n, n, n), 2.5, device=DEV)
    assert rel(langevin_sobolev(v, None, 0.0, taps), v) < 1e-6          # taps sum to one, replicate border
    x = torch.arange(n, dtype=torch.float32, device=DEV).view(1, 1, 1, 1, n).expand(1, 3, n, n, n).contiguous()
    sm = langevin_sobolev(x, None, 0.0, taps)
    assert rel(sm[..., 3:-3], x[..., 3:-3]) < 1e-6                      # symmetric kernel reproduces linear fields inside
    im = torch.rand(1, 1, n, n, n, device=DEV)
    zn, a, rs = ops.lcc_normalise(im, 2)
    zn2, _, _ = ops.lcc_normalise(3.0 * im + 0.7, 2)
    assert rel(zn2, zn) < 1e-3                                          # LCC is invariant to affine intensity changes
    y = ops.reg_energy(v)
    assert float(y.abs().max()) == 0.0
    assert abs(float(ops.reg_energy(x)[0]) / (3 * n ** 3) - 1.0) < 1e-5  # unit slope along x in all 3 components, last diff twice


def test_full_transition_fullsize_finite_and_reproducible(built):
    """two samplers with the same seed produce bit-identical chains (no floating-point atomics on the path)"""
    from irsgmcmc_b200.sampler import SGLDSampler, SGLDConfig
    from irsgmcmc_b200.data_loader.synthetic import make_pair
    n, C = 128, 2
    fixed, moving, vp = make_pair(n)
    outs = []
    for _ in range(2):
        s = SGLDSampler(fixed, moving, C, SGLDConfig(), device=DEV)
        s.init_chains('VI', vp, generator=torch.Generator(device=DEV).manual_seed(5))
        s.init_gmm()
        s.step(6)
        torch.cuda.synchronize()
        assert torch.isfinite(s.v).all() and torch.isfinite(s.stats).all()
        outs.append((s.v.clone(), s.hyper.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    # chains are keyed by their global index: a sampler that owns only chain 1 reproduces it
    s1 = SGLDSampler(fixed, moving, 1, SGLDConfig(), device=DEV, chain_offset=1)
    e_full = s.__class__  # noqa: F841 (documentation: same class)
    from irsgmcmc_b200.utils.functions import langevin_sobolev
    a = langevin_sobolev(torch.zeros(2, 3, 8, 8, 8, device=DEV), None, 1.0, [], seed=123, iteration=4, chain0=0)
    b = langevin_sobolev(torch.zeros(1, 3, 8, 8, 8, device=DEV), None, 1.0, [], seed=123, iteration=4, chain0=1)
    assert torch.equal(a[1], b[0])


@pytest.mark.parametrize('n', [4, 7, 9, 33, 40])
def test_odd_cubic_volumes(ops, n):
    """every op on cubes below / not a multiple of any tile size"""
    C = 2
    torch.manual_seed(1)
    v = smooth_field((C, 3, n, n, n), min(1.5, 0.2 * n), 3)
    hist, maxabs = ops.svf_exp_fwd(v.to(DEV), 12)
    v64 = v.double().requires_grad_(True)
    T64, d64 = O.svf_exp_aten(v64, 12, exact_grid=True)
    assert rel(hist[-1], d64) < 1e-5
    lin = [torch.linspace(-1, 1, steps=n).to(DEV)] * 3
    T32, d32 = O.svf_exp_aten(v.clone(), 12)
    assert rel(ops.svf_outputs(hist[-1], lin), T32) < 1e-5
    G = torch.randn(C, 3, n, n, n, generator=torch.Generator().manual_seed(2))
    g64, = torch.autograd.grad((d64 * G.double()).sum(), v64)
    v32 = v.clone().requires_grad_(True)
    _, d32b = O.svf_exp_aten(v32, 12)
    g32, = torch.autograd.grad((d32b * G).sum(), v32)
    assert grad_ok(ops.svf_exp_bwd(v.to(DEV), hist, maxabs, G.to(DEV)), g32, g64, f'svf adjoint n={n}')


@pytest.mark.parametrize('dims', [(12, 16, 20), (9, 33, 8), (5, 6, 7), (40, 24, 36)])
def test_non_cubic_volumes(ops, dims):
    """the operators that are defined for any D, H, W (everything except the SVF integrator, which the reference only
    supports on cubes -- see IRS_CHECK_CUBE)"""
    D, H, W = dims
    C = 2
    torch.manual_seed(1)
    v = smooth_field((C, 3, D, H, W), 1.5, 3)
    with pytest.raises(RuntimeError):
        ops.svf_exp_fwd(v.to(DEV), 12)
    T32 = O.identity_grid((W, H, D)).permute(0, 4, 1, 2, 3) + O.to_normalised(v)
    T32 = T32.contiguous()
    im = torch.rand(1, 1, D, H, W)
    assert rel(ops.warp3d(im.to(DEV), T32.to(DEV)), O.warp_aten(im.expand(C, -1, -1, -1, -1), T32)) < 1e-5
    g_out = torch.randn(C, 1, D, H, W)
    T64 = T32.double().requires_grad_(True)
    (O.warp_aten(im.double().expand(C, -1, -1, -1, -1), T64) * g_out.double()).sum().backward()
    assert rel(ops.warp3d_bwd_grid(im.to(DEV), T32.to(DEV), g_out.to(DEV)), T64.grad) < 1e-4
    seg = (torch.rand(1, 1, D, H, W) * 30).short()
    assert torch.equal(ops.warp3d_nearest(seg.to(DEV), T32.to(DEV)).cpu(), O.warp_nearest(seg, T32))
    counts, log_det = ops.log_det_jacobian(T32.to(DEV))
    ref = O.det_jacobian(O.forward_differences(T32.double(), transformation=True)).log()
    ok = torch.isfinite(ref)
    assert counts.cpu().tolist() == torch.isnan(ref).sum(dim=(1, 2, 3)).tolist() and rel(log_det.cpu()[ok], ref[ok]) < 1e-4
    if min(dims) >= 5:
        img = torch.rand(C, 1, D, H, W)
        zn, a, rs = ops.lcc_normalise(img.to(DEV), 2)
        i64 = img.double().requires_grad_(True)
        z64 = O.lcc_normalise(i64, 2)
        assert rel(zn, z64) < 1e-4
        Gz = torch.randn(C, 1, D, H, W)
        (z64 * Gz.double()).sum().backward()
        assert rel(ops.lcc_normalise_bwd(Gz.to(DEV), a, rs, 2), i64.grad) < 1e-3
    assert rel(ops.reg_energy(v.to(DEV)), O.reg_energy(v.double())) < 1e-6
    assert rel(ops.diff_fwd(v.to(DEV), True), O.forward_differences(v.double(), True)) < 1e-6
    from irsgmcmc_b200.utils.functions import langevin_sobolev, Sobolev_kernel_1D
    taps = Sobolev_kernel_1D(min(3, (min(dims) - 1) // 2), 0.5)[0].astype('float32')
    assert rel(langevin_sobolev(v.to(DEV), None, 0.0, list(taps)), O.sobolev_smooth(v.double(), taps)) < 1e-6


def test_many_chains_small_volume(built):
    """BASELINE.json configs[4] in miniature: many chains at a small volume (SSD, K = 1)"""
    from irsgmcmc_b200.sampler import SGLDSampler, SGLDConfig
    from irsgmcmc_b200.data_loader.synthetic import make_pair
    n, C = 16, 24
    fixed, moving, vp = make_pair(n)
    s = SGLDSampler(fixed, moving, C, SGLDConfig(data_loss='ssd', reg_loss='RegLoss_L2', w_reg=1.4, reg_learnable=False), device=DEV)
    s.init_chains('VI', vp)
    s.init_gmm()
    s.step(5)
    terms = s.loss_terms()
    assert all(torch.isfinite(t).all() for t in terms.values()) and terms['data'].shape == (C,)
    assert int(s.hyper[0].item()) == 25 + 5 * C   # the shared mixture was stepped once per chain per transition
