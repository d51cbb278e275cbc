"""GPU parity of every op-level entry point of the C ABI against the oracle (oracle/sgld_oracle.py) on seeded inputs."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import sgld_oracle as O
from tests.util import grad_ok, rel, smooth_field, three_numbers

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.fixture(scope='module')
def ops(built):
    from irsgmcmc_b200 import ops
    return ops


def make_T(n, C, amp, seed):
    v = smooth_field((C, 3, n, n, n), amp, seed)
    T, disp = O.svf_exp_aten(v, 6)
    return T.contiguous()


@pytest.mark.parametrize('n,C', [(16, 2), (33, 1)])
def test_warp_trilinear_fwd_bwd(ops, n, C):
    torch.manual_seed(1)
    im = torch.rand(1, 1, n, n, n)
    T = make_T(n, C, 2.5, 3)
    T[0, :, 0, 0, :] = -1.3     # out-of-range coordinates exercise the border clamp
    T[-1, :, 1, :, 0] = 1.0      # exactly on the border: zero grid gradient
    out = ops.warp3d(im.to(DEV), T.to(DEV))
    ref64 = O.warp_aten(im.double().expand(C, -1, -1, -1, -1), T.double())
    ref32 = O.warp_aten(im.expand(C, -1, -1, -1, -1), T)
    e = three_numbers(out, ref32, ref64)
    print('warp fwd', e)
    assert e[0] < 1e-5 and e[2] < 1e-5

    g_out = torch.randn(C, 1, n, n, n)
    T64 = T.double().requires_grad_(True)
    (O.warp_aten(im.double().expand(C, -1, -1, -1, -1), T64) * g_out.double()).sum().backward()
    T32 = T.clone().requires_grad_(True)
    (O.warp_aten(im.expand(C, -1, -1, -1, -1), T32) * g_out).sum().backward()
    g = ops.warp3d_bwd_grid(im.to(DEV), T.to(DEV), g_out.to(DEV))
    assert grad_ok(g, T32.grad, T64.grad, 'warp bwd')


def test_warp_jitter(ops):
    n, C, alpha = 16, 2, 0.1
    torch.manual_seed(2)
    im = torch.rand(1, 1, n, n, n)
    T = make_T(n, C, 1.5, 5)
    ju = torch.rand(C, 3, n, n, n)
    out = ops.warp3d(im.to(DEV), T.to(DEV), ju.to(DEV), alpha)
    ref = O.warp_aten(im.expand(C, -1, -1, -1, -1), T + O.uniform_jitter_normalised(ju, alpha, T.shape))
    assert rel(out, ref) < 1e-5


@pytest.mark.parametrize('dtype', [torch.int16, torch.bool])
def test_warp_nearest_bit_exact(ops, dtype):
    """bit-exact against ATen's nearest sampler on the same device (what the reference runs) and against the oracle"""
    n, C = 24, 2
    torch.manual_seed(3)
    seg = (torch.rand(1, 1, n, n, n) * 60).to(torch.int16) if dtype == torch.int16 else torch.rand(1, 1, n, n, n) > 0.5
    T = make_T(n, C, 3.0, 7)
    # adversarial coordinates: exact .5 voxel positions, borders, far outside, +-inf
    k = torch.arange(n, dtype=torch.float32)
    half = (2.0 * (k + 0.5) / (n - 1) - 1.0)
    T[0, 0, 0, 0, :] = half
    T[0, 1, 0, :, 0] = half
    T[0, 2, :, 0, 0] = half
    T[1, :, 1, 1, :4] = torch.tensor([-1.0, 1.0, -7.0, 9.0])
    T[1, :, 2, 2, :2] = torch.tensor([float('inf'), float('-inf')])
    out = ops.warp3d_nearest(seg.to(DEV), T.to(DEV))
    aten = F.grid_sample(seg.to(DEV).float().expand(C, -1, -1, -1, -1), T.to(DEV).permute(0, 2, 3, 4, 1), mode='nearest',
                         padding_mode='border', align_corners=True).to(dtype)
    assert torch.equal(out, aten)
    assert torch.equal(out.cpu(), O.warp_nearest(seg, T))
    assert torch.equal(out.cpu(), O.warp_nearest_aten(seg.expand(C, -1, -1, -1, -1), T))


# 16 / 24 / 20 / 40: TMA kernels (40 = several tiles in x and y, partial last tile, several z segments; 4.0 and 9.0 cross
# |u| = 1, 2 inside the pass and reach the ring / scatter paths); 18 = row pitch the TMA unit cannot address (ring kernels)
@pytest.mark.parametrize('n,C,amp', [(16, 2, 0.8), (24, 1, 4.0), (20, 2, 9.0), (40, 1, 1.5), (18, 1, 1.0)])
def test_svf_fwd_bwd(ops, n, C, amp):
    v = smooth_field((C, 3, n, n, n), amp, 11)
    hist, maxabs = ops.svf_exp_fwd(v.to(DEV), 12)
    v64 = v.double().requires_grad_(True)
    T64, d64 = O.svf_exp_aten(v64, 12, exact_grid=True)
    v32 = v.clone().requires_grad_(True)
    T32, d32 = O.svf_exp_aten(v32, 12)
    e = three_numbers(hist[-1], d32, d64)
    print('svf fwd', amp, e, maxabs.cpu().tolist()[-3:])
    assert e[0] < 1e-5

    lin = [torch.linspace(-1, 1, steps=n).to(DEV)] * 3
    T = ops.svf_outputs(hist[-1], lin)
    assert rel(T, T32) < 1e-6

    G = torch.randn(C, 3, n, n, n, generator=torch.Generator().manual_seed(5))
    g64, = torch.autograd.grad((d64 * G.double()).sum(), v64)
    g32, = torch.autograd.grad((d32 * G).sum(), v32)
    for radius_max in (8, 0):   # gather everywhere / atomic scatter everywhere
        g = ops.svf_exp_bwd(v.to(DEV), hist, maxabs, G.to(DEV), radius_max)
        assert grad_ok(g, g32, g64, f'svf bwd amp={amp} radius_max={radius_max}')


@pytest.mark.parametrize('n,bump', [(48, 3.0), (40, 7.0)])
def test_svf_adjoint_window_per_tile(ops, n, bump):
    """a deformation that exceeds one voxel only inside a local bump: tiles far from it keep the TMA kernel (window 1) while
    the tiles around it take the wider ring (3.0: |u| < 2, gather regime) or everything falls back (7.0: scatter regime
    in the last step, per-tile windows in the one before) -- forward and adjoint must stay exact in every mix"""
    C = 1
    v = smooth_field((C, 3, n, n, n), 0.6, 21)
    z, y, x = torch.meshgrid(torch.arange(n), torch.arange(n), torch.arange(n), indexing='ij')
    c = torch.tensor([0.72 * n, 0.3 * n, 0.25 * n])
    w = torch.exp(-((z - c[0]) ** 2 + (y - c[1]) ** 2 + (x - c[2]) ** 2) / (2 * (0.09 * n) ** 2))
    v = v + bump * w * torch.tensor([1.0, -0.7, 0.5]).view(1, 3, 1, 1, 1)
    hist, maxabs = ops.svf_exp_fwd(v.to(DEV), 12)
    assert float(maxabs[-1]) > 1.0 and float(maxabs[0]) < 0.01
    v64 = v.double().requires_grad_(True)
    T64, d64 = O.svf_exp_aten(v64, 12, exact_grid=True)
    v32 = v.clone().requires_grad_(True)
    T32, d32 = O.svf_exp_aten(v32, 12)
    assert three_numbers(hist[-1], d32, d64)[0] < 1e-5
    G = torch.randn(C, 3, n, n, n, generator=torch.Generator().manual_seed(6))
    g64, = torch.autograd.grad((d64 * G.double()).sum(), v64)
    g32, = torch.autograd.grad((d32 * G).sum(), v32)
    for radius_max in (2, 3):
        g = ops.svf_exp_bwd(v.to(DEV), hist, maxabs, G.to(DEV), radius_max)
        assert grad_ok(g, g32, g64, f'svf bwd local bump {bump} radius_max={radius_max}')


@pytest.mark.parametrize('s', [1, 2, 3])
def test_langevin_sobolev(ops, s):
    from irsgmcmc_b200.utils.functions import langevin_sobolev, Sobolev_kernel_1D
    n, C = 18, 2
    torch.manual_seed(4)
    v, sigma, eps = torch.randn(C, 3, n, n, n), torch.rand(1, 3, n, n, n) + 0.5, torch.randn(C, 3, n, n, n)
    taps = Sobolev_kernel_1D(s, 0.5)[0].astype('float32')
    coef = math.sqrt(2 * 0.4)
    out = langevin_sobolev(v.to(DEV), sigma.to(DEV), coef, list(taps), eps=eps.to(DEV))
    ref = O.sobolev_smooth(O.langevin_proposal(v.double(), sigma.double(), 0.4, eps.double()), taps)
    assert rel(out, ref) < 1e-6
    out0 = langevin_sobolev(v.to(DEV), None, 0.0, [])
    assert torch.equal(out0.cpu(), v)


def test_philox_noise_statistics(ops):
    from irsgmcmc_b200.utils.functions import langevin_sobolev
    n, C = 64, 2
    v = torch.zeros(C, 3, n, n, n, device=DEV)
    e1 = langevin_sobolev(v, None, 1.0, [], seed=7, iteration=3)
    e2 = langevin_sobolev(v, None, 1.0, [], seed=7, iteration=3)
    e3 = langevin_sobolev(v, None, 1.0, [], seed=7, iteration=4)
    assert torch.equal(e1, e2) and not torch.equal(e1, e3)
    x = e1.double().flatten()
    N = x.numel()
    assert abs(x.mean()) < 5 / math.sqrt(N) and abs(x.var() - 1) < 5 * math.sqrt(2 / N)
    assert abs((x ** 3).mean()) < 5 * math.sqrt(15 / N) and abs((x ** 4).mean() - 3) < 5 * math.sqrt(96 / N)
    flat = e1.view(C, 3, -1)
    for a, b in ((flat[0, 0], flat[0, 1]), (flat[0, 0], flat[1, 0]), (flat[0, 2, :-1], flat[0, 2, 1:]),
                 (e1.flatten(), e3.flatten())):
        corr = float((a.double() * b.double()).mean())
        assert abs(corr) < 5 / math.sqrt(a.numel())


@pytest.mark.parametrize('n', [16, 21])
def test_diff_op_and_energy(ops, n):
    C = 2
    torch.manual_seed(6)
    v = torch.randn(C, 3, n, n, n)
    for tr in (False, True):
        nabla = ops.diff_fwd(v.to(DEV), tr)
        assert rel(nabla, O.forward_differences(v.double(), tr)) < 1e-6
        G = torch.randn(C, 3, n, n, n, 3)
        v64 = v.double().requires_grad_(True)
        (O.forward_differences(v64, tr) * G.double()).sum().backward()
        assert rel(ops.diff_bwd(G.to(DEV), tr), v64.grad) < 1e-6
    y = ops.reg_energy(v.to(DEV))
    v64 = v.double().requires_grad_(True)
    y64 = O.reg_energy(v64)
    assert rel(y, y64) < 1e-6
    coef = torch.tensor([0.7, -1.3], dtype=torch.float64)
    (y64 * coef).sum().backward()
    assert rel(ops.reg_energy_grad(v.to(DEV), coef.to(DEV)), v64.grad) < 1e-6


@pytest.mark.parametrize('n,s', [(16, 1), (16, 2), (35, 2), (12, 3)])
def test_lcc_normalise_fwd_bwd(ops, n, s):
    C = 2
    torch.manual_seed(8)
    im = smooth_field((C, 1, n, n, n), 1.0, 2, passes=1) + 0.05 * torch.randn(C, 1, n, n, n)
    zn, a, rs = ops.lcc_normalise(im.to(DEV), s)
    im64 = im.double().requires_grad_(True)
    zn64 = O.lcc_normalise(im64, s)
    im32 = im.clone().requires_grad_(True)
    zn32 = O.lcc_normalise(im32, s)
    e = three_numbers(zn, zn32, zn64)
    print('lcc fwd', n, s, e)
    assert e[0] <= max(1e-5, 2 * e[1])
    G = torch.randn(C, 1, n, n, n)
    (zn64 * G.double()).sum().backward()
    (zn32 * G).sum().backward()
    g = ops.lcc_normalise_bwd(G.to(DEV), a, rs, s)
    assert grad_ok(g, im32.grad, im64.grad, f'lcc bwd n={n} s={s}')


@pytest.mark.parametrize('K', [1, 4])
def test_gmm_log_pdf_and_vd(ops, K):
    n = 20
    torch.manual_seed(9)
    z = smooth_field((1, 1, n, n, n), 2.0, 4, passes=1) + 0.3 * torch.randn(1, 1, n, n, n)
    mask = torch.rand(1, 1, n, n, n) > 0.3
    log_std = torch.linspace(math.log(0.01), math.log(5.0), K) if K > 1 else torch.tensor([0.2])
    logits = 0.3 * torch.randn(K)
    w = torch.rand(z.numel())
    logp, dz, gp = ops.gmm_log_pdf(z.to(DEV).flatten(), log_std, logits, True, w.to(DEV), True)
    z64 = z.double().flatten().requires_grad_(True)
    ls64, lg64 = log_std.double().requires_grad_(True), logits.double().requires_grad_(True)
    lp64 = O.gmm_log_pdf(z64, ls64, lg64)[0]
    assert rel(logp, lp64) < 1e-5
    (lp64 * w.double()).sum().backward()
    gz_ref = z64.grad / w.double()
    assert rel(dz, gz_ref) < 1e-5
    g_ls = gp[:K].cpu()
    g_lg = gp[8:8 + K].cpu() - torch.softmax(lg64.detach() + 1e-2, 0) * w.double().sum()
    assert rel(g_ls, ls64.grad) < 1e-5 and (K == 1 or rel(g_lg, lg64.grad) < 1e-4)
    if K > 1:
        alpha = ops.vd_factor(z.to(DEV), mask.to(DEV), log_std, logits)
        ref = O.vd_factor(O.vd_residual(z.double(), mask, log_std.double(), logits.double()), mask)
        print('vd', float(alpha), float(ref))
        assert abs(float(alpha) - float(ref)) < 1e-5 * abs(float(ref))
    ms = ops.masked_mean_std(z.to(DEV), mask.to(DEV)).cpu()
    assert abs(ms[0] - z[mask].double().mean()) < 1e-6 and abs(ms[1] - z[mask].double().std()) < 1e-6
    assert ms[2] == mask.sum()


def test_welford(ops):
    torch.manual_seed(10)
    samples = torch.randn(12, 3, 8, 8, 8)
    mean, m2 = torch.zeros(3, 8, 8, 8, device=DEV), torch.zeros(3, 8, 8, 8, device=DEV)
    count = 0
    for chunk in samples.split(4):
        count = ops.welford_update(chunk.to(DEV).contiguous(), count, mean, m2)
    ref_mean, ref_std = O.posterior_statistics(samples)
    assert rel(mean, ref_mean) < 1e-6 and rel(ops.welford_std(m2, count), ref_std) < 1e-6


def test_no_cpu_fallback(ops):
    with pytest.raises(RuntimeError):
        ops.warp3d(torch.rand(1, 1, 8, 8, 8), torch.rand(1, 3, 8, 8, 8))
