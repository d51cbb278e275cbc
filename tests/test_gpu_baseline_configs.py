"""
Oracle parity AT the BASELINE.json sizes: one full transition with injected noise against the
fp32 and fp64 oracle at 64^3 SSD + RegLoss_L2 (configs[0]) and at 128^3 LCC + RegLoss_LogNormal (configs[1]), same
three-number protocol as tests/test_gpu_sgld.py; bit-exact nearest-neighbour warp + Dice counts at 256^3 against ATen on
the same device (configs[3]).  The 128^3 kernels take multi-wave grids, 10-13-plane z segments and (with the large
velocity case) the per-tile cell-map path that the 16^3..40^3 cases barely touch.
"""
import pytest
import torch
import torch.nn.functional as F

from oracle import sgld_oracle as O
from tests.test_gpu_sgld import check, run_pair
from tests.util import grad_ok, rel, smooth_field

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def test_config0_64_ssd_l2_transition_vs_oracle(built):
    """BASELINE configs[0]: 64^3, SSD data term (K = 1, map = F - M) + diffusion regulariser (RegLoss_L2, w = 1.4)"""
    check(run_pair(64, 1, 'ssd', 'l2', False, iters=1))


def test_config0_64_ssd_l2_learnable_two_chains(built):
    check(run_pair(64, 2, 'ssd', 'l2', True, iters=1))


def test_config1_128_lcc_lognormal_transition_vs_oracle(built):
    """BASELINE configs[1]: 128^3, LCC + 4-component mixture, virtual decimation, jitter, RegLoss_LogNormal learnable"""
    check(run_pair(128, 1, 'lcc', 'lognormal', True, iters=1))


@pytest.mark.parametrize('n,amp', [(64, 6.0), (128, 10.0)])
def test_svf_exp_and_adjoint_fullsize_vs_fp64_oracle(built, n, amp):
    """scaling and squaring + its adjoint on a smooth velocity of several voxels at the BASELINE sizes: the last squaring
    steps exceed one voxel locally, so the forward kernel's in-kernel fallbacks and the adjoint's per-tile window choice
    (cell map) run on multi-wave grids; the gradient is judged with the three-number protocol"""
    from irsgmcmc_b200 import ops
    v = smooth_field((1, 3, n, n, n), amp, seed=5, passes=6)
    g = smooth_field((1, 3, n, n, n), 1.0, seed=6, passes=2)
    hist, maxabs = ops.svf_exp_fwd(v.to(DEV), 12)
    g_v = ops.svf_exp_bwd(v.to(DEV), hist, maxabs, g.to(DEV))
    torch.cuda.synchronize()
    out = {}
    for dtype in (torch.float32, torch.float64):
        vv = v.to(dtype).clone().requires_grad_(True)
        _, disp = O.svf_exp_aten(vv, 12, exact_grid=(dtype == torch.float64))
        (disp * g.to(dtype)).sum().backward()
        out[dtype] = (disp.detach(), vv.grad)
    print('max |u_12| =', float(out[torch.float64][0].abs().max()), 'maxabs per step', maxabs[:12].tolist())
    assert float(maxabs[11]) > 1.0          # the multi-voxel regime is really exercised
    assert rel(hist[-1], out[torch.float64][0]) < 1e-5
    assert grad_ok(g_v, out[torch.float32][1], out[torch.float64][1], f'svf adjoint {n}^3')


def test_config3_256_nearest_warp_and_dice_bit_exact(built):
    """BASELINE configs[3]: 256^3, nearest-neighbour warp of the int16 segmentation and of the bool mask with a
    sampler-like transformation, bit-exact against ATen on the same device; Dice counts equal to integer counting"""
    from irsgmcmc_b200 import ops
    from irsgmcmc_b200.data_loader.synthetic import make_pair
    n = 256
    fixed, moving, _ = make_pair(n)
    v = smooth_field((1, 3, n, n, n), 5.0, seed=11, passes=4).to(DEV)
    hist, _ = ops.svf_exp_fwd(v, 12)
    lin = [torch.linspace(-1, 1, steps=n).to(DEV)] * 3   # the reference's fp32 identity grid (utils/util.py:270-272)
    T = ops.svf_outputs(hist[-1], lin)
    grid = T.permute(0, 2, 3, 4, 1)
    seg = moving['seg'].to(DEV)
    mask = moving['mask'].to(DEV)
    w_seg = ops.warp3d_nearest(seg, T)
    ref_seg = F.grid_sample(seg.float(), grid, mode='nearest', padding_mode='border', align_corners=True).short()
    assert torch.equal(w_seg, ref_seg)
    w_mask = ops.warp3d_nearest(mask, T)
    ref_mask = F.grid_sample(mask.float(), grid, mode='nearest', padding_mode='border', align_corners=True).bool()
    assert w_mask.dtype == torch.bool and torch.equal(w_mask, ref_mask)
    assert int((w_seg != seg).sum()) > 1000   # the warp really moved labels

    labels = [10, 11, 12, 13, 16, 17, 18, 26, 49, 50, 51, 52, 53, 54, 58]   # reference parse_config.py:54-58
    seg_f = fixed['seg'].to(DEV)
    counts = ops.dice_counts(seg_f, w_seg, labels)
    for i, lab in enumerate(labels):
        a, b = seg_f == lab, w_seg == lab
        expect = (int(a.sum()), int(b.sum()), int((a & b).sum()))
        got = tuple(int(x) for x in counts[0, i])
        assert got == expect, (lab, got, expect)
