"""Times individual C-ABI ops with CUDA events at a given volume size (development aid; not part of the bench contract)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from irsgmcmc_b200 import ops  # noqa: E402
from irsgmcmc_b200.utils.functions import langevin_sobolev, Sobolev_kernel_1D  # noqa: E402


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3  # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--size', type=int, default=128)
    ap.add_argument('--chains', type=int, default=1)
    ap.add_argument('--amp', type=float, default=1.5, help='max |v| in voxels')
    ap.add_argument('--ffd', type=int, default=0, help='control point spacing: time only the B-spline FFD ops')
    args = ap.parse_args()
    n, C = args.size, args.chains
    dev = 'cuda:0'
    torch.manual_seed(0)
    if args.ffd:
        from irsgmcmc_b200.utils import B_spline_1D_kernel, get_control_grid_size
        cps, dims = (args.ffd,) * 3, (n,) * 3
        grid = get_control_grid_size(dims, cps)
        ks = [tuple(float(x) for x in B_spline_1D_kernel(args.ffd))] * 3
        cp, G = torch.randn(C, 3, *grid, device=dev), torch.randn(C, 3, *dims, device=dev)
        for name, fn in ((f'ffd_fwd cps={args.ffd}', lambda: ops.ffd_fwd(cp, ks, cps, dims)),
                         (f'ffd_bwd cps={args.ffd}', lambda: ops.ffd_bwd(G, ks, cps, grid))):
            us = timeit(fn)
            print(f'{name:55s} {us:10.1f} us   {us * 1e3 / (C * n ** 3):8.3f} ns/voxel')
        return
    taps = list(Sobolev_kernel_1D(3, 0.5)[0].astype('float32'))
    v = langevin_sobolev(torch.randn(C, 3, n, n, n, device=dev), None, 0.0, taps)
    v = v / v.abs().max() * args.amp
    res = {}
    res['langevin+sobolev'] = timeit(lambda: langevin_sobolev(v, None, 0.9, taps, seed=1, iteration=2))
    hist, maxabs = ops.svf_exp_fwd(v, 12)
    print('maxabs per step', [round(x, 3) for x in maxabs.tolist()])
    res['svf_fwd (12 steps)'] = timeit(lambda: ops.svf_exp_fwd(v, 12))
    G = torch.randn(C, 3, n, n, n, device=dev)
    mask = torch.zeros(C, 3, n, n, n, device=dev)
    q = n // 4
    mask[..., q:-q, q:-q, q:-q] = 1
    for name, g in (('dense g', G), ('g on the central 1/8', G * mask)):
        for rm in (2, 0):
            res[f'svf_bwd (12 steps) radius_max={rm} {name}'] = timeit(lambda: ops.svf_exp_bwd(v, hist, maxabs, g, rm), n=5)
    im = torch.rand(C, 1, n, n, n, device=dev)
    res['lcc_normalise s=2'] = timeit(lambda: ops.lcc_normalise(im, 2))
    zn, a, rs = ops.lcc_normalise(im, 2)
    res['lcc_normalise_bwd s=2'] = timeit(lambda: ops.lcc_normalise_bwd(zn, a, rs, 2))
    res['reg_energy'] = timeit(lambda: ops.reg_energy(v))
    lin = [torch.linspace(-1, 1, n, device=dev)] * 3
    T = ops.svf_outputs(hist[-1], lin)
    res['warp3d'] = timeit(lambda: ops.warp3d(im[:1], T))
    res['warp3d_bwd_grid'] = timeit(lambda: ops.warp3d_bwd_grid(im[:1], T, im))
    seg = (torch.rand(1, 1, n, n, n, device=dev) * 50).short()
    res['warp3d_nearest'] = timeit(lambda: ops.warp3d_nearest(seg, T))
    V = C * n ** 3
    for k, us in res.items():
        print(f'{k:55s} {us:10.1f} us   {us * 1e3 / V:8.3f} ns/voxel')


if __name__ == '__main__':
    main()
