"""
The configuration matrix of BASELINE.json / SURVEY.md section 8(d) on ONE GPU (development aid; bench.py stays the contract):
  1  64^3  SSD + RegLoss_L2, 1 chain
  2  128^3 LCC, 1 chain                                     (the headline; bench.py measures it properly)
  3  128^3 LCC, 64 / 32 / 16 / 8 chains                     (the per-GPU shards of 64 chains over 1 / 2 / 4 / 8 GPUs)
  4  256^3 LCC, 1 chain, + nearest-neighbour segmentation warp and Dice counts per kept sample
  5  64^3  SSD, 1024 chains;  256^3 LCC, 16 chains          (bounded by --max-gib)
Prints one line per configuration: ms per transition and voxel-steps/s (graph replay, CUDA events, 5 warm-up transitions).

    python tools/bench_matrix.py [--only 1,3] [--steps 20] [--max-gib 120]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from irsgmcmc_b200 import ops  # noqa: E402
from irsgmcmc_b200.data_loader.synthetic import STRUCTURE_LABELS, make_pair  # noqa: E402
from irsgmcmc_b200.sampler import SGLDConfig, SGLDSampler  # noqa: E402

DEV = 'cuda:0'


def gib_needed(n, chains, steps=12):
    """state + workspace of SGLDSampler: (5 + steps) fields of 3 V floats and 6 volumes per chain"""
    V = n ** 3
    return chains * ((5 + steps) * 3 * V + 6 * V) * 4 / 2 ** 30


def run(tag, n, chains, data, steps, with_seg=False):
    pair = make_pair(n)
    fixed, moving, vp = pair
    cfg = SGLDConfig() if data == 'lcc' else SGLDConfig(data_loss='ssd', reg_loss='RegLoss_L2', w_reg=1.4, reg_learnable=False)
    s = SGLDSampler(fixed, moving, chains, cfg, device=DEV)
    s.init_chains('VI', vp, generator=torch.Generator(device=DEV).manual_seed(123))
    s.init_gmm()
    s.step(5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    seg_f = fixed['seg'].to(DEV).expand(chains, -1, -1, -1, -1).contiguous() if with_seg else None
    e0.record()
    for _ in range(steps if with_seg else 1):
        s.step(1 if with_seg else steps)
        if with_seg:   # what a kept sample costs on top of the transition (trainer/trainer.py:414-445 of the reference)
            seg_w = s.warp_segmentation()
            ops.dice_counts(seg_f, seg_w, list(STRUCTURE_LABELS))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    extra = ' (+ segmentation warp + Dice counts every transition)' if with_seg else ''
    print(f'config {tag}: {n}^3 {data.upper()} x {chains} chain(s){extra}: {ms:.3f} ms per transition, '
          f'{chains * n ** 3 / ms / 1e6:.3f} G voxel-steps/s, {torch.cuda.max_memory_allocated() / 2 ** 30:.1f} GiB peak',
          flush=True)
    del s
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--only', default='1,2,3,4,5')
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--max-gib', type=float, default=120.0, help='skip configurations whose buffers need more than this')
    a = ap.parse_args()
    only = {int(x) for x in a.only.split(',')}
    matrix = [(1, '1', 64, 1, 'ssd', False), (2, '2', 128, 1, 'lcc', False)]
    matrix += [(3, f'3/{c}', 128, c, 'lcc', False) for c in (64, 32, 16, 8)]
    matrix += [(4, '4', 256, 1, 'lcc', True), (5, '5a', 64, 1024, 'ssd', False), (5, '5b', 256, 16, 'lcc', False)]
    for group, tag, n, chains, data, with_seg in matrix:
        if group not in only:
            continue
        need = gib_needed(n, chains)
        if need > a.max_gib:
            print(f'config {tag}: skipped, needs {need:.0f} GiB (> --max-gib {a.max_gib:.0f})', flush=True)
            continue
        run(tag, n, chains, data, a.steps, with_seg)


if __name__ == '__main__':
    if not torch.cuda.is_available():
        raise SystemExit('tools/bench_matrix.py needs a CUDA device')
    main()
