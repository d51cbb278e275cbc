#!/usr/bin/env python
"""
bench.py -- SGLD voxel-steps/s of the B200-native registration step (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # own arm (under torchrun for N > 1)
    python bench.py --impl reference --steps K --warmup W     # the reference's CPU path (oracle port) on the host cores

One "step" = one SGLD transition (reference Trainer._SGLD_transition) of every chain resident on the GPU.
Workload (N = 1): BASELINE.json configs[1] -- 128^3 synthetic brain-MRI-shaped pair, LCC + 4-component GMM data term,
virtual decimation, uniform jitter, Sobolev s=3, RegLoss_LogNormal, 12 SVF steps, ONE chain.  N > 1: every rank runs
the same number of chains on its own GPU (weak scaling), no per-iteration collective; the Welford moments are merged
with NCCL once after the timed region (reported, not part of the metric).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_VOXEL_STEP_LCC = 895.0  # 175 + 60 * n_svf
BYTES_PER_VOXEL_STEP_SSD = 875.0
SVF_BWD_BYTES_PER_VOXEL = 36.0    # read g_{k+1} 12 + u_k 12, write g_k 12
# dram__bytes_read.sum + dram__bytes_write.sum of ONE svf_step_bwd_tma2_kernel launch from the committed ncu --set full capture
# (profiles/r2_ncu_final_summary.txt; 128^3, one chain).  Below the algorithmic 75.5 MB: the 24 MB result stays in the L2.
NCU_TRAFFIC_BYTES = {(128, 1, 'lcc'): 50.900e6 + 4.270e6}


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(path):
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region"""
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,' \
        'clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v == 'Active'})
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx[0] if mx else None, 'reasons': reasons,
                'samples': len(sm)}


def oracle_transition_rate(n, steps, warmup, chains=1, data='lcc', cps=0):
    """times the oracle port of Trainer._SGLD_transition on the host cores; returns (voxel-steps/s, ms/step, threads)
    cps > 0: SVFFD_3D with that control point spacing (the secondary configuration of --cps)"""
    import torch
    from oracle import sgld_oracle as O
    from irsgmcmc_b200.data_loader.synthetic import make_pair
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(123)
    fixed, moving, vp = make_pair(n)
    cfg = O.Config(data=data, K=4 if data == 'lcc' else 1, reg='lognormal' if data == 'lcc' else 'l2',
                   w_reg=1.6 if data == 'lcc' else 1.4, cps=(cps,) * 3 if cps else None)
    sdims = O.control_grid_size((n, n, n), cfg.cps) if cps else (n, n, n)
    sigma = torch.full((chains, 3, *sdims), 0.5)       # exp(log_var / 2) of make_pair's variational parameters
    v0 = sigma * torch.randn(chains, 3, *sdims) + 0.1 * torch.randn(1)
    st = O.State(cfg, v0, sigma, (n, n, n))
    O.gmm_init(st, fixed, moving, v0[:1], warm_up=5)
    for _ in range(warmup):
        O.sgld_transition(st, fixed, moving)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.sgld_transition(st, fixed, moving)
    dt = time.perf_counter() - t0
    return chains * n ** 3 * steps / dt, 1e3 * dt / max(steps, 1), torch.get_num_threads()


def aten_gpu_transition_rate(n, steps, warmup, data='lcc', cps=0, device='cuda'):
    """
    SURVEY section 8(d), "the reference running its own PyTorch path on the B200": the oracle port of Trainer._SGLD_transition
    (the reference's operators: F.grid_sample with its scatter-atomics backward, dense 125-tap convolutions, autograd) with
    every tensor on the GPU.  A reported baseline like cpu_baseline, never part of `value`.  The oracle creates its helper
    tensors with plain factory calls, so the default device is switched for the duration of the measurement.
    """
    import torch
    from oracle import sgld_oracle as O
    from irsgmcmc_b200.data_loader.synthetic import make_pair
    torch.manual_seed(123)
    fixed, moving, _ = make_pair(n)                       # host generators inside: before the default device changes
    fixed = {k: v.to(device) for k, v in fixed.items()}
    moving = {k: v.to(device) for k, v in moving.items()}
    torch.set_default_device(device)
    try:
        cfg = O.Config(data=data, K=4 if data == 'lcc' else 1, reg='lognormal' if data == 'lcc' else 'l2',
                       w_reg=1.6 if data == 'lcc' else 1.4, cps=(cps,) * 3 if cps else None)
        sdims = O.control_grid_size((n, n, n), cfg.cps) if cps else (n, n, n)
        sigma = torch.full((1, 3, *sdims), 0.5)
        v0 = sigma * torch.randn(1, 3, *sdims) + 0.1 * torch.randn(1)
        st = O.State(cfg, v0, sigma, (n, n, n))
        O.gmm_init(st, fixed, moving, v0[:1], warm_up=5)
        sync = torch.cuda.synchronize if str(device).startswith('cuda') else (lambda: None)
        for _ in range(warmup):
            O.sgld_transition(st, fixed, moving)
        sync()
        t0 = time.perf_counter()
        for _ in range(steps):
            O.sgld_transition(st, fixed, moving)
        sync()
        dt = time.perf_counter() - t0
    finally:
        torch.set_default_device('cpu')
    return n ** 3 * steps / dt, 1e3 * dt / max(steps, 1)


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path = the oracle port (the reference is Python on
    PyTorch; /root/reference does not travel to the GPU box), all host threads, rank 0 only."""
    if int(os.environ.get('RANK', '0')) != 0:
        return
    n = args.size
    # Bounded sample WITHOUT changing the workload: the volume stays the requested one (a smaller volume is another
    # workload: the driver's ratio must be like for like).  When K + W transitions of the CPU port would not finish within the
    # budget, fewer are timed -- each transition is the same deterministic amount of work, so voxel-steps/s does not depend on
    # the count -- and the line says so (`steps_timed`, `sample`).
    budget_s = float(os.environ.get('IRS_REF_BUDGET_S', '150'))
    probe_value, probe_ms, threads = oracle_transition_rate(n, 1, 0, 1, args.data, args.cps)   # one untimed-region warm-up
    steps_run = int(max(1, min(args.steps, (budget_s - probe_ms / 1e3) // (probe_ms / 1e3) - min(args.warmup, 1))))
    warm_run = min(args.warmup, 1) if steps_run < args.steps else args.warmup
    n_run = n
    value, ms, threads = oracle_transition_rate(n_run, steps_run, warm_run, 1, args.data, args.cps)
    sample = (f'{steps_run} timed + {warm_run} warm-up oracle transitions at {n_run}^3, 1 chain, fp32, {threads} threads'
              + ('' if steps_run == args.steps else f' (of the {args.steps} + {args.warmup} asked: time bound {budget_s:.0f} s; '
                                                    f'same volume, fewer repetitions)'))
    line = {'impl': 'reference', 'metric': 'SGLD voxel-steps/s', 'value': value, 'unit': 'voxel-steps/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'steps_timed': steps_run, 'ms_per_step': ms,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args, 1, n_run),
            'cpu_baseline': {'value': value, 'unit': 'voxel-steps/s', 'cores': threads, 'kind': 'port', 'sample': sample},
            'e2e': {'value': value, 'unit': 'voxel-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, chains, n):
    return {'workload': f'{n}^3 synthetic brain-MRI-shaped pair, {"LCC+GMM(K=4,s=2)" if args.data == "lcc" else "SSD(K=1)"} '
                        f'data term, virtual decimation, uniform jitter 0.1, Sobolev s=3, '
                        f'{"RegLoss_LogNormal w=1.6 learnable" if args.data == "lcc" else "RegLoss_L2 w=1.4"}, SVF 12 steps, '
                        f'tau 0.4, VI-shaped init; {chains} SGLD chain(s) per GPU (BASELINE.json configs[1])'
                        + (f'; SECONDARY configuration: SVFFD_3D transformation, control point spacing {args.cps}'
                           if getattr(args, 'cps', 0) else ''),
            'volume': [n, n, n], 'chains_per_gpu': chains, 'parallelism': 'independent chains sharded by rank',
            'l2_policy': 'working set per transition (SVF history 288 MiB/chain at 128^3 + 20 field-sized buffers) '
                         'exceeds the 126 MB L2; no explicit flush'}


def pin_rank_to_cores(local_rank, local_world):
    """one contiguous share of the host cores per rank: with 8 processes on one node an unpinned rank that loses its core for a
    millisecond shows up as 5 % of a 20 ms timed region (max over ranks)"""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = len(cores) // max(local_world, 1)
        if local_world > 1 and per >= 1:
            os.sched_setaffinity(0, cores[local_rank * per:(local_rank + 1) * per])
            return per
    except (AttributeError, OSError):
        pass
    return None


def graph_length(steps):
    """transitions per CUDA graph for a timed region of `steps`: the whole region in ONE replay up to 40 transitions,
    otherwise the largest divisor of `steps` in [10, 40] (10 with an eager remainder when there is none)"""
    if steps <= 40:
        return max(steps, 1)
    for g in range(40, 9, -1):
        if steps % g == 0:
            return g
    return 10


def run_config(tag, n, chains, data, world, rank, dev, peak, barrier, seg_dice=False, moments=False, flush_l2=False,
               target_ms=1200.0, vi=False, cpu_vi=False):
    """
    One entry of the `configs` sub-record: a BASELINE.json configuration other than the headline, measured in the same run with
    the same rules (CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks, >= 3 warm-up
    transitions).  `seg_dice`: every transition is followed by the nearest-neighbour warp of the int16 segmentation and the
    Dice counts of the 15 structures (the reference's own speed loop includes the segmentation warp, trainer.py:467-476).
    `moments`: also times the Welford update of all chains and the NCCL merge of the moments, warm.
    `flush_l2`: the working set fits the 126 MB L2 (64^3, one chain): a 256 MB buffer is overwritten between transitions and
    every transition is timed on its own.
    """
    import gc
    import torch
    import torch.distributed as dist
    from irsgmcmc_b200 import ops
    from irsgmcmc_b200.sampler import SGLDSampler, SGLDConfig
    from irsgmcmc_b200.data_loader.synthetic import make_pair, STRUCTURE_LABELS
    V = n ** 3
    fixed, moving, vp = make_pair(n, device=dev)
    lcc = data == 'lcc'
    cfg = SGLDConfig(data_loss=data, reg_loss='RegLoss_LogNormal' if lcc else 'RegLoss_L2', w_reg=1.6 if lcc else 1.4,
                     reg_learnable=lcc)
    sampler = SGLDSampler(fixed, moving, chains, cfg, device=dev, chain_offset=rank * chains)
    sampler.init_chains('VI', vp, generator=torch.Generator(device=dev).manual_seed(123))   # same draw on every rank (see main)
    sampler.init_gmm()
    seg_f = fixed['seg'].to(dev) if seg_dice else None
    counts = None

    def one():
        nonlocal counts
        sampler.step(1)
        if seg_dice:
            counts = ops.dice_counts(seg_f, sampler.warp_segmentation(), STRUCTURE_LABELS)

    def allmax(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(20):     # the same warm-up as the headline (>= 20 transitions): the speed of the SVF kernels depends on the state
        one()               # of the chain (sign pattern of the field), so every configuration is timed equally far from its start
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(); one(); one(); e1.record()
    barrier()
    est = allmax(e0.elapsed_time(e1)) / 2
    K = int(min(50, max(3, target_ms / max(est, 1e-3))))
    if flush_l2:
        flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        barrier()
        for a, b in evs:
            flush.zero_()
            a.record(); one(); b.record()
        barrier()
        ms = allmax(sum(a.elapsed_time(b) for a, b in evs))
        del flush
    else:
        barrier()
        e0.record()
        for _ in range(K):
            one()
        e1.record()
        barrier()
        ms = allmax(e0.elapsed_time(e1))
    value = world * chains * V * K / (ms * 1e-3)
    bytes_step = BYTES_PER_VOXEL_STEP_LCC if lcc else BYTES_PER_VOXEL_STEP_SSD
    out = {'config': tag, 'volume': [n, n, n], 'data_term': 'LCC+GMM(K=4)' if lcc else 'SSD(K=1)',
           'chains_per_gpu': chains, 'chains_total': chains * world, 'steps': K, 'warmup': 22,
           'ms_per_step': ms / K, 'voxel_steps_per_s': value, 'iterations_per_s': K / (ms * 1e-3),
           'step_roofline_frac': bytes_step * (value / world) / 1e9 / peak, 'bytes_per_voxel_step': bytes_step,
           'gpu_launches_per_step': sampler.launches_per_step() + (3 if seg_dice else 0),
           'l2': 'flushed between transitions (256 MB overwrite), each transition timed on its own' if flush_l2
                 else 'working set exceeds the 126 MB L2'}
    if seg_dice:
        c = counts[0].double().cpu()
        dice = (2 * c[:, 2] / (c[:, 0] + c[:, 1]).clamp(min=1)).tolist()
        out['per_step_extras'] = 'transformation + nearest-neighbour int16 segmentation warp + Dice counts (15 structures)'
        out['dice_mean_last_sample'] = sum(dice) / len(dice)
        if rank == 0:   # ASD of the last sample on the host, outside the timed region (the reference computes it there too, with
            try:        # SimpleITK: utils/util.py:171-176; here scipy's exact distance transform -- parity unpinned)
                from irsgmcmc_b200.utils.util import calc_metrics
                asd = calc_metrics(seg_f, sampler.warp_segmentation(), dict(enumerate(STRUCTURE_LABELS)), (1.0, 1.0, 1.0))[0][0]
                out['asd_mean_last_sample_voxels'] = float(asd[asd < float('inf')].mean())
                out['asd'] = 'host, scipy distance transform of the label contours, outside the timed region (parity unpinned)'
            except Exception as exc:
                out['asd'] = f'unavailable ({type(exc).__name__}: {exc})'[:200]
    if vi:
        # configs[0] says "VI warm start then 1 SGLD chain": the VI iteration of reference trainer.py:119-171 on the fused device
        # path (two antithetic samples through the step's operators + closed-form entropy + field-sized Adam), graph replays
        from irsgmcmc_b200.vi import VIWarmStart
        w = VIWarmStart(fixed, moving, vp, cfg, device=dev)
        w.sampler.hyper.copy_(sampler.hyper)
        w.step(5)
        barrier()
        n_vi = 50
        e0.record(); w.step(n_vi); e1.record()
        barrier()
        ms_vi = allmax(e0.elapsed_time(e1)) / n_vi
        out['vi'] = {'ms_per_iteration': ms_vi, 'iterations_per_s': 1e3 / ms_vi, 'iterations_timed': n_vi,
                     'gpu_launches_per_iteration': w.sampler.launches_per_step() + 3}
        if rank == 0 and cpu_vi:
            import time as _t
            from oracle import sgld_oracle as O
            st = O.State(O.Config(data=data, K=4 if lcc else 1, reg='lognormal' if lcc else 'l2', w_reg=1.6 if lcc else 1.4),
                         torch.zeros(1, 3, n, n, n), torch.ones(1, 3, n, n, n), (n, n, n))
            st.init_gmm(0.7)
            torch.set_num_threads(os.cpu_count() or 1)
            t0 = _t.perf_counter()
            O.vi_iteration(st, fixed, moving, vp, torch.randn(1, 3, n, n, n), torch.randn(1), torch.rand(1, 3, n, n, n),
                           torch.rand(1, 3, n, n, n))
            out['vi']['cpu_port_ms_per_iteration'] = 1e3 * (_t.perf_counter() - t0)
            out['vi']['cpu_port_threads'] = torch.get_num_threads()
        del w
    if moments:
        sampler.accumulate()
        sampler.posterior_moments()          # warm: buffers allocated, NCCL channels for this size set up
        barrier()
        e0.record(); sampler.accumulate(); e1.record()
        barrier()
        out['welford_update_ms'] = allmax(e0.elapsed_time(e1))
        barrier()
        e0.record(); mom = sampler.posterior_moments(); e1.record()
        barrier()
        out['moments_merge_ms'] = allmax(e0.elapsed_time(e1))
        out['moments_payload_bytes'] = 2 * 4 * 4 * V
        out['moments_n'] = int(mom['n'])
    del sampler
    gc.collect()
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='own', choices=['own', 'reference'])
    ap.add_argument('--size', type=int, default=128)
    ap.add_argument('--chains', type=int, default=1, help='chains per GPU')
    ap.add_argument('--data', default='lcc', choices=['lcc', 'ssd'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--hyper-mode', default='reference', choices=['reference', 'per_chain', 'frozen'])
    ap.add_argument('--as-rank', type=int, default=None, help='development: use the chain ids / seeds of this rank on one GPU')
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--e2e-steps', type=int, default=0, help='default: min(steps, 50)')
    ap.add_argument('--aten-gpu-baseline', action='store_true', help='(default at N = 1) also time the oracle port of the '
                    'reference with all tensors on the GPU (ATen kernels): the GPU-vs-GPU comparison of SURVEY section 8(d)')
    ap.add_argument('--no-aten-gpu-baseline', action='store_true')
    ap.add_argument('--no-configs', action='store_true', help='skip the `configs` sub-record (the other BASELINE.json '
                    'configurations, measured after the headline in the same run)')
    ap.add_argument('--cps', type=int, default=0, help='secondary configuration: SVFFD_3D with this control point '
                    'spacing as the transformation model (reference configs/experiment5); 0 = SVF_3D, the headline')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    from irsgmcmc_b200.sampler import SGLDSampler, SGLDConfig
    from irsgmcmc_b200.data_loader.synthetic import make_pair
    from irsgmcmc_b200 import parallel

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device: there is no CPU path (use --impl reference for the CPU baseline)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    cores_per_rank = pin_rank_to_cores(local_rank, int(os.environ.get('LOCAL_WORLD_SIZE', world)))
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)

    n, C = args.size, args.chains
    V = n ** 3
    if cores_per_rank:
        torch.set_num_threads(cores_per_rank)   # host-side tensor code must not oversubscribe this rank's cores
    fixed, moving, vp = make_pair(n, device=dev)   # filtering on the GPU; the pair itself lives on the host like a loader's
    reg = 'RegLoss_LogNormal' if args.data == 'lcc' else 'RegLoss_L2'
    ffd = dict(transformation='SVFFD_3D', cps=(args.cps,) * 3) if args.cps else {}
    cfg = SGLDConfig(data_loss=args.data, reg_loss=reg, w_reg=1.6 if args.data == 'lcc' else 1.4,
                     reg_learnable=args.data == 'lcc', hyper_mode=args.hyper_mode, **ffd)
    if args.cps:   # the variational parameters live on the control grid (reference data_loader/datasets.py:23-27,57-68)
        from irsgmcmc_b200.utils import get_control_grid_size
        gdims = (1, 3, *get_control_grid_size((n, n, n), cfg.cps))
        vp = {'mu': torch.zeros(gdims), 'log_var': torch.full(gdims, math.log(0.5 ** 2)), 'u': torch.full(gdims, 0.1)}
    # Weak scaling = the same work on every GPU: all ranks start from the SAME draw of q(v) (seed 123) and differ in their
    # Langevin / jitter streams (Philox keyed by the global chain id).  With a different draw per rank the forward squaring steps
    # differ by up to 16 % between ranks through the sign pattern of the velocity field alone (the rank-1 term x u of the draw
    # makes the signs more or less coherent, which changes the shared-memory bank conflicts of the gather; same max |u|, same
    # kernels: profiles/r2_rank_data_dependence.txt) -- the N > 1 curve then measures the luck of the draws, not the system.
    # --as-rank R (development) reproduces rank R's old per-rank draw on one GPU.
    data_rank = rank if args.as_rank is None else args.as_rank
    sampler = SGLDSampler(fixed, moving, C, cfg, device=dev, chain_offset=data_rank * C)
    gen = torch.Generator(device=dev).manual_seed(123 + (0 if args.as_rank is None else args.as_rank))
    sampler.init_chains('VI', vp, generator=gen)
    sampler.init_gmm()
    use_graph = not args.no_graph

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ----
    # The timed region is exactly K transitions, replayed as CUDA graphs of `g` transitions each (one replay for K <= 40): all
    # state lives on the device, so the host issues K / g launches and cannot perturb the region (at 8 ranks x 20 single-step
    # replays a 1 ms host hiccup on any rank cost 5 %).  Warm-up: W transitions as asked, topped up to >= 20 and to whole
    # replays of the same graph, so that clocks have ramped and the graph is resident.
    g_len = graph_length(args.steps) if use_graph else 1
    if use_graph:
        sampler.capture(1)
        sampler.capture(g_len)
    n_warm = max(args.warmup, 20)
    n_warm = -(-n_warm // g_len) * g_len
    # every rank samples its own GPU's clocks (the regions are graph replays: the poller cannot delay them).  The sampler
    # starts before the warm-up -- the same transitions, back to back with the timed ones -- because nvidia-smi needs ~0.1 s to
    # deliver its first line and a 20-transition region lasts 18 ms; rank 0's record is the line's `clocks`, the per-rank
    # medians show whether a slow rank is a slow GPU
    clocks = ClockSampler(local_rank)
    t_load = time.perf_counter()
    while True:    # at least the asked warm-up, and at least 0.6 s under load so that the clock samples are taken under load
        sampler.step(n_warm, use_graph=use_graph)
        torch.cuda.synchronize()
        if time.perf_counter() - t_load > 0.6:
            break
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sampler.step(args.steps, use_graph=use_graph)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clock_info = clocks.stop()
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    per_rank_ms, per_rank_mhz = [ms], [clock_info.get('sm_mhz')]
    if world > 1:
        mine = torch.tensor([ms, float(clock_info.get('sm_mhz') or 0)], device=dev, dtype=torch.float64)
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        per_rank_ms = [float(x[0].item()) for x in gathered]
        per_rank_mhz = [int(x[1].item()) for x in gathered]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * C * V * args.steps / (ms_max * 1e-3)

    # ---- end to end through the public API with host buffers ----
    e2e_steps = args.e2e_steps or min(args.steps, 50)
    pin = lambda x: x.contiguous().pin_memory()
    h_fixed, h_moving, h_mask = pin(fixed['im']), pin(moving['im']), pin(fixed['mask'].view(torch.uint8))
    h_stats = torch.empty(C, 8, dtype=torch.float64).pin_memory()
    h2d = h_fixed.numel() * 4 + h_moving.numel() * 4 + h_mask.numel()
    d2h = h_stats.numel() * 8

    # Every step's image pair crosses PCIe inside the timed region and every step's result is read back.  The sampler's
    # input pipeline is double-buffered: the upload of step i+1's pair runs on a copy stream while step i computes
    # (prefetch_images / commit_images).  The read-back is pipelined by one step: the statistics of step i are copied to
    # pinned memory behind the transition and waited for after step i+1 has been enqueued, so the GPU never idles on the
    # host.  `serial_value` below is the fully synchronous closed loop (upload, compute, read, wait) for comparison.
    h_stats2 = [h_stats, torch.empty_like(h_stats).pin_memory()]
    ev = [torch.cuda.Event(), torch.cuda.Event()]

    def e2e_step(i):
        sampler.commit_images()                                  # the uploaded pair becomes the resident set (pointer swap)
        sampler.prefetch_images(h_fixed, h_moving, h_mask)       # next step's pair: pinned host -> device, overlapped
        sampler.step(1, use_graph=use_graph)
        h_stats2[i & 1].copy_(sampler.stats, non_blocking=True)  # loss terms / alpha / energy of every chain
        ev[i & 1].record()
        if i > 0:
            ev[(i - 1) & 1].synchronize()                        # the caller reads the previous step's result

    sampler.prefetch_images(h_fixed, h_moving, h_mask)
    for i in range(4):
        e2e_step(i)
    barrier()
    e0.record()
    for i in range(e2e_steps):
        e2e_step(i + 4)
    ev[(e2e_steps + 3) & 1].synchronize()                        # ... and the last one
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * C * V * e2e_steps / (float(t.item()) * 1e-3)

    # the same without overlap (upload, then compute, on one stream) for reference
    def e2e_serial_step():
        sampler.load_images(h_fixed, h_moving, h_mask)
        sampler.step(1, use_graph=use_graph)
        h_stats.copy_(sampler.stats, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(3):
        e2e_serial_step()
    barrier()
    e0.record()
    for _ in range(e2e_steps):
        e2e_serial_step()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_serial_value = world * C * V * e2e_steps / (float(t.item()) * 1e-3)

    # ---- dominant kernel: one SVF adjoint step (svf_step_bwd_tma2_kernel), CUDA events around the 12-step adjoint ----
    peak, peak_src = measured_peaks()
    stage_ms = {k: 0.0 for k in sampler.STAGES}
    n_prof = 5
    for _ in range(n_prof):
        for k, v in sampler.profile_stages().items():
            stage_ms[k] += v / n_prof
    svf_steps = cfg.svf_steps
    kernel_ms = stage_ms['svf_adjoint'] / svf_steps
    achieved = SVF_BWD_BYTES_PER_VOXEL * C * V / (kernel_ms * 1e-3) / 1e9
    bytes_step = BYTES_PER_VOXEL_STEP_LCC if args.data == 'lcc' else BYTES_PER_VOXEL_STEP_SSD
    if args.cps:   # proposal (36 B) and update (60 B) act on the control grid; the FFD writes / its adjoint reads 12 B per voxel
        bytes_step = bytes_step - 96 + 24 + 96.0 * sampler.v[0, 0].numel() / V
    step_gbs = bytes_step * (value / world) / 1e9

    # ---- posterior moments: Welford update + NCCL merge (once per run / log period; outside the metric), timed WARM ----
    sampler.accumulate()
    sampler.posterior_moments()
    barrier()
    e0.record()
    mom = sampler.posterior_moments()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    merge_ms = float(t.item())
    merge_bytes = 2 * 4 * 4 * V      # two phases x (3 + 1) fp32 volumes
    # ring-equivalent bus bandwidth of the two all-reduces (NCCL convention: 2 (N-1)/N x bytes / time)
    merge_bus_gbs = (2.0 * (world - 1) / world) * merge_bytes / (merge_ms * 1e-3) / 1e9 if world > 1 else None

    # ---- the other BASELINE.json configurations, same run, same rules (the headline above is not touched by them) ----
    configs = None
    default_headline = args.size == 128 and args.chains == 1 and args.data == 'lcc' and not args.cps and use_graph
    if default_headline and not args.no_configs:
        configs = []
        todo = [dict(tag='configs[0]: 64^3 SSD + RegLoss_L2, VI warm start iterations and 1 SGLD chain per GPU', n=64, chains=1, data='ssd',
                     flush_l2=True, target_ms=300.0, vi=True, cpu_vi=world == 1 and not args.no_cpu_baseline),
                dict(tag=f'configs[2]: 128^3 LCC, 64 chains sharded over {world} GPU(s) (strong scaling), Welford + NCCL merge',
                     n=128, chains=max(64 // world, 1), data='lcc', moments=True),
                dict(tag='configs[3]: 256^3 LCC, 1 chain per GPU, segmentation warp + Dice per transition', n=256, chains=1,
                     data='lcc', seg_dice=True, target_ms=600.0)]
        if world == 1:
            todo += [dict(tag='configs[4]a: 1024 chains at 64^3 SSD', n=64, chains=1024, data='ssd'),
                     dict(tag='configs[4]b: 16 chains at 256^3 LCC', n=256, chains=16, data='lcc')]
        for kw in todo:
            tag = kw.pop('tag')
            try:
                configs.append(run_config(tag, kw.pop('n'), kw.pop('chains'), kw.pop('data'), world, rank, dev, peak, barrier, **kw))
            except Exception as exc:   # a secondary configuration must not cost the headline line
                configs.append({'config': tag, 'unavailable': f'{type(exc).__name__}: {exc}'[:300]})
                if world > 1:
                    raise

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:   # the contract: rank 0 at N = 1 only (at N > 1 the ranks are
        n_cpu = n if n <= 128 else 128                          # pinned to a share of the cores: a CPU baseline there is meaningless)
        v_cpu, ms_cpu, threads = oracle_transition_rate(n_cpu, 2, 0, 1, args.data, args.cps)
        cpu = {'value': v_cpu, 'unit': 'voxel-steps/s', 'cores': threads, 'kind': 'port',
               'sample': f'2 oracle transitions (oracle/sgld_oracle.py, torch CPU fp32, {threads} threads) at {n_cpu}^3, '
                         f'1 chain, {ms_cpu:.0f} ms each'}

    aten_gpu = None
    if rank == 0 and (args.aten_gpu_baseline or (world == 1 and not args.no_aten_gpu_baseline)):
        try:
            v_gpu, ms_gpu = aten_gpu_transition_rate(n if n <= 128 else 128, 3, 1, args.data, args.cps, device=str(dev))
            aten_gpu = {'value': v_gpu, 'unit': 'voxel-steps/s', 'ms_per_step': ms_gpu, 'kind': 'port on the GPU (ATen)',
                        'sample': '3 timed + 1 warm-up oracle transitions, 1 chain, fp32, wall clock around synchronize'}
        except Exception as exc:   # a baseline must not cost the bench line
            aten_gpu = {'unavailable': f'{type(exc).__name__}: {exc}'[:300]}

    if rank == 0:
        line = {'metric': 'SGLD voxel-steps/s', 'value': value, 'unit': 'voxel-steps/s', 'n_gpus': world,
                'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_max / args.steps,
                'iterations_per_s': args.steps / (ms_max * 1e-3), 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(args, C, n),
                'e2e': {'value': e2e_value, 'unit': 'voxel-steps/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                        'steps': e2e_steps, 'pipeline': 'two resident image sets: the upload of the next pair runs on a copy stream while this step computes, '
                                    'committing it is a pointer swap (a CUDA graph per set); per-step read-back pipelined by one step',
                        'serial_value': e2e_serial_value},
                'gpu_launches': sampler.launches_per_step() * args.steps,
                'roofline': {'bound': 'hbm', 'kernel': 'svf_step_bwd_tma2_kernel', 'achieved': achieved, 'peak': peak,
                             'unit': 'GB/s', 'frac': achieved / peak,
                             'traffic': NCU_TRAFFIC_BYTES.get((n, C, args.data)), 'peak_source': peak_src,
                             'algorithmic_bytes_per_launch': SVF_BWD_BYTES_PER_VOXEL * C * V, 'kernel_ms': kernel_ms},
                'step_roofline': {'bytes_per_voxel_step': bytes_step, 'achieved_gbs_per_gpu': step_gbs,
                                  'frac': step_gbs / peak},
                'stage_ms': {k: round(v, 4) for k, v in stage_ms.items()},
                'moments_merge_ms': merge_ms, 'moments_merge_bytes': merge_bytes, 'moments_merge_bus_gbs': merge_bus_gbs,
                'svf_max_abs_u_per_step': [round(float(x), 4) for x in sampler._maxabs[:cfg.svf_steps].tolist()],
                'graph': use_graph, 'transitions_per_graph': g_len, 'per_rank_ms_per_step': [x / args.steps for x in per_rank_ms], 'per_rank_sm_mhz': per_rank_mhz,
                'host_cores_per_rank': cores_per_rank, 'clocks': clock_info, 'cpu_baseline': cpu}
        if configs is not None:
            line['configs'] = configs
        if aten_gpu is not None:
            line['aten_gpu_baseline'] = aten_gpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
