"""the N > 1 path on the CPU: chain sharding and the Welford moment merge over a world_size-2 gloo group"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sgld_oracle as O
from tests.util import rel


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, samples, counts, out_dir):
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from irsgmcmc_b200 import parallel
    offset, count = parallel.chain_shard(sum(counts))
    assert count == counts[rank] and offset == sum(counts[:rank])
    mine = samples[offset:offset + count].double()
    mean = mine.mean(0).float() if count else torch.zeros_like(samples[0])
    m2 = ((mine - mine.mean(0)) ** 2).sum(0).float() if count else torch.zeros_like(samples[0])
    mean2 = (2 * mine[:, :1]).mean(0).float() if count else torch.zeros_like(samples[0, :1])
    m22 = ((2 * mine[:, :1] - 2 * mine[:, :1].mean(0)) ** 2).sum(0).float() if count else torch.zeros_like(samples[0, :1])
    n, (gm, gm2), (gm_b, gm2_b) = parallel.merge_moments(count, [(mean, m2), (mean2, m22)])
    torch.save({'n': n, 'mean': gm, 'm2': gm2, 'mean_b': gm_b, 'm2_b': gm2_b}, os.path.join(out_dir, f'r{rank}.pt'))
    dist.destroy_process_group()


@pytest.mark.parametrize('counts', [(5, 4), (7, 0), (1, 1)])
def test_merge_moments_world_size_2(tmp_path, counts):
    torch.manual_seed(0)
    samples = torch.randn(sum(counts), 3, 6, 5, 4) * 2.0 + 1.0
    if counts == (7, 0):
        pytest.skip('chain_shard never produces an empty rank for totals >= world size')
    port = _free_port()
    mp.spawn(_worker, args=(2, port, samples, counts, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / 'r0.pt'), torch.load(tmp_path / 'r1.pt')
    mean, std = O.posterior_statistics(samples)
    for r in (r0, r1):
        assert r['n'] == sum(counts)
        assert rel(r['mean'], mean) < 1e-6 and rel(r['mean_b'], 2 * mean[:1]) < 1e-6
        if sum(counts) > 1:
            assert rel((r['m2'] / (r['n'] - 1)).sqrt(), std) < 1e-5
    assert torch.equal(r0['mean'], r1['mean']) and torch.equal(r0['m2'], r1['m2'])


def test_chain_shard_partition():
    from irsgmcmc_b200.parallel import chain_shard
    for total, world in ((64, 8), (64, 1), (10, 4), (3, 8)):
        spans = [chain_shard(total, r, world) for r in range(world)]
        assert sum(c for _, c in spans) == total
        pos = 0
        for off, c in spans:
            assert off == pos
            pos += c


def test_merge_moments_single_process():
    from irsgmcmc_b200.parallel import merge_moments
    m, m2 = torch.randn(3, 4), torch.rand(3, 4)
    n, (a, b) = merge_moments(5, [(m, m2)])
    assert n == 5 and torch.equal(a, m) and torch.equal(b, m2)
