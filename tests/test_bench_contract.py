"""bench.py's reference arm on the CPU: the JSON line the driver parses (keys of the bench contract, tier section 4)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('extra', [[], ['--cps', '4']])
def test_reference_arm_prints_the_contract_line(extra):
    res = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--size', '16', '--steps', '1',
                          '--warmup', '1'] + extra, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'SGLD voxel-steps/s' and d['unit'] == 'voxel-steps/s'
    assert d['higher_is_better'] is True and d['scaling'] == 'weak' and d['vs_baseline'] is None and d['dtype'] == 'f32'
    assert d['steps'] == 1 and d['warmup'] == 1 and d['value'] > 0 and d['ms_per_step'] > 0 and d['data'] == 'synthetic'
    assert 'workload' in d['config'] and ('SVFFD_3D' in d['config']['workload']) == bool(extra)
    cb, e2e = d['cpu_baseline'], d['e2e']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == d['value'] and 'sample' in cb
    assert e2e == {'value': d['value'], 'unit': 'voxel-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def test_other_ranks_of_the_reference_arm_exit_without_work():
    res = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2'],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env={**os.environ, 'RANK': '1'})
    assert res.returncode == 0 and res.stdout.strip() == ''


def test_graph_length_and_core_pinning_helpers():
    """the timed region is K transitions in K / g graph replays with g | K wherever a divisor in [10, 40] exists; ranks get disjoint,
    equal shares of the cores"""
    sys.path.insert(0, ROOT)
    import bench
    assert bench.graph_length(20) == 20 and bench.graph_length(1) == 1 and bench.graph_length(40) == 40
    assert bench.graph_length(200) == 40 and bench.graph_length(100) == 25 and bench.graph_length(50) == 25
    assert bench.graph_length(97) == 10           # prime: ten-transition replays and an eager remainder
    for k in (20, 60, 200, 1000):
        assert k % bench.graph_length(k) == 0
    before = os.sched_getaffinity(0)
    try:
        per = bench.pin_rank_to_cores(0, 1)       # a single rank keeps every core
        assert per is None and os.sched_getaffinity(0) == before
        if len(before) >= 2:
            per = bench.pin_rank_to_cores(1, 2)
            mine = os.sched_getaffinity(0)
            assert per == len(before) // 2 and len(mine) == per and mine == set(sorted(before)[per:2 * per])
    finally:
        os.sched_setaffinity(0, before)


def test_own_arm_fails_loudly_without_a_gpu():
    """no CPU path: without CUDA the own arm raises instead of measuring anything"""
    import torch
    if torch.cuda.is_available():
        pytest.skip('needs a machine without a GPU')
    res = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '1'], capture_output=True, text=True,
                         timeout=300, cwd=ROOT)
    assert res.returncode != 0 and 'no CPU path' in res.stderr
