"""Host-side logic of bench.py that needs no GPU: the sharding of a decode batch over ranks (independent videos, no
collective) and the CPU reference arms' JSON contract."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench


@pytest.mark.parametrize('world', [1, 2, 3, 4, 8])
def test_shard_range_covers_every_video_once(world):
    for batch in (1, 2, 5, 7, 8, 64, 1000, 1024):
        parts = [bench.shard_range(batch, r, world) for r in range(world)]
        covered = [v for lo, hi in parts for v in range(lo, hi)]
        assert covered == list(range(batch))                                  # contiguous, ordered, nothing twice
        sizes = [hi - lo for lo, hi in parts]
        assert max(sizes) == (batch + world - 1) // world and all(s >= 0 for s in sizes)
        assert all(lo <= hi for lo, hi in parts)


def test_reference_arm_of_the_beam_workload_prints_the_contract_line():
    """bench.py --impl reference --workload beam: the reference's batch-1 beam loop (NumPy restatement) on the host cores."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--workload', 'beam', '--steps', '1', '--warmup', '1',
                          '--ref-videos', '1', '--frames', '5'], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'beam5_decode_captions_per_s' and d['unit'] == 'captions/s'
    assert d['value'] > 0 and d['higher_is_better'] is True and d['scaling'] == 'strong'
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'captions/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def test_non_zero_ranks_of_the_reference_arms_exit_without_work():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2')
    for extra in ([], ['--workload', 'beam']):
        out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference'] + extra, capture_output=True, text=True, timeout=300,
                             cwd=ROOT, env=env)
        assert out.returncode == 0 and out.stdout.strip() == ''
