"""CPU: the reference arm of bench.py prints one JSON line with the contract's
keys (the GPU arm shares the metric / unit / config strings with it)."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_json_line():
    out = subprocess.run(
        [sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
         '--steps', '1', '--warmup', '1', '--J_space', '6'],
        capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'DoF-applies/s'
    assert d['metric'].startswith('space-time DoF-applies/sec')
    assert d['higher_is_better'] is True and d['value'] > 0
    assert d['steps'] == 1 and d['warmup'] == 1 and d['n_gpus'] == 1
    assert d['config']['workload'].startswith('BASELINE.json configs[3]')
    # the unmodified reference classes (oracle/_ref, copied by build()) when
    # present, else the oracle port; the sample is stated, not implied
    assert d['cpu_baseline']['kind'] in ('reference', 'port')
    assert d['cpu_baseline']['cores'] >= 1
    smp = d['config']['sample']
    assert smp['J_space'] == 6 and smp['J_time'] <= d['config']['J_time']
    assert smp['ranks'] == d['cpu_baseline']['cores']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'],
                        'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2')
    out = subprocess.run(
        [sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
         '--gpus', '2', '--steps', '1', '--warmup', '1'],
        capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ''
