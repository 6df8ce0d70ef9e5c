"""CPU: the reference arm of bench.py (the oracle port on the host cores) prints ONE JSON line with the keys of the
contract, and the GPU arm refuses to run without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), *args], capture_output=True, text=True, env=e, timeout=600)


def test_reference_arm_json_contract():
    # a small sample of the workload so the test stays in seconds; torchrun-style env to check the thread override
    r = run_bench('--impl', 'reference', '--steps', '1', '--warmup', '0', '--batch', '4', env={'OMP_NUM_THREADS': '1'})
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'scenes_per_sec' and d['unit'] == 'scenes/s'
    assert d['higher_is_better'] is True and d['value'] > 0 and d['n_gpus'] == 1
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == dict(value=d['value'], unit='scenes/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert 'workload' in d['config'] and 'RandomSplitQualitativeWorld' in d['config']['workload']


def test_reference_arm_other_ranks_exit_quietly():
    r = run_bench('--impl', 'reference', '--steps', '1', '--warmup', '0', '--batch', '4', '--gpus', '2', env={'RANK': '1', 'WORLD_SIZE': '2'})
    assert r.returncode == 0 and r.stdout.strip() == ''


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_gpu_arm_needs_a_gpu():
    r = run_bench('--steps', '1', '--warmup', '0')
    assert r.returncode != 0 and 'no CPU fallback' in (r.stderr + r.stdout)
