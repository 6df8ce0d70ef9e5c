"""CPU: the reference arm of bench.py (the UNMODIFIED reference, staged under oracle/_ref, on the host cores; the numpy port
only when it is missing) prints ONE JSON line with the keys of the
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
    from oracle import ref_shim
    want = 'reference' if ref_shim.reference_available() else 'port'
    assert d['cpu_baseline']['kind'] == want
    assert d['cpu_baseline']['cores'] == (os.cpu_count() or 1), 'the CPU arm must use every host core, whatever OMP_NUM_THREADS torchrun exports'
    assert d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == dict(value=d['value'], unit='scenes/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert 'workload' in d['config'] and 'RandomSplitQualitativeWorld' in d['config']['workload']


def test_reference_arm_other_ranks_exit_quietly():
    r = run_bench('--impl', 'reference', '--steps', '1', '--warmup', '0', '--batch', '4', '--gpus', '2', env={'RANK': '1', 'WORLD_SIZE': '2'})
    assert r.returncode == 0 and r.stdout.strip() == ''


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_gpu_arm_needs_a_gpu():
    r = run_bench('--steps', '1', '--warmup', '0')
    assert r.returncode != 0 and 'no CPU fallback' in (r.stderr + r.stdout)


def test_reference_arm_falls_back_to_the_port_and_says_so(tmp_path):
    """without /root/reference and without oracle/_ref the arm still answers, labelled kind='port'"""
    r = run_bench('--impl', 'reference', '--steps', '1', '--warmup', '0', '--batch', '4', env={'CCSP_REFERENCE_ROOT': str(tmp_path)})
    assert r.returncode == 0, r.stderr
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d['cpu_baseline']['kind'] == 'port' and 'make_ref' in d['cpu_baseline']['code']


def test_staged_reference_is_byte_identical_to_the_checkout():
    """oracle/_ref holds unmodified copies (sha256 in the manifest); the check against the checkout runs in the build container"""
    from oracle import make_ref
    dst = make_ref.make_ref(verbose=False)
    if dst is None:
        pytest.skip('no reference checkout and nothing staged')
    man = json.load(open(os.path.join(dst, 'MANIFEST.json')))
    assert set(man) == set(make_ref.FILES)
    for rel, info in man.items():
        assert make_ref._sha(os.path.join(dst, rel)) == info['sha256']
        src = os.path.join(make_ref.REF_SRC, rel)
        if os.path.isfile(src):
            assert make_ref._sha(src) == info['sha256'], rel
