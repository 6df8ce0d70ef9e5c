"""Pipelined-chains probe: bit-equality against the launch-per-evaluation path and per-evaluation timing.
   python scripts/pipe_probe.py B N T K [node_ctas ...]"""
import ctypes
import os
import sys
import time

import torch

sys.path.insert(0, '.')
from diffusion_ccsp_b200 import _abi, scenes, synthetic
from diffusion_ccsp_b200.ddpm import GaussianDiffusion
from diffusion_ccsp_b200.denoise_fn import ConstraintDiffuser

lib = _abi.load_library()
lib.ccsp_debug_trap_info.restype = ctypes.c_uint64
B, N, T, K = (int(v) for v in sys.argv[1:5])
ctas = [int(v) for v in sys.argv[5:]] or [0]
dims = synthetic.DIMS['qualitative']
batch = scenes.qualitative_batch(B, N)


def same_bits(a, b):
    return torch.equal(a.contiguous().view(torch.int32), b.contiguous().view(torch.int32))


def model():
    den = ConstraintDiffuser(dims=dims, input_mode='qualitative', device='cuda', verbose=False, math='bf16x3')
    gd = GaussianDiffusion(den, timesteps=T, EBM='ULA', samples_per_step=K).eval()
    gd.load_state_dict(synthetic.load_trained_checkpoint(), strict=False)
    return den, gd


def timed(gd, seed, reps=2):
    best = 1e9
    out = None
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = gd.sample(batch, seed=seed)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return out, best * 1e3 / (T * (1 + K))


os.environ['CCSP_CHAINS'] = '1'
os.environ['CCSP_PERSIST'] = '0'
den0, gd0 = model()
ref, ms0 = timed(gd0, 3)
print(f'one chain, launch per evaluation: {ms0:.4f} ms/evaluation  max|x| {float(ref.abs().max()):.3f}', flush=True)
den0.drop_plans()
os.environ['CCSP_CHAINS'] = os.environ.get('PROBE_CHAINS', '2')
den, gd = model()
a, ms1 = timed(gd, 3)
print(f'chain-sorted plan, launch per evaluation: {ms1:.4f} ms/evaluation  equal: {same_bits(a, ref)}', flush=True)
os.environ.pop('CCSP_PERSIST')
for c in ctas:
    if c:
        os.environ['CCSP_PIPE_NODE_CTAS'] = str(c)
    else:
        os.environ.pop('CCSP_PIPE_NODE_CTAS', None)
    try:
        _abi.reset_launch_count()
        b, ms2 = timed(gd, 3)
        print(f'pipelined, {c} node CTAs: {ms2:.4f} ms/evaluation  equal: {same_bits(b, ref)}  launches {_abi.launch_count()}', flush=True)
    except Exception as ex:
        info = int(lib.ccsp_debug_trap_info())
        print('FAILED', repr(ex)[:160], '\ntrap info: code/line', info >> 40, 'block', (info >> 24) & 0xFFFF, 'thread', info & 0xFFFFFF, flush=True)
        break

if os.environ.get('CCSP_PERSIST_TRACE'):
    lib.ccsp_debug_persist_trace.restype = ctypes.c_uint64
    lib.ccsp_debug_persist_trace.argtypes = [ctypes.c_int, ctypes.c_int]
    tr = [[int(lib.ccsp_debug_persist_trace(e, i)) for i in range(32)] for e in range(8)]
    base = tr[4][4]
    nc = int(os.environ['CCSP_CHAINS'])
    print('chain evaluation q (CTA 0 of each kernel), us relative to q = 4:')
    for q in range(4, 20):
        f = lambda v: f'{(v - base) / 1e3:8.1f}'
        nq = q + nc          # the node iteration that consumes chain evaluation q
        print(f'q {q:2d} (chain {q % nc}): edge sees node flag {f(tr[4][q])} | D1 of first unit {f(tr[6][q])} | edge signalled {f(tr[7][q])} || '
              f'node it {nq}: top {f(tr[0][nq]) if nq < 32 else "-"} flag seen {f(tr[1][nq]) if nq < 32 else "-"} first block updated {f(tr[2][nq]) if nq < 32 else "-"} signalled {f(tr[3][nq]) if nq < 32 else "-"}')
