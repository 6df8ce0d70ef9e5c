import os, sys, ctypes
import torch
sys.path.insert(0, '.')
from diffusion_ccsp_b200 import _abi, scenes, synthetic
from diffusion_ccsp_b200.ddpm import GaussianDiffusion
from diffusion_ccsp_b200.denoise_fn import ConstraintDiffuser
lib = _abi.load_library(); lib.ccsp_debug_trap_info.restype = ctypes.c_uint64
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4
T, K = 6, 3
dims = synthetic.DIMS['qualitative']
batch = scenes.qualitative_batch(B, N)
den = ConstraintDiffuser(dims=dims, input_mode='qualitative', device='cuda', verbose=False, math='bf16x3')
gd = GaussianDiffusion(den, timesteps=T, EBM='ULA', samples_per_step=K).eval()
gd.load_state_dict(synthetic.load_trained_checkpoint(), strict=False)
os.environ['CCSP_PERSIST'] = '0'
a = gd.sample(batch, seed=3); torch.cuda.synchronize(); print('plain ok', float(a.abs().max()), flush=True)
os.environ['CCSP_PERSIST'] = '1'
try:
    b = gd.sample(batch, seed=3); torch.cuda.synchronize()
    print('persistent ok, equal:', torch.equal(a, b), 'launches', _abi.launch_count(), flush=True)
except Exception as ex:
    info = int(lib.ccsp_debug_trap_info())
    print('FAILED', repr(ex)[:120], '\ntrap info: code/line', info >> 40, 'block', (info >> 24) & 0xFFFF, 'thread', info & 0xFFFFFF, flush=True)
if os.environ.get('CCSP_PERSIST_TRACE'):
    lib.ccsp_debug_persist_trace.restype = ctypes.c_uint64
    lib.ccsp_debug_persist_trace.argtypes = [ctypes.c_int, ctypes.c_int]
    names = ['node: iteration top', 'node: edge flag seen', 'node: update done', 'node: signalled', 'edge: node flag seen', 'edge: GEMM1 committed',
             'edge: D1 ready (epilogue)', 'edge: signalled']
    tr = [[int(lib.ccsp_debug_persist_trace(e, i)) for i in range(32)] for e in range(8)]
    # node iteration i (>= 1) follows edge evaluation i - 1; print relative to the node iteration's flag
    for i in range(4, 9):
        base = tr[4][i]      # edge evaluation i starts when it sees node iteration i's flag... node iteration i signalled at tr[3][i]
        print(f'eval {i}: node signalled {tr[3][i] - base:+6d} | edge sees flag 0 | GEMM1 committed {tr[5][i] - base:+6d} | D1 ready {tr[6][i] - base:+6d} | '
              f'edge signalled {tr[7][i] - base:+6d} || node(i+1): top {tr[0][i + 1] - base:+6d} flag seen {tr[1][i + 1] - base:+6d} update {tr[2][i + 1] - base:+6d} '
              f'signalled {tr[3][i + 1] - base:+6d} | next edge flag {tr[4][i + 1] - base:+6d}  (ns)')
    base = tr[4][5]
    ch = [int(lib.ccsp_debug_persist_trace(8, k)) - base for k in range(16)]
    ep = [int(lib.ccsp_debug_persist_trace(9, k)) - base for k in range(8)]
    print('evaluation 5, CTA 0: ring stage seen full by the GEMM1 issuer, chunk 0..15 (ns after the node flag):', ch)
    print('   epilogue thread 0: D1 in registers', ep[0], '| decoder chunk 0 / 1 published', ep[1], ep[2], '| D2 seen ready (units 0..3 mod 4)', ep[3:7])
