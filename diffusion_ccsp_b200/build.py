"""In-tree build of libccsp_b200.so (the C-ABI library) with nvcc for sm_100a.

    python -m diffusion_ccsp_b200.build [--force]

The .so lands in diffusion_ccsp_b200/lib/ (git-ignored, but it travels to the GPU box with the
gpurun snapshot).  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, 'csrc')
LIBDIR = os.path.join(PKG, 'lib')
LIB = os.path.join(LIBDIR, 'libccsp_b200.so')

NVCC_FLAGS = [
    '-O3', '-std=c++17', '-lineinfo',
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-Xcompiler', '-fPIC', '-shared',
    '-Xptxas', '-v',
]
if os.environ.get('CCSP_EXTRA_DEFS'):          # experiments: extra -D flags
    NVCC_FLAGS += ['-D' + d for d in os.environ['CCSP_EXTRA_DEFS'].split()]
if os.environ.get('CCSP_DEBUG') == '1':      # development builds: mbarrier spin loops trap after 2^22 polls instead of hanging
    NVCC_FLAGS.append('-DCCSP_DEBUG_SPIN_TRAP')


def _nvcc() -> str:
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found')


def _sources():
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith('.cu')]
    deps = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if os.path.isfile(os.path.join(CSRC, f))]
    deps.append(os.path.join(ROOT, 'include', 'ccsp_b200.h'))
    return srcs, deps


def _fingerprint(deps) -> str:
    h = hashlib.sha256(' '.join(NVCC_FLAGS).encode())
    for d in deps:
        with open(d, 'rb') as f:
            h.update(f.read())
    return h.hexdigest()


def build_library(force: bool = False, verbose: bool = False) -> str:
    srcs, deps = _sources()
    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, 'libccsp_b200.stamp')
    fp = _fingerprint(deps)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == fp:
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + ['-I', os.path.join(ROOT, 'include'), '-o', LIB] + srcs
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(LIBDIR, 'build.log')
    with open(log, 'w') as f:
        f.write(' '.join(cmd) + '\n' + r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f'nvcc failed (exit {r.returncode}); see {log}')
    if verbose:
        print(r.stdout + r.stderr)
    with open(stamp, 'w') as f:
        f.write(fp)
    return LIB


def build_tc_test() -> str:
    """Standalone numerics/throughput harness for the tcgen05 GEMM kernels (csrc/tests/tc_gemm_test.cu)."""
    src = os.path.join(CSRC, 'tests', 'tc_gemm_test.cu')
    out = os.path.join(LIBDIR, 'tc_gemm_test')
    os.makedirs(LIBDIR, exist_ok=True)
    cmd = [_nvcc(), '-O3', '-std=c++17', '-lineinfo', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', out, src]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError('nvcc failed for tc_gemm_test')
    return out


if __name__ == '__main__':
    path = build_library(force='--force' in sys.argv, verbose=True)
    print('built', path)
    if '--tests' in sys.argv:
        print('built', build_tc_test())
