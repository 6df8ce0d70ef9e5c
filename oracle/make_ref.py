"""TEST / BASELINE INFRASTRUCTURE ONLY — recipe that stages the UNMODIFIED reference sources of the hot path
under oracle/_ref/ so that they travel to the GPU box (oracle/_ref/ is git-ignored, NOT gpurun-ignored).

    python -m oracle.make_ref            # copies, prints a manifest with sha256 per file

Nothing is edited: files are byte-for-byte copies of /root/reference/{networks,envs}/... (sha256 recorded in
oracle/_ref/MANIFEST.json next to the reference commit-less path they came from).  They are loaded through
oracle/ref_shim.py (stub modules for the absent third-party imports) by

  * bench.py --impl reference / the cpu_baseline leg       (the reference's own CPU implementation, kind "reference")
  * tests/ (goldens, oracle pinning)

and never by the product package.  The reference has no build system for these files (plain Python), so the
"build" is the copy.  If /root/reference is absent (the GPU box) this script is a no-op and the staged copy is used.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get('CCSP_REFERENCE_SRC', '/root/reference')
DST = os.path.join(HERE, '_ref')

# the sampling path (networks/ddpm.py, networks/denoise_fn.py), what it imports (envs/data_utils.py), and the reference's
# own scene generator / labeller / batch transform used as oracles for the N1 and N3 rows
FILES = [
    'networks/ddpm.py',
    'networks/denoise_fn.py',
    'networks/data_transforms.py',
    'envs/data_utils.py',
    'envs/builders.py',
]


def _sha(path):
    h = hashlib.sha256()
    with open(path, 'rb') as f:
        h.update(f.read())
    return h.hexdigest()


def make_ref(verbose: bool = True) -> str | None:
    if not os.path.isfile(os.path.join(REF_SRC, 'networks', 'ddpm.py')):
        if verbose:
            print(f'[make_ref] {REF_SRC} not present; keeping staged copy' if os.path.isdir(DST) else
                  f'[make_ref] {REF_SRC} not present and nothing staged')
        return DST if os.path.isdir(DST) else None
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF_SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = dict(sha256=_sha(dst), bytes=os.path.getsize(dst), source=src)
        assert _sha(src) == manifest[rel]['sha256']
    with open(os.path.join(DST, 'MANIFEST.json'), 'w') as f:
        json.dump(manifest, f, indent=1)
    if verbose:
        print(f'[make_ref] staged {len(FILES)} unmodified reference files under {DST}')
    return DST


if __name__ == '__main__':
    make_ref()
