"""TEST INFRASTRUCTURE ONLY — loader for the UNMODIFIED reference sampler.

Imports /root/reference/networks/{denoise_fn,ddpm}.py (and envs/{builders,data_utils}.py)
exactly as they lie on disk, after registering stub modules for the third-party
packages that are absent from this image.  It exists so that

  * tests/golden/make_golden.py can generate golden vectors from the real reference, and
  * tests/test_oracle_vs_reference.py can pin oracle/ccsp_oracle.py against it,

The sources are taken from /root/reference when it exists (the build container) and otherwise from the
byte-identical staged copy oracle/_ref/ made by `python -m oracle.make_ref` (git-ignored, travels to the GPU
box with the snapshot) — that is how `bench.py --impl reference` and the `cpu_baseline` leg time the
reference's OWN implementation on the GPU box's host cores.
Nothing in the product package (diffusion_ccsp_b200/) imports anything from oracle/.
"""
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REFERENCE_ROOT = os.environ.get("CCSP_REFERENCE_ROOT") or (
    "/root/reference" if os.path.isfile("/root/reference/networks/ddpm.py") else _STAGED)


def reference_kind() -> str:
    """'live' (the read-only checkout), 'staged' (oracle/_ref copy) or 'absent'."""
    if not reference_available():
        return "absent"
    return "staged" if os.path.abspath(REFERENCE_ROOT) == os.path.abspath(_STAGED) else "live"


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "networks", "ddpm.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_loaded = {}


class _DataStub:
    """torch_geometric.data.Data as the reference's transforms use it: a keyword bag (data_transforms.py:195-199)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def load_reference():
    """Returns (denoise_fn_module, ddpm_module) of the unmodified reference."""
    if "mods" in _loaded:
        return _loaded["mods"]
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")

    # jactorch.add_dim(t, dim, size): pure shape op, used at denoise_fn.py:328,334,397
    def add_dim(t, dim, size):
        return t.unsqueeze(dim).expand(*t.shape[:dim], size, *t.shape[dim:])

    _stub("ipdb")
    _stub("jactorch", add_dim=add_dim)
    _stub("jactorch.nn")
    _stub("matplotlib")
    _stub("matplotlib.pyplot")
    _stub("imageio")
    _stub("torch_geometric")
    _stub("torch_geometric.loader", DataLoader=object)
    _stub("torch_geometric.data", Data=_DataStub)
    for p in ("networks", "envs", ""):
        path = os.path.join(REFERENCE_ROOT, p) if p else REFERENCE_ROOT
        if path not in sys.path:
            sys.path.insert(0, path)
    import denoise_fn  # noqa: E402  (reference module)
    import ddpm  # noqa: E402  (reference module)
    _loaded["mods"] = (denoise_fn, ddpm)
    return _loaded["mods"]


def load_reference_envs():
    """Returns (builders, data_utils) of the unmodified reference (numpy/torch only)."""
    load_reference()
    import builders  # noqa: E402
    import data_utils  # noqa: E402
    return builders, data_utils


class injected_randn:
    """Context manager: serve torch.randn calls made by the reference from a pre-drawn tensor.

    The reference draws with torch.randn(shape, device=...) at ddpm.py:121-122, 273, 292;
    draw order = 1 (x_T) + per timestep [1 (p_sample) + K (ULA)]   (SURVEY.md §8a quirk 4).
    """

    def __init__(self, noise):
        self.noise = noise  # [n_draws, n, P] float32 tensor
        self.calls = 0

    def __enter__(self):
        import torch
        self._torch = torch
        self._real = torch.randn
        outer = self

        def fake(*shape, device=None, **kw):
            if len(shape) == 1 and not isinstance(shape[0], int):
                shape = tuple(shape[0])
            z = outer.noise[outer.calls]
            assert tuple(z.shape) == tuple(shape), (z.shape, shape)
            outer.calls += 1
            return z.clone()

        torch.randn = fake
        return self

    def __exit__(self, *exc):
        self._torch.randn = self._real
        self._torch.set_grad_enabled(True)  # p_sample_loop flips the global flag (ddpm.py:262-265)
        return False
