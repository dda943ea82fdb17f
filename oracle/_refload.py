"""Import the UNMODIFIED reference NeRF++ modules in this container (test infrastructure only).

The reference (``/root/reference/nerf-methods/nerfplusplus``) imports a few non-numeric packages at
module scope that are not installed here (matplotlib, tensorboardX, imageio, configargparse;
``utils.py:37-41``, ``ddp_train_nerf.py:13,17``).  They are replaced by the import-only stand-ins of
``outdoor-nerf-depth_b200/compat``; no arithmetic goes through them.  Used by
``oracle/gen_golden.py`` (from ``/root/reference``, build container only) to produce ``tests/golden/*.npz`` and by
``bench.py``'s CPU legs / ``oracle/ref_harness.py``, which on the GPU box find the git-ignored copy
``baseline/_ref/nerfplusplus`` made by ``oracle/install_reference.py`` (and fall back to the oracle port without it).
"""
import importlib
import os
import sys
import types

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CANDIDATES = (os.environ.get("NERFPP_REFERENCE"), "/root/reference/nerf-methods/nerfplusplus",
               os.path.join(_ROOT, "baseline", "_ref", "nerfplusplus"))      # the copy oracle/install_reference.py makes
REF_ROOT = next((p for p in _CANDIDATES if p and os.path.isdir(p)), _CANDIDATES[1])


def available():
    return os.path.isdir(REF_ROOT)


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_NAMES = ("ddp_model", "depth_loss", "utils", "nerf_network", "ddp_train_nerf", "data_loader_split", "nerf_sample_ray_split")
_COMPAT = os.path.join(_ROOT, "outdoor-nerf-depth_b200", "compat")
_loaded = None


def load_reference(isolated=True):
    """Returns (ddp_train_nerf, ddp_model, depth_loss, utils) -- the reference's modules, unmodified.

    The product ships drop-in modules with the SAME names (``ddp_model``, ``depth_loss``, ``data_loader_split``).  With
    ``isolated`` (default) the reference is imported while those names are taken out of ``sys.modules`` / its directory is
    first on ``sys.path``, and both are restored afterwards: the returned module objects keep working (their globals are
    bound), while ``import ddp_model`` elsewhere in the process still means the product's module."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    for name in ("matplotlib", "tensorboardX", "imageio", "configargparse"):      # non-numeric imports of the reference
        try:
            importlib.import_module(name)
        except Exception:
            if _COMPAT not in sys.path:
                sys.path.append(_COMPAT)                                              # the same stand-ins the launcher uses
            importlib.import_module(name)
    saved_mods = {n: sys.modules.pop(n) for n in _NAMES if n in sys.modules}
    saved_path = list(sys.path)
    sys.path.insert(0, REF_ROOT)
    try:
        import ddp_train_nerf, ddp_model, depth_loss, utils  # noqa: E401
        for m in (ddp_train_nerf, ddp_model, depth_loss, utils):
            assert m.__file__.startswith(REF_ROOT), m.__file__
        _loaded = (ddp_train_nerf, ddp_model, depth_loss, utils)
    finally:
        if isolated:
            for n in _NAMES:
                sys.modules.pop(n, None)
            sys.modules.update(saved_mods)
            sys.path[:] = saved_path
    return _loaded
