"""Import the UNMODIFIED reference NeRF++ modules in this container (test infrastructure only).

The reference (``/root/reference/nerf-methods/nerfplusplus``) imports a few non-numeric packages at
module scope that are not installed here (matplotlib, tensorboardX, imageio, configargparse;
``utils.py:37-41``, ``ddp_train_nerf.py:13,17``).  They are replaced by empty stub modules; no
arithmetic goes through them.  Used ONLY by ``oracle/gen_golden.py`` to
produce ``tests/golden/*.npz``; nothing under tests/, bench.py or the product imports this at run
time on the GPU box (the reference tree does not travel).
"""
import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("NERFPP_REFERENCE", "/root/reference/nerf-methods/nerfplusplus")


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load_reference():
    """Returns (ddp_train_nerf, ddp_model, depth_loss, utils) reference modules."""
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    if "matplotlib" not in sys.modules:
        mpl = _stub("matplotlib", use=lambda *a, **k: None)
        _stub("matplotlib.backends")
        _stub("matplotlib.backends.backend_agg", FigureCanvasAgg=object)
        _stub("matplotlib.figure", Figure=object)
        _stub("matplotlib.cm")
        _stub("matplotlib.pyplot")
        mpl.cm = sys.modules["matplotlib.cm"]
        mpl.colors = _stub("matplotlib.colors")
        mpl.colorbar = _stub("matplotlib.colorbar")
    for name in ("tensorboardX", "imageio", "configargparse"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                _stub(name, SummaryWriter=object)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    # make sure we do not pick up the drop-in shim modules of the same name
    for name in ("ddp_model", "depth_loss", "utils", "nerf_network", "ddp_train_nerf"):
        mod = sys.modules.get(name)
        if mod is not None and not getattr(mod, "__file__", "").startswith(REF_ROOT):
            del sys.modules[name]
    import ddp_train_nerf, ddp_model, depth_loss, utils  # noqa: E401
    for m in (ddp_train_nerf, ddp_model, depth_loss, utils):
        assert m.__file__.startswith(REF_ROOT), m.__file__
    return ddp_train_nerf, ddp_model, depth_loss, utils
