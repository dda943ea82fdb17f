"""Import-only stand-in for matplotlib (the reference's utils.py:37-41 imports it at module scope for its colour-bar
helpers, which only the every-50000-steps image dump calls).  Anything that would actually draw raises."""


def use(*a, **k):
    pass


class _Missing(object):
    def __init__(self, what):
        self._what = what

    def __getattr__(self, name):
        raise ImportError("matplotlib is not installed: %s.%s is a stand-in (outdoor-nerf-depth_b200/compat)" % (self._what, name))

    def __call__(self, *a, **k):
        raise ImportError("matplotlib is not installed: %s is a stand-in (outdoor-nerf-depth_b200/compat)" % self._what)


cm = _Missing("matplotlib.cm")
colors = _Missing("matplotlib.colors")
colorbar = _Missing("matplotlib.colorbar")
