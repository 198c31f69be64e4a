"""Load a reference program of dsarvan/simulation by path -- TEST INFRASTRUCTURE ONLY.

The reference scripts import matplotlib (absent in this image) and call ``plt.style.use`` at
import time, so permissive stub modules are installed first.  Only usable where
``/root/reference`` exists (the build container); the GPU box uses the committed goldens instead.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("FDTD_REFERENCE_ROOT", "/root/reference")


class _Stub(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        m = _Stub(self.__name__ + "." + k)
        setattr(self, k, m)
        return m

    def __call__(self, *a, **k):
        return self

    def __iter__(self):           # ``fig, ax = plt.subplots(...)``
        return iter((self, self))


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "fd2d", "program"))


def load(relpath: str):
    """Import e.g. ``fd2d/program/fd2d_3_3.py`` and return the module (main() is not run)."""
    for n in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation", "mpl_toolkits",
              "mpl_toolkits.mplot3d", "mpl_toolkits.mplot3d.axes3d"):
        sys.modules.setdefault(n, _Stub(n))
    path = os.path.join(REFERENCE_ROOT, relpath)
    name = "ref_" + relpath.replace("/", "_").replace(".py", "")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run_main(relpath: str, capture=("visualize", "surfaceplot", "contourplot", "amplitude",
                                    "amplitudeplot")):
    """Run the program's own ``main()`` with its plot helpers replaced by recorders.
    Returns {helper_name: positional args tuple} -- the final arrays the program would plot."""
    mod = load(relpath)
    seen = {}
    for fn in capture:
        if hasattr(mod, fn):
            setattr(mod, fn, (lambda name: (lambda *a, **k: seen.__setitem__(name, a)))(fn))
    mod.main()
    return seen
