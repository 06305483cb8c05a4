"""Stand-ins for HermesPy's import-time dependencies that are absent from GPU images.

HermesPy imports ``matplotlib``, ``h5py``, ``ray`` and ``sparse`` at module import time.  On a box without them,

    import hermespy_b200.shims as shims
    shims.install()          # before `import hermespy`

registers, ONLY for the packages that are really missing:

* ``matplotlib`` (+ the submodules HermesPy names) and ``mpl_toolkits``, ``h5py``: inert modules whose attributes are
  ``MagicMock`` objects -- plotting and HDF5 storage are outside the channel hot path;
* ``sparse``: a small functional stand-in (dense-backed ``SparseArray`` / ``COO`` / ``GCXS``) because the fading
  ``state()`` call site (hermespy/channel/fading/fading.py:362) and ``ChannelStateInformation``
  (hermespy/core/channel.py:144) need real types;
* ``ray``: the functional in-process stand-in of ``hermespy_b200.shims.ray`` so that ``Simulation.run()`` works.
"""
from __future__ import annotations

import importlib.util
import sys
import types
from unittest.mock import MagicMock

import numpy as np

_MPL_SUBMODULES = ("pyplot axes figure lines axis ticker colors tri collections container image projections "
                   "projections.polar animation patches text gridspec transforms cm backend_bases widgets").split()


class _Inert(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        m = MagicMock(name=f"{self.__name__}.{name}")
        setattr(self, name, m)
        return m


def _missing(name: str) -> bool:
    if name in sys.modules:
        return False
    try:
        return importlib.util.find_spec(name) is None
    except (ImportError, ValueError):
        return True


def make_sparse_module() -> types.ModuleType:
    mod = types.ModuleType("sparse")
    mod.__path__ = []  # type: ignore[attr-defined]

    class SparseArray(object):
        """Dense-backed stand-in for pydata/sparse arrays."""

        def __init__(self, data):
            self._d = np.asarray(data)

        @classmethod
        def from_numpy(cls, x, *a, **k):
            return cls(x)

        def todense(self):
            return self._d

        @property
        def shape(self):
            return self._d.shape

        @property
        def ndim(self):
            return self._d.ndim

        @property
        def dtype(self):
            return self._d.dtype

        def __getitem__(self, item):
            return type(self)(self._d[item])

        def copy(self):
            return type(self)(self._d.copy())

        def __array__(self, dtype=None, copy=None):
            return self._d if dtype is None else self._d.astype(dtype)

    class COO(SparseArray):
        pass

    class GCXS(SparseArray):
        pass

    mod.SparseArray = SparseArray
    mod.COO = COO
    mod.GCXS = GCXS
    mod.tensordot = lambda a, b, *args, **kw: np.tensordot(np.asarray(a), np.asarray(b), *args, **kw)
    mod.einsum = lambda s, *ops: np.einsum(s, *[np.asarray(o) for o in ops])
    return mod


def install() -> list:
    """Register stand-ins for the missing packages; returns the names that were registered."""
    done = []
    if _missing("matplotlib"):
        for n in ["matplotlib"] + ["matplotlib." + x for x in _MPL_SUBMODULES]:
            m = _Inert(n)
            m.__path__ = []  # type: ignore[attr-defined]
            sys.modules[n] = m
        done.append("matplotlib")
    if _missing("mpl_toolkits") or "matplotlib" in done:
        for n in ("mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.mplot3d.art3d", "mpl_toolkits.mplot3d.axes3d"):
            if n not in sys.modules:
                m = _Inert(n)
                m.__path__ = []  # type: ignore[attr-defined]
                sys.modules[n] = m
        done.append("mpl_toolkits")
    if _missing("h5py"):
        m = _Inert("h5py")
        m.__path__ = []  # type: ignore[attr-defined]
        sys.modules["h5py"] = m
        done.append("h5py")
    if _missing("sparse"):
        sys.modules["sparse"] = make_sparse_module()
        done.append("sparse")
    from . import ray as _ray

    if _ray.install():
        done.append("ray")
    return done
