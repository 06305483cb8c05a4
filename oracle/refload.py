"""Loader for the UNMODIFIED HermesPy reference (test infrastructure only).

Usable where ``/root/reference`` exists (the build container) or where its pip install
``baseline/_ref/`` travelled with the snapshot (the GPU box; made by ``tools/install_reference.py``).
It is used (a) by ``oracle/make_golden.py`` to generate the committed golden vectors under
``tests/golden/``, (b) by CPU tests that pin the numpy restatement in ``oracle/`` against the live
reference code, (c) by ``tests/test_dropin_gpu.py`` -- the unmodified reference drop loop with the CUDA
path patched in -- and (d) by ``bench.py --impl reference``.  Everything that uses it skips / falls back
to the numpy port when no reference is present.

The reference imports ``matplotlib``, ``h5py``, ``ray`` and ``sparse`` at module import time
and none of them is installed here.  ``matplotlib`` / ``h5py`` are replaced by inert stub modules, ``ray`` by the
functional in-process stand-in ``hermespy_b200.shims.ray`` (so ``Simulation.run()`` works); ``sparse`` gets
a small functional stand-in (dense-backed ``GCXS``/``COO``) because the fading ``state()``
call-site (hermespy/channel/fading/fading.py:362) and ``ChannelStateInformation``
(hermespy/core/channel.py:144) need real types.
"""
from __future__ import annotations

import os
import sys
import types
from unittest.mock import MagicMock

import numpy as np

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the source tree in the build container, else the pip install made from it (`tools/install_reference.py`,
# git-ignored `baseline/_ref/`, the only form of the reference that travels to the GPU box)
_CANDIDATES = [os.environ.get("HERMES_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")]
REFERENCE_ROOT = next((c for c in _CANDIDATES if c and os.path.isdir(os.path.join(c, "hermespy"))), "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "hermespy"))


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        m = MagicMock(name=f"{self.__name__}.{name}")
        setattr(self, name, m)
        return m


def _make_sparse_module() -> types.ModuleType:
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


_loaded = False


def load_reference():
    """Insert stubs + the reference root into ``sys.path`` and import the channel stack."""
    global _loaded
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if not _loaded:
        mpl = ["matplotlib"] + [
            "matplotlib." + x
            for x in (
                "pyplot axes figure lines axis ticker colors tri collections container image "
                "projections projections.polar animation patches text gridspec transforms cm "
                "backend_bases widgets"
            ).split()
        ]
        names = mpl + [
            "mpl_toolkits",
            "mpl_toolkits.mplot3d",
            "mpl_toolkits.mplot3d.art3d",
            "mpl_toolkits.mplot3d.axes3d",
            "h5py",
        ]
        for n in names:
            if n not in sys.modules:
                m = _Stub(n)
                m.__path__ = []  # type: ignore[attr-defined]
                sys.modules[n] = m
        # `ray`: a FUNCTIONAL in-process stand-in (hermespy_b200/shims/ray.py) so that Simulation.run() itself works
        from hermespy_b200.shims import ray as _ray_shim

        _ray_shim.install()
        if "sparse" not in sys.modules:
            sys.modules["sparse"] = _make_sparse_module()
        if REFERENCE_ROOT not in sys.path:
            # appended, not prepended: the reference tree has its own top-level `tests` package, which must not shadow
            # this repo's (spawned test workers re-import `tests.*` from the inherited path)
            sys.path.append(REFERENCE_ROOT)
        _loaded = True
    import hermespy.channel  # noqa: F401
    import hermespy.simulation  # noqa: F401
    import hermespy.modem  # noqa: F401
    import hermespy

    return hermespy
