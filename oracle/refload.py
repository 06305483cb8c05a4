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

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the source tree in the build container, else the pip install made from it (`tools/install_reference.py`,
# git-ignored `baseline/_ref/`, the only form of the reference that travels to the GPU box)
_CANDIDATES = [os.environ.get("HERMES_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")]
REFERENCE_ROOT = next((c for c in _CANDIDATES if c and os.path.isdir(os.path.join(c, "hermespy"))), "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "hermespy"))


_loaded = False


def load_reference():
    """Insert stubs + the reference root into ``sys.path`` and import the channel stack."""
    global _loaded
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if not _loaded:
        # matplotlib / h5py inert, sparse and ray functional -- only what is really missing (hermespy_b200/shims)
        from hermespy_b200 import shims

        shims.install()
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
