#!/usr/bin/env python
"""Install the UNMODIFIED reference (HermesPy 1.6.0, /root/reference) into the git-ignored ``baseline/_ref/``.

    python tools/install_reference.py [--force]

``pip install --no-index --no-build-isolation --no-deps --target baseline/_ref /root/reference`` with the
reference's own CMake option ``HERMES_BUILD_AFF3CT=OFF``: the optional AFF3CT forward-error-correction bindings
(C++, 5 minutes to build, not on the channel path, and their import crashes under this image's pybind11) are left
out -- the reference treats them as optional (hermespy/fec/__init__.py:21-45).  The install is what travels to the
GPU box: ``tests/test_dropin_gpu.py`` and ``bench.py --impl reference`` import it through ``oracle/refload.py``
(which stubs the reference's absent plotting / storage / cluster dependencies).  Nothing is copied into the
tracked tree.
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TARGET = os.path.join(ROOT, "baseline", "_ref")
SOURCE = os.environ.get("HERMES_REFERENCE_SOURCE", "/root/reference")


def install(force: bool = False) -> str:
    if os.path.isdir(os.path.join(TARGET, "hermespy")) and not force:
        return "present"
    if not os.path.isdir(os.path.join(SOURCE, "hermespy")):
        return "no source tree"
    shutil.rmtree(TARGET, ignore_errors=True)
    os.makedirs(os.path.dirname(TARGET), exist_ok=True)
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
           "--find-links", "/opt/wheelhouse", "--config-settings=cmake.define.HERMES_BUILD_AFF3CT=OFF",
           "--target", TARGET, SOURCE]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    if r.returncode != 0:
        shutil.rmtree(TARGET, ignore_errors=True)
        return "pip failed: " + (r.stderr.strip().splitlines() or ["?"])[-1]
    return "installed"


if __name__ == "__main__":
    print(f"[install_reference] {TARGET}: {install('--force' in sys.argv)}")
