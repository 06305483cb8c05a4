#!/usr/bin/env python
"""Fetch the reference's OWN channel unit tests into the git-ignored ``baseline/_ref_tests/`` (checker infrastructure).

    python tools/fetch_reference_tests.py [--force]

SURVEY 8(c): "the new repo must run these reference tests against the patched classes".  The files
(``tests/unit_tests/channel/test_fading.py``, ``test_cdl.py`` and the two helper modules they import) are copied
verbatim from the read-only reference tree where it is mounted (the build container); like the pip install under
``baseline/_ref/`` the copy is git-ignored -- no reference source enters the history -- but travels to the GPU box
with the snapshot, where ``tests/test_reference_suite_gpu.py`` runs them with ``hermespy_b200.dropin`` enabled.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TARGET = os.path.join(ROOT, "baseline", "_ref_tests")
SOURCE = os.path.join(os.environ.get("HERMES_REFERENCE_SOURCE", "/root/reference"), "tests", "unit_tests")
FILES = ("utils.py", "channel/test_fading.py", "channel/test_cdl.py", "core/test_factory.py")


def fetch(force: bool = False) -> str:
    if all(os.path.isfile(os.path.join(TARGET, "unit_tests", f)) for f in FILES) and not force:
        return "present"
    if not os.path.isdir(SOURCE):
        return "no source tree"
    shutil.rmtree(TARGET, ignore_errors=True)
    for f in FILES:
        dst = os.path.join(TARGET, "unit_tests", f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SOURCE, f), dst)
    for d in ("", "channel", "core"):
        open(os.path.join(TARGET, "unit_tests", d, "__init__.py"), "a").close()
    return "fetched"


if __name__ == "__main__":
    print(f"[fetch_reference_tests] {TARGET}: {fetch('--force' in sys.argv)}")
