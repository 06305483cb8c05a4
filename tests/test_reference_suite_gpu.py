"""The reference's OWN channel unit tests, executed against the patched classes on the GPU (last cell).

``tools/fetch_reference_tests.py`` copies ``tests/unit_tests/channel/test_fading.py`` and ``test_cdl.py`` (verbatim,
git-ignored) next to the reference install; here they run under ``hermespy_b200.dropin.enable()`` in both precisions:
every ``propagate`` / ``state`` call those tests make -- propagate == CSI-propagate (test_fading.py:167-176,
test_cdl.py:110-119), delay shift (:415-441), gain (:560-591), identity correlation (:593-614), seed (:385-396), the
1000-realization energy tests (:712-854, test_cdl.py:317-335), time of flight (:337-364) -- is served by the CUDA kernels.

Left out, by name: the serialization and plotting tests, which exercise h5py / matplotlib (absent from this image and
replaced by inert stubs) and never reach the channel path.  They fail identically without the drop-in.
"""
import json
import os
import re
import sys
import unittest

import pytest

from oracle.refload import load_reference, reference_available

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SUITE = os.path.join(ROOT, "baseline", "_ref_tests")
_have_suite = os.path.isfile(os.path.join(SUITE, "unit_tests", "channel", "test_fading.py"))

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (reference_available() and _have_suite),
                                 reason="needs baseline/_ref and baseline/_ref_tests (tools/fetch_reference_tests.py)")]

NEEDS_ABSENT_PACKAGES = re.compile(r"serializ|plot|visualization", re.I)  # h5py / matplotlib are inert stubs here

# complex64 arithmetic cannot meet an ELEMENT-WISE 6-decimal comparison where the reference test drives the Doppler to
# 50 x the sampling rate (test_fading.py:146-147: up to 50 rad of phase advance per sample, 60 sinusoid terms, amplitudes
# of several units): the f32 mode's contract is relative L2 <= 1e-5 (north_star), which the same configuration meets in
# tests/test_dropin_gpu.py (golden case "extreme_doppler_2x2").  The float64 mode passes the test as written.
#
# test_channel_gain (test_fading.py:560-591) compares two propagations of UNSEEDED Gaussian samples, one scaled by
# sqrt(10), element-wise to 6 decimals (1.5e-6 absolute): outputs reach |y| ~ 10, where one complex64 rounding is 6e-7, so the
# outcome depends on the draw (it passed in two of three GPU runs of this round).  Relative L2 of the pair is ~1e-7.
#
# test_cdl.py:110-119 (propagate vs dense CSI of a 2 x 2 CDL link at fs = 1 MHz, outputs of magnitude ~10, 6 decimals): the link
# moves at 5 m/s, which at that sampling rate is "fast" (1.1e-4 rad per sample).  Until the planner gave fast links shorter Taylor
# tiles it fell through to the per-ray FP64 kernel even in the f32 mode and passed by being exact; now it runs the FP32 kernel
# (relative L2 ~1e-7 against the oracle, tests/test_cdl_gpu.py::test_fast_links_stay_on_the_taylor_path) and misses the absolute
# 1.5e-6 of the element-wise check.
F32_PRECISION_EXCEPTIONS = {"unit_tests.channel.test_fading.TestMultipathFadingSample.test_propagate_state",
                            "unit_tests.channel.test_fading.TestMultipathFadingChannel.test_channel_gain",
                            "unit_tests.channel.test_cdl.TestClusterDelayLineSample.test_propagate_state"}


def _flatten(suite):
    for t in suite:
        if isinstance(t, unittest.TestSuite):
            yield from _flatten(t)
        else:
            yield t


def _load():
    load_reference()
    if SUITE not in sys.path:
        sys.path.insert(0, SUITE)
    loaded = unittest.defaultTestLoader.loadTestsFromNames(["unit_tests.channel.test_fading", "unit_tests.channel.test_cdl"])
    tests = [t for t in _flatten(loaded)]
    broken = [t for t in tests if type(t).__name__ == "_FailedTest"]
    assert not broken, f"reference test modules failed to import: {broken}"
    keep = [t for t in tests if not NEEDS_ABSENT_PACKAGES.search(t.id())]
    return unittest.TestSuite(keep), len(tests) - len(keep)


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_reference_channel_unit_tests_pass_on_the_cuda_path(precision):
    import hermespy_b200.dropin as dropin
    from hermespy_b200 import _lib

    suite, left_out = _load()
    before = _lib.launch_counts()
    fallbacks_before = sum(dropin.fallbacks.values())
    dropin.enable(precision=precision)
    try:
        result = unittest.TextTestRunner(stream=open(os.devnull, "w"), verbosity=0).run(suite)
    finally:
        dropin.disable()
    after = _lib.launch_counts()
    launched = {k: after[k] - before[k] for k in after if after[k] != before[k]}
    record = {"precision": precision, "run": result.testsRun, "failures": len(result.failures), "errors": len(result.errors),
              "skipped_by_reference": len(result.skipped), "left_out_h5py_matplotlib": left_out,
              "kernel_launches": launched, "reference_fallbacks": sum(dropin.fallbacks.values()) - fallbacks_before,
              "failed": [t.id() for t, _ in result.failures + result.errors]}
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, f"reference_suite_{precision}.json"), "w") as fh:
        json.dump(record, fh, indent=1)
    print(json.dumps(record))
    for t, tb in result.failures + result.errors:
        print(t.id(), tb[-800:])
    assert result.testsRun >= 100
    assert sum(launched.values()) >= 1000, launched  # the energy tests alone propagate thousands of realizations
    assert sum(dropin.fallbacks.values()) == fallbacks_before  # nothing was served by the reference's numpy code
    allowed = F32_PRECISION_EXCEPTIONS if precision == "f32" else set()
    assert not result.errors and set(record["failed"]) <= allowed, record["failed"]
