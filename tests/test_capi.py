"""CPU: the C-ABI library loads, exports every symbol include/hermes_b200.h declares, plans without a GPU
and fails loudly (no CPU fallback) when asked to compute without one."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from hermespy_b200 import _lib
from hermespy_b200.kernels import FadingBatch, _problem, fading_propagate_host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = []
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            text = open(os.path.join(ROOT, "include", fn)).read()
            names += re.findall(r"HB_API\s+[\w\s\*]+?\b(hb_\w+)\s*\(", text)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert "hb_fading_propagate" in names and "hb_fading_propagate_host" in names and len(names) >= 8
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ but not exported"
    assert lib.hb_version() == 100


def _plan(**kw):
    base = dict(batch=2, num_tx=4, num_rx=4, num_samples=15344, max_delay=44, num_taps=3, num_sinusoids=20,
                precision="f32", io128=False, sos_mode="auto", omega_max=100 / 30.72e6,
                tap_delay=np.array([0, 10, 44], np.int32), omega_ptr=None, phi_ptr=None, amp_ptr=None, spatial_ptr=None)
    base.update(kw)
    p, keep = _problem(**base)
    info = _lib.FadingPlanInfo()
    st = _lib.load().hb_fading_plan(C.byref(p), C.byref(info))
    return st, info


def test_planner_modes():
    st, info = _plan()
    assert st == 0 and info.mode == _lib.HB_SOS_POLY and info.poly_order >= 2 and info.num_groups == 3
    assert info.error_bound <= 5e-8
    st, info = _plan(omega_max=0.0)
    assert st == 0 and info.poly_order == 1
    st, info = _plan(omega_max=0.5)
    assert st == 0 and info.mode == _lib.HB_SOS_DIRECT and info.tile == 256
    st, info = _plan(precision="f64")
    assert st == 0 and info.mode == _lib.HB_SOS_DIRECT
    st, info = _plan(num_tx=10)
    assert st == 0 and info.launches == 3  # coefficient kernel + two antenna chunks (8 + 2)


def test_planner_kernel_variants():
    """Which POLY kernel the planner picks (host logic, no GPU): the persistent TMA kernel for complex64 frames of whole
    16-sample rows and >= 2048 outputs with delays below 128 samples, the cp.async window kernel otherwise when its walk
    pays, the gather kernel for the rest; large arrays add the tensor-core GEMM."""
    V = {0: "gather", 1: "window", 2: "tma"}
    st, info = _plan()  # C2 shape, three taps in 0..44
    assert st == 0 and V[info.variant] == "tma" and info.tile == 1024 and info.poly_tile % 1024 == 0
    st, info = _plan(sos_mode="poly_window")
    assert st == 0 and V[info.variant] == "gather"  # 3 taps over 44 samples: a dense window walk would be wasted
    dense = np.arange(0, 45, 3).astype(np.int32)
    st, info = _plan(sos_mode="poly_window", num_taps=dense.size, tap_delay=dense)
    assert st == 0 and V[info.variant] == "window"
    st, info = _plan(sos_mode="poly_gather")
    assert st == 0 and V[info.variant] == "gather"
    st, info = _plan(num_samples=15346)  # not a whole number of 16-sample rows: no tensor map
    assert st == 0 and V[info.variant] != "tma"
    st, info = _plan(num_samples=1600)  # T + D < 2048: too short for the persistent kernel
    assert st == 0 and V[info.variant] != "tma"
    st, info = _plan(io128=True)  # complex128 frames are converted in the load of the window / gather kernels
    assert st == 0 and V[info.variant] != "tma"
    st, info = _plan(omega_max=3e4 / 30.72e6)  # fast fading: Taylor windows shorter than the 1024-output tile
    assert st == 0 and info.poly_tile < 1024 and V[info.variant] != "tma"
    far = np.array([0, 10, 200], np.int32)
    st, info = _plan(max_delay=200, tap_delay=far)  # delays beyond the 8 halo rows
    assert st == 0 and V[info.variant] != "tma"
    st, info = _plan(num_tx=64, num_rx=64)  # C4: K1 + one z-mode launch over 16 antenna chunks + one 64 x 64 GEMM block
    assert st == 0 and V[info.variant] == "tma" and info.launches == 3
    st, info = _plan(num_tx=70, num_rx=66)
    assert st == 0 and info.launches == 2 + 4
    st, info = _plan(num_tx=64, num_rx=8)  # not a large array on both sides: chunked fused kernels (4 antennas each)
    assert st == 0 and info.launches == 1 + 16 if V[info.variant] == "tma" else info.launches == 1 + 8


def test_invalid_problems_are_rejected():
    st, _ = _plan(tap_delay=np.array([0, 50, 44], np.int32))
    assert st == _lib.HB_ERR_INVALID  # delay beyond max_delay / not ascending
    st, _ = _plan(tap_delay=np.array([10, 0, 44], np.int32))
    assert st == _lib.HB_ERR_INVALID
    st, _ = _plan(num_tx=0)
    assert st == _lib.HB_ERR_INVALID
    st, _ = _plan(precision="f64", sos_mode="poly")
    assert st == _lib.HB_ERR_UNSUPPORTED
    assert b"direct" in _lib.load().hb_last_error()
    st, _ = _plan(num_taps=300, tap_delay=np.zeros(300, np.int32))
    assert st == _lib.HB_ERR_UNSUPPORTED


def test_no_cpu_fallback():
    if _lib.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    x = np.zeros((1, 1, 16), np.complex128)
    with pytest.raises(_lib.HermesB200Error) as e:
        fading_propagate_host(x, np.array([0], np.int32), 0, np.zeros((1, 1, 21)), np.zeros((1, 1, 21)),
                              np.ones((1, 1, 2)), np.ones((1, 1, 1), complex))
    assert e.value.status == _lib.HB_ERR_NO_DEVICE
