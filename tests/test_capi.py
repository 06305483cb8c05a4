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
    assert lib.hb_version() == 110


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
    st, info = _plan(precision="f64")  # parity mode: the float64 Taylor path, truncation bound below 1e-14
    assert st == 0 and info.mode == _lib.HB_SOS_POLY and info.variant == 0 and info.poly_order in (4, 6, 8)
    assert info.error_bound <= 1e-14 and info.launches == 2
    st, info = _plan(precision="f64", sos_mode="direct")
    assert st == 0 and info.mode == _lib.HB_SOS_DIRECT
    st, info = _plan(precision="f64", omega_max=0.1)  # 1500 rad over the frame: the reference's argument rounding matters
    assert st == 0 and info.mode == _lib.HB_SOS_DIRECT
    st, info = _plan(precision="f64", omega_max=0.1, sos_mode="poly")  # ... and no (order, window) reaches 1e-14 either
    assert st == _lib.HB_ERR_UNSUPPORTED
    st, info = _plan(num_tx=10)
    assert st == 0 and info.launches == 3  # coefficient kernel + two antenna chunks (8 + 2)


def test_planner_falls_back_before_it_refuses():
    """AUTO / POLY requests whose preferred kernel overflows shared memory (thousands of receive antennas, very
    long delay spreads) go to the antenna-chunking gather kernel, then to per-sample evaluation -- staged or, for delay
    spreads no tile can hold, reading x from global memory -- before the problem is refused: the drop-in has no CPU path."""
    wide = dict(num_tx=8, num_rx=4096, num_taps=64, max_delay=200, num_samples=4096,
                tap_delay=np.linspace(0, 200, 64).astype(np.int32))
    st, info = _plan(**wide)
    assert st == 0 and info.mode == _lib.HB_SOS_POLY and info.variant == 0  # HB_VARIANT_GATHER
    st, info = _plan(sos_mode="poly_window", **wide)  # an explicit request is not second-guessed
    assert st == _lib.HB_ERR_UNSUPPORTED and b"shared memory" in _lib.load().hb_last_error()
    long_spread = dict(num_samples=100000, max_delay=60000, num_taps=23, tap_delay=np.linspace(0, 60000, 23).astype(np.int32))
    for precision in ("f32", "f64"):
        st, info = _plan(precision=precision, **long_spread)
        assert st == 0 and info.mode == _lib.HB_SOS_DIRECT and info.tile == 256, (precision, st)
    st, info = _plan(precision="f64", num_samples=20000, max_delay=16000, num_taps=20,
                     tap_delay=np.linspace(0, 16000, 20).astype(np.int32))
    assert st == 0


def test_planner_kernel_variants():
    """Which POLY kernel the planner picks (host logic, no GPU): the persistent TMA kernel for complex64 frames of whole
    16-sample rows and >= 2048 outputs with delays below 128 samples, the cp.async window kernel otherwise when its walk
    pays, the gather kernel for the rest; large arrays add the tensor-core GEMM."""
    V = {0: "gather", 1: "window", 2: "tma", 3: "fused", 4: "siso"}
    st, info = _plan(num_tx=1, num_rx=1, sos_mode="poly_siso")  # one antenna per side: the time-packed kernel on request
    assert st == 0 and V[info.variant] == "siso" and info.tile == 2048 and info.launches == 2
    st, info = _plan(num_tx=1, num_rx=1, num_samples=500, io128=True, sos_mode="poly_siso")
    assert st == 0 and V[info.variant] == "siso" and info.tile == 1024  # 544 outputs: one CTA
    dense1 = np.arange(0, 45, 3).astype(np.int32)
    st, info = _plan(num_tx=1, num_rx=1, num_taps=dense1.size, tap_delay=dense1)  # 1 x 1 AUTO: the window kernel, not TMA
    assert st == 0 and V[info.variant] == "window"
    st, info = _plan(num_tx=1, num_rx=1, num_taps=dense1.size, tap_delay=dense1, sos_mode="poly_tma")
    assert st == 0 and V[info.variant] == "tma"
    st, info = _plan(num_tx=2, num_rx=1, sos_mode="poly_siso")
    assert st == _lib.HB_ERR_UNSUPPORTED
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
    st, info = _plan(num_tx=64, num_rx=64, sos_mode="poly_fused")  # K1 + the single GEMM / delay-line kernel (opt-in)
    assert st == 0 and V[info.variant] == "fused" and info.launches == 2 and info.tile == 64
    st, info = _plan(num_tx=4, num_rx=4, sos_mode="poly_fused")  # not a large array: the request degrades to the default
    assert st == 0 and V[info.variant] == "tma"
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
    st, _ = _plan(precision="f64", sos_mode="poly_gather")  # the complex64 kernels have no float64 form
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


def test_planner_fuzz_never_crashes_and_respects_its_own_limits():
    """Host logic: random shapes / delay tables / Doppler through hb_fading_plan -- every answer is a status code, and an
    accepted plan is self-consistent (tile sizes, Taylor window multiples, launch counts, error bound)."""
    rng = np.random.default_rng(2026)
    V = {0: "gather", 1: "window", 2: "tma", 3: "fused", 4: "siso"}
    accepted = {"gather": 0, "window": 0, "tma": 0, "fused": 0, "siso": 0, "direct": 0}
    for _ in range(400):
        ntx, nrx = int(rng.choice([1, 2, 3, 4, 5, 8, 10, 16, 33, 64, 70])), int(rng.choice([1, 2, 4, 7, 8, 16, 40, 64, 66]))
        T = int(rng.choice([0, 1, 37, 500, 1024, 2048, 4100, 15344, 16384, 1 << 20]))
        L = int(rng.integers(1, 40))
        dmax = int(rng.choice([0, 1, 7, 44, 66, 115, 127, 128, 300, 1023, 1500]))
        delays = np.sort(rng.integers(0, dmax + 1, L)).astype(np.int32)
        delays[0] = 0 if rng.random() < 0.7 else delays[0]
        delays[-1] = dmax
        delays = np.sort(delays).astype(np.int32)
        omega = float(rng.choice([0.0, 1e-7, 3.3e-6, 1e-4, 1e-2, 0.5]))
        sos_mode = str(rng.choice(["auto", "auto", "poly", "direct", "poly_window", "poly_gather", "poly_tma", "poly_fused",
                                   "poly_fused", "poly_siso"]))
        if sos_mode == "poly_siso" and rng.random() < 0.7:
            ntx = nrx = 1  # the only shape that mode takes
        st, info = _plan(batch=int(rng.integers(0, 5000)), num_tx=ntx, num_rx=nrx, num_samples=T, max_delay=dmax,
                         num_taps=L, num_sinusoids=int(rng.choice([0, 1, 8, 20])), omega_max=omega, tap_delay=delays,
                         io128=bool(rng.random() < 0.3), precision=str(rng.choice(["f32", "f32", "f64"])),
                         sos_mode=sos_mode)
        assert st in (0, _lib.HB_ERR_INVALID, _lib.HB_ERR_UNSUPPORTED), st
        if st != 0:
            assert _lib.load().hb_last_error()  # a refused problem always says why
            continue
        assert info.num_groups == np.unique(delays).size and info.launches >= 1 and info.num_tiles >= 1
        if info.mode == _lib.HB_SOS_POLY:
            v = V[info.variant]
            accepted[v] += 1
            assert info.error_bound <= 2e-7 and info.poly_order in (1, 2, 3, 4, 6, 8)
            assert info.poly_tile % info.tile == 0 or v == "gather"
            if v == "tma":
                assert info.tile == 1024 and T % 16 == 0 and T + dmax >= 2048 and dmax <= 127
            if v == "siso":
                assert ntx == 1 and nrx == 1 and info.tile in (512, 1024, 2048) and info.poly_tile >= 512 and info.launches == 2
            if v == "fused":
                assert info.tile == 64 and 16 <= ntx <= 64 and 16 <= nrx <= 64 and dmax <= 128 and info.launches == 2
        else:
            accepted["direct"] += 1
            assert info.tile == 256
    assert all(n > 0 for n in accepted.values()), accepted  # the fuzz reaches every kernel family


def _cdl_plan(ntx, nrx, T, delays, speed=10.0, precision=0, variant=0, fs=30.72e6, fc=3.5e9):
    """hb_cdl_plan on a bare problem description (planning reads no device arrays)."""
    lib = _lib.load()
    p = _lib.CdlProblem()
    td = np.asarray(delays, dtype=np.int32)
    p.batch, p.num_tx, p.num_rx, p.num_samples = 4, ntx, nrx, T
    p.max_delay, p.num_terms = int(td.max()) + 1, td.size
    p.precision, p.variant = precision, variant
    p.carrier_frequency, p.sampling_rate, p.max_speed = fc, fs, speed
    p.term_delay = td.ctypes.data_as(C.POINTER(C.c_int32))
    info = _lib.FadingPlanInfo()
    rc = lib.hb_cdl_plan(C.byref(p), C.byref(info))
    return rc, info.as_dict()


def test_cdl_planner_kernel_variants():
    """Which K6 kernel the CDL planner picks (include/hermes_b200.h: hb_cdl_variant); 0 AUTO, 1 GATHER, 2 UMMA."""
    delays = np.repeat(np.arange(0, 80, 4), 20)  # 20 clusters x 20 rays -> 20 delay groups
    rc, d = _cdl_plan(32, 4, 2048, delays)  # config C3 shape: tensor cores, 4 M-tiles per Taylor window
    assert rc == 0 and d["mode"] == 1 and d["variant"] == 2 and d["tile"] == 512 and d["num_groups"] == 20
    assert 1 <= d["poly_order"] <= 4 and d["error_bound"] <= 5e-8
    rc, d = _cdl_plan(2, 2, 2048, delays)  # short K: the FP32-pipe kernel
    assert rc == 0 and d["mode"] == 1 and d["variant"] == 1
    rc, d = _cdl_plan(2, 2, 2048, delays, variant=2)  # ... unless asked for
    assert rc == 0 and d["variant"] == 2
    rc, d = _cdl_plan(32, 4, 2048, delays, variant=1)
    assert rc == 0 and d["variant"] == 1
    rc, d = _cdl_plan(32, 4, 2048, delays, precision=1)  # f64 parity mode: per-ray kernel
    assert rc == 0 and d["mode"] == 2 and d["variant"] == 0
    rc, d = _cdl_plan(32, 4, 2048, delays, speed=0.0)  # static link: one Taylor term
    assert rc == 0 and d["variant"] == 2 and d["poly_order"] == 1
    rc, d = _cdl_plan(32, 8, 2048, delays, speed=30.0)  # wider accumulator rows -> fewer M-tiles per window
    assert rc == 0 and d["variant"] == 2 and d["tile"] in (256, 512)
    # beyond the staging limits AUTO falls back to the gather kernel, an explicit request is refused
    long_delays = np.repeat(np.arange(0, 1000, 50), 20)
    rc, d = _cdl_plan(32, 4, 2048, long_delays)
    assert rc == 0 and d["variant"] == 1
    rc, _ = _cdl_plan(32, 4, 2048, long_delays, variant=2)
    assert rc == _lib.HB_ERR_UNSUPPORTED and "tensor-core" in _lib.load().hb_last_error().decode()
    rc, _ = _cdl_plan(32, 4, 2048, delays, variant=7)
    assert rc == _lib.HB_ERR_INVALID
    # fast links: four Taylor terms do not cover a 512-sample window any more -> the FP32-pipe kernel with shorter tiles,
    # and only beyond those the per-ray path
    rc, d = _cdl_plan(32, 4, 2048, delays, speed=60.0)
    assert rc == 0 and d["mode"] == 1 and d["variant"] == 1 and d["tile"] == 256 and d["error_bound"] <= 5e-8
    rc, d = _cdl_plan(32, 4, 2048, delays, speed=110.0)
    assert rc == 0 and d["mode"] == 1 and d["variant"] == 1 and d["tile"] == 128
    rc, d = _cdl_plan(32, 4, 2048, delays, speed=400.0)
    assert rc == 0 and d["mode"] == 2
