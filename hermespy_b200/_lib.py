"""ctypes binding of ``libhermes_b200.so`` (the C-ABI declared in ``include/hermes_b200.h``).

There is no CPU fallback: if the shared library has not been built (``python -m hermespy_b200.build``
or ``__graft_entry__.build()``) importing the compute API raises, and every compute call fails with
``HermesB200Error`` when no CUDA device is visible.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

HB_OK = 0
HB_ERR_INVALID = 1
HB_ERR_CUDA = 2
HB_ERR_UNSUPPORTED = 3
HB_ERR_NO_DEVICE = 4

HB_F32 = 0
HB_F64 = 1

HB_SOS_AUTO = 0
HB_SOS_POLY = 1
HB_SOS_DIRECT = 2
HB_SOS_POLY_GATHER = 3
HB_SOS_POLY_WINDOW = 4
HB_SOS_POLY_TMA = 5
HB_SOS_POLY_FUSED = 6

HB_VARIANT_GATHER = 0
HB_VARIANT_WINDOW = 1
HB_VARIANT_TMA = 2
HB_VARIANT_FUSED = 3

HB_MAX_TAPS = 256
HB_CDL_MAX_TERMS = 1024
HB_CDL_MAX_GROUPS = 64
HB_ELEMENT_STRIDE = 12
HB_ELEMENTS_IDEAL, HB_ELEMENTS_UNIFORM, HB_ELEMENTS_PER_ELEMENT = 0, 1, 2
HB_ELEMENT_IDEAL, HB_ELEMENT_LINEAR, HB_ELEMENT_PATCH, HB_ELEMENT_DIPOLE = 0, 1, 2, 3


class HermesB200Error(RuntimeError):
    """Raised for every non-zero status of the C-ABI (carries ``status``)."""

    def __init__(self, status: int, message: str) -> None:
        super().__init__(f"libhermes_b200 status {status}: {message}")
        self.status = status


class FadingProblem(C.Structure):
    """Mirror of ``hb_fading_problem``."""

    _fields_ = [
        ("batch", C.c_int32),
        ("num_tx", C.c_int32),
        ("num_rx", C.c_int32),
        ("num_samples", C.c_int32),
        ("max_delay", C.c_int32),
        ("num_taps", C.c_int32),
        ("num_sinusoids", C.c_int32),
        ("precision", C.c_int32),
        ("io_complex128", C.c_int32),
        ("sos_mode", C.c_int32),
        ("omega_max", C.c_double),
        ("tap_delay", C.POINTER(C.c_int32)),
        ("omega", C.c_void_p),
        ("phi", C.c_void_p),
        ("amp", C.c_void_p),
        ("spatial", C.c_void_p),
    ]


class FadingPlanInfo(C.Structure):
    """Mirror of ``hb_fading_plan_info``."""

    _fields_ = [
        ("mode", C.c_int32),
        ("tile", C.c_int32),
        ("poly_order", C.c_int32),
        ("num_groups", C.c_int32),
        ("num_tiles", C.c_int32),
        ("launches", C.c_int32),
        ("error_bound", C.c_double),
        ("variant", C.c_int32),
        ("poly_tile", C.c_int32),
    ]

    def as_dict(self) -> dict:
        return {k: getattr(self, k) for k, _ in self._fields_}


NUM_KERNEL_KINDS = 9
KERNEL_KINDS = ("sos_poly_coef", "tdl_poly", "tdl_direct", "sos_state", "cdl_rays", "cdl_propagate", "spatial_gemm",
                "stats", "misc")


class ReceiveInput(C.Structure):
    """Mirror of ``hb_receive_input``."""

    _fields_ = [("samples", C.c_void_p), ("num_samples", C.c_int32), ("offset", C.c_int32)]


class ProfileReport(C.Structure):
    """Mirror of ``hb_profile_report``."""

    _fields_ = [("ms", C.c_double * NUM_KERNEL_KINDS), ("launches", C.c_int64 * NUM_KERNEL_KINDS)]


_LIB = None


def library_path() -> Path:
    override = os.environ.get("HERMES_B200_LIB")
    if override:
        return Path(override)
    return Path(__file__).resolve().parent / "lib" / "libhermes_b200.so"


def _declare(lib: C.CDLL) -> None:
    lib.hb_version.restype = C.c_int
    lib.hb_version.argtypes = []
    lib.hb_last_error.restype = C.c_char_p
    lib.hb_last_error.argtypes = []
    lib.hb_device_count.restype = C.c_int
    lib.hb_device_count.argtypes = []
    lib.hb_set_device.restype = C.c_int
    lib.hb_set_device.argtypes = [C.c_int]
    lib.hb_fading_plan.restype = C.c_int
    lib.hb_fading_plan.argtypes = [C.POINTER(FadingProblem), C.POINTER(FadingPlanInfo)]
    lib.hb_fading_sinc_taps.restype = C.c_int
    lib.hb_fading_sinc_taps.argtypes = [C.POINTER(C.c_double), C.c_int32, C.c_int32, C.c_double, C.c_int32,
                                        C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_int32),
                                        C.POINTER(C.c_int32)]
    lib.hb_fading_propagate.restype = C.c_int
    lib.hb_fading_propagate.argtypes = [
        C.POINTER(FadingProblem),
        C.c_void_p,
        C.c_void_p,
        C.c_void_p,
        C.POINTER(FadingPlanInfo),
    ]
    lib.hb_fading_propagate_host.restype = C.c_int
    lib.hb_fading_propagate_host.argtypes = [
        C.POINTER(FadingProblem),
        C.c_void_p,
        C.c_void_p,
        C.c_int32,
        C.POINTER(FadingPlanInfo),
    ]
    lib.hb_fading_state.restype = C.c_int
    lib.hb_fading_state.argtypes = [C.POINTER(FadingProblem), C.c_void_p, C.POINTER(C.c_int32), C.c_void_p]
    lib.hb_bit_errors.restype = C.c_int
    lib.hb_bit_errors.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p]
    lib.hb_stats_accumulate.restype = C.c_int
    lib.hb_stats_accumulate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                        C.c_void_p, C.c_void_p]
    lib.hb_fading_sample.restype = C.c_int
    lib.hb_fading_sample.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                     C.c_void_p, C.c_double, C.c_double, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p]
    lib.hb_kron_mix.restype = C.c_int
    lib.hb_kron_mix.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                C.c_void_p]
    lib.hb_spatial_gemm_3xtf32.restype = C.c_int
    lib.hb_spatial_gemm_3xtf32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                           C.c_void_p]
    lib.hb_receive_combine.restype = C.c_int
    lib.hb_receive_combine.argtypes = [C.POINTER(ReceiveInput), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    lib.hb_launch_counts.restype = None
    lib.hb_launch_counts.argtypes = [C.POINTER(C.c_int64)]
    lib.hb_profile_begin.restype = C.c_int
    lib.hb_profile_begin.argtypes = []
    lib.hb_profile_end.restype = C.c_int
    lib.hb_profile_end.argtypes = [C.POINTER(ProfileReport)]
    lib.hb_reserve_sms.restype = C.c_int
    lib.hb_reserve_sms.argtypes = [C.c_int]
    lib.hb_release.restype = None
    lib.hb_release.argtypes = []


def load() -> C.CDLL:
    """Load the shared library once; raise loudly when it is missing."""
    global _LIB
    if _LIB is None:
        path = library_path()
        if not path.exists():
            raise HermesB200Error(
                HB_ERR_NO_DEVICE,
                f"{path} not found -- build it with `python -m hermespy_b200.build` "
                "(there is no CPU fallback for the channel kernels)",
            )
        lib = C.CDLL(str(path))
        _declare(lib)
        _declare_cdl(lib)
        _LIB = lib
    return _LIB


def check(status: int) -> None:
    if status != HB_OK:
        msg = load().hb_last_error()
        raise HermesB200Error(status, msg.decode("utf-8", "replace") if msg else "unknown error")


def device_count() -> int:
    return int(load().hb_device_count())


def set_device(device) -> None:
    """Make ``device`` (index, or None = leave as is) the CUDA device of the calling thread (``hb_set_device``)."""
    if device is not None:
        check(load().hb_set_device(int(device)))


def reserve_sms(num_sms: int) -> int:
    """Keep `num_sms` SMs out of the persistent kernels' grids (room for a concurrent NCCL collective)."""
    return int(load().hb_reserve_sms(int(num_sms)))


def launch_counts() -> dict:
    """Kernel launches per kind since the library was loaded."""
    arr = (C.c_int64 * NUM_KERNEL_KINDS)()
    load().hb_launch_counts(arr)
    return {k: int(arr[i]) for i, k in enumerate(KERNEL_KINDS)}


def profile_begin() -> None:
    check(load().hb_profile_begin())


def profile_end() -> dict:
    """Per-kind summed device time (ms) and launch counts since ``profile_begin``."""
    rep = ProfileReport()
    check(load().hb_profile_end(C.byref(rep)))
    return {k: {"ms": float(rep.ms[i]), "launches": int(rep.launches[i])} for i, k in enumerate(KERNEL_KINDS)}


class CdlProblem(C.Structure):
    """Mirror of ``hb_cdl_problem``."""

    _fields_ = [
        ("batch", C.c_int32),
        ("num_tx", C.c_int32),
        ("num_rx", C.c_int32),
        ("num_samples", C.c_int32),
        ("max_delay", C.c_int32),
        ("num_terms", C.c_int32),
        ("line_of_sight", C.c_int32),
        ("los_delay", C.c_int32),
        ("precision", C.c_int32),
        ("io_complex128", C.c_int32),
        ("carrier_frequency", C.c_double),
        ("sampling_rate", C.c_double),
        ("los_amplitude", C.c_double),
        ("max_speed", C.c_double),
        ("term_delay", C.POINTER(C.c_int32)),
        ("angles", C.c_void_p),
        ("jones", C.c_void_p),
        ("amplitude", C.c_void_p),
        ("tx_pose", C.c_void_p),
        ("rx_pose", C.c_void_p),
        ("rel_velocity", C.c_void_p),
        ("tx_topology", C.c_void_p),
        ("rx_topology", C.c_void_p),
        ("element_mode", C.c_int32),
        ("variant", C.c_int32),
        ("tx_elements", C.c_void_p),
        ("rx_elements", C.c_void_p),
        ("link_term_delay", C.POINTER(C.c_int32)),
        ("link_los_delay", C.POINTER(C.c_int32)),
        ("link_los_amplitude", C.POINTER(C.c_double)),
    ]


def _declare_cdl(lib: C.CDLL) -> None:
    lib.hb_cdl_plan.restype = C.c_int
    lib.hb_cdl_plan.argtypes = [C.POINTER(CdlProblem), C.POINTER(FadingPlanInfo)]
    lib.hb_cdl_propagate.restype = C.c_int
    lib.hb_cdl_propagate.argtypes = [C.POINTER(CdlProblem), C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(FadingPlanInfo)]
    lib.hb_cdl_propagate_host.restype = C.c_int
    lib.hb_cdl_propagate_host.argtypes = [C.POINTER(CdlProblem), C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(FadingPlanInfo)]
    lib.hb_cdl_state.restype = C.c_int
    lib.hb_cdl_state.argtypes = [C.POINTER(CdlProblem), C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_void_p]
