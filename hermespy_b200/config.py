"""Process-wide compute settings of the GPU channel kernels."""
from __future__ import annotations

from contextlib import contextmanager

#: "f32": complex64 arithmetic (rel. L2 <= 1e-5 vs the float64 reference, throughput mode).
#: "f64": float64 parity mode (direct evaluation; bit-exact bit-error counts against the reference).
precision = "f32"
#: "auto" | "poly" | "direct" -- sum-of-sinusoids evaluation path (see include/hermes_b200.h).
sos_mode = "auto"
#: CUDA device index used by the host-facing plugin API.
device = 0
#: extension (NOT reference behaviour): serve ``InterpolationMode.SINC`` requests to the fading channel with windowed-sinc
#: fractional delays instead of the reference's rounding (fading.py:297 ignores the interpolation argument)
sinc_extension = False
#: batched drop runner (hermespy_b200/runner.py): drops in flight per Simulation actor (0 = the reference's serial loop)
batch_drops = 0
#: forked helper processes that run the lanes' CPU stages (0 = lanes run in the actor's own process)
workers = 0


@contextmanager
def compute(precision_: str = None, sos_mode_: str = None):
    """Temporarily override ``precision`` / ``sos_mode``."""
    global precision, sos_mode
    old = (precision, sos_mode)
    if precision_ is not None:
        if precision_ not in ("f32", "f64"):
            raise ValueError("precision must be 'f32' or 'f64'")
        precision = precision_
    if sos_mode_ is not None:
        if sos_mode_ not in ("auto", "poly", "direct"):
            raise ValueError("sos_mode must be 'auto', 'poly' or 'direct'")
        sos_mode = sos_mode_
    try:
        yield
    finally:
        precision, sos_mode = old
