"""Float64 numpy restatement of HermesPy's multipath-fading hot path (test oracle).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Pinned against the live reference
(``tests/test_oracle_vs_reference.py``) and against committed golden vectors generated from it
(``tests/golden/fading_*.npz``).

Reference functions restated here (paths relative to the reference root):

* tap gains from Rice factors ............ hermespy/channel/fading/fading.py:721-733
* consistent-variable layout ............. hermespy/channel/fading/fading.py:739-754
* normals -> sample parameters ........... hermespy/channel/fading/fading.py:468-515,
                                           hermespy/channel/consistent.py:475-485
* sum-of-sinusoids tap impulse ........... hermespy/channel/fading/fading.py:293-343
* tap-delay-line propagate + spatial mix . hermespy/channel/fading/fading.py:371-406
* channel state information .............. hermespy/channel/fading/fading.py:345-369
* reciprocal sample ...................... hermespy/channel/fading/fading.py:517-538
"""
from __future__ import annotations

from dataclasses import dataclass, replace
from typing import Optional

import numpy as np
from scipy.stats import norm

TWO_PI = 2.0 * np.pi

#: The reference hard-codes a (10, 10) antenna phase variable (fading.py:742).
ANTENNA_VARIABLE_DIM = 10


@dataclass
class FadingParams:
    """All numbers a ``MultipathFadingSample`` holds (fading.py:156-215), as plain arrays."""

    power: np.ndarray  # [L]
    delay: np.ndarray  # [L] seconds
    los_gain: np.ndarray  # [L]
    nlos_gain: np.ndarray  # [L]
    los_angle: np.ndarray  # [L]
    nlos_angle: np.ndarray  # [L, N]
    los_phase: np.ndarray  # [L]
    nlos_phase: np.ndarray  # [L, N]
    los_doppler: float  # used as angular rate, no 2*pi (fading.py:303-304)
    nlos_doppler: float
    spatial: np.ndarray  # [>=Nrx, >=Ntx] complex128
    gain: float
    fs: float  # LinkState.bandwidth == sampling rate
    num_rx: int
    num_tx: int

    @property
    def num_taps(self) -> int:
        return int(self.power.shape[0])

    @property
    def num_sinusoids(self) -> int:
        return int(self.nlos_angle.shape[1])

    def reciprocal(self) -> "FadingParams":
        """fading.py:517-538 -- identical fading, transposed spatial response."""
        return replace(self, spatial=self.spatial.T, num_rx=self.num_tx, num_tx=self.num_rx)


def rice_gains(rice_factors: np.ndarray, num_sinusoids: int):
    """LOS / NLOS amplitude factors per tap (fading.py:721-733)."""
    k = np.asarray(rice_factors, dtype=np.float64)
    los = np.ones_like(k)
    nlos = np.zeros_like(k)
    fin = ~np.isposinf(k)
    los[fin] = np.sqrt(k[fin] / (1.0 + k[fin]))
    nlos[fin] = np.sqrt(1.0 / ((1.0 + k[fin]) * num_sinusoids))
    return los, nlos


def num_scalars(num_taps: int, num_sinusoids: int) -> int:
    """Number of standard normals one realization draws (fading.py:739-754)."""
    a = ANTENNA_VARIABLE_DIM * ANTENNA_VARIABLE_DIM
    return a + 2 * num_taps + 2 * num_taps * num_sinusoids


def params_from_normals(
    g: np.ndarray,
    *,
    delays: np.ndarray,
    powers: np.ndarray,
    rice_factors: np.ndarray,
    num_sinusoids: int,
    doppler: float,
    los_doppler: Optional[float],
    gain: float,
    fs: float,
    num_rx: int,
    num_tx: int,
    num_rx_antennas: Optional[int] = None,
    num_tx_antennas: Optional[int] = None,
    cov_rx: Optional[np.ndarray] = None,
    cov_tx: Optional[np.ndarray] = None,
    antenna_dim: int = ANTENNA_VARIABLE_DIM,
) -> FadingParams:
    """Map a static realization's normals to sample parameters (fading.py:468-515).

    ``delays/powers/rice_factors`` must already be sorted by delay (fading.py:707-711).
    Variable offsets follow the declaration order at fading.py:742-754:
    antenna (dim x dim) | los angles (L) | nlos angles (L, N) | los phases (L) | nlos phases (L, N).
    ``antenna_dim`` > 10 is the documented extension for arrays beyond the reference's cap.
    """
    g = np.asarray(g, dtype=np.float64).ravel()
    L = int(len(delays))
    N = int(num_sinusoids)
    u = norm.cdf(g)  # consistent.py:485
    o = 0
    a = antenna_dim * antenna_dim
    nra = num_rx if num_rx_antennas is None else num_rx_antennas
    nta = num_tx if num_tx_antennas is None else num_tx_antennas
    spatial = np.exp(2j * np.pi * u[o : o + a].reshape(antenna_dim, antenna_dim))[:nra, :nta]
    o += a
    if cov_rx is not None or cov_tx is not None:
        # Note: covariance matrices themselves, not square roots (fading.py:480-489)
        spatial = cov_rx @ spatial @ cov_tx
    los_angle = TWO_PI * u[o : o + L]
    o += L
    nlos_angle = -np.pi + TWO_PI * u[o : o + L * N].reshape(L, N)
    o += L * N
    los_phase = -np.pi + TWO_PI * u[o : o + L]
    o += L
    nlos_phase = -np.pi + TWO_PI * u[o : o + L * N].reshape(L, N)
    los_gain, nlos_gain = rice_gains(rice_factors, N)
    return FadingParams(
        power=np.asarray(powers, dtype=np.float64),
        delay=np.asarray(delays, dtype=np.float64),
        los_gain=los_gain,
        nlos_gain=nlos_gain,
        los_angle=los_angle,
        nlos_angle=nlos_angle,
        los_phase=los_phase,
        nlos_phase=nlos_phase,
        los_doppler=float(doppler if los_doppler is None else los_doppler),
        nlos_doppler=float(doppler),
        spatial=spatial,
        gain=float(gain),
        fs=float(fs),
        num_rx=int(num_rx),
        num_tx=int(num_tx),
    )


def tap_delays_in_samples(p: FadingParams) -> np.ndarray:
    """numpy round-half-even of delay*fs (fading.py:297)."""
    return np.rint(p.delay * p.fs).astype(np.int64)


def max_delay_in_samples(p: FadingParams) -> int:
    """Python ``round`` of max(delay)*fs (fading.py:372)."""
    return int(round(float(p.delay.max()) * p.fs))


def sinusoid_rates(p: FadingParams):
    """Per-sample angular increments and start phases of every sinusoid.

    Returns ``(omega[L, N+1], phi[L, N+1], amp[L, N+1])`` with column 0 the LOS term, such that
    ``h_l[n] = sum_k amp[l,k] * exp(1j*(omega[l,k]*n + phi[l,k]))`` (fading.py:326-342).
    This is the flat parameter block the CUDA kernels consume; it is *not* how the oracle
    evaluates the impulse (see :func:`tap_impulses`, which follows the reference's operation order).
    """
    L, N = p.num_taps, p.num_sinusoids
    n = 1 + np.arange(N)
    omega = np.empty((L, N + 1))
    phi = np.empty((L, N + 1))
    amp = np.empty((L, N + 1))
    scale = np.sqrt(p.gain * p.power)
    omega[:, 0] = p.los_doppler * np.cos(p.los_angle) / p.fs
    omega[:, 1:] = p.nlos_doppler * np.cos((TWO_PI * n[None, :] + p.nlos_angle) / N) / p.fs
    phi[:, 0] = p.los_phase
    phi[:, 1:] = p.nlos_phase
    amp[:, 0] = p.los_gain * scale
    amp[:, 1:] = (p.nlos_gain * scale)[:, None]
    return omega, phi, amp


def tap_impulses(p: FadingParams, num_samples: int) -> np.ndarray:
    """``h[L, T]``: Rician sum-of-sinusoids coefficient of every tap (fading.py:293-343).

    Keeps the reference's operation order for the phase argument, ``(doppler * (n / fs)) * cos(.)``,
    so that agreement with the reference is at the 1e-15 level even for large Doppler rates.
    """
    L, N = p.num_taps, p.num_sinusoids
    T = int(num_samples)
    t = np.arange(T) / p.fs
    nlos_time = p.nlos_doppler * t
    los_time = p.los_doppler * t
    k = 1 + np.arange(N)
    h = np.empty((L, T), dtype=np.complex128)
    for l in range(L):
        c = np.cos((TWO_PI * k + p.nlos_angle[l]) / N)  # [N]
        arg = nlos_time[:, None] * c[None, :] + p.nlos_phase[l][None, :]
        acc = p.nlos_gain[l] * np.exp(1j * arg).sum(axis=1)
        acc = acc + p.los_gain[l] * np.exp(1j * (los_time * np.cos(p.los_angle[l]) + p.los_phase[l]))
        h[l] = acc * (p.gain * p.power[l]) ** 0.5
    return h


def propagate(p: FadingParams, x: np.ndarray) -> np.ndarray:
    """``y[Nrx, T+D] = S @ sum_l shift_{d_l}(x * h_l)`` (fading.py:371-406).

    ``x`` is complex128 ``[Ntx, T]``.  Guards follow fading.py:381,394-397.
    """
    x = np.asarray(x, dtype=np.complex128)
    T = x.shape[1]
    D = max_delay_in_samples(p)
    S = p.spatial[: p.num_rx, : p.num_tx]
    if T + D <= 0 or S.shape[0] == 0:
        return np.zeros((S.shape[0], T + D), dtype=np.complex128)
    d = tap_delays_in_samples(p)
    h = tap_impulses(p, T)
    z = np.zeros((S.shape[1], T + D), dtype=np.complex128)
    for l in range(p.num_taps):
        z[:, d[l] : d[l] + T] += x * h[l][None, :]
    return S @ z


def state(p: FadingParams, num_samples: int, max_num_taps: int) -> np.ndarray:
    """Dense CSI ``[Nrx?, Ntx?, T, taps]`` as the reference builds it (fading.py:345-369).

    Note the reference uses the *unsliced* spatial response in the outer product and skips only
    taps with ``d_l > num_taps`` (sic, ``>`` not ``>=``; a tap with d_l == num_taps would raise
    in the reference, which cannot happen because num_taps >= 1 + round(max_delay*fs) unless
    ``max_num_taps`` truncates).
    """
    taps = min(1 + max_delay_in_samples(p), int(max_num_taps))
    d = tap_delays_in_samples(p)
    h = tap_impulses(p, num_samples)
    siso = np.zeros((num_samples, taps), dtype=np.complex128)
    for l in range(p.num_taps):
        if d[l] > taps:
            continue
        siso[:, d[l]] += h[l]
    return np.einsum("ij,kl->ijkl", p.spatial, siso)


def expected_energy_scale(p: FadingParams) -> float:
    """fading.py:289-291."""
    return float(p.gain * np.sum(p.power))


def merged_delay_groups(p: FadingParams):
    """Distinct integer delays and the tap -> group map (host-side helper mirrored by the product).

    Taps are sorted by delay in the reference constructor (fading.py:707-711), so equal integer
    delays are contiguous after rounding as long as rint is monotone (it is).
    """
    d = tap_delays_in_samples(p)
    uniq, inv = np.unique(d, return_inverse=True)
    return uniq, inv


# ---- fractional delays: windowed-sinc extension (NOT reference behaviour: the reference rounds) -------------
# ``InterpolationMode.SINC`` is defined by the reference as the Whittaker-Shannon formula
# s^(tau) = sum_m s_m sinc(tau fs - m)  (hermespy/core/definitions.py:82-93) but no channel implements it.  The extension
# delays every tap's faded signal z_l[n] = x[n] h_l[n] by its TRUE delay tau_l fs = floor + eps with a Kaiser-windowed sinc
# of 2 W taps centred on the fractional position:
#     y[:, m] = S sum_l sum_{j = floor_l - W + 1}^{floor_l + W} g(j - tau_l fs) z_l[m - j],
#     g(u) = sinc(u) I0(beta sqrt(1 - (u / W)^2)) / I0(beta)   for |u| < W,   0 otherwise.
# Filter taps at negative delays j < 0 (the non-causal precursor of channel taps less than W - 1 samples into the frame)
# are dropped -- the channel stays causal; the output has T + D_s samples with D_s = max_l floor_l + W.  An integer
# delay (eps = 0) degenerates to the single tap g(0) = 1: NEAREST and SINC agree wherever the reference's rounding is exact.

SINC_HALF_WIDTH = 6
SINC_KAISER_BETA = 6.0


def sinc_kernel(u: np.ndarray, half_width: int = SINC_HALF_WIDTH, beta: float = SINC_KAISER_BETA) -> np.ndarray:
    u = np.asarray(u, dtype=np.float64)
    inside = np.abs(u) < half_width
    w = np.zeros_like(u)
    w[inside] = np.i0(beta * np.sqrt(1.0 - (u[inside] / half_width) ** 2)) / np.i0(beta)
    s = np.sinc(u)
    s[(u == np.floor(u)) & (u != 0)] = 0.0  # exact zeros at the integers (np.sinc leaves 1e-17 residues)
    return s * w


def sinc_max_delay_in_samples(p: FadingParams, half_width: int = SINC_HALF_WIDTH) -> int:
    return int(np.floor(p.delay * p.fs).max()) + int(half_width)


def propagate_sinc(p: FadingParams, x: np.ndarray, half_width: int = SINC_HALF_WIDTH,
                   beta: float = SINC_KAISER_BETA) -> np.ndarray:
    """Fractional-delay propagation by direct convolution of every tap's faded signal with its own windowed-sinc FIR
    (independent of the tap expansion the product uses)."""
    x = np.asarray(x, dtype=np.complex128)
    T = x.shape[1]
    W = int(half_width)
    Ds = sinc_max_delay_in_samples(p, W)
    S = p.spatial[: p.num_rx, : p.num_tx]
    h = tap_impulses(p, T)
    z = np.zeros((S.shape[1], T + Ds), dtype=np.complex128)
    for l in range(p.num_taps):
        tau = p.delay[l] * p.fs
        fl = int(np.floor(tau))
        j = np.arange(max(0, fl - W + 1), fl + W + 1)  # causal taps only
        g = sinc_kernel(j - tau, W, beta)
        zl = x * h[l][None, :]
        for a in range(S.shape[1]):
            full = np.convolve(zl[a], g)  # full[i] = sum_k zl[i - k] g[k]  -> lands on output index i + j[0]
            z[a, j[0]: j[0] + len(full)] += full
    return S @ z
