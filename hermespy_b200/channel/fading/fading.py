"""Stochastic multipath fading channel behind the reference's plugin API, computed on the B200.

Host-side mirror of hermespy/channel/fading/fading.py:
``MultipathFadingChannel`` (:585-972) -> ``MultipathFadingRealization`` (:425-538) ->
``MultipathFadingSample`` (:147-406).  Realization and sampling stay on the host with numpy because the
draw order of the shared generator is part of parity (consistent.py:431-432); ``_propagate`` and ``state``
-- 95 % of the reference's run time (fading.py:293-343, 371-406) -- go through the C-ABI
(``hb_fading_propagate_host`` / ``hb_fading_state``).  ``propagate_batch`` is the batched entry the drop
runner uses: many samples, one launch per shared delay profile.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Set

import numpy as np

from ... import config
from ...core import AntennaMode, InterpolationMode, SignalBlock
from ..channel import Channel, ChannelRealization, ChannelSample, ChannelSampleHook, LinkState
from ..consistent import ConsistentGenerator, ConsistentRealization, ConsistentUniform
from .correlation import AntennaCorrelation

pi = np.pi


class FadingChannelState(object):
    """Impulse-response channel state of a fading sample in factored form.

    The reference materializes ``einsum('ij,kl->ijkl', S, siso)`` with ``siso[T, taps]`` scattered from
    the tap gains (fading.py:351-364).  Here the factors are kept -- ``spatial[Nrx, Ntx]``,
    ``gains[G, T]`` (GPU result) and ``delays[G]`` -- and the dense ``[Nrx, Ntx, T, taps]`` array is only
    built on request.
    """

    def __init__(self, spatial: np.ndarray, gains: np.ndarray, delays: np.ndarray, num_taps: int) -> None:
        self.spatial = spatial
        self.gains = gains
        self.delays = delays
        self.num_delay_taps = int(num_taps)

    @property
    def num_samples(self) -> int:
        return int(self.gains.shape[1])

    @property
    def num_receive_streams(self) -> int:
        return int(self.spatial.shape[0])

    @property
    def num_transmit_streams(self) -> int:
        return int(self.spatial.shape[1])

    def siso_state(self) -> np.ndarray:
        """``[T, taps]`` tap gains placed at their delay columns."""
        siso = np.zeros((self.num_samples, self.num_delay_taps), dtype=np.complex128)
        for g, d in enumerate(self.delays):
            siso[:, d] = self.gains[g]
        return siso

    def dense_state(self) -> np.ndarray:
        """Layout of hermespy/core/channel.py ``ChannelStateInformation.dense_state()``: [Nrx, Ntx, T, taps]."""
        return self.spatial[:, :, None, None] * self.siso_state()[None, None, :, :]


def _tap_delay_samples(delay_profile: np.ndarray, bandwidth: float) -> np.ndarray:
    return np.rint(np.asarray(delay_profile) * bandwidth).astype(np.int32)  # fading.py:297


class MultipathFadingSample(ChannelSample):
    """Immutable sample of a multipath fading channel (fading.py:147-406); same constructor signature."""

    def __init__(self, power_profile, delay_profile, los_angles, nlos_angles, los_phases, nlos_phases, los_gains,
                 nlos_gains, los_doppler: float, nlos_doppler: float, spatial_response: np.ndarray, gain: float,
                 state: LinkState) -> None:
        ChannelSample.__init__(self, state)
        self.__power_profile = power_profile
        self.__delay_profile = delay_profile
        self.__los_angles = los_angles
        self.__nlos_angles = nlos_angles
        self.__los_phases = los_phases
        self.__nlos_phases = nlos_phases
        self.__los_gains = los_gains
        self.__nlos_gains = nlos_gains
        self.__los_doppler = los_doppler
        self.__nlos_doppler = nlos_doppler
        self.__spatial_response = spatial_response
        self.__gain = gain
        self.__max_delay = self.__delay_profile.max()
        self.__block = None

    power_profile = property(lambda self: self.__power_profile)
    delay_profile = property(lambda self: self.__delay_profile)
    los_angles = property(lambda self: self.__los_angles)
    nlos_angles = property(lambda self: self.__nlos_angles)
    los_phases = property(lambda self: self.__los_phases)
    nlos_phases = property(lambda self: self.__nlos_phases)
    los_gains = property(lambda self: self.__los_gains)
    nlos_gains = property(lambda self: self.__nlos_gains)
    los_doppler = property(lambda self: self.__los_doppler)
    nlos_doppler = property(lambda self: self.__nlos_doppler)
    spatial_response = property(lambda self: self.__spatial_response)
    gain = property(lambda self: self.__gain)
    max_delay = property(lambda self: self.__max_delay)

    @property
    def expected_energy_scale(self) -> float:
        return self.gain * np.sum(self.power_profile)  # fading.py:289-291

    @property
    def max_delay_in_samples(self) -> int:
        return round(self.__max_delay * self.bandwidth)  # Python round, fading.py:372

    # ---- kernel parameter block --------------------------------------------------------------------
    def kernel_block(self) -> dict:
        """Flat parameter block of this sample for the CUDA kernels (cached; the sample is immutable)."""
        if self.__block is None:
            from ...kernels import fading_param_block

            omega, phi, amp = fading_param_block(
                self.__power_profile, self.__delay_profile, self.__los_gains, self.__nlos_gains, self.__los_angles,
                self.__nlos_angles, self.__los_phases, self.__nlos_phases, self.__los_doppler, self.__nlos_doppler,
                self.__gain, self.bandwidth)
            spatial = np.ascontiguousarray(
                self.__spatial_response[: self.num_receive_antennas, : self.num_transmit_antennas], dtype=np.complex128)
            self.__block = dict(
                tap_delay=_tap_delay_samples(self.__delay_profile, self.bandwidth),
                max_delay=int(self.max_delay_in_samples),
                omega=omega, phi=phi, amp=amp, spatial=spatial,
                omega_max=float(max(abs(self.__los_doppler), abs(self.__nlos_doppler)) / self.bandwidth),
            )
        return self.__block

    # ---- plugin interface ---------------------------------------------------------------------------
    def _propagate(self, signal: SignalBlock, interpolation: InterpolationMode) -> SignalBlock:
        """GPU replacement of fading.py:371-406 for one block (complex128 in, complex128 out)."""
        from ...kernels import fading_propagate_host

        num_samples = signal.shape[1]
        D = self.max_delay_in_samples
        nrx = min(self.num_receive_antennas, self.__spatial_response.shape[0])
        if num_samples + D <= 0 or nrx == 0:
            out = np.zeros((nrx, num_samples + D), dtype=np.complex128)
        else:
            b = self.kernel_block()
            x = np.ascontiguousarray(np.asarray(signal, dtype=np.complex128))[None]
            out = fading_propagate_host(
                x, b["tap_delay"], b["max_delay"], b["omega"][None], b["phi"][None], b["amp"][None], b["spatial"][None],
                omega_max=b["omega_max"], precision=config.precision, sos_mode=config.sos_mode)[0]
        return SignalBlock(out.shape[0], out.shape[1], getattr(signal, "offset", 0), out.tobytes())

    def state(self, num_samples: int, max_num_taps: int,
              interpolation_mode: InterpolationMode = InterpolationMode.NEAREST) -> FadingChannelState:
        """GPU replacement of fading.py:345-369 (tap gains on the device, factored CSI on the host)."""
        from ...kernels import FadingBatch, fading_state

        b = self.kernel_block()
        num_taps = min(1 + self.max_delay_in_samples, max_num_taps)
        keep = b["tap_delay"] <= num_taps  # the reference skips taps with d_l > num_taps (fading.py:355)
        if np.any(b["tap_delay"][keep] >= num_taps):
            raise IndexError("tap delay equals the number of CSI taps (the reference raises here as well)")
        fb = FadingBatch.from_numpy(b["tap_delay"][keep], b["max_delay"], b["omega"][None][:, keep],
                                    b["phi"][None][:, keep], b["amp"][None][:, keep], b["spatial"][None],
                                    omega_max=b["omega_max"], device=f"cuda:{config.device}")
        h, gd = fading_state(fb, num_samples, precision=config.precision, io128=True)
        return FadingChannelState(self.__spatial_response, h[0].cpu().numpy(), gd, num_taps)


def propagate_batch(samples: Sequence[MultipathFadingSample], signals: Sequence[np.ndarray],
                    precision: Optional[str] = None, sos_mode: Optional[str] = None) -> List[np.ndarray]:
    """Propagate many (sample, signal) pairs with one launch per shared (delay profile, shape) group.

    This is the batched form of ``[s._propagate(x) for s, x in zip(samples, signals)]`` used by the drop
    runner; results are returned in input order as ``[Nrx, T + D]`` arrays of the signals' dtype.
    """
    from ...kernels import fading_propagate_host

    precision = config.precision if precision is None else precision
    sos_mode = config.sos_mode if sos_mode is None else sos_mode
    groups = {}
    for i, (s, x) in enumerate(zip(samples, signals)):
        b = s.kernel_block()
        x = np.asarray(x)
        key = (b["tap_delay"].tobytes(), b["max_delay"], b["omega"].shape, b["spatial"].shape, x.shape, x.dtype.str)
        groups.setdefault(key, []).append(i)
    out: List[Optional[np.ndarray]] = [None] * len(samples)
    for idx in groups.values():
        blocks = [samples[i].kernel_block() for i in idx]
        b0 = blocks[0]
        x = np.stack([np.asarray(signals[i]) for i in idx])
        if x.dtype not in (np.complex64, np.complex128):
            x = x.astype(np.complex128)
        if x.shape[1] != b0["spatial"].shape[1]:
            raise ValueError(
                "Number of signal streams to be propagated does not match the number of transmitter antennas "
                f"({x.shape[1]} != {b0['spatial'].shape[1]}))")
        y = fading_propagate_host(
            x, b0["tap_delay"], b0["max_delay"], np.stack([b["omega"] for b in blocks]),
            np.stack([b["phi"] for b in blocks]), np.stack([b["amp"] for b in blocks]),
            np.stack([b["spatial"] for b in blocks]), omega_max=max(b["omega_max"] for b in blocks),
            precision=precision, sos_mode=sos_mode)
        for k, i in enumerate(idx):
            out[i] = y[k]
    return out  # type: ignore[return-value]


class MultipathFadingRealization(ChannelRealization[MultipathFadingSample]):
    """Realization of a multipath fading channel (fading.py:425-538)."""

    def __init__(self, random_realization: ConsistentRealization, antenna_correlation_variable: ConsistentUniform,
                 los_angles_variable, nlos_angles_variable: ConsistentUniform, los_phases_variable: ConsistentUniform,
                 nlos_phases_variable: ConsistentUniform, power_profile, delay_profile, los_gains, nlos_gains,
                 los_doppler: float, nlos_doppler: float, antenna_correlation: Optional[AntennaCorrelation],
                 sample_hooks: Set[ChannelSampleHook], gain: float) -> None:
        ChannelRealization.__init__(self, sample_hooks, gain)
        self.random_realization = random_realization
        self.__antenna_variable = antenna_correlation_variable
        self.__los_angles_variable = los_angles_variable
        self.__nlos_angles_variable = nlos_angles_variable
        self.__los_phases_variable = los_phases_variable
        self.__nlos_phases_variable = nlos_phases_variable
        self.__power_profile = power_profile
        self.__delay_profile = delay_profile
        self.__los_gains = los_gains
        self.__nlos_gains = nlos_gains
        self.__los_doppler = los_doppler
        self.__nlos_doppler = nlos_doppler
        self.__antenna_correlation = antenna_correlation

    def _sample(self, state: LinkState) -> MultipathFadingSample:
        cs = self.random_realization.sample(state.transmitter.pose.translation, state.receiver.pose.translation)
        nrx_a = state.receiver.antennas.num_antennas
        ntx_a = state.transmitter.antennas.num_antennas
        dim = self.__antenna_variable.shape[0]
        if nrx_a > dim or ntx_a > dim:
            raise ValueError(
                f"link uses {nrx_a}x{ntx_a} antennas but the channel's antenna phase variable is {dim}x{dim}; "
                "construct the channel with max_antennas >= the largest array (extension of the reference's 10x10 cap)")
        spatial = np.exp(2j * np.pi * self.__antenna_variable.sample(cs))[:nrx_a, :ntx_a]
        if self.__antenna_correlation is not None:
            spatial = (self.__antenna_correlation.sample_covariance(state.receiver.antennas, AntennaMode.RX)
                       @ spatial
                       @ self.__antenna_correlation.sample_covariance(state.transmitter.antennas, AntennaMode.TX))
        if isinstance(self.__los_angles_variable, float):
            los_angles = self.__los_angles_variable * np.ones_like(self.__power_profile)
        else:
            los_angles = 2 * np.pi * self.__los_angles_variable.sample(cs)
        nlos_angles = -np.pi + 2 * np.pi * self.__nlos_angles_variable.sample(cs)
        los_phases = -np.pi + 2 * np.pi * self.__los_phases_variable.sample(cs)
        nlos_phases = -np.pi + 2 * np.pi * self.__nlos_phases_variable.sample(cs)
        return MultipathFadingSample(self.__power_profile, self.__delay_profile, los_angles, nlos_angles, los_phases,
                                     nlos_phases, self.__los_gains, self.__nlos_gains, self.__los_doppler,
                                     self.__nlos_doppler, spatial, self.gain, state)

    def _reciprocal_sample(self, sample: MultipathFadingSample, state: LinkState) -> MultipathFadingSample:
        return MultipathFadingSample(sample.power_profile, sample.delay_profile, sample.los_angles, sample.nlos_angles,
                                     sample.los_phases, sample.nlos_phases, sample.los_gains, sample.nlos_gains,
                                     sample.los_doppler, sample.nlos_doppler, sample.spatial_response.T, sample.gain,
                                     state)


class MultipathFadingChannel(Channel[MultipathFadingRealization, MultipathFadingSample]):
    """Base class of the stochastic multipath fading channels (fading.py:585-972).

    Constructor arguments, validation messages, tap sorting, Rice-factor gains and -- for generator
    parity -- the order of random draws and consistent-variable declarations follow the reference
    (fading.py:669-754).  ``max_antennas`` is an extension: the reference hard-codes a 10x10 antenna phase
    variable (fading.py:742); larger arrays need a larger variable, which changes the number of normals per
    realization, so the default keeps 10.
    """

    _DEFAULT_DECORRELATION_DISTANCE = float("inf")
    _DEFAULT_NUM_SINUSOIDS = 20
    _DEFAULT_DOPPLER_FREQUENCY = 0.0

    def __init__(self, delays, power_profile, rice_factors,
                 correlation_distance: float = _DEFAULT_DECORRELATION_DISTANCE,
                 num_sinusoids: int = _DEFAULT_NUM_SINUSOIDS, los_angle: Optional[float] = None,
                 doppler_frequency: float = _DEFAULT_DOPPLER_FREQUENCY, los_doppler_frequency: Optional[float] = None,
                 antenna_correlation: Optional[AntennaCorrelation] = None, gain: float = Channel._DEFAULT_GAIN,
                 seed: Optional[int] = None, max_antennas: int = 10) -> None:
        d = np.array(delays) if isinstance(delays, list) else delays
        p = np.array(power_profile) if isinstance(power_profile, list) else power_profile
        k = np.array(rice_factors) if isinstance(rice_factors, list) else rice_factors
        if d.ndim != 1 or p.ndim != 1 or k.ndim != 1:
            raise ValueError("Delays, power profile and rice factors must be vectors")
        if len(delays) < 1:
            raise ValueError("Configuration must contain at least one delay tap")
        if len(delays) != len(power_profile) or len(power_profile) != len(rice_factors):
            raise ValueError("Delays, power profile and rice factor vectors must be of equal length")
        if np.any(d < 0.0):
            raise ValueError("Delays must be greater or equal to zero")
        if np.any(p < 0.0):
            raise ValueError("Power profile factors must be greater or equal to zero")
        if np.any(k < 0.0):
            raise ValueError("Rice factors must be greater or equal to zero")

        self.__antenna_correlation = None
        Channel.__init__(self, gain, seed)

        order = np.argsort(delays)  # same call on the same input as fading.py:707 (tie order matters)
        self.__delays = d[order]
        self.__power_profile = p[order]
        self.__rice_factors = k[order]
        self.__num_sinusoids = num_sinusoids
        # one uniform is drawn when no angle is given (fading.py:713); the value itself is never used by the
        # math because a per-tap consistent variable replaces it below
        self.los_angle = self._rng.uniform(-pi, pi) if los_angle is None else los_angle
        self.doppler_frequency = doppler_frequency
        self.__los_doppler_frequency = los_doppler_frequency
        self.__max_delay = max(self.__delays)

        inf = np.isposinf(self.__rice_factors)
        fin = ~inf
        self.__los_gains = np.empty(len(self.__delays), dtype=float)
        self.__nlos_gains = np.empty(len(self.__delays), dtype=float)
        self.__los_gains[inf] = 1.0
        self.__los_gains[fin] = np.sqrt(self.__rice_factors[fin] / (1 + self.__rice_factors[fin]))
        self.__nlos_gains[fin] = np.sqrt(1 / ((1 + self.__rice_factors[fin]) * self.__num_sinusoids))
        self.__nlos_gains[inf] = 0.0

        self.antenna_correlation = antenna_correlation
        self.correlation_distance = correlation_distance

        L, N = len(self.__delays), self.__num_sinusoids
        self.__generator = ConsistentGenerator(self)
        self.__antenna_variable = self.__generator.uniform((int(max_antennas), int(max_antennas)))
        self.__los_angles_variable = self.__generator.uniform((L,)) if self.los_angle is not None else self.los_angle
        self.__nlos_angles_variable = self.__generator.uniform((L, N))
        self.__los_phases_variable = self.__generator.uniform((L,))
        self.__nlos_phases_variable = self.__generator.uniform((L, N))

    # ---- properties -----------------------------------------------------------------------------------
    @property
    def correlation_distance(self) -> float:
        return self.__correlation_distance

    @correlation_distance.setter
    def correlation_distance(self, distance: float) -> None:
        if distance < 0:
            raise ValueError("Correlation distance must be greater or equal to zero")
        self.__correlation_distance = distance

    delays = property(lambda self: self.__delays)
    power_profile = property(lambda self: self.__power_profile)
    rice_factors = property(lambda self: self.__rice_factors)
    los_gains = property(lambda self: self.__los_gains)
    nlos_gains = property(lambda self: self.__nlos_gains)
    max_delay = property(lambda self: self.__max_delay)
    num_resolvable_paths = property(lambda self: len(self.__delays))
    num_realization_scalars = property(lambda self: self.__generator.num_scalars)

    @property
    def num_sinusoids(self) -> int:
        return self.__num_sinusoids

    @num_sinusoids.setter
    def num_sinusoids(self, num: int) -> None:
        if num < 0:
            raise ValueError("Number of sinusoids must be greater or equal to zero")
        self.__num_sinusoids = num

    @property
    def los_doppler_frequency(self) -> float:
        return self.doppler_frequency if self.__los_doppler_frequency is None else self.__los_doppler_frequency

    @los_doppler_frequency.setter
    def los_doppler_frequency(self, frequency: Optional[float]) -> None:
        self.__los_doppler_frequency = frequency

    @property
    def antenna_correlation(self) -> Optional[AntennaCorrelation]:
        return self.__antenna_correlation

    @antenna_correlation.setter
    def antenna_correlation(self, value: Optional[AntennaCorrelation]) -> None:
        if value is not None:
            value.channel = self
        self.__antenna_correlation = value

    def _realize(self) -> MultipathFadingRealization:
        return MultipathFadingRealization(
            self.__generator.realize(self.correlation_distance), self.__antenna_variable, self.__los_angles_variable,
            self.__nlos_angles_variable, self.__los_phases_variable, self.__nlos_phases_variable,
            self.__power_profile, self.__delays, self.__los_gains, self.__nlos_gains, self.los_doppler_frequency,
            self.doppler_frequency, self.antenna_correlation, self.sample_hooks, self.gain)
