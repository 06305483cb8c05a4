"""Batched realization + sampling of fading links on the host (vectorized numpy, RNG-parity preserving).

``Channel.realize()`` + ``ChannelRealization.sample()`` cost ~0.3 ms per link in the reference
(fading.py:468-515, mostly ``scipy.stats.norm.cdf`` on five small arrays).  For Monte-Carlo batches the same
draws are made in ONE ``standard_normal((B, S))`` call -- numpy generators fill sequentially, so row b equals
the b-th sequential ``realize()`` of the reference -- and mapped to kernel parameter blocks with array math.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
from scipy.special import ndtr

from .channel.fading.fading import MultipathFadingChannel, _tap_delay_samples
from .core import AntennaMode


def sample_fading_links(channel: MultipathFadingChannel, num_links: int, num_tx: int, num_rx: int, bandwidth: float,
                        rng: Optional[np.random.Generator] = None, reciprocal: bool = False) -> dict:
    """Realize and sample ``num_links`` independent links of ``channel`` (static consistency only).

    Returns the stacked kernel parameter block ``dict(tap_delay, max_delay, omega, phi, amp, spatial, omega_max)``
    accepted by ``kernels.FadingBatch.from_numpy`` / ``kernels.fading_propagate_host``.  Row ``b`` is bit-identical
    to ``channel.realize().sample(tx, rx).kernel_block()`` executed ``b + 1`` times in sequence.
    """
    if channel.correlation_distance != float("inf"):
        raise ValueError("vectorized sampling needs a static realization (correlation_distance = inf)")
    rng = channel._rng if rng is None else rng
    L, N = channel.num_resolvable_paths, channel.num_sinusoids
    S = channel.num_realization_scalars
    dim = int(round((S - 2 * L - 2 * L * N) ** 0.5))
    if num_rx > dim or num_tx > dim:
        raise ValueError(f"channel antenna variable is {dim}x{dim}; construct it with max_antennas >= {max(num_rx, num_tx)}")
    g = rng.standard_normal((num_links, S))
    u = ndtr(g)  # == scipy.stats.norm.cdf (consistent.py:485)
    o = dim * dim
    spatial = np.exp(2j * np.pi * u[:, :o].reshape(num_links, dim, dim))[:, :num_rx, :num_tx]
    corr = channel.antenna_correlation
    mix = {}
    if corr is not None:
        r_rx = corr.sample_covariance(num_rx, AntennaMode.RX)
        r_tx = corr.sample_covariance(num_tx, AntennaMode.TX)
        if dim > 10 and not reciprocal:
            # large-array extension: the Kronecker mix R_rx S R_tx (fading.py:480-489) runs on the device (K2,
            # hb_kron_mix) when the block is uploaded -- FadingBatch.from_numpy / fading_propagate_host take the factors
            mix = dict(r_rx=np.ascontiguousarray(r_rx, dtype=np.complex128), r_tx=np.ascontiguousarray(r_tx, dtype=np.complex128))
        else:
            spatial = r_rx[None] @ spatial @ r_tx[None]
    los_angle = 2 * np.pi * u[:, o : o + L]
    o += L
    nlos_angle = -np.pi + 2 * np.pi * u[:, o : o + L * N].reshape(num_links, L, N)
    o += L * N
    los_phase = -np.pi + 2 * np.pi * u[:, o : o + L]
    o += L
    nlos_phase = -np.pi + 2 * np.pi * u[:, o : o + L * N].reshape(num_links, L, N)

    from .kernels import fading_param_block

    omega, phi, amp = fading_param_block(channel.power_profile, channel.delays, channel.los_gains, channel.nlos_gains,
                                         los_angle, nlos_angle, los_phase, nlos_phase, channel.los_doppler_frequency,
                                         channel.doppler_frequency, channel.gain, bandwidth)
    if reciprocal:
        spatial = np.ascontiguousarray(np.swapaxes(spatial, 1, 2))
    return dict(
        tap_delay=_tap_delay_samples(channel.delays, bandwidth),
        max_delay=int(round(channel.max_delay * bandwidth)),
        omega=omega, phi=phi, amp=amp, spatial=np.ascontiguousarray(spatial),
        omega_max=float(max(abs(channel.los_doppler_frequency), abs(channel.doppler_frequency)) / bandwidth),
        **mix,
    )
