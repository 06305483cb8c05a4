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


def sample_fading_links_device(channel: MultipathFadingChannel, num_links: int, num_tx: int, num_rx: int, bandwidth: float,
                               device="cuda", rng: Optional[np.random.Generator] = None, reciprocal: bool = False):
    """``sample_fading_links`` with everything but the random draw on the device: the host draws
    ``standard_normal((B, S))`` from the channel's generator (row b = the b-th sequential ``realize()`` of the reference),
    ships it (8 S bytes per link), and ``hb_fading_sample`` maps it to the kernel parameter block (Phi, angles, Doppler
    rates, antenna phases); antenna correlation follows on the device too (K2 ``hb_kron_mix``).  Returns a
    ``kernels.FadingBatch``.  Parameters agree with the host path to a few ulp (device ``erfc`` / ``cos`` vs scipy / numpy)."""
    import ctypes as C

    import torch

    from . import _lib
    from .kernels import FadingBatch
    from .montecarlo import kron_mix

    if channel.correlation_distance != float("inf"):
        raise ValueError("vectorized sampling needs a static realization (correlation_distance = inf)")
    rng = channel._rng if rng is None else rng
    L, N = channel.num_resolvable_paths, channel.num_sinusoids
    S = channel.num_realization_scalars
    dim = int(round((S - 2 * L - 2 * L * N) ** 0.5))
    if num_rx > dim or num_tx > dim:
        raise ValueError(f"channel antenna variable is {dim}x{dim}; construct it with max_antennas >= {max(num_rx, num_tx)}")
    dev = torch.device(device)
    g = torch.from_numpy(rng.standard_normal((num_links, S))).to(dev)
    scale = np.sqrt(channel.gain * np.asarray(channel.power_profile, dtype=np.float64))
    amp_tab = torch.from_numpy(np.ascontiguousarray(
        np.stack([np.asarray(channel.los_gains) * scale, np.asarray(channel.nlos_gains) * scale], axis=1))).to(dev)
    omega = torch.empty((num_links, L, N + 1), dtype=torch.float64, device=dev)
    phi = torch.empty_like(omega)
    amp = torch.empty((num_links, L, 2), dtype=torch.float64, device=dev)
    # reciprocal direction: the transposed (mixed) forward matrix (fading.py:517-538) -> [B, Ntx, Nrx]
    rows, cols = (num_tx, num_rx) if reciprocal else (num_rx, num_tx)
    spatial = torch.empty((num_links, rows, cols), dtype=torch.complex128, device=dev)
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.load().hb_fading_sample(
            g.data_ptr(), num_links, S, dim, rows, cols, L, N, amp_tab.data_ptr(),
            float(channel.los_doppler_frequency) / bandwidth, float(channel.doppler_frequency) / bandwidth,
            1 if reciprocal else 0, omega.data_ptr(), phi.data_ptr(), amp.data_ptr(), spatial.data_ptr(), C.c_void_p(st)))
    corr = channel.antenna_correlation
    if corr is not None:
        up = lambda m: torch.from_numpy(np.ascontiguousarray(m, dtype=np.complex128)).to(dev)
        r_rx, r_tx = corr.sample_covariance(num_rx, AntennaMode.RX), corr.sample_covariance(num_tx, AntennaMode.TX)
        if not reciprocal:  # R_rx S R_tx (fading.py:480-489)
            spatial = kron_mix(spatial, up(r_rx), up(r_tx), out=spatial)
        else:  # (R_rx S R_tx)^T = R_tx^T S^T R_rx^T
            spatial = kron_mix(spatial, up(r_tx.T), up(r_rx.T), out=spatial)
    return FadingBatch(tap_delay=_tap_delay_samples(channel.delays, bandwidth),
                       max_delay=int(round(channel.max_delay * bandwidth)), omega=omega, phi=phi, amp=amp, spatial=spatial,
                       omega_max=float(max(abs(channel.los_doppler_frequency), abs(channel.doppler_frequency)) / bandwidth))
