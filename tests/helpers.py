"""Shared test helpers: seeded synthetic fading links (independent of the reference tree)."""
from __future__ import annotations

import numpy as np

from oracle import fading_oracle as fo


def random_fading_params(rng, L, N, ntx, nrx, fs, doppler, max_delay_s, los_doppler=None, rice=None, gain=1.0,
                         delays=None, powers=None):
    if delays is None:
        delays = np.sort(rng.uniform(0, max_delay_s, L))
        delays[0] = 0.0
    if powers is None:
        powers = rng.uniform(0.05, 1.0, L)
        powers /= powers.sum()
    if rice is None:
        rice = np.zeros(L)
    g = rng.standard_normal(fo.num_scalars(L, N))
    return fo.params_from_normals(
        g, delays=np.asarray(delays, float), powers=np.asarray(powers, float), rice_factors=np.asarray(rice, float),
        num_sinusoids=N, doppler=doppler, los_doppler=los_doppler, gain=gain, fs=fs, num_rx=nrx, num_tx=ntx)


def random_signal(rng, ntx, T):
    return (rng.standard_normal((ntx, T)) + 1j * rng.standard_normal((ntx, T))) / np.sqrt(2.0)


def stack_param_blocks(plist):
    """Stack oracle parameter blocks of links sharing a delay profile into kernel arrays."""
    om, ph, am, S = [], [], [], []
    for p in plist:
        o, f, a = fo.sinusoid_rates(p)
        om.append(o)
        ph.append(f)
        am.append(np.stack([a[:, 0], a[:, 1] if a.shape[1] > 1 else np.zeros(len(a))], axis=1))
        S.append(p.spatial[: p.num_rx, : p.num_tx])
    p0 = plist[0]
    return dict(
        tap_delay=fo.tap_delays_in_samples(p0).astype(np.int32),
        max_delay=fo.max_delay_in_samples(p0),
        omega=np.stack(om), phi=np.stack(ph), amp=np.stack(am), spatial=np.stack(S),
    )


def rel_l2(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    d = np.linalg.norm((a - b).ravel())
    n = np.linalg.norm(b.ravel())
    return d / n if n > 0 else d
