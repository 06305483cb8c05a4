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


# ---- cluster delay line: oracle parameter block <-> kernel parameter block (independent of the host classes) -------

def cdl_params_from_golden(golden, name):
    """``oracle.cdl_oracle.CdlParams`` of a case stored by ``oracle/make_golden_elements.py``."""
    from oracle import cdl_oracle as co

    sc = golden[f"{name}/scalars"]

    def geom(side):
        return co.ArrayGeometry(rotation=golden[f"{name}/{side}_rotation"], translation=golden[f"{name}/{side}_translation"],
                                topology=golden[f"{name}/{side}_topology"], velocity=golden[f"{name}/{side}_velocity"],
                                elements=golden[f"{name}/{side}_elements"])

    return co.CdlParams(line_of_sight=bool(sc[0]), rice_factor_db=float(sc[1]), aoa=golden[f"{name}/aoa"],
                        zoa=golden[f"{name}/zoa"], aod=golden[f"{name}/aod"], zod=golden[f"{name}/zod"],
                        delay_offset=float(sc[2]), cluster_delays=golden[f"{name}/cluster_delays"],
                        cluster_delay_spread=float(sc[3]), cluster_powers=golden[f"{name}/cluster_powers"],
                        jones=golden[f"{name}/jones"], tx=geom("tx"), rx=geom("rx"), fc=float(sc[4]), fs=float(sc[5]))


def cdl_block_from_oracle_params(p):
    """``kernels.CdlBlock`` (B = 1) laid out from an oracle ``CdlParams`` -- the ray-term order of
    cluster_delay_lines.py:409-496 restated for the tests, so GPU parity does not lean on the product's own extraction."""
    from math import ceil

    from hermespy_b200.kernels import CdlBlock
    from oracle import cdl_oracle as co

    C, R = p.aoa.shape
    nsplit = min(2, C)
    sub = (p.cluster_delays[:nsplit, None] + p.cluster_delay_spread * np.array([0.0, 1.28, 2.56])[None, :]).ravel()
    vdelays = np.concatenate((sub, p.cluster_delays[nsplit:]))
    cs, rs, ds = [], [], []
    for v in range(3 * nsplit + max(0, C - 2)):
        c = int(v / 3) if v < 6 else v - 4
        for r in (co.SUBCLUSTER_RAYS[c] if c < nsplit else range(R)):
            cs.append(c), rs.append(r), ds.append(vdelays[v])
    c, r = np.array(cs), np.array(rs)
    rice = 10.0 ** (p.rice_factor_db / 10.0)
    nlos = (1.0 + rice) ** -0.5 if p.line_of_sight else 1.0
    pose = lambda g: np.concatenate([np.asarray(g.rotation).ravel(), np.asarray(g.translation)])[None]
    return CdlBlock(
        term_delay=np.array([int((d + p.delay_offset) * p.fs) for d in ds], dtype=np.int32),
        max_delay=ceil(p.max_delay * p.fs),
        angles=np.stack([p.aoa[c, r], p.zoa[c, r], p.aod[c, r], p.zod[c, r]], axis=1)[None],
        jones=np.ascontiguousarray(np.asarray(p.jones)[:, :, c, r].transpose(2, 0, 1))[None],
        amplitude=(np.sqrt(np.asarray(p.cluster_powers)[c] / R) * nlos)[None], tx_pose=pose(p.tx), rx_pose=pose(p.rx),
        rel_velocity=(np.asarray(p.rx.velocity, float) - np.asarray(p.tx.velocity, float))[None],
        tx_topology=p.tx.topology, rx_topology=p.rx.topology, carrier_frequency=p.fc, sampling_rate=p.fs,
        line_of_sight=p.line_of_sight, los_delay=int((p.cluster_delays[0] + p.delay_offset) * p.fs),
        los_amplitude=float((rice / (1 + rice)) ** 0.5), tx_elements=p.tx.elements, rx_elements=p.rx.elements)
