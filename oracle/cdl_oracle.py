"""Float64 numpy restatement of HermesPy's 3GPP cluster-delay-line hot path (test oracle).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Pinned against the live reference
(``tests/test_oracle_vs_reference.py``) and against golden vectors generated from it
(``tests/golden/cdl_golden.npz``).

Reference functions restated here:

* per-ray MIMO matrices, Doppler phasors, delays .... hermespy/channel/cdl/cluster_delay_lines.py:409-523
* propagate (truncating integer delays) ............. cluster_delay_lines.py:526-558
* dense channel state ............................... cluster_delay_lines.py:561-592
* antenna array response (distance phase x polarization), with the unit direction vector passed as
  a *global position*  ................. hermespy/core/antennas.py:954-1000 (array response),
                                                      :883-952 (phase response), :797-836 (characteristics),
                                                      :138-210 (polarization transformation, TR 38.901 7.1-11..14)
* static CDL sample construction (tables -> rays) .. hermespy/channel/cdl/cdl.py:222-299
* reciprocal sample ................................. cluster_delay_lines.py:732-756
"""
from __future__ import annotations

from dataclasses import dataclass, replace
from math import ceil
from typing import List, Optional, Tuple

import numpy as np

SPEED_OF_LIGHT = 299792458.0  # scipy.constants.speed_of_light

#: Sub-cluster ray partitions; the reference indexes this table by *cluster* index.
SUBCLUSTER_RAYS = ([0, 1, 2, 3, 4, 5, 6, 7, 18, 19], [8, 9, 10, 11, 16, 17], [12, 13, 14, 15])


@dataclass
class ArrayGeometry:
    """Antenna array frozen in the global frame: pose (rotation, translation) and element positions."""

    rotation: np.ndarray  # [3, 3] array-local -> global
    translation: np.ndarray  # [3]
    topology: np.ndarray  # [M, 3] element positions in the array frame
    velocity: np.ndarray  # [3] global
    #: per-element models [M, 12]: rotation element frame -> array frame (9, row-major), kind (0 ideal, 1 linear,
    #: 2 patch, 3 dipole), kind parameter (linear: slant), reserved; None = unrotated ideal elements
    elements: Optional[np.ndarray] = None


@dataclass
class CdlParams:
    line_of_sight: bool
    rice_factor_db: float
    aoa: np.ndarray  # [C, R] radians
    zoa: np.ndarray
    aod: np.ndarray
    zod: np.ndarray
    delay_offset: float
    cluster_delays: np.ndarray  # [C] seconds
    cluster_delay_spread: float
    cluster_powers: np.ndarray  # [C]
    jones: np.ndarray  # [2, 2, C, R] complex
    tx: ArrayGeometry
    rx: ArrayGeometry
    fc: float
    fs: float

    @property
    def max_delay(self) -> float:
        """cluster_delay_lines.py:316-319."""
        return max(np.max(self.cluster_delays[:3] + self.cluster_delay_spread * 2.56), self.cluster_delays.max()) + self.delay_offset

    def reciprocal(self) -> "CdlParams":
        """Arrival and departure angles swap, devices swap (cluster_delay_lines.py:732-756)."""
        return replace(self, aoa=self.aod, zoa=self.zod, aod=self.aoa, zod=self.zoa, tx=self.rx, rx=self.tx)


def unit_vector(azimuth: float, zenith: float) -> np.ndarray:
    """hermespy/core/transformation.py:29-44."""
    return np.array([np.sin(zenith) * np.cos(azimuth), np.sin(zenith) * np.sin(azimuth), np.cos(zenith)])


def to_spherical(v: np.ndarray) -> Tuple[float, float]:
    """transformation.py:66-84 (zenith = arccos(z) assumes a normalized vector)."""
    return float(np.arctan2(v[1], v[0])), float(np.arccos(v[2]))


def local_pattern(kind: int, param: float, azimuth: float, zenith: float) -> np.ndarray:
    """``Antenna.local_characteristics`` of the reference's four element models (antennas.py:435-436 ideal,
    :509-510 linear, :556-560 patch, :610-614 dipole).  The reference passes the ZENITH angle to the parameter it
    names "elevation" (antennas.py:157)."""
    if kind == 1:
        return np.array([np.cos(param), np.sin(param)])
    if kind == 2:
        vertical_azimuth = 0.1 + 0.9 * np.exp(-1.315 * azimuth**2)
        return np.array([max(0.1, vertical_azimuth * np.cos(zenith) ** 2), 0.0])
    if kind == 3:
        return np.array([0.0 if zenith == 0.0 else np.cos(0.5 * np.pi * np.cos(zenith)) / np.sin(zenith), 0.0])
    return np.array([2**-0.5, 2**-0.5])


def element_polarization(R: np.ndarray, global_direction: np.ndarray, kind: int = 0, param: float = 0.0) -> np.ndarray:
    """Polarization 2-vector of an element with orientation ``R`` (element frame -> global) seen from
    ``global_direction`` (antennas.py:138-210): the local pattern rotated into the global theta/phi basis by the
    2x2 matrix of TR 38.901 eq. 7.1-12."""
    local_direction = R.T @ global_direction
    az_g, ze_g = to_spherical(global_direction)
    az_l, ze_l = to_spherical(local_direction)
    phi_g = np.array([-np.sin(az_g), np.cos(az_g), 0.0])
    phi_l = np.array([-np.sin(az_l), np.cos(az_l), 0.0])
    th_g = np.array([np.cos(ze_g) * np.cos(az_g), np.cos(ze_g) * np.sin(az_g), -np.sin(ze_g)])
    th_l = np.array([np.cos(ze_l) * np.cos(az_l), np.cos(ze_l) * np.sin(az_l), -np.sin(ze_l)])
    th_lt, phi_lt = R @ th_l, R @ phi_l
    pt = np.array([[th_g @ th_lt, th_g @ phi_lt], [phi_g @ th_lt, phi_g @ phi_lt]])
    return pt @ local_pattern(kind, param, az_l, ze_l)


def ideal_polarization(geom: ArrayGeometry, global_direction: np.ndarray) -> np.ndarray:
    """Unrotated ideal isotropic element: constant local pattern [2^-1/2, 2^-1/2], orientation = the array's."""
    return element_polarization(geom.rotation, global_direction)


def array_response(geom: ArrayGeometry, fc: float, target_position: np.ndarray) -> np.ndarray:
    """``cartesian_array_response(fc, position, 'global', mode)`` -> [M, 2] (antennas.py:954-1000).

    phase_m = exp(-2j pi fc / c * || q_m - T^-1(position) ||), polarization of element m towards
    normalize(position - t) (antennas.py:797-836: one ``global_characteristics`` call per element).
    """
    local_position = geom.rotation.T @ (np.asarray(target_position, dtype=np.float64) - geom.translation)
    distances = np.linalg.norm(geom.topology.T - local_position[:, None], axis=0)
    phase = np.exp(-2j * np.pi * fc * distances / SPEED_OF_LIGHT)
    d = np.asarray(target_position, dtype=np.float64) - geom.translation
    d = d / np.linalg.norm(d)
    if geom.elements is None:
        pol = ideal_polarization(geom, d)
        return phase[:, None] * pol[None, :]
    pol = np.stack([element_polarization(geom.rotation @ e[:9].reshape(3, 3), d, int(e[9]), float(e[10]))
                    for e in np.asarray(geom.elements, dtype=np.float64)])
    return phase[:, None] * pol


def ray_terms(p: CdlParams) -> List[Tuple[np.ndarray, float, float]]:
    """All (H[Nrx, Ntx], radial_speed, delay_seconds) terms in the reference's order
    (cluster_delay_lines.py:409-523).  ``radial_speed = <wave vector, v_rx - v_tx>`` in m/s; the Doppler phasor of
    a term is exp(radial_speed * (fc / c * n / fs) * 2j pi), formed in that order as the reference does."""
    rice_lin = 10.0 ** (p.rice_factor_db / 10.0)
    nlos_scale = (1.0 + rice_lin) ** -0.5 if p.line_of_sight else 1.0
    C, R = p.aoa.shape
    nsplit = min(2, C)
    nvirtual = 3 * nsplit + max(0, C - 2)
    sub = (p.cluster_delays[:nsplit, None] + p.cluster_delay_spread * np.array([0.0, 1.28, 2.56])[None, :]).ravel()
    vdelays = np.concatenate((sub, p.cluster_delays[nsplit:]))
    wl = p.fc / SPEED_OF_LIGHT
    rel_v = p.rx.velocity - p.tx.velocity
    out = []
    for v in range(nvirtual):
        c = int(v / 3) if v < 6 else v - 4
        rays = SUBCLUSTER_RAYS[c] if c < nsplit else range(R)
        for r in rays:
            a_tx = array_response(p.tx, p.fc, unit_vector(p.aod[c, r], p.zod[c, r]))
            a_rx = array_response(p.rx, p.fc, unit_vector(p.aoa[c, r], p.zoa[c, r]))
            H = a_rx @ p.jones[:, :, c, r] @ a_tx.T * (np.sqrt(p.cluster_powers[c] / R) * nlos_scale)
            wave = unit_vector(p.aoa[c, r], p.zoa[c, r])
            out.append((H, float(np.inner(wave, rel_v)), float(vdelays[v])))
    if p.line_of_sight:
        dvec = p.rx.translation - p.tx.translation
        dist = np.linalg.norm(dvec, 2)
        a_tx = array_response(p.tx, p.fc, p.rx.translation)
        a_rx = array_response(p.rx, p.fc, p.tx.translation)
        H = a_rx @ np.array([[1, 0], [-1, 0]]) @ a_tx.T * (rice_lin / (1 + rice_lin)) ** 0.5
        H = H * np.exp(-2j * np.pi * dist * wl)
        out.append((H, float(np.inner(dvec / dist, rel_v)), float(p.cluster_delays[0])))
    return out


def max_delay_in_samples(p: CdlParams) -> int:
    return ceil(p.max_delay * p.fs)


def propagate(p: CdlParams, x: np.ndarray) -> np.ndarray:
    """``y[:, k : k + T] += H @ (x * e)`` for every ray term, k = int((tau + offset) fs) (cluster_delay_lines.py:526-558)."""
    x = np.asarray(x, dtype=np.complex128)
    T = x.shape[1]
    nrx = p.rx.topology.shape[0]
    y = np.zeros((nrx, T + max_delay_in_samples(p)), dtype=np.complex128)
    fast_fading = (p.fc / SPEED_OF_LIGHT) * np.arange(T) / p.fs
    for H, speed, tau in ray_terms(p):
        k = int((tau + p.delay_offset) * p.fs)
        e = np.exp(speed * fast_fading * 2j * np.pi)
        y[:, k : k + T] += H @ (x * e[None, :])
    return y


def state(p: CdlParams, num_samples: int, max_num_taps: int) -> np.ndarray:
    """Dense CSI [Nrx, Ntx, T, 1 + D] (cluster_delay_lines.py:561-592)."""
    D = min(max_num_taps, max_delay_in_samples(p))
    nrx, ntx = p.rx.topology.shape[0], p.tx.topology.shape[0]
    raw = np.zeros((nrx, ntx, num_samples, 1 + D), dtype=np.complex128)
    fast_fading = (p.fc / SPEED_OF_LIGHT) * np.arange(num_samples) / p.fs
    for H, speed, tau in ray_terms(p):
        k = int((tau + p.delay_offset) * p.fs)
        if k >= max_num_taps:
            continue
        raw[:, :, :, k] += H[:, :, None] * np.exp(speed * fast_fading * 2j * np.pi)[None, None, :]
    return raw


def expected_energy_scale(p: CdlParams) -> float:
    return float(np.sum(p.cluster_powers))
