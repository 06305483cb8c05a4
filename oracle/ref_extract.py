"""Helpers that turn live reference objects into oracle parameter blocks (build container only)."""
from __future__ import annotations

import numpy as np

from .fading_oracle import FadingParams


def fading_params_from_reference_sample(sample) -> FadingParams:
    """Read a reference ``MultipathFadingSample`` through its public properties (fading.py:217-291)."""
    return FadingParams(
        power=np.array(sample.power_profile, dtype=np.float64),
        delay=np.array(sample.delay_profile, dtype=np.float64),
        los_gain=np.array(sample.los_gains, dtype=np.float64),
        nlos_gain=np.array(sample.nlos_gains, dtype=np.float64),
        los_angle=np.array(sample.los_angles, dtype=np.float64),
        nlos_angle=np.array(sample.nlos_angles, dtype=np.float64),
        los_phase=np.array(sample.los_phases, dtype=np.float64),
        nlos_phase=np.array(sample.nlos_phases, dtype=np.float64),
        los_doppler=float(sample.los_doppler),
        nlos_doppler=float(sample.nlos_doppler),
        spatial=np.array(sample.spatial_response, dtype=np.complex128),
        gain=float(sample.gain),
        fs=float(sample.bandwidth),
        num_rx=int(sample.num_receive_antennas),
        num_tx=int(sample.num_transmit_antennas),
    )


def static_normals_of(realization) -> np.ndarray:
    """Private normals of a static fading realization (name-mangled)."""
    rr = realization._MultipathFadingRealization__random_realization
    return np.array(rr._StaticConsistentRealization__scalar_samples, dtype=np.float64)
