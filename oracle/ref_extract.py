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


def element_table_from_reference(antennas) -> np.ndarray:
    """[M, 12] element rows (rotation element -> array frame, kind, parameter) of reference ``Antenna`` objects."""
    from hermespy.core.antennas import Dipole, IdealAntenna, LinearAntenna, PatchAntenna

    rows = np.zeros((len(antennas), 12))
    for m, a in enumerate(antennas):
        rows[m, :9] = np.asarray(a.pose, dtype=np.float64)[:3, :3].ravel()
        if isinstance(a, LinearAntenna):
            rows[m, 9:11] = 1, a.slant
        elif isinstance(a, PatchAntenna):
            rows[m, 9] = 2
        elif isinstance(a, Dipole):
            rows[m, 9] = 3
        elif not isinstance(a, IdealAntenna):
            raise TypeError(f"no oracle pattern for antenna type {type(a).__name__}")
    return rows


def array_geometry_from_reference(antennas_state, velocity, mode):
    """``ArrayGeometry`` of a reference ``AntennaArrayState`` (pose = local -> global homogeneous matrix)."""
    from hermespy.core import AntennaMode

    from .cdl_oracle import ArrayGeometry

    fwd = np.asarray(antennas_state.forwards_transformation, dtype=np.float64)
    ants = antennas_state.transmit_antennas if mode == AntennaMode.TX else antennas_state.receive_antennas
    return ArrayGeometry(rotation=fwd[:3, :3].copy(), translation=fwd[:3, 3].copy(),
                         topology=np.asarray(antennas_state._topology(mode), dtype=np.float64).copy(),
                         velocity=np.asarray(velocity, dtype=np.float64).copy(),
                         elements=element_table_from_reference(list(ants)))


def cdl_params_from_reference_sample(sample):
    """Read a reference ``ClusterDelayLineSample`` through its public properties (cluster_delay_lines.py:321-403)."""
    from hermespy.core import AntennaMode

    from .cdl_oracle import CdlParams

    return CdlParams(
        line_of_sight=bool(sample.line_of_sight), rice_factor_db=float(sample.rice_factor),
        aoa=np.array(sample.azimuth_of_arrival), zoa=np.array(sample.zenith_of_arrival),
        aod=np.array(sample.azimuth_of_departure), zod=np.array(sample.zenith_of_departure),
        delay_offset=float(sample.delay_offset), cluster_delays=np.array(sample.cluster_delays, dtype=np.float64),
        cluster_delay_spread=float(sample.cluster_delay_spread), cluster_powers=np.array(sample.cluster_powers),
        jones=np.array(sample.polarization_transformations),
        tx=array_geometry_from_reference(sample.transmitter_antennas, sample.transmitter_velocity, AntennaMode.TX),
        rx=array_geometry_from_reference(sample.receiver_antennas, sample.receiver_velocity, AntennaMode.RX),
        fc=float(sample.carrier_frequency), fs=float(sample.bandwidth))
