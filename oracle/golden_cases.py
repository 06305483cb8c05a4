"""Case table shared by the golden-vector generator and the tests that replay the vectors.

Each case is (name, builder, ntx, nrx, fs, T, tx_pos, rx_pos); ``builder(M)`` constructs the channel from
a module ``M`` that exposes the reference's public names (``hermespy.channel`` when generating,
``hermespy_b200.channel`` when replaying) -- the same source text drives both implementations.
"""
import numpy as np


def _corr2():
    return np.array([[1.0, 0.5 + 0.1j], [0.5 - 0.1j, 1.0]], dtype=complex)


FADING_CASES = [
    ("tdl_b_4x4_c2", lambda M: M.TDL(M.TDLType.B, rms_delay=300e-9, doppler_frequency=100, seed=42,
                                      antenna_correlation=M.StandardAntennaCorrelation(M.CorrelationType.MEDIUM)),
     4, 4, 30.72e6, 384, (0, 0, 0), (50, 10, 0)),
    ("tdl_a_siso_flat_c1", lambda M: M.TDL(M.TDLType.A, seed=7), 1, 1, 4e8, 500, (0, 0, 0), (10, 0, 0)),
    ("tdl_e_2x4_medium_a", lambda M: M.TDL(M.TDLType.E, rms_delay=1e-7, doppler_frequency=10, seed=7,
                                            antenna_correlation=M.StandardAntennaCorrelation(M.CorrelationType.MEDIUM_A)),
     2, 4, 30.72e6, 200, (0, 0, 0), (10, 0, 0)),
    ("tdl_d_2x2_los", lambda M: M.TDL(M.TDLType.D, rms_delay=1e-7, doppler_frequency=1e4, seed=8),
     2, 2, 30.72e6, 300, (0, 0, 0), (10, 0, 0)),
    ("tdl_c_2x2_dual_consistent", lambda M: M.TDL(M.TDLType.C, rms_delay=1e-6, seed=9, correlation_distance=25.0),
     2, 2, 30.72e6, 128, (1.0, 2.0, 3.0), (40.0, -5.0, 1.0)),
    ("cost259_hilly_2x2", lambda M: M.Cost259(M.Cost259Type.HILLY, doppler_frequency=50, seed=5),
     2, 2, 30.72e6, 160, (0, 0, 0), (10, 0, 0)),
    ("cost259_urban_siso_c5", lambda M: M.Cost259(M.Cost259Type.URBAN, doppler_frequency=50, seed=5, gain=0.3),
     1, 1, 30.72e6, 1024, (0, 0, 0), (10, 0, 0)),
    ("cost259_rural_siso", lambda M: M.Cost259(M.Cost259Type.RURAL, seed=123), 1, 1, 30.72e6, 100, (0, 0, 0), (10, 0, 0)),
    ("exponential_1x3_fast", lambda M: M.Exponential(1e-7, 3e-7, doppler_frequency=1e6, seed=6),
     1, 3, 1e7, 99, (0, 0, 0), (10, 0, 0)),
    ("custom_rician_2x2", lambda M: M.MultipathFadingChannel(
        [3e-7, 0.0, 1e-7], [0.2, 1.0, 0.5], [0.0, 3.0, np.inf], num_sinusoids=7, doppler_frequency=30.0,
        los_doppler_frequency=5.0, seed=11, antenna_correlation=M.CustomAntennaCorrelation(_corr2())),
     2, 2, 30.72e6, 250, (0, 0, 0), (10, 0, 0)),
    # reference unit test set-up (tests/unit_tests/channel/test_fading.py:108-176): Doppler up to 50 * fs
    ("extreme_doppler_2x2", lambda M: M.MultipathFadingChannel(
        np.linspace(0, 99e-9, 10), np.linspace(1.0, 0.1, 10), np.r_[2.0, np.zeros(9)], num_sinusoids=5,
        doppler_frequency=37.3e9, los_doppler_frequency=11.9e9, seed=42),
     2, 2, 1e9, 100, (0, 0, 0), (10, 0, 0)),
]

SAMPLE_FIELDS = ("power_profile delay_profile los_angles nlos_angles los_phases nlos_phases los_gains "
                 "nlos_gains spatial_response").split()


def golden_signal(case_index: int, num_streams: int, num_samples: int) -> np.ndarray:
    rng = np.random.default_rng(1000 + case_index)
    return (rng.standard_normal((num_streams, num_samples)) + 1j * rng.standard_normal((num_streams, num_samples))) / np.sqrt(2)


# ---- cluster delay line ------------------------------------------------------------------------------------
# (name, channel builder, tx spec, rx spec, T); device spec = (array dims, roll-pitch-yaw, position, velocity)
CDL_FS = 30.72e6
CDL_FC = 3.5e9
CDL_SPACING = 0.5 * 299792458.0 / CDL_FC

CDL_CASES = [
    ("cdl_c_8x2_moving_c3", lambda M: M.CDL(M.CDLType.C, 300e-9, seed=42),
     ((4, 2, 1), (0, 0, 0), (0.0, 0.0, 10.0), (0, 0, 0)), ((2, 1, 1), (0, 0, 0), (100.0, 20.0, 1.5), (10.0, -3.0, 0.0)), 200),
    ("cdl_a_4x4_rotated", lambda M: M.CDL(M.CDLType.A, 300e-9, seed=43),
     ((2, 2, 1), (0.1, 0.2, 0.3), (0.0, 0.0, 10.0), (1.0, 2.0, 0.5)),
     ((2, 2, 1), (-0.4, 0.1, 2.0), (100.0, 20.0, 1.5), (10.0, -3.0, 0.0)), 128),
    ("cdl_d_los_4x2", lambda M: M.CDL(M.CDLType.D, 300e-9, rayleigh_factor=9.0, seed=44),
     ((2, 1, 2), (0.1, 0.2, 0.3), (0.0, 0.0, 10.0), (0, 0, 0)), ((1, 2, 1), (0, 0, 0), (100.0, 20.0, 1.5), (10.0, -3.0, 0.0)), 100),
    ("cdl_e_los_siso_static", lambda M: M.CDL(M.CDLType.E, 1e-7, rayleigh_factor=3.0, seed=45),
     ((1, 1, 1), (0, 0, 0), (0.0, 0.0, 10.0), (0, 0, 0)), ((1, 1, 1), (0, 0, 0), (30.0, -20.0, 1.5), (0, 0, 0)), 64),
    ("cdl_b_2x2_fast", lambda M: M.CDL(M.CDLType.B, 1e-6, seed=46),
     ((2, 1, 1), (0, 0, 0), (0.0, 0.0, 10.0), (0, 0, 0)), ((2, 1, 1), (0, 0, 0), (30.0, -20.0, 1.5), (30.0, 0.0, 0.0)), 64),
]

CDL_SAMPLE_FIELDS = ("azimuth_of_arrival zenith_of_arrival azimuth_of_departure zenith_of_departure cluster_delays "
                     "cluster_powers polarization_transformations").split()
