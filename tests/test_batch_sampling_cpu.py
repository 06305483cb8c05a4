"""Host logic (no GPU): the vectorized link sampler that feeds bench.py / tools/config_report.py draws exactly what B
sequential ``channel.realize().sample(tx, rx)`` calls draw -- in the mirror classes and in the live reference."""
import numpy as np
import pytest

import hermespy_b200.channel as MC
from hermespy_b200.batch import sample_fading_links
from hermespy_b200.core import SimulatedDevice, SimulatedIdealAntenna, SimulatedUniformArray

FS = 30.72e6

BUILDERS = {
    "tdl_b_medium_4x4": (lambda M, seed: M.TDL(M.TDLType.B, rms_delay=300e-9, doppler_frequency=100, seed=seed,
                                                antenna_correlation=M.StandardAntennaCorrelation(M.CorrelationType.MEDIUM)), 4, 4),
    "tdl_d_los_2x3": (lambda M, seed: M.TDL(M.TDLType.D, rms_delay=1e-7, doppler_frequency=1e3, seed=seed), 2, 3),
    "tdl_a_flat_siso": (lambda M, seed: M.TDL(M.TDLType.A, seed=seed), 1, 1),
    "cost259_urban_siso": (lambda M, seed: M.Cost259(M.Cost259Type.URBAN, doppler_frequency=50, seed=seed), 1, 1),
    "cost259_hilly_2x2": (lambda M, seed: M.Cost259(M.Cost259Type.HILLY, doppler_frequency=50, seed=seed), 2, 2),
    "exponential_1x3": (lambda M, seed: M.Exponential(1e-7, 3e-7, doppler_frequency=1e4, seed=seed), 1, 3),
}


def _dev(n):
    return SimulatedDevice(bandwidth=FS, antennas=SimulatedUniformArray(SimulatedIdealAntenna, 0.04, (n, 1, 1)))


@pytest.mark.parametrize("name", list(BUILDERS))
def test_batched_sampling_equals_sequential_mirror(name):
    build, ntx, nrx = BUILDERS[name]
    B = 5
    blk = sample_fading_links(build(MC, 42), B, ntx, nrx, FS)
    ch = build(MC, 42)
    tx, rx = _dev(ntx), _dev(nrx)
    for b in range(B):
        one = ch.realize().sample(tx, rx).kernel_block()
        assert np.array_equal(one["tap_delay"], blk["tap_delay"]) and one["max_delay"] == blk["max_delay"]
        for k in ("omega", "phi", "amp", "spatial"):
            assert np.array_equal(np.asarray(one[k]), blk[k][b]), (name, b, k)
    assert blk["omega"].shape[0] == B and blk["spatial"].shape == (B, nrx, ntx)


def test_reciprocal_batch_is_the_transposed_spatial_response():
    build, ntx, nrx = BUILDERS["tdl_d_los_2x3"]
    fwd = sample_fading_links(build(MC, 7), 3, ntx, nrx, FS)
    rev = sample_fading_links(build(MC, 7), 3, ntx, nrx, FS, reciprocal=True)
    assert np.array_equal(rev["spatial"], np.swapaxes(fwd["spatial"], 1, 2))
    for k in ("omega", "phi", "amp"):
        assert np.array_equal(rev[k], fwd[k])


def test_large_arrays_need_max_antennas():
    ch = MC.TDL(MC.TDLType.D, rms_delay=1e-7, seed=1)
    with pytest.raises(ValueError, match="max_antennas"):
        sample_fading_links(ch, 2, 16, 16, FS)
    big = MC.TDL(MC.TDLType.D, rms_delay=1e-7, seed=1, max_antennas=16)
    assert sample_fading_links(big, 2, 16, 16, FS)["spatial"].shape == (2, 16, 16)


def test_batched_sampling_equals_the_live_reference():
    """Same constructor text + seed through the UNMODIFIED reference classes, sampled one by one."""
    from oracle.refload import load_reference, reference_available

    if not reference_available():
        pytest.skip("reference tree not available")
    load_reference()
    import hermespy.channel as RC
    from hermespy.simulation import SimulatedDevice as RDev, SimulatedIdealAntenna as RAnt, SimulatedUniformArray as RArr

    from hermespy_b200.dropin import fading_block_from_reference

    for name in ("tdl_b_medium_4x4", "cost259_hilly_2x2"):
        build, ntx, nrx = BUILDERS[name]
        blk = sample_fading_links(build(MC, 11), 4, ntx, nrx, FS)
        ch = build(RC, 11)
        dev = lambda n: RDev(bandwidth=FS, oversampling_factor=1, carrier_frequency=3.5e9, antennas=RArr(RAnt, 0.04, (n, 1, 1)))
        tx, rx = dev(ntx), dev(nrx)
        for b in range(4):
            one = fading_block_from_reference(ch.realize().sample(tx, rx))
            assert np.array_equal(one["tap_delay"], blk["tap_delay"])
            for k in ("omega", "phi", "amp", "spatial"):
                assert np.array_equal(one[k], blk[k][b]), (name, b, k)
