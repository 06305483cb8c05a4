"""CPU: the CDL numpy oracle and the host-side CDL classes replay the reference's golden vectors."""
import os

import numpy as np
import pytest

import hermespy_b200.channel as MC
from hermespy_b200.core import SimulatedDevice, SimulatedIdealAntenna, SimulatedUniformArray, Transformation
from oracle import cdl_oracle as co
from oracle.golden_cases import CDL_CASES, CDL_FC, CDL_FS, CDL_SAMPLE_FIELDS, CDL_SPACING, golden_signal

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "cdl_golden.npz")


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def mirror_cdl_device(spec):
    dims, rpy, pos, vel = spec
    return SimulatedDevice(bandwidth=CDL_FS, oversampling_factor=1, carrier_frequency=CDL_FC,
                           antennas=SimulatedUniformArray(SimulatedIdealAntenna, CDL_SPACING, dims),
                           pose=Transformation.From_RPY(np.array(rpy, float), np.array(pos, float)),
                           velocity=np.array(vel, float))


def mirror_cdl_sample(case):
    name, build, txs, rxs, T = case
    ch = build(MC)
    tx, rx = mirror_cdl_device(txs), mirror_cdl_device(rxs)
    ch.realize()
    real = ch.realize()
    return real, real.sample(tx, rx), tx, rx


def geometry(state):
    return co.ArrayGeometry(rotation=state.pose.rotation.copy(), translation=state.pose.translation.copy(),
                            topology=state.antennas.topology, velocity=np.asarray(state.velocity, float))


def oracle_params(s) -> co.CdlParams:
    """Oracle parameter block from a (mirror) sample's public properties."""
    return co.CdlParams(
        line_of_sight=bool(s.line_of_sight), rice_factor_db=float(s.rice_factor), aoa=s.azimuth_of_arrival,
        zoa=s.zenith_of_arrival, aod=s.azimuth_of_departure, zod=s.zenith_of_departure,
        delay_offset=float(s.delay_offset), cluster_delays=np.asarray(s.cluster_delays, float),
        cluster_delay_spread=float(s.cluster_delay_spread), cluster_powers=s.cluster_powers,
        jones=s.polarization_transformations, tx=geometry(s.transmitter_state), rx=geometry(s.receiver_state),
        fc=s.carrier_frequency, fs=s.bandwidth)


@pytest.mark.parametrize("ci", range(len(CDL_CASES)), ids=[c[0] for c in CDL_CASES])
def test_host_classes_reproduce_reference_parameters(golden, ci):
    case = CDL_CASES[ci]
    name = case[0]
    _, s, _, _ = mirror_cdl_sample(case)
    for f in CDL_SAMPLE_FIELDS:
        want = golden[f"{name}/{f}"]
        got = np.asarray(getattr(s, f))
        assert got.shape == want.shape, f
        assert np.array_equal(got, want), f
    sc = golden[f"{name}/scalars"]
    assert bool(sc[0]) == s.line_of_sight and sc[1] == s.rice_factor and sc[2] == s.delay_offset
    assert sc[3] == s.cluster_delay_spread and sc[4] == s.max_delay and sc[5] == s.expected_energy_scale


@pytest.mark.parametrize("ci", range(len(CDL_CASES)), ids=[c[0] for c in CDL_CASES])
def test_oracle_reproduces_reference_outputs(golden, ci):
    case = CDL_CASES[ci]
    name, _, txs, rxs, T = case
    _, s, _, _ = mirror_cdl_sample(case)
    p = oracle_params(s)
    ntx, nrx = int(np.prod(txs[0])), int(np.prod(rxs[0]))
    y = co.propagate(p, golden_signal(200 + ci, ntx, T))
    ref = golden[f"{name}/y"]
    assert y.shape == ref.shape
    # distance phases are ~7e3 rad: rotated poses differ from the reference's kinematic chain at 1e-12
    assert np.linalg.norm(y - ref) <= 5e-12 * np.linalg.norm(ref)
    yr = co.propagate(p.reciprocal(), golden_signal(300 + ci, nrx, T))
    assert np.linalg.norm(yr - golden[f"{name}/y_reciprocal"]) <= 5e-12 * np.linalg.norm(ref)
    if f"{name}/csi" in golden.files:
        csi = golden[f"{name}/csi"]
        mine = co.state(p, T, 1000)
        assert mine.shape == csi.shape
        assert np.abs(mine - csi).max() <= 5e-12 * np.abs(csi).max()


def test_kernel_block_term_layout(golden):
    """488 ray terms for CDL-C and term order equal to the oracle's generator order."""
    case = CDL_CASES[0]
    _, s, _, _ = mirror_cdl_sample(case)
    blk = s.kernel_block()
    assert blk.term_delay.shape == (488,)
    terms = co.ray_terms(oracle_params(s))
    assert len(terms) == 488
    want = [int((t[2] + s.delay_offset) * s.bandwidth) for t in terms]
    assert np.array_equal(blk.term_delay, np.array(want))
    assert blk.max_delay == co.max_delay_in_samples(oracle_params(s))
    los = mirror_cdl_sample(CDL_CASES[2])[1].kernel_block()
    assert los.line_of_sight and 0 < los.los_amplitude < 1
