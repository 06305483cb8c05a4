"""CPU: the numpy oracle and the host-side channel classes replay the committed golden vectors
(generated from the unmodified reference by oracle/make_golden.py)."""
import os

import numpy as np
import pytest

import hermespy_b200.channel as MC
from hermespy_b200.core import SimulatedDevice, SimulatedIdealAntenna, SimulatedUniformArray, Transformation
from oracle import fading_oracle as fo
from oracle.golden_cases import FADING_CASES, SAMPLE_FIELDS, golden_signal

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "fading_golden.npz")


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def params_from_golden(g, name, ntx, nrx):
    sc = g[f"{name}/scalars"]
    return fo.FadingParams(
        power=g[f"{name}/power_profile"], delay=g[f"{name}/delay_profile"], los_gain=g[f"{name}/los_gains"],
        nlos_gain=g[f"{name}/nlos_gains"], los_angle=g[f"{name}/los_angles"], nlos_angle=g[f"{name}/nlos_angles"],
        los_phase=g[f"{name}/los_phases"], nlos_phase=g[f"{name}/nlos_phases"], los_doppler=float(sc[0]),
        nlos_doppler=float(sc[1]), spatial=g[f"{name}/spatial_response"], gain=float(sc[2]), fs=float(sc[4]),
        num_rx=nrx, num_tx=ntx)


def mirror_device(n, fs, pos):
    return SimulatedDevice(bandwidth=fs, oversampling_factor=1, carrier_frequency=3.5e9,
                           antennas=SimulatedUniformArray(SimulatedIdealAntenna, 0.04, (n, 1, 1)),
                           pose=Transformation.From_Translation(np.array(pos, dtype=float)))


def mirror_sample(case):
    name, build, ntx, nrx, fs, T, ptx, prx = case
    ch = build(MC)
    tx, rx = mirror_device(ntx, fs, ptx), mirror_device(nrx, fs, prx)
    ch.realize()
    real = ch.realize()
    return real, real.sample(tx, rx), tx, rx


@pytest.mark.parametrize("ci", range(len(FADING_CASES)), ids=[c[0] for c in FADING_CASES])
def test_oracle_reproduces_reference_outputs(golden, ci):
    name, _, ntx, nrx, fs, T, _, _ = FADING_CASES[ci]
    p = params_from_golden(golden, name, ntx, nrx)
    y = fo.propagate(p, golden_signal(ci, ntx, T))
    ref = golden[f"{name}/y"]
    assert y.shape == ref.shape
    # same operation order as the reference -> agreement to rounding of the final matrix product
    assert np.abs(y - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max())
    yr = fo.propagate(p.reciprocal(), golden_signal(100 + ci, nrx, T))
    assert np.abs(yr - golden[f"{name}/y_reciprocal"]).max() <= 1e-13 * max(1.0, np.abs(ref).max())
    assert fo.expected_energy_scale(p) == pytest.approx(float(golden[f"{name}/scalars"][3]), rel=1e-15)
    if f"{name}/csi" in golden.files:
        csi = golden[f"{name}/csi"]
        mine = fo.state(p, T, csi.shape[3])
        assert mine.shape == csi.shape
        assert np.abs(mine - csi).max() <= 1e-13


@pytest.mark.parametrize("ci", range(len(FADING_CASES)), ids=[c[0] for c in FADING_CASES])
def test_host_classes_reproduce_reference_parameters(golden, ci):
    """Same constructor text, same seeds -> bit-identical sample parameters (RNG draw order parity)."""
    case = FADING_CASES[ci]
    name = case[0]
    real, s, tx, rx = mirror_sample(case)
    for f in SAMPLE_FIELDS:
        want = golden[f"{name}/{f}"]
        got = np.asarray(getattr(s, f))
        assert got.shape == want.shape, f
        assert np.array_equal(got, want), f
    sc = golden[f"{name}/scalars"]
    assert (s.los_doppler, s.nlos_doppler, s.gain) == (sc[0], sc[1], sc[2])
    assert s.expected_energy_scale == sc[3]
    rs = real.reciprocal_sample(s, rx, tx)
    assert np.array_equal(rs.spatial_response, golden[f"{name}/spatial_response"].T)
    assert rs.num_transmit_antennas == case[3] and rs.num_receive_antennas == case[2]


def test_kernel_parameter_block_matches_oracle_rates(golden):
    """The flat (omega, phi, amp) block the kernels read equals the oracle's restatement of fading.py:326-342."""
    for ci, case in enumerate(FADING_CASES):
        name, _, ntx, nrx, fs, T, _, _ = case
        _, s, _, _ = mirror_sample(case)
        b = s.kernel_block()
        p = params_from_golden(golden, name, ntx, nrx)
        om, ph, am = fo.sinusoid_rates(p)
        assert np.array_equal(b["omega"], om)
        assert np.array_equal(b["phi"], ph)
        assert np.array_equal(b["amp"][:, 0], am[:, 0])
        assert np.array_equal(b["amp"][:, 1], am[:, 1])
        assert np.array_equal(b["tap_delay"], fo.tap_delays_in_samples(p))
        assert b["max_delay"] == fo.max_delay_in_samples(p)
        assert b["omega_max"] >= np.abs(om).max()


def test_sum_of_sinusoids_restatement_equals_operation_order_form(golden):
    """h via (omega, phi, amp) equals the reference-ordered evaluation to rounding (moderate Doppler cases)."""
    for ci, case in enumerate(FADING_CASES):
        name, _, ntx, nrx, fs, T, _, _ = case
        if name.startswith("extreme"):
            continue
        p = params_from_golden(golden, name, ntx, nrx)
        om, ph, am = fo.sinusoid_rates(p)
        n = np.arange(T)
        h = (am[:, :, None] * np.exp(1j * (om[:, :, None] * n[None, None, :] + ph[:, :, None]))).sum(axis=1)
        assert np.abs(h - fo.tap_impulses(p, T)).max() < 1e-11
