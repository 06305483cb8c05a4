"""CPU, build container only: pin the numpy oracle and the host classes against the LIVE reference."""
import numpy as np
import pytest

from oracle.refload import reference_available

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not reference_available(), reason="reference tree not mounted")]


@pytest.fixture(scope="module")
def ref():
    from oracle.refload import load_reference

    load_reference()
    import hermespy.channel as RC
    from hermespy.core import Signal, Transformation
    from hermespy.simulation import SimulatedDevice, SimulatedIdealAntenna, SimulatedUniformArray

    def device(n, fs, pos=(0, 0, 0)):
        return SimulatedDevice(bandwidth=fs, oversampling_factor=1, carrier_frequency=3.5e9,
                               antennas=SimulatedUniformArray(SimulatedIdealAntenna, 0.04, (n, 1, 1)),
                               pose=Transformation.From_Translation(np.array(pos, dtype=float)))

    return dict(RC=RC, Signal=Signal, device=device)


def test_oracle_matches_live_reference_random_configurations(ref):
    from oracle import fading_oracle as fo
    from oracle.ref_extract import fading_params_from_reference_sample, static_normals_of

    RC = ref["RC"]
    rng = np.random.default_rng(0)
    for trial in range(6):
        L = int(rng.integers(1, 9))
        N = int(rng.integers(1, 12))
        ntx, nrx = int(rng.integers(1, 5)), int(rng.integers(1, 5))
        fs = float(rng.choice([1e6, 30.72e6, 4e8]))
        T = int(rng.integers(1, 300))
        delays = rng.uniform(0, 20 / fs, L)
        ch = RC.MultipathFadingChannel(delays, rng.uniform(0.1, 1, L), rng.choice([0.0, 1.0, 10.0, np.inf], L),
                                       num_sinusoids=N, doppler_frequency=float(rng.choice([0.0, 50.0, 1e5])),
                                       gain=float(rng.uniform(0.1, 2)), seed=int(rng.integers(1 << 30)))
        tx, rx = ref["device"](ntx, fs), ref["device"](nrx, fs)
        real = ch.realize()
        s = real.sample(tx, rx)
        x = (rng.standard_normal((ntx, T)) + 1j * rng.standard_normal((ntx, T))) / np.sqrt(2)
        y = s.propagate(ref["Signal"].Create(x, fs, 3.5e9)).view(np.ndarray)
        p = fading_params_from_reference_sample(s)
        yo = fo.propagate(p, x)
        assert yo.shape == y.shape
        assert np.abs(yo - y).max() <= 1e-13 * max(1.0, np.abs(y).max())
        # normals -> parameters (fading.py:468-515)
        p2 = fo.params_from_normals(static_normals_of(real), delays=ch.delays, powers=ch.power_profile,
                                    rice_factors=ch.rice_factors, num_sinusoids=N, doppler=ch.doppler_frequency,
                                    los_doppler=None, gain=ch.gain, fs=fs, num_rx=nrx, num_tx=ntx)
        assert np.array_equal(p2.nlos_angle, p.nlos_angle) and np.array_equal(p2.spatial, p.spatial)
        assert fo.num_scalars(L, N) == static_normals_of(real).size


def test_golden_file_is_current(ref):
    """The committed golden vectors equal what the reference produces today."""
    import os

    from oracle.golden_cases import FADING_CASES, golden_signal

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "fading_golden.npz"))
    RC = ref["RC"]
    for ci, (name, build, ntx, nrx, fs, T, ptx, prx) in enumerate(FADING_CASES[:4]):
        ch = build(RC)
        tx, rx = ref["device"](ntx, fs, ptx), ref["device"](nrx, fs, prx)
        ch.realize()
        s = ch.realize().sample(tx, rx)
        y = s.propagate(ref["Signal"].Create(golden_signal(ci, ntx, T), fs, 3.5e9)).view(np.ndarray)
        assert np.array_equal(np.asarray(y), g[f"{name}/y"])
