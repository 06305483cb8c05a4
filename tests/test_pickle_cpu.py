"""Host logic (no GPU): channels, realizations and samples of the mirror package survive pickling -- what the reference
ships to its Ray actors (core/pymonte/monte_carlo.py:363-365) -- and the drop-in's replacement methods are module-level."""
import pickle

import numpy as np
import pytest

import hermespy_b200.channel as MC
from hermespy_b200 import dropin
from hermespy_b200.core import SimulatedDevice, SimulatedIdealAntenna, SimulatedUniformArray, Transformation


def _dev(n):
    return SimulatedDevice(bandwidth=30.72e6, carrier_frequency=3.5e9, antennas=SimulatedUniformArray(SimulatedIdealAntenna, 0.04, (n, 1, 1)),
                           pose=Transformation.From_Translation(np.array([10.0 * n, 0.0, 0.0])))


def _equal_blocks(a, b):
    if isinstance(a, dict):
        return all(np.array_equal(np.asarray(a[k]), np.asarray(b[k])) for k in a)
    return all(np.array_equal(np.asarray(getattr(a, k)), np.asarray(getattr(b, k)))
               for k in ("term_delay", "angles", "jones", "amplitude", "tx_pose", "rx_pose", "rel_velocity"))


@pytest.mark.parametrize("build", [
    lambda: MC.TDL(MC.TDLType.B, rms_delay=300e-9, doppler_frequency=100, seed=42,
                   antenna_correlation=MC.StandardAntennaCorrelation(MC.CorrelationType.MEDIUM)),
    lambda: MC.Cost259(MC.Cost259Type.URBAN, doppler_frequency=50, seed=5),
    lambda: MC.CDL(MC.CDLType.C, 300e-9, seed=42),
], ids=["tdl", "cost259", "cdl"])
def test_channel_realization_sample_round_trip(build):
    ch = build()
    tx, rx = _dev(2), _dev(4)
    clone = pickle.loads(pickle.dumps(ch))  # the generator state travels: both draw the same next realization
    real = ch.realize()
    sample = pickle.loads(pickle.dumps(real)).sample(tx, rx)
    assert _equal_blocks(real.sample(tx, rx).kernel_block(), sample.kernel_block())
    assert _equal_blocks(pickle.loads(pickle.dumps(sample)).kernel_block(), sample.kernel_block())
    assert _equal_blocks(clone.realize().sample(tx, rx).kernel_block(), sample.kernel_block())


def test_dropin_replacements_are_picklable_functions():
    for fn in (dropin._fading_propagate, dropin._fading_state, dropin._cdl_propagate, dropin._cdl_state):
        assert pickle.loads(pickle.dumps(fn)) is fn
