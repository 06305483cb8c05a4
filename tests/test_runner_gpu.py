"""GPU: the batched drop runner with the real device call -- many drops of an unmodified scenario in flight, ONE launch
per round, artifacts identical to the reference's serial numpy schedule in the float64 mode (bit-exact BER)."""
import numpy as np
import pytest

from oracle.refload import load_reference, reference_available
from tests.test_runner_cpu import _serial_reference, _simulation

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not reference_available(), reason="no reference install (baseline/_ref)")]


@pytest.mark.parametrize("workers", [0, 2])
def test_batched_lanes_bit_exact_and_one_launch_per_round(workers):
    import hermespy_b200.dropin as dropin
    from hermespy_b200 import _lib
    from hermespy_b200.runner import LaneSet, propagate_requests

    simulation, grid, evaluators = _simulation()
    scenario = simulation.scenario
    rounds = [[(0,), (1,), (2,), (2,), (1,), (0,)], [(2,), (2,), (1,), (0,), (0,), (1,)]]
    calls = []

    def propagate(requests):
        before = sum(_lib.launch_counts().values())
        out = propagate_requests(requests, precision="f64")
        calls.append((len(requests), sum(_lib.launch_counts().values()) - before))
        return out

    dropin.enable(precision="f64")
    try:
        lanes = LaneSet(scenario, grid, evaluators, 6, workers, base_seed=7, first_lane_is_original=False)
        try:
            got = [lanes.run_round(s, propagate) for s in rounds]
        finally:
            lanes.close()
    finally:
        dropin.disable()
    # 6 drops x 2 fading links per round: ONE batched device call = the float64 coefficient kernel + the float64 propagate kernel
    assert calls == [(12, 2), (12, 2)]
    for k in range(6):
        want = _serial_reference(scenario, grid, evaluators, k, 7, [r[k] for r in rounds])  # numpy channel, serial
        have = [[float(a.to_scalar()) for a in g[k]] for g in got]
        assert have == want, (k, have, want)


def test_simulation_run_batched_on_the_gpu():
    load_reference()
    import hermespy_b200.dropin as dropin
    from hermespy_b200 import _lib, runner

    simulation, _, _ = _simulation(num_samples=16)
    runner.stats.update(rounds=0, drops=0, links=0, max_links_per_round=0)
    before = _lib.launch_counts()
    dropin.enable(precision="f32", batch_drops=12, workers=2)
    try:
        result = simulation.run()
    finally:
        dropin.disable()
    after = _lib.launch_counts()
    ber = np.asarray(result.evaluation_results[0].to_array(), dtype=float).ravel()
    assert ber.shape == (3,) and np.all((ber >= 0) & (ber <= 0.5 + 1e-9)) and ber[0] < ber[2]
    # 12 lanes = two alternating groups of 6: 8 rounds of 6 drops x 2 links
    assert runner.stats["drops"] == 48 and runner.stats["links"] == 96 and runner.stats["rounds"] == 8
    assert runner.stats["max_links_per_round"] == 12
    launches = sum(after.values()) - sum(before.values())
    assert 8 <= launches <= 20, launches  # coefficient + propagate kernel per round (+ the warm-up drop) -- not per drop
