"""Host logic of the batched drop runner (hermespy_b200/runner.py), no GPU: lanes, the restated link loop, request
gathering / scattering, helper processes, and ``Simulation.run()`` routed through it -- all against the reference's own
serial schedule with identical seeds (artifacts must be identical bit for bit)."""
import numpy as np
import pytest

from oracle.refload import load_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not available")


def _simulation(num_samples=6, seed=42, snrs=(20, 10, 0)):
    """_examples/getting_started/simulation.py, minus the plots, with every random root pinned."""
    load_reference()
    from hermespy.channel import TDL, TDLType
    from hermespy.core import ConsoleMode, dB
    from hermespy.modem import (BitErrorEvaluator, RootRaisedCosineWaveform, SimplexLink,
                                SingleCarrierLeastSquaresChannelEstimation, SingleCarrierZeroForcingChannelEqualization)
    from hermespy.simulation import SNR, Simulation

    simulation = Simulation(console_mode=ConsoleMode.SILENT, num_samples=num_samples, seed=seed)
    tx = simulation.new_device(oversampling_factor=4)
    rx = simulation.new_device(oversampling_factor=4)
    rx.noise_level = SNR(dB(20), tx)
    simulation.set_channel(tx, rx, TDL(TDLType.A, doppler_frequency=100.0))
    link = SimplexLink(seed=seed + 1)
    tx.transmitters.add(link)
    rx.receivers.add(link)
    link.waveform = RootRaisedCosineWaveform(num_preamble_symbols=10, num_data_symbols=100, roll_off=0.9)
    link.waveform.channel_estimation = SingleCarrierLeastSquaresChannelEstimation()
    link.waveform.channel_equalization = SingleCarrierZeroForcingChannelEqualization()
    dim = simulation.new_dimension("noise_level", dB(*snrs), rx)
    ber = BitErrorEvaluator(link, link)
    simulation.add_evaluator(ber)
    return simulation, [dim], [ber]


def _serial_reference(scenario, grid, evaluators, lane_index, base_seed, sections):
    """What the reference's actor does with this lane's sections, one after the other (actors.py:367-441)."""
    from hermespy.simulation.simulation import SimulationRunner

    from hermespy_b200.runner import Lane

    lane = Lane.clone_of(scenario, grid, evaluators, lane_index, base_seed)
    runner = SimulationRunner(lane.scenario)
    out = []
    for s in sections:
        lane.configure(s)
        for stage in (runner.realize_channels, runner.sample_states, runner.transmit_operators, runner.generate_outputs,
                      runner.propagate, runner.process_inputs, runner.receive_operators):
            stage()
        out.append([float(e.evaluate().artifact().to_scalar()) for e in lane.evaluators])
    return out


def _oracle_propagate(requests):
    """Device stand-in for the CPU tests: the numpy restatement of fading.py:293-406 on the request's parameter block."""
    out = []
    for kind, blk, x, zero in requests:
        assert kind == "fading" and not zero
        T, D = x.shape[1], blk["max_delay"]
        n = np.arange(T)
        K = blk["omega"].shape[1]
        amp = blk["amp"][:, [0] + [1] * (K - 1), None]
        h = (amp * np.exp(1j * (blk["omega"][:, :, None] * n + blk["phi"][:, :, None]))).sum(1)
        z = np.zeros((x.shape[0], T + D), complex)
        for l, d in enumerate(blk["tap_delay"]):
            z[:, d: d + T] += x * h[l]
        out.append(blk["spatial"] @ z)
    return out


@pytest.mark.parametrize("workers", [0, 2])
def test_lanes_reproduce_the_serial_reference_schedule(workers):
    from hermespy_b200.runner import LaneSet

    simulation, grid, evaluators = _simulation()
    scenario = simulation.scenario
    rounds = [[(0,), (1,), (2,)], [(1,), (1,), (0,)], [(2,), (0,)]]
    lanes = LaneSet(scenario, grid, evaluators, 3, workers, base_seed=99, first_lane_is_original=False)
    try:
        got = [lanes.run_round(sections, lambda r: pytest.fail("no device requests expected") if r else [])
               for sections in rounds]
    finally:
        lanes.close()
    for k in range(3):
        mine = [r[k] for r in rounds if k < len(r)]
        want = _serial_reference(scenario, grid, evaluators, k, 99, mine)
        have = [[float(a.to_scalar()) for a in g[k]] for g in got if k < len(g)]
        assert have == want, (k, have, want)
    # lanes own independent random roots: three drops of one grid point differ
    assert len({tuple(_serial_reference(scenario, grid, evaluators, k, 99, [(2,)])[0]) for k in range(3)}) > 1


@pytest.mark.parametrize("workers", [0, 2])
def test_deferred_links_scatter_back_correctly(workers):
    """With the reference classes patched the fading links leave the lanes as (parameter block, block) requests and come
    back as propagated blocks; a numpy stand-in for the device must reproduce the unpatched run to rounding."""
    import hermespy_b200.dropin as dropin
    from hermespy_b200.runner import LaneSet

    simulation, grid, evaluators = _simulation()
    scenario = simulation.scenario
    rounds = [[(0,), (1,), (2,), (2,)], [(2,), (2,), (1,), (0,)]]
    seen = []

    def propagate(requests):
        seen.append(len(requests))
        return _oracle_propagate(requests)

    dropin.patch_reference()  # host-side patch only: makes the sample types "device served" (no launch happens here)
    try:
        lanes = LaneSet(scenario, grid, evaluators, 4, workers, base_seed=7, first_lane_is_original=False,
                        propagate=_oracle_propagate)  # (helpers: one warm-up drop in this process before the fork)
        try:
            got = [lanes.run_round(s, propagate) for s in rounds]
        finally:
            lanes.close()
    finally:
        dropin.disable()
    assert seen == [8, 8]  # 4 lanes x 2 fading links (both directions), ONE gathered call per round
    for k in range(4):
        want = _serial_reference(scenario, grid, evaluators, k, 7, [r[k] for r in rounds])
        have = [[float(a.to_scalar()) for a in g[k]] for g in got]
        assert have == want, (k, have, want)  # BER artifacts are decisions: identical although y differs in rounding


def test_simulation_run_routed_through_the_runner():
    load_reference()
    import ray

    from hermespy_b200 import config, runner
    from hermespy_b200.shims import ray as shim

    if getattr(ray, "__version__", "") != shim.__version__:
        pytest.skip("a real ray is installed")
    simulation, _, _ = _simulation(num_samples=7)
    config.batch_drops, config.workers = 5, 0
    runner.patch_actor()
    runner.stats.update(rounds=0, drops=0, links=0)
    try:
        result = simulation.run()
    finally:
        runner.unpatch_actor()
        config.batch_drops = 0
    ber = np.asarray(result.evaluation_results[0].to_array(), dtype=float).ravel()
    assert ber.shape == (3,) and np.all((ber >= 0) & (ber <= 0.5 + 1e-9)) and ber[0] < ber[2]
    assert runner.stats["drops"] == 21 and runner.stats["rounds"] == 5  # 3 grid points x 7 samples, 5 drops per round
    from hermespy.simulation.simulation import SimulationActor

    assert "run" not in SimulationActor.__dict__  # restored


def test_device_calls_from_helper_processes_are_served_by_the_gpu_owner():
    """``state()`` for ideal channel estimation is called deep inside a lane's receive stage.  A forked helper must not
    touch CUDA: the call travels to the process that owns the GPU (here a numpy stand-in records where it ran)."""
    import os

    load_reference()
    import hermespy_b200.dropin as dropin
    from hermespy_b200.runner import LaneSet
    from tests.test_dropin_gpu import _ofdm_2x1_alamouti_tdl_b  # 2x1 Alamouti OFDM, OFDMIdealChannelEstimation -> sample.state()

    scenario, tx, rx, ber = _ofdm_2x1_alamouti_tdl_b(42)
    grid, evaluators = [], [ber]
    served = []

    def fake_state(b, keep, num_samples):
        served.append(os.getpid())
        n = np.arange(num_samples)
        K = b["omega"].shape[1]
        amp = b["amp"][:, [0] + [1] * (K - 1), None]
        h = (amp * np.exp(1j * (b["omega"][:, :, None] * n + b["phi"][:, :, None]))).sum(1)[keep]
        gd = np.unique(b["tap_delay"][keep])
        return np.stack([h[b["tap_delay"][keep] == d].sum(0) for d in gd]), gd

    rounds = [[(), (), ()], [(), (), ()]]  # empty grid: every section is the empty coordinate tuple
    dropin.patch_reference()
    real = dropin.DEVICE_CALLS["fading_state"]
    dropin.DEVICE_CALLS["fading_state"] = fake_state
    try:
        lanes = LaneSet(scenario, grid, evaluators, 3, 2, base_seed=3, first_lane_is_original=False, propagate=_oracle_propagate)
        try:
            got = [lanes.run_round(s) for s in rounds]
        finally:
            lanes.close()
    finally:
        dropin.DEVICE_CALLS["fading_state"] = real
        dropin.disable()
    assert len(served) >= 1 + 6 and set(served) == {os.getpid()}  # warm-up + one per drop, all in THIS process
    for k in range(3):
        want = _serial_reference(scenario, grid, evaluators, k, 3, [r[k] for r in rounds])
        assert [[float(a.to_scalar()) for a in g[k]] for g in got] == want


def test_device_call_results_travel_through_the_helpers_shared_memory_windows():
    """Large device-call results (ideal-CSI ``state()`` arrays) are written by the GPU owner straight into the calling
    helper's shared-memory window; only a descriptor crosses the pipe.  Same artifacts as through the pipe."""
    load_reference()
    import hermespy_b200.dropin as dropin
    from hermespy_b200 import runner
    from hermespy_b200.runner import LaneSet
    from tests.test_dropin_gpu import _ofdm_2x1_alamouti_tdl_b

    scenario, tx, rx, ber = _ofdm_2x1_alamouti_tdl_b(42)
    through_window = []

    def fake_state(b, keep, num_samples, out_alloc=None):
        n = np.arange(num_samples)
        K = b["omega"].shape[1]
        amp = b["amp"][:, [0] + [1] * (K - 1), None]
        h = (amp * np.exp(1j * (b["omega"][:, :, None] * n + b["phi"][:, :, None]))).sum(1)[keep]
        gd = np.unique(b["tap_delay"][keep])
        res = np.stack([h[b["tap_delay"][keep] == d].sum(0) for d in gd])
        dst = None if out_alloc is None else out_alloc(res.shape, res.dtype)
        through_window.append(dst is not None)
        if dst is None:
            return res, gd
        dst[...] = res
        return dst, gd

    def run(window_bytes):
        old = runner.RPC_SHM_BYTES
        runner.RPC_SHM_BYTES = window_bytes
        dropin.patch_reference()
        real = dropin.DEVICE_CALLS["fading_state"]
        dropin.DEVICE_CALLS["fading_state"] = fake_state
        try:
            lanes = LaneSet(scenario, [], [ber], 3, 2, base_seed=3, first_lane_is_original=False, propagate=_oracle_propagate)
            try:
                return [[float(a.to_scalar()) for a in arts] for arts in lanes.run_round([(), (), ()])]
            finally:
                lanes.close()
        finally:
            dropin.DEVICE_CALLS["fading_state"] = real
            dropin.disable()
            runner.RPC_SHM_BYTES = old

    del through_window[:]
    with_window = run(64 << 20)
    assert through_window.count(True) >= 3  # one per helper-lane drop (the warm-up drop runs in this process: no window)
    del through_window[:]
    without = run(0)
    assert not any(through_window)
    del through_window[:]
    tiny = run(4096)  # a window too small for the array: falls back to the pipe, same result
    assert not any(through_window)
    assert with_window == without == tiny


def test_pipelined_stream_over_helper_processes():
    """``run_stream``: two lane groups alternate so that helpers always have a group's stages to run; every section is
    served exactly once and every lane's drops equal the serial reference schedule of that lane."""
    from hermespy_b200.runner import LaneSet

    simulation, grid, evaluators = _simulation()
    scenario = simulation.scenario
    todo = [(i % 3,) for i in range(13)]
    lanes = LaneSet(scenario, grid, evaluators, 4, 2, base_seed=21, first_lane_is_original=False)
    calls = []
    try:
        got = list(lanes.run_stream(iter(todo), lambda r: calls.append(len(r)) or [], groups=2))
    finally:
        lanes.close()
    assert sorted(s for s, _ in got) == sorted(todo) and len(calls) == 7  # 13 sections in groups of 2
    # the deterministic hand-out: groups take turns, two sections at a time
    per_lane = {k: [] for k in range(4)}
    for turn, o in enumerate(range(0, 13, 2)):
        g = turn % 2
        for i, sec in enumerate(todo[o: o + 2]):
            per_lane[2 * g + i].append(sec)
    have = {k: [] for k in range(4)}
    order = {k: iter(v) for k, v in per_lane.items()}
    by_section = {}
    for sec, art in got:
        by_section.setdefault(sec, []).append([float(a.to_scalar()) for a in art])
    want_all = []
    for k in range(4):
        want_all.extend(zip(per_lane[k], _serial_reference(scenario, grid, evaluators, k, 21, per_lane[k])))
    assert sorted((s, tuple(a)) for s, a in want_all) == sorted((s, tuple(a)) for s, v in by_section.items() for a in v)


def test_large_payloads_do_not_deadlock_the_pipeline():
    """Requests and results far larger than a pipe buffer, two lane groups in flight: the GPU owner keeps reading replies
    while its own payloads drain (writer threads), so helper and owner never block on each other's send."""
    import hermespy_b200.dropin as dropin
    from hermespy_b200.runner import LaneSet
    from tests.test_dropin_gpu import _ofdm_2x1_alamouti_tdl_b

    load_reference()
    scenario, tx, rx, ber = _ofdm_2x1_alamouti_tdl_b(42)  # 2 x 816-sample streams per link: ~26 KB per request, x 2 links x 4 lanes
    served = {"n": 0}

    def propagate(requests):
        served["n"] += len(requests)
        return [np.concatenate([r, r, r, r], axis=1) for r in _oracle_propagate(requests)] if False else _oracle_propagate(requests)

    def fake_state(b, keep, num_samples):
        n = np.arange(num_samples)
        K = b["omega"].shape[1]
        amp = b["amp"][:, [0] + [1] * (K - 1), None]
        h = (amp * np.exp(1j * (b["omega"][:, :, None] * n + b["phi"][:, :, None]))).sum(1)[keep]
        gd = np.unique(b["tap_delay"][keep])
        return np.stack([h[b["tap_delay"][keep] == d].sum(0) for d in gd]), gd

    dropin.patch_reference()
    real = dropin.DEVICE_CALLS["fading_state"]
    dropin.DEVICE_CALLS["fading_state"] = fake_state
    try:
        lanes = LaneSet(scenario, [], [ber], 8, 2, base_seed=5, first_lane_is_original=False, propagate=_oracle_propagate)
        try:
            got = list(lanes.run_stream(iter([()] * 24), propagate, groups=2))
        finally:
            lanes.close()
    finally:
        dropin.DEVICE_CALLS["fading_state"] = real
        dropin.disable()
    assert len(got) == 24 and served["n"] == 48


def test_a_called_off_stream_closes_at_once_and_leaves_no_windows_behind():
    """The reference's campaign loop collects what it needs and kills its actors while rounds are in flight
    (monte_carlo.py:404-470).  Helpers and writer threads are then blocked on each other's pipes: ``close(abort=True)``
    must not wait for them politely (it did: 5 s per thread and process, 80 s at interpreter exit with 16 helpers), and the
    helpers' shared-memory windows must be gone afterwards."""
    import os
    import time

    import hermespy_b200.dropin as dropin
    from hermespy_b200.runner import LaneSet
    from tests.test_dropin_gpu import _ofdm_2x1_alamouti_tdl_b

    load_reference()
    scenario, tx, rx, ber = _ofdm_2x1_alamouti_tdl_b(42)

    def fake_state(b, keep, num_samples, out_alloc=None):
        gd = np.unique(b["tap_delay"][keep])
        return np.ones((gd.size, num_samples), dtype=np.complex128), gd

    def sections():  # the engine goes away after ten sections
        for _ in range(10):
            yield ()
        raise RuntimeError("cannot schedule new futures after shutdown")

    shm_before = set(os.listdir("/dev/shm")) if os.path.isdir("/dev/shm") else set()
    dropin.patch_reference()
    real = dropin.DEVICE_CALLS["fading_state"]
    dropin.DEVICE_CALLS["fading_state"] = fake_state
    try:
        lanes = LaneSet(scenario, [], [ber], 8, 4, base_seed=5, first_lane_is_original=False, propagate=_oracle_propagate)
        procs = [p for p, _ in lanes.procs]
        assert len(lanes.windows) == 4 and all(w is not None for w in lanes.windows)
        got = []
        with pytest.raises(RuntimeError, match="after shutdown"):
            for item in lanes.run_stream(sections(), groups=2):
                got.append(item)
        t0 = time.perf_counter()
        lanes.close(abort=True)
        dt = time.perf_counter() - t0
    finally:
        dropin.DEVICE_CALLS["fading_state"] = real
        dropin.disable()
    assert dt < 4.0, f"close(abort=True) took {dt:.1f} s"
    assert not any(p.is_alive() for p in procs)
    if os.path.isdir("/dev/shm"):
        assert set(os.listdir("/dev/shm")) <= shm_before


def test_window_size_follows_what_dev_shm_can_back(monkeypatch):
    """Touching more shared memory than the tmpfs holds is a SIGBUS: the windows shrink to half of the free space over all
    helpers, and disappear below 1 MB each (container defaults are as small as 64 MB)."""
    import os
    from types import SimpleNamespace

    from hermespy_b200 import runner

    monkeypatch.setattr(runner, "RPC_SHM_BYTES", 64 << 20)
    monkeypatch.setattr(os, "statvfs", lambda path: SimpleNamespace(f_bavail=1 << 20, f_frsize=4096))  # 4 GB free
    assert runner._window_bytes(16) == 64 << 20
    monkeypatch.setattr(os, "statvfs", lambda path: SimpleNamespace(f_bavail=16384, f_frsize=4096))    # 64 MB free
    assert runner._window_bytes(16) == (64 << 20) // 32
    assert runner._window_bytes(64) == 0  # 512 KB each: not worth a window
    monkeypatch.setattr(runner, "RPC_SHM_BYTES", 0)
    assert runner._window_bytes(16) == 0

    def gone(path):
        raise OSError("no /dev/shm")

    monkeypatch.setattr(runner, "RPC_SHM_BYTES", 64 << 20)
    monkeypatch.setattr(os, "statvfs", gone)
    assert runner._window_bytes(16) == 0
