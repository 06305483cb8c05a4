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


def test_cdl_oracle_matches_live_reference(ref):
    from hermespy.core import Transformation
    from hermespy.simulation import SimulatedDevice, SimulatedIdealAntenna, SimulatedUniformArray

    from oracle import cdl_oracle as co
    from oracle.golden_cases import CDL_FC, CDL_FS, CDL_SPACING
    from oracle.ref_extract import cdl_params_from_reference_sample

    RC = ref["RC"]
    rng = np.random.default_rng(5)

    def dev(dims, rpy, pos, vel):
        return SimulatedDevice(bandwidth=CDL_FS, oversampling_factor=1, carrier_frequency=CDL_FC,
                               antennas=SimulatedUniformArray(SimulatedIdealAntenna, CDL_SPACING, dims),
                               pose=Transformation.From_RPY(np.array(rpy, float), np.array(pos, float)),
                               velocity=np.array(vel, float))

    for t, K in ((RC.CDLType.B, 0.0), (RC.CDLType.D, 7.0)):
        tx = dev((2, 2, 1), rng.uniform(-1, 1, 3), (0, 0, 20.0), (0, 0, 0))
        rx = dev((2, 1, 1), rng.uniform(-1, 1, 3), (60.0, -30.0, 1.5), rng.uniform(-20, 20, 3))
        s = RC.CDL(t, 200e-9, rayleigh_factor=K, seed=int(rng.integers(1 << 30))).realize().sample(tx, rx)
        x = (rng.standard_normal((4, 90)) + 1j * rng.standard_normal((4, 90))) / np.sqrt(2)
        y = s.propagate(ref["Signal"].Create(x, CDL_FS, CDL_FC)).view(np.ndarray)
        yo = co.propagate(cdl_params_from_reference_sample(s), x)
        assert np.linalg.norm(yo - y) <= 5e-12 * np.linalg.norm(y)


def test_cdl_oracle_matches_live_reference_stochastic_scenarios(ref):
    """The 3GPP scenario models (UMa / UMi / RMa / InH / InF) draw their samples differently but hand out the same
    ``ClusterDelayLineSample``: the oracle (and the kernels behind the drop-in) must follow any cluster count, LOS state
    and delay structure they produce."""
    from hermespy.core import Transformation
    from hermespy.simulation import SimulatedDevice, SimulatedIdealAntenna, SimulatedUniformArray

    from hermespy_b200 import dropin
    from oracle import cdl_oracle as co
    from oracle.golden_cases import CDL_FC, CDL_FS, CDL_SPACING
    from oracle.ref_extract import cdl_params_from_reference_sample

    RC = ref["RC"]
    rng = np.random.default_rng(9)

    def dev(dims, pos, vel):
        return SimulatedDevice(bandwidth=CDL_FS, oversampling_factor=1, carrier_frequency=CDL_FC,
                               antennas=SimulatedUniformArray(SimulatedIdealAntenna, CDL_SPACING, dims),
                               pose=Transformation.From_Translation(np.array(pos, float)), velocity=np.array(vel, float))

    builders = [
        lambda: RC.UrbanMacrocells(expected_state=RC.O2IState.LOS, seed=1),
        lambda: RC.UrbanMacrocells(expected_state=RC.O2IState.O2I, seed=3),
        lambda: RC.UrbanMicrocells(expected_state=RC.O2IState.NLOS, seed=4),
        lambda: RC.RuralMacrocells(expected_state=RC.O2IState.LOS, seed=5),
        lambda: RC.IndoorOffice(expected_state=RC.LOSState.LOS, seed=6),
        lambda: RC.IndoorFactory(2000.0, 1500.0, RC.FactoryType.DH, expected_state=RC.LOSState.NLOS, seed=7),
    ]
    counts = set()
    blocks = []
    for build in builders:
        tx, rx = dev((2, 2, 1), (0.0, 0.0, 25.0), (0, 0, 0)), dev((2, 1, 1), (120.0, 40.0, 1.5), (3.0, -1.0, 0.0))
        s = build().realize().sample(tx, rx)
        x = (rng.standard_normal((4, 120)) + 1j * rng.standard_normal((4, 120))) / np.sqrt(2)
        y = s.propagate(ref["Signal"].Create(x, CDL_FS, CDL_FC)).view(np.ndarray)
        yo = co.propagate(cdl_params_from_reference_sample(s), x)
        assert yo.shape == y.shape and np.linalg.norm(yo - y) <= 5e-12 * np.linalg.norm(y)
        blk = dropin.cdl_block_from_reference(s)  # what the patched _propagate hands to hb_cdl_propagate_host
        assert blk.batch == 1 and blk.num_tx == 4 and blk.num_rx == 2 and blk.max_delay == y.shape[1] - 120
        assert blk.term_delay.size == blk.amplitude.shape[1] == blk.angles.shape[1] and blk.term_delay.max() <= blk.max_delay
        counts.add(int(s.num_clusters))
        blocks.append(blk)
    assert len(counts) >= 4  # the scenarios really produce different cluster counts
    # ... and travel as ONE heterogeneous batch: rays padded with zero amplitude, per-link delay tables, largest delay spread
    from hermespy_b200.kernels import CdlBlock, cdl_plan

    stacked = CdlBlock.stack(blocks)
    rn = max(b.term_delay.size for b in blocks)
    assert stacked.batch == len(blocks) and stacked.link_term_delay.shape == (len(blocks), rn)
    assert stacked.max_delay == max(b.max_delay for b in blocks) and stacked.line_of_sight == any(b.line_of_sight for b in blocks)
    for k, b in enumerate(blocks):
        n = b.term_delay.size
        assert np.array_equal(stacked.link_term_delay[k, :n], b.term_delay) and np.array_equal(stacked.amplitude[k, :n], b.amplitude[0])
        assert not stacked.amplitude[k, n:].any() and np.isfinite(stacked.jones[k]).all()
        assert stacked.link_los_amplitude[k] == (b.los_amplitude if b.line_of_sight else 0.0)
        assert stacked.link_los_delay[k] == b.los_delay and stacked.link_max_delay[k] == b.max_delay
    plan = cdl_plan(stacked, 2304, precision="f32")  # the planner sizes the launch by the largest per-link group count
    groups = [np.unique(np.append(b.term_delay, b.los_delay) if b.line_of_sight else b.term_delay).size for b in blocks]
    assert plan["num_groups"] == max(groups)


def test_dropin_extraction_equals_mirror_blocks(ref):
    """The drop-in adapter reads reference samples into the same kernel blocks the mirror classes build."""
    import hermespy_b200.channel as MC
    from hermespy.core import Transformation
    from hermespy.simulation import SimulatedDevice, SimulatedIdealAntenna, SimulatedUniformArray
    from hermespy_b200 import dropin
    from oracle.golden_cases import CDL_CASES, CDL_FC, CDL_FS, CDL_SPACING, FADING_CASES
    from tests.test_cdl_golden import mirror_cdl_sample
    from tests.test_oracle_golden import mirror_sample

    RC = ref["RC"]
    for case in FADING_CASES[:6]:
        name, build, ntx, nrx, fs, T, ptx, prx = case
        ch = build(RC)
        tx, rx = ref["device"](ntx, fs, ptx), ref["device"](nrx, fs, prx)
        ch.realize()
        rs = ch.realize().sample(tx, rx)
        got = dropin.fading_block_from_reference(rs)
        want = mirror_sample(case)[1].kernel_block()
        for k in ("tap_delay", "omega", "phi", "amp", "spatial"):
            assert np.array_equal(got[k], want[k]), (name, k)
        assert got["max_delay"] == want["max_delay"] and got["omega_max"] == want["omega_max"]

    def dev(spec):
        dims, rpy, pos, vel = spec
        return SimulatedDevice(bandwidth=CDL_FS, oversampling_factor=1, carrier_frequency=CDL_FC,
                               antennas=SimulatedUniformArray(SimulatedIdealAntenna, CDL_SPACING, dims),
                               pose=Transformation.From_RPY(np.array(rpy, float), np.array(pos, float)),
                               velocity=np.array(vel, float))

    for case in CDL_CASES:
        name, build, txs, rxs, T = case
        ch = build(RC)
        ch.realize()
        rs = ch.realize().sample(dev(txs), dev(rxs))
        got = dropin.cdl_block_from_reference(rs)
        want = mirror_cdl_sample(case)[1].kernel_block()
        assert np.array_equal(got.term_delay, want.term_delay) and got.max_delay == want.max_delay
        for k in ("angles", "jones", "amplitude", "rel_velocity"):
            assert np.array_equal(getattr(got, k), getattr(want, k)), (name, k)
        for k in ("tx_pose", "rx_pose", "tx_topology", "rx_topology"):
            assert np.allclose(getattr(got, k), getattr(want, k), rtol=0, atol=1e-12), (name, k)
        assert (got.line_of_sight, got.los_delay, got.los_amplitude) == (want.line_of_sight, want.los_delay, want.los_amplitude)


def test_dropin_patch_and_restore(ref):
    from hermespy.channel.cdl.cluster_delay_lines import ClusterDelayLineSample
    from hermespy.channel.fading.fading import MultipathFadingSample
    from hermespy_b200 import dropin

    orig_f, orig_c, orig_s = MultipathFadingSample._propagate, ClusterDelayLineSample._propagate, MultipathFadingSample.state
    orig_cs = ClusterDelayLineSample.state
    dropin.patch_reference()
    try:
        assert MultipathFadingSample._propagate is dropin._fading_propagate
        assert MultipathFadingSample.state is dropin._fading_state
        assert ClusterDelayLineSample._propagate is dropin._cdl_propagate
        assert ClusterDelayLineSample.state is dropin._cdl_state
    finally:
        dropin.disable()
    assert MultipathFadingSample._propagate is orig_f and ClusterDelayLineSample._propagate is orig_c
    assert MultipathFadingSample.state is orig_s and ClusterDelayLineSample.state is orig_cs


def test_stats_oracle_matches_reference_evaluator(ref):
    """oracle/stats_oracle.py against BitErrorEvaluator.evaluate/.artifact (modem/evaluators.py:231-259)."""
    from hermespy.modem.evaluators import BitErrorEvaluator

    from oracle import stats_oracle as so

    class _Bits:
        def __init__(self, bits):
            self.bits = bits

    rng = np.random.default_rng(7)
    BitErrorEvaluator.__del__ = lambda self: None  # the bare test instance never registered its hooks
    for _ in range(20):
        nt, nr = int(rng.integers(0, 40)), int(rng.integers(1, 40))
        tx, rx = rng.integers(0, 2, nt), rng.integers(0, 2, nr)
        ev = BitErrorEvaluator.__new__(BitErrorEvaluator)
        ev._fetch_dsp_results = lambda tx=tx, rx=rx: (_Bits(tx), _Bits(rx))
        evaluation = ev.evaluate()
        errors, bits, artifact = so.bit_errors(tx, rx)
        assert errors == int(np.sum(evaluation.evaluation)) and bits == len(evaluation.evaluation)
        assert artifact == evaluation.artifact().to_scalar()


def test_dropin_extracts_element_tables_of_non_ideal_arrays(ref):
    """dropin.cdl_block_from_reference on dipole / patch / cross-polarized arrays: element tables, element mode and the
    whole parameter block equal the test-side layout from the oracle parameters; unknown element classes are refused."""
    import hermespy_b200.dropin as dropin
    from hermespy.core import Transformation
    from hermespy.simulation import (SimulatedCustomArray, SimulatedDevice, SimulatedDipole, SimulatedIdealAntenna,
                                     SimulatedLinearAntenna, SimulatedPatchAntenna, SimulatedUniformArray)
    from oracle.ref_extract import cdl_params_from_reference_sample
    from tests.helpers import cdl_block_from_oracle_params

    RC = ref["RC"]
    fs, fc = 30.72e6, 3.5e9

    def dev(arr, pos, rpy=(0, 0, 0)):
        return SimulatedDevice(bandwidth=fs, oversampling_factor=1, carrier_frequency=fc, antennas=arr,
                               pose=Transformation.From_RPY(np.array(rpy, float), np.array(pos, float)))

    xpol = SimulatedCustomArray([SimulatedLinearAntenna(slant=s_, pose=Transformation.From_RPY(
        np.array([0.0, 0.1 * i, 0.0]), np.array([0.0, 0.04 * i, 0.0]))) for i in range(2) for s_ in (0.7, -0.7)])
    cases = [(SimulatedUniformArray(SimulatedDipole, 0.04, (2, 2, 1)), SimulatedUniformArray(SimulatedPatchAntenna, 0.04, (2, 1, 1)), 1),
             (xpol, SimulatedUniformArray(SimulatedIdealAntenna, 0.04, (2, 1, 1)), 2),
             (SimulatedUniformArray(SimulatedIdealAntenna, 0.04, (2, 1, 1)), SimulatedUniformArray(SimulatedIdealAntenna, 0.04, (1, 1, 1)), 0)]
    for txa, rxa, mode in cases:
        s = RC.CDL(RC.CDLType.E, 1e-7, seed=3).realize().sample(dev(txa, (0, 0, 10.0), (0.1, 0.2, 0.3)), dev(rxa, (40.0, 5.0, 1.5)))
        got = dropin.cdl_block_from_reference(s)
        want = cdl_block_from_oracle_params(cdl_params_from_reference_sample(s))
        assert got.element_mode == want.element_mode == mode
        for f in ("term_delay", "angles", "jones", "amplitude", "tx_pose", "rx_pose", "rel_velocity", "tx_topology",
                  "rx_topology", "tx_elements", "rx_elements"):
            a, b = getattr(got, f), getattr(want, f)
            assert (a is None and b is None) or np.array_equal(a, b), f
        assert (got.max_delay, got.line_of_sight, got.los_delay, got.los_amplitude) == \
               (want.max_delay, want.line_of_sight, want.los_delay, want.los_amplitude)

    class Cardioid(SimulatedIdealAntenna):
        def local_characteristics(self, azimuth, elevation):
            return np.array([0.5 * (1 + np.cos(azimuth)), 0.0])

        def copy(self):
            return Cardioid(self.mode, self.pose.copy())

    s = RC.CDL(RC.CDLType.A, 1e-7, seed=3).realize().sample(
        dev(SimulatedUniformArray(Cardioid, 0.04, (2, 1, 1)), (0, 0, 10.0)), dev(SimulatedUniformArray(SimulatedIdealAntenna, 0.04, (1, 1, 1)), (40.0, 5.0, 1.5)))
    with pytest.raises(dropin.UnsupportedByKernels):
        dropin.cdl_block_from_reference(s)


def test_delay_radar_and_ideal_channels_pass_through_the_dropin_untouched(ref):
    """SURVEY 8(f)-4 / DESIGN section 7: delay, radar and ideal channels carry no sum-of-sinusoids stage (one static tap: a
    small matmul and a shift that numpy runs at memory speed) -- the drop-in leaves their sample classes alone.  Pinned
    here: with the patch applied their ``_propagate`` / ``state`` are still the reference's own functions and their
    outputs are bit-identical; the batched runner classifies them as not device-served."""
    import hermespy_b200.dropin as dropin
    from hermespy.channel import IdealChannel, RandomDelayChannel, SingleTargetRadarChannel, SpatialDelayChannel
    from hermespy.channel.delay.delay import DelayChannelSample
    from hermespy.channel.ideal import IdealChannelSample
    from hermespy.channel.radar.radar import RadarChannelSample
    from hermespy_b200 import runner

    Signal, device = ref["Signal"], ref["device"]
    fs = 1e8
    classes = (DelayChannelSample, RadarChannelSample, IdealChannelSample)
    before = {c: (c.__dict__.get("_propagate"), c.__dict__.get("state")) for c in classes}
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 200)) + 1j * rng.standard_normal((2, 200))
    sig = Signal.Create(x, fs, 3.5e9)
    tx, rx = device(2, fs, (0, 0, 0)), device(2, fs, (30.0, 4.0, 0))
    channels = [SpatialDelayChannel(seed=1), RandomDelayChannel(3e-8, seed=2), IdealChannel(seed=3),
                SingleTargetRadarChannel(25.0, 1.0, seed=4)]
    samples = [ch.realize().sample(tx, rx if not isinstance(ch, SingleTargetRadarChannel) else tx) for ch in channels]
    want = [np.asarray(s.propagate(sig).view(np.ndarray)) for s in samples]
    dropin.patch_reference()
    try:
        for c in classes:
            assert (c.__dict__.get("_propagate"), c.__dict__.get("state")) == before[c]  # untouched
        got = [np.asarray(s.propagate(sig).view(np.ndarray)) for s in samples]
        assert all(runner._gpu_kind(s) is None for s in samples)
    finally:
        dropin.disable()
    for a, b in zip(got, want):
        assert a.shape == b.shape and np.array_equal(a, b)
