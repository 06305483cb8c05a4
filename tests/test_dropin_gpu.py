"""The UNMODIFIED reference drop loop with the CUDA channel path patched in (north_star acceptance test).

Needs the reference install ``baseline/_ref`` (``tools/install_reference.py``; travels to the GPU box with the
snapshot, git-ignored) -- skipped where it is absent.  Every scenario is run twice from the same seed: once with
the reference's own numpy ``_propagate`` and once with ``hermespy_b200.dropin`` enabled.  Realization, sampling,
modems, noise, synchronization, equalization and ``BitErrorEvaluator`` are reference code in both runs.

* float64 parity mode: received signals agree to <= 1e-12 relative L2 and the per-drop bit-error vectors are
  IDENTICAL (bit-exact BER counts).
* float32 mode: propagated signals within the stated 1e-5 relative L2; bit-error totals are reported and may differ
  by isolated decisions, bounded here.
"""
import numpy as np
import pytest

from oracle.golden_cases import CDL_CASES, CDL_FC, CDL_FS, CDL_SPACING, FADING_CASES, golden_signal
from oracle.refload import load_reference, reference_available
from tests.helpers import rel_l2

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not reference_available(), reason="no reference install (baseline/_ref)")]


@pytest.fixture(scope="module")
def ref():
    load_reference()
    import hermespy_b200.dropin as dropin

    yield dropin
    dropin.disable()


def _launches():
    from hermespy_b200 import _lib

    return sum(_lib.launch_counts().values())


def _assert_cuda_path_ran(before, y_gpu, y_ref, at_least=1):
    """The patched call launched kernels of this library and did not merely replay the reference's numpy result."""
    assert _launches() - before >= at_least, "no CUDA kernel was launched by the patched call"
    assert not np.array_equal(y_gpu, y_ref), "bit-identical to the reference's numpy output: the CUDA path did not run"


# ---- scenarios (reference API only) ---------------------------------------------------------------------------

def _siso_rrc_tdl_a(seed):
    """BASELINE config C1: SISO RRC modem over TDL-A (getting_started/simulation.py)."""
    from hermespy.channel import TDL, TDLType
    from hermespy.modem import (BitErrorEvaluator, RootRaisedCosineWaveform, SimplexLink,
                                SingleCarrierLeastSquaresChannelEstimation, SingleCarrierZeroForcingChannelEqualization)
    from hermespy.simulation import SimulationScenario

    sc = SimulationScenario(seed=seed)
    tx = sc.new_device(oversampling_factor=4)
    rx = sc.new_device(oversampling_factor=4)
    sc.set_channel(tx, rx, TDL(TDLType.A, doppler_frequency=100.0))
    link = SimplexLink(seed=seed + 1)
    tx.transmitters.add(link)
    rx.receivers.add(link)
    link.waveform = RootRaisedCosineWaveform(num_preamble_symbols=10, num_data_symbols=100, roll_off=0.9)
    link.waveform.channel_estimation = SingleCarrierLeastSquaresChannelEstimation()
    link.waveform.channel_equalization = SingleCarrierZeroForcingChannelEqualization()
    return _seeded(sc, tx, rx, seed), tx, rx, BitErrorEvaluator(link, link)


def _seeded(sc, tx, rx, seed):
    """The reference leaves the modem and the devices' noise models (re-parented to an unseeded RF block,
    simulation/rf/block.py:102-106) as independent random roots; pin them so that two runs of the reference itself
    are reproducible."""
    tx.noise_model.seed = seed + 2
    rx.noise_model.seed = seed + 3
    return sc


def _ofdm_link(seed, channel_builder, ntx, nrx, coding=None, carrier=3.5e9, ideal_csi=False):
    from hermespy.core import Transformation
    from hermespy.modem import (BitErrorEvaluator, ElementType, GridElement, GridResource, OFDMWaveform,
                                OrthogonalLeastSquaresChannelEstimation, OrthogonalZeroForcingChannelEqualization,
                                SimplexLink, SymbolSection)
    from hermespy.simulation import SimulatedIdealAntenna, SimulatedUniformArray, SimulationScenario

    sc = SimulationScenario(seed=seed)
    bw = 128 * 240e3  # 30.72 MHz
    lam = 299792458.0 / carrier

    def dev(n, pos, vel):
        return sc.new_device(carrier_frequency=carrier, bandwidth=bw, oversampling_factor=1,
                             pose=Transformation.From_Translation(np.array(pos, float)), velocity=np.array(vel, float),
                             antennas=SimulatedUniformArray(SimulatedIdealAntenna, 0.5 * lam, [n, 1, 1]))

    tx, rx = dev(ntx, (0.0, 0.0, 25.0), (0, 0, 0)), dev(nrx, (100.0, 20.0, 1.5), (10.0, -3.0, 0.0))
    channel = channel_builder()
    sc.set_channel(tx, rx, channel)
    link = SimplexLink(seed=seed + 1)
    tx.transmitters.add(link)
    rx.receivers.add(link)
    res = [GridResource(128 // 5, prefix_ratio=0.0703, elements=[GridElement(ElementType.REFERENCE, 1), GridElement(ElementType.DATA, 4)]),
           GridResource(128 // 5, prefix_ratio=0.0703, elements=[GridElement(ElementType.DATA, 2), GridElement(ElementType.REFERENCE, 1),
                                                                  GridElement(ElementType.DATA, 2)])]
    link.waveform = OFDMWaveform(num_subcarriers=128, dc_suppression=True, grid_resources=res,
                                 grid_structure=[SymbolSection(64, [0, 1], 5)], modulation_order=4)
    if ideal_csi:  # MIMO: the reference's least-squares estimator is SISO only; ideal CSI calls sample.state()
        from hermespy.simulation import OFDMIdealChannelEstimation

        link.waveform.channel_estimation = OFDMIdealChannelEstimation(channel, tx, rx)
    else:
        link.waveform.channel_estimation = OrthogonalLeastSquaresChannelEstimation()
    link.waveform.channel_equalization = OrthogonalZeroForcingChannelEqualization()
    if coding is not None:  # as tests/integration_tests/test_mimo.py:130-143: the space-time decoder consumes the CSI itself
        from hermespy.modem import ChannelEqualization

        link.transmit_symbol_coding[0] = coding()
        link.receive_symbol_coding[0] = coding()
        link.waveform.channel_equalization = ChannelEqualization()
    return _seeded(sc, tx, rx, seed), tx, rx, BitErrorEvaluator(link, link)


def _ofdm_2x1_alamouti_tdl_b(seed):
    from hermespy.channel import TDL, TDLType
    from hermespy.modem import Alamouti

    return _ofdm_link(seed, lambda: TDL(TDLType.B, rms_delay=300e-9, doppler_frequency=100.0), 2, 1, Alamouti, ideal_csi=True)


def _ofdm_siso_cdl_c(seed):
    from hermespy.channel import CDL, CDLType

    return _ofdm_link(seed, lambda: CDL(CDLType.C, 300e-9), 1, 1)


def _run(build, seed, snrs_db, drops):
    """Drop loop over an SNR sweep: per-drop bit-error vectors and received device signals."""
    from hermespy.core import dB
    from hermespy.simulation import SNR

    sc, tx, rx, ber = build(seed)
    errors, received = [], []
    for snr in snrs_db:
        rx.noise_level = SNR(dB(snr), tx)
        for _ in range(drops):
            drop = sc.drop()
            errors.append(np.asarray(ber.evaluate().evaluation).copy())
            received.append(np.asarray(drop.device_receptions[1].impinging_signals[0].view(np.ndarray)).copy())
    return errors, received


# f64 tolerance: 1e-12 for fading; 1e-10 for CDL, whose distance / Doppler phases reach 1e3..1e4 rad (the reference's own
# rounding there is ~1e-12)
@pytest.mark.parametrize("build,snrs,drops,tol64", [(_siso_rrc_tdl_a, (0, 4, 8, 12, 16, 20), 25, 1e-12),
                                                     (_ofdm_2x1_alamouti_tdl_b, (0, 10, 20), 8, 1e-12),
                                                     (_ofdm_siso_cdl_c, (5, 15), 5, 1e-10)],
                         ids=["c1_siso_rrc_tdl_a", "ofdm_2x1_alamouti_tdl_b", "ofdm_siso_cdl_c"])
def test_simulation_drop_loop_bit_exact_ber(ref, build, snrs, drops, tol64):
    from hermespy_b200 import _lib

    ref.disable()
    e_ref, r_ref = _run(build, 42, snrs, drops)
    before = sum(_lib.launch_counts().values())
    ref.enable(precision="f64")
    e_64, r_64 = _run(build, 42, snrs, drops)
    ref.enable(precision="f32")
    e_32, r_32 = _run(build, 42, snrs, drops)
    ref.disable()
    assert sum(_lib.launch_counts().values()) - before >= 2 * len(snrs) * drops  # the CUDA path really ran
    assert len(e_ref) == len(e_64) == len(snrs) * drops
    for a, b in zip(r_ref, r_64):
        assert a.shape == b.shape and rel_l2(b, a) < tol64
    for a, b in zip(e_ref, e_64):
        assert np.array_equal(a, b)  # bit-exact per drop, not only in total
    for a, b in zip(r_ref, r_32):
        assert rel_l2(b, a) < 1e-5
    tot_ref, tot_32 = sum(int(e.sum()) for e in e_ref), sum(int(e.sum()) for e in e_32)
    bits = sum(e.size for e in e_ref)
    print(f"bits {bits}, errors: reference {tot_ref}, f64 {sum(int(e.sum()) for e in e_64)}, f32 {tot_32}")
    assert abs(tot_32 - tot_ref) <= max(3, tot_ref // 200)


@pytest.mark.parametrize("ci", range(len(FADING_CASES)), ids=[c[0] for c in FADING_CASES])
def test_reference_fading_samples_propagate_through_dropin(ref, ci):
    """Every golden fading configuration, built from the REFERENCE classes, propagated by both implementations."""
    import hermespy.channel as RC
    from hermespy.core import Signal, Transformation
    from hermespy.simulation import SimulatedDevice, SimulatedIdealAntenna, SimulatedUniformArray

    name, build, ntx, nrx, fs, T, ptx, prx = FADING_CASES[ci]

    def device(n, pos):
        return SimulatedDevice(bandwidth=fs, oversampling_factor=1, carrier_frequency=3.5e9,
                               antennas=SimulatedUniformArray(SimulatedIdealAntenna, 0.04, (n, 1, 1)),
                               pose=Transformation.From_Translation(np.array(pos, dtype=float)))

    s = build(RC).realize().sample(device(ntx, ptx), device(nrx, prx))
    sig = Signal.Create(golden_signal(ci, ntx, max(T, 2100)), fs, 3.5e9)
    ref.disable()
    y0 = np.asarray(s.propagate(sig).view(np.ndarray))
    before = _launches()
    ref.enable(precision="f64")
    y64 = np.asarray(s.propagate(sig).view(np.ndarray))
    mid = _launches()
    ref.enable(precision="f32")
    y32 = np.asarray(s.propagate(sig).view(np.ndarray))
    ref.disable()
    assert mid - before >= 1
    _assert_cuda_path_ran(mid, y32, y0)
    assert y0.shape == y64.shape == y32.shape
    assert rel_l2(y64, y0) < (1e-10 if "extreme" in name else 1e-12)
    assert rel_l2(y32, y0) < 1e-5
    # channel state (fading.py:345-369): the tap gains come from hb_fading_state, container and layout stay the reference's
    taps = 1 + y0.shape[1] - sig.num_samples
    c0 = np.asarray(s.state(150, taps).dense_state())
    before = _launches()
    ref.enable(precision="f64")
    st = s.state(150, taps)
    ref.disable()
    assert _launches() - before >= 1  # hb_fading_state
    c64 = np.asarray(st.dense_state())
    assert type(st).__name__ == "ChannelStateInformation" and c0.shape == c64.shape
    assert rel_l2(c64, c0) < (1e-10 if "extreme" in name else 1e-12)


@pytest.mark.parametrize("ci", range(len(CDL_CASES)), ids=[c[0] for c in CDL_CASES])
def test_reference_cdl_samples_propagate_through_dropin(ref, ci):
    import hermespy.channel as RC
    from hermespy.core import Signal, Transformation
    from hermespy.simulation import SimulatedDevice, SimulatedIdealAntenna, SimulatedUniformArray

    name, build, txs, rxs, T = CDL_CASES[ci]

    def dev(spec):
        dims, rpy, pos, vel = spec
        return SimulatedDevice(bandwidth=CDL_FS, oversampling_factor=1, carrier_frequency=CDL_FC,
                               antennas=SimulatedUniformArray(SimulatedIdealAntenna, CDL_SPACING, dims),
                               pose=Transformation.From_RPY(np.array(rpy, float), np.array(pos, float)),
                               velocity=np.array(vel, float))

    s = build(RC).realize().sample(dev(txs), dev(rxs))
    sig = Signal.Create(golden_signal(200 + ci, int(np.prod(txs[0])), T), CDL_FS, CDL_FC)
    ref.disable()
    y0 = np.asarray(s.propagate(sig).view(np.ndarray))
    before = _launches()
    ref.enable(precision="f64")
    y64 = np.asarray(s.propagate(sig).view(np.ndarray))
    mid = _launches()
    ref.enable(precision="f32")
    y32 = np.asarray(s.propagate(sig).view(np.ndarray))
    ref.disable()
    assert mid - before >= 2  # ray kernel + per-ray FP64 kernel
    _assert_cuda_path_ran(mid, y32, y0, at_least=3)  # rays, moments, K6
    assert not np.array_equal(y64, y0)
    assert rel_l2(y64, y0) < 1e-10
    assert rel_l2(y32, y0) < 1e-5
    c0 = np.asarray(s.state(40, 1000).dense_state())  # cluster_delay_lines.py:561-592
    before = _launches()
    ref.enable(precision="f64")
    c64 = np.asarray(s.state(40, 1000).dense_state())
    ref.disable()
    assert _launches() - before >= 2
    assert c0.shape == c64.shape and rel_l2(c64, c0) < 1e-10


_STOCHASTIC = {
    "uma_los": lambda C: C.UrbanMacrocells(expected_state=C.O2IState.LOS, seed=1),
    "uma_nlos": lambda C: C.UrbanMacrocells(expected_state=C.O2IState.NLOS, seed=2),
    "uma_o2i": lambda C: C.UrbanMacrocells(expected_state=C.O2IState.O2I, seed=3),
    "umi_nlos": lambda C: C.UrbanMicrocells(expected_state=C.O2IState.NLOS, seed=4),
    "rma_los": lambda C: C.RuralMacrocells(expected_state=C.O2IState.LOS, seed=5),
    "inh_los": lambda C: C.IndoorOffice(expected_state=C.LOSState.LOS, seed=6),
    "inf_nlos": lambda C: C.IndoorFactory(2000.0, 1500.0, C.FactoryType.DH, expected_state=C.LOSState.NLOS, seed=7),
}


@pytest.mark.parametrize("name", list(_STOCHASTIC))
def test_stochastic_cdl_scenarios_through_dropin(ref, name):
    """the 3GPP scenario models (UMa / UMi / RMa / InH / InF, cluster_delay_lines.py:1824-2013) only differ
    from the static CDL tables in how a sample is DRAWN; the sample type -- and therefore the patched ``_propagate`` /
    ``state`` -- is the same.  Cluster counts, delays and LOS state vary per realization (6..25 clusters here)."""
    import hermespy.channel as RC
    from hermespy.core import Signal, Transformation
    from hermespy.simulation import SimulatedDevice, SimulatedIdealAntenna, SimulatedUniformArray

    def dev(dims, pos, vel):
        return SimulatedDevice(bandwidth=CDL_FS, oversampling_factor=1, carrier_frequency=CDL_FC,
                               antennas=SimulatedUniformArray(SimulatedIdealAntenna, CDL_SPACING, dims),
                               pose=Transformation.From_Translation(np.array(pos, float)), velocity=np.array(vel, float))

    ch = _STOCHASTIC[name](RC)
    tx, rx = dev((2, 2, 1), (0.0, 0.0, 25.0), (0, 0, 0)), dev((2, 1, 1), (120.0, 40.0, 1.5), (3.0, -1.0, 0.0))
    for rep in range(2):  # two realizations: different cluster sets / delay structures
        s = ch.realize().sample(tx, rx)
        sig = Signal.Create(golden_signal(400 + rep, 4, 2304), CDL_FS, CDL_FC)
        ref.disable()
        y0 = np.asarray(s.propagate(sig).view(np.ndarray))
        c0 = np.asarray(s.state(32, 1000).dense_state())
        before = _launches()
        ref.enable(precision="f64")
        y64 = np.asarray(s.propagate(sig).view(np.ndarray))
        c64 = np.asarray(s.state(32, 1000).dense_state())
        mid = _launches()
        ref.enable(precision="f32")
        y32 = np.asarray(s.propagate(sig).view(np.ndarray))
        ref.disable()
        assert mid - before >= 4
        _assert_cuda_path_ran(mid, y32, y0, at_least=2)
        assert y0.shape == y64.shape == y32.shape and c0.shape == c64.shape
        assert rel_l2(y64, y0) < 1e-10 and rel_l2(c64, c0) < 1e-10
        assert rel_l2(y32, y0) < 1e-5


def test_stochastic_realizations_share_one_launch_set_in_the_runner(ref):
    """VERDICT r1 design note 14: realizations of a stochastic scenario draw their own cluster delays, cluster counts and
    LOS state.  The batched runner's device call stacks them into ONE heterogeneous batch (per-link delay tables in device
    memory) instead of one launch set per realization; every result equals the reference's numpy ``propagate``."""
    import hermespy.channel as RC
    from hermespy.core import Signal, Transformation
    from hermespy.simulation import SimulatedDevice, SimulatedIdealAntenna, SimulatedUniformArray

    from hermespy_b200 import _lib, dropin, runner

    def dev(dims, pos, vel):
        return SimulatedDevice(bandwidth=CDL_FS, oversampling_factor=1, carrier_frequency=CDL_FC,
                               antennas=SimulatedUniformArray(SimulatedIdealAntenna, CDL_SPACING, dims),
                               pose=Transformation.From_Translation(np.array(pos, float)), velocity=np.array(vel, float))

    tx, rx = dev((2, 2, 1), (0.0, 0.0, 25.0), (0, 0, 0)), dev((2, 1, 1), (120.0, 40.0, 1.5), (3.0, -1.0, 0.0))
    samples, sigs = [], []
    for k, name in enumerate(("uma_los", "uma_nlos", "umi_nlos", "uma_los", "rma_los", "uma_o2i")):
        samples.append(_STOCHASTIC[name](RC).realize().sample(tx, rx))
        sigs.append(Signal.Create(golden_signal(500 + k, 4, 1536), CDL_FS, CDL_FC))
    ref.disable()
    want = [np.asarray(s.propagate(x).view(np.ndarray)) for s, x in zip(samples, sigs)]
    blocks = [dropin.cdl_block_from_reference(s) for s in samples]
    assert len({b.delay_key() for b in blocks}) >= 5  # the realizations have their own delay structures (one seed repeats)
    requests = [("cdl", b, np.ascontiguousarray(np.asarray(x.blocks[0], dtype=np.complex128)), False)
                for b, x in zip(blocks, sigs)]
    for precision, tol in (("f64", 1e-10), ("f32", 1e-5)):
        before = _lib.launch_counts()
        got = runner.propagate_requests(requests, precision=precision)
        after = _lib.launch_counts()
        assert after["cdl_propagate"] - before["cdl_propagate"] == 1 and after["cdl_rays"] - before["cdl_rays"] <= 2
        for y, w in zip(got, want):
            assert y.shape == w.shape and rel_l2(y, w) < tol


def test_unmodified_simulation_run_with_gpu_channel(ref):
    """``Simulation.run()`` itself (the reference's Monte-Carlo engine with its queue manager / actor / collector, driven
    through the in-process ``ray`` stand-in) with the channel on the GPU: every drop launches CUDA kernels."""
    import ray

    from hermespy_b200 import _lib
    from hermespy_b200.shims import ray as shim

    if getattr(ray, "__version__", "") != shim.__version__:
        pytest.skip("a real ray is installed")
    from hermespy.channel import TDL
    from hermespy.core import ConsoleMode, dB
    from hermespy.modem import (BitErrorEvaluator, RootRaisedCosineWaveform, SimplexLink,
                                SingleCarrierLeastSquaresChannelEstimation, SingleCarrierZeroForcingChannelEqualization)
    from hermespy.simulation import SNR, Simulation

    simulation = Simulation(console_mode=ConsoleMode.SILENT, num_samples=5, seed=42)
    tx_device = simulation.new_device(oversampling_factor=4)
    rx_device = simulation.new_device(oversampling_factor=4)
    tx_device.noise_level = SNR(dB(20), tx_device)
    rx_device.noise_level = SNR(dB(20), tx_device)
    simulation.set_channel(tx_device, rx_device, TDL())
    link = SimplexLink()
    tx_device.transmitters.add(link)
    rx_device.receivers.add(link)
    link.waveform = RootRaisedCosineWaveform(num_preamble_symbols=10, num_data_symbols=100, roll_off=0.9)
    link.waveform.channel_estimation = SingleCarrierLeastSquaresChannelEstimation()
    link.waveform.channel_equalization = SingleCarrierZeroForcingChannelEqualization()
    simulation.new_dimension("noise_level", dB(20, 10, 0), rx_device)
    simulation.add_evaluator(BitErrorEvaluator(link, link))
    before = sum(_lib.launch_counts().values())
    ref.enable(precision="f64")
    try:
        result = simulation.run()
    finally:
        ref.disable()
    ber = np.asarray(result.evaluation_results[0].to_array(), dtype=float).ravel()
    assert ber.shape == (3,) and np.all((ber >= 0) & (ber <= 0.5 + 1e-9))
    assert sum(_lib.launch_counts().values()) - before >= 10  # 3 SNR points x 5 drops, one parity-mode kernel each (15 measured)


# ---- non-ideal antenna elements through the drop-in -----------------------------------------------

def _element_arrays():
    from hermespy.core import Transformation
    from hermespy.simulation import (SimulatedCustomArray, SimulatedDipole, SimulatedLinearAntenna, SimulatedPatchAntenna,
                                     SimulatedUniformArray)

    def xpol(n):
        ants = []
        for i in range(n):
            for slant in (np.pi / 4, -np.pi / 4):
                ants.append(SimulatedLinearAntenna(slant=slant, pose=Transformation.From_Translation(
                    np.array([0.0, i * CDL_SPACING, 0.0]))))
        return SimulatedCustomArray(ants)

    return {
        "dipole_uniform": (lambda: SimulatedUniformArray(SimulatedDipole, CDL_SPACING, (2, 2, 1)),
                           lambda: SimulatedUniformArray(SimulatedDipole, CDL_SPACING, (2, 1, 1))),
        "patch_uniform": (lambda: SimulatedUniformArray(SimulatedPatchAntenna, CDL_SPACING, (4, 1, 1)),
                          lambda: SimulatedUniformArray(SimulatedPatchAntenna, CDL_SPACING, (1, 2, 1))),
        "xpol_panels": (lambda: xpol(4), lambda: xpol(1)),
    }


@pytest.mark.parametrize("kind", ["dipole_uniform", "patch_uniform", "xpol_panels"])
@pytest.mark.parametrize("cdl", ["C", "D"])
def test_non_ideal_antenna_elements_through_dropin(ref, kind, cdl):
    """Dipole / patch / cross-polarized linear elements (core/antennas.py:447-622): the patched ``_propagate`` and
    ``state`` serve them from ``cdl_ray_kernel``'s element models -- no reference code on the path."""
    import hermespy.channel as RC
    from hermespy.core import Signal, Transformation
    from hermespy.simulation import SimulatedDevice

    txa, rxa = _element_arrays()[kind]
    tx = SimulatedDevice(bandwidth=CDL_FS, oversampling_factor=1, carrier_frequency=CDL_FC, antennas=txa(),
                         pose=Transformation.From_RPY(np.array([0.1, 0.2, 0.3]), np.array([0.0, 0.0, 25.0])))
    rx = SimulatedDevice(bandwidth=CDL_FS, oversampling_factor=1, carrier_frequency=CDL_FC, antennas=rxa(),
                         pose=Transformation.From_RPY(np.array([0.0, 0.0, 2.5]), np.array([100.0, 20.0, 1.5])),
                         velocity=np.array([10.0, -3.0, 0.0]))
    ch = RC.CDL(getattr(RC.CDLType, cdl), 300e-9, seed=11, **({"rayleigh_factor": 6.0} if cdl == "D" else {}))
    s = ch.realize().sample(tx, rx)
    ntx = s.num_transmit_antennas
    sig = Signal.Create(golden_signal(600, ntx, 1200), CDL_FS, CDL_FC)
    ref.disable()
    y0 = np.asarray(s.propagate(sig).view(np.ndarray))
    c0 = np.asarray(s.state(24, 1000).dense_state())
    before = _launches()
    ref.enable(precision="f64")
    y64 = np.asarray(s.propagate(sig).view(np.ndarray))
    c64 = np.asarray(s.state(24, 1000).dense_state())
    mid = _launches()
    ref.enable(precision="f32")
    y32 = np.asarray(s.propagate(sig).view(np.ndarray))
    ref.disable()
    assert mid - before >= 4
    _assert_cuda_path_ran(mid, y32, y0, at_least=3)
    assert rel_l2(y64, y0) < 1e-10 and rel_l2(c64, c0) < 1e-10
    assert rel_l2(y32, y0) < 1e-5


def test_unknown_antenna_element_raises_instead_of_falling_back(ref):
    """No silent CPU path: an element class without a device pattern is a hard error; the opt-in fallback is counted."""
    import hermespy.channel as RC
    from hermespy.core import Signal, Transformation
    from hermespy.simulation import SimulatedDevice, SimulatedIdealAntenna, SimulatedUniformArray

    from hermespy_b200 import _lib

    class Cardioid(SimulatedIdealAntenna):
        def local_characteristics(self, azimuth, elevation):
            return np.array([0.5 * (1 + np.cos(azimuth)), 0.0])

        def copy(self):
            return Cardioid(self.mode, self.pose.copy())

    def dev(element, pos):
        return SimulatedDevice(bandwidth=CDL_FS, oversampling_factor=1, carrier_frequency=CDL_FC,
                               antennas=SimulatedUniformArray(element, CDL_SPACING, (2, 1, 1)),
                               pose=Transformation.From_Translation(np.array(pos, float)))

    s = RC.CDL(RC.CDLType.A, 300e-9, seed=5).realize().sample(dev(Cardioid, (0, 0, 10.0)), dev(SimulatedIdealAntenna, (50.0, 5.0, 1.5)))
    sig = Signal.Create(golden_signal(601, 2, 256), CDL_FS, CDL_FC)
    ref.enable(precision="f32")
    try:
        with pytest.raises(_lib.HermesB200Error):
            s.propagate(sig)
        with pytest.raises(_lib.HermesB200Error):
            s.state(8, 100)
    finally:
        ref.disable()
    y0 = np.asarray(s.propagate(sig).view(np.ndarray))
    n0 = ref.fallbacks["cdl_propagate"]
    ref.enable(precision="f32", allow_reference_fallback=True)
    try:
        with pytest.warns(RuntimeWarning):
            ref._warned.discard("cdl_propagate")
            y1 = np.asarray(s.propagate(sig).view(np.ndarray))
    finally:
        ref.disable()
        ref.enable(precision="f32")  # clears the opt-in
        ref.disable()
    assert np.array_equal(y0, y1) and ref.fallbacks["cdl_propagate"] == n0 + 1


# ---- BASELINE shapes, built from the REFERENCE classes -------------------------------------------

def _baseline_c2(RC, S):
    from hermespy.core import Transformation

    dev = lambda pos: S.SimulatedDevice(bandwidth=30.72e6, oversampling_factor=1, carrier_frequency=3.5e9,
                                        antennas=S.SimulatedUniformArray(S.SimulatedIdealAntenna, 0.04, (4, 1, 1)),
                                        pose=Transformation.From_Translation(np.array(pos, float)))
    ch = RC.TDL(RC.TDLType.B, rms_delay=300e-9, doppler_frequency=100, seed=42,
                antenna_correlation=RC.StandardAntennaCorrelation(RC.CorrelationType.MEDIUM))
    return ch.realize().sample(dev((0, 0, 0)), dev((50, 10, 0))), 4, 15344, 30.72e6, 1e-12


def _baseline_c3(RC, S):
    from hermespy.core import Transformation

    fc, fs = 3.5e9, 30.72e6
    lam = 299792458.0 / fc
    tx = S.SimulatedDevice(bandwidth=fs, oversampling_factor=1, carrier_frequency=fc,
                           antennas=S.SimulatedUniformArray(S.SimulatedIdealAntenna, lam / 2, (8, 4, 1)),
                           pose=Transformation.From_Translation(np.array([0.0, 0.0, 25.0])))
    rx = S.SimulatedDevice(bandwidth=fs, oversampling_factor=1, carrier_frequency=fc,
                           antennas=S.SimulatedUniformArray(S.SimulatedIdealAntenna, lam / 2, (2, 2, 1)),
                           pose=Transformation.From_Translation(np.array([100.0, 20.0, 1.5])), velocity=np.array([10.0, -3.0, 0.0]))
    return RC.CDL(RC.CDLType.C, 300e-9, seed=42).realize().sample(tx, rx), 32, 2048, fs, 1e-10


def _baseline_c5(RC, S):
    dev = lambda: S.SimulatedDevice(bandwidth=30.72e6, oversampling_factor=1, carrier_frequency=3.5e9)
    return RC.Cost259(RC.Cost259Type.URBAN, doppler_frequency=50, seed=42).realize().sample(dev(), dev()), 1, 1 << 20, 30.72e6, 1e-12


@pytest.mark.parametrize("build", [_baseline_c2, _baseline_c3, _baseline_c5], ids=["C2_4x4_tdl_b_15344", "C3_cdl_c_32x4_2048",
                                                                                    "C5_cost259_siso_1M"])
def test_baseline_shapes_from_reference_samples(ref, build):
    """A reference-built sample of the BASELINE shape itself (4x4 TDL-B T = 15 344; CDL-C 32x4 T = 2 048, moving receiver;
    COST259 urban SISO T = 2^20), propagated by the reference's numpy code and by the patched CUDA path."""
    import hermespy.channel as RC
    import hermespy.simulation as S
    from hermespy.core import Signal

    s, ntx, T, fs, tol64 = build(RC, S)
    sig = Signal.Create(golden_signal(700, ntx, T), fs, 3.5e9)
    ref.disable()
    y0 = np.asarray(s.propagate(sig).view(np.ndarray))
    before = _launches()
    ref.enable(precision="f64")
    y64 = np.asarray(s.propagate(sig).view(np.ndarray))
    mid = _launches()
    ref.enable(precision="f32")
    y32 = np.asarray(s.propagate(sig).view(np.ndarray))
    ref.disable()
    assert mid - before >= 1
    _assert_cuda_path_ran(mid, y32, y0)
    assert y0.shape == y64.shape == y32.shape
    assert rel_l2(y64, y0) < tol64
    assert rel_l2(y32, y0) < 1e-5


def test_sinc_extension_is_opt_in_and_matches_its_oracle(ref):
    """The reference rounds fading delays whatever the interpolation mode (fading.py:297): so does the drop-in, unless
    ``enable(sinc_extension=True)``; then ``InterpolationMode.SINC`` requests get windowed-sinc fractional delays."""
    import hermespy.channel as RC
    import hermespy.simulation as S
    from hermespy.core import InterpolationMode, Signal
    from oracle import fading_oracle as fo
    from oracle.ref_extract import fading_params_from_reference_sample

    dev = lambda: S.SimulatedDevice(bandwidth=30.72e6, oversampling_factor=1, carrier_frequency=3.5e9)
    s = RC.Cost259(RC.Cost259Type.URBAN, doppler_frequency=50, seed=9).realize().sample(dev(), dev())
    sig = Signal.Create(golden_signal(800, 1, 4096), 30.72e6, 3.5e9)
    ref.disable()
    y_ref = np.asarray(s.propagate(sig, InterpolationMode.SINC).view(np.ndarray))  # the reference: rounds anyway
    ref.enable(precision="f64")
    y_default = np.asarray(s.propagate(sig, InterpolationMode.SINC).view(np.ndarray))
    ref.enable(precision="f64", sinc_extension=True)
    y_sinc = np.asarray(s.propagate(sig, InterpolationMode.SINC).view(np.ndarray))
    y_nearest = np.asarray(s.propagate(sig, InterpolationMode.NEAREST).view(np.ndarray))
    ref.disable()
    assert rel_l2(y_default, y_ref) < 1e-12 and rel_l2(y_nearest, y_ref) < 1e-12
    want = fo.propagate_sinc(fading_params_from_reference_sample(s), np.asarray(sig.view(np.ndarray)))
    assert y_sinc.shape == want.shape and rel_l2(y_sinc, want) < 1e-12
    assert rel_l2(y_sinc[:, : y_ref.shape[1]], y_ref) > 1e-2  # fractional delays do change the signal
