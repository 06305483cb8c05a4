"""GPU parity: CDL kernels (through the C-ABI) against the numpy oracle and the reference's golden vectors."""
import numpy as np
import pytest

import hermespy_b200.channel as MC
from hermespy_b200 import config
from hermespy_b200.core import Signal
from oracle import cdl_oracle as co
from oracle.golden_cases import CDL_CASES, CDL_FC, CDL_FS, golden_signal
from tests.helpers import rel_l2
from tests.test_cdl_golden import GOLDEN, mirror_cdl_device, mirror_cdl_sample, oracle_params

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


@pytest.mark.parametrize("ci", range(len(CDL_CASES)), ids=[c[0] for c in CDL_CASES])
def test_plugin_propagate_matches_reference_golden(golden, ci):
    case = CDL_CASES[ci]
    name, _, txs, rxs, T = case
    real, s, tx, rx = mirror_cdl_sample(case)
    ntx, nrx = int(np.prod(txs[0])), int(np.prod(rxs[0]))
    x = Signal.Create(golden_signal(200 + ci, ntx, T), CDL_FS, CDL_FC)
    ref = golden[f"{name}/y"]
    with config.compute("f64"):
        y64 = s.propagate(x).view(np.ndarray)
    assert y64.shape == ref.shape
    assert rel_l2(y64, ref) < 1e-10  # ~7e3 rad distance phases limit FP64 agreement to ~1e-12
    with config.compute("f32"):
        y32 = s.propagate(x).view(np.ndarray)
    assert rel_l2(y32, ref) < 1e-5
    rs = real.reciprocal_sample(s, rx, tx)
    with config.compute("f64"):
        yr = rs.propagate(Signal.Create(golden_signal(300 + ci, nrx, T), CDL_FS, CDL_FC)).view(np.ndarray)
    assert rel_l2(yr, golden[f"{name}/y_reciprocal"]) < 1e-10


@pytest.mark.parametrize("name", [c[0] for c in CDL_CASES if c[4] <= 100])
def test_plugin_state_matches_reference_golden(golden, name):
    case = next(c for c in CDL_CASES if c[0] == name)
    _, s, _, _ = mirror_cdl_sample(case)
    csi = golden[f"{name}/csi"]
    dense = s.state(case[4], 1000).dense_state()
    assert dense.shape == csi.shape
    assert np.abs(dense - csi).max() < 1e-10 * np.abs(csi).max()


@pytest.mark.parametrize("dims_tx,dims_rx,speed", [((8, 4, 1), (2, 2, 1), (10.0, -3.0, 0.0)),
                                                    ((4, 4, 1), (4, 2, 1), (0.0, 0.0, 0.0)),
                                                    ((3, 1, 1), (5, 2, 1), (100.0, 50.0, 5.0)),
                                                    ((1, 1, 1), (1, 1, 1), (300.0, 0.0, 0.0))])
def test_batched_kernels_against_oracle(dims_tx, dims_rx, speed):
    """Config C3 shape (32 x 4 UPA, CDL-C, moving receiver) and friends, batched, both precisions."""
    from hermespy_b200.channel.cdl import cdl_propagate_batch

    rng = np.random.default_rng(7)
    T = 700
    tx = mirror_cdl_device((dims_tx, (0.0, 0.1, 0.0), (0.0, 0.0, 25.0), (0, 0, 0)))
    ch = MC.CDL(MC.CDLType.C, 300e-9, seed=3)
    samples, signals, refs = [], [], []
    ntx = int(np.prod(dims_tx))
    for b in range(3):
        rx = mirror_cdl_device((dims_rx, (0, 0, 0.3 * b), (100.0 + 7 * b, 20.0, 1.5), speed))
        s = ch.realize().sample(tx, rx)
        x = (rng.standard_normal((ntx, T)) + 1j * rng.standard_normal((ntx, T))) / np.sqrt(2)
        samples.append(s)
        signals.append(x)
        refs.append(co.propagate(oracle_params(s), x))
    y32 = cdl_propagate_batch(samples, signals, precision="f32")
    y64 = cdl_propagate_batch(samples, signals, precision="f64")
    for a, b_, r in zip(y32, y64, refs):
        assert a.shape == r.shape
        assert rel_l2(b_, r) < 1e-10
        assert rel_l2(a, r) < 1e-5


def test_device_entry_and_plan():
    import torch
    from hermespy_b200.kernels import CdlBlock, CdlDeviceBlock, cdl_plan, cdl_propagate, cdl_propagate_host

    rng = np.random.default_rng(1)
    tx = mirror_cdl_device(((8, 4, 1), (0, 0, 0), (0.0, 0.0, 25.0), (0, 0, 0)))
    rx = mirror_cdl_device(((2, 2, 1), (0, 0, 0), (100.0, 20.0, 1.5), (10.0, -3.0, 0.0)))
    ch = MC.CDL(MC.CDLType.C, 300e-9, seed=42)
    blocks = [ch.realize().sample(tx, rx).kernel_block() for _ in range(4)]
    blk = CdlBlock.stack(blocks)
    T = 2048
    plan = cdl_plan(blk, T)
    assert plan["mode"] == "poly" and plan["num_groups"] == np.unique(blk.term_delay).size and plan["poly_order"] <= 4
    x = (rng.standard_normal((4, 32, T)) + 1j * rng.standard_normal((4, 32, T))).astype(np.complex64)
    yh = cdl_propagate_host(x, blk, precision="f32", chunk_links=3)
    yd = cdl_propagate(torch.from_numpy(x).cuda(), CdlDeviceBlock(blk), precision="f32").cpu().numpy()
    assert np.array_equal(yh, yd)


def test_non_nearest_interpolation_yields_zeros():
    """the reference accumulates nothing for interpolation modes other than NEAREST."""
    from hermespy_b200.core import InterpolationMode

    _, s, _, _ = mirror_cdl_sample(CDL_CASES[3])
    y = s.propagate(Signal.Create(np.ones((1, 16), complex), CDL_FS, CDL_FC), InterpolationMode.SINC).view(np.ndarray)
    assert y.shape[1] > 16 and not y.any()


@pytest.mark.parametrize("dims_tx,dims_rx,speed,T,cdl_type,io", [
    ((8, 4, 1), (2, 2, 1), (10.0, -3.0, 0.0), 2048, "C", np.complex64),    # config C3 shape: 4 M-tiles + the delay tail
    ((8, 4, 1), (2, 2, 1), (10.0, -3.0, 0.0), 700, "C", np.complex128),    # partial window, complex128 frames
    ((4, 4, 1), (4, 2, 1), (0.0, 0.0, 0.0), 300, "A", np.complex64),       # static: P = 1, 8 receive antennas
    ((4, 2, 1), (5, 2, 1), (30.0, 15.0, 2.0), 900, "D", np.complex128),     # LOS term, 10 receive antennas = two chunks of 8
    ((3, 1, 1), (3, 1, 1), (30.0, 0.0, 0.0), 64, "E", np.complex64),       # odd antenna counts, frame shorter than one M-tile
    ((1, 1, 1), (1, 1, 1), (25.0, 0.0, 0.0), 1500, "B", np.complex64),     # SISO: one half-empty K step
    ((5, 3, 1), (2, 1, 1), (3.0, 3.0, 0.0), 1100, "C", np.complex64),      # 15 transmit antennas: last K stage ragged
])
def test_tensor_core_variant_against_oracle_and_gather(dims_tx, dims_rx, speed, T, cdl_type, io):
    """K6 on tcgen05 (``variant="umma"``, cdl_umma.cuh) against the float64 oracle (<= 1e-5, north_star) and against the FP32-pipe
    kernel it replaces; the planner reports which kernel ran."""
    from hermespy_b200 import _lib
    from hermespy_b200.kernels import CdlBlock, cdl_plan, cdl_propagate_host

    rng = np.random.default_rng(11)
    tx = mirror_cdl_device((dims_tx, (0.0, 0.1, 0.0), (0.0, 0.0, 25.0), (0, 0, 0)))
    ch = MC.CDL(getattr(MC.CDLType, cdl_type), 300e-9, seed=5)
    ntx = int(np.prod(dims_tx))
    B = 5
    samples, xs, refs = [], [], []
    for b in range(B):
        rx = mirror_cdl_device((dims_rx, (0, 0, 0.3 * b), (100.0 + 7 * b, 20.0, 1.5), speed))
        s = ch.realize().sample(tx, rx)
        x = (rng.standard_normal((ntx, T)) + 1j * rng.standard_normal((ntx, T))) / np.sqrt(2)
        samples.append(s), xs.append(x), refs.append(co.propagate(oracle_params(s), x))
    groups = {}
    for b, s in enumerate(samples):  # realizations of one static CDL table share their delay table
        groups.setdefault(s.kernel_block().group_key(), []).append(b)
    for idx in groups.values():
        blk = CdlBlock.stack([samples[b].kernel_block() for b in idx])
        x = np.stack([xs[b] for b in idx]).astype(io)
        assert cdl_plan(blk, T, "f32", variant="umma")["variant"] == "umma"
        before = _lib.launch_counts()["cdl_propagate"]
        yu, info = cdl_propagate_host(x, blk, precision="f32", variant="umma", return_info=True)
        assert info["variant"] == "umma" and _lib.launch_counts()["cdl_propagate"] > before
        yg, info_g = cdl_propagate_host(x, blk, precision="f32", variant="gather", return_info=True)
        assert info_g["variant"] == "gather"
        for k, b in enumerate(idx):
            assert yu[k].shape == refs[b].shape
            assert rel_l2(yu[k], refs[b]) < 1e-5
            assert rel_l2(yu[k], yg[k]) < 3e-6


@pytest.mark.parametrize("speed,tile", [((60.0, 0.0, 0.0), 256), ((100.0, 40.0, 5.0), 128)])
def test_fast_links_stay_on_the_taylor_path(speed, tile):
    """Links too fast for four Taylor terms on a 512-sample window run the FP32-pipe kernel with shorter tiles (f32 mode) instead
    of dropping to the per-ray FP64 kernel."""
    from hermespy_b200.kernels import CdlBlock, cdl_propagate_host

    rng = np.random.default_rng(3)
    tx = mirror_cdl_device(((4, 2, 1), (0.0, 0.1, 0.0), (0.0, 0.0, 25.0), (0, 0, 0)))
    rx = mirror_cdl_device(((2, 1, 1), (0, 0, 0.2), (80.0, 20.0, 1.5), speed))
    s = MC.CDL(MC.CDLType.A, 300e-9, seed=9).realize().sample(tx, rx)
    x = (rng.standard_normal((8, 1500)) + 1j * rng.standard_normal((8, 1500))) / np.sqrt(2)
    ref = co.propagate(oracle_params(s), x)
    y, info = cdl_propagate_host(x[None], CdlBlock.stack([s.kernel_block()]), precision="f32", return_info=True)
    assert info["mode"] == "poly" and info["variant"] == "gather" and info["tile"] == tile, info
    assert rel_l2(y[0], ref) < 1e-5


@pytest.mark.parametrize("dims_tx,dims_rx,speed,T,cdl_type,io", [
    ((8, 4, 1), (2, 2, 1), (10.0, -3.0, 0.0), 2048, "C", np.complex64),    # config C3 shape: 8 full tiles of 256 + the delay tail
    ((8, 4, 1), (2, 2, 1), (10.0, -3.0, 0.0), 700, "C", np.complex128),    # partial tile, complex128 frames
    ((4, 4, 1), (4, 2, 1), (0.0, 0.0, 0.0), 300, "A", np.complex64),       # static: P = 1, 8 receive antennas (16 columns per term)
    ((3, 1, 1), (3, 1, 1), (30.0, 0.0, 0.0), 64, "E", np.complex64),       # odd antenna counts, frame shorter than one M-tile
    ((1, 1, 1), (1, 1, 1), (25.0, 0.0, 0.0), 1500, "B", np.complex64),     # SISO: one K stage with one antenna of eight
    ((5, 3, 1), (2, 1, 1), (3.0, 3.0, 0.0), 1100, "D", np.complex64),      # 15 transmit antennas (ragged second K stage), LOS term
])
def test_bf16x3_tensor_core_variant(dims_tx, dims_rx, speed, T, cdl_type, io):
    """K6 as BF16x3 on tcgen05 (``variant="umma_bf16"``, cdl_umma_bf16.cuh) against the float64 oracle and the 3xTF32 kernel."""
    from hermespy_b200 import _lib
    from hermespy_b200.kernels import CdlBlock, cdl_propagate_host

    rng = np.random.default_rng(12)
    tx = mirror_cdl_device((dims_tx, (0.0, 0.1, 0.0), (0.0, 0.0, 25.0), (0, 0, 0)))
    ch = MC.CDL(getattr(MC.CDLType, cdl_type), 300e-9, seed=6)
    ntx = int(np.prod(dims_tx))
    samples, xs, refs = [], [], []
    for b in range(4):
        rx = mirror_cdl_device((dims_rx, (0, 0, 0.3 * b), (100.0 + 7 * b, 20.0, 1.5), speed))
        s = ch.realize().sample(tx, rx)
        x = (rng.standard_normal((ntx, T)) + 1j * rng.standard_normal((ntx, T))) / np.sqrt(2)
        samples.append(s), xs.append(x), refs.append(co.propagate(oracle_params(s), x))
    groups = {}
    for b, s in enumerate(samples):
        groups.setdefault(s.kernel_block().group_key(), []).append(b)
    for idx in groups.values():
        blk = CdlBlock.stack([samples[b].kernel_block() for b in idx])
        x = np.stack([xs[b] for b in idx]).astype(io)
        before = _lib.launch_counts()["cdl_propagate"]
        yb, info = cdl_propagate_host(x, blk, precision="f32", variant="umma_bf16", return_info=True)
        assert info["variant"] == "umma_bf16" and info["tile"] == 256 and _lib.launch_counts()["cdl_propagate"] > before
        yt = cdl_propagate_host(x, blk, precision="f32", variant="umma")
        for k, b in enumerate(idx):
            assert yb[k].shape == refs[b].shape
            assert rel_l2(yb[k], refs[b]) < 1e-5
            assert rel_l2(yb[k], yt[k]) < 3e-6


def test_bf16x3_variant_is_refused_beyond_32_columns():
    from hermespy_b200 import _lib
    from hermespy_b200.kernels import CdlBlock, cdl_plan

    tx = mirror_cdl_device(((4, 2, 1), (0.0, 0.1, 0.0), (0.0, 0.0, 25.0), (0, 0, 0)))
    rx = mirror_cdl_device(((4, 2, 1), (0, 0, 0.0), (100.0, 20.0, 1.5), (20.0, 0.0, 0.0)))  # 8 receive antennas, P > 2
    blk = CdlBlock.stack([MC.CDL(MC.CDLType.C, 300e-9, seed=1).realize().sample(tx, rx).kernel_block()])
    with pytest.raises(_lib.HermesB200Error):
        cdl_plan(blk, 2048, "f32", variant="umma_bf16")
    assert cdl_plan(blk, 2048, "f32", variant="umma")["variant"] == "umma"
