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
