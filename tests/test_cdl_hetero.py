"""Heterogeneous CDL batches: links with their OWN delay structure (cluster delays, ray count, line-of-sight state, delay
spread -- what every realization of a stochastic 3GPP scenario draws, cluster_delay_lines.py:1824-2013) in ONE launch set
through per-link delay tables (``hb_cdl_problem.link_term_delay``).  CPU part: the padding of ``CdlBlock.stack`` and the
planner; GPU part: every K6 kernel family against the float64 oracle and against the per-link launches it replaces."""
import ctypes as C

import numpy as np
import pytest

import hermespy_b200.channel as MC
from hermespy_b200 import _lib
from oracle import cdl_oracle as co
from tests.helpers import rel_l2
from tests.test_cdl_golden import mirror_cdl_device, oracle_params

#: (model, rms delay spread): 23 / 24 clusters without a line of sight, 13 / 14 with one, three delay spreads
MIX = [("C", 300e-9), ("D", 300e-9), ("A", 100e-9), ("E", 1000e-9), ("C", 30e-9), ("B", 300e-9)]


def _mixed_samples(dims_tx, dims_rx, speed, T, seed=3, mix=None):
    rng = np.random.default_rng(seed)
    tx = mirror_cdl_device((dims_tx, (0.0, 0.1, 0.0), (0.0, 0.0, 25.0), (0, 0, 0)))
    ntx = int(np.prod(dims_tx))
    samples, xs = [], []
    for b, (model, spread) in enumerate(MIX if mix is None else mix):
        rx = mirror_cdl_device((dims_rx, (0, 0, 0.3 * b), (100.0 + 7 * b, 20.0, 1.5), speed))
        samples.append(MC.CDL(getattr(MC.CDLType, model), spread, seed=seed + b).realize().sample(tx, rx))
        xs.append((rng.standard_normal((ntx, T)) + 1j * rng.standard_normal((ntx, T))) / np.sqrt(2))
    return samples, xs


def test_stack_pads_to_one_heterogeneous_block_and_the_planner_takes_it():
    from hermespy_b200.kernels import CdlBlock, cdl_plan

    samples, _ = _mixed_samples((2, 2, 1), (2, 1, 1), (3.0, 0.0, 0.0), 64)
    blocks = [s.kernel_block() for s in samples]
    assert len({b.delay_key() for b in blocks}) == len(blocks) and len({b.geometry_key() for b in blocks}) == 1
    blk = CdlBlock.stack(blocks)
    rn = max(b.term_delay.shape[0] for b in blocks)
    assert blk.batch == len(blocks) and blk.link_term_delay.shape == (len(blocks), rn) and blk.angles.shape == (len(blocks), rn, 4)
    assert blk.max_delay == max(b.max_delay for b in blocks) and blk.line_of_sight
    for k, b in enumerate(blocks):
        n = b.term_delay.shape[0]
        assert np.array_equal(blk.link_term_delay[k, :n], b.term_delay) and not blk.amplitude[k, n:].any()
        assert np.array_equal(blk.amplitude[k, :n], b.amplitude[0]) and np.isfinite(blk.angles[k]).all()
        assert blk.link_los_amplitude[k] == (b.los_amplitude if b.line_of_sight else 0.0)
        assert blk.link_max_delay[k] == b.max_delay
    # uniform blocks keep the launch-uniform table
    same = CdlBlock.stack([blocks[0], blocks[0]])
    assert same.link_term_delay is None and same.batch == 2
    # a heterogeneous block stacks again (nested batching)
    again = CdlBlock.stack([blk, blocks[1]])
    assert again.batch == blk.batch + 1 and again.link_term_delay.shape[1] == rn
    # blocks of another array geometry do not belong in the batch
    other, _ = _mixed_samples((2, 1, 1), (2, 1, 1), (3.0, 0.0, 0.0), 64, mix=MIX[:1])
    with pytest.raises(ValueError, match="share array topologies"):
        CdlBlock.stack([blocks[0], other[0].kernel_block()])
    # the planner: groups = the largest per-link count, no device needed
    plan = cdl_plan(blk, 2048)
    per_link = [np.unique(np.append(b.term_delay, b.los_delay) if b.line_of_sight else b.term_delay).size for b in blocks]
    assert plan["num_groups"] == max(per_link) and plan["mode"] == "poly"


def test_per_link_tables_are_validated_on_the_host():
    from hermespy_b200.kernels import CdlBlock, _cdl_problem

    samples, _ = _mixed_samples((2, 1, 1), (1, 1, 1), (0.0, 0.0, 0.0), 32)
    blk = CdlBlock.stack([s.kernel_block() for s in samples[:3]])
    lib = _lib.load()
    info = _lib.FadingPlanInfo()
    blk.link_term_delay[1, 0] = blk.max_delay + 1  # a delay index beyond the batch's max_delay
    p = _cdl_problem(blk, 32, "f32", False, {k: None for k in CdlBlock.ARRAYS})
    assert lib.hb_cdl_plan(C.byref(p), C.byref(info)) == _lib.HB_ERR_INVALID
    assert b"delay index" in lib.hb_last_error()
    blk.link_term_delay[1, 0] = 0
    p = _cdl_problem(blk, 32, "f64", True, {k: None for k in CdlBlock.ARRAYS})
    assert lib.hb_cdl_plan(C.byref(p), C.byref(info)) == _lib.HB_OK
    # channel state is indexed by delay group: one delay structure per call
    G = C.c_int32(0)
    assert lib.hb_cdl_state(C.byref(p), None, None, C.byref(G), None) == _lib.HB_ERR_UNSUPPORTED


def _k6_launches():
    c = _lib.launch_counts()
    return c.get("cdl_propagate", 0), c.get("cdl_rays", 0)


@pytest.mark.gpu
@pytest.mark.parametrize("dims_tx,dims_rx,speed,T,variant", [
    ((2, 2, 1), (2, 1, 1), (3.0, -1.0, 0.0), 700, "auto"),      # FP32-pipe K6 (below 8 transmit antennas)
    ((8, 4, 1), (2, 2, 1), (10.0, -3.0, 0.0), 2048, "auto"),    # config C3 shape: tensor-core K6
    ((4, 2, 1), (2, 2, 1), (10.0, -3.0, 0.0), 900, "umma_bf16"),  # BF16x3 tensor-core K6
    ((3, 1, 1), (5, 2, 1), (100.0, 50.0, 5.0), 300, "auto"),    # fast link, 10 receive antennas: two receive chunks
    ((1, 1, 1), (1, 1, 1), (300.0, 0.0, 0.0), 500, "auto"),     # Doppler beyond four Taylor terms: per-ray path in f32 too
])
def test_heterogeneous_batch_matches_oracle_and_per_link_launches(dims_tx, dims_rx, speed, T, variant):
    from hermespy_b200.kernels import CdlBlock, cdl_propagate_host

    # the BF16x3 kernel stages 256-sample tiles: the 1 us delay spread's 640-sample halo does not fit its operand images
    mix = [m for m in MIX if m[1] <= 300e-9] if variant == "umma_bf16" else MIX
    samples, xs = _mixed_samples(dims_tx, dims_rx, speed, T, mix=mix)
    blocks = [s.kernel_block() for s in samples]
    blk = CdlBlock.stack(blocks)
    assert blk.link_term_delay is not None
    x = np.stack(xs)
    refs = [co.propagate(oracle_params(s), xi) for s, xi in zip(samples, xs)]
    for precision, tol in (("f64", 1e-10), ("f32", 1e-5)):
        v = variant if precision == "f32" else "auto"
        before = _k6_launches()
        y, info = cdl_propagate_host(x, blk, precision=precision, variant=v, return_info=True)
        after = _k6_launches()
        nrx_chunks = -(-blk.num_rx // 8) if info["mode"] == "poly" else 1
        assert after[0] - before[0] == nrx_chunks, "one K6 launch set for the whole heterogeneous batch"
        if variant != "auto" and precision == "f32":
            assert info["variant"] == variant
        for k, (b, r) in enumerate(zip(blocks, refs)):
            own = y[k][:, : T + b.max_delay]
            assert own.shape == r.shape
            assert rel_l2(own, r) < tol, (precision, k, mix[k])
            assert not y[k][:, T + b.max_delay:].any()  # nothing beyond the link's own delay spread
            alone = cdl_propagate_host(xs[k][None], b, precision=precision, variant=v)[0]
            # same kernels, same summation order; the zero-amplitude padding adds exact zeros (the Taylor window / order
            # of the batch may differ from the single link's, hence not bitwise in f32)
            assert rel_l2(own, alone) < (1e-13 if precision == "f64" else 3e-6), (precision, k)


@pytest.mark.gpu
def test_device_entry_equals_host_entry_and_chunks():
    import torch
    from hermespy_b200.kernels import CdlBlock, CdlDeviceBlock, cdl_propagate, cdl_propagate_host

    samples, xs = _mixed_samples((8, 2, 1), (2, 2, 1), (5.0, 2.0, 0.0), 1024)
    blk = CdlBlock.stack([s.kernel_block() for s in samples])
    x = np.stack(xs).astype(np.complex64)
    whole = cdl_propagate_host(x, blk, precision="f32")
    chunked = cdl_propagate_host(x, blk, precision="f32", chunk_links=4)  # 4 + 2 links: the tables are sliced per chunk
    dev = cdl_propagate(torch.from_numpy(x).cuda(), CdlDeviceBlock(blk), precision="f32").cpu().numpy()
    assert np.array_equal(whole, chunked) and np.array_equal(whole, dev)


@pytest.mark.gpu
def test_mirror_batch_call_groups_by_geometry():
    from hermespy_b200.channel.cdl import cdl_propagate_batch

    samples, xs = _mixed_samples((2, 2, 1), (2, 1, 1), (3.0, 0.0, 0.0), 256)
    before = _k6_launches()
    ys = cdl_propagate_batch(samples, xs, precision="f64")
    assert _k6_launches()[0] - before[0] == 1
    for s, xi, y in zip(samples, xs, ys):
        r = co.propagate(oracle_params(s), xi)
        assert y.shape == r.shape and rel_l2(y, r) < 1e-10


@pytest.mark.gpu
def test_heterogeneous_batch_with_per_element_patterns():
    """Rank-two ray matrices (cross-polarized linear elements with their own orientations) in a heterogeneous batch: the
    links differ in their cluster delays, delay spread and line-of-sight state."""
    from dataclasses import replace

    from hermespy_b200.kernels import CdlBlock, cdl_propagate_host
    from tests.helpers import cdl_block_from_oracle_params, cdl_params_from_golden
    from tests.test_cdl_elements import GOLDEN as ELEMENT_GOLDEN

    golden = np.load(ELEMENT_GOLDEN)
    p = cdl_params_from_golden(golden, "xpol_linear_custom")
    rng = np.random.default_rng(9)
    blocks, xs, refs = [], [], []
    T = 300
    for b in range(4):
        q = replace(p, cluster_delays=p.cluster_delays * (1.0 + 0.45 * b), cluster_delay_spread=p.cluster_delay_spread * (1 + b),
                    line_of_sight=bool(b % 2), rice_factor_db=3.0 + b,
                    rx=replace(p.rx, velocity=p.rx.velocity * (1 + b)))
        blocks.append(cdl_block_from_oracle_params(q))
        x = (rng.standard_normal((blocks[-1].num_tx, T)) + 1j * rng.standard_normal((blocks[-1].num_tx, T))) / np.sqrt(2)
        xs.append(x)
        refs.append(co.propagate(q, x))
    assert len({b.delay_key() for b in blocks}) == 4
    blk = CdlBlock.stack(blocks)
    assert blk.link_term_delay is not None and blk.element_mode == _lib.HB_ELEMENTS_PER_ELEMENT
    for precision, tol in (("f64", 1e-10), ("f32", 1e-5)):
        y = cdl_propagate_host(np.stack(xs), blk, precision=precision)
        for k, (b, r) in enumerate(zip(blocks, refs)):
            assert rel_l2(y[k][:, : T + b.max_delay], r) < tol, (precision, k)


@pytest.mark.gpu
def test_heterogeneous_batch_on_a_large_array():
    """96 transmit antennas: the steering phases of a pass no longer fit the all-windows moment kernel's shared memory, the
    per-window moment kernel (global-memory steering phases) serves the heterogeneous batch."""
    from hermespy_b200.kernels import CdlBlock, cdl_propagate_host

    samples, xs = _mixed_samples((12, 8, 1), (2, 1, 1), (4.0, 1.0, 0.0), 260, mix=MIX[:4])
    blocks = [s.kernel_block() for s in samples]
    blk = CdlBlock.stack(blocks)
    y, info = cdl_propagate_host(np.stack(xs), blk, precision="f32", return_info=True)
    assert info["mode"] == "poly"
    for k, (s, xi, b) in enumerate(zip(samples, xs, blocks)):
        r = co.propagate(oracle_params(s), xi)
        assert rel_l2(y[k][:, : 260 + b.max_delay], r) < 1e-5, k
