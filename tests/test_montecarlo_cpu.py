"""CPU: drop sharding, rank seeds and the evaluator-statistics all-reduce under gloo with world_size 2."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hermespy_b200 import _lib
from hermespy_b200.montecarlo import GridStatistics, rank_seed, shard_drops, shard_links
from oracle import stats_oracle as so


def test_shards_partition_every_cell():
    for W in (1, 2, 3, 8):
        for n in (0, 1, 7, 1000):
            seen = np.concatenate([np.asarray(shard_drops(n, r, W), dtype=np.int64) for r in range(W)])
            assert sorted(seen.tolist()) == list(range(n))
            sizes = [len(shard_drops(n, r, W)) for r in range(W)]
            assert max(sizes) - min(sizes) <= 1  # balanced per grid cell, like the round-robin of actors.py:91-96
    with pytest.raises(ValueError):
        shard_drops(10, 2, 2)
    cell, drop = shard_links((7,), 10, 1, 4)
    assert cell.shape == drop.shape == (7 * 3,) and set(drop.tolist()) == {1, 5, 9}
    assert (np.bincount(cell) == 3).all()


def test_rank_seed_scheme():
    assert rank_seed(42, 0) == 42 and rank_seed(42, 3) == 42 + 3 * 12345678  # simulation.py:220-223


def test_statistics_need_a_gpu():
    gs = GridStatistics((3,), device="cpu")
    with pytest.raises(_lib.HermesB200Error) as e:
        gs.accumulate(torch.zeros(2, dtype=torch.float64), torch.zeros(2, dtype=torch.int32))
    assert e.value.status == _lib.HB_ERR_NO_DEVICE


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, grid, drops, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cells = int(np.prod(grid))
        cell, drop = shard_links(grid, drops, rank, world)
        # synthetic per-drop evaluator output, a pure function of (cell, drop) so that every world size agrees
        rng_bits = 64 + (drop % 5) * 8
        errors = (cell.astype(np.int64) * 7 + drop * 13) % 11
        art = errors / rng_bits
        gs = GridStatistics(grid, device="cpu")
        # the local reduction is GPU-only in the product; here the oracle stands in for it (CPU test of the exchange)
        st, ct = so.add_artifacts(art, cell, cells, errors, rng_bits)
        gs.stats.copy_(torch.from_numpy(st))
        gs.counts.copy_(torch.from_numpy(ct))
        gs.all_reduce()
        if rank == 0:
            np.savez(out, stats=gs.stats.numpy(), counts=gs.counts.numpy(), mean=gs.mean(), ber=gs.bit_error_rate())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_all_reduce_of_sharded_statistics_equals_single_process(tmp_path, world):
    grid, drops = (3, 2), 37
    out = str(tmp_path / "reduced.npz")
    mp.spawn(_worker, args=(world, _free_port(), grid, drops, out), nprocs=world, join=True)
    got = np.load(out)
    cells = int(np.prod(grid))
    cell, drop = shard_links(grid, drops, 0, 1)
    bits = 64 + (drop % 5) * 8
    errors = (cell.astype(np.int64) * 7 + drop * 13) % 11
    st, ct = so.add_artifacts(errors / bits, cell, cells, errors, bits)
    np.testing.assert_array_equal(got["counts"], ct)  # integer counters: bit exact for any sharding
    np.testing.assert_allclose(got["stats"], st, rtol=1e-13)
    np.testing.assert_array_equal(got["stats"][:, 2], np.full(cells, drops))
    np.testing.assert_allclose(got["ber"].ravel(), ct[:, 0] / ct[:, 1], rtol=0, atol=0)
