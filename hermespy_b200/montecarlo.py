"""Sharding of the Monte-Carlo campaign over one process per GPU and the evaluator statistics they exchange.

Mirrors what the reference does with Ray actors (hermespy/core/pymonte/monte_carlo.py:358-371, actors.py:77-225):
every (grid cell, drop) is independent, rank ``r`` of ``W`` takes drops ``r, r + W, ...`` of every grid cell with the
scenario seed ``base + r * 12345678`` (hermespy/simulation/simulation.py:220-223), and the only thing that ever
crosses GPUs is the per-cell running statistics ``(sum, sum^2, count)`` of ``ScalarEvaluationResult.add_artifact``
(hermespy/core/pymonte/scalar.py:109-115) plus exact integer ``(bit errors, bits)`` counters -- one
``all_reduce(SUM)`` over NCCL / NVLink at the end of a campaign (or every K batches).

The local reduction runs on the GPU (``hb_bit_errors`` / ``hb_stats_accumulate``) straight into the tensors handed
to the collective; there is no CPU fallback for it.  The collective itself is ``torch.distributed`` plumbing and
therefore also runs under ``gloo`` (CPU tests of the sharding logic).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence

import numpy as np

from . import _lib

RANK_SEED_STRIDE = 12345678  # hermespy/simulation/simulation.py:220-223


def rank_seed(base_seed: int, rank: int) -> int:
    """Scenario seed of the process with the given rank (same scheme as the reference's per-actor seeds)."""
    return int(base_seed) + int(rank) * RANK_SEED_STRIDE


def shard_drops(num_drops: int, rank: int, world_size: int) -> range:
    """Drop indices of every grid cell owned by ``rank``: ``rank, rank + W, ...`` (balanced per cell)."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside a world of {world_size}")
    if num_drops < 0:
        raise ValueError("number of drops must be non-negative")
    return range(rank, num_drops, world_size)


def shard_links(grid_shape: Sequence[int], num_drops: int, rank: int, world_size: int):
    """(cell, drop) pairs of this rank in launch order: cells outer, drops inner -- one link batch per call."""
    cells = int(np.prod(grid_shape)) if len(grid_shape) else 1
    drops = np.asarray(shard_drops(num_drops, rank, world_size), dtype=np.int64)
    cell = np.repeat(np.arange(cells, dtype=np.int32), len(drops))
    drop = np.tile(drops, cells)
    return cell, drop


def _torch():
    import torch

    return torch


class GridStatistics:
    """Running evaluator statistics of one rank, laid out as the collective reads them.

    ``stats``  float64 ``[cells, 3]``: sum, sum of squares, count of the scalar artifacts (scalar.py:109-115);
    ``counts`` int64 ``[cells, 2]``: bit errors, compared bits (exact).
    """

    def __init__(self, grid_shape: Sequence[int], device=None) -> None:
        torch = _torch()
        self.grid_shape = tuple(int(g) for g in grid_shape)
        self.num_cells = int(np.prod(self.grid_shape)) if self.grid_shape else 1
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.stats = torch.zeros((self.num_cells, 3), dtype=torch.float64, device=self.device)
        self.counts = torch.zeros((self.num_cells, 2), dtype=torch.int64, device=self.device)

    # ---- local reduction (GPU only) ---------------------------------------------------------------------------
    def _require_cuda(self) -> None:
        if self.device.type != "cuda":
            raise _lib.HermesB200Error(_lib.HB_ERR_NO_DEVICE, "evaluator statistics are reduced on the GPU "
                                                               "(no CPU fallback); create GridStatistics on a cuda device")

    def accumulate(self, artifact, cell, errors=None, bits=None) -> None:
        """``stats[cell[i]] += (a_i, a_i^2, 1)``, ``counts[cell[i]] += (errors_i, bits_i)`` for device tensors."""
        torch = _torch()
        self._require_cuda()
        n = int(artifact.shape[0])
        if int(cell.shape[0]) != n:
            raise ValueError("artifact and cell index differ in length")
        artifact = artifact.to(torch.float64).contiguous()
        cell = cell.to(torch.int32).contiguous()
        if n and (int(cell.min()) < 0 or int(cell.max()) >= self.num_cells):
            raise ValueError("grid cell index outside the grid")
        e = errors.to(torch.int64).contiguous() if errors is not None else None
        b = bits.to(torch.int64).contiguous() if bits is not None else None
        st = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().hb_stats_accumulate(
                artifact.data_ptr(), cell.data_ptr(), e.data_ptr() if e is not None else None,
                b.data_ptr() if b is not None else None, n, self.num_cells, self.stats.data_ptr(),
                self.counts.data_ptr(), C.c_void_p(st)))

    def accumulate_bits(self, tx_bits, rx_bits, cell, tx_len=None, rx_len=None):
        """Bit-error artifacts of a batch of drops (uint8 ``[drops, bits]`` device tensors) and their statistics."""
        torch = _torch()
        self._require_cuda()
        if tx_bits.shape != rx_bits.shape or tx_bits.dim() != 2:
            raise ValueError("bit tensors must be [drops, bits] of equal shape (pad and pass the lengths)")
        n, nb = int(tx_bits.shape[0]), int(tx_bits.shape[1])
        tx = tx_bits.to(torch.uint8).contiguous()
        rx = rx_bits.to(torch.uint8).contiguous()
        errors = torch.empty(n, dtype=torch.int64, device=self.device)
        bits = torch.empty(n, dtype=torch.int64, device=self.device)
        art = torch.empty(n, dtype=torch.float64, device=self.device)
        tl = tx_len.to(torch.int32).contiguous() if tx_len is not None else None
        rl = rx_len.to(torch.int32).contiguous() if rx_len is not None else None
        st = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().hb_bit_errors(
                tx.data_ptr(), rx.data_ptr(), tl.data_ptr() if tl is not None else None,
                rl.data_ptr() if rl is not None else None, n, nb, errors.data_ptr(), bits.data_ptr(), art.data_ptr(),
                C.c_void_p(st)))
        self.accumulate(art, cell, errors, bits)
        return errors, bits, art

    # ---- the collective ---------------------------------------------------------------------------------------
    def all_reduce(self, group=None, async_op: bool = False):
        """Sum the statistics over all ranks (NCCL over NVLink on GPUs, gloo in the CPU tests).

        ONE collective: the integer counters ride along as float64 (exact below 2^53 bits) in a packed ``[cells, 5]``
        buffer.  With ``async_op`` the collective is enqueued behind the work already on the current stream and runs on
        the backend's own stream, so the next batch's kernels overlap it; call the returned function to finish
        (waits and unpacks).  Without it the call is complete on return.
        """
        import torch.distributed as dist

        torch = _torch()
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return (lambda: None) if async_op else None
        packed = torch.cat((self.stats, self.counts.to(torch.float64)), dim=1)
        work = dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group, async_op=True)

        def finish():
            work.wait()
            self.stats.copy_(packed[:, :3])
            self.counts.copy_(packed[:, 3:].round().to(torch.int64))

        if async_op:
            return finish
        finish()
        return None

    # ---- read-out (what ScalarEvaluationResult reports) ---------------------------------------------------------
    def mean(self) -> np.ndarray:
        s = self.stats.detach().cpu().numpy()
        with np.errstate(invalid="ignore", divide="ignore"):
            return (s[:, 0] / s[:, 2]).reshape(self.grid_shape or (1,))

    def bit_error_rate(self) -> np.ndarray:
        """Exact ratio of the integer counters (differs from ``mean`` when drops carry different bit counts)."""
        c = self.counts.detach().cpu().numpy().astype(np.float64)
        with np.errstate(invalid="ignore", divide="ignore"):
            return (c[:, 0] / c[:, 1]).reshape(self.grid_shape or (1,))

    def confident(self, cell: int, accuracy: float, confidence: float, min_num_samples: int = 1) -> bool:
        """Stopping rule of ScalarEvaluationResult.add_artifact (scalar.py:117-123) on the reduced statistics.

        As in the reference the rule is only evaluated when the cell's sample count is a multiple of
        ``min_num_samples`` (scalar.py:117); between those counts the cell is never declared confident."""
        s, s2, n = (float(v) for v in self.stats[cell].detach().cpu())
        if n < 2 or min_num_samples < 1 or int(round(n)) % int(min_num_samples) != 0:
            return False
        var = (s2 - s * s / n) / (n - 1)
        std = math.sqrt(var) if var > 0 else 0.0
        if std <= 0.0:
            return False
        from scipy.stats import norm

        return 2.0 * (1.0 - norm.cdf(math.sqrt(n) * accuracy / std)) < confidence


def kron_mix(spatial, r_rx=None, r_tx=None, out=None):
    """``R_rx @ S @ R_tx`` for a batch of device spatial matrices (complex128), fading.py:480-489."""
    torch = _torch()
    if not spatial.is_cuda:
        raise _lib.HermesB200Error(_lib.HB_ERR_NO_DEVICE, "hb_kron_mix needs device tensors (no CPU fallback)")
    s = spatial.to(torch.complex128).contiguous()
    B, nrx, ntx = (int(v) for v in s.shape)
    rr = r_rx.to(torch.complex128).to(s.device).contiguous() if r_rx is not None else None
    rt = r_tx.to(torch.complex128).to(s.device).contiguous() if r_tx is not None else None
    if rr is not None and tuple(rr.shape) != (nrx, nrx):
        raise ValueError("receive correlation must be [Nrx, Nrx]")
    if rt is not None and tuple(rt.shape) != (ntx, ntx):
        raise ValueError("transmit correlation must be [Ntx, Ntx]")
    out = torch.empty_like(s) if out is None else out
    st = torch.cuda.current_stream(s.device).cuda_stream
    with torch.cuda.device(s.device):
        _lib.check(_lib.load().hb_kron_mix(rr.data_ptr() if rr is not None else None, s.data_ptr(),
                                           rt.data_ptr() if rt is not None else None, out.data_ptr(), B, nrx, ntx,
                                           C.c_void_p(st)))
    return out
