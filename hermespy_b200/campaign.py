"""Monte-Carlo BER campaign over an UNMODIFIED HermesPy scenario, one process per GPU (first form).

What ``Simulation.run()`` does with Ray actors (hermespy/simulation/simulation.py:184-246, core/pymonte/actors.py:255-452)
-- every actor loops ``configure grid cell -> scenario.drop() -> evaluator.evaluate().artifact()`` and a collector sums the
artifacts per cell -- is done here by the ranks of a ``torchrun`` job:

* rank ``r`` of ``W`` takes drops ``r, r + W, ...`` of every grid cell (``montecarlo.shard_drops``, the reference's
  round-robin), with the channel hot path routed to the CUDA kernels by ``hermespy_b200.dropin``;
* transmitted / received bits of the local drops go through ``hb_bit_errors`` / ``hb_stats_accumulate`` into the rank's
  ``GridStatistics`` on the GPU, and ONE all-reduce (NCCL over NVLink) at the end yields the campaign result -- nothing
  else crosses GPUs;
* every (cell, drop) is seeded on its own (scenario, modem, noise models), so the result does not depend on how the drops
  are sharded: a single-process run of the reference's numpy channel and an N-GPU run give IDENTICAL bit-error counts in
  the float64 mode.  (The reference's own campaigns are not reproducible across Ray scheduling)

The scenario is built by a user function from the reference's public API; modems, noise, synchronization, equalization
and the evaluator are reference code.

This is the REPRODUCIBLE form (a result independent of the number of ranks, bit for bit): one drop at a time per rank, so the
rank's Python modem code -- not the channel -- sets its pace (~140 drops/s on C1).  Throughput is the business of the batched
drop runner (``hermespy_b200/runner.py``, second form): B drops in flight per actor, the modem stages spread over forked
helper processes, one channel launch per round for all their links, driven by the unmodified ``Simulation.run()``.
"""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np

from .montecarlo import GridStatistics, shard_drops

DROP_SEED_STRIDE = 7919  # distinct seed blocks per (cell, drop)


def drop_seed(base_seed: int, cell: int, drop: int, num_drops: int) -> int:
    return int(base_seed) + DROP_SEED_STRIDE * (int(cell) * int(num_drops) + int(drop))


def run_ber_campaign(build: Callable, snrs_db: Sequence[float], num_drops: int, *, base_seed: int = 42,
                     rank: int = 0, world_size: int = 1, device=None, use_gpu_channel: bool = True,
                     precision: str = "f64", reseed: Callable = None) -> GridStatistics:
    """Run the local share of a BER-over-SNR campaign and return the (not yet all-reduced) statistics.

    ``build(seed) -> (scenario, tx_device, rx_device, link, evaluator)`` constructs the scenario from the reference API;
    ``evaluator`` is the reference's ``BitErrorEvaluator(link, link)``: its hooks capture the modem's transmission and
    reception of every drop, whose bit vectors are what ``evaluate()`` compares (modem/evaluators.py:239-258).
    ``reseed(scenario, tx, rx, link, seed)`` pins every random root for one drop (default: scenario, modem and the two
    noise models, which the reference leaves as independent roots).
    """
    import torch
    from hermespy.core import dB  # type: ignore
    from hermespy.simulation import SNR  # type: ignore

    from . import dropin

    if use_gpu_channel:
        dropin.enable(precision=precision)
    else:
        dropin.disable()
    dev = torch.device(device if device is not None else "cuda:0")
    stats = GridStatistics((len(snrs_db),), device=dev)
    scenario, tx, rx, link, evaluator = build(base_seed)

    def default_reseed(sc, tx_, rx_, link_, seed):
        sc.seed = seed
        link_.seed = seed + 1
        tx_.noise_model.seed = seed + 2
        rx_.noise_model.seed = seed + 3

    reseed = default_reseed if reseed is None else reseed
    mine = list(shard_drops(num_drops, rank, world_size))
    for cell, snr in enumerate(snrs_db):
        rx.noise_level = SNR(dB(snr), tx)
        tx_bits, rx_bits = [], []
        for d in mine:
            reseed(scenario, tx, rx, link, drop_seed(base_seed, cell, d, num_drops))
            scenario.drop()
            transmission, reception = evaluator._fetch_dsp_results()
            tx_bits.append(np.asarray(transmission.bits, dtype=np.uint8))
            rx_bits.append(np.asarray(reception.bits, dtype=np.uint8))
        if not mine:
            continue
        # ragged drops are padded; the kernel applies the reference's zero-padding rule (modem/evaluators.py:245-252)
        n = max(max(len(b) for b in tx_bits), max(len(b) for b in rx_bits))
        tb = np.zeros((len(mine), n), dtype=np.uint8)
        rb = np.zeros((len(mine), n), dtype=np.uint8)
        for i, (a, b) in enumerate(zip(tx_bits, rx_bits)):
            tb[i, :len(a)] = a
            rb[i, :len(b)] = b
        tl = torch.tensor([len(b) for b in tx_bits], dtype=torch.int32, device=dev)
        rl = torch.tensor([len(b) for b in rx_bits], dtype=torch.int32, device=dev)
        stats.accumulate_bits(torch.from_numpy(tb).to(dev), torch.from_numpy(rb).to(dev),
                              torch.full((len(mine),), cell, dtype=torch.int32, device=dev), tl, rl)
    if use_gpu_channel:
        dropin.disable()
    return stats
