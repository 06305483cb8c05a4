#!/usr/bin/env python
"""BASELINE config C1 end to end: SISO RRC modem over 5G TDL-A, BER sweep dB(0, 2, ..., 20), num_samples = 1000.

    python tools/run_campaign_c1.py [--drops 1000] [--check-drops 100]
    python -m torch.distributed.run --nproc-per-node N ... tools/run_campaign_c1.py      (one rank per GPU)

The scenario is the reference's getting-started simulation (``_examples/getting_started/simulation.py``), built from
the unmodified reference in ``baseline/_ref``; the drop loop is ``hermespy_b200.campaign.run_ber_campaign`` (drops
sharded over the ranks, channel on the GPU in the float64 parity mode, statistics all-reduced over NCCL).  Rank 0 then
re-runs the first ``--check-drops`` drops of every SNR point in ONE process with the reference's numpy channel and
compares the bit-error counts of those drops: they must be identical.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build(seed):
    from hermespy.channel import TDL, TDLType
    from hermespy.modem import (BitErrorEvaluator, RootRaisedCosineWaveform, SimplexLink,
                                SingleCarrierLeastSquaresChannelEstimation, SingleCarrierZeroForcingChannelEqualization)
    from hermespy.simulation import SimulationScenario

    sc = SimulationScenario(seed=seed)
    tx = sc.new_device(oversampling_factor=4)
    rx = sc.new_device(oversampling_factor=4)
    sc.set_channel(tx, rx, TDL(TDLType.A))
    link = SimplexLink(seed=seed + 1)
    tx.transmitters.add(link)
    rx.receivers.add(link)
    link.waveform = RootRaisedCosineWaveform(num_preamble_symbols=10, num_data_symbols=100, roll_off=0.9)
    link.waveform.channel_estimation = SingleCarrierLeastSquaresChannelEstimation()
    link.waveform.channel_equalization = SingleCarrierZeroForcingChannelEqualization()
    return sc, tx, rx, link, BitErrorEvaluator(link, link)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--drops", type=int, default=1000)
    ap.add_argument("--check-drops", type=int, default=100)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "campaign_c1.json"))
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    from oracle.refload import load_reference, reference_available

    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    if not reference_available():
        if rank == 0:
            print(json.dumps({"unavailable": "no reference install (baseline/_ref)"}))
        return
    load_reference()
    from hermespy_b200.campaign import run_ber_campaign

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    snrs = list(range(0, 21, 2))
    t0 = time.perf_counter()
    stats = run_ber_campaign(build, snrs, args.drops, rank=rank, world_size=world, device=dev, precision="f64")
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    stats.all_reduce()
    dt = time.perf_counter() - t0
    if rank == 0:
        counts = stats.counts.cpu().numpy()
        out = {"config": "C1: SISO RRC (10 + 100 symbols, QPSK-class default) over TDL-A, oversampling 4, SNR 0..20 dB step 2",
               "n_gpus": world, "drops_per_point": args.drops, "seconds": dt, "drops_per_s": len(snrs) * args.drops / dt,
               "snr_db": snrs, "bit_errors": counts[:, 0].tolist(), "bits": counts[:, 1].tolist(),
               "ber": (counts[:, 0] / np.maximum(counts[:, 1], 1)).tolist()}
        # parity: the first check-drops drops of every point, one process, reference numpy channel vs GPU channel
        n = min(args.check_drops, args.drops)
        t1 = time.perf_counter()
        ref = run_ber_campaign(lambda s: build(s), snrs, n, device=dev, use_gpu_channel=False)
        t_ref = time.perf_counter() - t1
        # same drops on the GPU path need the same seeds: drop_seed depends on num_drops, so re-run with n drops
        gpu = run_ber_campaign(build, snrs, n, device=dev, precision="f64")
        rc, gc = ref.counts.cpu().numpy(), gpu.counts.cpu().numpy()
        out["parity"] = {"drops_per_point": n, "reference_bit_errors": rc[:, 0].tolist(), "gpu_f64_bit_errors": gc[:, 0].tolist(),
                         "bit_exact": bool(np.array_equal(rc, gc)), "reference_drops_per_s_one_process": len(snrs) * n / t_ref}
        if n == args.drops:  # the sharded, all-reduced campaign against the single-process reference run
            out["parity"]["sharded_campaign_equals_reference"] = bool(np.array_equal(rc, counts))
        print(json.dumps(out, indent=1))
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        json.dump(out, open(args.out, "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
