#!/usr/bin/env python
"""CPU-only: drops/s of the stock reference (P processes) against the batched runner driving the reference's NUMPY channel
through its lanes (channel patch off), same script.  What differs is the runner's own cost: lane hand-over, pickling, rounds.

    python tools/experiments/runner_overhead.py [--config c1] [--samples 200] [--workers 8] [--lanes 32] [--fake-channel]

``--fake-channel``: the round's device call returns zeros of the right shape at once (a free channel): the runner's ceiling.
"""
import argparse
import json
import os
import sys
import time

os.environ.setdefault("OMP_NUM_THREADS", "1")
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c1")
    ap.add_argument("--samples", type=int, default=200)
    ap.add_argument("--workers", type=int, default=os.cpu_count())
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--skip-reference", action="store_true")
    ap.add_argument("--fake-channel", action="store_true")
    ap.add_argument("--profile", action="store_true")
    args = ap.parse_args()
    import simulation_campaign as sc
    from oracle.refload import load_reference

    out = {}
    if not args.skip_reference:
        out["reference"] = sc.run_reference(args.config, args.samples, os.cpu_count())
        print(json.dumps(out["reference"]), flush=True)
    load_reference()
    from hermespy_b200 import config, runner

    lanes = args.lanes or 4 * args.workers
    config.batch_drops, config.workers = lanes, args.workers
    if args.fake_channel:
        import hermespy_b200.dropin as dropin

        dropin.patch_reference()  # requests are deferred to the round's device call ...

        def fake(requests, precision=None, device=None):  # ... which costs nothing here
            res = []
            for kind, blk, x, zero in requests:
                nrx = blk["spatial"].shape[0] if kind == "fading" else blk.num_rx
                md = blk["max_delay"] if kind == "fading" else blk.max_delay
                res.append(np.zeros((nrx, x.shape[1] + md), dtype=np.complex128))
            return res

        runner.propagate_requests = fake
    runner.patch_actor()
    points = sc.BUILDERS[args.config][1]
    sc._fast_polling(sc.BUILDERS[args.config][0](2, 5)).run()
    sim = sc._fast_polling(sc.BUILDERS[args.config][0](args.samples, 1000))
    t0 = time.perf_counter()
    if args.profile:
        import cProfile
        import pstats

        pr = cProfile.Profile()
        pr.enable()
    sim.run()
    dt = time.perf_counter() - t0
    if args.profile:
        pr.disable()
        pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
    out["runner"] = dict(lanes=lanes, workers=args.workers, drops=args.samples * points, seconds=dt,
                         drops_per_s=args.samples * points / dt, owner_seconds=runner.stats.get("seconds"),
                         fake_channel=args.fake_channel)
    print(json.dumps(out["runner"]))


if __name__ == "__main__":
    main()
