#!/usr/bin/env python
"""EXPERIMENT (never run yet): replay the two-kernel C2 step from a CUDA graph and compare with eager launches.

    python tools/experiments/graph_step.py [--links 2048] [--steps 200]

Hypothesis (tools/experiments/NEXT.md item 1): at 8 ranks per box the eager launch path (ctypes call, plan, tensor-map
encode, cudaMallocAsync / cudaFreeAsync, 2 launches per step) competes with NCCL's service threads for host cores; a graph
replay is one driver call per step.  Everything the library does inside `hb_fading_propagate` is capturable: stream-ordered
allocation nodes, kernel nodes with by-value parameters (plans, delay tables, the CUtensorMap).  The per-kernel event
brackets of `hb_profile_begin` must stay off while capturing.
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--links", type=int, default=2048)
    ap.add_argument("--steps", type=int, default=200)
    args = ap.parse_args()
    import torch

    import bench
    from hermespy_b200.batch import sample_fading_links
    from hermespy_b200.kernels import FadingBatch, fading_propagate

    B, T, n = args.links, bench.C2["T"], 4
    blk = sample_fading_links(bench.make_channel(42), B, n, n, bench.C2["fs"])
    fb = FadingBatch.from_numpy(device="cuda", **blk)
    x = torch.view_as_complex(torch.randn((B, n, T, 2), device="cuda", dtype=torch.float32))
    y = torch.empty((B, n, T + blk["max_delay"]), dtype=torch.complex64, device="cuda")

    def timed(fn, label):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(args.steps):
            fn()
        host = time.perf_counter() - t0
        e1.record()
        torch.cuda.synchronize()
        print(f"{label}: {e0.elapsed_time(e1) / args.steps:.4f} ms/step on the device, host enqueue {1e3 * host / args.steps:.4f} ms/step")

    timed(lambda: fading_propagate(x, fb, out=y), "eager")
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fading_propagate(x, fb, out=y)  # warm-up on the capture stream (pool growth, function attributes)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        fading_propagate(x, fb, out=y)
    ref = y.clone()
    y.zero_()
    graph.replay()
    torch.cuda.synchronize()
    print("replay reproduces the eager result:", bool(torch.equal(y, ref)))
    timed(graph.replay, "graph")


if __name__ == "__main__":
    main()
