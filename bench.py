#!/usr/bin/env python
"""Benchmark of the channel hot path: propagated complex samples/s per link (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--links B]

A *step* is one pass of the hot path over one batch of ``--links`` synthetic links of config C2
(BASELINE.json configs[1]: 4x4 MIMO, 5G TDL-B, rms delay 300 ns, Doppler 100, medium antenna correlation,
T = 15 344 samples at 30.72 MHz).  The whole job (10 000 drops x 7 SNR points = 70 000 links, every one
re-realized) is 70000 / links steps of identical shape; K steps are timed.

* ``value``  device-resident throughput: inputs and parameters already in HBM, CUDA events around K steps,
             max over ranks, aggregate over all ranks (weak scaling: same links per GPU).
* ``e2e``    the same work through the host-buffer C-ABI (``hb_fading_propagate_host``): complex128 host
             buffers (the reference's SignalBlock dtype) in pinned memory, H2D + kernels + D2H inside the timed
             region.
* ``roofline`` achieved algorithmic HBM bytes/s of the dominant kernel (the K3+K4 kernel the planner picked: tdl_tma for
             this shape; accounting kind "tdl_poly") from per-launch CUDA events
             recorded by the library on the launch stream during the timed region, against MEASURED_PEAKS.json.
* ``cpu_baseline`` the numpy oracle (a restatement of the reference's CPU path) on the host cores, bounded sample.

``--impl reference`` times that CPU path alone (rank 0 only): the UNMODIFIED reference classes when the pip install
``baseline/_ref`` (tools/install_reference.py, git-ignored, travels with the snapshot) is present (kind
"reference"), else the numpy port under ``oracle/`` (kind "port").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# ---- workload: BASELINE.json configs[1] (C2) -------------------------------------------------------------
C2 = dict(
    name="C2: 4x4 MIMO OFDM frame (14 x (1024+72) = 15344 samples @30.72 MHz) over 5G TDL-B, rms_delay 300 ns, "
         "doppler 100, 3GPP medium antenna correlation; 10k drops x 7 SNR points = 70000 links",
    ntx=4, nrx=4, T=15344, fs=30.72e6, rms_delay=300e-9, doppler=100.0, total_links=70000,
)
UNIT = "complex samples/s per link direction"
METRIC = "propagated complex samples/s per link"


def make_channel(seed):
    import hermespy_b200.channel as MC

    return MC.TDL(MC.TDLType.B, rms_delay=C2["rms_delay"], doppler_frequency=C2["doppler"], seed=seed,
                  antenna_correlation=MC.StandardAntennaCorrelation(MC.CorrelationType.MEDIUM))


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---- clocks ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock / power / throttle reasons through NVML (ctypes calls release the GIL; no fork inside
    the timed region -- spawning nvidia-smi from a thread stalled the launch loop by milliseconds)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self._stop_evt = threading.Event()
        self.h = None
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def run(self):
        if self.h is None:
            return
        nv = self.nv
        while not self._stop_evt.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    reasons = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    reasons = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                power = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.rows.append((sm, reasons, power))
            except Exception:
                pass
            self._stop_evt.wait(0.005)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0,
                    "note": getattr(self, "err", "no NVML samples")}
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        seen = set()
        for _, r, _ in self.rows:
            for bit, name in names.items():
                if r & bit:
                    seen.add(name)
        return {"sm_mhz": float(np.median([r[0] for r in self.rows])), "sm_max_mhz": self.max_sm,
                "reasons": sorted(seen), "samples": len(self.rows), "power_w_max": max(r[2] for r in self.rows)}


# ---- CPU baseline: the reference's own code when its install travelled (baseline/_ref), else the oracle port ----
def cpu_kind():
    from oracle.refload import REFERENCE_ROOT, reference_available

    # /root/reference is never read at run time on the GPU box; only the pip install under baseline/_ref counts
    if reference_available() and os.path.realpath(REFERENCE_ROOT).startswith(os.path.realpath(os.path.join(ROOT, "baseline"))):
        return "reference"
    return "port"


def _cpu_worker_reference(args):
    """realize + sample + propagate of ONE C2 link per iteration through the unmodified reference classes
    (hermespy/channel/fading/fading.py:371-406 and its generator :293-343), complex128."""
    seed, n = args
    from oracle.refload import load_reference

    load_reference()
    from hermespy.channel import TDL, CorrelationType, StandardAntennaCorrelation, TDLType
    from hermespy.core import Signal
    from hermespy.simulation import SimulatedDevice, SimulatedIdealAntenna, SimulatedUniformArray

    ch = TDL(TDLType.B, rms_delay=C2["rms_delay"], doppler_frequency=C2["doppler"], seed=seed,
             antenna_correlation=StandardAntennaCorrelation(CorrelationType.MEDIUM))
    dev = lambda n_: SimulatedDevice(bandwidth=C2["fs"], oversampling_factor=1, carrier_frequency=3.5e9,
                                     antennas=SimulatedUniformArray(SimulatedIdealAntenna, 0.04, (n_, 1, 1)))
    tx, rx = dev(C2["ntx"]), dev(C2["nrx"])
    rng = np.random.default_rng(seed)
    done = 0
    for _ in range(n):
        s = ch.realize().sample(tx, rx)
        x = (rng.standard_normal((C2["ntx"], C2["T"])) + 1j * rng.standard_normal((C2["ntx"], C2["T"]))) / np.sqrt(2)
        y = s.propagate(Signal.Create(x, C2["fs"], 3.5e9))
        done += y.num_samples > 0
    return done


def _cpu_worker(args):
    if cpu_kind() == "reference":
        return _cpu_worker_reference(args)
    seed, n = args
    from oracle import fading_oracle as fo

    ch = make_channel(seed)
    rng = np.random.default_rng(seed)
    done = 0
    # same per-link work as the reference: realize + sample + propagate, complex128
    from hermespy_b200.core import SimulatedDevice, SimulatedIdealAntenna, SimulatedUniformArray

    dev = lambda n_: SimulatedDevice(bandwidth=C2["fs"], antennas=SimulatedUniformArray(SimulatedIdealAntenna, 0.04, (n_, 1, 1)))
    tx, rx = dev(C2["ntx"]), dev(C2["nrx"])
    for _ in range(n):
        s = ch.realize().sample(tx, rx)
        p = fo.FadingParams(power=s.power_profile, delay=s.delay_profile, los_gain=s.los_gains, nlos_gain=s.nlos_gains,
                            los_angle=s.los_angles, nlos_angle=s.nlos_angles, los_phase=s.los_phases,
                            nlos_phase=s.nlos_phases, los_doppler=s.los_doppler, nlos_doppler=s.nlos_doppler,
                            spatial=s.spatial_response, gain=s.gain, fs=C2["fs"], num_rx=C2["nrx"], num_tx=C2["ntx"])
        x = (rng.standard_normal((C2["ntx"], C2["T"])) + 1j * rng.standard_normal((C2["ntx"], C2["T"]))) / np.sqrt(2)
        y = fo.propagate(p, x)
        done += y.shape[1] > 0
    return done


class CpuPool:
    """`cores` worker processes (one per host core, as Ray's one actor per core, monte_carlo.py:176,621), warmed up
    once (imports, page-in); every pass maps a bounded number of links onto them."""

    def __init__(self, cores):
        import multiprocessing as mp

        self.cores = cores
        self.pool = mp.get_context("fork").Pool(cores)
        self.pool.map(_cpu_worker, [(1000 + i, 1) for i in range(cores)])
        self.seed = 0

    def run(self, links_per_core):
        t0 = time.perf_counter()
        done = self.pool.map(_cpu_worker, [(self.seed + i, links_per_core) for i in range(self.cores)])
        dt = time.perf_counter() - t0
        self.seed += self.cores
        return int(sum(done)), dt

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_reference_pass(links_per_core, cores):
    """One bounded pass of the CPU path on `cores` processes; returns (links, seconds)."""
    pool = CpuPool(cores)
    try:
        return pool.run(links_per_core)
    finally:
        pool.close()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    # a step = a bounded sample of the C2 workload: per_core links on every host core, sized so that K steps end
    # within a few minutes (one reference link costs ~0.7 s of one core)
    per_core = max(1, min(8, 240 // max(1, args.steps + args.warmup)))
    pool = CpuPool(cores)
    for _ in range(args.warmup):
        pool.run(per_core)
    t_total = 0.0
    links_total = 0
    for _ in range(args.steps):
        links, dt = pool.run(per_core)
        t_total += dt
        links_total += links
    pool.close()
    value = links_total * C2["T"] / t_total
    kind = cpu_kind()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": C2["name"], "links_per_step": per_core * cores},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{per_core * cores} links of C2 per step ({per_core} per core process), "
                                   "realize+sample+propagate in complex128 numpy (" + (
                                       "the unmodified reference classes from baseline/_ref" if kind == "reference"
                                       else "oracle port of fading.py:293-406") + ")"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---- GPU arm ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from hermespy_b200 import _lib
    from hermespy_b200.batch import sample_fading_links
    from hermespy_b200.kernels import FadingBatch, fading_propagate, fading_propagate_host

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        _lib.reserve_sms(args.reserve_sms)  # room for the statistics all-reduce next to the persistent kernels
    B, T, ntx, nrx = args.links, C2["T"], C2["ntx"], C2["nrx"]

    # ---- realize + sample the step's links on the host (numpy RNG, rank-dependent seed as simulation.py:220-223)
    ch = make_channel(42 + rank * 12345678)
    blk = sample_fading_links(ch, B, ntx, nrx, C2["fs"])
    fb = FadingBatch.from_numpy(device=dev, **blk)
    D = blk["max_delay"]
    gen = torch.Generator(device=dev)
    gen.manual_seed(rank)
    x = torch.view_as_complex(torch.randn((B, ntx, T, 2), device=dev, generator=gen, dtype=torch.float32) * (0.5 ** 0.5))
    y = torch.empty((B, nrx, T + D), dtype=torch.complex64, device=dev)
    from hermespy_b200.montecarlo import GridStatistics

    # evaluator statistics of the 7 SNR points of C2: the only data that crosses GPUs
    grid_stats = GridStatistics((7,), device=dev)

    pending = []
    counter = [0]

    def step():
        fading_propagate(x, fb, precision="f32", sos_mode=args.sos_mode, out=y)
        counter[0] += 1
        if world > 1 and counter[0] % args.stats_every == 0 and not os.environ.get("HB_BENCH_SKIP_COLLECTIVE"):
            # the only collective of the path: (sum, sum^2, count) + (bit errors, bits) of the grid cells, ONE packed
            # all-reduce every `stats_every` batches (the reference's collector polls its actors between batches, never
            # per drop: actors.py:219-225), enqueued behind this step's kernels and overlapped with the next step's.
            # Measured at 8 GPUs: a collective per step costs 0.078 ms of rank-synchronization jitter per 0.69 ms step.
            if pending:
                pending.pop()()
            pending.append(grid_stats.all_reduce(async_op=True))

    _, info = fading_propagate(x, fb, out=y, sos_mode=args.sos_mode, return_info=True)
    for _ in range(max(3, args.warmup)):
        step()
    if world > 1:
        grid_stats.all_reduce()  # warm-up of the collective too (communicator set-up happens on first use)
        counter[0] = 0
    # Everything with a variable host cost (NVML set-up takes milliseconds) happens BEFORE the barrier: ranks that leave
    # it must start their timed region together, or the rank that started first pays the others' start skew at the
    # first collective (measured at 8 GPUs: ~8 ms per run, 11 % of 100 steps).
    sampler = ClockSampler(local_rank)
    _lib.profile_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        step()
    if world > 1 and not os.environ.get("HB_BENCH_SKIP_COLLECTIVE"):
        if pending:
            pending.pop()()
        grid_stats.all_reduce()  # the final reduction of the campaign statistics, inside the timed region
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    prof = _lib.profile_end()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * B * T * args.steps / (ms_total * 1e-3)

    # ---- roofline of the dominant kernel --------------------------------------------------------------------
    peak, peak_src = measured_peaks()
    alg_bytes = 8.0 * B * (ntx * T + nrx * (T + D))  # complex64 in + out, SURVEY 8(d): 8 (Ntx + Nrx) B / sample
    k = prof["tdl_poly"] if prof["tdl_poly"]["launches"] else prof["tdl_direct"]
    kname = {"window": "tdl_window_kernel", "gather": "tdl_poly_kernel", "tma": "tdl_tma_kernel"}.get(info.get("variant"), "tdl_poly_kernel") \
        if prof["tdl_poly"]["launches"] else "tdl_direct_kernel"
    traffic = None  # dram bytes of one launch from the committed ncu --set full capture (same kernel, same launch shape)
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_ncu_traffic.json")) as fh:
            tr = json.load(fh)
        if tr["kernel"].startswith(kname) and B == 2048:
            traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
    except (OSError, KeyError, ValueError):
        traffic = None
    k_ms = k["ms"] / max(1, k["launches"])
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src, "kernel_ms": k_ms,
                "algorithmic_bytes_per_launch": alg_bytes,
                "kernel_share_of_step": k["ms"] / ms_total if ms_total > 0 else None,
                "other_kernels_ms": {n: v["ms"] / max(1, v["launches"]) for n, v in prof.items()
                                     if v["launches"] and n not in ("tdl_poly", "tdl_direct")}}
    launches = sum(v["launches"] for v in prof.values())

    # ---- end to end through the host-buffer C-ABI (complex128 host buffers, as the reference's SignalBlock) ----
    Be = min(B, args.e2e_links)
    if args.no_e2e:  # profiling runs only (tools/ncu_one.sh); such a line is not a bench result
        Be = 1
    xh = torch.empty((Be, ntx, T), dtype=torch.complex128).pin_memory()
    xh.copy_(x[:Be].to(torch.complex128).cpu())
    yh = torch.empty((Be, nrx, T + D), dtype=torch.complex128).pin_memory()
    sub = {k_: (v[:Be] if isinstance(v, np.ndarray) and v.ndim == 3 else v) for k_, v in blk.items()}

    def e2e_step():
        fading_propagate_host(xh.numpy(), out=yh.numpy(), precision="f32", **sub)

    e2e_step()
    e2e_step()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    n_e2e = max(1, min(args.steps, 5))
    for _ in range(n_e2e):
        e2e_step()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * Be * T * n_e2e / float(t.item())
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(xh.numel() * 16 + sum(
        sub[k_].nbytes for k_ in ("omega", "phi", "amp", "spatial"))), "d2h_bytes_per_step": int(yh.numel() * 16),
           "links_per_step": Be, "host_dtype": "complex128", "timing": "host wall clock around the blocking C-ABI call"}

    # ---- CPU baseline: rank 0, N = 1 only ---------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        per_core = 16 if cpu_kind() == "reference" else 4  # 10-30 s of host work
        links, dtc = cpu_reference_pass(per_core, cores)
        cpu = {"value": links * T / dtc, "unit": UNIT, "cores": cores, "kind": cpu_kind(),
               "sample": f"{links} links of C2 ({per_core} per core process, {cores} processes), realize+sample+propagate "
                         f"in complex128 numpy ({'unmodified reference from baseline/_ref' if cpu_kind() == 'reference' else 'oracle port'}); {dtc:.1f} s"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": C2["name"], "links_per_step_per_gpu": B, "steps_for_whole_job": int(np.ceil(
                C2["total_links"] / (B * world))), "l2_policy": f"inputs larger than L2 ({x.numel() * 8 / 1e6:.0f} MB in, "
                f"{y.numel() * 8 / 1e6:.0f} MB out per step)", "plan": info,
                "stats_allreduce": (f"NCCL, packed [7, 5] float64, every {args.stats_every} steps + once at the end, inside "
                                    "the timed region") if world > 1 else "none (one rank)"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--links", type=int, default=2048, help="links per step per GPU")
    ap.add_argument("--e2e-links", type=int, default=512, help="links per end-to-end step per GPU")
    ap.add_argument("--stats-every", type=int, default=10, help="N > 1: batches between statistics all-reduces")
    ap.add_argument("--reserve-sms", type=int, default=0, help="N > 1: SMs left to the NCCL all-reduce of the statistics")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs: shrink the end-to-end leg to one link")
    ap.add_argument("--sos-mode", default="auto", choices=["auto", "poly", "poly_window", "poly_gather", "poly_tma", "direct"],
                    help="kernel selection (profiling / A-B runs); the default lets the planner choose")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
