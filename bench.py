#!/usr/bin/env python
"""Benchmark of the channel hot path: propagated complex samples/s per link (BASELINE.json metric), every BASELINE config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--config C1|C2|C3|C4|C5] [--precision f32|f64] [--only]

A *step* is one pass of the hot path over one batch of synthetic links of a BASELINE.json config:

    C1  SISO RRC frame (500 samples @ 400 MHz) over 5G TDL-A, 11 SNR points x 1000 drops        (configs[0])
    C2  4x4 MIMO OFDM frame (15 344 samples @ 30.72 MHz) over 5G TDL-B, 10k drops x 7 SNR points  (configs[1], headline)
    C3  3GPP CDL-C, 32x4 UPA link, moving receiver, 2048-sample frames                            (configs[2])
    C4  64x64 Rician (TDL-D profile) with Kronecker correlation, tensor-core spatial GEMM        (configs[3])
    C5  COST259 typical urban, SISO, 2^20-sample frames                                          (configs[4])

The printed JSON line is the headline config (default C2, complex64 arithmetic).  Unless ``--only`` is given it also
carries ``"configs"``: one sub-record per other config plus the float64 parity mode of the headline config, each with
its own ``value`` / ``roofline`` / ``e2e`` / ``cpu_baseline`` / ``parity`` (north_star: "throughput on synthetic signals of
each named shape").  Per record:

* ``value``    device-resident throughput: inputs and parameters already in HBM, CUDA events around K steps, max over
               ranks, aggregate over all ranks (weak scaling: same links per GPU).
* ``e2e``      the same work through the host-buffer C-ABI (``hb_fading_propagate_host`` / ``hb_cdl_propagate_host``):
               complex128 pinned host buffers (the reference's SignalBlock dtype), H2D + kernels + D2H inside the timed region.
* ``roofline`` the dominant kernel from per-launch CUDA events recorded by the library on the launch stream during the
               timed region: algorithmic HBM bytes/s against MEASURED_PEAKS.json (HBM-bound kernels), or flop/s against the
               FP32 / tensor pipe peak for the CDL contraction.
* ``parity``   the gate beside the timing: relative L2 of the first links of the SAME batch against the numpy oracle.
* ``cpu_baseline`` the reference's CPU path on the host cores, bounded sample: the UNMODIFIED reference classes from
               ``baseline/_ref`` (kind "reference"), or the oracle port where the reference cannot run the shape (C4).

``--impl reference`` times that CPU path alone (rank 0 only) for ``--config``.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

# One BLAS / OpenMP thread per process, set BEFORE numpy loads its BLAS: every CPU leg of this file (cpu_baseline, the
# reference arm, the reference side of simulation_api) runs one worker PROCESS per host core, like Ray's one actor per core
# (monte_carlo.py:176,621).  Left at the default, 16 processes x 16 BLAS threads oversubscribe the box and the reference's
# matmul-heavy CDL code runs 50x slower than it should (measured: UMa 0.53 against 26 drops/s) -- unfair to the reference.
for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ.setdefault(_v, "1")

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "complex samples/s per link direction"
METRIC = "propagated complex samples/s per link"

# ---- workloads: BASELINE.json configs -----------------------------------------------------------------
CONFIGS = {
    "C1": dict(id="C1", kind="fading", ntx=1, nrx=1, T=500, fs=4e8, links=11000, e2e_links=11000, total_links=11000,
               oracle_links=2, cpu_links_per_core=64,
               name="C1: SISO RRC frame (500 samples @ 400 MHz) over 5G TDL-A, doppler 100; 11 SNR points x 1000 drops "
                    "= 11000 links"),
    "C2": dict(id="C2", kind="fading", ntx=4, nrx=4, T=15344, fs=30.72e6, links=2048, e2e_links=512, total_links=70000,
               oracle_links=1, cpu_links_per_core=16,
               name="C2: 4x4 MIMO OFDM frame (14 x (1024+72) = 15344 samples @30.72 MHz) over 5G TDL-B, rms_delay 300 ns, "
                    "doppler 100, 3GPP medium antenna correlation; 10k drops x 7 SNR points = 70000 links"),
    "C3": dict(id="C3", kind="cdl", ntx=32, nrx=4, T=2048, fs=30.72e6, fc=3.5e9, links=4096, e2e_links=512,
               total_links=100000, distinct=128, oracle_links=1, cpu_links_per_core=1,
               name="C3: 3GPP CDL-C (rms delay 300 ns), 32 tx (8x4 UPA) x 4 rx (2x2) ideal elements, fc 3.5 GHz, receiver "
                    "moving at (10, -3, 0) m/s, 2048-sample frames @30.72 MHz; 100k drops"),
    "C4": dict(id="C4", kind="fading", ntx=64, nrx=64, T=16384, fs=30.72e6, links=256, e2e_links=16, total_links=256,
               oracle_links=1, cpu_links_per_core=1,
               name="C4: 64x64 Rician sum-of-sinusoids (TDL-D profile, rms_delay 300 ns, doppler 100) with exponential "
                    "Kronecker correlation rho = 0.7, 16384-sample frames; 256 links"),
    "C5": dict(id="C5", kind="fading", ntx=1, nrx=1, T=1 << 20, fs=30.72e6, links=64, e2e_links=16, total_links=64,
               oracle_links=1, oracle_samples=1 << 16, cpu_links_per_core=1,
               name="C5: COST259 typical urban, doppler 50, SISO, 2^20-sample frames @30.72 MHz (NEAREST delays); 64 links"),
}


def make_channel(cfg, M, seed):
    """The config's channel from a module exposing the reference's public names (``hermespy.channel`` for the CPU
    arm, ``hermespy_b200.channel`` for the GPU arm): one source text drives both implementations."""
    cid = cfg["id"]
    if cid == "C1":
        return M.TDL(M.TDLType.A, doppler_frequency=100.0, seed=seed)
    if cid == "C2":
        return M.TDL(M.TDLType.B, rms_delay=300e-9, doppler_frequency=100, seed=seed,
                     antenna_correlation=M.StandardAntennaCorrelation(M.CorrelationType.MEDIUM))
    if cid == "C3":
        return M.CDL(M.CDLType.C, 300e-9, seed=seed)
    if cid == "C4":
        n = cfg["ntx"]
        R = 0.7 ** np.abs(np.subtract.outer(np.arange(n), np.arange(n))).astype(complex)
        return M.TDL(M.TDLType.D, rms_delay=300e-9, doppler_frequency=100, seed=seed, max_antennas=n,
                     antenna_correlation=M.CustomAntennaCorrelation(R))
    if cid == "C5":
        return M.Cost259(M.Cost259Type.URBAN, doppler_frequency=50, seed=seed)
    raise KeyError(cid)


def make_devices(cfg, S, T):
    """(tx, rx) devices of the config from a module exposing SimulatedDevice & co; ``T`` = Transformation class."""
    fs = cfg["fs"]
    if cfg["kind"] == "cdl":
        fc = cfg["fc"]
        lam = 299792458.0 / fc
        tx = S.SimulatedDevice(bandwidth=fs, oversampling_factor=1, carrier_frequency=fc,
                               antennas=S.SimulatedUniformArray(S.SimulatedIdealAntenna, lam / 2, (8, 4, 1)),
                               pose=T.From_Translation(np.array([0.0, 0.0, 25.0])))
        rx = S.SimulatedDevice(bandwidth=fs, oversampling_factor=1, carrier_frequency=fc,
                               antennas=S.SimulatedUniformArray(S.SimulatedIdealAntenna, lam / 2, (2, 2, 1)),
                               pose=T.From_Translation(np.array([100.0, 20.0, 1.5])), velocity=np.array([10.0, -3.0, 0.0]))
        return tx, rx
    dev = lambda n: S.SimulatedDevice(bandwidth=fs, oversampling_factor=1, carrier_frequency=3.5e9,
                                      antennas=S.SimulatedUniformArray(S.SimulatedIdealAntenna, 0.04, (n, 1, 1)))
    return dev(cfg["ntx"]), dev(cfg["nrx"])


def public_config(cfg, precision):
    """The ``config`` object of the JSON line: identical for both arms (the driver compares them)."""
    return {"workload": cfg["name"], "id": cfg["id"], "num_tx": cfg["ntx"], "num_rx": cfg["nrx"],
            "samples_per_frame": cfg["T"], "sampling_rate_hz": cfg["fs"], "gpu_arithmetic": precision}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return float(d["hbm_gbs"]), float(d.get("sm_max_mhz", 1965.0)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


def _measured_bf16_tflops():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["bf16_tflops"]), "measured burst, MEASURED_PEAKS.json"
    except Exception:
        return 2250.0, "fallback: nominal dense bf16 (B200_PROFILING.md)"


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


# ---- CPU arm: the reference's own code when its install travelled (baseline/_ref), else the oracle port ----------
def cpu_kind(cfg):
    from oracle.refload import REFERENCE_ROOT, reference_available

    if cfg["id"] == "C4":
        return "port"  # the reference caps fading at 10x10 antennas: float64 restatement, labelled
    # /root/reference is never read at run time on the GPU box; only the pip install under baseline/_ref counts
    if reference_available() and os.path.realpath(REFERENCE_ROOT).startswith(os.path.realpath(os.path.join(ROOT, "baseline"))):
        return "reference"
    return "port"


def _signal(rng, ntx, T):
    return (rng.standard_normal((ntx, T)) + 1j * rng.standard_normal((ntx, T))) / np.sqrt(2)


def _cpu_worker(args):
    """realize + sample + propagate of ``n`` links of a config, complex128, one host process."""
    cid, seed, n = args
    cfg = CONFIGS[cid]
    rng = np.random.default_rng(seed)
    done = 0
    if cpu_kind(cfg) == "reference":
        # the unmodified reference classes (fading.py:371-406 + generator :293-343; cluster_delay_lines.py:409-558)
        from oracle.refload import load_reference

        load_reference()
        import hermespy.channel as RC
        import hermespy.simulation as RS
        from hermespy.core import Signal, Transformation

        ch = make_channel(cfg, RC, seed)
        tx, rx = make_devices(cfg, RS, Transformation)
        fc = cfg.get("fc", 3.5e9)
        for _ in range(n):
            s = ch.realize().sample(tx, rx)
            y = s.propagate(Signal.Create(_signal(rng, cfg["ntx"], cfg["T"]), cfg["fs"], fc))
            done += y.num_samples > 0
        return done
    # oracle port (float64 numpy restatement) through the host mirror classes
    import hermespy_b200.channel as MC
    import hermespy_b200.core as MS
    from oracle import cdl_oracle as co
    from oracle import fading_oracle as fo

    ch = make_channel(cfg, MC, seed)
    tx, rx = make_devices(cfg, MS, MS.Transformation)
    for _ in range(n):
        s = ch.realize().sample(tx, rx)
        x = _signal(rng, cfg["ntx"], cfg["T"])
        if cfg["kind"] == "cdl":
            from tests.test_cdl_golden import oracle_params

            y = co.propagate(oracle_params(s), x)
        else:
            y = fo.propagate(_fading_oracle_params(s, cfg), x)
        done += y.shape[1] > 0
    return done


def _fading_oracle_params(s, cfg):
    from oracle import fading_oracle as fo

    return fo.FadingParams(power=s.power_profile, delay=s.delay_profile, los_gain=s.los_gains, nlos_gain=s.nlos_gains,
                           los_angle=s.los_angles, nlos_angle=s.nlos_angles, los_phase=s.los_phases,
                           nlos_phase=s.nlos_phases, los_doppler=s.los_doppler, nlos_doppler=s.nlos_doppler,
                           spatial=s.spatial_response, gain=s.gain, fs=cfg["fs"], num_rx=cfg["nrx"], num_tx=cfg["ntx"])


class CpuPool:
    """`cores` worker processes (one per host core, as Ray's one actor per core, monte_carlo.py:176,621), warmed up
    once (imports, page-in); every pass maps a bounded number of links onto them."""

    def __init__(self, cid, cores):
        import multiprocessing as mp

        self.cid, self.cores = cid, cores
        self.pool = mp.get_context("fork").Pool(cores)
        self.pool.map(_cpu_worker, [(cid, 1000 + i, 1 if CONFIGS[cid]["T"] <= 20000 else 0) for i in range(cores)])
        self.seed = 0

    def run(self, links_per_core):
        t0 = time.perf_counter()
        done = self.pool.map(_cpu_worker, [(self.cid, self.seed + i, links_per_core) for i in range(self.cores)])
        dt = time.perf_counter() - t0
        self.seed += self.cores
        return int(sum(done)), dt

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_sample_text(cfg, links, per_core, cores, dt=None):
    kind = cpu_kind(cfg)
    what = ("the unmodified reference classes from baseline/_ref" if kind == "reference" else
            "float64 numpy port under oracle/ (the reference cannot run this shape)" if cfg["id"] == "C4" else "oracle port")
    return (f"{links} links of {cfg['id']} ({per_core} per core process, {cores} processes), realize+sample+propagate in "
            f"complex128 numpy ({what})" + (f"; {dt:.1f} s" if dt is not None else ""))


def cpu_baseline(cfg):
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    per_core = cfg["cpu_links_per_core"]
    pool = CpuPool(cfg["id"], cores)
    try:
        links, dt = pool.run(per_core)
    finally:
        pool.close()
    return {"value": links * cfg["T"] / dt, "unit": UNIT, "cores": cores, "kind": cpu_kind(cfg),
            "sample": cpu_sample_text(cfg, links, per_core, cores, dt)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = CONFIGS[args.config]
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    # a step = a bounded sample of the workload: per_core links on every host core, sized so that K steps end within
    # a few minutes (one C2 reference link costs ~0.7 s of one core, a C3 link 4.5 s, a C5 link ~8 s)
    budget = {"C1": 2048, "C2": 240, "C3": 40, "C4": 16, "C5": 20}[cfg["id"]]
    per_core = max(1, min(cfg["cpu_links_per_core"], budget // max(1, args.steps + args.warmup)))
    pool = CpuPool(cfg["id"], cores)
    for _ in range(args.warmup):
        pool.run(per_core)
    t_total, links_total = 0.0, 0
    for _ in range(args.steps):
        links, dt = pool.run(per_core)
        t_total += dt
        links_total += links
    pool.close()
    value = links_total * cfg["T"] / t_total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": public_config(cfg, args.precision), "run": {"links_per_step": per_core * cores},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": cpu_kind(cfg),
                         "sample": cpu_sample_text(cfg, per_core * cores, per_core, cores) + " per step"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---- GPU arm ------------------------------------------------------------------------------------------------
class Workload:
    """One config on one rank: device-resident step, end-to-end step, algorithmic work, parity gate."""

    def __init__(self, cfg, precision, rank, dev, links, e2e_links, sos_mode="auto", sinc=False):
        import torch

        self.cfg, self.precision, self.dev, self.rank, self.sinc = cfg, precision, dev, rank, sinc
        self.B, self.T, self.ntx, self.nrx = links, cfg["T"], cfg["ntx"], cfg["nrx"]
        self.sos_mode = sos_mode
        self.cdtype = torch.complex128 if precision == "f64" else torch.complex64
        gen = torch.Generator(device=dev)
        gen.manual_seed(rank)
        rdtype = torch.float64 if precision == "f64" else torch.float32
        self.x = torch.view_as_complex(torch.randn((self.B, self.ntx, self.T, 2), device=dev, generator=gen, dtype=rdtype)
                                       * (0.5 ** 0.5))
        t0 = time.perf_counter()
        self.device_sampling_s = None
        self._sample_links(42 + rank * 12345678)  # rank-dependent seed as simulation.py:220-223
        self.host_sampling_s = time.perf_counter() - t0 - (self.device_sampling_s or 0.0)
        self.y = torch.empty((self.B, self.nrx, self.T + self.D), dtype=self.cdtype, device=dev)
        self.Be = max(1, min(self.B, e2e_links))

    # -- realize + sample the step's links on the host (numpy RNG, reference draw order) ---------------------------
    def _sample_links(self, seed):
        import hermespy_b200.channel as MC
        import hermespy_b200.core as MS

        cfg = self.cfg
        ch = make_channel(cfg, MC, seed)
        if cfg["kind"] == "fading":
            from hermespy_b200.batch import sample_fading_links, sample_fading_links_device
            from hermespy_b200.kernels import FadingBatch

            # the timed batch is realized by the DEVICE sampler (hb_fading_sample + hb_kron_mix: host draws the normals
            # only); the host sampler below supplies the parameter block of the end-to-end leg and of the parity oracle
            # from an identically seeded channel -- the two agree to a few ulp (tests/test_sampling_gpu.py)
            t1 = time.perf_counter()
            fb_dev = None if self.sinc else sample_fading_links_device(make_channel(cfg, MC, seed), self.B, self.ntx, self.nrx,
                                                                       cfg["fs"], device=self.dev)
            self.device_sampling_s = time.perf_counter() - t1
            self.blk = sample_fading_links(ch, self.B, self.ntx, self.nrx, cfg["fs"])
            if self.sinc:
                # fractional delays (extension): every tap at its true delay -> 12 windowed-sinc taps at integer delays;
                # the kernels are the NEAREST ones (equal delays merge), the tap table carries the interpolation
                from hermespy_b200.kernels import sinc_expand

                self.blk.update(sinc_expand(ch.delays, cfg["fs"], self.blk["omega"], self.blk["phi"], self.blk["amp"]))
            self.fb = fb_dev if fb_dev is not None else FadingBatch.from_numpy(device=self.dev, **self.blk)
            self.D = self.blk["max_delay"]
            self.channel = ch
        else:
            from hermespy_b200.kernels import CdlBlock, CdlDeviceBlock

            tx, rx = make_devices(cfg, MS, MS.Transformation)
            distinct = min(self.B, cfg["distinct"])  # host sampling of CDL links costs 10 ms each: tile a distinct set
            self.samples = [ch.realize().sample(tx, rx) for _ in range(distinct)]
            blocks = [s.kernel_block() for s in self.samples]
            self.blk = CdlBlock.stack([blocks[i % distinct] for i in range(self.B)])
            self.dblk = CdlDeviceBlock(self.blk, device=self.dev)
            self.D = self.blk.max_delay

    def step(self, return_info=False):
        if self.cfg["kind"] == "fading":
            from hermespy_b200.kernels import fading_propagate

            return fading_propagate(self.x, self.fb, precision=self.precision, sos_mode=self.sos_mode, out=self.y,
                                    return_info=return_info)
        from hermespy_b200.kernels import cdl_propagate

        return cdl_propagate(self.x, self.dblk, precision=self.precision, out=self.y, return_info=return_info,
                             variant=os.environ.get("HB_BENCH_CDL_VARIANT", "auto"))  # experiments: gather | umma | umma_bf16

    # -- algorithmic work ------------------------------------------------------------------------------
    def algorithmic_bytes(self):
        esz = 16.0 if self.precision == "f64" else 8.0
        return esz * self.B * (self.ntx * self.T + self.nrx * (self.T + self.D))

    def dominant(self, prof, info):
        """(accounting kind, kernel name, bound) of the step's dominant kernel."""
        if self.cfg["kind"] == "cdl":
            if info.get("variant") in ("umma", "umma_bf16"):
                return "cdl_propagate", "cdl_umma_kernel" if info["variant"] == "umma" else "cdl_umma_bf16_kernel", "tensor"
            return "cdl_propagate", ("cdl_poly_kernel" if info.get("mode") == "poly" else "cdl_direct_f64_kernel"), "fp32"
        if info.get("variant") == "fused":
            return "spatial_gemm", "fused_gemm_tdl_kernel", "hbm"
        if prof["tdl_poly"]["launches"]:
            if prof["spatial_gemm"]["launches"] and prof["spatial_gemm"]["ms"] > prof["tdl_poly"]["ms"]:
                return "spatial_gemm", "spatial_gemm_3xtf32_kernel", "hbm"
            name = {"window": "tdl_window_kernel", "gather": "tdl_poly_kernel", "tma": "tdl_tma_kernel", "siso": "tdl_siso_kernel"}.get(
                info.get("variant"), "tdl_poly_kernel")
            if self.precision == "f64":
                name = "tdl_poly64_kernel"  # the float64 Taylor path (fading_poly64.cuh)
            return "tdl_poly", name, "hbm"
        return "tdl_direct", "tdl_direct_kernel", "hbm"

    # -- end to end through the host-buffer C-ABI (complex128 host buffers, as the reference's SignalBlock) -----------
    def e2e_prepare(self):
        import torch

        Be = self.Be
        self.xh = torch.empty((Be, self.ntx, self.T), dtype=torch.complex128).pin_memory()
        self.xh.copy_(self.x[:Be].to(torch.complex128).cpu())
        self.yh = torch.empty((Be, self.nrx, self.T + self.D), dtype=torch.complex128).pin_memory()
        if self.cfg["kind"] == "fading":
            self.sub = {k: (v[:Be] if isinstance(v, np.ndarray) and v.ndim == 3 else v) for k, v in self.blk.items()}
            self.h2d = int(self.xh.numel() * 16 + sum(self.sub[k].nbytes for k in ("omega", "phi", "amp", "spatial")))
        else:
            from hermespy_b200.kernels import CdlBlock

            blocks = [s.kernel_block() for s in self.samples]
            self.sub = CdlBlock.stack([blocks[i % len(blocks)] for i in range(Be)])
            self.h2d = int(self.xh.numel() * 16 + sum(getattr(self.sub, k).nbytes for k in CdlBlock.ARRAYS))
        self.d2h = int(self.yh.numel() * 16)

    def e2e_step(self):
        if self.cfg["kind"] == "fading":
            from hermespy_b200.kernels import fading_propagate_host

            fading_propagate_host(self.xh.numpy(), out=self.yh.numpy(), precision=self.precision, device=self.dev.index,
                                  **self.sub)
        else:
            from hermespy_b200.kernels import cdl_propagate_host

            cdl_propagate_host(self.xh.numpy(), self.sub, out=self.yh.numpy(), precision=self.precision,
                               device=self.dev.index)

    # -- parity gate: the first links of THIS batch against the numpy oracle ------------------------------------------
    def parity(self):
        cfg = self.cfg
        n = cfg["oracle_links"]
        Tn = min(self.T, cfg.get("oracle_samples", self.T) // (4 if self.sinc else 1))  # causal channel: a prefix of x determines the same prefix of y
        y = self.y[:n, :, :Tn].cpu().numpy().astype(np.complex128)
        x = self.x[:n, :, :Tn].cpu().numpy().astype(np.complex128)
        worst = 0.0
        for b in range(n):
            if cfg["kind"] == "fading":
                blk = self.blk
                nn = np.arange(Tn)
                K = blk["omega"].shape[2]
                ampk = blk["amp"][b][:, [0] + [1] * (K - 1), None]
                h = (ampk * np.exp(1j * (blk["omega"][b][:, :, None] * nn + blk["phi"][b][:, :, None]))).sum(1)  # fading.py:326-342
                z = np.zeros((self.ntx, Tn + self.D), complex)
                for l, d in enumerate(blk["tap_delay"]):
                    z[:, d: d + Tn] += x[b] * h[l]  # fading.py:385-390
                S = blk["spatial"][b]
                if "r_rx" in blk:  # large arrays: the Kronecker mix runs on the device (K2), fading.py:480-489
                    S = blk["r_rx"] @ S @ blk["r_tx"]
                ref = (S @ z)[:, :Tn]  # fading.py:393
            else:
                from oracle import cdl_oracle as co
                from tests.test_cdl_golden import oracle_params

                ref = co.propagate(oracle_params(self.samples[b % len(self.samples)]), x[b])[:, :Tn]
            worst = max(worst, float(np.linalg.norm(y[b] - ref) / np.linalg.norm(ref)))
        tol = 1e-5 if self.precision == "f32" else (1e-10 if cfg["kind"] == "cdl" else 1e-12)
        return {"rel_l2_vs_oracle": worst, "tolerance": tol, "links_checked": n, "samples_checked": Tn,
                "ok": bool(worst <= tol), "oracle": "numpy float64 restatement (oracle/), same links and signal as the timed batch"}


def measure(cfg, precision, args, world, rank, dev, steps, with_cpu, stats_allreduce, sinc=False):
    """One config on this rank (all ranks call it together): returns the record (rank 0) or None."""
    import torch
    import torch.distributed as dist

    from hermespy_b200 import _lib
    from hermespy_b200.montecarlo import GridStatistics

    links = args.links if (args.links and cfg["id"] == args.config) else cfg["links"]
    if precision == "f64":
        links = max(1, links // 8)  # the FP64 direct kernels are ~10x slower per link: keep the step a fraction of a second
    wl = Workload(cfg, precision, rank, dev, links, min(args.e2e_links or cfg["e2e_links"], cfg["e2e_links"]), args.sos_mode,
                  sinc=sinc)
    B, T = wl.B, wl.T
    grid_stats = GridStatistics((7,), device=dev)  # evaluator statistics: the only data that crosses GPUs
    pending, counter = [], [0]

    def step():
        wl.step()
        counter[0] += 1
        if stats_allreduce and counter[0] % args.stats_every == 0:
            # the only collective of the path: (sum, sum^2, count) + (bit errors, bits) of the grid cells, ONE packed
            # all-reduce every `stats_every` batches, enqueued behind this step's kernels, overlapped with the next step's
            if pending:
                pending.pop()()
            pending.append(grid_stats.all_reduce(async_op=True))

    _, info = wl.step(return_info=True)
    for _ in range(max(3, args.warmup)):
        step()
    if stats_allreduce:
        grid_stats.all_reduce()
        counter[0] = 0
    # everything with a variable host cost happens BEFORE the barrier (start skew is paid at the first collective)
    sampler = ClockSampler(dev.index)
    _lib.profile_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        step()
    if stats_allreduce:
        if pending:
            pending.pop()()
        grid_stats.all_reduce()  # the final reduction of the campaign statistics, inside the timed region
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    prof = _lib.profile_end()
    clocks = sampler.stop()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * B * T * steps / (ms_total * 1e-3)

    # ---- roofline of the dominant kernel ---------------------------------------------------------------------------
    hbm_peak, sm_max_mhz, peak_src = measured_peaks()
    kind, kname, bound = wl.dominant(prof, info)
    k = prof[kind]
    k_ms = k["ms"] / max(1, k["launches"]) * (k["launches"] / max(1, steps))  # per step (a step may launch it per chunk)
    alg_bytes = wl.algorithmic_bytes()
    others = {n: v["ms"] / steps for n, v in prof.items() if v["launches"] and n != kind}
    if bound == "hbm":
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
        roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                    "frac": achieved / hbm_peak, "traffic": None if sinc else _traffic(cfg, kname, B), "peak_source": peak_src}
    elif bound == "tensor":
        # CDL contraction on tcgen05 (cdl_umma_kernel): ALGORITHMIC flops = 8 P G Nrx Ntx per output sample; the kernel runs
        # them as 3xTF32 (three TF32 products per FP32 product), so the tensor pipe executes 3x that.  Peak: dense TF32 =
        # half the measured cuBLAS bf16 burst figure of MEASURED_PEAKS.json (the kernel is timed alone per launch).
        flops = 8.0 * info["poly_order"] * info["num_groups"] * wl.nrx * wl.ntx * B * (T + wl.D)
        bf16_tf = _measured_bf16_tflops()
        peak_tf = 0.5 * bf16_tf[0]
        achieved = flops / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
        roofline = {"bound": "tensor", "kernel": kname, "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": achieved / peak_tf, "traffic": _traffic(cfg, kname, B),
                    "peak_source": f"dense TF32 = 0.5 x bf16_tflops ({bf16_tf[1]})",
                    "tensor_flops_executed_per_algorithmic_flop": 3.0,
                    "tensor_pipe_frac_3xtf32": 3.0 * achieved / peak_tf,
                    "binding_unit": "shared-memory operand reads of the MMAs (128 B/clk/SM; N = 64 columns per instruction): "
                                    "88 clk per (delay group, 4 antennas, 128 samples) pair measured by tools/microbench/umma_shift_probe.cu",
                    "hbm_frac_of_step": alg_bytes / (ms_total / steps * 1e-3) / 1e9 / hbm_peak}
    else:
        # CDL contraction: 8 P G Nrx Ntx real flop per output sample, on the FP32 pipe
        flops = 8.0 * info["poly_order"] * info["num_groups"] * wl.nrx * wl.ntx * B * (T + wl.D) if info.get("mode") == "poly" else 0.0
        peak_tf = 148 * 128 * 2 * sm_max_mhz * 1e6 / 1e12
        achieved = flops / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
        roofline = {"bound": "fp32", "kernel": kname, "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": achieved / peak_tf, "traffic": None,
                    "peak_source": f"148 SMs x 128 FMA lanes x 2 x {sm_max_mhz:.0f} MHz (sm_max_mhz of MEASURED_PEAKS.json)",
                    "hbm_frac_of_step": alg_bytes / (ms_total / steps * 1e-3) / 1e9 / hbm_peak}
    roofline.update({"kernel_ms_per_step": k_ms, "algorithmic_bytes_per_step": alg_bytes,
                     "kernel_share_of_step": k["ms"] / ms_total if ms_total > 0 else None,
                     "step_frac_of_hbm_peak": alg_bytes / (ms_total / steps * 1e-3) / 1e9 / hbm_peak,
                     "other_kernels_ms_per_step": others})
    launches = sum(v["launches"] for v in prof.values())

    # ---- parity gate beside the timing ---------------------------------------------------------------------------
    parity = wl.parity() if rank == 0 else None

    # ---- end to end -------------------------------------------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        wl.e2e_prepare()
        wl.e2e_step()
        wl.e2e_step()
        if world > 1:
            dist.barrier()
        n_e2e = max(1, min(steps, 5))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            wl.e2e_step()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": world * wl.Be * T * n_e2e / float(t.item()), "unit": UNIT, "h2d_bytes_per_step": wl.h2d,
               "d2h_bytes_per_step": wl.d2h, "links_per_step": wl.Be, "host_dtype": "complex128",
               "timing": "host wall clock around the blocking C-ABI call, max over ranks"}

    cpu = cpu_baseline(cfg) if (with_cpu and rank == 0 and world == 1) else None
    host_sampling_s, device_sampling_s = wl.host_sampling_s, wl.device_sampling_s
    del wl
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": precision, "data": "synthetic", "config": public_config(cfg, precision),
        "delay_interpolation": ("SINC: 12-tap Kaiser-windowed sinc per channel tap at its true delay (extension with its own "
                                "oracle; the reference rounds delays)") if sinc else "NEAREST (the reference's rounding, parity mode)",
        "run": {"links_per_step_per_gpu": B, "steps_for_whole_job": int(np.ceil(cfg["total_links"] / (B * world))),
                "l2_policy": f"inputs larger than L2 ({alg_bytes / 1e6:.0f} MB in + out per step)" if alg_bytes > 2 * 126e6
                else f"in + out {alg_bytes / 1e6:.0f} MB per step: L2-resident between steps (the whole job of this config is one step)",
                "plan": info, "host_sampling_s": round(host_sampling_s, 3),
                "device_sampling_s": None if device_sampling_s is None else round(device_sampling_s, 4),
                "stats_allreduce": (f"NCCL, packed [7, 5] float64, every {args.stats_every} steps + once at the end, inside "
                                    "the timed region") if stats_allreduce else "none"},
        "roofline": roofline, "parity": parity, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
    }


def _traffic(cfg, kname, B):
    """dram bytes of one launch from the committed ncu --set full captures (same kernel, same launch shape), or None."""
    for name in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as fh:
                tr = json.load(fh)
            rows = tr if isinstance(tr, list) else [tr]
            for r in rows:
                if r.get("config", "C2") == cfg["id"] and r["kernel"].startswith(kname) and r.get("links", 2048) == B:
                    return r["dram_bytes_read"] + r["dram_bytes_write"]
        except (OSError, KeyError, ValueError):
            continue
    return None


#: set when a leg had to be abandoned on its time limit: threads / helper processes of the hung run may still be alive, so the
#: process leaves through os._exit once the JSON line is out (an interpreter exit would join them)
_FORCE_EXIT = False
SIMULATION_TIME_LIMIT_S = 180  # per script and arm; the three scripts take 2 .. 12 s each


def _time_limited(seconds, fn, *a, **k):
    """``fn(*a, **k)`` in the main thread under a SIGALRM limit: ``TimeoutError`` instead of a bench run that never prints
    its line.  (The campaign engine polls from the main thread, so the exception surfaces there.)"""
    import signal

    def on_alarm(signum, frame):
        raise TimeoutError(f"exceeded its {seconds} s limit")

    old = signal.signal(signal.SIGALRM, on_alarm)
    signal.alarm(int(seconds))
    try:
        return fn(*a, **k)
    finally:
        signal.alarm(0)
        signal.signal(signal.SIGALRM, old)


def measure_simulation(args, world, rank, dev):
    """End to end through the UNMODIFIED ``Simulation.run()`` API (north_star): drops/s of the reference's own scripts with
    the channel on the GPU and the batched drop runner (hermespy_b200/runner.py).  Three scripts: BASELINE config C1 (SISO RRC
    over TDL-A, 11 SNR points), the C2 frame through a modem the reference can demodulate (2x1 Alamouti OFDM, 1024
    subcarriers, ideal CSI, TDL-B) and the same link over the stochastic 3GPP UMa scenario (heterogeneous CDL batches).  N = 1: helper processes on the host cores, and the stock reference on all host cores
    beside it.  N > 1: every rank runs its own campaign share with in-process lanes (no fork next to NCCL)."""
    global _FORCE_EXIT
    import torch
    import torch.distributed as dist

    from oracle.refload import load_reference, reference_available

    if not reference_available():
        return {"unavailable": "no reference install (baseline/_ref)"}
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import simulation_campaign as sc

    load_reference()
    from hermespy_b200 import config as hb_config

    hb_config.device = dev.index
    cores = os.cpu_count() or 1
    workers = max(1, cores - 1) if world == 1 else 0
    out = {"api": "hermespy.simulation.Simulation.run() (unmodified script), dropin.enable(precision='f64', batch_drops, workers)",
           "host_cores": cores, "helper_processes_per_rank": workers}
    for name, samples, lanes in (("c1", 1000 if world == 1 else 40, 64 if world == 1 else 32),  # N = 1: BASELINE's 11 x 1000 drops
                                 ("ofdm", 64 if world == 1 else 8, 64 if world == 1 else 16),
                                 ("uma", 64 if world == 1 else 8, 64 if world == 1 else 16)):
        try:
            rec, err = _time_limited(SIMULATION_TIME_LIMIT_S, sc.run_gpu, name, samples, "f64", lanes, workers), None
        except Exception as e:  # every rank still takes part in the reduction below: one control flow for all of them
            rec, err = None, repr(e)
        t = torch.tensor([rec["seconds"] if rec else float("inf")], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if not np.isfinite(float(t.item())):
            out[name] = {"error": err or "failed on another rank"}
            _FORCE_EXIT = True
            break
        entry = {"drops_per_s": world * rec["drops"] / float(t.item()), "drops_per_rank": rec["drops"], "lanes": lanes,
                 "links_per_launch_round": rec["links_per_round"], "kernel_launches_per_rank": rec["kernel_launches"],
                 "ber": rec["ber"], "owner_seconds": rec.get("owner_seconds")}
        if world == 1 and not args.no_cpu_baseline:
            try:
                ref = _time_limited(SIMULATION_TIME_LIMIT_S, sc.run_reference, name, max(cores, samples), cores)
            except Exception as e:
                entry["reference_error"] = repr(e)
                out[name] = entry
                _FORCE_EXIT = True
                break
            entry["reference_all_host_cores_drops_per_s"] = ref["drops_per_s"]
            entry["reference_ber"] = ref["ber"]
            entry["reference_propagate_share_of_run"] = ref["propagate_share_of_run"]
            entry["amdahl_bound_if_propagate_were_free"] = ref["amdahl_bound_if_propagate_were_free"]
            entry["speedup_vs_reference_all_host_cores"] = entry["drops_per_s"] / ref["drops_per_s"]
        out[name] = entry
    return out if rank == 0 else None


def run_ours(args):
    import torch
    import torch.distributed as dist

    from hermespy_b200 import _lib

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
    allreduce = world > 1 and not os.environ.get("HB_BENCH_SKIP_COLLECTIVE")
    head = measure(CONFIGS[args.config], args.precision, args, world, rank, dev, args.steps, not args.no_cpu_baseline, allreduce,
                   sinc=args.sinc)
    subs = {}
    if not args.only:
        sub_steps = max(1, min(args.steps, 20))
        todo = [(c, "f32", False) for c in CONFIGS if c != args.config] + \
               [(args.config, "f64" if args.precision == "f32" else "f32", False), ("C5", "f32", True)]
        for cid, prec, sinc in todo:
            try:
                rec = measure(CONFIGS[cid], prec, args, world, rank, dev, sub_steps,
                              not args.no_cpu_baseline and prec == "f32" and not sinc, allreduce, sinc=sinc)
            except Exception as e:  # one config must not hide the others; the failure is part of the record
                rec = {"error": repr(e), "config": public_config(CONFIGS[cid], prec)}
                if world > 1:
                    raise
            if rank == 0:
                subs[f"{cid}_sinc" if sinc else (cid if prec == "f32" or cid != args.config else f"{cid}_{prec}")] = rec
    sim = None
    if not args.only and not args.no_simulation:
        try:
            sim = measure_simulation(args, world, rank, dev)
        except Exception as e:
            sim = {"error": repr(e)}
            if world > 1:
                raise
    if rank == 0:
        if sim is not None:
            head["simulation_api"] = sim
        if subs:
            head["configs"] = subs
            head["gpu_launches_all_configs"] = head["gpu_launches"] + sum(r.get("gpu_launches", 0) for r in subs.values())
            head["parity_all_ok"] = bool(head["parity"]["ok"] and all(r.get("parity", {}).get("ok", False) for r in subs.values()))
        print(json.dumps(head))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=sorted(CONFIGS), help="headline config of the printed line")
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"],
                    help="f32: complex64 arithmetic (rel-L2 <= 1e-5); f64: float64 parity mode (bit-exact BER counts)")
    ap.add_argument("--only", action="store_true", help="measure the headline config only (no per-config sub-records)")
    ap.add_argument("--sinc", action="store_true", help="headline config with fractional (windowed-sinc) delays (extension)")
    ap.add_argument("--links", type=int, default=0, help="links per step per GPU of the headline config (0 = config default)")
    ap.add_argument("--e2e-links", type=int, default=0, help="links per end-to-end step per GPU (0 = config default)")
    ap.add_argument("--stats-every", type=int, default=10, help="N > 1: batches between statistics all-reduces")
    ap.add_argument("--reserve-sms", type=int, default=0, help="N > 1: SMs left to the NCCL all-reduce of the statistics")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs: skip the end-to-end leg")
    ap.add_argument("--no-simulation", action="store_true", help="skip the Simulation.run() drops/s record")
    ap.add_argument("--sos-mode", default="auto", choices=["auto", "poly", "poly_window", "poly_gather", "poly_tma", "poly_fused", "poly_siso", "direct"],
                    help="kernel selection (profiling / A-B runs); the default lets the planner choose")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    rc = run_ours(args)
    if _FORCE_EXIT:
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(rc)
    return rc


if __name__ == "__main__":
    sys.exit(main())
