#!/usr/bin/env python
"""GPU box: every BASELINE.json config through the device-resident C-ABI, with per-kernel CUDA-event times.

    python tools/config_report.py [--configs C1,C2,C3,C4,C5] [--steps 20] [--out gpurun_out/config_report.json]

For each config: links are realized + sampled on the host by the mirror classes (numpy RNG, as the reference),
the signal is synthetic complex64 resident in HBM, `steps` propagations are timed with CUDA events on the launch
stream after 3 warm-up steps (inputs + outputs exceed the 126 MB L2 for every config), and the library's own
per-kernel events (hb_profile_begin/end) give the time of each kernel.  Reported per config: propagated complex
samples/s per link direction, the algorithmic HBM bytes per launch (8 (Ntx + Nrx) B per sample +
delay tail) and the fraction of the measured HBM peak for the whole step and for its dominant kernel.
The batch sizes are the per-GPU share of the config where that fits in seconds of host sampling (C3: 256 distinct
CDL links tiled to the batch -- the kernels do not know).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def time_steps(fn, steps):
    import torch
    from hermespy_b200 import _lib

    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    _lib.profile_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    prof = _lib.profile_end()
    return e0.elapsed_time(e1) / steps, {k: v["ms"] / steps for k, v in prof.items() if v["launches"]}, \
        {k: v["launches"] // steps for k, v in prof.items() if v["launches"]}


def rand_x(B, n, T, seed=0):
    import torch

    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    return torch.view_as_complex(torch.randn((B, n, T, 2), device="cuda", generator=g, dtype=torch.float32) * 0.5 ** 0.5)


def fading_config(name, ch, B, ntx, nrx, T, fs, steps, note=""):
    import torch
    from hermespy_b200.batch import sample_fading_links
    from hermespy_b200.kernels import FadingBatch, fading_propagate

    t0 = time.perf_counter()
    blk = sample_fading_links(ch, B, ntx, nrx, fs)
    t_host = time.perf_counter() - t0
    fb = FadingBatch.from_numpy(device="cuda", **blk)
    D = blk["max_delay"]
    x = rand_x(B, ntx, T)
    y = torch.empty((B, nrx, T + D), dtype=torch.complex64, device="cuda")
    _, info = fading_propagate(x, fb, out=y, return_info=True)
    ms, per_kernel, launches = time_steps(lambda: fading_propagate(x, fb, out=y), steps)
    alg = 8.0 * B * (ntx * T + nrx * (T + D))
    return dict(config=name, note=note, links=B, ntx=ntx, nrx=nrx, T=T, D=D, plan=info, ms_per_step=ms,
                samples_per_s=B * T / (ms * 1e-3), algorithmic_bytes=alg, step_gbs=alg / (ms * 1e-3) / 1e9,
                step_frac_of_hbm_peak=alg / (ms * 1e-3) / 1e9 / peak(), kernel_ms=per_kernel, kernel_launches=launches,
                host_sampling_s=t_host)


def c1(steps):
    import hermespy_b200.channel as MC

    return fading_config("C1", MC.TDL(MC.TDLType.A, doppler_frequency=100.0, seed=42), 11000, 1, 1, 500, 4e8, steps,
                         "SISO TDL-A, rms_delay 0 (all 23 taps in one delay group), 11 SNR points x 1000 drops")


def c2(steps):
    import hermespy_b200.channel as MC

    ch = MC.TDL(MC.TDLType.B, rms_delay=300e-9, doppler_frequency=100, seed=42,
                antenna_correlation=MC.StandardAntennaCorrelation(MC.CorrelationType.MEDIUM))
    return fading_config("C2", ch, 2048, 4, 4, 15344, 30.72e6, steps, "4x4 TDL-B 300 ns, 2048 of the 70000 links per step")


def c4(steps):
    import hermespy_b200.channel as MC

    n = 64
    R = 0.7 ** np.abs(np.subtract.outer(np.arange(n), np.arange(n))).astype(complex)
    ch = MC.TDL(MC.TDLType.D, rms_delay=300e-9, doppler_frequency=100, seed=42, max_antennas=n,
                antenna_correlation=MC.CustomAntennaCorrelation(R))
    return fading_config("C4", ch, 256, n, n, 16384, 30.72e6, steps,
                         "64x64 Rician (TDL-D profile), exponential Kronecker correlation rho = 0.7; z mode + tcgen05 3xTF32 GEMM")


def c5(steps):
    import hermespy_b200.channel as MC

    ch = MC.Cost259(MC.Cost259Type.URBAN, doppler_frequency=50, seed=42)
    return fading_config("C5", ch, 64, 1, 1, 1 << 20, 30.72e6, steps, "COST259 urban SISO, 1M-sample frames, NEAREST delays")


def c3(steps):
    import torch
    import hermespy_b200.channel as MC
    from hermespy_b200.core import SimulatedDevice, SimulatedIdealAntenna, SimulatedUniformArray, Transformation
    from hermespy_b200.kernels import CdlBlock, CdlDeviceBlock, cdl_propagate

    fc, fs, T = 3.5e9, 30.72e6, 2048
    lam = 299792458.0 / fc
    tx = SimulatedDevice(bandwidth=fs, carrier_frequency=fc, antennas=SimulatedUniformArray(SimulatedIdealAntenna, lam / 2, (8, 4, 1)),
                         pose=Transformation.From_Translation(np.array([0.0, 0.0, 25.0])))
    rx = SimulatedDevice(bandwidth=fs, carrier_frequency=fc, antennas=SimulatedUniformArray(SimulatedIdealAntenna, lam / 2, (2, 2, 1)),
                         pose=Transformation.From_Translation(np.array([100.0, 20.0, 1.5])), velocity=np.array([10.0, -3.0, 0.0]))
    ch = MC.CDL(MC.CDLType.C, 300e-9, seed=42)
    t0 = time.perf_counter()
    distinct, B = 128, 4096
    blocks = [ch.realize().sample(tx, rx).kernel_block() for _ in range(distinct)]
    t_host = time.perf_counter() - t0
    blk = CdlBlock.stack([blocks[i % distinct] for i in range(B)])
    dblk = CdlDeviceBlock(blk)
    x = rand_x(B, 32, T)
    y, info = cdl_propagate(x, dblk, return_info=True)
    ms, per_kernel, launches = time_steps(lambda: cdl_propagate(x, dblk, out=y), steps)
    D = blk.max_delay
    alg = 8.0 * B * (32 * T + 4 * (T + D))
    return dict(config="C3", note=f"CDL-C 32x4 UPA, moving receiver, {B} links per step ({distinct} distinct samples tiled)",
                links=B, ntx=32, nrx=4, T=T, D=D, plan=info, ms_per_step=ms, samples_per_s=B * T / (ms * 1e-3),
                algorithmic_bytes=alg, step_gbs=alg / (ms * 1e-3) / 1e9, step_frac_of_hbm_peak=alg / (ms * 1e-3) / 1e9 / peak(),
                kernel_ms=per_kernel, kernel_launches=launches, host_sampling_s=t_host)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="C1,C2,C3,C4,C5")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "config_report.json"))
    args = ap.parse_args()
    import torch

    assert torch.cuda.is_available(), "needs a GPU"
    table = {"C1": c1, "C2": c2, "C3": c3, "C4": c4, "C5": c5}
    rows = []
    for name in args.configs.split(","):
        try:
            r = table[name](args.steps)
        except Exception as e:  # keep going: one config must not hide the others
            r = dict(config=name, error=repr(e))
        rows.append(r)
        print(json.dumps(r))
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(dict(hbm_peak_gbs=peak(), rows=rows), open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
