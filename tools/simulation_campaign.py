#!/usr/bin/env python
"""GPU box: end-to-end drops/s of UNMODIFIED ``Simulation.run()`` scripts -- stock reference on all host cores vs the batched
drop runner with the channel on the GPU.

    python tools/simulation_campaign.py [--config c1|ofdm] [--samples N] [--lanes B] [--workers W] [--out file.json]

* ``c1``   BASELINE config C1: SISO RRC modem over 5G TDL-A, BER sweep dB(0..20) (11 points), ``num_samples`` drops each
           (the reference's getting-started simulation).
* ``ofdm`` the C2 frame through a modem the reference can demodulate: 2x1 Alamouti OFDM, 1024 subcarriers x 14 symbols
           (15 344 samples @30.72 MHz), ideal CSI, TDL-B 300 ns with Doppler, 7 SNR points.
* ``uma``  the same link over the stochastic 3GPP Urban Macrocell model (cluster_delay_lines.py:1824-2013): every drop draws
           its own line-of-sight state, cluster count and cluster delays; 3 SNR points.

Arms, same script text, same box:
  reference   ``cores`` processes, each running ``Simulation.run()`` on its share of the samples with the reference's numpy
              channel (what Ray's one-actor-per-core does, monte_carlo.py:176,621; ray itself is not in this image)
  gpu_serial  one process, ``dropin.enable(precision)``: the per-drop loop, launches of one link
  gpu_batched one process, ``dropin.enable(precision, batch_drops=B, workers=W)``: B drops in flight, one launch per
              round for all their links, W forked helpers for the modem stages
Reported: drops/s, bit-error rates per SNR point (statistical agreement), links per launch.
"""
import argparse
import json
import os
import sys
import time

os.environ.setdefault("OMP_NUM_THREADS", "1")  # one thread per process in BOTH arms: processes are the parallelism
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
os.environ.setdefault("MKL_NUM_THREADS", "1")

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build_c1(num_samples, seed):
    from hermespy.channel import TDL, TDLType
    from hermespy.core import ConsoleMode, dB
    from hermespy.modem import (BitErrorEvaluator, RootRaisedCosineWaveform, SimplexLink,
                                SingleCarrierLeastSquaresChannelEstimation, SingleCarrierZeroForcingChannelEqualization)
    from hermespy.simulation import SNR, Simulation

    sim = Simulation(console_mode=ConsoleMode.SILENT, num_samples=num_samples, seed=seed)
    tx = sim.new_device(oversampling_factor=4)
    rx = sim.new_device(oversampling_factor=4)
    rx.noise_level = SNR(dB(20), tx)
    sim.set_channel(tx, rx, TDL(TDLType.A, doppler_frequency=100.0))
    link = SimplexLink(seed=seed + 1)
    tx.transmitters.add(link)
    rx.receivers.add(link)
    link.waveform = RootRaisedCosineWaveform(num_preamble_symbols=10, num_data_symbols=100, roll_off=0.9)
    link.waveform.channel_estimation = SingleCarrierLeastSquaresChannelEstimation()
    link.waveform.channel_equalization = SingleCarrierZeroForcingChannelEqualization()
    sim.new_dimension("noise_level", dB(*range(0, 21, 2)), rx)
    sim.add_evaluator(BitErrorEvaluator(link, link))
    return sim


def build_uma(num_samples, seed):
    """The OFDM link of ``build_ofdm`` over a STOCHASTIC 3GPP scenario (UMa, line-of-sight state drawn per realization):
    every drop has its own cluster delays / cluster count, so a batch of drops is a heterogeneous CDL batch."""
    from hermespy.channel import UrbanMacrocells

    return build_ofdm(num_samples, seed, channel=UrbanMacrocells(seed=seed + 7), snrs=(0, 10, 20), rx_velocity=(3.0, -1.0, 0.0))


def build_ofdm(num_samples, seed, channel=None, snrs=(0, 5, 10, 15, 20, 25, 30), rx_velocity=None):
    from hermespy.channel import TDL, TDLType
    from hermespy.core import ConsoleMode, Transformation, dB
    from hermespy.modem import (Alamouti, BitErrorEvaluator, ChannelEqualization, ElementType, GridElement, GridResource,
                                OFDMWaveform, SimplexLink, SymbolSection)
    from hermespy.simulation import (SNR, OFDMIdealChannelEstimation, SimulatedIdealAntenna, SimulatedUniformArray,
                                     Simulation)

    sim = Simulation(console_mode=ConsoleMode.SILENT, num_samples=num_samples, seed=seed)
    fc, bw = 3.5e9, 30.72e6
    lam = 299792458.0 / fc

    def dev(n, pos):
        return sim.new_device(carrier_frequency=fc, bandwidth=bw, oversampling_factor=1,
                              pose=Transformation.From_Translation(np.array(pos, float)),
                              antennas=SimulatedUniformArray(SimulatedIdealAntenna, 0.5 * lam, [n, 1, 1]))

    tx, rx = dev(2, (0.0, 0.0, 25.0)), dev(1, (100.0, 20.0, 1.5))
    if rx_velocity is not None:
        rx.velocity = np.array(rx_velocity, float)
    ch = TDL(TDLType.B, rms_delay=300e-9, doppler_frequency=100.0) if channel is None else channel
    # scenario models carry a path loss: the noise level follows the sample's expected energy scale (noise/level.py:207-241)
    rx.noise_level = SNR(dB(15), tx) if channel is None else SNR(dB(15), tx, ch)
    sim.set_channel(tx, rx, ch)
    link = SimplexLink(seed=seed + 1)
    tx.transmitters.add(link)
    rx.receivers.add(link)
    res = [GridResource(1024 // 8, prefix_ratio=72 / 1024, elements=[GridElement(ElementType.REFERENCE, 1), GridElement(ElementType.DATA, 7)])]
    link.waveform = OFDMWaveform(num_subcarriers=1024, dc_suppression=False, grid_resources=res,
                                 grid_structure=[SymbolSection(14, [0], 1)], modulation_order=4)
    link.waveform.channel_estimation = OFDMIdealChannelEstimation(ch, tx, rx)
    link.waveform.channel_equalization = ChannelEqualization()
    link.transmit_symbol_coding[0] = Alamouti()
    link.receive_symbol_coding[0] = Alamouti()
    sim.new_dimension("noise_level", dB(*snrs), rx)
    sim.add_evaluator(BitErrorEvaluator(link, link))
    return sim


BUILDERS = {"c1": (build_c1, 11), "ofdm": (build_ofdm, 7), "uma": (build_uma, 3)}


def _fast_polling(sim):
    """``MonteCarlo.simulate`` sleeps ``progress_log_interval`` = 4 s between its progress polls (monte_carlo.py:404; not a
    constructor argument of ``Simulation``): a campaign's wall time is quantised to 4 s.  BOTH arms poll every 50 ms so
    that the clock measures drops, not the poll."""
    sim._MonteCarlo__progress_log_interval = 0.05
    return sim


def _reference_process(args):
    """One host process of the reference arm: Simulation.run() on its share of the samples (numpy channel)."""
    cfg, samples, seed = args
    try:  # one thread per process also when the importing program loaded its BLAS with the default thread count
        from threadpoolctl import threadpool_limits

        threadpool_limits(1)
    except Exception:
        pass
    from oracle.refload import load_reference

    load_reference()
    from hermespy.channel.channel import ChannelSample

    sim = _fast_polling(BUILDERS[cfg][0](samples, seed))
    spent = [0.0]
    orig = ChannelSample.propagate

    def timed(self, *a, **k):  # share of the run spent inside the channel: what a faster channel can remove (Amdahl)
        t = time.perf_counter()
        try:
            return orig(self, *a, **k)
        finally:
            spent[0] += time.perf_counter() - t

    ChannelSample.propagate = timed
    try:
        t0 = time.perf_counter()
        res = sim.run()
        dt = time.perf_counter() - t0
    finally:
        ChannelSample.propagate = orig
    return np.asarray(res.evaluation_results[0].to_array(), dtype=float).ravel().tolist(), dt, spent[0]


def run_reference(cfg, samples, cores):
    import multiprocessing as mp

    per = max(1, samples // cores)
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_reference_process, [(cfg, 1, 900 + i) for i in range(cores)])  # warm-up: imports, numba
        t0 = time.perf_counter()
        out = pool.map(_reference_process, [(cfg, per, 1000 + i) for i in range(cores)])
        dt = time.perf_counter() - t0
    points = BUILDERS[cfg][1]
    share = float(np.mean([o[2] / o[1] for o in out]))  # includes the ideal self-links; channel-model share is below it
    return dict(processes=cores, samples_per_point=per * cores, drops=per * cores * points, seconds=dt,
                drops_per_s=per * cores * points / dt, ber=np.mean([o[0] for o in out], axis=0).tolist(),
                propagate_share_of_run=share, amdahl_bound_if_propagate_were_free=1.0 / max(1e-9, 1.0 - share))


def run_gpu(cfg, samples, precision, lanes, workers):
    import hermespy_b200.dropin as dropin
    from hermespy_b200 import _lib, runner

    points = BUILDERS[cfg][1]
    dropin.enable(precision=precision, batch_drops=lanes, workers=workers)
    try:
        _fast_polling(BUILDERS[cfg][0](2, 5)).run()  # warm-up
        runner.stats.update(rounds=0, drops=0, links=0, max_links_per_round=0)
        before = sum(_lib.launch_counts().values())
        sim = _fast_polling(BUILDERS[cfg][0](samples, 1000))
        t0 = time.perf_counter()
        res = sim.run()
        dt = time.perf_counter() - t0
        launches = sum(_lib.launch_counts().values()) - before
    finally:
        dropin.disable()
    return dict(precision=precision, lanes=lanes, workers=workers, samples_per_point=samples, drops=samples * points,
                seconds=dt, drops_per_s=samples * points / dt, kernel_launches=launches,
                links=runner.stats["links"] if lanes else None, rounds=runner.stats["rounds"] if lanes else None,
                links_per_round=runner.stats["max_links_per_round"] if lanes else 1,
                owner_seconds=runner.stats.get("seconds") if lanes else None,
                ber=np.asarray(res.evaluation_results[0].to_array(), dtype=float).ravel().tolist())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c1", choices=sorted(BUILDERS))
    ap.add_argument("--samples", type=int, default=1000)
    ap.add_argument("--lanes", type=int, default=0, help="drops in flight (0 = 4 x workers)")
    ap.add_argument("--workers", type=int, default=-1, help="helper processes (-1 = host cores)")
    ap.add_argument("--precision", default="f64")
    ap.add_argument("--skip-reference", action="store_true")
    ap.add_argument("--skip-serial", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    from oracle.refload import load_reference, reference_available

    if not reference_available():
        print(json.dumps({"unavailable": "no reference install (baseline/_ref)"}))
        return
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    workers = cores if args.workers < 0 else args.workers
    lanes = args.lanes or max(8, 4 * max(1, workers))
    out = {"config": args.config, "host_cores": cores}
    if not args.skip_reference:
        out["reference"] = run_reference(args.config, args.samples, cores)
        print(json.dumps({"reference": out["reference"]}), flush=True)
    load_reference()
    if not args.skip_serial:
        out["gpu_serial"] = run_gpu(args.config, max(8, args.samples // 8), args.precision, 0, 0)
        print(json.dumps({"gpu_serial": out["gpu_serial"]}), flush=True)
    out["gpu_batched"] = run_gpu(args.config, args.samples, args.precision, lanes, workers)
    print(json.dumps({"gpu_batched": out["gpu_batched"]}), flush=True)
    if "reference" in out:
        out["speedup_vs_reference_all_cores"] = out["gpu_batched"]["drops_per_s"] / out["reference"]["drops_per_s"]
    path = args.out or os.path.join(ROOT, "gpurun_out", f"simulation_campaign_{args.config}.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps({k: v for k, v in out.items() if not isinstance(v, dict)}), flush=True)
    if os.environ.get("HB_EXIT_WATCHDOG"):  # diagnose a slow interpreter exit: all thread stacks after N seconds, then leave
        import faulthandler

        faulthandler.dump_traceback_later(float(os.environ["HB_EXIT_WATCHDOG"]), exit=True, file=sys.stderr)
        print("main() returned at", time.perf_counter(), file=sys.stderr, flush=True)


if __name__ == "__main__":
    main()
