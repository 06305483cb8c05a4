#!/usr/bin/env python
"""The reference's getting-started script (`_examples/getting_started/simulation.py`), UNMODIFIED in its simulation part,
through `Simulation.run()` with the channel hot path on the GPU.

    python tools/run_simulation_example.py [--num-samples 50] [--no-gpu]

Needs the reference install (baseline/_ref).  `ray` is the in-process stand-in `hermespy_b200.shims.ray`, matplotlib / h5py
are inert stubs (`oracle/refload.py`); `hermespy_b200.dropin.enable("f64")` routes `MultipathFadingSample._propagate` to the
CUDA kernels.  Prints the BER per SNR point and the number of kernel launches.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--num-samples", type=int, default=50)
    ap.add_argument("--no-gpu", action="store_true")
    args = ap.parse_args()
    from oracle.refload import load_reference, reference_available

    if not reference_available():
        print(json.dumps({"unavailable": "no reference install"}))
        return
    load_reference()
    from hermespy_b200 import _lib, dropin

    if not args.no_gpu:
        dropin.enable(precision="f64")

    # ---- from here on: the reference's example, verbatim except for plotting and the sample count ----------------------
    from hermespy.core import ConsoleMode, dB
    from hermespy.channel import TDL
    from hermespy.simulation import Simulation, SNR
    from hermespy.modem import (BitErrorEvaluator, SimplexLink, RootRaisedCosineWaveform,
                                SingleCarrierLeastSquaresChannelEstimation, SingleCarrierZeroForcingChannelEqualization)

    simulation = Simulation(console_mode=ConsoleMode.SILENT, num_samples=args.num_samples, seed=42)
    tx_device = simulation.new_device(oversampling_factor=4)
    rx_device = simulation.new_device(oversampling_factor=4)
    tx_device.noise_level = SNR(dB(20), tx_device)
    rx_device.noise_level = SNR(dB(20), tx_device)
    simulation.set_channel(tx_device, rx_device, TDL())
    link = SimplexLink()
    tx_device.transmitters.add(link)
    rx_device.receivers.add(link)
    link.waveform = RootRaisedCosineWaveform(num_preamble_symbols=10, num_data_symbols=100, roll_off=.9)
    link.waveform.channel_estimation = SingleCarrierLeastSquaresChannelEstimation()
    link.waveform.channel_equalization = SingleCarrierZeroForcingChannelEqualization()
    ber = BitErrorEvaluator(link, link)
    simulation.new_dimension('noise_level', dB(20, 16, 12, 8, 4, 0), rx_device)
    simulation.add_evaluator(ber)
    before = sum(_lib.launch_counts().values())
    t0 = time.perf_counter()
    result = simulation.run()
    dt = time.perf_counter() - t0
    # ---------------------------------------------------------------------------------------------------------------------
    launches = sum(_lib.launch_counts().values()) - before
    print(json.dumps({"ber": [float(v) for v in result.evaluation_results[0].to_array().ravel()], "snr_db": [20, 16, 12, 8, 4, 0],
                      "drops": 6 * args.num_samples, "seconds": dt, "gpu_kernel_launches": int(launches),
                      "channel": "numpy (reference)" if args.no_gpu else "CUDA (hermespy_b200.dropin, f64)"}))


if __name__ == "__main__":
    main()
