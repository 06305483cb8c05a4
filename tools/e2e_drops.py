#!/usr/bin/env python
"""GPU box: end-to-end drops/s of the UNMODIFIED reference drop loop, with and without the CUDA channel path.

    python tools/e2e_drops.py [--drops 20] [--out gpurun_out/e2e_drops.json]

Scenario (reference API only, from baseline/_ref): a 2x1 Alamouti OFDM link, 1024 subcarriers, 14 symbols with a 72-sample
cyclic prefix (15 344 samples at 30.72 MHz, the frame of BASELINE config C2), ideal CSI, over TDL-B 300 ns with Doppler;
`SimulationScenario.drop()` + `BitErrorEvaluator` per drop, single process.  Reported: drops/s and the share of the drop
spent inside `ChannelSample.propagate` for (a) the reference numpy path, (b) `dropin.enable("f64")`, (c) `"f32"`, and
whether the bit-error totals agree.  This is the SURVEY 8(d) "end-to-end" figure; the per-kernel numbers are bench.py's.
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
    from hermespy.core import Transformation
    from hermespy.modem import (Alamouti, BitErrorEvaluator, ChannelEqualization, ElementType, GridElement, GridResource,
                                OFDMWaveform, SimplexLink, SymbolSection)
    from hermespy.simulation import OFDMIdealChannelEstimation, SimulatedIdealAntenna, SimulatedUniformArray, SimulationScenario

    sc = SimulationScenario(seed=seed)
    fc, bw = 3.5e9, 30.72e6
    lam = 299792458.0 / fc

    def dev(n, pos):
        return sc.new_device(carrier_frequency=fc, bandwidth=bw, oversampling_factor=1,
                             pose=Transformation.From_Translation(np.array(pos, float)),
                             antennas=SimulatedUniformArray(SimulatedIdealAntenna, 0.5 * lam, [n, 1, 1]))

    tx, rx = dev(2, (0.0, 0.0, 25.0)), dev(1, (100.0, 20.0, 1.5))
    ch = TDL(TDLType.B, rms_delay=300e-9, doppler_frequency=100.0)
    sc.set_channel(tx, rx, ch)
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
    tx.noise_model.seed = seed + 2
    rx.noise_model.seed = seed + 3
    return sc, tx, rx, BitErrorEvaluator(link, link)


def run(drops, seed=42):
    from hermespy.channel.channel import ChannelSample
    from hermespy.core import dB
    from hermespy.simulation import SNR

    sc, tx, rx, ber = build(seed)
    rx.noise_level = SNR(dB(15), tx)
    spent = [0.0]
    orig = ChannelSample.propagate

    def timed(self, *a, **k):
        t0 = time.perf_counter()
        try:
            return orig(self, *a, **k)
        finally:
            spent[0] += time.perf_counter() - t0

    ChannelSample.propagate = timed
    try:
        sc.drop()  # warm-up (numba, page-in)
        ber.evaluate()
        spent[0] = 0.0
        errors = bits = 0
        samples = 0
        t0 = time.perf_counter()
        for _ in range(drops):
            d = sc.drop()
            e = np.asarray(ber.evaluate().evaluation)
            errors += int(e.sum())
            bits += e.size
            samples = d.device_transmissions[0].mixed_signal.num_samples
        dt = time.perf_counter() - t0
    finally:
        ChannelSample.propagate = orig
    return dict(drops=drops, seconds=dt, drops_per_s=drops / dt, propagate_share=spent[0] / dt,
                propagate_ms_per_drop=1e3 * spent[0] / drops, bit_errors=errors, bits=bits, frame_samples=int(samples))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--drops", type=int, default=20)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "e2e_drops.json"))
    args = ap.parse_args()
    from oracle.refload import load_reference, reference_available

    if not reference_available():
        print(json.dumps({"unavailable": "no reference install (baseline/_ref)"}))
        return
    load_reference()
    import hermespy_b200.dropin as dropin

    out = {"scenario": "2x1 Alamouti OFDM, 1024 subcarriers x 14 symbols (15344 samples), ideal CSI, TDL-B 300 ns Doppler 100, SNR 15 dB",
           "cores": 1}
    dropin.disable()
    out["reference"] = run(args.drops)
    dropin.enable(precision="f64")
    out["dropin_f64"] = run(args.drops)
    dropin.enable(precision="f32")
    out["dropin_f32"] = run(args.drops)
    dropin.disable()
    out["bit_exact_f64"] = out["reference"]["bit_errors"] == out["dropin_f64"]["bit_errors"]
    out["speedup_drop_loop_f64"] = out["dropin_f64"]["drops_per_s"] / out["reference"]["drops_per_s"]
    out["speedup_propagate_f64"] = out["reference"]["propagate_ms_per_drop"] / out["dropin_f64"]["propagate_ms_per_drop"]
    print(json.dumps(out, indent=1))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
