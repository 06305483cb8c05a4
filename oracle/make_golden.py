"""Generate tests/golden/fading_golden.npz from the UNMODIFIED reference (build container only).

    python -m oracle.make_golden

For every case of ``oracle/golden_cases.py`` the reference channel is realized twice; the second realization is
sampled (forward and reciprocal direction), a seeded signal is propagated and the channel state is taken.
Stored per case: all public sample parameters, the input, ``propagate`` output, reciprocal output and
(for small cases) the dense CSI.  The file is what ``-m gpu`` parity tests and the CPU oracle tests replay
on machines without the reference tree.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.refload import load_reference  # noqa: E402

load_reference()
import hermespy.channel as RC  # noqa: E402
from hermespy.core import Signal, Transformation  # noqa: E402
from hermespy.simulation import SimulatedDevice, SimulatedIdealAntenna, SimulatedUniformArray  # noqa: E402

from oracle.golden_cases import FADING_CASES, SAMPLE_FIELDS, golden_signal  # noqa: E402


def device(n, fs, pos):
    return SimulatedDevice(bandwidth=fs, oversampling_factor=1, carrier_frequency=3.5e9,
                           antennas=SimulatedUniformArray(SimulatedIdealAntenna, 0.04, (n, 1, 1)),
                           pose=Transformation.From_Translation(np.array(pos, dtype=float)))


def main():
    out = {}
    for ci, (name, build, ntx, nrx, fs, T, ptx, prx) in enumerate(FADING_CASES):
        ch = build(RC)
        tx, rx = device(ntx, fs, ptx), device(nrx, fs, prx)
        ch.realize()
        real = ch.realize()  # second realization: checks that the generator state carries over
        s = real.sample(tx, rx)
        for f in SAMPLE_FIELDS:
            out[f"{name}/{f}"] = np.asarray(getattr(s, f))
        out[f"{name}/scalars"] = np.array([s.los_doppler, s.nlos_doppler, s.gain, s.expected_energy_scale, fs])
        x = golden_signal(ci, ntx, T)
        y = s.propagate(Signal.Create(x, fs, 3.5e9)).view(np.ndarray)
        out[f"{name}/y"] = np.asarray(y)
        sr = real.reciprocal_sample(s, rx, tx)
        xr = golden_signal(100 + ci, nrx, T)
        out[f"{name}/y_reciprocal"] = np.asarray(sr.propagate(Signal.Create(xr, fs, 3.5e9)).view(np.ndarray))
        if T <= 200:
            taps = 1 + y.shape[1] - T
            out[f"{name}/csi"] = np.asarray(s.state(T, taps).dense_state()).astype(np.complex128)
        print(f"{name}: y {y.shape}")
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "fading_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()


def cdl_main():
    from oracle.golden_cases import CDL_CASES, CDL_FC, CDL_FS, CDL_SAMPLE_FIELDS, CDL_SPACING

    def dev(spec):
        dims, rpy, pos, vel = spec
        return SimulatedDevice(bandwidth=CDL_FS, oversampling_factor=1, carrier_frequency=CDL_FC,
                               antennas=SimulatedUniformArray(SimulatedIdealAntenna, CDL_SPACING, dims),
                               pose=Transformation.From_RPY(np.array(rpy, float), np.array(pos, float)),
                               velocity=np.array(vel, float))

    out = {}
    for ci, (name, build, txs, rxs, T) in enumerate(CDL_CASES):
        ch = build(RC)
        tx, rx = dev(txs), dev(rxs)
        ch.realize()
        real = ch.realize()
        s = real.sample(tx, rx)
        for f in CDL_SAMPLE_FIELDS:
            out[f"{name}/{f}"] = np.asarray(getattr(s, f))
        out[f"{name}/scalars"] = np.array([float(s.line_of_sight), s.rice_factor, s.delay_offset, s.cluster_delay_spread,
                                           s.max_delay, s.expected_energy_scale])
        ntx, nrx = int(np.prod(txs[0])), int(np.prod(rxs[0]))
        x = golden_signal(200 + ci, ntx, T)
        y = s.propagate(Signal.Create(x, CDL_FS, CDL_FC)).view(np.ndarray)
        out[f"{name}/y"] = np.asarray(y)
        sr = real.reciprocal_sample(s, rx, tx)
        out[f"{name}/y_reciprocal"] = np.asarray(
            sr.propagate(Signal.Create(golden_signal(300 + ci, nrx, T), CDL_FS, CDL_FC)).view(np.ndarray))
        if T <= 100:
            out[f"{name}/csi"] = np.asarray(s.state(T, 1000).dense_state()).astype(np.complex128)
        print(f"{name}: y {y.shape}")
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "cdl_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    cdl_main()
