"""Generate tests/golden/cdl_elements_golden.npz from the UNMODIFIED reference (build container only).

    python -m oracle.make_golden_elements

CDL links whose arrays carry the reference's non-ideal element models -- ``Dipole``, ``PatchAntenna``,
``LinearAntenna`` (hermespy/core/antennas.py:447-622) -- in uniform arrays and in custom arrays with per-element
slants / orientations (cross-polarized pairs).  Stored per case, self-contained (no host classes needed to replay):
the sample's public parameters, both array geometries with their element tables, the input, ``propagate`` output and
the dense channel state.  TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.refload import load_reference  # noqa: E402

load_reference()
import hermespy.channel as RC  # noqa: E402
from hermespy.core import Signal, Transformation  # noqa: E402
from hermespy.simulation import (SimulatedCustomArray, SimulatedDevice, SimulatedDipole, SimulatedIdealAntenna,  # noqa: E402
                                 SimulatedLinearAntenna, SimulatedPatchAntenna, SimulatedUniformArray)

from oracle.golden_cases import CDL_FC, CDL_FS, CDL_SPACING, golden_signal  # noqa: E402
from oracle.ref_extract import cdl_params_from_reference_sample  # noqa: E402

PARAM_FIELDS = ("aoa zoa aod zod cluster_delays cluster_powers jones").split()
GEOMETRY_FIELDS = ("rotation translation topology velocity elements").split()


def _pose(rpy, pos):
    return Transformation.From_RPY(np.array(rpy, float), np.array(pos, float))


def _xpol(n):
    """n co-located +-45 degree pairs along y (the usual 3GPP cross-polarized panel)."""
    ants = []
    for i in range(n):
        for slant in (np.pi / 4, -np.pi / 4):
            ants.append(SimulatedLinearAntenna(slant=slant, pose=_pose((0, 0, 0), (0.0, i * CDL_SPACING, 0.0))))
    return SimulatedCustomArray(ants)


def _mixed():
    return SimulatedCustomArray([
        SimulatedDipole(pose=_pose((0.3, 0.0, 0.0), (0.0, 0.0, 0.0))),
        SimulatedPatchAntenna(pose=_pose((0.0, -0.4, 0.9), (0.0, CDL_SPACING, 0.0))),
        SimulatedIdealAntenna(pose=_pose((0.2, 0.1, -0.5), (0.0, 0.0, CDL_SPACING))),
        SimulatedLinearAntenna(slant=0.6, pose=_pose((0.0, 0.7, 0.0), (CDL_SPACING, CDL_SPACING, 0.0))),
    ])


ELEMENT_CASES = [
    # (name, channel, tx array, tx rpy, tx position, tx velocity, rx array, rx rpy, rx position, rx velocity, T)
    ("dipole_uniform_4x2", lambda: RC.CDL(RC.CDLType.C, 300e-9, seed=51),
     lambda: SimulatedUniformArray(SimulatedDipole, CDL_SPACING, (2, 2, 1)), (0.1, 0.2, 0.3), (0.0, 0.0, 10.0), (0, 0, 0),
     lambda: SimulatedUniformArray(SimulatedDipole, CDL_SPACING, (2, 1, 1)), (0, 0, 0), (100.0, 20.0, 1.5), (10.0, -3.0, 0.0), 96),
    ("patch_tx_ideal_rx", lambda: RC.CDL(RC.CDLType.A, 300e-9, seed=52),
     lambda: SimulatedUniformArray(SimulatedPatchAntenna, CDL_SPACING, (4, 1, 1)), (0, 0, 0.4), (0.0, 0.0, 10.0), (0, 0, 0),
     lambda: SimulatedUniformArray(SimulatedIdealAntenna, CDL_SPACING, (2, 1, 1)), (0, 0.1, 0), (60.0, -35.0, 1.5), (0.0, 5.0, 0.0), 80),
    ("xpol_linear_custom", lambda: RC.CDL(RC.CDLType.B, 100e-9, seed=53),
     lambda: _xpol(2), (0, 0.15, 0), (0.0, 0.0, 25.0), (0, 0, 0),
     lambda: SimulatedCustomArray([SimulatedLinearAntenna(slant=0.0), SimulatedLinearAntenna(slant=np.pi / 2)]), (0, 0, 1.0),
     (80.0, 10.0, 1.5), (3.0, 1.0, 0.0), 72),
    ("mixed_rotated_los", lambda: RC.CDL(RC.CDLType.D, 300e-9, rayleigh_factor=7.0, seed=54),
     _mixed, (0.05, -0.1, 0.2), (0.0, 0.0, 12.0), (1.0, 0.0, 0.0),
     lambda: SimulatedUniformArray(SimulatedDipole, CDL_SPACING, (1, 2, 1)), (0.0, 0.0, -2.0), (45.0, 30.0, 2.0), (-4.0, 2.0, 0.0), 64),
]


def main():
    out = {}
    for ci, (name, channel, txa, trpy, tpos, tvel, rxa, rrpy, rpos, rvel, T) in enumerate(ELEMENT_CASES):
        def dev(arr, rpy, pos, vel):
            return SimulatedDevice(bandwidth=CDL_FS, oversampling_factor=1, carrier_frequency=CDL_FC, antennas=arr(),
                                   pose=_pose(rpy, pos), velocity=np.array(vel, float))

        tx, rx = dev(txa, trpy, tpos, tvel), dev(rxa, rrpy, rpos, rvel)
        ch = channel()
        s = ch.realize().sample(tx, rx)
        p = cdl_params_from_reference_sample(s)
        for f in PARAM_FIELDS:
            out[f"{name}/{f}"] = np.asarray(getattr(p, f))
        out[f"{name}/scalars"] = np.array([float(p.line_of_sight), p.rice_factor_db, p.delay_offset, p.cluster_delay_spread,
                                           p.fc, p.fs])
        for side, g in (("tx", p.tx), ("rx", p.rx)):
            for f in GEOMETRY_FIELDS:
                out[f"{name}/{side}_{f}"] = np.asarray(getattr(g, f), dtype=np.float64)
        ntx = p.tx.topology.shape[0]
        x = golden_signal(500 + ci, ntx, T)
        y = s.propagate(Signal.Create(x, CDL_FS, CDL_FC)).view(np.ndarray)
        out[f"{name}/y"] = np.asarray(y)
        out[f"{name}/csi"] = np.asarray(s.state(T, 1000).dense_state()).astype(np.complex128)
        print(f"{name}: ntx {ntx} y {y.shape} |y| {np.linalg.norm(y):.4f}")
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                        "cdl_elements_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
