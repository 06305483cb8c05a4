"""Minimal data model either side of the channel hot path.

Mirrors the handful of ``hermespy.core`` / ``hermespy.simulation`` types the channel plugin API
touches, with the same names and argument meaning, so that parity tests read like the reference's
own tests (reference files: hermespy/core/signal_model.py:395-460 ``SignalBlock``,
hermespy/core/definitions.py:74-93 ``InterpolationMode``, hermespy/core/antennas.py ``AntennaMode``,
hermespy/simulation/simulated_device.py:1516-1546 ``SimulatedDeviceState``).  This is host-side
plumbing only; it is not a re-implementation of HermesPy's device model.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field
from enum import Enum
from typing import List, Optional, Sequence, Tuple

import numpy as np


class InterpolationMode(Enum):
    """hermespy/core/definitions.py:74-93."""

    NEAREST = 0
    SINC = 1


class AntennaMode(Enum):
    TX = 0
    RX = 1
    DUPLEX = 2


class SignalBlock(np.ndarray):
    """``[streams, samples]`` complex128 block with a sample offset (signal_model.py:395-428)."""

    _DEFAULT_OFFSET = 0

    def __new__(cls, num_streams: int, num_samples: int, offset: int = 0, buffer=None):
        obj = np.ndarray.__new__(cls, (num_streams, num_samples), np.complex128, buffer=buffer)
        obj.offset = offset
        return obj

    def __array_finalize__(self, obj) -> None:
        if obj is None:
            return
        self.offset = getattr(obj, "offset", 0)

    @classmethod
    def from_array(cls, samples: np.ndarray, offset: int = 0) -> "SignalBlock":
        a = np.ascontiguousarray(samples, dtype=np.complex128)
        if a.ndim == 1:
            a = a[None, :]
        if a.ndim != 2:
            raise ValueError("HermesPy signal models must be two-dimensional")
        blk = a.view(cls)
        blk.offset = int(offset)
        return blk

    @property
    def num_streams(self) -> int:
        return self.shape[0] if self.ndim > 0 else 0

    @property
    def num_samples(self) -> int:
        return self.shape[1] if self.ndim > 1 else 0


class Signal(object):
    """Base-band signal: a list of :class:`SignalBlock` plus sampling metadata.

    ``Signal.Create(samples, sampling_rate, carrier_frequency)`` follows the reference's factory
    (hermespy/core/signal_model.py ``Signal.Create``); dense signals have a single block at offset 0.
    """

    def __init__(self, blocks: Sequence[SignalBlock], sampling_rate: float, carrier_frequency: float = 0.0,
                 noise_power: float = 0.0, delay: float = 0.0) -> None:
        self.blocks: List[SignalBlock] = list(blocks)
        self.sampling_rate = float(sampling_rate)
        self.carrier_frequency = float(carrier_frequency)
        self.noise_power = float(noise_power)
        self.delay = float(delay)

    @classmethod
    def Create(cls, samples, sampling_rate: float = 1.0, carrier_frequency: float = 0.0, noise_power: float = 0.0,
               delay: float = 0.0, offsets: Optional[Sequence[int]] = None) -> "Signal":
        if isinstance(samples, (list, tuple)):
            offs = list(offsets) if offsets is not None else [getattr(b, "offset", 0) for b in samples]
            blocks = [SignalBlock.from_array(np.asarray(b), o) for b, o in zip(samples, offs)]
        else:
            blocks = [SignalBlock.from_array(np.asarray(samples), 0 if offsets is None else offsets[0])]
        return cls(blocks, sampling_rate, carrier_frequency, noise_power, delay)

    @classmethod
    def Empty(cls, sampling_rate: float, num_streams: int = 0, num_samples: int = 0, **kwargs) -> "Signal":
        return cls([SignalBlock(num_streams, num_samples, 0, np.zeros((num_streams, num_samples), np.complex128))],
                   sampling_rate, **kwargs)

    @property
    def num_streams(self) -> int:
        return self.blocks[0].num_streams if self.blocks else 0

    @property
    def num_samples(self) -> int:
        if not self.blocks:
            return 0
        return max(b.offset + b.num_samples for b in self.blocks)

    def to_dense(self) -> np.ndarray:
        out = np.zeros((self.num_streams, self.num_samples), dtype=np.complex128)
        for b in self.blocks:
            out[:, b.offset : b.offset + b.num_samples] = b
        return out

    def __array__(self, dtype=None, copy=None):
        a = self.to_dense()
        return a if dtype is None else a.astype(dtype)

    def view(self, _type=np.ndarray):
        return self.to_dense()

    def __getitem__(self, item):
        return self.to_dense()[item]

    @property
    def power(self) -> np.ndarray:
        a = self.to_dense()
        return np.mean(np.abs(a) ** 2, axis=1) if a.shape[1] else np.zeros(a.shape[0])

    @property
    def energy(self) -> np.ndarray:
        return np.sum(np.abs(self.to_dense()) ** 2, axis=1)


# --------------------------------------------------------------------------------------------------
# Geometry / device state (only what LinkState consumers read)


def rotation_from_rpy(roll: float, pitch: float, yaw: float) -> np.ndarray:
    """Rotation matrix R = Rz(yaw) @ Ry(pitch) @ Rx(roll) (hermespy/core/transformation.py:264-290)."""
    cr, sr = np.cos(roll), np.sin(roll)
    cp, sp = np.cos(pitch), np.sin(pitch)
    cy, sy = np.cos(yaw), np.sin(yaw)
    return np.array(
        [
            [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
            [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
            [-sp, cp * sr, cp * cr],
        ]
    )


class Transformation(object):
    """Rigid pose: 4x4 homogeneous matrix with ``translation`` / ``rotation`` accessors."""

    def __init__(self, matrix: Optional[np.ndarray] = None) -> None:
        self.matrix = np.eye(4) if matrix is None else np.array(matrix, dtype=np.float64)

    @classmethod
    def From_RPY(cls, rpy: np.ndarray, pos: np.ndarray) -> "Transformation":
        m = np.eye(4)
        m[:3, :3] = rotation_from_rpy(*np.asarray(rpy, dtype=np.float64))
        m[:3, 3] = np.asarray(pos, dtype=np.float64)
        return cls(m)

    @classmethod
    def From_Translation(cls, pos: np.ndarray) -> "Transformation":
        return cls.From_RPY(np.zeros(3), pos)

    @classmethod
    def No(cls) -> "Transformation":
        return cls()

    @property
    def translation(self) -> np.ndarray:
        return self.matrix[:3, 3]

    @property
    def rotation(self) -> np.ndarray:
        return self.matrix[:3, :3]

    def invert(self) -> "Transformation":
        return Transformation(np.linalg.inv(self.matrix))

    def transform_position(self, p: np.ndarray) -> np.ndarray:
        return self.matrix[:3, :3] @ np.asarray(p, dtype=np.float64) + self.matrix[:3, 3]

    def transform_direction(self, d: np.ndarray) -> np.ndarray:
        return self.matrix[:3, :3] @ np.asarray(d, dtype=np.float64)


class SimulatedIdealAntenna(object):
    """Marker type: isotropic antenna with [2^-1/2, 2^-1/2] polarization (core/antennas.py ideal antenna)."""


class SimulatedUniformArray(object):
    """Uniform rectangular array of identical antennas (hermespy/core/antennas.py:1299-1410).

    Element positions are ``spacing * (ix, iy, iz)`` -- not centred -- enumerated exactly like the reference's
    ``np.meshgrid(arange(nx), arange(ny), arange(nz))`` (default 'xy' indexing) flattened in C order
    (antennas.py:1344-1351): z fastest, then x, then y.
    """

    def __init__(self, element=SimulatedIdealAntenna, spacing: float = 1.0, dimensions: Tuple[int, ...] = (1, 1, 1)) -> None:
        dims = tuple(int(d) for d in dimensions) + (1,) * (3 - len(dimensions))
        self.element = element
        self.spacing = float(spacing)
        self.dimensions = dims

    @property
    def num_antennas(self) -> int:
        return int(np.prod(self.dimensions))

    num_transmit_antennas = num_antennas
    num_receive_antennas = num_antennas

    @property
    def topology(self) -> np.ndarray:
        """Element positions ``[M, 3]`` in the array frame."""
        nx, ny, nz = self.dimensions
        grid = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz))
        return self.spacing * np.vstack((grid[0].flat, grid[1].flat, grid[2].flat)).T.astype(np.float64)

    def state(self, pose: Transformation) -> "AntennaArrayState":
        return AntennaArrayState(self, pose)


@dataclass
class AntennaArrayState(object):
    """Antenna array frozen at a global pose (hermespy/core/antennas.py ``AntennaArrayState``)."""

    array: SimulatedUniformArray
    pose: Transformation

    @property
    def num_antennas(self) -> int:
        return self.array.num_antennas

    @property
    def num_transmit_antennas(self) -> int:
        return self.array.num_antennas

    @property
    def num_receive_antennas(self) -> int:
        return self.array.num_antennas

    @property
    def topology(self) -> np.ndarray:
        return self.array.topology


_device_ids = itertools.count()


@dataclass
class SimulatedDeviceState(object):
    """What a channel reads from a device at sampling time (simulated_device.py:1516-1546)."""

    device_id: int
    pose: Transformation
    velocity: np.ndarray
    antennas: AntennaArrayState
    carrier_frequency: float
    sampling_rate: float
    timestamp: float = 0.0

    @property
    def position(self) -> np.ndarray:
        return self.pose.translation


class SimulatedDevice(object):
    """Device stub: antennas + pose + velocity + sampling parameters; ``state(t)`` freezes it."""

    def __init__(self, bandwidth: float = 1.0, oversampling_factor: int = 1, carrier_frequency: float = 0.0,
                 antennas: Optional[SimulatedUniformArray] = None, pose: Optional[Transformation] = None,
                 velocity: Optional[np.ndarray] = None) -> None:
        self.bandwidth = float(bandwidth)
        self.oversampling_factor = int(oversampling_factor)
        self.carrier_frequency = float(carrier_frequency)
        self.antennas = antennas if antennas is not None else SimulatedUniformArray(SimulatedIdealAntenna, 1.0, (1, 1, 1))
        self.pose = pose if pose is not None else Transformation()
        self.velocity = np.zeros(3) if velocity is None else np.asarray(velocity, dtype=np.float64)
        self.device_id = next(_device_ids)

    @property
    def sampling_rate(self) -> float:
        return self.bandwidth * self.oversampling_factor

    def state(self, timestamp: float = 0.0) -> SimulatedDeviceState:
        # straight-line motion, as the reference's LinearTrajectory does for a constant velocity
        pose = Transformation(self.pose.matrix.copy())
        pose.matrix[:3, 3] = pose.matrix[:3, 3] + self.velocity * timestamp
        return SimulatedDeviceState(self.device_id, pose, self.velocity.copy(), self.antennas.state(pose),
                                    self.carrier_frequency, self.sampling_rate, float(timestamp))
