"""Spatially consistent random variables (host side; RNG draw order is part of parity).

Mirror of hermespy/channel/consistent.py: variables are slices of one vector of scalars per
realization (``ConsistentVariable`` :24-103, ``ConsistentGenerator.add_variable`` :353-367); a
realization is either static normals (decorrelation distance = inf, :236-251, :431-432) or a sum of
30 sinusoids in the two end-point positions (``DualConsistentRealization`` :159-200, drawn at :434-459
with the radial-velocity CDF of :272-285 tabulated by 1000 bisections, :369-408).
All randomness is drawn with numpy on the host so that draw count and order match the reference.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
from scipy.optimize import bisect
from scipy.stats import norm


class ConsistentSample(object):
    def __init__(self, scalars: np.ndarray) -> None:
        self.__scalars = scalars

    def fetch_scalars(self, offset: int, num_scalars: int) -> np.ndarray:
        return self.__scalars[offset : offset + num_scalars]

    @property
    def scalars(self) -> np.ndarray:
        return self.__scalars


class ConsistentRealization(object):
    def sample(self, position_a: np.ndarray, position_b: np.ndarray) -> ConsistentSample:
        raise NotImplementedError


class StaticConsistentRealization(ConsistentRealization):
    def __init__(self, scalar_samples: np.ndarray) -> None:
        self.scalar_samples = np.asarray(scalar_samples, dtype=np.float64).ravel()

    def sample(self, position_a, position_b) -> ConsistentSample:
        return ConsistentSample(self.scalar_samples)


class DualConsistentRealization(ConsistentRealization):
    def __init__(self, frequencies: np.ndarray, phases: np.ndarray) -> None:
        self.frequencies = frequencies  # [3, S, M, 2]
        self.phases = phases  # [S, M]

    def sample(self, position_a, position_b) -> ConsistentSample:
        pa = np.asarray(position_a, dtype=np.float64)
        pb = np.asarray(position_b, dtype=np.float64)
        M = self.frequencies.shape[2]
        arg = (np.tensordot(pa, self.frequencies[..., 0], (0, 0)) + np.tensordot(pb, self.frequencies[..., 1], (0, 0))
               + self.phases)
        return ConsistentSample((2.0 / M) ** 0.5 * np.cos(arg).sum(axis=-1))


class ConsistentVariable(object):
    def __init__(self, shape: Tuple[int, ...], offset: int = 0) -> None:
        self.shape = (1,) if shape is None else tuple(shape)
        self.size = int(np.prod(self.shape))
        self.offset = offset

    def sample(self, sample: ConsistentSample) -> np.ndarray:
        return sample.fetch_scalars(self.offset, self.size).reshape(self.shape)


class ConsistentGaussian(ConsistentVariable):
    def sample(self, sample: ConsistentSample, mean: float = 0.0, std: float = 1.0) -> np.ndarray:
        return mean + std * ConsistentVariable.sample(self, sample)


class ConsistentUniform(ConsistentVariable):
    def sample(self, sample: ConsistentSample) -> np.ndarray:
        return norm.cdf(ConsistentVariable.sample(self, sample))


class ConsistentBoolean(ConsistentVariable):
    def sample(self, sample: ConsistentSample) -> np.ndarray:
        return ConsistentVariable.sample(self, sample) > 0.0


def _radial_velocity_cdf(fr: float, a: float, u: float) -> float:
    return 2 / np.pi * np.arctan(2 * np.pi * fr / a) - 4 * a * fr / (4 * np.pi**2 * fr**2 + a**2) - u


class ConsistentGenerator(object):
    """Allocates variables and draws realizations from the owner's numpy generator."""

    _cdf_cache: Dict[Tuple[float, int], np.ndarray] = dict()

    def __init__(self, rng) -> None:
        self.__rng = rng  # np.random.Generator or an object exposing ``_rng``
        self.__offset = 0
        self.__variables: List[ConsistentVariable] = []

    def _add(self, variable: ConsistentVariable) -> ConsistentVariable:
        variable.offset = self.__offset
        self.__offset += variable.size
        self.__variables.append(variable)
        return variable

    def gaussian(self, shape: Optional[Tuple[int, ...]] = None) -> ConsistentGaussian:
        return self._add(ConsistentGaussian((1,) if shape is None else shape))

    def uniform(self, shape: Optional[Tuple[int, ...]] = None) -> ConsistentUniform:
        return self._add(ConsistentUniform((1,) if shape is None else shape))

    def boolean(self, shape: Optional[Tuple[int, ...]] = None) -> ConsistentBoolean:
        return self._add(ConsistentBoolean((1,) if shape is None else shape))

    @property
    def num_scalars(self) -> int:
        return self.__offset

    @classmethod
    def _sample_cdf(cls, decorrelation_distance: float, num_samples: int = 1000) -> np.ndarray:
        key = (decorrelation_distance, num_samples)
        hit = cls._cdf_cache.get(key)
        if hit is not None:
            return hit
        u_candidates = np.linspace(0, 1, 1 + num_samples, endpoint=True, dtype=np.float64)[:-1]
        out = np.empty(num_samples, dtype=np.float64)
        a = 1 / decorrelation_distance
        fr_max = 1
        for i, u in enumerate(u_candidates):
            while _radial_velocity_cdf(fr_max, a, u) < 0:
                fr_max *= 2
            out[i] = bisect(_radial_velocity_cdf, 0, fr_max, args=(a, u))
        cls._cdf_cache[key] = out
        return out

    def realize(self, decorrelation_distance: float, num_sinusoids: int = 30) -> ConsistentRealization:
        S = self.__offset
        rng = self.__rng if isinstance(self.__rng, np.random.Generator) else self.__rng._rng
        if decorrelation_distance == float("inf"):
            return StaticConsistentRealization(rng.standard_normal(S))
        dims = (S, num_sinusoids, 2)
        # draw order: radial velocities, azimuth, zenith, phases (consistent.py:436-458)
        fr = rng.choice(self._sample_cdf(decorrelation_distance), size=dims)
        az = rng.uniform(0, 2 * np.pi, size=dims)
        ze = np.arccos(1 - rng.uniform(0, 2, size=dims))
        sz = np.sin(ze)
        direction = np.array([np.cos(az) * sz, np.sin(az) * sz, np.cos(ze)])
        frequencies = 2 * np.pi * fr * direction
        phases = rng.uniform(0, 2 * np.pi, size=(S, num_sinusoids))
        return DualConsistentRealization(frequencies, phases)
