"""Antenna correlation models (hermespy/channel/fading/correlation.py:20-140, fading.py:41-144).

``sample_covariance(antennas, mode)`` returns the covariance matrix that is multiplied onto the
spatial response from the left (RX) or right (TX) -- the matrices themselves, not their square
roots (fading.py:480-489).
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from enum import Enum

import numpy as np

from ...core import AntennaMode


class DeviceType(Enum):
    BASE_STATION = 0
    TERMINAL = 1


class CorrelationType(Enum):
    """(base-station factor, terminal factor) per 3GPP correlation level (correlation.py:30-44)."""

    LOW = (0.0, 0.0)
    MEDIUM = (0.3, 0.3)
    MEDIUM_A = (0.3, 0.3874)
    HIGH = (0.9, 0.9)

    @classmethod
    def from_parameters(cls, value) -> "CorrelationType":
        if isinstance(value, cls):
            return value
        if isinstance(value, str):
            return cls[value.upper()]
        return cls(value)


def _count(antennas, mode: AntennaMode) -> int:
    if isinstance(antennas, (int, np.integer)):
        return int(antennas)
    return antennas.num_transmit_antennas if mode == AntennaMode.TX else antennas.num_receive_antennas


class AntennaCorrelation(ABC):
    def __init__(self, channel=None, device=None) -> None:
        self.channel = channel
        self.device = device

    @abstractmethod
    def sample_covariance(self, antennas, mode: AntennaMode) -> np.ndarray:
        ...


class CustomAntennaCorrelation(AntennaCorrelation):
    """User-supplied hermitian positive-definite covariance (fading.py:84-144)."""

    def __init__(self, covariance: np.ndarray, **kwargs) -> None:
        AntennaCorrelation.__init__(self, **kwargs)
        self.covariance = covariance

    @property
    def covariance(self) -> np.ndarray:
        return self.__covariance

    @covariance.setter
    def covariance(self, value: np.ndarray) -> None:
        value = np.asarray(value)
        if value.ndim != 2 or not np.allclose(value, value.T.conj()):
            raise ValueError("Antenna correlation must be a hermitian matrix")
        if np.any(np.linalg.eigvals(value) <= 0.0):
            raise ValueError("Antenna correlation matrix must be positive definite")
        self.__covariance = value

    def sample_covariance(self, antennas, mode: AntennaMode) -> np.ndarray:
        n = _count(antennas, mode)
        if self.__covariance.shape[0] < n:
            raise ValueError("Antenna correlation matrix does not match the number of antennas")
        return self.__covariance[:n, :n]


class StandardAntennaCorrelation(AntennaCorrelation):
    """3GPP standardized correlations for 1, 2 and 4 antennas (correlation.py:47-106).

    As in the reference the receiving side always uses the TERMINAL factor and the transmitting side the
    BASE_STATION factor.
    """

    def __init__(self, correlation, **kwargs) -> None:
        self.correlation = CorrelationType.from_parameters(correlation)
        AntennaCorrelation.__init__(self, **kwargs)

    def sample_covariance(self, antennas, mode: AntennaMode) -> np.ndarray:
        side = DeviceType.TERMINAL if mode == AntennaMode.RX else DeviceType.BASE_STATION
        f = self.correlation.value[side.value]
        n = _count(antennas, mode)
        if n == 1:
            return np.ones((1, 1), dtype=complex)
        if n == 2:
            return np.array([[1, f], [f, 1]], dtype=complex)
        if n == 4:
            a, b = f ** (1 / 9), f ** (4 / 9)
            return np.array([[1, a, b, f], [a, 1, a, b], [b, a, 1, a], [f, b, a, 1]], dtype=complex)
        raise RuntimeError(
            f"3GPP standard antenna covariance is only defined for 1, 2 and 4 antennas, device has {n} antennas"
        )
