"""Standard parameterizations of the multipath fading channel.

* ``TDL``          5G tapped-delay-line models A-E      (hermespy/channel/fading/tdl.py:40-397)
* ``Cost259``      COST 259 urban / rural / hilly       (hermespy/channel/fading/cost259.py:42-227)
* ``Exponential``  exponentially decaying profile       (hermespy/channel/fading/exponential.py:29-117)

Only the constructor differs from :class:`MultipathFadingChannel`; tables live in ``profiles.py``.
"""
from __future__ import annotations

from enum import Enum
from typing import Optional

import numpy as np

from .correlation import AntennaCorrelation
from .fading import MultipathFadingChannel
from .profiles import COST259_PROFILES, TDL_PROFILES

_INF = MultipathFadingChannel._DEFAULT_DECORRELATION_DISTANCE
_NSIN = MultipathFadingChannel._DEFAULT_NUM_SINUSOIDS
_DOPP = MultipathFadingChannel._DEFAULT_DOPPLER_FREQUENCY


class TDLType(Enum):
    """Enum values as in tdl.py:25-29 (note the gap: D = 4, E = 5)."""

    A = 0
    B = 1
    C = 2
    D = 4
    E = 5


class TDL(MultipathFadingChannel):
    """5G TDL multipath fading channel."""

    def __init__(self, model_type: TDLType = TDLType.A, rms_delay: float = 0.0, correlation_distance: float = _INF,
                 num_sinusoids: int = _NSIN, los_angle: Optional[float] = None, doppler_frequency: float = _DOPP,
                 los_doppler_frequency: Optional[float] = None, antenna_correlation: Optional[AntennaCorrelation] = None,
                 gain: float = 1.0, seed: Optional[int] = None, **kwargs) -> None:
        if rms_delay < 0.0:
            raise ValueError("Root-Mean-Squared delay must be greater or equal to zero")
        try:
            model_type = TDLType(model_type) if not isinstance(model_type, TDLType) else model_type
            profile = TDL_PROFILES[model_type.value]
        except (ValueError, KeyError):
            raise ValueError("Requested model type not supported")
        if profile["los_doppler"] is not None:
            # models D and E fix the line-of-sight Doppler (tdl.py:276-279, 322, 325-328, 369)
            if los_doppler_frequency is not None:
                raise ValueError(
                    f"Model type {model_type.name} does not support line of sight doppler frequency configuration")
            los_doppler_frequency = profile["los_doppler"]
        self.__model_type = model_type
        self.__rms_delay = rms_delay
        MultipathFadingChannel.__init__(self, rms_delay * profile["delay"], profile["power"].copy(),
                                        profile["rice"].copy(), correlation_distance, num_sinusoids, los_angle,
                                        doppler_frequency, los_doppler_frequency, antenna_correlation, gain, seed,
                                        **kwargs)

    model_type = property(lambda self: self.__model_type)
    rms_delay = property(lambda self: self.__rms_delay)


class Cost259Type(Enum):
    URBAN = 0
    RURAL = 1
    HILLY = 2


class Cost259(MultipathFadingChannel):
    """COST 259 multipath fading channel."""

    def __init__(self, model_type: Cost259Type = Cost259Type.URBAN, correlation_distance: float = _INF,
                 num_sinusoids: int = _NSIN, los_angle: Optional[float] = None, doppler_frequency: float = _DOPP,
                 los_doppler_frequency: Optional[float] = None, antenna_correlation: Optional[AntennaCorrelation] = None,
                 gain: float = 1.0, seed: Optional[int] = None, **kwargs) -> None:
        try:
            model_type = Cost259Type(model_type) if not isinstance(model_type, Cost259Type) else model_type
            profile = COST259_PROFILES[model_type.value]
        except (ValueError, KeyError):
            raise ValueError("Requested model type not supported")
        if model_type == Cost259Type.HILLY:
            # cost259.py:202 overrides the argument silently (the docstring promises a ValueError, the code
            # does not raise); the value never reaches the math
            los_angle = np.arccos(0.7)
        self.__model_type = model_type
        MultipathFadingChannel.__init__(self, profile["delay"].copy(), profile["power"].copy(), profile["rice"].copy(),
                                        correlation_distance, num_sinusoids, los_angle, doppler_frequency,
                                        los_doppler_frequency, antenna_correlation, gain, seed, **kwargs)

    model_type = property(lambda self: self.__model_type)


class Exponential(MultipathFadingChannel):
    """Exponentially decaying power-delay profile, truncated at 1e-5 (exponential.py:29-117)."""

    _TRUNCATION = 1e-5

    def __init__(self, tap_interval: float, rms_delay: float, correlation_distance: float = _INF,
                 num_sinusoids: int = _NSIN, los_angle: Optional[float] = None, doppler_frequency: float = _DOPP,
                 los_doppler_frequency: Optional[float] = None, antenna_correlation: Optional[AntennaCorrelation] = None,
                 gain: float = 1.0, seed: Optional[int] = None, **kwargs) -> None:
        if tap_interval <= 0.0:
            raise ValueError("Tap interval must be greater than zero")
        if rms_delay <= 0.0:
            raise ValueError("Root-Mean-Squared delay must be greater than zero")
        self.__tap_interval = tap_interval
        self.__rms_delay = rms_delay
        rms_norm = rms_delay / tap_interval
        # decay exponent of an infinite geometric profile with that RMS delay (exponential.py:96-100)
        alpha = -2 * np.log((-1 + np.sqrt(1 + 4 * rms_norm**2)) / (2 * rms_norm))
        last = int(-np.ceil(np.log(Exponential._TRUNCATION) / alpha))
        taps = np.arange(last + 1)
        MultipathFadingChannel.__init__(self, taps * tap_interval, np.exp(-alpha * taps), np.zeros(taps.shape),
                                        correlation_distance, num_sinusoids, los_angle, doppler_frequency,
                                        los_doppler_frequency, antenna_correlation, gain, seed, **kwargs)

    tap_interval = property(lambda self: self.__tap_interval)
    rms_delay = property(lambda self: self.__rms_delay)
