from .correlation import (AntennaCorrelation, CorrelationType, CustomAntennaCorrelation, DeviceType,
                          StandardAntennaCorrelation)
from .fading import (FadingChannelState, MultipathFadingChannel, MultipathFadingRealization,
                     MultipathFadingSample, propagate_batch)
from .templates import TDL, Cost259, Cost259Type, Exponential, TDLType

__all__ = [
    "AntennaCorrelation", "CorrelationType", "CustomAntennaCorrelation", "DeviceType",
    "StandardAntennaCorrelation", "FadingChannelState", "MultipathFadingChannel",
    "MultipathFadingRealization", "MultipathFadingSample", "propagate_batch", "TDL", "TDLType", "Cost259",
    "Cost259Type", "Exponential",
]
