"""Channel models of the hot path (mirror of ``hermespy.channel`` for fading and CDL)."""
from .channel import Channel, ChannelRealization, ChannelSample, ChannelSampleHook, LinkState
from .consistent import (ConsistentBoolean, ConsistentGaussian, ConsistentGenerator, ConsistentUniform,
                         DualConsistentRealization, StaticConsistentRealization)
from .fading import *  # noqa: F401,F403
from .fading import __all__ as _fading_all
from .cdl import *  # noqa: F401,F403
from .cdl import __all__ as _cdl_all

__all__ = [
    "Channel", "ChannelRealization", "ChannelSample", "ChannelSampleHook", "LinkState", "ConsistentBoolean",
    "ConsistentGaussian", "ConsistentGenerator", "ConsistentUniform", "DualConsistentRealization",
    "StaticConsistentRealization",
] + list(_fading_all) + list(_cdl_all)
