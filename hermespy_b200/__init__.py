"""hermespy_b200 -- B200-native (sm_100a) implementation of HermesPy's Monte-Carlo channel hot path.

Scope: stochastic multipath-fading (TDL / COST259 / Exponential) and 3GPP CDL channel
realization + signal propagation, batched over drops and sweep points, behind the reference's
``Channel.realize() -> ChannelRealization.sample() -> ChannelSample.propagate(Signal)`` plugin API.
The arithmetic runs in hand-written CUDA kernels reached through the C-ABI in ``include/hermes_b200.h``;
there is no CPU fallback.
"""
__version__ = "0.1.0"
