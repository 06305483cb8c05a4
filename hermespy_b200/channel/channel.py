"""Channel plugin interface: ``Channel.realize() -> ChannelRealization.sample() -> ChannelSample.propagate()``.

Host-side mirror of hermespy/channel/channel.py (same method names, argument meaning and error
behaviour): ``LinkState`` (:107-178), ``ChannelSample`` (:181-417, ``propagate`` :306-379),
``ChannelRealization`` (:420-727, ``sample`` :530-577, ``reciprocal_sample`` :644-701) and
``Channel`` (:730-986, ``realize`` :877-898, ``propagate`` :910-973).  Subclasses implement
``_realize`` / ``_sample`` / ``_reciprocal_sample`` / ``_propagate`` / ``state`` exactly like
reference plugins do; the arithmetic behind ``_propagate`` / ``state`` runs on the GPU.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Callable, Generic, Optional, Set, TypeVar

import numpy as np

from ..core import InterpolationMode, Signal, SignalBlock, SimulatedDevice, SimulatedDeviceState

CST = TypeVar("CST", bound="ChannelSample")
CRT = TypeVar("CRT", bound="ChannelRealization")


class LinkState(object):
    """Physical parameters of one directed link at sampling time (channel.py:107-178)."""

    def __init__(self, transmitter: SimulatedDeviceState, receiver: SimulatedDeviceState, carrier_frequency: float,
                 bandwidth: float, time: float) -> None:
        self.__transmitter = transmitter
        self.__receiver = receiver
        self.__carrier_frequency = float(carrier_frequency)
        self.__bandwidth = float(bandwidth)
        self.__time = float(time)

    transmitter = property(lambda self: self.__transmitter)
    receiver = property(lambda self: self.__receiver)
    carrier_frequency = property(lambda self: self.__carrier_frequency)
    bandwidth = property(lambda self: self.__bandwidth)  # == sampling rate of the propagated signal
    time = property(lambda self: self.__time)


class ChannelSampleHook(Generic[CST]):
    """Callback fired after a sample is generated, optionally filtered by device (channel.py:44-104)."""

    def __init__(self, callback: Callable[[CST], None], transmitter=None, receiver=None) -> None:
        self.__callback = callback
        self.__transmitter = transmitter
        self.__receiver = receiver

    def __call__(self, sample: CST, transmitter, receiver) -> None:
        def _id(d):
            return d.device_id if hasattr(d, "device_id") else d

        if self.__transmitter is not None and _id(self.__transmitter) != _id(transmitter):
            return
        if self.__receiver is not None and _id(self.__receiver) != _id(receiver):
            return
        self.__callback(sample)


class ChannelSample(ABC):
    """Immutable sample of a channel in time and space (channel.py:181-417)."""

    def __init__(self, state: LinkState) -> None:
        self.__state = state

    link_state = property(lambda self: self.__state)
    transmitter_state = property(lambda self: self.__state.transmitter)
    receiver_state = property(lambda self: self.__state.receiver)
    transmitter_pose = property(lambda self: self.__state.transmitter.pose)
    receiver_pose = property(lambda self: self.__state.receiver.pose)
    transmitter_velocity = property(lambda self: self.__state.transmitter.velocity)
    receiver_velocity = property(lambda self: self.__state.receiver.velocity)
    transmitter_antennas = property(lambda self: self.__state.transmitter.antennas)
    receiver_antennas = property(lambda self: self.__state.receiver.antennas)
    num_transmit_antennas = property(lambda self: self.__state.transmitter.antennas.num_transmit_antennas)
    num_receive_antennas = property(lambda self: self.__state.receiver.antennas.num_receive_antennas)
    carrier_frequency = property(lambda self: self.__state.carrier_frequency)
    bandwidth = property(lambda self: self.__state.bandwidth)
    time = property(lambda self: self.__state.time)

    @property
    @abstractmethod
    def expected_energy_scale(self) -> float:
        ...

    @abstractmethod
    def _propagate(self, signal: SignalBlock, interpolation: InterpolationMode) -> SignalBlock:
        ...

    @abstractmethod
    def state(self, num_samples: int, max_num_taps: int,
              interpolation_mode: InterpolationMode = InterpolationMode.NEAREST):
        ...

    def propagate(self, signal: Signal, interpolation_mode: InterpolationMode = InterpolationMode.NEAREST) -> Signal:
        """Propagate a signal over this sample, block by block (channel.py:306-379)."""
        if hasattr(signal, "mixed_signal"):  # DeviceOutput
            signal = signal.mixed_signal
        if not isinstance(signal, Signal):
            raise ValueError("Signal is of unsupported type")
        if self.expected_energy_scale <= 0.0:
            return Signal.Empty(signal.sampling_rate, self.num_receive_antennas, 0,
                                carrier_frequency=signal.carrier_frequency, noise_power=signal.noise_power,
                                delay=signal.delay)
        if signal.num_streams != self.num_transmit_antennas:
            raise ValueError(
                "Number of signal streams to be propagated does not match the number of transmitter antennas "
                f"({signal.num_streams} != {self.num_transmit_antennas}))"
            )
        blocks = [self._propagate(b, interpolation_mode) for b in signal.blocks]
        return Signal.Create(blocks, self.bandwidth, self.carrier_frequency, signal.noise_power, signal.delay,
                             offsets=[b.offset for b in blocks])


class ChannelRealization(ABC, Generic[CST]):
    """Realization of all random processes of a channel model (channel.py:420-727)."""

    _DEFAULT_GAIN = 1.0

    def __init__(self, sample_hooks: Optional[Set[ChannelSampleHook]] = None, gain: float = _DEFAULT_GAIN) -> None:
        self.__sample_hooks = set() if sample_hooks is None else sample_hooks
        self.__gain = gain

    @property
    def sample_hooks(self) -> Set[ChannelSampleHook]:
        return self.__sample_hooks.copy()

    @property
    def gain(self) -> float:
        return self.__gain

    @staticmethod
    def _resolve(transmitter, receiver, timestamp):
        if isinstance(transmitter, SimulatedDevice) and isinstance(receiver, SimulatedDevice):
            return transmitter.state(timestamp), receiver.state(timestamp)
        if isinstance(transmitter, SimulatedDeviceState) and isinstance(receiver, SimulatedDeviceState):
            return transmitter, receiver
        raise ValueError("Invalid input argument types for channel sampling.")

    def sample(self, transmitter, receiver, timestamp: float = 0.0, carrier_frequency: Optional[float] = None,
               bandwidth: Optional[float] = None) -> CST:
        """Sample the realization for a directed link (channel.py:530-577)."""
        tx, rx = self._resolve(transmitter, receiver, timestamp)
        fc = tx.carrier_frequency if carrier_frequency is None else float(carrier_frequency)
        bw = tx.sampling_rate if bandwidth is None else float(bandwidth)
        sample = self._sample(LinkState(tx, rx, fc, bw, timestamp))
        for hook in self.sample_hooks:
            hook(sample, tx.device_id, rx.device_id)
        return sample

    def reciprocal_sample(self, sample: CST, transmitter, receiver, *args) -> CST:
        """Sample the reverse direction of ``sample`` (channel.py:644-701)."""
        if isinstance(transmitter, SimulatedDevice):
            timestamp = float(args[0]) if len(args) > 0 else 0.0
            rest = args[1:]
        else:
            timestamp = sample.time
            rest = args
        tx, rx = self._resolve(transmitter, receiver, timestamp)
        fc = float(rest[0]) if len(rest) > 0 and rest[0] is not None else tx.carrier_frequency
        bw = float(rest[1]) if len(rest) > 1 and rest[1] is not None else tx.sampling_rate
        out = self._reciprocal_sample(sample, LinkState(tx, rx, fc, bw, timestamp))
        for hook in self.sample_hooks:
            hook(out, tx.device_id, rx.device_id)
        return out

    @abstractmethod
    def _sample(self, state: LinkState) -> CST:
        ...

    @abstractmethod
    def _reciprocal_sample(self, sample: CST, state: LinkState) -> CST:
        ...


class Channel(ABC, Generic[CRT, CST]):
    """Base class of channel models (channel.py:730-986).

    ``seed`` initialises a private ``numpy`` generator (the reference's ``RandomNode`` root behaviour,
    hermespy/core/random_node.py:69-124); without a seed the channel draws from ``scenario_rng`` when one
    has been attached (shared scenario generator) or from a fresh unseeded generator.
    """

    _DEFAULT_GAIN = 1.0

    def __init__(self, gain: float = _DEFAULT_GAIN, seed: Optional[int] = None) -> None:
        self.__seed = seed
        self.__generator = np.random.default_rng(seed) if seed is not None else None
        self.scenario_rng: Optional[np.random.Generator] = None
        self.gain = gain
        self.__sample_hooks: Set[ChannelSampleHook] = set()
        self.__last_realization = None

    @property
    def seed(self) -> Optional[int]:
        return self.__seed

    @seed.setter
    def seed(self, value: Optional[int]) -> None:
        self.__seed = value
        self.__generator = np.random.default_rng(value)

    @property
    def _rng(self) -> np.random.Generator:
        if self.__generator is not None:
            return self.__generator
        if self.scenario_rng is None:
            self.scenario_rng = np.random.default_rng()
        return self.scenario_rng

    @property
    def gain(self) -> float:
        return self.__gain

    @gain.setter
    def gain(self, value: float) -> None:
        if value < 0.0:
            raise ValueError("Channel gain must be greater or equal to zero")
        self.__gain = value

    @property
    def sample_hooks(self) -> Set[ChannelSampleHook]:
        return self.__sample_hooks.copy()

    def add_sample_hook(self, callback, transmitter=None, receiver=None) -> ChannelSampleHook:
        hook = ChannelSampleHook(callback, transmitter, receiver)
        self.__sample_hooks.add(hook)
        return hook

    def remove_sample_hook(self, hook: ChannelSampleHook) -> None:
        self.__sample_hooks.discard(hook)

    @abstractmethod
    def _realize(self) -> CRT:
        ...

    def realize(self, cache: bool = True) -> CRT:
        realization = self._realize()
        if cache:
            self.__last_realization = realization
        return realization

    @property
    def realization(self) -> Optional[CRT]:
        return self.__last_realization

    def propagate(self, signal: Signal, transmitter, receiver, timestamp: float = 0.0,
                  interpolation_mode: InterpolationMode = InterpolationMode.NEAREST) -> Signal:
        """realize -> sample -> propagate convenience (channel.py:910-973)."""
        realization = self.realize()
        sample = realization.sample(transmitter, receiver, timestamp, signal.carrier_frequency, signal.sampling_rate)
        return sample.propagate(signal, interpolation_mode)
