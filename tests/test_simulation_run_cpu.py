"""Host logic (no GPU): the reference's UNMODIFIED ``Simulation.run()`` executes in this process through the in-process
``ray`` stand-in (``hermespy_b200/shims/ray.py``) -- queue manager, simulation actor and collector of
``hermespy/core/pymonte`` included -- and the stand-in honours the semantics the engine relies on."""
import threading
import time

import numpy as np
import pytest

from hermespy_b200.shims import ray as shim


def test_actor_semantics():
    class Counter(object):
        def __init__(self, payload, peer=None):
            self.payload, self.peer, self.value = payload, peer, 0

        def add(self, n):
            time.sleep(0.01)
            self.value += n
            return self.value

        def thread(self):
            return threading.get_ident()

    payload = {"a": [1, 2, 3]}
    a = shim.remote(Counter).options(num_cpus=0).remote(payload)
    b = shim.remote(Counter).options(max_concurrency=2).remote(payload, peer=a)
    assert a._instance.payload == payload and a._instance.payload is not payload  # arguments are copied ...
    assert b._instance.peer is a                                                  # ... actor handles are not
    refs = [a.add.remote(1) for _ in range(5)]
    assert shim.get(refs) == [1, 2, 3, 4, 5]  # calls on one actor are serialized in submission order
    assert shim.get(a.thread.remote()) != threading.get_ident()
    ready, pending = shim.wait([a.add.remote(1), shim.put(7)], num_returns=1)
    assert len(ready) == 1 and len(pending) == 1
    assert shim.get(shim.put(np.arange(3))).tolist() == [0, 1, 2]
    shim.init(logging_level=0)
    assert shim.is_initialized() and shim.available_resources()["CPU"] >= 1


def test_unmodified_simulation_run():
    from oracle.refload import load_reference, reference_available

    if not reference_available():
        pytest.skip("reference tree not available")
    load_reference()
    import ray

    if getattr(ray, "__version__", "") != shim.__version__:
        pytest.skip("a real ray is installed")
    from hermespy.channel import TDL
    from hermespy.core import ConsoleMode, dB
    from hermespy.modem import (BitErrorEvaluator, RootRaisedCosineWaveform, SimplexLink,
                                SingleCarrierLeastSquaresChannelEstimation, SingleCarrierZeroForcingChannelEqualization)
    from hermespy.simulation import SNR, Simulation

    # _examples/getting_started/simulation.py, minus the plots
    simulation = Simulation(console_mode=ConsoleMode.SILENT, num_samples=8, seed=42)
    tx_device = simulation.new_device(oversampling_factor=4)
    rx_device = simulation.new_device(oversampling_factor=4)
    tx_device.noise_level = SNR(dB(20), tx_device)
    rx_device.noise_level = SNR(dB(20), tx_device)
    simulation.set_channel(tx_device, rx_device, TDL())
    link = SimplexLink()
    tx_device.transmitters.add(link)
    rx_device.receivers.add(link)
    link.waveform = RootRaisedCosineWaveform(num_preamble_symbols=10, num_data_symbols=100, roll_off=0.9)
    link.waveform.channel_estimation = SingleCarrierLeastSquaresChannelEstimation()
    link.waveform.channel_equalization = SingleCarrierZeroForcingChannelEqualization()
    simulation.new_dimension("noise_level", dB(20, 10, 0), rx_device)
    simulation.add_evaluator(BitErrorEvaluator(link, link))
    result = simulation.run()
    ber = np.asarray(result.evaluation_results[0].to_array(), dtype=float).ravel()
    assert ber.shape == (3,) and np.all((ber >= 0) & (ber <= 0.5 + 1e-9))
    assert ber[0] < ber[2]  # 20 dB beats 0 dB
