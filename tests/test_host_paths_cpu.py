"""CPU: error paths and small host-side behaviours fixed after the round-1 review."""
import numpy as np
import pytest
import torch

from hermespy_b200 import _lib


def test_spatial_gemm_refuses_cpu_tensors_with_the_documented_error():
    from hermespy_b200.kernels import spatial_gemm

    with pytest.raises(_lib.HermesB200Error) as e:
        spatial_gemm(torch.zeros((1, 16, 16), dtype=torch.complex128), torch.zeros((1, 16, 64), dtype=torch.complex64))
    assert e.value.status == _lib.HB_ERR_NO_DEVICE


def test_cdl_propagate_refuses_cpu_tensors():
    from hermespy_b200.kernels import cdl_propagate

    with pytest.raises(_lib.HermesB200Error):
        cdl_propagate(torch.zeros((1, 2, 8), dtype=torch.complex64), None)


def test_set_device_without_gpu_is_a_hard_error():
    if _lib.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(_lib.HermesB200Error) as e:
        _lib.set_device(0)
    assert e.value.status == _lib.HB_ERR_NO_DEVICE
    _lib.set_device(None)  # None = leave the thread's device alone


def test_dropin_enable_without_gpu_is_a_hard_error():
    if _lib.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    import hermespy_b200.dropin as dropin

    with pytest.raises(_lib.HermesB200Error):
        dropin.enable()


def test_ray_wait_edge_cases_and_pool_release():
    from hermespy_b200.shims import ray as shim

    assert shim.wait([], num_returns=1) == ([], [])
    ref = shim.put(3)
    ready, pending = shim.wait([ref], num_returns=5)  # more than available: everything that exists
    assert ready == [ref] and pending == []

    class Actor:
        def f(self, x):
            return x + 1

    import gc

    gc.collect()  # actors of earlier Simulation.run() calls retire here, not inside the measurement
    n0 = len(shim._live_pools)
    h = shim.remote(Actor).remote()
    assert shim.get(h.f.remote(1)) == 2 and len(shim._live_pools) == n0 + 1
    del h

    gc.collect()
    assert len(shim._live_pools) == n0
    h = shim.remote(Actor).remote()
    shim.shutdown()
    assert not shim._live_pools and not shim.is_initialized()


def test_confidence_rule_is_gated_on_min_num_samples():
    """scalar.py:117: the stopping rule runs only when count % min_num_samples == 0."""
    from hermespy_b200.montecarlo import GridStatistics

    st = GridStatistics((1,))
    x = np.array([0.10, 0.11, 0.09, 0.10, 0.105, 0.095])
    st.stats[0, 0], st.stats[0, 1], st.stats[0, 2] = float(x.sum()), float((x ** 2).sum()), len(x)
    assert st.confident(0, accuracy=0.05, confidence=0.5, min_num_samples=1)
    assert st.confident(0, accuracy=0.05, confidence=0.5, min_num_samples=3)
    assert not st.confident(0, accuracy=0.05, confidence=0.5, min_num_samples=4)


def test_cdl_block_validates_element_tables():
    from hermespy_b200.kernels import CdlBlock, ideal_elements

    base = dict(term_delay=np.zeros(2, np.int32), max_delay=0, angles=np.zeros((1, 2, 4)), jones=np.zeros((1, 2, 2, 2), complex),
                amplitude=np.ones((1, 2)), tx_pose=np.zeros((1, 12)), rx_pose=np.zeros((1, 12)), rel_velocity=np.zeros((1, 3)),
                tx_topology=np.zeros((3, 3)), rx_topology=np.zeros((2, 3)), carrier_frequency=1e9, sampling_rate=1e6)
    assert CdlBlock(**base).element_mode == _lib.HB_ELEMENTS_IDEAL
    dip = ideal_elements(3)
    dip[:, 9] = _lib.HB_ELEMENT_DIPOLE
    b = CdlBlock(**base, tx_elements=dip)  # the other side becomes a table of ideal elements
    assert b.element_mode == _lib.HB_ELEMENTS_UNIFORM and b.rx_elements.shape == (2, 12)
    dip[1, 10] = 0.3
    assert CdlBlock(**base, tx_elements=dip).element_mode == _lib.HB_ELEMENTS_PER_ELEMENT
    with pytest.raises(ValueError):
        CdlBlock(**base, tx_elements=ideal_elements(4), rx_elements=ideal_elements(2))


def test_bench_time_limit_turns_a_hang_into_an_error():
    """bench.py runs the Simulation.run() scripts under a SIGALRM limit: a leg that hangs becomes an error entry of the JSON
    line instead of a bench run that never prints it."""
    import signal
    import time

    import bench

    assert bench._time_limited(5, lambda a, b=1: a + b, 2, b=3) == 5
    before = signal.getsignal(signal.SIGALRM)
    t0 = time.perf_counter()
    with pytest.raises(TimeoutError, match="1 s limit"):
        bench._time_limited(1, time.sleep, 30)
    assert time.perf_counter() - t0 < 5
    assert signal.getsignal(signal.SIGALRM) == before and signal.alarm(0) == 0  # handler restored, no alarm left pending
