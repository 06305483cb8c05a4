"""GPU: a3 on the device (``hb_fading_sample`` + ``hb_kron_mix``) against the host sampler, which is pinned to the reference
(tests/test_batch_sampling_cpu.py: row b equals the b-th sequential ``realize().sample()``).  Same generator seed -> the same
normals; parameters agree to a few ulp (device erfc / cos / sincospi vs scipy ndtr / numpy), propagated signals to 1e-12."""
import numpy as np
import pytest

import hermespy_b200.channel as MC
from hermespy_b200.batch import sample_fading_links, sample_fading_links_device
from tests.test_batch_sampling_cpu import BUILDERS, FS

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("reciprocal", [False, True])
@pytest.mark.parametrize("name", list(BUILDERS))
def test_device_sampling_equals_host_sampling(name, reciprocal):
    from hermespy_b200 import _lib

    build, ntx, nrx = BUILDERS[name]
    B = 7
    host = sample_fading_links(build(MC, 42), B, ntx, nrx, FS, reciprocal=reciprocal)
    before = _lib.launch_counts()["misc"]
    dev = sample_fading_links_device(build(MC, 42), B, ntx, nrx, FS, reciprocal=reciprocal)
    assert _lib.launch_counts()["misc"] > before
    assert np.array_equal(dev.tap_delay, host["tap_delay"]) and dev.max_delay == host["max_delay"]
    assert dev.omega_max == host["omega_max"]
    for k in ("omega", "phi", "amp", "spatial"):
        got, want = getattr(dev, k).cpu().numpy(), host[k]
        assert got.shape == want.shape, k
        scale = max(1e-300, float(np.abs(want).max()))
        # angles near +-pi: 2 pi Phi(g) is formed from erfc with ~1 ulp of relative error in u, i.e. 1e-16 * 2 pi absolute
        assert np.abs(got - want).max() <= 4e-15 * max(scale, 1.0), (k, np.abs(got - want).max())


def test_device_sampling_large_array_with_correlation_and_propagation():
    """C4 ingredients: 64 x 64 antenna phases + exponential Kronecker correlation on the device, then propagation: equal to
    the host-sampled block through the same kernels to 1e-12 (f64) -- the sampler, not the kernels, is under test."""
    import torch

    from hermespy_b200.kernels import FadingBatch, fading_propagate

    n, B, T = 64, 2, 512
    R = 0.7 ** np.abs(np.subtract.outer(np.arange(n), np.arange(n))).astype(complex)
    build = lambda: MC.TDL(MC.TDLType.D, rms_delay=300e-9, doppler_frequency=100, seed=3, max_antennas=n,
                           antenna_correlation=MC.CustomAntennaCorrelation(R))
    host = sample_fading_links(build(), B, n, n, FS)
    dev = sample_fading_links_device(build(), B, n, n, FS)
    want = FadingBatch.from_numpy(device="cuda", **host)  # applies R_rx S R_tx on the device as well (K2)
    assert np.abs(dev.spatial.cpu().numpy() - want.spatial.cpu().numpy()).max() <= 1e-12 * np.abs(want.spatial.cpu().numpy()).max()
    rng = np.random.default_rng(0)
    x = torch.from_numpy((rng.standard_normal((B, n, T)) + 1j * rng.standard_normal((B, n, T))) / np.sqrt(2)).cuda()
    y_dev = fading_propagate(x, dev, precision="f64").cpu().numpy()
    y_host = fading_propagate(x, want, precision="f64").cpu().numpy()
    assert np.linalg.norm(y_dev - y_host) <= 1e-12 * np.linalg.norm(y_host)
