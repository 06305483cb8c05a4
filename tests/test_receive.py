"""Receive superposition + AWGN: oracle pinned against the live reference on the CPU, the fused kernel
``hb_receive_combine`` against the oracle on the GPU -- complex128 bit for bit (integer-offset superposition and the
``sqrt(P/2) (n_re + j n_im)`` noise of hermespy/simulation/rf/noise/model.py:140-160 are exact operations)."""
import numpy as np
import pytest

from oracle import receive_oracle as ro
from oracle.refload import reference_available


def _case(rng, nrx, lens, offsets):
    sigs = [(rng.standard_normal((nrx, T)) + 1j * rng.standard_normal((nrx, T))) for T in lens]
    return sigs, list(offsets)


@pytest.mark.reference
@pytest.mark.skipif(not reference_available(), reason="reference tree not mounted")
def test_oracle_equals_reference_superposition_and_awgn():
    from oracle.refload import load_reference

    load_reference()
    from hermespy.core import Signal
    from hermespy.core.signal_model import SparseSignal
    from hermespy.simulation.rf.noise.model import AWGN

    rng = np.random.default_rng(1)
    fs, fc = 1e6, 1e9
    for lens, offs in (((50, 58), (0, 3)), ((64,), (0,)), ((20, 31, 40), (5, 0, 2))):
        sigs, offs = _case(rng, 2, lens, offs)
        mixed = SparseSignal.Empty(fs, 2, carrier_frequency=fc, delay=0.0)
        for s, o in zip(sigs, offs):  # simulated_device.py:1905-1915
            mixed = mixed.superimpose(Signal.Create(s, fs, fc, delay=o / fs))
        dense = mixed.to_dense() if mixed.num_blocks < 2 else mixed.to_dense()
        mine = ro.superimpose(sigs, offs)
        assert np.array_equal(np.asarray(dense.view(np.ndarray)), mine)
        for power in (0.25, 3.7e-3, 0.0):
            real = AWGN(seed=5).realize(power)
            noisy = np.asarray(real.add_to(dense).view(np.ndarray))
            re, im = ro.noise_normals(real.seed, mine.shape)
            assert np.array_equal(noisy, ro.add_awgn(mine, power, re, im))


def test_wrapper_refuses_host_tensors():
    import torch

    from hermespy_b200 import _lib
    from hermespy_b200.kernels import receive_combine

    with pytest.raises(_lib.HermesB200Error):
        receive_combine([torch.zeros((1, 1, 4), dtype=torch.complex128)])


@pytest.mark.gpu
@pytest.mark.parametrize("lens,offs", [((50, 58), (0, 3)), ((4097,), (0,)), ((20, 31, 40, 17), (5, 0, 2, 30))])
def test_fused_receive_kernel_is_bit_exact_in_complex128(lens, offs):
    import torch

    from hermespy_b200 import _lib
    from hermespy_b200.kernels import receive_combine

    rng = np.random.default_rng(2)
    B, nrx = 5, 3
    sigs = [(rng.standard_normal((B, nrx, T)) + 1j * rng.standard_normal((B, nrx, T))) for T in lens]
    T = max(o + n for o, n in zip(offs, lens))
    powers = rng.uniform(0.01, 2.0, B)
    re, im = rng.standard_normal((B, nrx, T)), rng.standard_normal((B, nrx, T))
    want = np.stack([ro.add_awgn(ro.superimpose([s[b] for s in sigs], offs), powers[b], re[b], im[b]) for b in range(B)])
    dev = [torch.from_numpy(s).cuda() for s in sigs]
    before = _lib.launch_counts()["misc"]
    got = receive_combine(dev, offs, re, im, powers).cpu().numpy()
    assert _lib.launch_counts()["misc"] == before + 1  # ONE pass: superposition and noise fused
    assert np.array_equal(got, want)
    plain = receive_combine(dev, offs).cpu().numpy()  # superposition only
    assert np.array_equal(plain, np.stack([ro.superimpose([s[b] for s in sigs], offs) for b in range(B)]))
    # complex64 I/O: float64 arithmetic inside, rounded once on the way out
    got32 = receive_combine([d.to(torch.complex64) for d in dev], offs, re, im, powers).cpu().numpy()
    want32 = np.stack([ro.add_awgn(ro.superimpose([s[b].astype(np.complex64).astype(np.complex128) for s in sigs], offs),
                                   powers[b], re[b], im[b]) for b in range(B)]).astype(np.complex64)
    assert np.array_equal(got32, want32)


@pytest.mark.gpu
def test_receive_after_propagate_matches_the_reference_receive_chain():
    """propagate (CUDA) -> superimpose + AWGN (CUDA) on device-resident data == the reference's numpy chain to 1e-12."""
    import torch

    from hermespy_b200.kernels import FadingBatch, fading_propagate, receive_combine
    from oracle import fading_oracle as fo
    from tests.helpers import random_fading_params, random_signal, rel_l2, stack_param_blocks

    rng = np.random.default_rng(4)
    fs, T = 30.72e6, 600
    plist = [random_fading_params(rng, 6, 8, 2, 2, fs, 200.0, 20 / fs) for _ in range(3)]
    for p in plist[1:]:
        p.delay = plist[0].delay
    xs = [random_signal(rng, 2, T) for _ in plist]
    blk = stack_param_blocks(plist)
    fb = FadingBatch.from_numpy(device="cuda", **blk)
    y = fading_propagate(torch.from_numpy(np.stack(xs)).cuda(), fb, precision="f64")
    own = torch.from_numpy(np.stack([random_signal(rng, 2, T) for _ in plist])).cuda()  # e.g. the self-interference link
    Tout = y.shape[2]
    re, im = rng.standard_normal((3, 2, Tout)), rng.standard_normal((3, 2, Tout))
    got = receive_combine([own, y], [0, 0], re, im, 0.1).cpu().numpy()
    for b, p in enumerate(plist):
        ref = ro.add_awgn(ro.superimpose([own[b].cpu().numpy(), fo.propagate(p, xs[b])], [0, 0]), 0.1, re[b], im[b])
        assert rel_l2(got[b], ref) < 1e-12
