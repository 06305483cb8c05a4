"""Fractional-delay (``InterpolationMode.SINC``) extension of the fading path -- BASELINE config C5 "with fractional delays".

The reference rounds tap delays (fading.py:297); ``InterpolationMode.SINC`` is defined by it as the
Whittaker-Shannon formula (core/definitions.py:82-93) but implemented by no channel.  This extension has its OWN oracle
(``oracle.fading_oracle.propagate_sinc``: per-tap convolution with a Kaiser-windowed sinc centred on the true delay) and
stated tolerances: the product path (tap-table expansion + the NEAREST kernels) equals the oracle to 1e-12 (f64) / 1e-5
(f32); against the IDEAL band-limited fractional delay the 12-tap filter (Kaiser beta = 6) is accurate to 1e-3 relative
L2 for signals occupying 70 % of the band (channel taps at least 6 samples into the frame: closer taps lose their
non-causal filter precursor).  NEAREST parity is untouched: integer delays expand to themselves.
"""
import numpy as np
import pytest

from oracle import fading_oracle as fo
from tests.helpers import random_fading_params, random_signal, rel_l2


def _expanded_oracle_params(p):
    """The product's expansion evaluated by the NEAREST oracle (pins hb_fading_sinc_taps on the CPU)."""
    from dataclasses import replace

    from hermespy_b200.kernels import sinc_expand

    om, ph, am = fo.sinusoid_rates(p)
    amp2 = np.stack([am[:, 0], am[:, 1] if am.shape[1] > 1 else np.zeros(len(am))], axis=1)
    return sinc_expand(p.delay, p.fs, om, ph, amp2)


def _nearest_eval(blk, x, spatial):
    T, D = x.shape[1], blk["max_delay"]
    n = np.arange(T)
    K = blk["omega"].shape[1]
    amp = blk["amp"][:, [0] + [1] * (K - 1), None]
    h = (amp * np.exp(1j * (blk["omega"][:, :, None] * n + blk["phi"][:, :, None]))).sum(1)
    z = np.zeros((x.shape[0], T + D), complex)
    for l, d in enumerate(blk["tap_delay"]):
        z[:, d: d + T] += x * h[l]
    return spatial @ z


def test_expansion_equals_the_convolution_oracle():
    rng = np.random.default_rng(0)
    fs = 30.72e6
    for trial in range(4):
        L = int(rng.integers(1, 8))
        p = random_fading_params(rng, L, 6, 2, 2, fs, 80.0, 40 / fs)
        x = random_signal(rng, 2, 300)
        blk = _expanded_oracle_params(p)
        assert np.all(np.diff(blk["tap_delay"]) >= 0) and blk["max_delay"] == fo.sinc_max_delay_in_samples(p)
        y = _nearest_eval(blk, x, p.spatial[: p.num_rx, : p.num_tx])
        ref = fo.propagate_sinc(p, x)
        assert y.shape == ref.shape and rel_l2(y, ref) < 1e-13


def test_integer_delays_expand_to_themselves():
    from hermespy_b200.kernels import sinc_expand

    fs = 1e6
    delays = np.array([0.0, 3.0, 3.0, 17.0]) / fs
    om, ph, am = np.zeros((4, 3)), np.zeros((4, 3)), np.ones((4, 2))
    blk = sinc_expand(delays, fs, om, ph, am)
    assert blk["tap_delay"].tolist() == [0, 3, 3, 17] and np.array_equal(blk["amp"], am)  # NEAREST == SINC there
    assert blk["max_delay"] == 17 + 6  # the output always carries the filter's tail


def test_against_the_ideal_bandlimited_fractional_delay():
    """A static single-tap channel IS a pure fractional delay: compare with the exact band-limited shift (FFT phase ramp)."""
    rng = np.random.default_rng(1)
    fs, T = 1e6, 2048
    X = np.zeros(T, complex)
    band = int(0.35 * T)  # 70 % of the band occupied
    X[:band] = rng.standard_normal(band) + 1j * rng.standard_normal(band)
    X[-band:] = rng.standard_normal(band) + 1j * rng.standard_normal(band)
    x = np.fft.ifft(X)[None, :] * np.hanning(T)[None, :]  # tapered: the circular shift equals the linear one
    for tau in (10.37, 6.5, 25.99):
        p = fo.FadingParams(power=np.ones(1), delay=np.array([tau / fs]), los_gain=np.ones(1), nlos_gain=np.zeros(1),
                            los_angle=np.zeros(1), nlos_angle=np.zeros((1, 1)), los_phase=np.zeros(1),
                            nlos_phase=np.zeros((1, 1)), los_doppler=0.0, nlos_doppler=0.0, spatial=np.eye(1, dtype=complex),
                            gain=1.0, fs=fs, num_rx=1, num_tx=1)
        y = fo.propagate_sinc(p, x)[0]
        f = np.fft.fftfreq(T)
        ideal = np.fft.ifft(np.fft.fft(x[0]) * np.exp(-2j * np.pi * f * tau))
        assert rel_l2(y[:T], ideal) < 1e-3
        nearest = fo.propagate(p, x)[0][:T]
        assert rel_l2(nearest, ideal) > 10 * rel_l2(y[:T], ideal)  # what rounding the delay costs


def test_capacity_is_reported():
    from hermespy_b200 import _lib
    from hermespy_b200.kernels import sinc_expand

    with pytest.raises(_lib.HermesB200Error):
        sinc_expand((np.arange(40) + 0.5) / 1e6, 1e6, np.zeros((40, 2)), np.zeros((40, 2)), np.ones((40, 2)))  # 480 taps


@pytest.mark.gpu
@pytest.mark.parametrize("T", [1024, 70000])
def test_cost259_urban_sinc_on_the_gpu(T):
    """Config C5 shape: COST259 typical urban, SISO, fs = 30.72 MHz (all 20 delays fractional), both precisions."""
    import torch

    from hermespy_b200 import _lib
    from hermespy_b200.channel.fading.profiles import COST259_PROFILES
    from hermespy_b200.kernels import FadingBatch, fading_propagate, fading_propagate_host, sinc_expand

    rng = np.random.default_rng(5)
    fs = 30.72e6
    delays = np.asarray(COST259_PROFILES[0]["delay"], float)
    powers = np.asarray(COST259_PROFILES[0]["power"], float)
    powers = powers / powers.sum()
    plist = [random_fading_params(rng, 20, 20, 1, 1, fs, 50.0, 0.0, delays=delays, powers=powers) for _ in range(3)]
    xs = [random_signal(rng, 1, T) for _ in plist]
    blocks = [_expanded_oracle_params(p) for p in plist]
    blk = dict(tap_delay=blocks[0]["tap_delay"], max_delay=blocks[0]["max_delay"], omega=np.stack([b["omega"] for b in blocks]),
               phi=np.stack([b["phi"] for b in blocks]), amp=np.stack([b["amp"] for b in blocks]),
               spatial=np.stack([p.spatial[:1, :1] for p in plist]))
    # 20 taps x 12 windowed-sinc taps, minus the precursor taps at negative delays of the taps closest to zero delay
    assert 200 <= blk["tap_delay"].shape[0] <= 240
    before = sum(_lib.launch_counts().values())
    y64 = fading_propagate_host(np.stack(xs), precision="f64", **blk)
    y32 = fading_propagate(torch.from_numpy(np.stack(xs).astype(np.complex64)).cuda(),
                           FadingBatch.from_numpy(device="cuda", **blk), precision="f32").cpu().numpy()
    assert sum(_lib.launch_counts().values()) - before >= 2
    for b, p in enumerate(plist):
        ref = fo.propagate_sinc(p, xs[b])
        assert y64[b].shape == ref.shape
        assert rel_l2(y64[b], ref) < 1e-12
        assert rel_l2(y32[b], ref) < 1e-5
