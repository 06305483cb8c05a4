"""GPU parity: CUDA fading kernels (through the C-ABI) against the float64 numpy oracle."""
import numpy as np
import pytest

from oracle import fading_oracle as fo
from tests.helpers import random_fading_params, random_signal, rel_l2, stack_param_blocks

pytestmark = pytest.mark.gpu

F32_TOL = 1e-5   # north_star: relative L2 <= 1e-5 against the reference's float64 path
F64_TOL = 1e-12  # parity mode


def _run_case(B, L, N, ntx, nrx, T, fs, doppler, max_delay_s, precision, sos_mode, io, seed=0, los_doppler=None,
              rice=None, same_profile=True, large=False):
    import torch
    from hermespy_b200.kernels import FadingBatch, fading_propagate

    rng = np.random.default_rng(seed)
    p0 = random_fading_params(rng, L, N, ntx, nrx, fs, doppler, max_delay_s, los_doppler, rice)
    plist = [p0]
    for _ in range(B - 1):
        plist.append(random_fading_params(rng, L, N, ntx, nrx, fs, doppler, max_delay_s, los_doppler, rice,
                                          delays=p0.delay, powers=p0.power))
    if large:  # beyond the reference's 10 x 10 antenna variable: a dense random spatial response per link
        import dataclasses

        plist = [dataclasses.replace(p, spatial=(rng.standard_normal((nrx, ntx)) + 1j * rng.standard_normal((nrx, ntx))) / np.sqrt(2 * ntx))
                 for p in plist]
    xs = [random_signal(rng, ntx, T) for _ in range(B)]
    ref = np.stack([fo.propagate(p, x) for p, x in zip(plist, xs)])
    blk = stack_param_blocks(plist)
    fb = FadingBatch.from_numpy(**blk)
    x = torch.from_numpy(np.stack(xs).astype(io)).cuda()
    y, info = fading_propagate(x, fb, precision=precision, sos_mode=sos_mode, return_info=True)
    torch.cuda.synchronize()
    y = y.cpu().numpy()
    assert y.shape == ref.shape
    return rel_l2(y, ref), info


@pytest.mark.parametrize("ntx,nrx", [(1, 1), (2, 2), (4, 4), (4, 2), (3, 5), (8, 8), (10, 10)])
@pytest.mark.parametrize("sos_mode", ["poly", "poly_window", "poly_gather", "direct"])
def test_f32_small_doppler(ntx, nrx, sos_mode):
    err, info = _run_case(B=3, L=12, N=20, ntx=ntx, nrx=nrx, T=1500, fs=30.72e6, doppler=100.0,
                          max_delay_s=1.5e-6, precision="f32", sos_mode=sos_mode, io=np.complex64)
    assert info["mode"] == sos_mode.split("_")[0]
    if sos_mode != "direct":
        # 12 taps over 46 samples: the sliding window reads less shared memory than the gather kernel
        want = {"poly": "window", "poly_window": "window", "poly_gather": "gather"}[sos_mode]
        assert info["variant"] == want, info  # T + D < 2048: too short for the persistent TMA kernel
    assert err < F32_TOL, (err, info)


@pytest.mark.parametrize("io", [np.complex64, np.complex128])
@pytest.mark.parametrize("T,max_delay_s,L", [(37, 0.0, 3), (255, 2e-7, 5), (256, 2.6e-7, 9), (1023, 1e-6, 16),
                                              (1025, 3e-6, 40), (5000, 3.3e-5, 200), (9000, 1e-5, 2)])
@pytest.mark.parametrize("ntx,nrx", [(1, 1), (2, 3), (4, 4), (8, 8), (5, 9)])
def test_window_variant_shapes(ntx, nrx, T, max_delay_s, L, io):
    """Sliding-window kernel over ragged frame lengths, tile edges, dense/sparse delay sets, rx chunking."""
    err, info = _run_case(B=2, L=L, N=8, ntx=ntx, nrx=nrx, T=T, fs=30.72e6, doppler=300.0,
                          max_delay_s=max_delay_s, precision="f32", sos_mode="auto", io=io, seed=T)
    assert info["mode"] == "poly"
    if L == 2:
        assert info["variant"] == "gather", info  # two taps 307 samples apart: a window walk would be wasted
    elif not (L == 200 and ntx > 4):  # 8-antenna chunks slide R = 4 outputs: 1014 delays / 4 > 183 groups
        assert info["variant"] == "window", info
    assert err < F32_TOL, (err, info)


def test_window_and_gather_agree_on_c2_shape():
    """Both POLY kernels evaluate the same Taylor model; they differ only by FP32 summation order."""
    e1, i1 = _run_case(B=2, L=23, N=20, ntx=4, nrx=4, T=15344, fs=30.72e6, doppler=100.0, max_delay_s=1.44e-6,
                       precision="f32", sos_mode="poly", io=np.complex64)
    e2, i2 = _run_case(B=2, L=23, N=20, ntx=4, nrx=4, T=15344, fs=30.72e6, doppler=100.0, max_delay_s=1.44e-6,
                       precision="f32", sos_mode="poly_gather", io=np.complex64)
    e3, i3 = _run_case(B=2, L=23, N=20, ntx=4, nrx=4, T=15344, fs=30.72e6, doppler=100.0, max_delay_s=1.44e-6,
                       precision="f32", sos_mode="poly_window", io=np.complex64)
    assert i1["variant"] == "tma" and i2["variant"] == "gather" and i3["variant"] == "window"
    assert i1["tile"] == 1024 and i1["poly_tile"] == 2048, i1
    assert i1["poly_tile"] % i1["tile"] == 0
    assert e1 < 1e-6 and e2 < 1e-6 and e3 < 1e-6, (e1, e2, e3)


@pytest.mark.parametrize("ntx,nrx", [(1, 1), (2, 2), (3, 2), (4, 4), (8, 4), (6, 7)])
@pytest.mark.parametrize("T,B", [(1024, 3), (2048, 5), (4094, 2), (15344, 3), (20000, 1)])
def test_window_variant_tiles_and_edges(ntx, nrx, T, B):
    """Sliding-window kernel: first / last tiles, frame lengths around tile multiples, partial antenna chunks."""
    err, info = _run_case(B=B, L=14, N=12, ntx=ntx, nrx=nrx, T=T, fs=30.72e6, doppler=250.0, max_delay_s=1.4e-6,
                          precision="f32", sos_mode="auto", io=np.complex64, seed=T + ntx)
    tma_shape = ntx <= 4 and T % 16 == 0 and T + 43 >= 2048
    assert info["variant"] == ("window" if (ntx, nrx) == (1, 1) else "tma" if tma_shape else "window"), info  # 1 x 1: window
    assert info["poly_tile"] % info["tile"] == 0 and info["tile"] in (256, 512, 1024, 2048), info
    assert err < F32_TOL, (err, info)


@pytest.mark.parametrize("ntx,nrx", [(1, 1), (2, 2), (3, 2), (4, 4), (4, 9), (2, 64)])
@pytest.mark.parametrize("T,B,max_delay_s,doppler", [(2048, 5, 1.4e-6, 250.0), (4096, 3, 4.1e-6, 2e3), (15344, 40, 1.44e-6, 100.0),
                                                      (3056, 7, 0.0, 50.0), (20000, 2, 2.0e-6, 2e4), (16384, 150, 3.0e-7, 0.0)])
def test_tma_variant(ntx, nrx, T, B, max_delay_s, doppler):
    """Persistent TMA-pipelined window kernel: ring reuse over many tiles per CTA (B x tiles > 2 x 444 CTAs), halo
    sizes from 0 to 126 samples, first / last tiles, partial antenna chunks, every compiled Taylor order."""
    if B >= 40 and ntx * nrx > 16:
        pytest.skip("oracle time")
    err, info = _run_case(B=B, L=20, N=12, ntx=ntx, nrx=nrx, T=T, fs=30.72e6, doppler=doppler, max_delay_s=max_delay_s,
                          precision="f32", sos_mode="poly_tma", io=np.complex64, seed=T + ntx)
    if info["poly_tile"] % 1024 == 0:
        assert info["variant"] == "tma" and info["tile"] == 1024, info
    else:  # fast fading: Taylor windows shorter than the 1024-output tile of the persistent kernel
        assert info["variant"] == "window", info
    assert err < F32_TOL, (err, info)


@pytest.mark.parametrize("ntx,nrx,T,B", [(16, 16, 2048, 2), (64, 64, 4096, 2), (24, 40, 2048, 1), (33, 17, 3072, 1), (70, 66, 2048, 1)])
def test_large_array_tensor_core_path(ntx, nrx, T, B):
    """Config C4 shape and friends: tap delay lines per transmit antenna (z mode of the TMA kernel, chunks of 4, even-pitch
    workspace), then the spatial product on the tcgen05 tensor cores in 3xTF32 -- one C-ABI call.  ``sos_mode="poly_fused"``
    (up to 64 x 64): ONE kernel, the GEMM first and the delay lines on its accumulator (fused_gemm_tdl_kernel)."""
    kw = dict(B=B, L=12, N=20, ntx=ntx, nrx=nrx, T=T, fs=30.72e6, doppler=100.0, max_delay_s=1.5e-6, precision="f32",
              io=np.complex64, seed=ntx, rice=np.r_[3.0, np.zeros(11)], large=True)
    err, info = _run_case(sos_mode="auto", **kw)
    gemms = ((nrx + 63) // 64) * ((ntx + 63) // 64)
    assert info["variant"] == "tma" and info["launches"] == 2 + gemms, info  # K1, one z-mode launch, GEMM blocks
    assert err < F32_TOL, (err, info)
    if ntx <= 64 and nrx <= 64:  # the single-kernel variant (z never reaches HBM; opt-in, see profiles/r02_c4.md)
        err1, info1 = _run_case(sos_mode="poly_fused", **kw)
        assert info1["variant"] == "fused" and info1["launches"] == 2, info1
        assert err1 < F32_TOL, (err1, info1)
    # the same problem through the chunked kernels (no tensor cores)
    err2, info2 = _run_case(sos_mode="poly_window", **kw)
    assert info2["variant"] in ("window", "gather") and err2 < F32_TOL


@pytest.mark.parametrize("ntx,nrx,T,B,L,delay_s,doppler", [
    (64, 64, 16384, 3, 23, 3.7e-6, 100.0),   # C4 shape: D = 114, 258 tiles, three links
    (64, 64, 1000, 1, 4, 4.1e-6, 0.0),       # static channel (P = 1), delay 126 of the 128-sample history, T % 64 != 0
    (32, 48, 70000, 1, 9, 1e-6, 3e3),        # one long link: several segments with history tiles, faster Doppler
    (17, 64, 4096 + 37, 5, 6, 2e-6, 800.0),  # odd sizes, K and N padding
    (64, 16, 64, 2, 3, 0.0, 50.0),           # one tile, all taps at delay 0
])
def test_fused_gemm_delay_line_kernel(ntx, nrx, T, B, L, delay_s, doppler):
    """fused_gemm_tdl_kernel against the oracle: segment starts (history tiles), ring wrap-around, the frame tail where u
    is zero, padded antenna counts, every polynomial order the planner picks."""
    err, info = _run_case(B=B, L=L, N=12, ntx=ntx, nrx=nrx, T=T, fs=30.72e6, doppler=doppler, max_delay_s=delay_s,
                          precision="f32", sos_mode="poly_fused", io=np.complex64, seed=7 * ntx + nrx, large=True)
    assert info["variant"] == "fused" and info["launches"] == 2, info
    assert err < F32_TOL, (err, info)


def test_tma_matches_window_kernel_bitwise_model():
    """The TMA and cp.async window kernels evaluate the same Taylor model in the same summation order."""
    import torch
    from hermespy_b200.kernels import FadingBatch, fading_propagate

    rng = np.random.default_rng(5)
    p0 = random_fading_params(rng, 23, 20, 4, 4, 30.72e6, 100.0, 1.44e-6, None, None)
    blk = stack_param_blocks([p0] * 3)
    fb = FadingBatch.from_numpy(**blk)
    x = torch.from_numpy(np.stack([random_signal(rng, 4, 15344) for _ in range(3)]).astype(np.complex64)).cuda()
    y1, i1 = fading_propagate(x, fb, sos_mode="poly_tma", return_info=True)
    y2, i2 = fading_propagate(x, fb, sos_mode="poly_window", return_info=True)
    assert i1["variant"] == "tma" and i2["variant"] == "window"
    assert torch.equal(y1, y2)


@pytest.mark.parametrize("T", [2048, 4080])
def test_tma_schedule_covers_every_tile(T):
    """The persistent kernel's dynamic schedule (static first tiles, global counter, in-order slot retirement, end
    markers) must process every tile exactly once for any tile count around multiples of the SM count and ring depth,
    including frames whose last tile holds a handful of outputs (it retires almost immediately).  Bitwise against the
    one-tile-per-CTA window kernel."""
    import torch
    from hermespy_b200.kernels import FadingBatch, fading_propagate

    rng = np.random.default_rng(11)
    p0 = random_fading_params(rng, 9, 8, 2, 2, 30.72e6, 100.0, 4e-7, None, None)
    blk1 = stack_param_blocks([p0])
    for B in (1, 2, 49, 50, 147, 148, 149, 295, 296, 297, 443, 445, 889, 1200):
        blk = {k: (np.repeat(v, B, axis=0) if isinstance(v, np.ndarray) and v.ndim == 3 else v) for k, v in blk1.items()}
        fb = FadingBatch.from_numpy(**blk)
        x = torch.view_as_complex(torch.randn((B, 2, T, 2), device="cuda", dtype=torch.float32))
        y1, i1 = fading_propagate(x, fb, sos_mode="poly_tma", return_info=True)
        y2, i2 = fading_propagate(x, fb, sos_mode="poly_window", return_info=True)
        assert i1["variant"] == "tma" and i2["variant"] == "window"
        assert torch.equal(y1, y2), (B, T)


def test_tma_falls_back_for_unaligned_frames():
    err, info = _run_case(B=2, L=14, N=12, ntx=4, nrx=4, T=4100, fs=30.72e6, doppler=100.0, max_delay_s=1.4e-6,
                          precision="f32", sos_mode="poly_tma", io=np.complex64)
    assert info["variant"] == "window", info  # T % 16 != 0
    assert err < F32_TOL


@pytest.mark.parametrize("io", [np.complex64, np.complex128])
@pytest.mark.parametrize("doppler,expect", [(0.0, "poly"), (1e3, "poly"), (1e4, "poly"), (3e5, "direct"), (5e7, "direct")])
def test_f32_auto_mode(doppler, expect, io):
    err, info = _run_case(B=2, L=23, N=20, ntx=4, nrx=4, T=4096, fs=30.72e6, doppler=doppler,
                          max_delay_s=1.5e-6, precision="f32", sos_mode="auto", io=io, los_doppler=0.7 * doppler,
                          rice=np.r_[13.3, np.zeros(22)])
    assert info["mode"] == expect, info
    assert err < F32_TOL, (err, info)


@pytest.mark.parametrize("ntx,nrx", [(1, 1), (2, 2), (4, 4), (8, 3)])
def test_f64_parity(ntx, nrx):
    err, info = _run_case(B=2, L=10, N=5, ntx=ntx, nrx=nrx, T=700, fs=1e9, doppler=2.5e10, max_delay_s=1e-7,
                          precision="f64", sos_mode="auto", io=np.complex128, los_doppler=1.3e10,
                          rice=np.r_[2.0, np.zeros(9)])
    assert info["mode"] == "direct"
    # |phase| reaches ~2e4 rad here (reference test_fading.py:139-140 uses dopplers up to 50*fs); the
    # reference itself only resolves such arguments to ~1e-12
    assert err < 1e-10, (err, info)


def test_f64_parity_realistic():
    err, info = _run_case(B=2, L=23, N=20, ntx=4, nrx=4, T=2048, fs=30.72e6, doppler=100.0, max_delay_s=1.5e-6,
                          precision="f64", sos_mode="auto", io=np.complex128)
    assert err < F64_TOL, (err, info)


def test_flat_channel_all_taps_one_group():
    # TDL default rms_delay = 0: every tap lands on delay 0 (config C1)
    import torch
    from hermespy_b200.kernels import FadingBatch, fading_propagate

    rng = np.random.default_rng(5)
    L, N = 23, 20
    p = random_fading_params(rng, L, N, 1, 1, 4e8, 0.0, 0.0, delays=np.zeros(L))
    x = random_signal(rng, 1, 500)
    ref = fo.propagate(p, x)
    fb = FadingBatch.from_numpy(**stack_param_blocks([p]))
    y, info = fading_propagate(torch.from_numpy(x[None].astype(np.complex64)).cuda(), fb, return_info=True)
    assert info["num_groups"] == 1 and info["poly_order"] == 1
    assert rel_l2(y.cpu().numpy()[0], ref) < F32_TOL


def test_long_frame_many_tiles():
    err, info = _run_case(B=1, L=20, N=20, ntx=1, nrx=1, T=70000, fs=30.72e6, doppler=50.0, max_delay_s=2.14e-6,
                          precision="f32", sos_mode="auto", io=np.complex64)
    assert info["num_tiles"] > 8
    assert err < F32_TOL, (err, info)


def test_host_buffer_entry_matches_device_entry():
    import torch
    from hermespy_b200.kernels import FadingBatch, fading_propagate, fading_propagate_host

    rng = np.random.default_rng(11)
    B, L, N, ntx, nrx, T = 7, 8, 20, 2, 2, 900
    p0 = random_fading_params(rng, L, N, ntx, nrx, 30.72e6, 200.0, 1e-6)
    plist = [p0] + [random_fading_params(rng, L, N, ntx, nrx, 30.72e6, 200.0, 1e-6, delays=p0.delay, powers=p0.power)
                    for _ in range(B - 1)]
    xs = np.stack([random_signal(rng, ntx, T) for _ in range(B)])
    blk = stack_param_blocks(plist)
    ref = np.stack([fo.propagate(p, x) for p, x in zip(plist, xs)])
    for dt, prec, tol in [(np.complex128, "f64", F64_TOL), (np.complex128, "f32", F32_TOL), (np.complex64, "f32", F32_TOL)]:
        yh = fading_propagate_host(xs.astype(dt), precision=prec, chunk_links=3, **blk)
        assert yh.dtype == dt
        assert rel_l2(yh, ref) < tol
        fb = FadingBatch.from_numpy(**blk)
        yd = fading_propagate(torch.from_numpy(xs.astype(dt)).cuda(), fb, precision=prec).cpu().numpy()
        np.testing.assert_array_equal(yh, yd)  # same kernels, same bits


def test_state_matches_oracle():
    import torch
    from hermespy_b200.kernels import FadingBatch, fading_state

    rng = np.random.default_rng(3)
    p = random_fading_params(rng, 12, 20, 2, 2, 30.72e6, 1e4, 1.5e-6)
    fb = FadingBatch.from_numpy(**stack_param_blocks([p]))
    T = 333
    h, gd = fading_state(fb, T, precision="f64")
    h = h.cpu().numpy()[0]
    hl = fo.tap_impulses(p, T)
    d = fo.tap_delays_in_samples(p)
    for g, dg in enumerate(gd):
        want = hl[d == dg].sum(axis=0)
        assert rel_l2(h[g], want) < 1e-12
    h32, _ = fading_state(fb, T, precision="f32", io128=False)
    for g, dg in enumerate(gd):
        assert rel_l2(h32.cpu().numpy()[0][g], hl[d == dg].sum(axis=0)) < F32_TOL


def test_stream_count_mismatch_raises():
    import torch
    from hermespy_b200.kernels import FadingBatch, fading_propagate

    rng = np.random.default_rng(2)
    p = random_fading_params(rng, 4, 20, 2, 2, 1e6, 0.0, 1e-6)
    fb = FadingBatch.from_numpy(**stack_param_blocks([p]))
    with pytest.raises(ValueError):
        fading_propagate(torch.zeros((1, 3, 10), dtype=torch.complex64, device="cuda"), fb)


@pytest.mark.parametrize("io", [np.complex64, np.complex128])
@pytest.mark.parametrize("T,max_delay_s,L,doppler,B", [
    (37, 0.0, 3, 300.0, 2),            # shorter than one pair row, no delay
    (500, 1e-7, 23, 100.0, 9),         # config C1 shape: one delay group, one 512-output tile
    (1025, 3e-6, 40, 300.0, 3),        # odd frame length (unaligned pair stores on odd links), odd and even delays
    (4094, 2.15e-6, 20, 50.0, 2),      # COST259-like spread, last tile nearly empty
    (70000, 2.15e-6, 20, 50.0, 2),     # many 2048-output tiles, several Taylor windows
    (9000, 1e-5, 2, 2e3, 2),           # two taps 307 samples apart, fast fading (short Taylor windows)
    (3000, 3.3e-5, 200, 300.0, 1),     # 1014-sample delay spread: halo as long as half a tile
])
def test_siso_kernel(T, max_delay_s, L, doppler, B, io):
    """tdl_siso_kernel (fading_siso.cuh): planar staging, pairs of consecutive outputs in one FFMA2 -- frame edges, delay
    parities, tile sizes 512 / 1024 / 2048, both frame element types, against the float64 oracle."""
    err, info = _run_case(B=B, L=L, N=12, ntx=1, nrx=1, T=T, fs=30.72e6, doppler=doppler, max_delay_s=max_delay_s,
                          precision="f32", sos_mode="poly_siso", io=io, seed=T)
    assert info["mode"] == "poly" and info["variant"] == "siso" and info["launches"] == 2, info
    assert info["tile"] in (512, 1024, 2048) and info["poly_tile"] % info["tile"] == 0, info
    assert err < F32_TOL, (err, info)
    err_w, info_w = _run_case(B=B, L=L, N=12, ntx=1, nrx=1, T=T, fs=30.72e6, doppler=doppler, max_delay_s=max_delay_s,
                              precision="f32", sos_mode="poly_gather", io=io, seed=T)
    assert info_w["variant"] == "gather" and err_w < F32_TOL


def test_siso_mode_is_refused_for_arrays():
    from hermespy_b200 import _lib

    with pytest.raises(_lib.HermesB200Error):
        _run_case(B=1, L=4, N=8, ntx=2, nrx=1, T=600, fs=30.72e6, doppler=10.0, max_delay_s=1e-7, precision="f32",
                  sos_mode="poly_siso", io=np.complex64)


@pytest.mark.parametrize("precision,io,tol", [("f64", np.complex128, F64_TOL), ("f32", np.complex64, F32_TOL)])
def test_delay_spread_beyond_any_staged_tile(precision, io, tol):
    """A 60 000-sample delay spread (2 ms at 30.72 MHz): no kernel can stage the halo, the planner's last resort evaluates
    per sample in FP64 with x read from global memory.  The reference handles such links; so must the drop-in."""
    err, info = _run_case(B=2, L=5, N=8, ntx=3, nrx=2, T=9000, fs=30.72e6, doppler=200.0, max_delay_s=60000 / 30.72e6,
                          precision=precision, sos_mode="auto", io=io, seed=5)
    assert info["mode"] == "direct"
    assert err < tol, (err, info)


def test_thousands_of_receive_antennas_take_the_gather_kernel():
    """8 x 3500 antennas overflow the window kernel's shared memory (its [Nrx x 8] spatial block alone is 224 KB);
    AUTO re-plans with the gather kernel, which halves its antenna chunk until the block fits."""
    err, info = _run_case(B=1, L=6, N=8, ntx=8, nrx=3500, T=300, fs=30.72e6, doppler=100.0, max_delay_s=1e-6,
                          precision="f32", sos_mode="auto", io=np.complex64, seed=7, large=True)
    assert info["mode"] == "poly" and info["variant"] == "gather", info
    assert err < F32_TOL, (err, info)


def test_host_entry_runs_on_the_requested_device_from_a_worker_thread():
    """CUDA's current device is per thread and defaults to 0, so a rank that pinned its GPU in the main thread used
    to run the drop-in's calls -- made from an actor's worker thread -- on GPU 0.  The host entries take the device explicitly:
    a call from a fresh thread with ``device=1`` allocates and launches on GPU 1 and leaves GPU 0 alone."""
    import torch
    from concurrent.futures import ThreadPoolExecutor

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from hermespy_b200 import _lib
    from hermespy_b200.kernels import fading_propagate_host

    rng = np.random.default_rng(3)
    plist = [random_fading_params(rng, 6, 8, 2, 2, 30.72e6, 100.0, 1e-6)]
    blk = stack_param_blocks(plist)
    x = random_signal(rng, 2, 40000)[None]  # 1.3 MB in, 1.3 MB out: the pipeline's slot buffers are visible in mem_get_info
    ref = fo.propagate(plist[0], x[0])

    def call(device):
        y = fading_propagate_host(x, precision="f64", device=device, **blk)
        return y, torch.cuda.current_device()

    with ThreadPoolExecutor(1) as pool:
        y0, _ = pool.submit(call, 0).result()
    torch.cuda.synchronize(0)
    free0 = torch.cuda.mem_get_info(0)[0]
    free1 = torch.cuda.mem_get_info(1)[0]
    with ThreadPoolExecutor(1) as pool:  # a FRESH thread: its current device starts at 0
        y1, dev_after = pool.submit(call, 1).result()
    assert dev_after == 1
    assert np.array_equal(y0, y1) and rel_l2(y1[0], ref) < F64_TOL
    assert torch.cuda.mem_get_info(1)[0] < free1            # context / slot buffers appeared on GPU 1 ...
    assert torch.cuda.mem_get_info(0)[0] >= free0 - (1 << 20)  # ... and nothing was allocated on GPU 0 (its slots were released)
    _lib.set_device(0)
