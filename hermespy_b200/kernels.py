"""Batched, device-facing Python API over the C-ABI (torch tensors carry the device memory).

This is the layer the host-side channel classes (``hermespy_b200.fading``) and the batched drop
runner call.  Nothing here computes on the CPU: numpy is used only to lay out parameter blocks.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from . import _lib
from ._lib import HB_F32, HB_F64, HB_SOS_AUTO, HB_SOS_DIRECT, HB_SOS_POLY, FadingPlanInfo, FadingProblem

_PRECISION = {"f32": HB_F32, "f64": HB_F64, HB_F32: HB_F32, HB_F64: HB_F64}
_SOS_MODE = {"auto": HB_SOS_AUTO, "poly": HB_SOS_POLY, "direct": HB_SOS_DIRECT, "poly_gather": 3, "poly_window": 4, "poly_tma": 5, "poly_fused": 6, "poly_siso": 7}
_SOS_NAME = {HB_SOS_POLY: "poly", HB_SOS_DIRECT: "direct"}


def _torch():
    import torch

    return torch


def fading_param_block(
    power: np.ndarray,
    delay: np.ndarray,
    los_gain: np.ndarray,
    nlos_gain: np.ndarray,
    los_angle: np.ndarray,
    nlos_angle: np.ndarray,
    los_phase: np.ndarray,
    nlos_phase: np.ndarray,
    los_doppler,
    nlos_doppler,
    gain,
    fs: float,
) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Lay out the kernel parameter block for B links (leading axes broadcast).

    ``h_l[n] = amp[l,0] e^{j(omega[l,0] n + phi[l,0])} + amp[l,1] sum_{k>=1} e^{j(omega[l,k] n + phi[l,k])}``
    reproduces the reference's tap impulse (hermespy/channel/fading/fading.py:326-342) with
    ``omega[l,0] = w_los cos(theta_l0)/fs`` and ``omega[l,k] = w_nlos cos((2 pi k + theta_lk)/N)/fs``.
    Angle/phase arrays may carry a leading batch axis ``[B, L(, N)]``; profile arrays are ``[L]``.
    """
    nlos_angle = np.asarray(nlos_angle, dtype=np.float64)
    nlos_phase = np.asarray(nlos_phase, dtype=np.float64)
    los_angle = np.asarray(los_angle, dtype=np.float64)
    los_phase = np.asarray(los_phase, dtype=np.float64)
    N = nlos_angle.shape[-1]
    lead = nlos_angle.shape[:-2]
    L = nlos_angle.shape[-2]
    k = 1.0 + np.arange(N)
    w_los = np.asarray(los_doppler, dtype=np.float64).reshape(lead + (1,)) if np.ndim(los_doppler) else float(los_doppler)
    w_nlos = (
        np.asarray(nlos_doppler, dtype=np.float64).reshape(lead + (1, 1)) if np.ndim(nlos_doppler) else float(nlos_doppler)
    )
    g = np.asarray(gain, dtype=np.float64).reshape(lead + (1,)) if np.ndim(gain) else float(gain)
    omega = np.empty(lead + (L, N + 1), dtype=np.float64)
    phi = np.empty(lead + (L, N + 1), dtype=np.float64)
    amp = np.empty(lead + (L, 2), dtype=np.float64)
    omega[..., 0] = w_los * np.cos(los_angle) / fs
    omega[..., 1:] = w_nlos * np.cos((2.0 * np.pi * k + nlos_angle) / N) / fs if N > 0 else 0.0
    phi[..., 0] = los_phase
    phi[..., 1:] = nlos_phase
    scale = np.sqrt(g * np.asarray(power, dtype=np.float64))
    amp[..., 0] = np.asarray(los_gain, dtype=np.float64) * scale
    amp[..., 1] = np.asarray(nlos_gain, dtype=np.float64) * scale
    return omega, phi, amp


SINC_HALF_WIDTH = 6     # 12-tap fractional-delay filters: 20 COST259 taps expand to 240 <= HB_MAX_TAPS
SINC_KAISER_BETA = 6.0


def sinc_expand(delay_seconds: np.ndarray, fs: float, omega: np.ndarray, phi: np.ndarray, amp: np.ndarray,
                half_width: int = SINC_HALF_WIDTH, beta: float = SINC_KAISER_BETA) -> dict:
    """Fractional-delay (``InterpolationMode.SINC``) form of a fading parameter block: every tap at its TRUE delay
    ``delay * fs`` becomes up to ``2 * half_width`` Kaiser-windowed-sinc taps at integer delays (``hb_fading_sinc_taps``).

    ``omega / phi [..., L, N+1]`` and ``amp [..., L, 2]`` are gathered per expanded tap, amplitudes scaled by the filter
    weight.  Returns ``dict(tap_delay, max_delay, omega, phi, amp)`` for ``FadingBatch.from_numpy`` /
    ``fading_propagate_host`` -- the propagation kernels are the NEAREST ones (equal delays merge into one group); the
    extension lives entirely in the tap table.  The reference itself rounds delays (fading.py:297): this is an extension
    with its own oracle (``oracle.fading_oracle.propagate_sinc``)."""
    lib = _lib.load()
    d = np.ascontiguousarray(np.asarray(delay_seconds, dtype=np.float64) * float(fs))
    L = d.shape[0]
    cap = 2 * int(half_width) * L
    od, ow, osrc = np.zeros(cap, np.int32), np.zeros(cap, np.float64), np.zeros(cap, np.int32)
    n = C.c_int32(0)
    _lib.check(lib.hb_fading_sinc_taps(d.ctypes.data_as(C.POINTER(C.c_double)), L, int(half_width), float(beta), cap,
                                       od.ctypes.data_as(C.POINTER(C.c_int32)), ow.ctypes.data_as(C.POINTER(C.c_double)),
                                       osrc.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(n)))
    n = int(n.value)
    if n > _lib.HB_MAX_TAPS:
        raise _lib.HermesB200Error(_lib.HB_ERR_UNSUPPORTED, f"windowed-sinc expansion yields {n} taps (limit {_lib.HB_MAX_TAPS}); "
                                   "use a smaller half_width")
    od, ow, osrc = od[:n], ow[:n], osrc[:n]
    omega, phi, amp = (np.asarray(a, dtype=np.float64) for a in (omega, phi, amp))
    # output length T + D_s with D_s = max floor(delay) + half_width, whether or not the last filter taps are non-zero
    return dict(tap_delay=od.copy(), max_delay=int(np.floor(d).max()) + int(half_width) if L else 0,
                omega=np.ascontiguousarray(omega[..., osrc, :]), phi=np.ascontiguousarray(phi[..., osrc, :]),
                amp=np.ascontiguousarray(amp[..., osrc, :] * ow[:, None]))


@dataclass
class FadingBatch:
    """Parameters of B fading links that share one delay profile (one kernel launch).

    Device tensors: ``omega/phi [B, L, N+1]`` float64, ``amp [B, L, 2]`` float64,
    ``spatial [B, Nrx, Ntx]`` complex128.  ``tap_delay`` is a host int32 vector (ascending).
    """

    tap_delay: np.ndarray
    max_delay: int
    omega: "object"
    phi: "object"
    amp: "object"
    spatial: "object"
    omega_max: float

    @property
    def batch(self) -> int:
        return int(self.omega.shape[0])

    @property
    def num_taps(self) -> int:
        return int(self.omega.shape[1])

    @property
    def num_sinusoids(self) -> int:
        return int(self.omega.shape[2]) - 1

    @property
    def num_rx(self) -> int:
        return int(self.spatial.shape[1])

    @property
    def num_tx(self) -> int:
        return int(self.spatial.shape[2])

    @classmethod
    def from_numpy(cls, tap_delay, max_delay, omega, phi, amp, spatial, omega_max=None, device="cuda", r_rx=None,
                   r_tx=None):
        """Upload a host parameter block.  ``r_rx`` / ``r_tx``: antenna-correlation factors still to be applied,
        ``spatial <- r_rx @ spatial @ r_tx`` on the device (K2 ``hb_kron_mix``, fading.py:480-489)."""
        torch = _torch()
        omega = np.ascontiguousarray(omega, dtype=np.float64)
        if omega_max is None:
            omega_max = float(np.abs(omega).max()) if omega.size else 0.0
        dev = torch.device(device)

        def up(a, dt):
            return torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev, non_blocking=False)

        s_dev = up(spatial, np.complex128)
        if r_rx is not None or r_tx is not None:
            from .montecarlo import kron_mix

            s_dev = kron_mix(s_dev, None if r_rx is None else up(r_rx, np.complex128),
                             None if r_tx is None else up(r_tx, np.complex128), out=s_dev)
        return cls(
            tap_delay=np.ascontiguousarray(tap_delay, dtype=np.int32),
            max_delay=int(max_delay),
            omega=up(omega, np.float64),
            phi=up(phi, np.float64),
            amp=up(amp, np.float64),
            spatial=s_dev,
            omega_max=float(omega_max),
        )


def _problem(
    *,
    batch,
    num_tx,
    num_rx,
    num_samples,
    max_delay,
    num_taps,
    num_sinusoids,
    precision,
    io128,
    sos_mode,
    omega_max,
    tap_delay: np.ndarray,
    omega_ptr,
    phi_ptr,
    amp_ptr,
    spatial_ptr,
) -> Tuple[FadingProblem, np.ndarray]:
    td = np.ascontiguousarray(tap_delay, dtype=np.int32)
    if td.ndim != 1 or td.shape[0] != num_taps:
        raise ValueError("tap_delay must be a vector with one entry per tap")
    p = FadingProblem()
    p.batch = int(batch)
    p.num_tx = int(num_tx)
    p.num_rx = int(num_rx)
    p.num_samples = int(num_samples)
    p.max_delay = int(max_delay)
    p.num_taps = int(num_taps)
    p.num_sinusoids = int(num_sinusoids)
    p.precision = _PRECISION[precision]
    p.io_complex128 = 1 if io128 else 0
    p.sos_mode = _SOS_MODE[sos_mode]
    p.omega_max = float(omega_max)
    p.tap_delay = td.ctypes.data_as(C.POINTER(C.c_int32))
    p.omega = omega_ptr
    p.phi = phi_ptr
    p.amp = amp_ptr
    p.spatial = spatial_ptr
    return p, td  # td must outlive the call


def _info_dict(info: FadingPlanInfo) -> dict:
    d = info.as_dict()
    d["mode"] = _SOS_NAME.get(d["mode"], d["mode"])
    d["variant"] = ("gather", "window", "tma", "fused", "siso")[d["variant"]] if d["mode"] == "poly" else None
    return d


def fading_plan(batch: FadingBatch, num_samples: int, precision="f32", sos_mode="auto", io128=False) -> dict:
    """Ask the library which kernel path / tile / polynomial order it would use (no GPU needed)."""
    lib = _lib.load()
    p, keep = _problem(
        batch=batch.batch,
        num_tx=batch.num_tx,
        num_rx=batch.num_rx,
        num_samples=num_samples,
        max_delay=batch.max_delay,
        num_taps=batch.num_taps,
        num_sinusoids=batch.num_sinusoids,
        precision=precision,
        io128=io128,
        sos_mode=sos_mode,
        omega_max=batch.omega_max,
        tap_delay=batch.tap_delay,
        omega_ptr=None,
        phi_ptr=None,
        amp_ptr=None,
        spatial_ptr=None,
    )
    info = FadingPlanInfo()
    _lib.check(lib.hb_fading_plan(C.byref(p), C.byref(info)))
    return _info_dict(info)


def fading_propagate(x, batch: FadingBatch, precision="f32", sos_mode="auto", out=None, return_info=False):
    """Propagate ``x[B, Ntx, T]`` (cuda complex64/complex128) over B fading links -> ``y[B, Nrx, T+D]``.

    Enqueued on torch's current stream; no synchronization.  GPU counterpart of
    ``MultipathFadingSample._propagate`` (hermespy/channel/fading/fading.py:371-406) for a whole batch.
    """
    torch = _torch()
    lib = _lib.load()
    if not x.is_cuda:
        raise _lib.HermesB200Error(_lib.HB_ERR_NO_DEVICE, "x must be a CUDA tensor (no CPU fallback)")
    if x.dtype not in (torch.complex64, torch.complex128):
        raise ValueError("x must be complex64 or complex128")
    if x.dim() != 3:
        raise ValueError("x must have shape [B, Ntx, T]")
    B, ntx, T = (int(s) for s in x.shape)
    if B != batch.batch:
        raise ValueError(f"batch mismatch: x has {B} links, parameters have {batch.batch}")
    if ntx != batch.num_tx:
        raise ValueError(
            f"Number of signal streams to be propagated does not match the number of transmitter antennas ({ntx} != {batch.num_tx}))"
        )
    x = x.contiguous()
    io128 = x.dtype == torch.complex128
    Tout = T + batch.max_delay
    if out is None:
        out = torch.empty((B, batch.num_rx, Tout), dtype=x.dtype, device=x.device)
    else:
        if tuple(out.shape) != (B, batch.num_rx, Tout) or out.dtype != x.dtype or not out.is_contiguous():
            raise ValueError("out has the wrong shape / dtype / layout")
    p, keep = _problem(
        batch=B,
        num_tx=ntx,
        num_rx=batch.num_rx,
        num_samples=T,
        max_delay=batch.max_delay,
        num_taps=batch.num_taps,
        num_sinusoids=batch.num_sinusoids,
        precision=precision,
        io128=io128,
        sos_mode=sos_mode,
        omega_max=batch.omega_max,
        tap_delay=batch.tap_delay,
        omega_ptr=batch.omega.data_ptr(),
        phi_ptr=batch.phi.data_ptr(),
        amp_ptr=batch.amp.data_ptr(),
        spatial_ptr=batch.spatial.data_ptr(),
    )
    info = FadingPlanInfo()
    with torch.cuda.device(x.device):
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(
            lib.hb_fading_propagate(C.byref(p), x.data_ptr(), out.data_ptr(), C.c_void_p(stream), C.byref(info))
        )
    del keep
    if return_info:
        return out, _info_dict(info)
    return out


def fading_propagate_host(
    x: np.ndarray,
    tap_delay: np.ndarray,
    max_delay: int,
    omega: np.ndarray,
    phi: np.ndarray,
    amp: np.ndarray,
    spatial: np.ndarray,
    omega_max: Optional[float] = None,
    precision="f32",
    sos_mode="auto",
    out: Optional[np.ndarray] = None,
    chunk_links: int = 0,
    return_info=False,
    device: Optional[int] = None,
    r_rx: Optional[np.ndarray] = None,
    r_tx: Optional[np.ndarray] = None,
):
    """Host-buffer entry (what a drop-in plugin calls): numpy in, numpy out, copies inside the call.
    ``r_rx`` / ``r_tx``: correlation factors still to be applied to ``spatial`` (mixed on the device by K2 first).

    ``x`` is ``[B, Ntx, T]`` complex64 or complex128 (the reference's ``SignalBlock`` dtype).  ``device``: CUDA device
    index the call runs on (``hb_set_device`` on the calling thread; None = the thread's current device).
    """
    lib = _lib.load()
    _lib.set_device(device)
    x = np.ascontiguousarray(x)
    if x.dtype not in (np.complex64, np.complex128):
        raise ValueError("x must be complex64 or complex128")
    if x.ndim != 3:
        raise ValueError("x must have shape [B, Ntx, T]")
    B, ntx, T = x.shape
    omega = np.ascontiguousarray(omega, dtype=np.float64)
    phi = np.ascontiguousarray(phi, dtype=np.float64)
    amp = np.ascontiguousarray(amp, dtype=np.float64)
    spatial = np.ascontiguousarray(spatial, dtype=np.complex128)
    if r_rx is not None or r_tx is not None:
        torch = _torch()
        from .montecarlo import kron_mix

        dev = f"cuda:{device}" if device is not None else "cuda"
        to = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a, dtype=np.complex128)).to(dev)
        spatial = kron_mix(to(spatial), to(r_rx), to(r_tx)).cpu().numpy()
    if omega.shape != phi.shape or omega.ndim != 3 or omega.shape[0] != B:
        raise ValueError("omega/phi must have shape [B, L, N+1]")
    L, K = omega.shape[1], omega.shape[2]
    if amp.shape != (B, L, 2):
        raise ValueError("amp must have shape [B, L, 2]")
    if spatial.ndim != 3 or spatial.shape[0] != B or spatial.shape[2] != ntx:
        raise ValueError(
            f"Number of signal streams to be propagated does not match the number of transmitter antennas ({ntx} != {spatial.shape[-1]}))"
        )
    nrx = spatial.shape[1]
    if omega_max is None:
        omega_max = float(np.abs(omega).max()) if omega.size else 0.0
    Tout = T + int(max_delay)
    if out is None:
        out = np.empty((B, nrx, Tout), dtype=x.dtype)
    elif out.shape != (B, nrx, Tout) or out.dtype != x.dtype or not out.flags.c_contiguous:
        raise ValueError("out has the wrong shape / dtype / layout")
    p, keep = _problem(
        batch=B,
        num_tx=ntx,
        num_rx=nrx,
        num_samples=T,
        max_delay=max_delay,
        num_taps=L,
        num_sinusoids=K - 1,
        precision=precision,
        io128=x.dtype == np.complex128,
        sos_mode=sos_mode,
        omega_max=omega_max,
        tap_delay=tap_delay,
        omega_ptr=omega.ctypes.data,
        phi_ptr=phi.ctypes.data,
        amp_ptr=amp.ctypes.data,
        spatial_ptr=spatial.ctypes.data,
    )
    info = FadingPlanInfo()
    _lib.check(
        lib.hb_fading_propagate_host(C.byref(p), x.ctypes.data, out.ctypes.data, int(chunk_links), C.byref(info))
    )
    del keep
    if return_info:
        return out, _info_dict(info)
    return out


def fading_state(batch: FadingBatch, num_samples: int, precision="f32", io128=True):
    """SISO tap gains ``h[B, G, T]`` and the G distinct integer delays (fading.py:345-358)."""
    torch = _torch()
    lib = _lib.load()
    td = np.ascontiguousarray(batch.tap_delay, dtype=np.int32)
    G = int(np.unique(td).size)
    dev = batch.omega.device
    h = torch.empty((batch.batch, G, int(num_samples)), dtype=torch.complex128 if io128 else torch.complex64, device=dev)
    p, keep = _problem(
        batch=batch.batch,
        num_tx=batch.num_tx,
        num_rx=batch.num_rx,
        num_samples=num_samples,
        max_delay=batch.max_delay,
        num_taps=batch.num_taps,
        num_sinusoids=batch.num_sinusoids,
        precision=precision,
        io128=io128,
        sos_mode="auto",
        omega_max=batch.omega_max,
        tap_delay=td,
        omega_ptr=batch.omega.data_ptr(),
        phi_ptr=batch.phi.data_ptr(),
        amp_ptr=batch.amp.data_ptr(),
        spatial_ptr=batch.spatial.data_ptr(),
    )
    gd = np.zeros(G, dtype=np.int32)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(
            lib.hb_fading_state(C.byref(p), h.data_ptr(), gd.ctypes.data_as(C.POINTER(C.c_int32)), C.c_void_p(stream))
        )
    del keep
    return h, gd


def receive_combine(signals, offsets=None, noise_re=None, noise_im=None, noise_power=None, num_samples=None, out=None):
    """Superimpose the signals impinging on a device and add white Gaussian noise in one pass (``hb_receive_combine``).

    ``signals``: device tensors ``[B, Nrx, T_k]`` of one complex dtype; ``offsets``: their whole-sample delays;
    ``noise_re`` / ``noise_im``: the standard normals ``rng.standard_normal(shape)`` drawn TWICE in the reference's order
    (rf/noise/model.py:149-151), float64 ``[B, Nrx, T]`` (host arrays are uploaded); ``noise_power``: scalar or ``[B]``.
    Counterpart of the superposition in ``SimulatedDevice.process_input`` (simulated_device.py:1899-1915) followed by
    ``AWGNRealization.add_to`` (model.py:140-160); complex128 results equal numpy's bit for bit.
    """
    torch = _torch()
    lib = _lib.load()
    signals = list(signals)
    if not signals:
        raise ValueError("at least one impinging signal is required")
    if any(not s.is_cuda for s in signals):
        raise _lib.HermesB200Error(_lib.HB_ERR_NO_DEVICE, "hb_receive_combine needs device tensors (no CPU fallback)")
    dt, dev = signals[0].dtype, signals[0].device
    if dt not in (torch.complex64, torch.complex128) or any(s.dtype != dt or s.dim() != 3 for s in signals):
        raise ValueError("signals must be [B, Nrx, T] tensors of one complex dtype")
    B, nrx = int(signals[0].shape[0]), int(signals[0].shape[1])
    if any(tuple(s.shape[:2]) != (B, nrx) for s in signals):
        raise ValueError("signals differ in batch size or stream count")
    offsets = [0] * len(signals) if offsets is None else [int(o) for o in offsets]
    if len(offsets) != len(signals) or min(offsets) < 0:
        raise ValueError("one non-negative offset per signal")
    signals = [s.contiguous() for s in signals]
    T = max(o + int(s.shape[2]) for o, s in zip(offsets, signals)) if num_samples is None else int(num_samples)
    desc = (_lib.ReceiveInput * len(signals))()
    for k, (s, o) in enumerate(zip(signals, offsets)):
        desc[k].samples, desc[k].num_samples, desc[k].offset = s.data_ptr(), int(s.shape[2]), o
    nre = nim = scale = None
    if noise_re is not None:
        up = lambda a: (a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))).to(
            dev, torch.float64).contiguous()
        nre, nim = up(noise_re), up(noise_im)
        if tuple(nre.shape) != (B, nrx, T) or tuple(nim.shape) != (B, nrx, T):
            raise ValueError(f"noise planes must be [B={B}, Nrx={nrx}, T={T}]")
        p = np.broadcast_to(np.asarray(noise_power, dtype=np.float64), (B,))
        scale = torch.from_numpy(np.ascontiguousarray((0.5 * p) ** 0.5)).to(dev)  # (0.5 * power) ** 0.5, model.py:149
    if out is None:
        out = torch.empty((B, nrx, T), dtype=dt, device=dev)
    elif tuple(out.shape) != (B, nrx, T) or out.dtype != dt or not out.is_contiguous():
        raise ValueError("out has the wrong shape / dtype / layout")
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.hb_receive_combine(desc, len(signals), nre.data_ptr() if nre is not None else None,
                                          nim.data_ptr() if nim is not None else None,
                                          scale.data_ptr() if scale is not None else None, out.data_ptr(), B, nrx, T,
                                          1 if dt == torch.complex128 else 0, C.c_void_p(st)))
    return out


# ------------------------------------------------------------------------------------------------------------
# Cluster delay line


def spatial_gemm(spatial, z, out=None):
    """``y[b] = spatial[b] @ z[b]`` on the tcgen05 tensor cores in 3xTF32 (``hb_spatial_gemm_3xtf32``): the
    ``spatial_response @ propagated`` product of fading.py:395 for large arrays.  ``spatial``: device complex128
    ``[B, Nrx, Ntx]``, ``z``: device complex64 ``[B, Ntx, T]``; returns device complex64 ``[B, Nrx, T]``."""
    torch = _torch()
    if not (spatial.is_cuda and z.is_cuda):
        raise _lib.HermesB200Error(_lib.HB_ERR_NO_DEVICE, "hb_spatial_gemm_3xtf32 needs device tensors (no CPU fallback)")
    s = spatial.to(torch.complex128).contiguous()
    zz = z.to(torch.complex64).contiguous()
    B, nrx, ntx = (int(v) for v in s.shape)
    if zz.dim() != 3 or int(zz.shape[0]) != B or int(zz.shape[1]) != ntx:
        raise ValueError(f"z must be [B={B}, Ntx={ntx}, T], got {tuple(zz.shape)}")
    T = int(zz.shape[2])
    if out is None:
        out = torch.empty((B, nrx, T), dtype=torch.complex64, device=zz.device)
    st = torch.cuda.current_stream(zz.device).cuda_stream
    with torch.cuda.device(zz.device):
        _lib.check(_lib.load().hb_spatial_gemm_3xtf32(s.data_ptr(), zz.data_ptr(), out.data_ptr(), B, nrx, ntx, T,
                                                      C.c_void_p(st)))
    return out



def ideal_elements(n: int) -> np.ndarray:
    """Element table of ``n`` unrotated ideal isotropic elements."""
    t = np.zeros((int(n), _lib.HB_ELEMENT_STRIDE))
    t[:, [0, 4, 8]] = 1.0
    return t


def _all_plain_ideal(table: np.ndarray) -> bool:
    return bool(np.all(table == ideal_elements(table.shape[0])))


@dataclass
class CdlBlock:
    """Host-side parameter block of B CDL links sharing one delay structure (numpy arrays, see hb_cdl_problem)."""

    term_delay: np.ndarray  # int32 [Rn]
    max_delay: int
    angles: np.ndarray  # f64 [B, Rn, 4] aoa, zoa, aod, zod
    jones: np.ndarray  # c128 [B, Rn, 2, 2]
    amplitude: np.ndarray  # f64 [B, Rn]
    tx_pose: np.ndarray  # f64 [B, 12]
    rx_pose: np.ndarray  # f64 [B, 12]
    rel_velocity: np.ndarray  # f64 [B, 3]
    tx_topology: np.ndarray  # f64 [Ntx, 3]
    rx_topology: np.ndarray  # f64 [Nrx, 3]
    carrier_frequency: float
    sampling_rate: float
    line_of_sight: bool = False
    los_delay: int = 0
    los_amplitude: float = 0.0
    max_speed: Optional[float] = None
    #: antenna element models, rows of ``HB_ELEMENT_STRIDE`` doubles (rotation element -> array frame, kind, parameter);
    #: None = unrotated ideal isotropic elements (include/hermes_b200.h, hb_element_mode)
    tx_elements: Optional[np.ndarray] = None
    rx_elements: Optional[np.ndarray] = None
    #: heterogeneous batch (``hb_cdl_problem.link_term_delay``): per-link delay indices int32 [B, Rn], line-of-sight delay
    #: int32 [B] and amplitude f64 [B] (0 = link without a line of sight) and the links' own ``max_delay`` int [B] (the
    #: batch runs with the largest; a link's output is the first ``T + link_max_delay[b]`` samples).  None = uniform batch.
    link_term_delay: Optional[np.ndarray] = None
    link_los_delay: Optional[np.ndarray] = None
    link_los_amplitude: Optional[np.ndarray] = None
    link_max_delay: Optional[np.ndarray] = None

    def __post_init__(self):
        self.term_delay = np.ascontiguousarray(self.term_delay, dtype=np.int32)
        if self.link_term_delay is not None:
            B = np.asarray(self.angles).shape[0]
            self.link_term_delay = np.ascontiguousarray(self.link_term_delay, dtype=np.int32).reshape(B, -1)
            if self.link_term_delay.shape[1] != self.term_delay.shape[0]:
                raise ValueError("link_term_delay must have shape [B, Rn]")
            full = lambda v, fill, dt: np.ascontiguousarray(np.full(B, fill) if v is None else v, dtype=dt).reshape(B)
            self.link_los_delay = full(self.link_los_delay, self.los_delay, np.int32)
            self.link_los_amplitude = full(self.link_los_amplitude, self.los_amplitude, np.float64)
            self.link_max_delay = full(self.link_max_delay, self.max_delay, np.int64)
        self.angles = np.ascontiguousarray(self.angles, dtype=np.float64)
        self.jones = np.ascontiguousarray(self.jones, dtype=np.complex128)
        self.amplitude = np.ascontiguousarray(self.amplitude, dtype=np.float64)
        self.tx_pose = np.ascontiguousarray(self.tx_pose, dtype=np.float64)
        self.rx_pose = np.ascontiguousarray(self.rx_pose, dtype=np.float64)
        self.rel_velocity = np.ascontiguousarray(self.rel_velocity, dtype=np.float64)
        self.tx_topology = np.ascontiguousarray(self.tx_topology, dtype=np.float64)
        self.rx_topology = np.ascontiguousarray(self.rx_topology, dtype=np.float64)
        if self.max_speed is None:
            self.max_speed = float(np.linalg.norm(self.rel_velocity, axis=-1).max()) if self.rel_velocity.size else 0.0
        if (self.tx_elements is None) != (self.rx_elements is None):
            ideal = ideal_elements  # one side given: the other is an array of unrotated ideal elements
            self.tx_elements = ideal(self.tx_topology.shape[0]) if self.tx_elements is None else self.tx_elements
            self.rx_elements = ideal(self.rx_topology.shape[0]) if self.rx_elements is None else self.rx_elements
        if self.tx_elements is not None:
            self.tx_elements = np.ascontiguousarray(self.tx_elements, dtype=np.float64).reshape(-1, _lib.HB_ELEMENT_STRIDE)
            self.rx_elements = np.ascontiguousarray(self.rx_elements, dtype=np.float64).reshape(-1, _lib.HB_ELEMENT_STRIDE)
            if self.tx_elements.shape[0] != self.num_tx or self.rx_elements.shape[0] != self.num_rx:
                raise ValueError("element tables need one row per antenna")
            if _all_plain_ideal(self.tx_elements) and _all_plain_ideal(self.rx_elements):
                self.tx_elements = self.rx_elements = None

    @property
    def element_mode(self) -> int:
        """``hb_element_mode``: ideal / uniform (all rows of each table equal: rank-one ray matrices) / per element."""
        if self.tx_elements is None:
            return _lib.HB_ELEMENTS_IDEAL
        uniform = all(np.all(t == t[:1]) for t in (self.tx_elements, self.rx_elements))
        return _lib.HB_ELEMENTS_UNIFORM if uniform else _lib.HB_ELEMENTS_PER_ELEMENT

    @property
    def batch(self) -> int:
        return int(self.angles.shape[0])

    @property
    def num_tx(self) -> int:
        return int(self.tx_topology.shape[0])

    @property
    def num_rx(self) -> int:
        return int(self.rx_topology.shape[0])

    ARRAYS = ("angles", "jones", "amplitude", "tx_pose", "rx_pose", "rel_velocity", "tx_topology", "rx_topology")
    ELEMENT_ARRAYS = ("tx_elements", "rx_elements")

    def pointers(self, getter) -> dict:
        """Field -> address through ``getter(array)`` (host: ``a.ctypes.data``, device: tensor ``data_ptr``)."""
        d = {k: getter(getattr(self, k)) for k in self.ARRAYS}
        for k in self.ELEMENT_ARRAYS:
            d[k] = getter(getattr(self, k)) if self.tx_elements is not None else None
        return d

    def per_link(self):
        """``(term_delay [B, Rn], los_delay [B], los_amplitude [B], max_delay [B])`` of this block, uniform or not."""
        if self.link_term_delay is not None:
            return self.link_term_delay, self.link_los_delay, self.link_los_amplitude, self.link_max_delay
        B = self.batch
        return (np.broadcast_to(self.term_delay, (B, self.term_delay.shape[0])), np.full(B, self.los_delay, dtype=np.int32),
                np.full(B, self.los_amplitude if self.line_of_sight else 0.0), np.full(B, self.max_delay, dtype=np.int64))

    @classmethod
    def stack(cls, blocks) -> "CdlBlock":
        """Blocks of one array geometry along the batch axis.  Blocks that differ in their delay structure (cluster delays,
        ray count, line-of-sight state, delay spread: stochastic scenario realizations) give a heterogeneous batch: rays
        padded with zero-amplitude copies of each link's first ray, per-link delay tables, the largest ``max_delay``."""
        b0 = blocks[0]
        if any(b.geometry_key() != b0.geometry_key() for b in blocks[1:]):
            raise ValueError("blocks of one batch must share array topologies, element models, carrier frequency and sampling rate")
        cat = lambda f: np.concatenate([getattr(b, f) for b in blocks])
        common = dict(tx_pose=cat("tx_pose"), rx_pose=cat("rx_pose"), rel_velocity=cat("rel_velocity"),
                      tx_topology=b0.tx_topology, rx_topology=b0.rx_topology, carrier_frequency=b0.carrier_frequency,
                      sampling_rate=b0.sampling_rate, max_speed=max(b.max_speed for b in blocks),
                      tx_elements=b0.tx_elements, rx_elements=b0.rx_elements)
        if all(b.link_term_delay is None and b.delay_key() == b0.delay_key() for b in blocks):
            return cls(term_delay=b0.term_delay, max_delay=b0.max_delay, angles=cat("angles"), jones=cat("jones"),
                       amplitude=cat("amplitude"), line_of_sight=b0.line_of_sight, los_delay=b0.los_delay,
                       los_amplitude=b0.los_amplitude, **common)
        rn = max(1, max(b.term_delay.shape[0] for b in blocks))

        def pad(a, fill_first: bool):  # [B, Rn_b, ...] -> [B, rn, ...]: zero amplitudes, finite angles / polarizations
            if a.shape[1] == rn:
                return a
            out = np.zeros((a.shape[0], rn) + a.shape[2:], dtype=a.dtype)
            out[:, : a.shape[1]] = a
            if fill_first:
                out[:, a.shape[1]:] = a[:, :1] if a.shape[1] else 1.0
            return out

        parts = [b.per_link() for b in blocks]
        return cls(term_delay=np.zeros(rn, dtype=np.int32), max_delay=int(max(p[3].max() for p in parts)),
                   angles=np.concatenate([pad(b.angles, True) for b in blocks]),
                   jones=np.concatenate([pad(b.jones, True) for b in blocks]),
                   amplitude=np.concatenate([pad(b.amplitude, False) for b in blocks]),
                   line_of_sight=any(b.line_of_sight for b in blocks), los_delay=0, los_amplitude=0.0,
                   link_term_delay=np.concatenate([pad(np.ascontiguousarray(p[0]), False) for p in parts]),
                   link_los_delay=np.concatenate([p[1] for p in parts]),
                   link_los_amplitude=np.concatenate([p[2] for p in parts]),
                   link_max_delay=np.concatenate([p[3] for p in parts]), **common)

    def delay_key(self):
        """What a launch-uniform delay table is built from."""
        return (self.term_delay.tobytes(), self.max_delay, self.line_of_sight, self.los_delay, self.los_amplitude)

    def geometry_key(self):
        """What blocks must share to travel in one (possibly heterogeneous) batch: arrays, elements, carrier, rate."""
        el = b"" if self.tx_elements is None else self.tx_elements.tobytes() + self.rx_elements.tobytes()
        return (self.tx_topology.tobytes(), self.rx_topology.tobytes(), self.carrier_frequency, self.sampling_rate, el)

    def group_key(self):
        return self.delay_key() + self.geometry_key()


_CDL_VARIANT = {"auto": 0, "gather": 1, "umma": 2, "umma_bf16": 3}


def _cdl_info_dict(info: FadingPlanInfo) -> dict:
    d = info.as_dict()
    d["mode"] = _SOS_NAME.get(d["mode"], d["mode"])
    d["variant"] = {1: "gather", 2: "umma", 3: "umma_bf16"}.get(d["variant"]) if d["mode"] == "poly" else None
    return d


def _cdl_problem(blk: CdlBlock, num_samples: int, precision, io128: bool, ptrs: dict, variant="auto"):
    from ._lib import CdlProblem

    p = CdlProblem()
    p.variant = _CDL_VARIANT[variant]
    p.batch = blk.batch
    p.num_tx = blk.num_tx
    p.num_rx = blk.num_rx
    p.num_samples = int(num_samples)
    p.max_delay = int(blk.max_delay)
    p.num_terms = int(blk.term_delay.shape[0])
    p.line_of_sight = 1 if blk.line_of_sight else 0
    p.los_delay = int(blk.los_delay)
    p.precision = _PRECISION[precision]
    p.io_complex128 = 1 if io128 else 0
    p.carrier_frequency = float(blk.carrier_frequency)
    p.sampling_rate = float(blk.sampling_rate)
    p.los_amplitude = float(blk.los_amplitude)
    p.max_speed = float(blk.max_speed)
    p.term_delay = blk.term_delay.ctypes.data_as(C.POINTER(C.c_int32))
    if blk.link_term_delay is not None:  # host arrays in every entry (include/hermes_b200.h)
        p.link_term_delay = blk.link_term_delay.ctypes.data_as(C.POINTER(C.c_int32))
        p.link_los_delay = blk.link_los_delay.ctypes.data_as(C.POINTER(C.c_int32))
        p.link_los_amplitude = blk.link_los_amplitude.ctypes.data_as(C.POINTER(C.c_double))
    p.angles = ptrs["angles"]
    p.jones = ptrs["jones"]
    p.amplitude = ptrs["amplitude"]
    p.tx_pose = ptrs["tx_pose"]
    p.rx_pose = ptrs["rx_pose"]
    p.rel_velocity = ptrs["rel_velocity"]
    p.tx_topology = ptrs["tx_topology"]
    p.rx_topology = ptrs["rx_topology"]
    p.element_mode = blk.element_mode
    p.tx_elements = ptrs.get("tx_elements")
    p.rx_elements = ptrs.get("rx_elements")
    return p


def cdl_plan(blk: CdlBlock, num_samples: int, precision="f32", variant="auto") -> dict:
    """What ``hb_cdl_plan`` decides.  ``variant``: "auto" | "gather" (FP32-pipe K6) | "umma" (tensor-core K6)."""
    lib = _lib.load()
    p = _cdl_problem(blk, num_samples, precision, False, {k: None for k in CdlBlock.ARRAYS}, variant)  # planning reads no arrays
    info = FadingPlanInfo()
    _lib.check(lib.hb_cdl_plan(C.byref(p), C.byref(info)))
    return _cdl_info_dict(info)


def cdl_propagate_host(x: np.ndarray, blk: CdlBlock, precision="f32", out: Optional[np.ndarray] = None,
                       chunk_links: int = 0, return_info=False, device: Optional[int] = None, variant="auto"):
    """Host-buffer CDL propagation: ``x[B, Ntx, T]`` numpy complex64/128 -> ``y[B, Nrx, T + D]``.

    GPU counterpart of ``ClusterDelayLineSample._propagate`` (cluster_delay_lines.py:526-558) for a batch.
    """
    lib = _lib.load()
    _lib.set_device(device)
    x = np.ascontiguousarray(x)
    if x.dtype not in (np.complex64, np.complex128):
        raise ValueError("x must be complex64 or complex128")
    if x.ndim != 3 or x.shape[0] != blk.batch:
        raise ValueError("x must have shape [B, Ntx, T] with B matching the parameter block")
    if x.shape[1] != blk.num_tx:
        raise ValueError(
            f"Number of signal streams to be propagated does not match the number of transmitter antennas ({x.shape[1]} != {blk.num_tx}))")
    T = x.shape[2]
    Tout = T + blk.max_delay
    if out is None:
        out = np.empty((blk.batch, blk.num_rx, Tout), dtype=x.dtype)
    elif out.shape != (blk.batch, blk.num_rx, Tout) or out.dtype != x.dtype or not out.flags.c_contiguous:
        raise ValueError("out has the wrong shape / dtype / layout")
    p = _cdl_problem(blk, T, precision, x.dtype == np.complex128, blk.pointers(lambda a: a.ctypes.data), variant)
    info = FadingPlanInfo()
    _lib.check(lib.hb_cdl_propagate_host(C.byref(p), x.ctypes.data, out.ctypes.data, int(chunk_links), C.byref(info)))
    if return_info:
        return out, _cdl_info_dict(info)
    return out


class CdlDeviceBlock(object):
    """Device-resident copy of a :class:`CdlBlock` (torch tensors) for the device-pointer entry."""

    def __init__(self, blk: CdlBlock, device="cuda") -> None:
        torch = _torch()
        self.host = blk
        names = CdlBlock.ARRAYS + (CdlBlock.ELEMENT_ARRAYS if blk.tx_elements is not None else ())
        self.tensors = {k: torch.from_numpy(getattr(blk, k)).to(device) for k in names}
        self.device = self.tensors["angles"].device


def cdl_propagate(x, dblk: CdlDeviceBlock, precision="f32", out=None, return_info=False, variant="auto"):
    """Device-resident CDL propagation on torch's current stream (no synchronization)."""
    torch = _torch()
    lib = _lib.load()
    if not x.is_cuda:
        raise _lib.HermesB200Error(_lib.HB_ERR_NO_DEVICE, "x must be a CUDA tensor (no CPU fallback)")
    blk = dblk.host
    if x.dtype not in (torch.complex64, torch.complex128):
        raise ValueError("x must be complex64 or complex128")
    if x.dim() != 3 or x.shape[0] != blk.batch or x.shape[1] != blk.num_tx:
        raise ValueError("x must have shape [B, Ntx, T] matching the parameter block")
    x = x.contiguous()
    T = int(x.shape[2])
    Tout = T + blk.max_delay
    if out is None:
        out = torch.empty((blk.batch, blk.num_rx, Tout), dtype=x.dtype, device=x.device)
    elif tuple(out.shape) != (blk.batch, blk.num_rx, Tout) or out.dtype != x.dtype or not out.is_contiguous() \
            or out.device != x.device:
        raise ValueError("out has the wrong shape / dtype / layout / device")
    p = _cdl_problem(blk, T, precision, x.dtype == torch.complex128, {k: v.data_ptr() for k, v in dblk.tensors.items()},
                     variant)
    info = FadingPlanInfo()
    with torch.cuda.device(x.device):
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(lib.hb_cdl_propagate(C.byref(p), x.data_ptr(), out.data_ptr(), C.c_void_p(stream), C.byref(info)))
    if return_info:
        return out, _cdl_info_dict(info)
    return out


def cdl_state(dblk: CdlDeviceBlock, num_samples: int, io128=True):
    """Per-delay-group MIMO impulse responses ``h[B, G, Nrx, Ntx, T]`` and the G delay indices."""
    torch = _torch()
    lib = _lib.load()
    blk = dblk.host
    p = _cdl_problem(blk, num_samples, "f64", io128, {k: v.data_ptr() for k, v in dblk.tensors.items()})
    G = C.c_int32(0)
    gd = np.zeros(_lib.HB_CDL_MAX_GROUPS, dtype=np.int32)
    _lib.check(lib.hb_cdl_state(C.byref(p), None, gd.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(G), None))
    G = int(G.value)
    h = torch.empty((blk.batch, G, blk.num_rx, blk.num_tx, int(num_samples)),
                    dtype=torch.complex128 if io128 else torch.complex64, device=dblk.device)
    with torch.cuda.device(dblk.device):
        stream = torch.cuda.current_stream(dblk.device).cuda_stream
        _lib.check(lib.hb_cdl_state(C.byref(p), h.data_ptr(), gd.ctypes.data_as(C.POINTER(C.c_int32)), None,
                                    C.c_void_p(stream)))
    return h, gd[:G].copy()
