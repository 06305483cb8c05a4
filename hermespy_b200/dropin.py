"""Drop-in adapter: route an installed, UNMODIFIED HermesPy's channel hot path through the B200 kernels.

    import hermespy_b200.dropin as dropin
    dropin.enable()            # raises unless libhermes_b200.so is built and a CUDA device is visible
    ... run any Simulation script ...
    dropin.disable()

``enable()`` rebinds four methods of the reference -- nothing else is touched:

* ``hermespy.channel.fading.fading.MultipathFadingSample._propagate`` (fading.py:371-406) and ``.state`` (:345-369)
* ``hermespy.channel.cdl.cluster_delay_lines.ClusterDelayLineSample._propagate`` (cluster_delay_lines.py:526-558)
  and ``.state`` (:561-592)

The replacements read the reference sample through its public properties (fading.py:217-287,
cluster_delay_lines.py:321-403), lay out the same kernel parameter blocks the mirror classes of this package
use, and call the host-buffer C-ABI (``hb_fading_propagate_host`` / ``hb_cdl_propagate_host``).  Realization,
sampling, hooks, Signal bookkeeping, serialization and the whole drop loop remain the reference's own code.
The functions are module-level (picklable), so Ray actors that import this module inherit the patch.

There is NO fallback into the reference's numpy code: a sample the kernels cannot serve (a user-defined antenna
element class without a device pattern, a delay spread beyond the planner's limits) raises ``HermesB200Error``.
``enable(allow_reference_fallback=True)`` is the only way to let such samples run the saved reference methods; every use
is counted in ``fallbacks`` and warned about once.
"""
from __future__ import annotations

from math import ceil

import numpy as np

from . import config

_ORIGINALS = {}
#: opt-in only (``enable(allow_reference_fallback=True)``): samples served by the saved reference methods, per kind
fallbacks = {"fading_propagate": 0, "fading_state": 0, "cdl_propagate": 0, "cdl_state": 0}
_allow_fallback = False
_warned = set()


class UnsupportedByKernels(Exception):
    """A reference sample the CUDA kernels have no model for (carries the reason)."""


def _unsupported(kind: str, why: str, self, *args):
    """No silent CPU path: raise, unless the user opted into the reference fallback (counted + warned once)."""
    from . import _lib

    if not _allow_fallback:
        raise _lib.HermesB200Error(_lib.HB_ERR_UNSUPPORTED, f"{kind}: {why} (hermespy_b200 has no CPU fallback; "
                                   "dropin.enable(allow_reference_fallback=True) opts into the reference code)")
    fallbacks[kind] += 1
    if kind not in _warned:
        import warnings

        _warned.add(kind)
        warnings.warn(f"hermespy_b200.dropin: {kind} served by the reference's CPU code ({why})", RuntimeWarning)
    return _ORIGINALS[kind](self, *args)


# ---- parameter extraction from reference objects (pure host code, testable without a GPU) --------------------

def fading_block_from_reference(sample, sinc: bool = False) -> dict:
    """Kernel parameter block of a reference ``MultipathFadingSample`` (same layout as the mirror class).
    ``sinc``: fractional-delay extension -- taps at their true delays, expanded to windowed-sinc filters."""
    from .kernels import fading_param_block, sinc_expand

    fs = sample.bandwidth
    omega, phi, amp = fading_param_block(sample.power_profile, sample.delay_profile, sample.los_gains,
                                         sample.nlos_gains, sample.los_angles, sample.nlos_angles, sample.los_phases,
                                         sample.nlos_phases, sample.los_doppler, sample.nlos_doppler, sample.gain, fs)
    spatial = np.ascontiguousarray(
        np.asarray(sample.spatial_response)[: sample.num_receive_antennas, : sample.num_transmit_antennas],
        dtype=np.complex128)
    if sinc:
        blk = sinc_expand(np.asarray(sample.delay_profile, dtype=np.float64), fs, omega, phi, amp)
        blk.update(spatial=spatial, omega_max=float(max(abs(sample.los_doppler), abs(sample.nlos_doppler)) / fs))
        return blk
    tap_delay = np.rint(np.asarray(sample.delay_profile) * fs).astype(np.int32)
    if np.any(np.diff(tap_delay) < 0):
        # a sample built by hand (the reference's own unit tests do) may list its taps in any order; the channel classes
        # sort theirs (fading.py:707-711).  The kernels want ascending delays: a tap sum does not care about the order.
        order = np.argsort(tap_delay, kind="stable")
        tap_delay, omega, phi, amp = tap_delay[order], omega[order], phi[order], amp[order]
    return dict(tap_delay=tap_delay, max_delay=int(round(float(np.max(sample.delay_profile)) * fs)), omega=omega, phi=phi,
                amp=amp, spatial=spatial, omega_max=float(max(abs(sample.los_doppler), abs(sample.nlos_doppler)) / fs))


def element_table(antennas) -> np.ndarray:
    """Rows ``[rotation element -> array frame (9), hb_element_kind, parameter, 0]`` of reference ``Antenna`` objects
    (core/antennas.py:392-622).  Subclasses of the four models that override ``local_characteristics`` -- and any other
    user-defined element -- have no device pattern."""
    from hermespy.core.antennas import Dipole, IdealAntenna, LinearAntenna, PatchAntenna  # type: ignore

    from . import _lib

    kinds = ((LinearAntenna, _lib.HB_ELEMENT_LINEAR), (PatchAntenna, _lib.HB_ELEMENT_PATCH),
             (Dipole, _lib.HB_ELEMENT_DIPOLE), (IdealAntenna, _lib.HB_ELEMENT_IDEAL))
    rows = np.zeros((len(antennas), _lib.HB_ELEMENT_STRIDE))
    for m, a in enumerate(antennas):
        for base, kind in kinds:
            if isinstance(a, base) and type(a).local_characteristics is base.local_characteristics:
                rows[m, 9] = kind
                rows[m, 10] = a.slant if kind == _lib.HB_ELEMENT_LINEAR else 0.0
                break
        else:
            raise UnsupportedByKernels(f"antenna element {type(a).__name__} has no CUDA pattern model "
                                       "(supported: IdealAntenna, LinearAntenna, PatchAntenna, Dipole)")
        rows[m, :9] = np.asarray(a.pose, dtype=np.float64)[:3, :3].ravel()
    return rows


def cdl_block_from_reference(sample):
    """``kernels.CdlBlock`` (B = 1) of a reference ``ClusterDelayLineSample``.

    Raises ``UnsupportedByKernels`` for antenna element classes without a device pattern (``element_table``).
    """
    from hermespy.core import AntennaMode  # type: ignore

    from .channel.cdl.cdl import SUBCLUSTER_INDICES
    from .kernels import CdlBlock

    tx_a, rx_a = sample.transmitter_antennas, sample.receiver_antennas
    tx_el, rx_el = element_table(list(tx_a.transmit_antennas)), element_table(list(rx_a.receive_antennas))
    fs = sample.bandwidth
    C_, R_ = sample.num_clusters, sample.num_rays
    nsplit = min(2, C_)
    nvirtual = 3 * nsplit + max(0, C_ - 2)
    sub = (np.repeat(np.asarray(sample.cluster_delays)[:nsplit, None], 3, axis=1)
           + sample.cluster_delay_spread * np.array([0.0, 1.28, 2.56]))
    vdelays = np.concatenate((sub.flatten(), np.asarray(sample.cluster_delays)[nsplit:]))
    cs, rs, ds = [], [], []
    for v in range(nvirtual):
        c = int(v / 3) if v < 6 else v - 4
        for r in (SUBCLUSTER_INDICES[c] if c < nsplit else range(R_)):
            cs.append(c)
            rs.append(r)
            ds.append(vdelays[v])
    c, r = np.array(cs), np.array(rs)
    rice_lin = 10.0 ** (sample.rice_factor / 10.0)
    nlos_scale = (1.0 + rice_lin) ** -0.5 if sample.line_of_sight else 1.0

    def pose12(antennas_state):
        m = np.asarray(antennas_state.forwards_transformation, dtype=np.float64)
        return np.concatenate([m[:3, :3].ravel(), m[:3, 3]])[None]

    return CdlBlock(
        term_delay=np.array([int((d + sample.delay_offset) * fs) for d in ds], dtype=np.int32),
        max_delay=ceil(sample.max_delay * fs),
        angles=np.stack([sample.azimuth_of_arrival[c, r], sample.zenith_of_arrival[c, r],
                         sample.azimuth_of_departure[c, r], sample.zenith_of_departure[c, r]], axis=1)[None],
        jones=np.ascontiguousarray(np.asarray(sample.polarization_transformations)[:, :, c, r].transpose(2, 0, 1))[None],
        amplitude=(np.sqrt(np.asarray(sample.cluster_powers)[c] / R_) * nlos_scale)[None],
        tx_pose=pose12(tx_a), rx_pose=pose12(rx_a),
        rel_velocity=(np.asarray(sample.receiver_velocity, float) - np.asarray(sample.transmitter_velocity, float))[None],
        tx_topology=np.asarray(tx_a._topology(AntennaMode.TX), dtype=np.float64),
        rx_topology=np.asarray(rx_a._topology(AntennaMode.RX), dtype=np.float64),
        carrier_frequency=sample.carrier_frequency, sampling_rate=fs, line_of_sight=bool(sample.line_of_sight),
        los_delay=int((sample.cluster_delays[0] + sample.delay_offset) * fs),
        los_amplitude=float((rice_lin / (1 + rice_lin)) ** 0.5), tx_elements=tx_el, rx_elements=rx_el)


# ---- device work of the four replacements: plain arrays in, plain arrays out --------------------------------------------
# Kept apart from the methods because a lane worker of the batched drop runner (a forked helper process, which must not
# touch CUDA) forwards exactly these calls to the process that owns the GPU (runner.device_call).

def _dev_fading_propagate(b: dict, x: np.ndarray) -> np.ndarray:
    from .kernels import fading_propagate_host

    return fading_propagate_host(x[None], b["tap_delay"], b["max_delay"], b["omega"][None], b["phi"][None], b["amp"][None],
                                 b["spatial"][None], omega_max=b["omega_max"], precision=config.precision,
                                 sos_mode=config.sos_mode, device=config.device)[0]


def _to_host(t, out_alloc=None) -> np.ndarray:
    """Device tensor -> numpy; ``out_alloc(shape, dtype)`` supplies the destination (the runner hands out slices of the
    calling helper's shared-memory window: the device-to-host copy lands where the helper reads it, nothing is pickled)."""
    if out_alloc is None:
        return t.cpu().numpy()
    import torch

    dst = out_alloc(tuple(t.shape), np.complex128 if t.dtype == torch.complex128 else np.complex64)
    if dst is None:
        return t.cpu().numpy()
    torch.from_numpy(dst).copy_(t)
    return dst


def _dev_fading_state(b: dict, keep: np.ndarray, num_samples: int, out_alloc=None):
    from . import _lib
    from .kernels import FadingBatch, fading_state

    _lib.set_device(config.device)
    fb = FadingBatch.from_numpy(b["tap_delay"][keep], b["max_delay"], b["omega"][None][:, keep], b["phi"][None][:, keep],
                                b["amp"][None][:, keep], b["spatial"][None], omega_max=b["omega_max"],
                                device=f"cuda:{config.device}")
    h, group_delay = fading_state(fb, int(num_samples), precision=config.precision, io128=True)
    return _to_host(h[0], out_alloc), group_delay


def _dev_cdl_propagate(blk, x: np.ndarray) -> np.ndarray:
    from .kernels import cdl_propagate_host

    return cdl_propagate_host(x[None], blk, precision=config.precision, device=config.device)[0]


def _dev_cdl_state(blk, num_samples: int, out_alloc=None):
    from . import _lib
    from .kernels import CdlDeviceBlock, cdl_state

    _lib.set_device(config.device)
    h, gd = cdl_state(CdlDeviceBlock(blk, device=f"cuda:{config.device}"), int(num_samples))
    return _to_host(h[0], out_alloc), gd  # [G, Nrx, Ntx, T]


DEVICE_CALLS = {"fading_propagate": _dev_fading_propagate, "fading_state": _dev_fading_state,
                "cdl_propagate": _dev_cdl_propagate, "cdl_state": _dev_cdl_state}
#: calls whose large result can be written straight into a caller-supplied buffer (keyword ``out_alloc``)
DEVICE_CALLS_WITH_OUT = frozenset(("fading_state", "cdl_state"))


def _device(name: str, *args):
    from . import runner

    return runner.device_call(name, *args)


# ---- replacement methods (module level => picklable) ------------------------------------------------------------

def _fading_propagate(self, signal, interpolation):
    from hermespy.core import InterpolationMode  # type: ignore
    from hermespy.core.signal_model import SignalBlock  # type: ignore

    # the reference ignores `interpolation` here (fading.py:371-406 rounds every delay): so does this path, unless the
    # fractional-delay extension was switched on explicitly
    b = fading_block_from_reference(self, sinc=config.sinc_extension and interpolation == InterpolationMode.SINC)
    T = signal.num_samples
    nrx = b["spatial"].shape[0]
    if T + b["max_delay"] <= 0 or nrx == 0:
        out = np.zeros((nrx, T + b["max_delay"]), dtype=np.complex128)
    else:
        out = _device("fading_propagate", b, np.ascontiguousarray(np.asarray(signal, dtype=np.complex128)))
    return SignalBlock(out.shape[0], out.shape[1], signal.offset, out.tobytes())


def _fading_state(self, num_samples, max_num_taps, interpolation_mode=None):
    """``MultipathFadingSample.state`` (fading.py:345-369): the SISO tap gains -- the sum-of-sinusoids cost of the call --
    come from ``hb_fading_state``; the outer product with the spatial response and the sparse container stay the
    reference's own expressions, so consumers (``OFDMIdealChannelEstimation`` ...) see the type they expect."""
    from hermespy.core import ChannelStateFormat, ChannelStateInformation  # type: ignore
    from sparse import GCXS  # type: ignore

    b = fading_block_from_reference(self)
    num_taps = min(1 + b["max_delay"], max_num_taps)
    siso_csi = np.zeros((num_samples, num_taps), dtype=np.complex128)
    keep = b["tap_delay"] <= num_taps  # the reference skips taps with d_l > num_taps (fading.py:355) ...
    if np.any(b["tap_delay"][keep] >= num_taps):  # ... and indexes out of bounds for d_l == num_taps (:358)
        raise IndexError(f"index {num_taps} is out of bounds for axis 1 with size {num_taps}")
    if num_samples >= 1 and num_taps >= 1 and np.any(keep):
        h, group_delay = _device("fading_state", b, keep, int(num_samples))
        siso_csi[:, group_delay] = h.T
    mimo_csi = GCXS.from_numpy(np.einsum("ij,kl->ijkl", self.spatial_response, siso_csi), compressed_axes=(0, 1, 2))
    return ChannelStateInformation(ChannelStateFormat.IMPULSE_RESPONSE, mimo_csi, num_delay_taps=num_taps)


def _cdl_propagate(self, signal, interpolation):
    from hermespy.core import InterpolationMode  # type: ignore
    from hermespy.core.signal_model import SignalBlock  # type: ignore

    try:
        blk = cdl_block_from_reference(self)
    except UnsupportedByKernels as e:
        return _unsupported("cdl_propagate", str(e), self, signal, interpolation)
    if interpolation != InterpolationMode.NEAREST:
        out = np.zeros((self.num_receive_antennas, signal.num_samples + blk.max_delay), dtype=np.complex128)
    else:
        out = _device("cdl_propagate", blk, np.ascontiguousarray(np.asarray(signal, dtype=np.complex128)))
    return SignalBlock(out.shape[0], out.shape[1], signal._offset, out.tobytes())


def _cdl_state(self, num_samples, max_num_taps, interpolation_mode=None):
    """``ClusterDelayLineSample.state`` (cluster_delay_lines.py:561-592): per-delay MIMO impulse responses from
    ``hb_cdl_state`` (FP64 ray synthesis), scattered into the reference's dense [Nrx, Ntx, T, 1 + D] container."""
    from hermespy.core import ChannelStateFormat, ChannelStateInformation  # type: ignore

    try:
        blk = cdl_block_from_reference(self)
    except UnsupportedByKernels as e:
        return _unsupported("cdl_state", str(e), self, num_samples, max_num_taps)
    D = min(max_num_taps, blk.max_delay)
    raw_state = np.zeros((blk.num_rx, blk.num_tx, num_samples, 1 + D), dtype=np.complex128)
    if num_samples >= 1:
        h, gd = _device("cdl_state", blk, int(num_samples))
        for g, d in enumerate(gd):
            if d < max_num_taps:  # cluster_delay_lines.py:583-584
                raw_state[:, :, :, d] = h[g]
    return ChannelStateInformation(ChannelStateFormat.IMPULSE_RESPONSE, raw_state)


def enabled() -> bool:
    """True while the reference classes are patched."""
    return bool(_ORIGINALS)


def enable(precision: str = "f32", device: int | None = None, allow_reference_fallback: bool = False,
           batch_drops: int = 0, workers: int = 0, sinc_extension: bool = False) -> None:
    """Patch the reference classes.  Fails loudly when the library or a CUDA device is missing.

    ``device``: CUDA device index every patched call runs on, from whatever thread it is made (None keeps
    ``config.device``; one process per GPU passes its local rank).  ``allow_reference_fallback``: see module docstring.
    ``batch_drops`` > 0 additionally routes ``Simulation.run()`` through the batched drop runner
    (``hermespy_b200.runner``): that many drops in flight per actor, their links propagated in one launch per stage;
    ``workers`` forked helper processes run the lanes' modem / RF stages.  ``sinc_extension``: fading samples asked to
    propagate with ``InterpolationMode.SINC`` use windowed-sinc fractional delays (an extension: the reference rounds
    whatever the mode, and that stays the default).
    """
    global _allow_fallback
    from . import _lib

    if _lib.device_count() < 1:
        raise _lib.HermesB200Error(_lib.HB_ERR_NO_DEVICE, "no CUDA device visible; hermespy_b200 has no CPU fallback")
    if precision not in ("f32", "f64"):
        raise ValueError("precision must be 'f32' or 'f64'")
    if device is not None:
        if not 0 <= int(device) < _lib.device_count():
            raise ValueError(f"device {device} outside the {_lib.device_count()} visible CUDA devices")
        config.device = int(device)
    config.precision = precision
    config.sinc_extension = bool(sinc_extension)
    _allow_fallback = bool(allow_reference_fallback)
    patch_reference()
    config.batch_drops, config.workers = max(0, int(batch_drops)), max(0, int(workers))
    from . import runner

    if config.batch_drops > 0:
        runner.patch_actor()
    else:
        runner.unpatch_actor()


def patch_reference() -> None:
    """Rebind the reference methods (split from ``enable`` so host-side tests can patch without a GPU)."""
    from hermespy.channel.cdl.cluster_delay_lines import ClusterDelayLineSample  # type: ignore
    from hermespy.channel.fading.fading import MultipathFadingSample  # type: ignore

    if not _ORIGINALS:
        _ORIGINALS["fading_propagate"] = MultipathFadingSample._propagate
        _ORIGINALS["fading_state"] = MultipathFadingSample.state
        _ORIGINALS["cdl_propagate"] = ClusterDelayLineSample._propagate
        _ORIGINALS["cdl_state"] = ClusterDelayLineSample.state
    ClusterDelayLineSample.state = _cdl_state
    MultipathFadingSample._propagate = _fading_propagate
    MultipathFadingSample.state = _fading_state
    ClusterDelayLineSample._propagate = _cdl_propagate


def disable() -> None:
    from . import runner

    runner.unpatch_actor()
    config.batch_drops = config.workers = 0
    config.sinc_extension = False
    if not _ORIGINALS:
        return
    from hermespy.channel.cdl.cluster_delay_lines import ClusterDelayLineSample  # type: ignore
    from hermespy.channel.fading.fading import MultipathFadingSample  # type: ignore

    MultipathFadingSample._propagate = _ORIGINALS["fading_propagate"]
    MultipathFadingSample.state = _ORIGINALS["fading_state"]
    ClusterDelayLineSample.state = _ORIGINALS["cdl_state"]
    ClusterDelayLineSample._propagate = _ORIGINALS["cdl_propagate"]
    _ORIGINALS.clear()
