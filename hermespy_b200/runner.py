"""Batched drop runner: many Monte-Carlo drops of an UNMODIFIED HermesPy scenario in flight, ONE channel launch per stage
(SURVEY 8(f)-1).

The reference runs one drop at a time per Ray actor (hermespy/core/pymonte/actors.py:340-441): seven stages, all Python /
numpy, and inside stage five one ``propagate`` call per link (hermespy/simulation/scenario.py:576-609).  A GPU served that
way sees batches of one.  Here the actor that ``Simulation.run()`` creates keeps ``B`` *lanes* -- deep copies of its
``(scenario, grid dimensions, evaluators)`` tuple, exactly what Ray ships to every actor (monte_carlo.py:363-365), each
with its own seeds -- and runs them stage-synchronously:

    stages 0-3   realize_channels, sample_states, transmit_operators, generate_outputs   per lane (reference code)
    stage  4     the link loop of scenario.py:576-609, restated: every lane samples its links (reference code) and hands
                 the (sample parameters, transmit block) pairs over; ALL pairs of ALL lanes that share a delay structure
                 go to the device in one ``hb_fading_propagate_host`` / ``hb_cdl_propagate_host`` call
    stages 5-6   process_inputs, receive_operators, evaluators                           per lane (reference code)

Lanes can live in ``workers`` forked helper processes (the modem / RF stages are single-threaded Python: this is what
uses the host cores), while the process that owns the GPU gathers their link requests and launches.  The Monte-Carlo
engine around the actor (queue manager, collector, result containers, confidence stop) is the reference's own.

Enable with ``hermespy_b200.dropin.enable(batch_drops=B, workers=W)``; ``Simulation.run()`` scripts need no change.
Statistically a run with B lanes equals a reference run with B actors: every lane owns independent random roots.
With the channel patch disabled the same runner drives the reference's numpy channel code through the same lanes and
seeds, which is how the tests pin it (identical artifacts, bit for bit, in the float64 mode).
"""
from __future__ import annotations

import copy
import multiprocessing as mp
import os
import queue
import threading
import time
import traceback
from typing import Any, List, Sequence

import numpy as np

from . import config

LANE_SEED_STRIDE = 12345678  # the reference's per-actor stride (hermespy/simulation/simulation.py:220-223)
_PROPAGATE_STAGE = 4
_STAGES = ("realize_channels", "sample_states", "transmit_operators", "generate_outputs", "propagate", "process_inputs",
           "receive_operators")


# ---- stage 4, restated -----------------------------------------------------------------------------------------------

class _Pending(object):
    """One propagate call of the link loop whose blocks went to the device: where its result belongs."""

    __slots__ = ("row", "col", "first", "count", "sample", "signal", "offsets")

    def __init__(self, row, col, first, count, sample, signal, offsets):
        self.row, self.col, self.first, self.count = row, col, first, count
        self.sample, self.signal, self.offsets = sample, signal, offsets


def _gpu_kind(sample) -> str | None:
    """'fading' / 'cdl' when the drop-in serves this sample type on the device, else None."""
    from . import dropin

    if not dropin.enabled():
        return None
    from hermespy.channel.cdl.cluster_delay_lines import ClusterDelayLineSample  # type: ignore
    from hermespy.channel.fading.fading import MultipathFadingSample  # type: ignore

    if isinstance(sample, MultipathFadingSample):
        return "fading"
    if isinstance(sample, ClusterDelayLineSample):
        return "cdl"
    return None


def _submit(matrix, pending, requests, row, col, sample, transmission, interpolation) -> None:
    """``matrix[row, col] = sample.propagate(transmission)`` -- now, or after the batched launch.

    For device-served samples this restates the wrapper ``ChannelSample.propagate`` (hermespy/channel/channel.py:344-379):
    signal type resolution, the zero-energy short circuit, the stream-count check, one ``_propagate`` per signal block."""
    from hermespy.core import DeviceOutput, InterpolationMode, Signal  # type: ignore

    from . import dropin

    kind = _gpu_kind(sample)
    if kind is None:
        matrix[row, col] = sample.propagate(transmission, interpolation)
        return
    if isinstance(transmission, DeviceOutput):
        signal = transmission.mixed_signal
    elif isinstance(transmission, Signal):
        signal = transmission
    else:
        raise ValueError("Signal is of unsupported type")
    if sample.expected_energy_scale <= 0.0:
        matrix[row, col] = Signal.Empty(signal.sampling_rate, sample.num_receive_antennas, 0,
                                        carrier_frequency=signal.carrier_frequency, noise_power=signal.noise_power,
                                        delay=signal.delay)
        return
    if signal.num_streams != sample.num_transmit_antennas:
        raise ValueError("Number of signal streams to be propagated does not match the number of transmitter antennas "
                         f"({signal.num_streams} != {sample.num_transmit_antennas}))")
    try:
        if kind == "fading":
            block = dropin.fading_block_from_reference(
                sample, sinc=config.sinc_extension and interpolation == InterpolationMode.SINC)
        else:
            block = dropin.cdl_block_from_reference(sample)
    except dropin.UnsupportedByKernels as e:
        matrix[row, col] = _serve_unsupported(kind, e, sample, signal, interpolation)
        return
    first = len(requests)
    offsets = []
    for b in signal.blocks:
        offsets.append(b.offset if kind == "fading" else b._offset)
        zero = kind == "cdl" and interpolation != InterpolationMode.NEAREST  # cluster_delay_lines.py:547: nothing accumulates
        requests.append((kind, block, np.ascontiguousarray(np.asarray(b, dtype=np.complex128)), zero))
    pending.append(_Pending(row, col, first, len(offsets), sample, signal, offsets))


def _serve_unsupported(kind, error, sample, signal, interpolation):
    """A sample without a device model: hard error unless the user opted into the reference fallback (dropin)."""
    from hermespy.core import Signal  # type: ignore

    from . import dropin

    blocks = [dropin._unsupported(f"{kind}_propagate", str(error), sample, b, interpolation) for b in signal.blocks]
    return Signal.Create(blocks, sample.bandwidth, sample.carrier_frequency, signal.noise_power, signal.delay,
                         offsets=[b.offset for b in blocks])


def collect_links(scenario, transmissions, device_states, channel_realizations, timestamp: float = 0.0, interpolation=None):
    """The link loop of ``SimulationScenario.propagate`` (hermespy/simulation/scenario.py:576-609) with the device-served
    ``propagate`` calls deferred: returns ``(matrix, pending, requests)``.  Sampling order -- and therefore every random
    draw and every sample hook -- is the reference's."""
    from hermespy.core import InterpolationMode  # type: ignore

    interpolation = InterpolationMode.NEAREST if interpolation is None else interpolation
    devices = scenario.devices
    n = len(devices)
    if len(transmissions) != n:
        raise ValueError(f"Number of transmit signals ({len(transmissions)}) does not match the number of registered devices ({n})")
    matrix = np.empty((n, n), dtype=np.object_)
    pending: List[_Pending] = []
    requests: list = []
    channels = scenario.channels
    for a, (alpha, alpha_state) in enumerate(zip(devices, device_states)):
        for b, (beta, beta_state) in enumerate(zip(devices[: 1 + a], device_states[: 1 + a])):
            realization = channel_realizations[channels.index(scenario.channel(alpha, beta))]
            ab = realization.sample(alpha_state, beta_state, timestamp)
            _submit(matrix, pending, requests, b, a, ab, transmissions[a], interpolation)
            if a == b:
                continue
            ba = realization.reciprocal_sample(ab, beta_state, alpha_state)
            _submit(matrix, pending, requests, a, b, ba, transmissions[b], interpolation)
    return matrix, pending, requests


def finish_links(matrix, pending: Sequence[_Pending], results: Sequence[np.ndarray]) -> None:
    """Wrap the propagated blocks exactly like ``ChannelSample.propagate`` does (channel.py:369-379)."""
    from hermespy.core import Signal  # type: ignore
    from hermespy.core.signal_model import SignalBlock  # type: ignore

    for p in pending:
        blocks = []
        for k in range(p.count):
            y = results[p.first + k]
            blocks.append(SignalBlock(y.shape[0], y.shape[1], p.offsets[k], np.ascontiguousarray(y).tobytes()))
        matrix[p.row, p.col] = Signal.Create(blocks, p.sample.bandwidth, p.sample.carrier_frequency, p.signal.noise_power,
                                             p.signal.delay, offsets=[b.offset for b in blocks])


def propagate_requests(requests: Sequence[tuple], precision: str | None = None, device: int | None = None) -> List[np.ndarray]:
    """All (kind, parameter block, x, zero) requests -> propagated blocks, ONE host-buffer launch per delay structure.

    ``hb_fading_propagate_host`` takes a launch-uniform delay table: fading requests are grouped by (delay table, antenna
    counts, block length).  CDL requests are grouped by array geometry and block length only: realizations with their own
    cluster delays / cluster counts / line-of-sight state (the stochastic 3GPP scenarios) travel as ONE heterogeneous batch
    with per-link delay tables (``hb_cdl_problem.link_term_delay``)."""
    from .kernels import CdlBlock, cdl_propagate_host, fading_propagate_host

    precision = config.precision if precision is None else precision
    device = config.device if device is None else device
    out: List[Any] = [None] * len(requests)
    groups: dict = {}
    for i, (kind, blk, x, zero) in enumerate(requests):
        if kind == "fading":
            nrx = blk["spatial"].shape[0]
            Tout = x.shape[1] + blk["max_delay"]
            if Tout <= 0 or nrx == 0:  # fading.py:381,394-397
                out[i] = np.zeros((nrx, Tout), dtype=np.complex128)
                continue
            key = ("fading", blk["tap_delay"].tobytes(), blk["max_delay"], blk["omega"].shape, blk["spatial"].shape, x.shape)
        else:
            if zero:
                out[i] = np.zeros((blk.num_rx, x.shape[1] + blk.max_delay), dtype=np.complex128)
                continue
            key = ("cdl", blk.geometry_key(), x.shape)  # delay structures may differ: CdlBlock.stack pads, per-link tables
        groups.setdefault(key, []).append(i)
    for key, idx in groups.items():
        x = np.stack([requests[i][2] for i in idx])
        if key[0] == "fading":
            b0 = requests[idx[0]][1]
            stack = lambda f: np.stack([requests[i][1][f] for i in idx])
            y = fading_propagate_host(x, b0["tap_delay"], b0["max_delay"], stack("omega"), stack("phi"), stack("amp"),
                                      stack("spatial"), omega_max=max(requests[i][1]["omega_max"] for i in idx),
                                      precision=precision, sos_mode=config.sos_mode, device=device)
        else:
            blk = CdlBlock.stack([requests[i][1] for i in idx])
            y = cdl_propagate_host(x, blk, precision=precision, device=device)
            if blk.link_max_delay is not None:  # heterogeneous batch: every link keeps its own T + max_delay samples
                T = x.shape[2]
                for k, i in enumerate(idx):
                    out[i] = np.ascontiguousarray(y[k][:, : T + int(blk.link_max_delay[k])])
                continue
        for k, i in enumerate(idx):
            out[i] = y[k]
    return out


# ---- lanes -----------------------------------------------------------------------------------------------------------

def _reseed_random_roots(objects, seed: int) -> None:
    """Give every random root among ``objects`` (RandomNode without a mother node: the scenario, modems and noise models
    the reference leaves as independent roots) its own deterministic seed, so that lanes do not replay each other."""
    from hermespy.core.random_node import RandomNode  # type: ignore

    k = 0
    seen = set()
    for o in objects:
        if isinstance(o, RandomNode) and id(o) not in seen and o.random_mother is None:
            seen.add(id(o))
            o.seed = int(seed) + k
            k += 1


class Lane(object):
    """One clone of the investigated tuple ``(scenario, grid, evaluators)`` with the reference's stage runner on it."""

    def __init__(self, scenario, grid, evaluators, stage_arguments=None) -> None:
        from hermespy.simulation.simulation import SimulationRunner  # type: ignore

        self.scenario, self.grid, self.evaluators = scenario, grid, evaluators
        self.runner = SimulationRunner(scenario)
        self.stage_arguments = stage_arguments or {}
        self.recent = None
        self._matrix = self._pending = None
        self._first, self._last = 0, len(_STAGES) - 1

    @classmethod
    def clone_of(cls, scenario, grid, evaluators, lane_index: int, base_seed: int, stage_arguments=None) -> "Lane":
        """Deep copy of the tuple AS ONE OBJECT (dimensions and evaluators keep pointing into their own scenario copy --
        what Ray's serialization does per actor, monte_carlo.py:365), then independent seeds for every random root."""
        memo: dict = {}
        sc, gr, ev = copy.deepcopy((scenario, grid, evaluators), memo)
        seed = int(base_seed) + int(lane_index) * LANE_SEED_STRIDE
        _reseed_random_roots([sc] + [o for o in memo.values() if not isinstance(o, (list, tuple, dict))], seed)
        return cls(sc, gr, ev, stage_arguments)

    # -- the reference's section bookkeeping (actors.py:381-424), per lane ----------------------------------------------
    def configure(self, section) -> None:
        idx = np.asarray(section, dtype=int)
        n = len(_STAGES)
        if self.recent is None:
            for d, i in enumerate(idx):
                self.grid[d].configure_point(int(i))
            changed = np.array([], dtype=int)
        else:
            changed = np.argwhere(idx != self.recent).flatten()
            for d in changed:
                self.grid[int(d)].configure_point(int(idx[d]))
        first, last = n, 0
        for d in changed:
            dim = self.grid[int(d)]
            if dim.first_impact is None:
                first = 0
            elif dim.first_impact in _STAGES:
                first = min(first, _STAGES.index(dim.first_impact))
            if dim.last_impact is None:
                last = n - 1
            elif dim.last_impact in _STAGES:
                last = max(last, _STAGES.index(dim.last_impact))
        self._first = 0 if first >= n else first
        self._last = n - 1 if last <= 0 else last
        self.recent = idx

    def before_propagate(self) -> list:
        """Stages up to the link loop; returns this lane's device requests (possibly none)."""
        r = self.runner
        for s, stage in enumerate((r.realize_channels, r.sample_states, r.transmit_operators, r.generate_outputs)):
            if self._first <= s <= self._last:
                stage()
        self._matrix = self._pending = None
        if not self._first <= _PROPAGATE_STAGE <= self._last:
            return []
        states = r._SimulationRunner__device_states
        realizations = r._SimulationRunner__channel_realizations
        outputs = r._SimulationRunner__device_outputs
        if states is None or realizations is None:
            raise RuntimeError("Propagation simulation stage called without prior channel or device realization")
        if outputs is None:
            raise RuntimeError("Propagation simulation stage called without prior device transmission")
        self._matrix, self._pending, requests = collect_links(self.scenario, outputs, states, realizations)
        return requests

    def after_propagate(self, results: Sequence[np.ndarray]) -> list:
        r = self.runner
        if self._matrix is not None:
            finish_links(self._matrix, self._pending, results)
            r._SimulationRunner__propagation = self._matrix.tolist()
            self._matrix = self._pending = None
        for s, stage in ((5, r.process_inputs), (6, r.receive_operators)):
            if self._first <= s <= self._last:
                stage()
        return [e.evaluate().artifact() for e in self.evaluators]


# ---- helper processes for the CPU stages -----------------------------------------------------------------------------------

_rpc_conn = None  # inside a lane worker: the pipe to the process that owns the GPU
_rpc_stash: list = []  # work messages that arrived while a device call was waiting for its answer (pipelined rounds)
_rpc_shm = None  # inside a lane worker: its shared-memory window for the large results of device calls

#: bytes of the shared-memory window each helper offers for device-call results (ideal-CSI ``state()`` arrays: megabytes
#: per drop).  The GPU owner copies device -> window and only a descriptor crosses the pipe.  0 disables the windows
#: (results are pickled through the pipe).
RPC_SHM_BYTES = int(os.environ.get("HB_RPC_SHM_BYTES", 64 << 20))


class _ShmRef(object):
    """Descriptor of an array that lives in the helper's shared-memory window."""

    __slots__ = ("offset", "shape", "dtype")

    def __init__(self, offset, shape, dtype):
        self.offset, self.shape, self.dtype = int(offset), tuple(shape), str(dtype)


def _window_bytes(helpers: int) -> int:
    """Window size per helper: ``RPC_SHM_BYTES``, cut down to what /dev/shm can actually back (half of its free space over
    all helpers) -- pages of a shared-memory segment are allocated on first touch, and touching more than the tmpfs holds is
    a SIGBUS, not an exception (container defaults are as small as 64 MB).  Below 1 MB per helper: no windows."""
    want = int(RPC_SHM_BYTES)
    if want <= 0 or helpers <= 0:
        return 0
    try:
        st = os.statvfs("/dev/shm")
        free = int(st.f_bavail) * int(st.f_frsize)
    except OSError:
        return 0
    size = min(want, free // (2 * helpers))
    return size if size >= (1 << 20) else 0


class _Window(object):
    """One helper's shared-memory window, created by the GPU owner BEFORE the fork (the helper inherits the mapping); the
    owner unlinks it, whatever happens to the helper.  Deliberately NOT page-locked: ``cudaHostRegister`` of 16 x 64 MB costs
    0.5 .. 1.5 s per ``Simulation.run()`` and holds the context lock against the propagate launches when done in the
    background (measured: OFDM campaign 267 -> 141 drops/s), for 0.35 ms per drop of faster copies the owner does not need."""

    def __init__(self, nbytes: int) -> None:
        from multiprocessing import shared_memory

        self.shm = shared_memory.SharedMemory(create=True, size=int(nbytes))

    def release(self) -> None:
        for step in (self.shm.close, self.shm.unlink):
            try:
                step()
            except Exception:  # views handed out earlier may still hold the mapping: the name is gone, the pages follow
                pass


def _serve_device_call(name: str, args, shm) -> Any:
    """One helper's device call, in the GPU owner.  With a window, large results are written into it and replaced by
    ``_ShmRef`` descriptors."""
    from . import dropin

    import inspect

    call = dropin.DEVICE_CALLS[name]
    if shm is None or name not in dropin.DEVICE_CALLS_WITH_OUT or "out_alloc" not in inspect.signature(call).parameters:
        return call(*args)
    used = [0]
    handed: list = []

    def out_alloc(shape, dtype):
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        start = (used[0] + 255) & ~255
        if start + nbytes > shm.size:
            return None  # does not fit: this array travels through the pipe
        used[0] = start + nbytes
        arr = np.ndarray(shape, dtype=dtype, buffer=shm.buf, offset=start)
        handed.append((arr, start))
        return arr

    result = call(*args, out_alloc=out_alloc)

    def pack(o):
        if isinstance(o, tuple):
            return tuple(pack(v) for v in o)
        for arr, start in handed:
            if o is arr:
                return _ShmRef(start, arr.shape, arr.dtype.str)
        return o

    return pack(result)


def _unpack_shm(o):
    """Helper side: descriptors -> array views on the own window (valid until the next device call of this process)."""
    if isinstance(o, tuple):
        return tuple(_unpack_shm(v) for v in o)
    if isinstance(o, _ShmRef):
        return np.ndarray(o.shape, dtype=np.dtype(o.dtype), buffer=_rpc_shm.buf, offset=o.offset)
    return o


def device_call(name: str, *args):
    """Run one of ``dropin.DEVICE_CALLS`` -- here, or, from inside a lane worker (a forked process must not touch CUDA),
    in the process that owns the GPU.  Serves the calls that are not part of the batched link loop: ``state()`` for ideal
    channel estimation, ``propagate`` from sample hooks."""
    if _rpc_conn is None:
        from . import dropin

        return dropin.DEVICE_CALLS[name](*args)
    _rpc_conn.send(("rpc", (name, args)))
    while True:
        msg = _rpc_conn.recv()
        if msg[0] in ("rpc_ok", "rpc_error"):
            break
        _rpc_stash.append(msg)  # the owner already queued this helper's next stage: keep it for the main loop
    if msg[0] != "rpc_ok":
        raise RuntimeError(f"device call {name} failed in the GPU process: {msg[1]}")
    return _unpack_shm(msg[1])


def _worker_main(conn, make_lanes, window=None) -> None:
    """Serve the lanes of one helper process: ('pre', {lane: section}) -> requests; ('post', {lane: results}) -> artifacts.
    The helper builds its own lanes (``make_lanes()``: deep copies of the tuple it inherited by fork), so the copies of all
    helpers are made in parallel instead of one after the other in the parent.  ``window``: the shared-memory object the
    GPU owner writes large device-call results into (inherited mapping)."""
    global _rpc_conn, _rpc_shm
    _rpc_conn = conn
    _rpc_shm = window
    lanes = make_lanes()
    try:  # one thread per helper: the helpers ARE the parallelism (BLAS / OpenMP pools would oversubscribe the cores).
        # NOT through torch: ``torch.set_num_threads`` in a forked child whose parent already ran a torch thread pool
        # blocks ~5 s (measured: tools/experiments/runner_overhead.py) and the helpers never call torch.
        from threadpoolctl import threadpool_limits

        threadpool_limits(1)
    except Exception:
        pass
    while True:
        try:
            msg = _rpc_stash.pop(0) if _rpc_stash else conn.recv()
        except EOFError:
            return
        op, tag, payload = msg
        try:
            if op == "stop":
                return
            if op == "pre":
                reply = {}
                for lid, section in payload.items():
                    lanes[lid].configure(section)
                    reply[lid] = lanes[lid].before_propagate()
            else:
                reply = {lid: lanes[lid].after_propagate(results) for lid, results in payload.items()}
            conn.send(("ok", (op, tag, reply)))
        except Exception as e:  # ship the failure to the owner of the GPU; it decides (catch_exceptions)
            conn.send(("error", (op, tag, f"{type(e).__name__}: {e}\n{traceback.format_exc()}")))


def _writer_main(conn, outbox, lock) -> None:
    """Messages to one helper leave through this thread: the GPU owner must keep READING the helpers' replies while a
    large payload (propagated blocks) is still draining into a pipe whose other end is busy sending its own reply --
    two blocking sends facing each other would never finish."""
    while True:
        msg = outbox.get()
        if msg is None:
            return
        try:
            with lock:
                conn.send(msg)
        except Exception:  # the helper is gone or the pipe was closed under us: nothing left to deliver
            return


class LaneSet(object):
    """B lanes, in this process or spread over ``workers`` forked helper processes."""

    def __init__(self, scenario, grid, evaluators, num_lanes: int, workers: int, base_seed: int, stage_arguments=None,
                 first_lane_is_original: bool = True, propagate=None) -> None:
        """``propagate(requests) -> results``: the device call of a round (default ``propagate_requests``)."""
        self.num_lanes = max(1, int(num_lanes))
        self.propagate = propagate_requests if propagate is None else propagate
        t_setup = time.perf_counter()

        def make(k):
            if k == 0 and first_lane_is_original:
                return Lane(scenario, grid, evaluators, stage_arguments)  # lane 0 IS the actor's own tuple and seed
            return Lane.clone_of(scenario, grid, evaluators, k, base_seed, stage_arguments)

        self.local: dict = {}
        self.procs: list = []
        self.outbox: list = []
        self.send_lock: list = []  # per helper pipe: one sender at a time (writer thread / device-call answers)
        self.writers: list = []
        self._mail: dict = {}
        self.windows: list = []  # per helper: _Window or None
        #: where the GPU owner's wall time went (run_stream): waiting for the helpers' stages, the device call, hand-over
        self.seconds = {"setup": 0.0, "sections": 0.0, "wait_pre": 0.0, "propagate": 0.0, "send": 0.0, "wait_post": 0.0}
        workers = min(int(workers), self.num_lanes)
        if workers > 0:
            # One throwaway drop in THIS process first: numba compiles the reference's jitted helpers (resampling, dB
            # conversions, rotations) on first use, and the helpers inherit the compiled code by fork instead of each
            # compiling its own copy.  The clone is discarded: no lane's random stream is touched.
            warm = Lane.clone_of(scenario, grid, evaluators, 10**6, base_seed, stage_arguments)
            warm.after_propagate(self.propagate(warm.before_propagate()))
            del warm
            self.seconds["setup_warm_drop"] = time.perf_counter() - t_setup
            ctx = mp.get_context("fork")  # lanes travel by fork: no pickling of scenarios, exactly the parent's objects
            window_bytes = _window_bytes(workers)
            self.owner = {}
            for w in range(workers):
                mine = list(range(w, self.num_lanes, workers))
                parent, child = ctx.Pipe()
                window = None
                if window_bytes > 0:
                    try:
                        window = _Window(window_bytes)
                    except Exception:
                        window = None  # no /dev/shm: results travel through the pipe
                self.windows.append(window)
                p = ctx.Process(target=_worker_main, daemon=True,
                                args=(child, lambda mine=mine: {k: make(k) for k in mine}, None if window is None else window.shm))
                p.start()
                child.close()
                self.procs.append((p, parent))
                box: queue.SimpleQueue = queue.SimpleQueue()
                self.send_lock.append(threading.Lock())
                th = threading.Thread(target=_writer_main, args=(parent, box, self.send_lock[-1]), daemon=True)
                th.start()
                self.outbox.append(box)
                self.writers.append(th)
                for k in mine:
                    self.owner[k] = w
            self.seconds["setup_fork"] = time.perf_counter() - t_setup - self.seconds.get("setup_warm_drop", 0.0)
        else:
            self.local = {k: make(k) for k in range(self.num_lanes)}
        self.seconds["setup"] = time.perf_counter() - t_setup

    def run_round(self, sections: Sequence[tuple], propagate=None) -> List[list]:
        """One stage-synchronous round: ``sections[k]`` runs on lane k; ``propagate(requests) -> results`` is called ONCE
        with the requests of all lanes.  Returns the artifacts per section."""
        propagate = self.propagate if propagate is None else propagate
        n = len(sections)
        if n > self.num_lanes:
            raise ValueError("more sections than lanes")
        lanes = list(range(n))
        self._send("pre", 0, lanes, sections)
        per_lane = self._await("pre", 0, lanes)
        self._send("post", 0, lanes, self._propagate_split(per_lane, propagate))
        return self._await("post", 0, lanes)

    def run_stream(self, sections, propagate=None, groups: int = 2):
        """Pipelined rounds: the lanes are split into ``groups`` lane groups that alternate, so the helpers compute one
        group's stages while the GPU owner gathers, launches and scatters the other's.  ``sections``: iterable of grid
        sections; yields ``(section, artifacts)`` in completion order."""
        propagate = self.propagate if propagate is None else propagate
        groups = max(1, min(int(groups), self.num_lanes)) if self.procs else 1
        size = (self.num_lanes + groups - 1) // groups
        it = iter(sections)
        pre: dict = {}   # group -> (lanes, sections) whose stages 0-3 are running
        post: dict = {}  # group -> (lanes, sections) whose stages 5-6 are running

        def start(g):
            lanes, secs = [], []
            t_s = time.perf_counter()
            for lane in range(g * size, min(self.num_lanes, (g + 1) * size)):
                try:
                    secs.append(next(it))
                except StopIteration:
                    break
                lanes.append(lane)
            self.seconds["sections"] += time.perf_counter() - t_s
            if lanes:
                self._send("pre", g, lanes, secs)
                pre[g] = (lanes, secs)

        for g in range(groups):
            start(g)
        clock = time.perf_counter
        while pre or post:
            for g in range(groups):
                if g in post:
                    lanes, secs = post.pop(g)
                    t0 = clock()
                    arts = self._await("post", g, lanes)
                    self.seconds["wait_post"] += clock() - t0
                    for sec, art in zip(secs, arts):
                        yield sec, art
                if g in pre:
                    lanes, secs = pre.pop(g)
                    t0 = clock()
                    per_lane = self._await("pre", g, lanes)
                    t1 = clock()
                    slices = self._propagate_split(per_lane, propagate)
                    t2 = clock()
                    self._send("post", g, lanes, slices)
                    post[g] = (lanes, secs)
                    start(g)  # queued behind the post message: the helpers run it as soon as they are through
                    self.seconds["wait_pre"] += t1 - t0
                    self.seconds["propagate"] += t2 - t1
                    self.seconds["send"] += clock() - t2

    @staticmethod
    def _propagate_split(per_lane, propagate):
        flat = [r for reqs in per_lane for r in reqs]
        results = propagate(flat)  # also for an empty round: the caller's accounting sees every round
        slices, o = [], 0
        for reqs in per_lane:
            slices.append(results[o: o + len(reqs)])
            o += len(reqs)
        return slices

    def _send(self, op: str, tag: int, lanes, payloads) -> None:
        """Hand ``payloads[i]`` to lane ``lanes[i]``: a message per helper process, or -- lanes in this process -- run it."""
        if not self.procs:
            out = []
            for lane, payload in zip(lanes, payloads):
                if op == "pre":
                    self.local[lane].configure(payload)
                    out.append(self.local[lane].before_propagate())
                else:
                    out.append(self.local[lane].after_propagate(payload))
            self._mail[(op, tag)] = dict(zip(lanes, out))
            return
        by_worker: dict = {}
        for lane, payload in zip(lanes, payloads):
            by_worker.setdefault(self.owner[lane], {})[lane] = payload
        for w, payload in by_worker.items():
            self.outbox[w].put((op, tag, payload))

    def _await(self, op: str, tag: int, lanes) -> list:
        """Replies for ``lanes`` to the (op, tag) message, serving the helpers' device calls (``device_call``) and stashing
        replies to other messages while waiting."""
        from multiprocessing.connection import wait

        from . import dropin

        box = self._mail.setdefault((op, tag), {})
        conns = {conn: w for w, (_, conn) in enumerate(self.procs)}
        while any(lane not in box for lane in lanes):
            if not conns:
                raise RuntimeError("lane workers are gone")
            for conn in wait(list(conns)):
                try:
                    status, reply = conn.recv()
                except EOFError:
                    raise RuntimeError(f"lane worker {conns.pop(conn)} died") from None
                if status == "rpc":
                    name, args = reply
                    window = self.windows[conns[conn]]
                    # answered from this thread (the helper is waiting): under the pipe's send lock, because the helper's
                    # writer thread may be in the middle of a message and two senders would interleave their bytes.  No
                    # deadlock: a helper that asked is reading its pipe until the answer arrives, so the writer drains.
                    try:
                        shm = None if window is None else window.shm
                        answer = ("rpc_ok", _serve_device_call(name, args, shm))
                    except Exception as e:
                        answer = ("rpc_error", f"{type(e).__name__}: {e}")
                    with self.send_lock[conns[conn]]:
                        conn.send(answer)
                elif status == "ok":
                    rop, rtag, payload = reply
                    self._mail.setdefault((rop, rtag), {}).update(payload)
                else:
                    raise RuntimeError(f"lane worker {conns[conn]} failed in {reply[0]}: {reply[2]}")
        return [box.pop(lane) for lane in lanes]

    def close(self, abort: bool = False) -> None:
        """Stop the helpers.  ``abort``: the stream ended early (an exception, or the campaign was called off while rounds
        were in flight): helpers and writer threads may be blocked on full pipes facing each other, so nobody is asked
        politely -- the helpers are terminated first, which unblocks every sender."""
        if abort:
            for p, _ in self.procs:
                if p.is_alive():
                    p.terminate()
        for box in self.outbox:
            box.put(("stop", 0, None))
            box.put(None)
        deadline = time.monotonic() + 5.0  # one budget for all of them, not one per thread / process
        for p, conn in self.procs:
            p.join(timeout=max(0.0, deadline - time.monotonic()))
            if p.is_alive():
                p.terminate()
                p.join(timeout=1.0)
            conn.close()
        for th in self.writers:  # their pipes are closed now: a blocked send has failed, a waiting get has its None
            th.join(timeout=max(0.1, deadline - time.monotonic()))
        for window in self.windows:
            if window is not None:
                window.release()
        self.procs, self.outbox, self.writers, self.windows, self.send_lock = [], [], [], [], []


# ---- the actor's run loop (replaces MonteCarloActor.run for SimulationActor while the runner is enabled) -----------------

#: accounting of the last / current batched run in this process (tests and the campaign report read it)
stats = {"rounds": 0, "drops": 0, "links": 0, "launch_groups": 0, "max_links_per_round": 0}


def batched_actor_run(self) -> None:
    """``SimulationActor.run`` with ``config.batch_drops`` drops in flight (see module docstring)."""
    from hermespy.core.pymonte.artifact import MonteCarloSample  # type: ignore
    from hermespy.core.pymonte.definitions import UnmatchableException  # type: ignore
    from ray import get, put  # type: ignore

    queue = self._MonteCarloActor__queue_manager
    results = self._MonteCarloActor__results
    stage_arguments = self._MonteCarloActor__stage_arguments
    if stage_arguments:  # stages iterating over argument lists nest drops inside drops: the serial schedule handles them
        return _original_run(self)
    scenario = self._investigated_object

    def propagate(requests):
        stats["rounds"] += 1
        stats["links"] += len(requests)
        stats["max_links_per_round"] = max(stats["max_links_per_round"], len(requests))
        return propagate_requests(requests)

    lanes = LaneSet(scenario, self._MonteCarloActor__grid, self._MonteCarloActor__evaluators, config.batch_drops,
                    config.workers, base_seed=scenario.seed if scenario.seed is not None else 0)
    stats["seconds"] = lanes.seconds  # live: ``Simulation.run()`` returns when the last result is in, before this thread ends
    finished = False

    try:
        def sections():  # the queue hands out one section per active grid point per call, until the campaign is served
            while True:
                batch = get(queue.next_batch.remote())
                if len(batch) < 1:
                    return
                for s_ in batch:
                    yield tuple(s_)

        done: list = []
        try:
            for section, artifacts in lanes.run_stream(sections(), propagate):
                done.append(MonteCarloSample(section, 0, artifacts))
                stats["drops"] += 1
                if len(done) >= lanes.num_lanes:
                    results.append(put(done))
                    done = []
            finished = True
        except Exception as e:
            # The campaign loop collects what it needs and shuts the engine down without waiting for its actors
            # (monte_carlo.py:404-470 kills them): a queue call failing on the closed engine is the normal end of a run
            # that still had rounds in flight, not an error of this actor.
            called_off = isinstance(e, RuntimeError) and "after shutdown" in str(e)
            if not called_off:
                if not self.catch_exceptions:
                    raise UnmatchableException(f"Actor #{self.index} encountered an error during run: {e}") from e
                print(e)
        if done and finished:
            results.append(put(done))
    finally:
        lanes.close(abort=not finished)  # rounds in flight: helpers and writers may be blocked on each other's pipes


_original_run = None


def patch_actor() -> None:
    global _original_run
    from hermespy.core.pymonte.actors import MonteCarloActor  # type: ignore
    from hermespy.simulation.simulation import SimulationActor  # type: ignore

    if _original_run is None:
        _original_run = MonteCarloActor.run
    SimulationActor.run = batched_actor_run  # on the subclass: other Monte-Carlo actors keep the reference loop


def unpatch_actor() -> None:
    global _original_run
    if _original_run is None:
        return
    from hermespy.simulation.simulation import SimulationActor  # type: ignore

    if "run" in SimulationActor.__dict__:
        del SimulationActor.run
    _original_run = None
