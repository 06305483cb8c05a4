"""In-process stand-in for the subset of ``ray`` that HermesPy's Monte-Carlo engine uses.

``hermespy.core.pymonte`` drives a campaign with three kinds of Ray actors -- a queue manager, N simulation actors and a
collector (monte_carlo.py:355-371, actors.py) -- through ``ray.remote(cls).options(...).remote(*args)``,
``handle.method.remote(*args)``, ``ray.get`` / ``ray.wait`` / ``ray.put``.  Here every actor is an object in THIS process
with its own small thread pool (``max_concurrency`` threads, default 1: calls on one actor are serialized exactly as Ray
does), a remote call returns a future-backed ``ObjectRef``, and constructor arguments are deep-copied the way Ray's
serialization would copy them (actor handles stay references), so every simulation actor owns its scenario clone.

With one process per GPU (``torchrun``) and the channel hot path on the device (``hermespy_b200.dropin``), this turns
``Simulation.run()`` into the single-process drop loop the GPU path wants: no worker processes that would each need
their own CUDA context.  Install with ``hermespy_b200.shims.ray.install()`` BEFORE ``import hermespy``; it registers this
module as ``ray`` only when the real package is not importable.
"""
from __future__ import annotations

import copy
import sys
import threading
import time
import weakref
from concurrent.futures import FIRST_COMPLETED, Future, ThreadPoolExecutor
from concurrent.futures import wait as _futures_wait
from typing import Any, Sequence

__version__ = "0.0-hermespy_b200-inprocess"
_initialized = False


class ObjectRef(object):
    """Future-backed reference to the result of a remote call (or to a ``put`` value)."""

    __class_getitem__ = classmethod(lambda cls, item: cls)  # ``ObjectRef[list[...]]`` annotations in actors.py

    def __init__(self, future: Future) -> None:
        self._future = future

    def result(self):
        return self._future.result()


def _ready(value: Any) -> ObjectRef:
    f: Future = Future()
    f.set_result(value)
    return ObjectRef(f)


#: first exception raised inside an actor method since the last check.  The reference's driver loop discards the refs
#: of ``actor.run.remote()`` (monte_carlo.py:385-389) and polls the queue manager for progress: a dead actor would make
#: it spin forever.  Here the failure surfaces at the driver's next remote call instead.
_actor_failure: list = []


def _guarded(bound, name):
    def call(*args, **kwargs):
        try:
            return bound(*args, **kwargs)
        except BaseException as e:
            if not _actor_failure:
                _actor_failure.append((name, e))
            raise

    return call


class _RemoteMethod(object):
    def __init__(self, handle: "ActorHandle", name: str) -> None:
        self._handle, self._name = handle, name

    def remote(self, *args, **kwargs) -> ObjectRef:
        if _actor_failure and threading.current_thread() is threading.main_thread():
            name, error = _actor_failure.pop()
            _actor_failure.clear()
            raise RuntimeError(f"actor method {name} failed: {error!r}") from error
        bound = getattr(self._handle._instance, self._name)
        label = f"{type(self._handle._instance).__name__}.{self._name}"
        return ObjectRef(self._handle._pool.submit(_guarded(bound, label), *args, **kwargs))


def _retire_pool(pool) -> None:
    pool.shutdown(wait=False)
    try:
        _live_pools.remove(pool)
    except ValueError:
        pass


_live_pools: list = []  # thread pools of the actors created since the last shutdown()


class ActorHandle(object):
    """Reference to an in-process actor; attribute access yields ``.remote``-callable methods."""

    def __init__(self, instance: Any, max_concurrency: int) -> None:
        self._instance = instance
        self._pool = ThreadPoolExecutor(max_workers=max(1, int(max_concurrency)),
                                        thread_name_prefix=type(instance).__name__)
        _live_pools.append(self._pool)
        # the reference never calls ray.shutdown() (monte_carlo.py:252-262): release the actor's threads when the last
        # handle goes away, so repeated Simulation.run() calls do not accumulate idle threads
        weakref.finalize(self, _retire_pool, self._pool)

    def __getattr__(self, name: str) -> _RemoteMethod:
        if name.startswith("_"):
            raise AttributeError(name)
        return _RemoteMethod(self, name)

    def __deepcopy__(self, memo):  # handles travel by reference, like Ray actor handles
        return self

    def __reduce__(self):
        raise TypeError("in-process actor handles cannot leave the process")


def _copy_argument(value: Any) -> Any:
    """What Ray's serialization does to an actor constructor argument: objects are copied, actor handles are not."""
    return copy.deepcopy(value)


class _RemoteClass(object):
    def __init__(self, cls: type, options: dict | None = None) -> None:
        self._cls, self._options = cls, dict(options or {})

    def options(self, **kwargs) -> "_RemoteClass":
        merged = dict(self._options)
        merged.update(kwargs)
        return _RemoteClass(self._cls, merged)

    def remote(self, *args, **kwargs) -> ActorHandle:
        args = tuple(_copy_argument(a) for a in args)
        kwargs = {k: _copy_argument(v) for k, v in kwargs.items()}
        return ActorHandle(self._cls(*args, **kwargs), self._options.get("max_concurrency", 1))


def remote(*args, **kwargs):
    """``ray.remote(cls)`` and ``@ray.remote(**options)``."""
    if len(args) == 1 and isinstance(args[0], type) and not kwargs:
        return _RemoteClass(args[0])
    return lambda cls: _RemoteClass(cls, kwargs)


def get(refs, timeout: float | None = None):
    if isinstance(refs, ObjectRef):
        return refs._future.result(timeout)
    return [r._future.result(timeout) for r in refs]


def put(value: Any) -> ObjectRef:
    return _ready(value)


def wait(refs: Sequence[ObjectRef], num_returns: int = 1, timeout: float | None = None):
    """Block until ``num_returns`` of ``refs`` are done; returns (ready, pending) in the order of ``refs``."""
    refs = list(refs)
    if not refs:
        return [], []
    num_returns = min(int(num_returns), len(refs))  # Ray raises for num_returns > len(refs); never spin on it
    deadline = None if timeout is None else time.monotonic() + timeout
    while True:
        ready = [r for r in refs if r._future.done()]
        if len(ready) >= num_returns or (deadline is not None and time.monotonic() >= deadline):
            ready = ready[:num_returns] if len(ready) > num_returns else ready
            return ready, [r for r in refs if r not in ready]
        _futures_wait([r._future for r in refs if not r._future.done()], timeout=0.05, return_when=FIRST_COMPLETED)
        # A collector polling its actors in a tight loop (actors.py:189-199) would otherwise starve the actor threads
        # of the interpreter lock; a real cluster pays a network round trip here.
        time.sleep(0.0005)


class _Context(object):
    dashboard_url = None
    address_info: dict = {}


def init(*args, **kwargs) -> _Context:
    global _initialized
    _initialized = True
    return _Context()


def is_initialized() -> bool:
    return _initialized


def shutdown(*args, **kwargs) -> None:
    """End the 'cluster': actor threads are released (pending calls are dropped, running ones finish)."""
    global _initialized
    _initialized = False
    _actor_failure.clear()
    while _live_pools:
        _live_pools.pop().shutdown(wait=False, cancel_futures=True)


def available_resources() -> dict:
    # one actor: the GPU path wants a single drop loop per process (ranks are the unit of parallelism)
    return {"CPU": 1.0}


def cluster_resources() -> dict:
    return available_resources()


def install(force: bool = False) -> bool:
    """Register this module as ``ray`` when the real package is absent (or ``force``).  Returns True if registered."""
    if not force:
        if "ray" in sys.modules and getattr(sys.modules["ray"], "__version__", "") != __version__ \
                and type(sys.modules["ray"]).__name__ == "module":
            return False
        try:
            import importlib.util

            if "ray" not in sys.modules and importlib.util.find_spec("ray") is not None:
                return False
        except (ImportError, ValueError):
            pass
    sys.modules["ray"] = sys.modules[__name__]
    return True
