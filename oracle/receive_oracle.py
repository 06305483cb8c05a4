"""Float64 numpy restatement of the receive-side superposition and AWGN (test oracle).  TEST INFRASTRUCTURE ONLY.

Restates, for dense signals of one sampling rate / carrier frequency with whole-sample delays,

* the superposition loop of ``SimulatedDevice.process_input`` .... hermespy/simulation/simulated_device.py:1899-1915
  (``SparseSignal.Empty(...).superimpose(s)`` per impinging signal, then ``to_dense()``)
* ``AWGNRealization.add_to`` ..................................... hermespy/simulation/rf/noise/model.py:140-160

Pinned against the live reference in ``tests/test_receive.py`` (bit-identical for the same generator seed).
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np


def superimpose(signals: Sequence[np.ndarray], offsets: Sequence[int], num_samples: Optional[int] = None) -> np.ndarray:
    """Sum of ``signals[k]`` ([Nrx, T_k]) placed at sample ``offsets[k]``, in the order given, zeros elsewhere."""
    T = max(o + s.shape[1] for o, s in zip(offsets, signals)) if num_samples is None else num_samples
    mixed = np.zeros((signals[0].shape[0], T), dtype=np.complex128)
    for s, o in zip(signals, offsets):
        mixed[:, o: o + s.shape[1]] += s
    return mixed


def noise_normals(seed: int, shape) -> tuple:
    """The two standard-normal planes of one realization, drawn in the reference's order (model.py:147-151)."""
    rng = np.random.default_rng(seed)
    re = rng.standard_normal(shape)
    im = rng.standard_normal(shape)
    return re, im


def add_awgn(signal: np.ndarray, power: float, normals_re: np.ndarray, normals_im: np.ndarray) -> np.ndarray:
    """``signal + (0.5 * power) ** 0.5 * (re + 1j * im)`` (model.py:143-156); zero power returns the signal itself."""
    if power == 0.0:
        return signal
    return signal + (0.5 * power) ** 0.5 * (normals_re + 1j * normals_im)
