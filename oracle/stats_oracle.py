"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the evaluator statistics on the hot path's far end.

* ``bit_errors``  follows ``BitErrorEvaluator.evaluate`` / ``BitErrorEvaluation.artifact``
  (hermespy/modem/evaluators.py:231-259): zero-pad the shorter bit sequence, ``|tx - rx|`` on int8, artifact = mean.
* ``add_artifacts`` follows ``ScalarEvaluationResult.add_artifact`` (hermespy/core/pymonte/scalar.py:101-125):
  per grid cell running ``sum``, ``squared_sum``, ``count`` -- applied drop by drop in index order.
* ``kron_mix`` follows ``MultipathFadingRealization._sample`` (hermespy/channel/fading/fading.py:476-489):
  ``R_rx @ S @ R_tx`` with the covariance matrices themselves.

Pinned against the live reference in tests/test_oracle_vs_reference.py::test_stats_oracle_matches_reference_evaluator.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import numpy as np


def bit_errors(tx_bits, rx_bits):
    """(errors, bits, artifact) of one drop; inputs are 0/1 sequences of possibly different length."""
    t = np.asarray(tx_bits).astype(np.int8)
    r = np.asarray(rx_bits).astype(np.int8)
    n = max(len(t), len(r))
    tp = np.append(t, np.zeros(n - len(t), dtype=np.int8))
    rp = np.append(r, np.zeros(n - len(r), dtype=np.int8))
    e = np.abs(tp - rp)
    return int(e.astype(np.int64).sum()), int(n), float(np.mean(e)) if n else 0.0


def add_artifacts(artifact, cell, num_cells, errors=None, bits=None, stats=None, counts=None):
    """Sequential accumulation (sum, sum^2, count) + exact integer counters per grid cell."""
    stats = np.zeros((num_cells, 3)) if stats is None else stats
    counts = np.zeros((num_cells, 2), dtype=np.int64) if counts is None else counts
    for i, (a, c) in enumerate(zip(artifact, cell)):
        stats[c, 0] = stats[c, 0] + a
        stats[c, 1] = stats[c, 1] + a ** 2
        stats[c, 2] = stats[c, 2] + 1
        if errors is not None:
            counts[c, 0] += int(errors[i])
        if bits is not None:
            counts[c, 1] += int(bits[i])
    return stats, counts


def kron_mix(r_rx, spatial, r_tx):
    s = np.asarray(spatial)
    if r_rx is not None:
        s = np.asarray(r_rx) @ s
    if r_tx is not None:
        s = s @ np.asarray(r_tx)
    return s
