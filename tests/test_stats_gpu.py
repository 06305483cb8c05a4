"""GPU parity: evaluator statistics (K7) and Kronecker mixing (K2) through the C-ABI against the numpy oracle."""
import numpy as np
import pytest

from oracle import stats_oracle as so

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("drops,nbits,cells", [(1, 1, 1), (5, 37, 3), (200, 1000, 11), (1031, 256, 77)])
def test_bit_errors_and_statistics_match_oracle(drops, nbits, cells):
    import torch
    from hermespy_b200.montecarlo import GridStatistics

    rng = np.random.default_rng(drops)
    tx = rng.integers(0, 2, (drops, nbits)).astype(np.uint8)
    rx = tx ^ (rng.random((drops, nbits)) < 0.03).astype(np.uint8)
    tl = rng.integers(0, nbits + 1, drops).astype(np.int32)
    rl = rng.integers(0, nbits + 1, drops).astype(np.int32)
    cell = rng.integers(0, cells, drops).astype(np.int32)
    want = [so.bit_errors(tx[i, : tl[i]], rx[i, : rl[i]]) for i in range(drops)]
    we, wb, wa = (np.array(v) for v in zip(*want))

    gs = GridStatistics((cells,), device="cuda")
    d = lambda a: torch.from_numpy(a).cuda()
    e, b, a = gs.accumulate_bits(d(tx), d(rx), d(cell), d(tl), d(rl))
    np.testing.assert_array_equal(e.cpu().numpy(), we)  # integer work: bit exact
    np.testing.assert_array_equal(b.cpu().numpy(), wb)
    np.testing.assert_array_equal(a.cpu().numpy(), wa)  # one IEEE division each
    st, ct = so.add_artifacts(wa, cell, cells, we, wb)
    np.testing.assert_array_equal(gs.counts.cpu().numpy(), ct)
    np.testing.assert_allclose(gs.stats.cpu().numpy(), st, rtol=1e-13, atol=1e-300)
    np.testing.assert_array_equal(gs.stats.cpu().numpy()[:, 2], st[:, 2])
    # a second batch accumulates on top; the result is reproducible bit for bit (fixed summation order)
    gs2 = GridStatistics((cells,), device="cuda")
    gs2.accumulate_bits(d(tx), d(rx), d(cell), d(tl), d(rl))
    np.testing.assert_array_equal(gs2.stats.cpu().numpy(), gs.stats.cpu().numpy())
    gs.accumulate(a, d(cell), e, b)
    np.testing.assert_array_equal(gs.counts.cpu().numpy(), 2 * ct)


def test_equal_length_default_and_empty():
    import torch
    from hermespy_b200.montecarlo import GridStatistics

    gs = GridStatistics((2, 2), device="cuda")
    tx = torch.tensor([[0, 1, 1, 0], [1, 1, 1, 1]], dtype=torch.uint8, device="cuda")
    rx = torch.tensor([[0, 1, 0, 0], [0, 0, 1, 1]], dtype=torch.uint8, device="cuda")
    e, b, a = gs.accumulate_bits(tx, rx, torch.tensor([3, 3], device="cuda"))
    assert e.tolist() == [1, 2] and b.tolist() == [4, 4] and a.tolist() == [0.25, 0.5]
    assert gs.mean().shape == (2, 2) and gs.mean()[1, 1] == 0.375 and gs.bit_error_rate()[1, 1] == 0.375
    gs.accumulate(torch.zeros(0, dtype=torch.float64, device="cuda"), torch.zeros(0, dtype=torch.int32, device="cuda"))
    with pytest.raises(ValueError):
        gs.accumulate(torch.zeros(1, dtype=torch.float64, device="cuda"), torch.tensor([4], device="cuda"))


@pytest.mark.parametrize("nrx,ntx", [(1, 1), (2, 4), (4, 4), (10, 10), (64, 64)])
def test_kron_mix_matches_oracle(nrx, ntx):
    import torch
    from hermespy_b200.montecarlo import kron_mix

    rng = np.random.default_rng(nrx * 100 + ntx)
    B = 5
    S = np.exp(2j * np.pi * rng.random((B, nrx, ntx)))
    rho = 0.7
    Rrx = rho ** np.abs(np.subtract.outer(np.arange(nrx), np.arange(nrx))) * np.exp(0.3j * np.subtract.outer(np.arange(nrx), np.arange(nrx)))
    Rtx = rho ** np.abs(np.subtract.outer(np.arange(ntx), np.arange(ntx))).astype(complex)
    for rr, rt in [(Rrx, Rtx), (None, Rtx), (Rrx, None), (None, None)]:
        want = np.stack([so.kron_mix(rr, S[b], rt) for b in range(B)])
        got = kron_mix(torch.from_numpy(S).cuda(), None if rr is None else torch.from_numpy(rr),
                       None if rt is None else torch.from_numpy(rt)).cpu().numpy()
        err = np.linalg.norm(got - want) / np.linalg.norm(want)
        assert err < 1e-14, err
