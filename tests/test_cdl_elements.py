"""Non-ideal antenna elements in the CDL path: Dipole / PatchAntenna / LinearAntenna, uniform arrays and
custom arrays with per-element slant and orientation, against golden vectors of the UNMODIFIED reference
(``oracle/make_golden_elements.py`` -> ``tests/golden/cdl_elements_golden.npz``).

CPU: the numpy oracle reproduces the reference.  GPU: ``hb_cdl_propagate_host`` / ``hb_cdl_state`` reproduce it in both
precisions (f64 <= 1e-10, f32 <= 1e-5 relative L2), through the rank-one (uniform elements) and rank-two (per element)
ray-matrix paths.
"""
import os

import numpy as np
import pytest

from oracle import cdl_oracle as co
from oracle.golden_cases import golden_signal
from tests.helpers import cdl_block_from_oracle_params, cdl_params_from_golden, rel_l2

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "cdl_elements_golden.npz")
CASES = ["dipole_uniform_4x2", "patch_tx_ideal_rx", "xpol_linear_custom", "mixed_rotated_los"]
MODES = {"dipole_uniform_4x2": 1, "patch_tx_ideal_rx": 1, "xpol_linear_custom": 2, "mixed_rotated_los": 2}


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


@pytest.mark.parametrize("ci", range(len(CASES)), ids=CASES)
def test_oracle_reproduces_reference_with_element_models(golden, ci):
    name = CASES[ci]
    p = cdl_params_from_golden(golden, name)
    ref = golden[f"{name}/y"]
    T = ref.shape[1] - co.max_delay_in_samples(p)
    y = co.propagate(p, golden_signal(500 + ci, p.tx.topology.shape[0], T))
    assert y.shape == ref.shape and rel_l2(y, ref) < 5e-12
    csi = golden[f"{name}/csi"]
    mine = co.state(p, T, 1000)
    assert mine.shape == csi.shape and np.abs(mine - csi).max() <= 5e-12 * np.abs(csi).max()


@pytest.mark.parametrize("ci", range(len(CASES)), ids=CASES)
def test_block_layout_and_element_mode(golden, ci):
    name = CASES[ci]
    p = cdl_params_from_golden(golden, name)
    blk = cdl_block_from_oracle_params(p)
    assert blk.element_mode == MODES[name]
    assert blk.term_delay.shape[0] == len(co.ray_terms(p)) - (1 if p.line_of_sight else 0)
    # an array of unrotated ideal elements collapses to the table-free fast path
    from hermespy_b200.kernels import ideal_elements
    from dataclasses import replace

    ideal = replace(blk, tx_elements=ideal_elements(blk.num_tx), rx_elements=ideal_elements(blk.num_rx))
    assert ideal.element_mode == 0 and ideal.tx_elements is None


@pytest.mark.gpu
@pytest.mark.parametrize("ci", range(len(CASES)), ids=CASES)
def test_kernels_reproduce_reference_with_element_models(golden, ci):
    import torch

    from hermespy_b200 import _lib
    from hermespy_b200.kernels import CdlDeviceBlock, cdl_propagate_host, cdl_state

    name = CASES[ci]
    p = cdl_params_from_golden(golden, name)
    blk = cdl_block_from_oracle_params(p)
    ref = golden[f"{name}/y"]
    T = ref.shape[1] - blk.max_delay
    x = golden_signal(500 + ci, blk.num_tx, T)[None]
    before = _lib.launch_counts()
    y64 = cdl_propagate_host(x, blk, precision="f64")[0]
    y32 = cdl_propagate_host(x, blk, precision="f32")[0]
    after = _lib.launch_counts()
    assert after["cdl_rays"] - before["cdl_rays"] >= 3 and after["cdl_propagate"] - before["cdl_propagate"] >= 2
    assert y64.shape == ref.shape
    assert rel_l2(y64, ref) < 1e-10
    assert rel_l2(y32, ref) < 1e-5
    h, gd = cdl_state(CdlDeviceBlock(blk, device="cuda"), T)
    csi = golden[f"{name}/csi"]
    dense = np.zeros_like(csi)
    hh = h[0].cpu().numpy()
    for g, d in enumerate(gd):
        dense[:, :, :, d] = hh[g]
    assert np.abs(dense - csi).max() < 1e-10 * np.abs(csi).max()
    torch.cuda.synchronize()


@pytest.mark.gpu
def test_batched_mixed_links_with_element_models(golden):
    """Several links of one delay structure in one launch, element tables shared by the batch."""
    from hermespy_b200.kernels import CdlBlock, cdl_propagate_host

    name = "xpol_linear_custom"
    p = cdl_params_from_golden(golden, name)
    rng = np.random.default_rng(3)
    blocks, xs, refs = [], [], []
    T = 300
    for b in range(3):
        from dataclasses import replace

        q = replace(p, aoa=p.aoa + 0.05 * b, zod=p.zod - 0.02 * b,
                    rx=replace(p.rx, velocity=p.rx.velocity * (1 + b), translation=p.rx.translation + np.array([b, 0.0, 0.0])))
        blocks.append(cdl_block_from_oracle_params(q))
        x = (rng.standard_normal((4, T)) + 1j * rng.standard_normal((4, T))) / np.sqrt(2)
        xs.append(x)
        refs.append(co.propagate(q, x))
    blk = CdlBlock.stack(blocks)
    y64 = cdl_propagate_host(np.stack(xs), blk, precision="f64")
    y32 = cdl_propagate_host(np.stack(xs), blk, precision="f32")
    for b in range(3):
        assert rel_l2(y64[b], refs[b]) < 1e-10 and rel_l2(y32[b], refs[b]) < 1e-5
