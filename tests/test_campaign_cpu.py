"""Host logic of the campaign runner (no GPU): per-drop seeds do not depend on the sharding."""
import numpy as np

from hermespy_b200.campaign import drop_seed
from hermespy_b200.montecarlo import shard_drops


def test_drop_seeds_are_unique_and_sharding_invariant():
    num_drops, cells = 37, 5
    seeds = {(c, d): drop_seed(42, c, d, num_drops) for c in range(cells) for d in range(num_drops)}
    assert len(set(seeds.values())) == cells * num_drops
    assert min(np.diff(sorted(seeds.values()))) >= 4  # scenario, modem and two noise models take seed .. seed + 3
    for world in (1, 2, 3, 8):
        owned = sorted(d for r in range(world) for d in shard_drops(num_drops, r, world))
        assert owned == list(range(num_drops))  # every drop exactly once, whatever the world size
