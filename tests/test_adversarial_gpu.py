"""GPU: inputs a foreign binding could hand the C ABI that ``TownBatch.validate`` would have rejected, and the
largest legal towns.  compute-sanitizer is not available on the pool, so the device-side clamps are pinned by their
results: the launch must finish and equal the oracle run on the input the clamp is documented to reduce it to.

* edge counts above the row stride (``tie_count`` / ``riv_count`` > ``max_edges``) are clamped to the stride;
* entities outside the declared town grid are ignored (the dense-grid form never indexes shared memory with them);
* towns of 1 024 entities and more walk the entity list in passes of 32 (entity form) or stage it in the grid form.
"""

from __future__ import annotations

import numpy as np
import pytest
from helpers import assert_obs_equal, assert_reward_close

from oracle.oracle import oracle_tick
from townlet_b200 import capi
from townlet_b200.params import ObsSpec, RewardSpec
from townlet_b200.soa import pack_entity
from townlet_b200.synth import OBJECT_TYPES, assemble, synth_batch, synth_town

pytestmark = pytest.mark.gpu
FORMS = {"ent": 0, "grid": 1}


def _device_run(batch, obs, rew, n_ticks=2):
    import torch

    from townlet_b200.engine import DeviceTownBatch

    dev = DeviceTownBatch(batch, obs, rew)
    out = dev.alloc_outputs(capi.TL_PHASE_ALL, n_ticks, summary=True, comps=True)
    dev.tick(out, capi.TL_PHASE_ALL, n_ticks)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in vars(out).items() if v is not None}


@pytest.fixture
def form(request):
    capi.set_option("kernel_form", FORMS[request.param])
    yield request.param
    capi.reset_options()


@pytest.mark.parametrize("form", list(FORMS), indirect=True)
@pytest.mark.parametrize("variant", ["hybrid", "compact"])
def test_edge_counts_above_the_row_stride_are_clamped(variant, form) -> None:
    obs, rew = ObsSpec(variant=variant), RewardSpec(fuse_faint=True)
    clean = synth_batch(24, 48, 8, obs)
    k = clean.layout.max_edges
    hostile = clean.copy()
    hostile.tie_count[::3] = 200  # far beyond the row: the kernel must read K entries, not 200
    hostile.riv_count[1::4] = 127
    expect = clean.copy()
    expect.tie_count[::3] = k
    expect.riv_count[1::4] = k
    ref = oracle_tick(expect, obs, rew, capi.TL_PHASE_ALL, 2)
    got = _device_run(hostile, obs, rew)
    assert_obs_equal(got, ref, f"edge-clamp-{variant}-{form}")
    assert_reward_close(got, ref, f"edge-clamp-{variant}-{form}")


@pytest.mark.parametrize("form", list(FORMS), indirect=True)
@pytest.mark.parametrize("variant", ["hybrid", "full", "compact"])
def test_entities_outside_the_grid_are_ignored(variant, form) -> None:
    obs, rew = ObsSpec(variant=variant), RewardSpec(fuse_faint=True)
    clean = synth_batch(32, 48, 8, obs)
    kind = (clean.ent >> 20) & 3
    victims = np.nonzero(kind == capi.TL_KIND_RESERVED)[0][::2]
    assert victims.size > 4
    hostile = clean.copy()
    # reserved tiles far outside the 48 x 48 grid (and outside every window): x = 1000 .. 1023, y = 1001
    hostile.ent[victims] = pack_entity(1000 + (victims % 24), np.full(victims.size, 1001), capi.TL_KIND_RESERVED)
    expect = clean.copy()
    expect.ent[victims] = pack_entity(np.zeros(victims.size, dtype=np.int64), np.zeros(victims.size, dtype=np.int64), capi.TL_KIND_EMPTY)
    with pytest.raises(ValueError):
        hostile.validate()  # the host-side check a Python caller would have hit
    ref = oracle_tick(expect, obs, rew, capi.TL_PHASE_ALL, 2)
    got = _device_run(hostile, obs, rew)
    assert_obs_equal(got, ref, f"out-of-grid-{variant}-{form}")
    assert_reward_close(got, ref, f"out-of-grid-{variant}-{form}")


def _crowded_town(index: int, extra_objects: int):
    """A 128 x 128 town of 64 agents with ``extra_objects`` more objects on free tiles (a quarter of them reserved)."""
    town = synth_town(index, 128, 64)
    rng = np.random.default_rng(99 + index)
    taken = {(int(x), int(y)) for x, y in town.positions} | {(int(x), int(y)) for x, y in town.object_positions}
    free = [(x, y) for y in range(128) for x in range(128) if (x, y) not in taken]
    picks = rng.choice(len(free), size=extra_objects, replace=False)
    base = len(town.object_ids)
    new_pos = np.array([free[int(i)] for i in picks], dtype=np.int64)
    town.object_positions = np.concatenate([town.object_positions, new_pos])
    for j in range(extra_objects):
        kind = OBJECT_TYPES[(base + j) % 4]
        town.object_types.append(kind)
        town.object_ids.append(f"{kind}_{base + j}")
        if j % 4 == 0:
            town.reservations[town.object_ids[-1]] = town.agent_ids[j % 64]
    return town


@pytest.mark.parametrize("form", list(FORMS), indirect=True)
@pytest.mark.parametrize("variant", ["hybrid", "compact"])
def test_towns_of_more_than_1024_entities(variant, form) -> None:
    obs, rew = ObsSpec(variant=variant), RewardSpec(fuse_faint=True)
    batch = assemble([_crowded_town(0, 1100), synth_town(1, 128, 64), _crowded_town(2, 2050)], obs)
    batch.validate()
    assert np.diff(batch.town_ent_off).max() >= 2048
    ref = oracle_tick(batch.copy(), obs, rew, capi.TL_PHASE_ALL, 2)
    got = _device_run(batch.copy(), obs, rew)
    assert_obs_equal(got, ref, f"crowded-{variant}-{form}")
    assert_reward_close(got, ref, f"crowded-{variant}-{form}")
