"""GPU: TL_PHASE_RESERVATIONS — refresh_reservations (world/systems/queues.py:55-86) on the device, against the oracle
(which tests/test_oracle_golden.py pins to the reference's own refresh on the ``reservations_*`` fixtures)."""

from __future__ import annotations

import numpy as np
import pytest
from helpers import assert_obs_equal, assert_reward_close

from townlet_b200 import capi
from townlet_b200.params import DynamicsSpec, ObsSpec, RewardSpec
from townlet_b200.synth import random_actions, synth_batch

pytestmark = pytest.mark.gpu


def _shuffle_holders(batch, rng, collide: bool) -> None:
    """What the host's queue manager does between ticks: grants and releases; optionally objects sharing a tile."""
    slot = batch.ent_holder >= -1
    idx = np.nonzero(slot)[0]
    n_per_town = np.diff(batch.town_agent_off)
    town_of = np.repeat(np.arange(batch.n_towns), np.diff(batch.town_ent_off))
    holders = rng.integers(-1, n_per_town[town_of[idx]].clip(min=1))
    holders[rng.random(idx.size) < 0.5] = -1
    batch.ent_holder[idx] = holders
    if collide:  # move some slots onto the tile of an earlier slot of the same town (two objects on one tile)
        for k in idx[rng.random(idx.size) < 0.3]:
            earlier = idx[(idx < k) & (town_of[idx] == town_of[k])]
            if earlier.size:
                src = int(rng.choice(earlier))
                batch.ent[k] = (batch.ent[k] & ~np.uint32(0xFFFFF)) | (batch.ent[src] & np.uint32(0xFFFFF))


@pytest.mark.parametrize("mode", ["ent", "grid", "ent-generic", "ent-chunk8"])
@pytest.mark.parametrize("variant,grid,agents,towns", [("hybrid", 48, 8, 40), ("compact", 48, 10, 21), ("full", 200, 5, 6), ("hybrid", 128, 64, 3)])
def test_refresh_matches_oracle(variant, grid, agents, towns, mode, kernel_form) -> None:
    import torch

    from oracle.oracle import oracle_tick
    from townlet_b200.engine import DeviceTownBatch

    kernel_form.setenv("TL_MODE", mode.split("-")[0])
    if mode.endswith("generic"):
        kernel_form.setenv("TL_GENERIC", "1")
    if mode.endswith("chunk8"):
        kernel_form.setenv("TL_CHUNK", "8")
    obs, rew, dyn = ObsSpec(variant=variant), RewardSpec(fuse_faint=True), DynamicsSpec()
    batch = synth_batch(towns, grid, agents, obs, device_reservations=True)  # grid 200: 50 objects per town (two 32-slot steps)
    rng = np.random.default_rng(towns)
    _shuffle_holders(batch, rng, collide=True)
    ref_batch = batch.copy()
    dev = DeviceTownBatch(batch, obs, rew, dyn=dyn)
    for step, phases in enumerate((capi.TL_PHASE_ALL | capi.TL_PHASE_RESERVATIONS, capi.TL_PHASE_WORLD | capi.TL_PHASE_RESERVATIONS,
                                   capi.TL_PHASE_OBS | capi.TL_PHASE_RESERVATIONS)):
        n_ticks = 2
        words = random_actions(batch, n_ticks, seed=step) if phases & capi.TL_PHASE_ACTIONS else None
        out = dev.alloc_outputs(phases, n_ticks, summary=True, comps=True)
        actions = torch.from_numpy(words.view(np.int32)).cuda() if words is not None else None
        dev.tick(out, phases, n_ticks, actions=actions)
        torch.cuda.synchronize()
        ref = oracle_tick(ref_batch, obs, rew, phases, n_ticks, dyn=dyn, actions=words)
        got = {k: v.cpu().numpy() for k, v in vars(out).items() if v is not None}
        assert_obs_equal(got, ref, f"step {step}")
        assert_reward_close(got, ref, f"step {step}")
        state = dev.download_state(full=True)
        assert np.array_equal(state.ent, ref_batch.ent), f"step {step}: entity words"
        assert np.array_equal(state.flags, ref_batch.flags), f"step {step}: flags"
        # the host's queue manager grants and releases between the launches
        _shuffle_holders(ref_batch, rng, collide=False)
        dev.field("ent_holder").copy_(torch.from_numpy(ref_batch.ent_holder))
    slots = ref_batch.ent_holder >= -1
    kinds = (ref_batch.ent >> 20) & 3
    assert slots.any() and set(np.unique(kinds[slots])) <= {capi.TL_KIND_RESERVED, capi.TL_KIND_EMPTY}
    assert (kinds[slots] == capi.TL_KIND_RESERVED).any() and (kinds[slots] == capi.TL_KIND_EMPTY).any()


def test_refresh_through_the_host_path() -> None:
    from oracle.oracle import oracle_tick
    from townlet_b200.engine import HostEngine

    obs, rew = ObsSpec(), RewardSpec(fuse_faint=True)
    batch = synth_batch(16, 48, 8, obs, device_reservations=True)
    _shuffle_holders(batch, np.random.default_rng(1), collide=True)
    ref_batch = batch.copy()
    phases = capi.TL_PHASE_ALL | capi.TL_PHASE_RESERVATIONS
    eng = HostEngine(obs, rew)
    out = eng.alloc_outputs(batch, phases, 2)
    eng.submit(0, batch, out, phases, 2, wire="bits")
    eng.wait(0)
    ref = oracle_tick(ref_batch, obs, rew, phases, 2)
    assert_obs_equal(out, ref, "host path")
    assert np.array_equal(batch.ent, ref_batch.ent) and np.array_equal(batch.flags, ref_batch.flags)  # written back
    eng.close()


def test_the_phase_needs_the_holder_array() -> None:
    import ctypes as C

    import torch

    from townlet_b200.engine import DeviceTownBatch

    obs, rew = ObsSpec(), RewardSpec(fuse_faint=True)
    dev = DeviceTownBatch(synth_batch(4, 24, 4, obs), obs, rew)
    out = dev.alloc_outputs(capi.TL_PHASE_OBS, 1)
    dev._struct.ent_holder = None
    with pytest.raises(capi.TownletB200Error, match="ent_holder"):
        dev.tick(out, capi.TL_PHASE_OBS | capi.TL_PHASE_RESERVATIONS, 1)
    del C, torch
