"""World dynamics around the hot path: actions and ledger decay inside the tick.

The fixtures under tests/golden/world/ were produced by running the reference per tick in
``WorldContext.tick`` order (world/core/context.py:229-289): ``apply_actions`` + ``process_actions``
(world/actions.py:44, world/systems/affordances.py:105), need decay, ``RelationshipService.decay``
(world/agents/relationships_service.py:165), lifecycle, reward, observe.  CPU tests pin the C oracle to
them; GPU tests compare the kernels (both forms) with the fixtures and with the oracle on larger seeded
batches.  Maps, features, summaries, positions, entity words and ledger rows are bit-exact.
"""

from __future__ import annotations

import numpy as np
import pytest
from fixture_io import load_fixture, world_fixture_paths
from helpers import REL_TOL, assert_obs_equal, assert_reward_close

from oracle.oracle import oracle_tick
from townlet_b200 import capi
from townlet_b200.actions import encode_actions
from townlet_b200.params import DynamicsSpec, ObsSpec, RewardSpec, pack_actions
from townlet_b200.synth import random_actions, synth_batch

FIXTURES = world_fixture_paths()
LEDGER_FIELDS = ("tie_count", "tie_trust", "tie_fam", "tie_riv", "tie_embed", "riv_count", "riv_score", "riv_embed")
ACTION_FIELDS = ("pos", "ent", "last_action_duration", "last_action_hash")


def _check_final_state(batch, ref, what: str) -> None:
    for name in LEDGER_FIELDS + ACTION_FIELDS:
        got, want = np.asarray(getattr(batch, name)), ref[f"final_{name}"]
        assert got.shape == want.shape, (what, name)
        same = got.view(np.uint8) == want.view(np.uint8) if got.dtype.kind == "f" else got == want
        assert np.all(same), f"{what}: {name} differs at {np.argwhere(~np.asarray(same))[:4].tolist()}"
    got_success = np.asarray(batch.flags) & capi.TL_F_LAST_SUCCESS
    assert np.array_equal(got_success, ref["final_last_success"]), f"{what}: last_action_success"
    np.testing.assert_allclose(batch.needs, ref["needs"][-1], rtol=REL_TOL, atol=1e-15)
    np.testing.assert_allclose(batch.episode_total, ref["episode_total"][-1], rtol=REL_TOL, atol=1e-12)


# ---------------------------------------------------------------------------
# CPU: oracle against the reference, host logic
# ---------------------------------------------------------------------------
def test_world_fixtures_present() -> None:
    assert {p.stem for p in FIXTURES} >= {"world_hybrid_48x8", "world_compact_24x10", "world_full_tie_decay", "world_hybrid_128x64"}


@pytest.mark.parametrize("path", FIXTURES, ids=[p.stem for p in FIXTURES])
def test_oracle_matches_reference_world(path) -> None:
    batch, obs, rew, meta, ref = load_fixture(path)
    assert meta["phases"] == capi.TL_PHASE_WORLD
    out = oracle_tick(batch, obs, rew, meta["phases"], meta["n_ticks"], dyn=meta["dyn_spec"], actions=meta["actions"])
    assert_obs_equal(out, ref, path.stem)
    assert_reward_close(out, ref, path.stem)
    _check_final_state(batch, ref, path.stem)


def test_fixture_sequences_evict_edges() -> None:
    """The sequences are only a test of compaction if edges actually leave the ledgers."""
    for path in FIXTURES:
        batch, _, _, _, ref = load_fixture(path)
        assert ref["final_riv_count"].sum() < batch.riv_count.sum(), path.stem
        assert ref["final_tie_count"].sum() < batch.tie_count.sum(), path.stem
        assert np.any(ref["final_pos"] != batch.pos), path.stem


def test_encode_actions_follows_process_actions() -> None:
    ids = ["a", "b", "c", "d", "e", "f"]
    actions = {
        "a": {"kind": "move", "position": (7, 9), "duration": 3},
        "b": {"kind": "move"},  # no position: recorded, unsuccessful (affordances.py:157)
        "c": {"kind": "wait", "duration": 12},
        "d": {"kind": "request", "object": "stove_1"},
        "e": "not a mapping",  # skipped (affordances.py:129-130)
    }
    words = encode_actions(ids, actions, origin=(2, 4), success={"d": True})
    kind = words & 15
    assert kind.tolist() == [1, 1, 3, 4, 0, 0]
    assert (words[0] >> 5) & 1 == 1 and (words[0] >> 6) & 1023 == 5 and (words[0] >> 16) & 1023 == 5 and words[0] >> 26 == 3
    assert (words[1] >> 5) & 1 == 0
    assert words[2] >> 26 == 12 and (words[2] >> 4) & 1 == 0
    assert (words[3] >> 4) & 1 == 1
    with pytest.raises(ValueError):
        encode_actions(ids, {"a": {"kind": "teleport"}})
    with pytest.raises(ValueError):
        pack_actions(1, 0, 1, 5, 5, 64)
    with pytest.raises(ValueError):
        pack_actions(1, 0, 1, 2000, 5, 1)


def test_oracle_no_action_and_no_decay_is_identity() -> None:
    obs, rew = ObsSpec(), RewardSpec(fuse_faint=True)
    dyn = DynamicsSpec(tie_rivalry_decay=0.0, rivalry_decay_per_tick=0.0, rivalry_eviction_threshold=-1.0)
    base = synth_batch(5, 24, 6, obs)
    a, b = base.copy(), base.copy()
    o1 = oracle_tick(a, obs, rew, capi.TL_PHASE_ALL, 3)
    o2 = oracle_tick(b, obs, rew, capi.TL_PHASE_WORLD, 3, dyn=dyn, actions=np.zeros(base.n_agents, np.uint32))
    for key in o1:
        assert np.array_equal(o1[key], o2[key], equal_nan=True), key
    assert np.array_equal(a.arena, b.arena)


# ---------------------------------------------------------------------------
# GPU: kernels against the reference fixtures and the oracle
# ---------------------------------------------------------------------------
def _device_run(batch, obs, rew, dyn, phases, n_ticks, actions, tick_by_tick=False):
    import torch

    from townlet_b200.engine import DeviceTownBatch

    dev = DeviceTownBatch(batch, obs, rew, dyn=dyn)
    out = dev.alloc_outputs(phases, n_ticks, summary=True, comps=True)
    act = torch.from_numpy(np.ascontiguousarray(actions).view(np.int32)).to(dev.device) if actions is not None else None
    if tick_by_tick:
        for it in range(n_ticks):
            dev.tick(out, phases, 1, tick0=it, actions=None if act is None else (act[it].contiguous() if act.dim() == 2 else act))
    else:
        dev.tick(out, phases, n_ticks, actions=act)
    torch.cuda.synchronize()
    dev.download_state()
    return {k: v.cpu().numpy() for k, v in vars(out).items() if v is not None}


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["ent", "grid", "ent-generic"])
@pytest.mark.parametrize("path", FIXTURES, ids=[p.stem for p in FIXTURES])
def test_device_matches_reference_world(path, mode, kernel_form) -> None:
    kernel_form.setenv("TL_MODE", mode.split("-")[0])
    if mode.endswith("generic"):
        kernel_form.setenv("TL_GENERIC", "1")
    batch, obs, rew, meta, ref = load_fixture(path)
    got = _device_run(batch, obs, rew, meta["dyn_spec"], meta["phases"], meta["n_ticks"], meta["actions"])
    assert_obs_equal(got, ref, path.stem)
    assert_reward_close(got, ref, path.stem)
    _check_final_state(batch, ref, path.stem)


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=[p.stem for p in FIXTURES])
def test_host_path_matches_reference_world(path) -> None:
    """The same fixtures through tl_host_tick_world: HOST buffers in, HOST buffers and mutated state out."""
    from townlet_b200.engine import HostEngine

    batch, obs, rew, meta, ref = load_fixture(path)
    eng = HostEngine(obs, rew)
    out = eng.alloc_outputs(batch, meta["phases"], meta["n_ticks"])
    eng.tick(batch, out, meta["phases"], meta["n_ticks"], dyn=meta["dyn_spec"], actions=meta["actions"])
    eng.close()
    assert_obs_equal(out, ref, path.stem)
    assert_reward_close(out, ref, path.stem)
    _check_final_state(batch, ref, path.stem)


WORLD_SYNTH = [
    # name, variant, grid, agents, towns, ticks, ObsSpec kwargs, DynamicsSpec kwargs
    ("hybrid_48x8", "hybrid", 48, 8, 96, 12, {}, {}),
    ("full_48x10", "full", 48, 10, 40, 6, {}, {"trust_decay": 0.05, "familiarity_decay": 0.02}),
    ("compact_48x10", "compact", 48, 10, 40, 6, {"object_channels": ("stove", "bed")}, {"rivalry_decay_per_tick": 0.03}),
    ("hybrid_128x64", "hybrid", 128, 64, 5, 4, {}, {"tie_rivalry_decay": 0.08}),
    ("hybrid_no_social", "hybrid", 16, 4, 11, 5, {"top_friends": 0, "top_rivals": 0}, {"rivalry_decay_per_tick": 0.02}),
]


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["ent", "grid", "ent-chunk8"])
@pytest.mark.parametrize("case", WORLD_SYNTH, ids=[c[0] for c in WORLD_SYNTH])
def test_device_matches_oracle_world(case, mode, kernel_form) -> None:
    kernel_form.setenv("TL_MODE", mode.split("-")[0])
    if mode.endswith("chunk8"):
        kernel_form.setenv("TL_CHUNK", "8")
    name, variant, grid, agents, towns, ticks, okw, dkw = case
    obs, rew, dyn = ObsSpec(variant=variant, **okw), RewardSpec(fuse_faint=True), DynamicsSpec(**dkw)
    base = synth_batch(towns, grid, agents, obs)
    actions = random_actions(base, ticks, seed=len(name) + towns)
    cpu, gpu = base.copy(), base.copy()
    want = oracle_tick(cpu, obs, rew, capi.TL_PHASE_WORLD, ticks, dyn=dyn, actions=actions)
    got = _device_run(gpu, obs, rew, dyn, capi.TL_PHASE_WORLD, ticks, actions)
    assert_obs_equal(got, want, name)
    assert_reward_close(got, want, name)
    assert np.array_equal(got["kpi"].view(np.uint32), want["kpi"].view(np.uint32)) or np.allclose(got["kpi"], want["kpi"], rtol=1e-6)
    for field in LEDGER_FIELDS + ACTION_FIELDS + ("flags",):
        assert np.array_equal(getattr(gpu, field), getattr(cpu, field)), f"{name}: {field}"
    assert cpu.riv_count.sum() < base.riv_count.sum()  # the sequence evicted something


@pytest.mark.gpu
def test_world_phases_tick_by_tick_equals_fused() -> None:
    obs, rew, dyn = ObsSpec(), RewardSpec(fuse_faint=True), DynamicsSpec(trust_decay=0.03)
    base = synth_batch(64, 48, 8, obs)
    actions = random_actions(base, 6, seed=5)
    fused, stepped = base.copy(), base.copy()
    a = _device_run(fused, obs, rew, dyn, capi.TL_PHASE_WORLD, 6, actions)
    b = _device_run(stepped, obs, rew, dyn, capi.TL_PHASE_WORLD, 6, actions, tick_by_tick=True)
    for key in a:
        assert np.array_equal(a[key], b[key], equal_nan=True), key
    assert np.array_equal(fused.arena, stepped.arena)


@pytest.mark.gpu
def test_world_phases_without_observation() -> None:
    """Ledger decay and actions also run when no observation is requested (state-only rollouts)."""
    obs, rew, dyn = ObsSpec(), RewardSpec(fuse_faint=True), DynamicsSpec()
    base = synth_batch(33, 24, 6, obs)
    actions = random_actions(base, 4, seed=9)
    phases = capi.TL_PHASE_WORLD & ~capi.TL_PHASE_OBS
    cpu, gpu = base.copy(), base.copy()
    want = oracle_tick(cpu, obs, rew, phases, 4, dyn=dyn, actions=actions)
    got = _device_run(gpu, obs, rew, dyn, phases, 4, actions)
    assert_reward_close(got, want, "state-only")
    for field in LEDGER_FIELDS + ACTION_FIELDS + ("flags",):
        assert np.array_equal(getattr(gpu, field), getattr(cpu, field)), field


@pytest.mark.gpu
def test_world_phases_need_dynamics_params() -> None:
    import torch

    from townlet_b200.engine import DeviceTownBatch

    obs, rew = ObsSpec(), RewardSpec()
    base = synth_batch(2, 16, 4, obs)
    dev = DeviceTownBatch(base, obs, rew)  # no DynamicsSpec
    out = dev.alloc_outputs(capi.TL_PHASE_ALL, 1)
    with pytest.raises(capi.TownletB200Error):
        dev.tick(out, capi.TL_PHASE_WORLD, 1)
    dev2 = DeviceTownBatch(base, obs, rew, dyn=DynamicsSpec())
    with pytest.raises(capi.TownletB200Error):
        dev2.tick(out, capi.TL_PHASE_WORLD, 1)  # TL_PHASE_ACTIONS without action words
    with pytest.raises(capi.TownletB200Error):
        dev2.tick(out, capi.TL_PHASE_WORLD, 1, actions=torch.zeros(3, dtype=torch.int32, device="cuda"))
