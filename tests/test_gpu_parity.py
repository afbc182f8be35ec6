"""GPU: CUDA kernels (through the C ABI) against the reference goldens and the C oracle.

Bars (north_star): map tensors, masks, integer summaries and the float32
feature vector bit-exact; float64 need / reward state within 1e-6 relative.
"""

from __future__ import annotations

import numpy as np
import pytest
from fixture_io import fixture_paths, load_fixture
from helpers import REL_TOL, assert_kpi_close, assert_obs_equal, assert_refreshed_state, assert_reward_close

from townlet_b200 import capi
from townlet_b200.params import ObsSpec, RewardSpec
from townlet_b200.synth import synth_batch

pytestmark = pytest.mark.gpu
FIXTURES = fixture_paths()


def _device_run(batch, obs, rew, phases, n_ticks, tick_by_tick=False):
    import torch

    from townlet_b200.engine import DeviceTownBatch

    dev = DeviceTownBatch(batch, obs, rew)
    out = dev.alloc_outputs(phases, n_ticks, summary=True, comps=True)
    if tick_by_tick:
        for it in range(n_ticks):
            dev.tick(out, phases, 1, tick0=it)
    else:
        dev.tick(out, phases, n_ticks)
    torch.cuda.synchronize()
    dev.download_state(full=bool(phases & capi.TL_PHASE_RESERVATIONS))
    return {k: v.cpu().numpy() for k, v in vars(out).items() if v is not None}


def _host_run(batch, obs, rew, phases, n_ticks):
    from townlet_b200.engine import HostEngine

    eng = HostEngine(obs, rew)
    out = eng.alloc_outputs(batch, phases, n_ticks)
    eng.tick(batch, out, phases, n_ticks)
    eng.close()
    return out


def _check_state(batch, ref):
    assert_refreshed_state(batch, ref, "state")
    if "needs" in ref:
        np.testing.assert_allclose(batch.needs, ref["needs"][-1], rtol=REL_TOL, atol=1e-15)
    if "episode_total" in ref:
        np.testing.assert_allclose(batch.episode_total, ref["episode_total"][-1], rtol=REL_TOL, atol=1e-12)


@pytest.mark.parametrize("mode", ["ent", "grid"])
@pytest.mark.parametrize("path", FIXTURES, ids=[p.stem for p in FIXTURES])
def test_device_path_matches_reference_golden(path, mode, kernel_form) -> None:
    kernel_form.setenv("TL_MODE", mode)  # both kernel forms: entity list and dense shared-memory grid
    batch, obs, rew, meta, ref = load_fixture(path)
    got = _device_run(batch, obs, rew, meta["phases"], meta["n_ticks"])
    assert_obs_equal(got, ref, path.stem)
    assert_reward_close(got, ref, path.stem)
    assert_kpi_close(got, ref, batch, path.stem)
    _check_state(batch, ref)


@pytest.mark.parametrize("path", FIXTURES, ids=[p.stem for p in FIXTURES])
def test_host_path_matches_reference_golden(path) -> None:
    batch, obs, rew, meta, ref = load_fixture(path)
    got = _host_run(batch, obs, rew, meta["phases"], meta["n_ticks"])
    assert_obs_equal(got, ref, path.stem)
    assert_reward_close(got, ref, path.stem)
    assert_kpi_close(got, ref, batch, path.stem)
    _check_state(batch, ref)


SYNTH_CASES = [
    # name, variant, grid, agents, towns, ticks, extra ObsSpec kwargs
    ("hybrid_48x8", "hybrid", 48, 8, 96, 4, {}),
    ("hybrid_targets_48x8", "hybrid", 48, 8, 37, 2, {"include_targets": True, "personality_channels": True}),
    ("full_48x10", "full", 48, 10, 64, 3, {}),
    ("compact_48x10", "compact", 48, 10, 64, 3, {}),
    ("compact_types_48x10", "compact", 48, 10, 33, 2, {"object_channels": ("stove", "bed", "fridge"), "normalize_counts": False}),
    ("hybrid_128x64", "hybrid", 128, 64, 6, 3, {}),
    ("hybrid_small_window", "hybrid", 16, 3, 50, 2, {"local_window": 5, "top_friends": 1, "top_rivals": 3, "embed_dim": 4}),
    ("compact_big_window", "compact", 20, 5, 21, 2, {"compact_window": 13, "top_friends": 4, "top_rivals": 4, "embed_dim": 16}),
    ("hybrid_no_social", "hybrid", 16, 4, 11, 2, {"top_friends": 0, "top_rivals": 0}),
    ("single_agent_towns", "hybrid", 8, 1, 70, 2, {}),
]


@pytest.mark.parametrize("mode", ["ent", "grid", "ent-chunk8", "ent-generic"])
@pytest.mark.parametrize("case", SYNTH_CASES, ids=[c[0] for c in SYNTH_CASES])
def test_cuda_matches_oracle_on_synthetic_towns(case, mode, kernel_form) -> None:
    from oracle.oracle import oracle_tick

    kernel_form.setenv("TL_MODE", mode.split("-")[0])
    if mode == "ent-chunk8":
        kernel_form.setenv("TL_CHUNK", "8")
    if mode.endswith("-generic"):
        kernel_form.setenv("TL_GENERIC", "1")  # runtime-window kernels instead of the unrolled specialisations

    _, variant, grid, agents, towns, ticks, extra = case
    obs = ObsSpec(variant=variant, **extra)
    rew = RewardSpec(fuse_faint=True, personality_scaling=bool(extra.get("personality_channels")))
    batch = synth_batch(towns, grid, agents, obs, terminated_frac=0.05)
    batch.needs[0, ::7] = 0.045  # some agents faint during the run
    batch.episode_total[::5] = 49.95
    batch.episode_total[1::5] = -49.9
    b_dev, b_ora = batch.copy(), batch.copy()
    got = _device_run(b_dev, obs, rew, capi.TL_PHASE_ALL, ticks)
    ref = oracle_tick(b_ora, obs, rew, capi.TL_PHASE_ALL, ticks)
    assert_obs_equal(got, ref, case[0])
    assert_reward_close(got, ref, case[0])
    np.testing.assert_allclose(got["kpi"], ref["kpi"], rtol=1e-6, atol=1e-7)
    for name in ("needs", "episode_total"):
        np.testing.assert_allclose(getattr(b_dev, name), getattr(b_ora, name), rtol=REL_TOL, atol=1e-15)
    for name in ("term_block_tick", "episode_tick", "flags", "town_tick"):
        assert np.array_equal(getattr(b_dev, name), getattr(b_ora, name)), name


def test_one_launch_of_n_ticks_equals_n_launches() -> None:
    obs, rew = ObsSpec(), RewardSpec(fuse_faint=True)
    batch = synth_batch(40, 48, 8, obs)
    b1, b2 = batch.copy(), batch.copy()
    fused = _device_run(b1, obs, rew, capi.TL_PHASE_ALL, 5)
    stepped = _device_run(b2, obs, rew, capi.TL_PHASE_ALL, 5, tick_by_tick=True)
    for key in fused:
        assert np.array_equal(fused[key], stepped[key]), key
    assert np.array_equal(b1.arena[: b1.layout.mutable_bytes], b2.arena[: b2.layout.mutable_bytes])


def test_separate_entry_points_match_oracle() -> None:
    import torch

    from oracle.oracle import oracle_tick
    from townlet_b200.engine import DeviceTownBatch

    obs, rew = ObsSpec(variant="full"), RewardSpec()
    batch = synth_batch(24, 48, 10, obs, terminated_frac=0.1)
    ref_batch = batch.copy()
    dev = DeviceTownBatch(batch, obs, rew)
    out_r = dev.alloc_outputs(capi.TL_PHASE_DECAY | capi.TL_PHASE_REWARD, 1, comps=True)
    dev.decay_reward(out_r, capi.TL_PHASE_DECAY)
    dev.decay_reward(out_r, capi.TL_PHASE_REWARD)
    out_o = dev.alloc_outputs(capi.TL_PHASE_OBS, 1, summary=True)
    dev.obs_build(out_o)
    torch.cuda.synchronize()
    oracle_tick(ref_batch, obs, rew, capi.TL_PHASE_DECAY, 1)
    ref_r = oracle_tick(ref_batch, obs, rew, capi.TL_PHASE_REWARD, 1)
    ref_o = oracle_tick(ref_batch, obs, rew, capi.TL_PHASE_OBS, 1)
    got = {k: v.cpu().numpy() for k, v in vars(out_o).items() if v is not None}
    got.update({k: v.cpu().numpy() for k, v in vars(out_r).items() if v is not None})
    assert_obs_equal(got, ref_o, "separate obs")
    assert_reward_close(got, ref_r, "separate reward")


def test_obs_build_is_idempotent_and_stateless() -> None:
    """Size-independent property at BASELINE config 2 size (1,024 towns x 8 agents)."""
    import torch

    from townlet_b200.engine import DeviceTownBatch

    obs = ObsSpec()
    batch = synth_batch(1024, 48, 8, obs)
    dev = DeviceTownBatch(batch, obs, RewardSpec())
    before = dev.arena.clone()
    o1 = dev.alloc_outputs(capi.TL_PHASE_OBS, 1)
    o2 = dev.alloc_outputs(capi.TL_PHASE_OBS, 1)
    dev.obs_build(o1)
    dev.obs_build(o2)
    torch.cuda.synchronize()
    assert torch.equal(o1.map, o2.map) and torch.equal(o1.features, o2.features)
    assert torch.equal(before, dev.arena), "observation build must not modify state"
    m = o1.map[0]
    assert torch.all((m == 0) | (m == 1)), "hybrid channels are {0,1}"
    assert torch.all(m[:, 0].sum(dim=(1, 2)) == 1), "exactly one self cell"
    assert torch.all(m[:, 1, 5, 5] == 1), "observer counts on its own tile (map.py:109-110)"


@pytest.mark.parametrize("variant,towns,grid,agents", [("compact", 16384, 48, 10), ("full", 16384, 48, 10), ("hybrid", 4096, 128, 64)])
def test_full_size_configs_match_oracle(variant, towns, grid, agents) -> None:
    """BASELINE configs 3 and 4 at full size against the multi-threaded oracle (towns tiled from 256 distinct)."""
    import os

    from oracle.oracle import oracle_tick
    from townlet_b200.synth import tile_batch

    obs, rew = ObsSpec(variant=variant), RewardSpec(fuse_faint=True)
    base = synth_batch(256, grid, agents, obs)
    batch = tile_batch(base, towns // 256)
    b_dev, b_ora = batch.copy(), batch.copy()
    got = _device_run(b_dev, obs, rew, capi.TL_PHASE_ALL, 1)
    ref = oracle_tick(b_ora, obs, rew, capi.TL_PHASE_ALL, 1, n_threads=os.cpu_count() or 1, want_comps=True)
    assert_obs_equal(got, ref, f"{variant} full size")
    assert_reward_close(got, ref, f"{variant} full size")
    np.testing.assert_allclose(b_dev.needs, b_ora.needs, rtol=REL_TOL, atol=1e-15)


def _ragged_batch(obs, sizes, grid, extra_objects=0):
    """Towns of different sizes in one batch (ragged agent / entity ranges)."""
    from townlet_b200.synth import assemble, synth_town

    towns = []
    for t, n in enumerate(sizes):
        town = synth_town(1000 + t, grid, n, terminated_frac=0.1)
        if extra_objects:
            rng = np.random.default_rng(t)
            k = extra_objects
            town.object_ids += [f"extra_{i}" for i in range(k)]
            town.object_types += [("stove", "crate", "bed", "Fridge")[i % 4] for i in range(k)]
            town.object_positions = np.concatenate([town.object_positions, rng.integers(0, grid, size=(k, 2))])
        towns.append(town)
    return assemble(towns, obs)


@pytest.mark.parametrize("mode", ["ent", "grid", "ent-generic"])
@pytest.mark.parametrize(
    "variant,sizes,grid,extra",
    [
        ("hybrid", [1, 33, 7, 64, 2, 2, 2, 40, 5, 9, 70, 3], 40, 0),  # ragged, several passes per town
        ("compact", [3, 1, 1, 1, 17, 32, 31, 33, 8], 24, 30),  # stacked extra objects, unknown types
        ("full", [12, 5, 65], 64, 150),  # more than 128 entities per town: the uncached entity loop
    ],
)
def test_ragged_towns_match_oracle(variant, sizes, grid, extra, mode, kernel_form) -> None:
    from oracle.oracle import oracle_tick

    kernel_form.setenv("TL_MODE", mode.split("-")[0])
    if mode.endswith("-generic"):
        kernel_form.setenv("TL_GENERIC", "1")
    obs = ObsSpec(variant=variant, include_targets=True, object_channels=("stove", "fridge") if variant == "compact" else ())
    rew = RewardSpec(fuse_faint=True)
    batch = _ragged_batch(obs, sizes, grid, extra)
    batch.validate(len(obs.object_channels) if variant == "compact" else 0)
    b_dev, b_ora = batch.copy(), batch.copy()
    got = _device_run(b_dev, obs, rew, capi.TL_PHASE_ALL, 3)
    ref = oracle_tick(b_ora, obs, rew, capi.TL_PHASE_ALL, 3)
    assert_obs_equal(got, ref, f"ragged {variant}")
    assert_reward_close(got, ref, f"ragged {variant}")
    np.testing.assert_allclose(got["kpi"], ref["kpi"], rtol=1e-6, atol=1e-7)
    assert np.array_equal(b_dev.flags, b_ora.flags)


def test_batches_with_empty_towns_and_no_agents() -> None:
    """Empty inputs: towns without agents inside a batch, and a batch with no agents at all."""
    from oracle.oracle import oracle_tick
    from townlet_b200.soa import BatchLayout, TownBatch

    obs, rew = ObsSpec(), RewardSpec(fuse_faint=True)
    base = _ragged_batch(obs, [4, 6], 32)
    # splice two empty towns around the populated ones by rewriting the offset tables
    lay = BatchLayout.make(4, base.n_agents, base.n_entities, 6, 1)
    batch = TownBatch(lay, 32, 32)
    for name, view in base.views().items():
        if name not in ("town_agent_off", "town_ent_off", "town_tick"):
            getattr(batch, name)[...] = view
    a, e = base.town_agent_off, base.town_ent_off
    batch.town_agent_off[...] = [0, 0, a[1], a[1], a[2]]
    batch.town_ent_off[...] = [0, 0, e[1], e[1], e[2]]
    batch.town_tick[...] = [5, base.town_tick[0], 7, base.town_tick[1]]
    b_dev, b_ora = batch.copy(), batch.copy()
    got = _device_run(b_dev, obs, rew, capi.TL_PHASE_ALL, 2)
    ref = oracle_tick(b_ora, obs, rew, capi.TL_PHASE_ALL, 2)
    assert_obs_equal(got, ref, "empty towns")
    assert_reward_close(got, ref, "empty towns")
    np.testing.assert_allclose(got["kpi"], ref["kpi"], rtol=1e-6, atol=1e-7)
    assert np.array_equal(b_dev.town_tick, b_ora.town_tick)
    # nothing to do at all: zero towns
    empty = TownBatch(BatchLayout.make(0, 0, 0, 6, 1), 16, 16)
    out = _device_run(empty, obs, rew, capi.TL_PHASE_ALL, 1)
    assert out["map"].shape[1] == 0 and out["reward"].shape[1] == 0


@pytest.mark.parametrize("world_phases", [False, True], ids=["hot-path", "world-phases"])
def test_long_horizon_rollout_matches_oracle(world_phases) -> None:
    """1 500 ticks in one launch: needs run to zero, agents faint, episode totals reach the caps and the death
    window opens and closes — every reward, flag and KPI row of every tick against the oracle, observations of the
    last tick (the map / feature outputs overwrite one slot: tick stride 0) and the final state."""
    import torch

    from oracle.oracle import oracle_tick
    from townlet_b200.engine import DeviceTownBatch
    from townlet_b200.params import DynamicsSpec
    from townlet_b200.synth import random_actions

    ticks = 1500
    obs, rew = ObsSpec(), RewardSpec(fuse_faint=True)
    dyn = DynamicsSpec(rivalry_decay_per_tick=0.0002, tie_rivalry_decay=0.0003) if world_phases else None
    phases = capi.TL_PHASE_WORLD if world_phases else capi.TL_PHASE_ALL
    base = synth_batch(6, 48, 8, obs)
    actions = random_actions(base, ticks, seed=77) if world_phases else None
    cpu, gpu = base.copy(), base.copy()
    want = oracle_tick(cpu, obs, rew, phases, ticks, dyn=dyn, actions=actions, want_comps=False)

    dev = DeviceTownBatch(gpu, obs, rew, dyn=dyn)
    out = dev.alloc_outputs(phases, ticks, summary=False, comps=False)
    n = gpu.n_agents
    out.map = torch.empty((1, n, *obs.map_shape), dtype=torch.float32, device=dev.device).expand(ticks, -1, -1, -1, -1)
    out.features = torch.empty((1, n, obs.n_features), dtype=torch.float32, device=dev.device).expand(ticks, -1, -1)
    act = torch.from_numpy(actions.view(np.int32)).to(dev.device) if actions is not None else None
    dev.tick(out, phases, ticks, actions=act)
    torch.cuda.synchronize()
    dev.download_state()

    got_reward = out.reward.cpu().numpy()
    np.testing.assert_allclose(got_reward, want["reward"], rtol=REL_TOL, atol=1e-9)
    assert np.array_equal(out.terminated.cpu().numpy(), want["terminated"])
    np.testing.assert_allclose(out.kpi.cpu().numpy(), want["kpi"], rtol=1e-5, atol=1e-6)
    assert np.array_equal(out.map[0].cpu().numpy().view(np.uint32), want["map"][-1].view(np.uint32))
    assert np.array_equal(out.features[0].cpu().numpy().view(np.uint32), want["features"][-1].view(np.uint32))
    np.testing.assert_allclose(gpu.needs, cpu.needs, rtol=REL_TOL, atol=1e-15)
    np.testing.assert_allclose(gpu.episode_total, cpu.episode_total, rtol=REL_TOL, atol=1e-12)
    assert np.array_equal(gpu.flags, cpu.flags) and np.array_equal(gpu.term_block_tick, cpu.term_block_tick)
    assert want["terminated"].any() and not want["terminated"].all()  # the horizon crosses the faint threshold
    if world_phases:
        for field in ("pos", "ent", "riv_count", "riv_score", "tie_count", "tie_riv"):
            assert np.array_equal(getattr(gpu, field), getattr(cpu, field)), field


def test_ctx_reset_request_is_reported_once() -> None:
    """A pending ctx-reset request is consumed by the observation that reports it (service.py:224-229,
    grid.py:932-936): in a fused launch the flag shows on the first tick only (unless the agent is terminated)."""
    from oracle.oracle import oracle_tick

    obs, rew = ObsSpec(), RewardSpec(fuse_faint=True)
    base = synth_batch(20, 24, 6, obs)
    base.needs[0, :] = 0.9  # nobody faints during the sequence
    base.flags[:] &= ~np.uint32(capi.TL_F_TERMINATED | capi.TL_F_FAINTED)
    marked = np.arange(base.n_agents) % 3 == 0
    base.flags[marked] |= np.uint32(capi.TL_F_CTX_RESET)
    cpu, gpu = base.copy(), base.copy()
    want = oracle_tick(cpu, obs, rew, capi.TL_PHASE_ALL, 3)
    got = _device_run(gpu, obs, rew, capi.TL_PHASE_ALL, 3)
    assert_obs_equal(got, want, "ctx reset")
    ctx = got["features"][:, :, 16]  # TL_FEAT_CTX_RESET
    assert np.array_equal(ctx[0], marked.astype(np.float32)) and not ctx[1:].any()
    assert not (gpu.flags & capi.TL_F_CTX_RESET).any() and np.array_equal(gpu.flags, cpu.flags)


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("TL_FUZZ_SEEDS", "24"))))
def test_kernels_against_oracle_on_random_configurations(seed, kernel_form) -> None:
    """Fuzz: random variant, window, social layout, object channels, town sizes (ragged), phases and kernel form;
    every output and the whole final state against the oracle."""
    from oracle.oracle import oracle_tick
    from townlet_b200.params import DynamicsSpec
    from townlet_b200.synth import assemble, random_actions, synth_town

    rng = np.random.default_rng(1000 + seed)
    variant = ("hybrid", "full", "compact")[seed % 3]
    kw = dict(
        variant=variant,
        top_friends=int(rng.integers(0, 5)),
        top_rivals=int(rng.integers(0, 4)),
        embed_dim=int(rng.choice([4, 8, 8, 12, 16])),
        include_aggregates=bool(rng.random() < 0.7),
        include_targets=bool(rng.random() < 0.3),
        personality_channels=bool(rng.random() < 0.3),
    )
    if variant == "compact":
        kw["compact_window"] = int(rng.choice([3, 5, 7, 7, 9, 13]))
        kw["normalize_counts"] = bool(rng.random() < 0.5)
        kw["object_channels"] = tuple(rng.choice(["stove", "bed", "fridge", "shower"], size=int(rng.integers(0, 4)), replace=False))
    else:
        kw["local_window"] = int(rng.choice([5, 9, 11, 11, 15, 21]))
    obs, rew = ObsSpec(**kw), RewardSpec(fuse_faint=True, personality_scaling=bool(rng.random() < 0.3))
    grid = int(rng.choice([8, 24, 48, 96]))
    towns = [synth_town(int(rng.integers(0, 5000)), grid, int(rng.integers(1, 41)), terminated_frac=0.1) for _ in range(int(rng.integers(1, 30)))]
    base = assemble(towns, obs)
    ticks = int(rng.integers(1, 6))
    world_phases = bool(rng.random() < 0.5)
    dyn = DynamicsSpec(trust_decay=float(rng.choice([0.0, 0.05])), rivalry_decay_per_tick=float(rng.choice([0.005, 0.05]))) if world_phases else None
    phases = capi.TL_PHASE_WORLD if world_phases else capi.TL_PHASE_ALL
    actions = random_actions(base, ticks, seed=seed) if world_phases else None
    mode = ("ent", "grid", "ent")[int(rng.integers(0, 3))]
    kernel_form.setenv("TL_MODE", mode)
    if rng.random() < 0.3:
        kernel_form.setenv("TL_CHUNK", str(int(rng.choice([8, 16, 24]))))
    if rng.random() < 0.3:
        kernel_form.setenv("TL_GENERIC", "1")

    import torch

    from townlet_b200.engine import DeviceTownBatch

    cpu, gpu = base.copy(), base.copy()
    want = oracle_tick(cpu, obs, rew, phases, ticks, dyn=dyn, actions=actions)
    dev = DeviceTownBatch(gpu, obs, rew, dyn=dyn)
    out = dev.alloc_outputs(phases, ticks, summary=True, comps=True)
    act = torch.from_numpy(actions.view(np.int32)).to(dev.device) if actions is not None else None
    dev.tick(out, phases, ticks, actions=act)
    torch.cuda.synchronize()
    dev.download_state()
    got = {k: v.cpu().numpy() for k, v in vars(out).items() if v is not None}
    what = f"seed {seed}: {kw}, grid {grid}, {len(towns)} towns, {base.n_agents} agents, {ticks} ticks, mode {mode}, world {world_phases}"
    assert_obs_equal(got, want, what)
    assert_reward_close(got, want, what)
    np.testing.assert_allclose(gpu.needs, cpu.needs, rtol=REL_TOL, atol=1e-15, err_msg=what)
    for field in ("flags", "pos", "ent", "episode_tick", "term_block_tick", "tie_count", "tie_trust", "riv_count", "riv_score"):
        assert np.array_equal(getattr(gpu, field), getattr(cpu, field)), f"{what}: {field}"


def test_episode_tick_wraps_like_python_modulo() -> None:
    """(episode_tick + 1) % ticks_per_day with Python semantics (world/core/context.py:283-289), also for inputs
    outside [0, ticks_per_day)."""
    from oracle.oracle import oracle_tick

    obs, rew = ObsSpec(), RewardSpec(fuse_faint=True)
    base = synth_batch(4, 16, 6, obs)
    base.episode_tick[:] = np.array([-7, -1, 0, 1438, 1439, 1440, 5000, -1441] * 3, dtype=np.int32)[: base.n_agents]
    cpu, gpu = base.copy(), base.copy()
    want = oracle_tick(cpu, obs, rew, capi.TL_PHASE_ALL, 2)
    got = _device_run(gpu, obs, rew, capi.TL_PHASE_ALL, 2)
    assert_obs_equal(got, want, "episode tick")
    expect = (((base.episode_tick.astype(np.int64) + 1) % 1440) + 1) % 1440
    assert np.array_equal(cpu.episode_tick, expect) and np.array_equal(gpu.episode_tick, expect)
