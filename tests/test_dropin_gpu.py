"""GPU: the drop-in classes end to end (fake world -> ingest -> tl_host_tick -> dicts) against the oracle."""

from __future__ import annotations

import numpy as np
import pytest
from fake_world import FakeWorld, fake_config

from townlet_b200 import capi
from townlet_b200.params import ObsSpec, RewardSpec
from townlet_b200.synth import assemble, synth_town

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("variant", ["hybrid", "full", "compact"])
def test_observation_service_on_fake_world(variant) -> None:
    from oracle.oracle import oracle_tick
    from townlet_b200.observation_service import B200ObservationService

    config = fake_config(variant)
    town = synth_town(3, 48, 10)
    world = FakeWorld(town, config)
    world.request_ctx_reset(town.agent_ids[2])
    terminated = {town.agent_ids[1]: True}
    service = B200ObservationService(config=config)
    batch_out = service.build_batch(world, terminated)
    assert list(batch_out) == town.agent_ids
    assert world.ctx_reset_consumed == 1
    assert world.embedding_allocator.released == [town.agent_ids[1]]

    spec = ObsSpec.from_config(config)
    town.terminated[:] = False
    ref_batch = assemble([town], spec)
    ref_batch.flags[1] |= capi.TL_F_TERMINATED
    ref_batch.flags[2] |= capi.TL_F_CTX_RESET
    ref = oracle_tick(ref_batch, spec, RewardSpec(), capi.TL_PHASE_OBS, 1)
    for i, aid in enumerate(town.agent_ids):
        entry = batch_out[aid]
        assert entry["map"].shape == spec.map_shape and entry["map"].dtype == np.float32
        assert np.array_equal(entry["map"], ref["map"][0, i]), aid
        assert np.array_equal(entry["features"].view(np.uint32), ref["features"][0, i].view(np.uint32)), aid
        meta = entry["metadata"]
        assert meta["variant"] == variant and meta["embedding_slot"] == i
        assert meta["local_summary"]["agent_count"] == float(ref["summary"][0, i, 0])
        assert meta["social_context"]["filled_slots"] == int(ref["summary"][0, i, 4])
    with pytest.raises(KeyError):
        service.build_single(world, "nobody")


def test_reward_engine_on_fake_world() -> None:
    from oracle.oracle import oracle_tick
    from townlet_b200.reward_engine import B200RewardEngine

    config = fake_config("hybrid")
    town = synth_town(5, 48, 8)
    world = FakeWorld(town, config)
    a0, a1, a2 = town.agent_ids[:3]
    world.chat_events = [{"event": "chat_success", "speaker": a0, "listener": a1, "quality": 0.5}, {"event": "chat_failure", "speaker": a2, "listener": a0}]
    world.avoidance_events = [{"agent": a1, "object_id": "o", "reason": "queue_rival"}]
    engine = B200RewardEngine(config)
    out = engine.compute(world, {a2: True}, {a2: "faint"})
    assert list(out) == town.agent_ids
    assert world.events_consumed == 1 and not world.chat_events
    assert out[a0].social_bonus > 0 and out[a0].social_penalty < 0 and out[a1].social_avoidance > 0
    assert out[a2].terminal_penalty == -1.0 and engine._termination_block == {a2: world.tick}
    # same numbers from the oracle on the equivalent batch
    spec = RewardSpec.from_config(config)
    from townlet_b200.reward_events import reduce_social_events

    world2 = FakeWorld(town, config)
    sums = reduce_social_events(
        world2,
        [{"event": "chat_success", "speaker": a0, "listener": a1, "quality": 0.5}, {"event": "chat_failure", "speaker": a2, "listener": a0}],
        [{"agent": a1, "object_id": "o", "reason": "queue_rival"}],
        spec,
    )
    town.terminated[:] = False
    batch = assemble([town])
    batch.flags[2] |= capi.TL_F_TERMINATED | (1 << capi.TL_F_REASON_SHIFT)
    for k in range(3):
        for i, aid in enumerate(town.agent_ids):
            batch.social_in[k, i] = sums[k].get(aid, 0.0)
    ref = oracle_tick(batch, ObsSpec(), spec, capi.TL_PHASE_REWARD, 1)
    for i, aid in enumerate(town.agent_ids):
        assert out[aid].total == pytest.approx(ref["comps"][0, i, 0], rel=1e-6, abs=1e-12)
        assert out[aid].needs_penalty == pytest.approx(ref["comps"][0, i, 2], rel=1e-6, abs=1e-12)
    # second tick: death window guard for the terminated agent, episode totals carried on the host
    world.tick += 1
    out2 = engine.compute(world, {}, {})
    assert out2[a2].total <= 0.0
    assert engine._episode_totals[a0] == pytest.approx(out[a0].total + out2[a0].total, rel=1e-9)


def test_need_decay_hook_on_fake_world() -> None:
    from townlet_b200.reward_engine import install_need_decay

    config = fake_config("hybrid")
    town = synth_town(9, 48, 8)
    world = FakeWorld(town, config)
    before = {aid: dict(s.needs) for aid, s in world.agents.items()}
    hook = install_need_decay(world)
    hook()
    for aid, snap in world.agents.items():
        assert snap.needs["hunger"] == max(0.0, before[aid]["hunger"] - 0.01)
        assert snap.needs["hygiene"] == max(0.0, before[aid]["hygiene"] - 0.005)
        assert snap.needs["energy"] == max(0.0, before[aid]["energy"] - 0.008)


def test_parallel_env_steps() -> None:
    import torch

    from oracle.oracle import oracle_tick
    from townlet_b200.env import TownletParallelEnv
    from townlet_b200.synth import synth_batch

    obs, rew = ObsSpec(variant="compact"), RewardSpec(fuse_faint=True)
    batch = synth_batch(12, 48, 10, obs)
    env = TownletParallelEnv(batch, obs, rew, max_ticks=3)
    observations, infos = env.reset()
    assert observations["map"].shape == (batch.n_agents, *obs.map_shape) and infos == {}
    ref_batch = batch.copy()
    for step in range(3):
        observations, rewards, terminations, truncations, infos = env.step(None)
        ref = oracle_tick(ref_batch, obs, rew, capi.TL_PHASE_ALL, 1)
        torch.cuda.synchronize()
        assert np.array_equal(observations["map"].cpu().numpy(), ref["map"][0])
        np.testing.assert_allclose(rewards.cpu().numpy(), ref["reward"][0], rtol=1e-6, atol=1e-9)
        assert np.array_equal(terminations.cpu().numpy(), ref["terminated"][0].astype(bool))
        assert bool(truncations.all()) == (step == 2)
        assert infos["kpi"].shape == (batch.n_towns, capi.TL_N_KPI)


def test_parallel_env_applies_action_words_on_the_device() -> None:
    """With a DynamicsSpec the env moves agents and decays the ledgers inside the tick."""
    import torch

    from oracle.oracle import oracle_tick
    from townlet_b200.env import TownletParallelEnv
    from townlet_b200.params import DynamicsSpec, pack_actions
    from townlet_b200.synth import synth_batch

    obs, rew, dyn = ObsSpec(), RewardSpec(fuse_faint=True), DynamicsSpec()
    batch = synth_batch(9, 32, 6, obs)
    env = TownletParallelEnv(batch, obs, rew, dyn=dyn)
    env.reset()
    ref_batch = batch.copy()
    rng = np.random.default_rng(3)
    for step in range(4):
        n = batch.n_agents
        words = pack_actions(rng.choice([0, 1, 3], size=n), 0, 1, rng.integers(0, 32, n), rng.integers(0, 32, n), 1)
        act = torch.from_numpy(words.view(np.int32)).cuda() if step else None  # first step: nobody acts
        observations, rewards, _, _, _ = env.step(act)
        ref = oracle_tick(
            ref_batch, obs, rew, capi.TL_PHASE_WORLD, 1, dyn=dyn, actions=words if step else np.zeros(n, np.uint32)
        )
        torch.cuda.synchronize()
        assert np.array_equal(observations["map"].cpu().numpy(), ref["map"][0])
        assert np.array_equal(observations["features"].cpu().numpy().view(np.uint32), ref["features"][0].view(np.uint32))
        np.testing.assert_allclose(rewards.cpu().numpy(), ref["reward"][0], rtol=1e-6, atol=1e-9)
    state = env.state().download_state()
    assert np.array_equal(state.pos, ref_batch.pos) and np.array_equal(state.riv_score, ref_batch.riv_score)


def test_env_step_loop_with_employment_and_reservation_refresh() -> None:
    """The env's tick = shift state machine (tl_employment_step) for the coming tick, then actions, reservation refresh, ledger decay,
    need decay, reward and observation in one launch — against the oracle doing the same on the host, tick by tick."""
    import torch

    from oracle.oracle import oracle_employment_step, oracle_tick
    from townlet_b200.employment import EmploymentArrays, EmploymentSpec
    from townlet_b200.env import TownletParallelEnv
    from townlet_b200.params import DynamicsSpec
    from townlet_b200.synth import random_actions, synth_batch

    obs, rew, dyn = ObsSpec(), RewardSpec(fuse_faint=True), DynamicsSpec()
    rng = np.random.default_rng(4)
    batch = synth_batch(30, 16, 7, obs, device_reservations=True)
    batch.town_tick[:] = rng.integers(0, 50, batch.n_towns)
    spec = EmploymentSpec(
        job_ids=("a", "b", "c"), job_start=(10, 25, 40), job_end=(30, 45, 40), job_wage=(0.02, 0.03, 0.01), job_lateness_penalty=(0.1, 0.2, 0.0),
        job_location=((3, 3), None, (7, 2)), arrival_buffer=4, ticks_per_day=60, grace_ticks=2, absent_cutoff=9, absence_slack=4,
        attendance_window=3, late_tick_penalty=0.005, absence_penalty=0.2,
    )
    arrays = EmploymentArrays(batch.n_agents)
    arrays.job[:] = rng.integers(-1, 3, batch.n_agents)
    env = TownletParallelEnv(batch, obs, rew, dyn=dyn, employment=(spec, arrays))
    ref_batch, ref_arrays = batch.copy(), arrays.copy()
    env.reset()
    phases = capi.TL_PHASE_WORLD | capi.TL_PHASE_RESERVATIONS
    seen_events = 0
    for t in range(40):
        words = random_actions(ref_batch, 1, seed=100 + t)
        observations, reward, terminations, _, infos = env.step(torch.from_numpy(words[0].view(np.int32)).cuda())
        torch.cuda.synchronize()
        want_events = oracle_employment_step(ref_batch, spec, ref_arrays, 1, 1)
        ref = oracle_tick(ref_batch, obs, rew, phases, 1, dyn=dyn, actions=words)
        assert np.array_equal(infos["employment_events"].cpu().numpy().view(np.uint32), want_events[0]), t
        assert np.array_equal(observations["map"].cpu().numpy(), ref["map"][0]), t
        assert np.array_equal(observations["features"].cpu().numpy().view(np.uint32), ref["features"][0].view(np.uint32)), t
        np.testing.assert_allclose(reward.cpu().numpy(), ref["reward"][0], rtol=1e-6, atol=1e-9)
        assert np.array_equal(terminations.cpu().numpy(), ref["terminated"][0].astype(bool)), t
        seen_events |= int(np.bitwise_or.reduce(want_events, axis=None))
    state = env.state().download_state(full=True)
    assert np.array_equal(state.arena, ref_batch.arena)
    got = env.employment_state().download()
    for name, _, _ in capi.EMPLOYMENT_STATE_FIELDS:
        assert np.array_equal(got.arrays[name], ref_arrays.arrays[name]), name
    assert seen_events != 0  # shifts started, ran late or were missed along the way
    # reset restores towns and employment context
    env.reset()
    assert np.array_equal(env.employment_state().download().arrays["state"], arrays.arrays["state"])
