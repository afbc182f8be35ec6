"""CPU: host logic of the drop-in classes against the REAL reference classes.

Runs only where the reference is present (/root/reference in the build
container).  The CUDA library is replaced by the C oracle through the
test-only ``OracleHostEngine``; what is under test is everything around it:
ingest, metadata dictionaries, destructive reads, slot allocation, reward
bookkeeping — compared with ``WorldObservationService`` / ``RewardEngine`` /
``WorldState._apply_need_decay`` of the reference on the same worlds.
"""

from __future__ import annotations

import os
import shutil
import sys
from pathlib import Path

import numpy as np
import pytest

REF = Path(os.environ.get("TOWNLET_REFERENCE", "/root/reference"))
pytestmark = pytest.mark.skipif(not (REF / "src" / "townlet").exists(), reason="reference checkout not available")


@pytest.fixture(scope="module")
def ref(tmp_path_factory):
    work = tmp_path_factory.mktemp("ref")
    shutil.copytree(REF, work / "ref", symlinks=True)
    os.system(f"chmod -R u+w {work / 'ref'}")
    cwd = os.getcwd()
    os.chdir(work / "ref")
    sys.path.insert(0, str(work / "ref" / "src"))
    import townlet_b200.observation_service as obs_mod
    import townlet_b200.reward_engine as rew_mod
    from oracle_engine import OracleHostEngine

    saved = (obs_mod.HostEngine, rew_mod.HostEngine)
    obs_mod.HostEngine = OracleHostEngine
    rew_mod.HostEngine = OracleHostEngine
    yield work
    obs_mod.HostEngine, rew_mod.HostEngine = saved
    os.chdir(cwd)
    sys.path.remove(str(work / "ref" / "src"))


def _config(variant="hybrid", **overrides):
    from townlet.config import load_config

    config = load_config(Path("configs/examples/poc_hybrid.yaml"))
    data = config.model_dump()
    data["features"]["systems"]["observations"] = variant
    for dotted, value in overrides.items():
        node = data
        keys = dotted.split("__")
        for key in keys[:-1]:
            node = node[key]
        node[keys[-1]] = value
    return type(config).model_validate(data)


def _populate(world):
    from townlet.world.grid import AgentSnapshot

    world.tick = 321
    for oid, otype, pos in (("stove_a", "stove", (2, 1)), ("fridge_a", "fridge", (2, 1)), ("bed_a", "bed", (-3, 2)), ("shower_a", "shower", (0, -4))):
        world.register_object(object_id=oid, object_type=otype, position=pos)
    spots = {"a": (0, 0), "b": (0, 0), "c": (2, 0), "d": (2, 1), "e": (-5, -5)}
    for k, (aid, pos) in enumerate(spots.items()):
        world.agents[aid] = AgentSnapshot(
            agent_id=aid,
            position=pos,
            needs={"hunger": 0.15 * (k + 1), "hygiene": 0.9 - 0.1 * k, "energy": 0.5},
            wallet=float(k),
            last_action_id=("", "move", "wait", "chat", "noop")[k],
            last_action_duration=k,
            episode_tick=100 * k,
        )
    world.rebuild_spatial_index()
    world._active_reservations["stove_a"] = "d"
    world._spatial_index.set_reservation((2, 1), True)
    world.update_relationship("a", "b", trust=0.4, familiarity=0.2)
    world.update_relationship("a", "c", trust=0.1, familiarity=0.5, rivalry=0.3)
    world.register_rivalry_conflict("c", "d")
    world.request_ctx_reset("b")


def _assert_same_batch(got, want):
    assert list(got) == list(want)
    for aid in want:
        assert got[aid]["map"].dtype == np.float32 and got[aid]["features"].dtype == np.float32
        np.testing.assert_array_equal(got[aid]["map"], want[aid]["map"], err_msg=aid)
        np.testing.assert_array_equal(got[aid]["features"].view(np.uint32), want[aid]["features"].view(np.uint32), err_msg=aid)
        assert got[aid]["metadata"] == want[aid]["metadata"], aid


@pytest.mark.parametrize(
    "variant,overrides",
    [
        ("hybrid", {}),
        ("hybrid", {"observations_config__hybrid__include_targets": True, "features__observations__personality_channels": True}),
        ("full", {}),
        ("compact", {}),
        ("compact", {"observations_config__compact__object_channels": ["stove", "Fridge "], "observations_config__compact__normalize_counts": False}),
    ],
)
def test_observation_service_matches_reference(ref, variant, overrides):
    from townlet.world.grid import WorldState
    from townlet.world.observations.interfaces import ObservationServiceProtocol
    from townlet.world.observations.service import WorldObservationService

    from townlet_b200.observation_service import B200ObservationService

    config = _config(variant, **overrides)
    w_ref, w_new = WorldState.from_config(config), WorldState.from_config(config)
    _populate(w_ref)
    _populate(w_new)
    service_ref = WorldObservationService(config=config)
    service_new = B200ObservationService(config=config)
    assert isinstance(service_new, ObservationServiceProtocol)
    assert service_new.variant == service_ref.variant
    assert service_new.feature_names == service_ref.feature_names
    assert service_new.social_vector_length == service_ref.social_vector_length
    terminated = {"c": True}
    _assert_same_batch(service_new.build_batch(w_new, terminated), service_ref.build_batch(w_ref, terminated))
    # side effects: ctx resets consumed once, slot of the terminated agent released
    assert not w_new._ctx_reset_requests and not w_ref._ctx_reset_requests
    assert w_new.embedding_allocator.has_assignment("c") == w_ref.embedding_allocator.has_assignment("c") is False
    # second call: flags cleared, the released agent gets the slot the reference hands out
    _assert_same_batch(service_new.build_batch(w_new, {}), service_ref.build_batch(w_ref, {}))
    single_new = service_new.build_single(w_new, "a", slot=7)
    single_ref = service_ref.build_single(w_ref, "a", slot=7)
    assert single_new["metadata"] == single_ref["metadata"]
    with pytest.raises(KeyError):
        service_new.build_single(w_new, "nobody")


def test_relationship_stage_required(ref):
    from townlet_b200.observation_service import B200ObservationService

    config = _config("hybrid", features__stages__relationships="OFF")
    with pytest.raises(ValueError, match="relationships stage"):
        B200ObservationService(config=config)


def test_reference_stored_goldens_through_the_service(ref):
    """tests/test_observation_baselines.py:49-76 with the drop-in service."""
    from townlet.core.sim_loop import SimulationLoop
    from townlet.world.grid import AgentSnapshot

    from townlet_b200.observation_service import B200ObservationService

    for variant in ("hybrid", "full", "compact"):
        config = _config(variant)
        loop = SimulationLoop(config)
        world = loop.world
        world.agents.clear()
        world.agents["alice"] = AgentSnapshot(agent_id="alice", position=(0, 0), needs={"hunger": 0.4, "hygiene": 0.6, "energy": 0.8}, wallet=5.0)
        world.rebuild_spatial_index()
        payload = B200ObservationService(config=loop.config).build_batch(world, terminated={})["alice"]
        stored = np.load(Path("tests/data/observations") / f"baseline_{variant}.npz", allow_pickle=True)
        np.testing.assert_allclose(payload["map"], stored["map"])
        np.testing.assert_allclose(payload["features"], stored["features"])
        assert payload["metadata"] == stored["metadata"][0]


def test_reward_engine_matches_reference(ref):
    from townlet.rewards.engine import RewardEngine
    from townlet.world.grid import WorldState

    from townlet_b200.reward_engine import B200RewardEngine

    config = _config("hybrid", features__stages__social_rewards="C2")
    w_ref, w_new = WorldState.from_config(config), WorldState.from_config(config)
    _populate(w_ref)
    _populate(w_new)
    e_ref, e_new = RewardEngine(config), B200RewardEngine(config)
    script = {
        0: ({}, {}),
        1: ({"b": True}, {"b": "faint"}),
        2: ({"c": True}, {"c": "eviction"}),
        3: ({}, {}),
        12: ({"a": True}, {}),
    }
    for it in range(16):
        for w in (w_ref, w_new):
            w.tick += 1
            if it == 0:
                w.record_chat_success("a", "b", 0.8)
                w.record_chat_failure("c", "a")
            if it == 3:
                w.record_relationship_guard_block(agent_id="d", reason="queue_rival", target_agent="c", object_id="stove_a")
        terminated, reasons = script.get(it, ({}, {}))
        r_ref = e_ref.compute(w_ref, dict(terminated), dict(reasons))
        r_new = e_new.compute(w_new, dict(terminated), dict(reasons))
        assert list(r_new) == list(r_ref)
        for aid in r_ref:
            a, b = r_new[aid].model_dump(), r_ref[aid].model_dump()
            for key, value in b.items():
                if isinstance(value, float):
                    assert a[key] == pytest.approx(value, rel=1e-6, abs=1e-12), (it, aid, key)
                else:
                    assert a[key] == value, (it, aid, key)
        assert e_new._termination_block == e_ref._termination_block
        assert e_new._episode_totals == pytest.approx(e_ref._episode_totals, rel=1e-6)
        assert e_new.latest_social_events() == e_ref.latest_social_events()
        assert e_new.latest_reward_breakdown().keys() == e_ref.latest_reward_breakdown().keys()
        assert not w_new.consume_chat_events() and not w_new.consume_relationship_avoidance_events()


def test_need_decay_hook_matches_reference(ref):
    from townlet.world.grid import WorldState

    from townlet_b200.reward_engine import install_need_decay

    config = _config("hybrid")
    w_ref, w_new = WorldState.from_config(config), WorldState.from_config(config)
    _populate(w_ref)
    _populate(w_new)
    w_new.agents["e"].needs["hunger"] = w_ref.agents["e"].needs["hunger"] = 0.004  # clamps at 0 (tests/test_need_decay.py:39-60)
    hook = install_need_decay(w_new)
    assert w_new._affordance_service.environment.apply_need_decay is hook
    for _ in range(3):
        w_ref._apply_need_decay()
        w_new._apply_need_decay()
        for aid in w_ref.agents:
            assert w_new.agents[aid].needs == pytest.approx(w_ref.agents[aid].needs, rel=1e-12, abs=0.0), aid
