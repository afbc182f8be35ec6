#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ by RUNNING THE REFERENCE.

Run in the build container (the reference does not exist on the GPU box):

    python tests/golden/make_golden.py            # reference at /root/reference

The script copies the reference to a writable temp dir (it writes ``logs/``
and uses relative config paths), imports it, builds populated worlds the way
the reference's own tests do, runs ``WorldObservationService.build_batch``
(src/townlet/world/observations/service.py:201), ``WorldState._apply_need_decay``
(world/grid.py:1439), ``LifecycleManager.evaluate`` (lifecycle/manager.py:32)
and ``RewardEngine.compute`` (rewards/engine.py:34), ingests the SAME world
states into the SoA batch and stores inputs + reference outputs.

It also re-checks the reference against its own stored fixtures
(tests/data/observations/baseline_*.npz, social_snippet_gold.npz) before
copying those vectors, so the oracle is pinned to the reference's goldens.
"""

from __future__ import annotations

import os
import shutil
import sys
import tempfile
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests" / "golden"))

REF_SRC = Path(os.environ.get("TOWNLET_REFERENCE", "/root/reference"))
if not (REF_SRC / "src" / "townlet").is_dir():  # the GPU box: the copy made by oracle/make_ref.sh
    REF_SRC = REPO / "oracle" / "_ref"
WORK = Path(tempfile.mkdtemp(prefix="townlet_ref_"))
shutil.copytree(REF_SRC, WORK / "ref", symlinks=True)
os.system(f"chmod -R u+w {WORK / 'ref'}")
os.chdir(WORK / "ref")
sys.path.insert(0, str(WORK / "ref" / "src"))
sys.path.insert(0, str(WORK / "ref"))

from fixture_io import GOLDEN_DIR, save_fixture  # noqa: E402

from townlet.agents.models import Personality  # noqa: E402
from townlet.config import load_config  # noqa: E402
from townlet.core.sim_loop import SimulationLoop  # noqa: E402
from townlet.lifecycle.manager import LifecycleManager  # noqa: E402
from townlet.rewards.engine import RewardEngine  # noqa: E402
from townlet.world.grid import AgentSnapshot, WorldState  # noqa: E402
from townlet.world.observations.service import WorldObservationService  # noqa: E402

from townlet_b200 import capi  # noqa: E402
from townlet_b200.ingest import ingest_town, rows_to_batch  # noqa: E402
from townlet_b200.actions import encode_actions  # noqa: E402
from townlet_b200.params import DynamicsSpec, ObsSpec, RewardSpec  # noqa: E402
from townlet_b200.reward_events import normalise_events, reduce_social_events  # noqa: E402
from townlet_b200.synth import TownDescription, assemble, synth_town  # noqa: E402

BASE_CONFIG = Path("configs/examples/poc_hybrid.yaml")
RELATION_SOURCE = {"empty": 0, "disabled": 0, "relationships": 1, "rivalry_fallback": 2}


def make_config(variant: str = "hybrid", **overrides):
    config = load_config(BASE_CONFIG)
    data = config.model_dump()
    data["features"]["systems"]["observations"] = variant
    for dotted, value in overrides.items():
        node = data
        keys = dotted.split("__")
        for key in keys[:-1]:
            node = node[key]
        node[keys[-1]] = value
    return type(config).model_validate(data)


def summary_from_metadata(meta: dict) -> np.ndarray:
    row = np.zeros(capi.TL_N_SUMMARY, dtype=np.int32)
    ls = meta["local_summary"]
    row[0] = int(ls["agent_count"])
    row[1] = int(ls["object_count"])
    row[2] = int(ls["reserved_tiles"])
    row[3] = int(round(ls["nearest_agent_distance"] ** 2))
    sc = meta.get("social_context") or {}
    row[4] = int(sc.get("filled_slots", 0))
    row[5] = RELATION_SOURCE[sc.get("relation_source", "empty")]
    return row


def reference_obs(world, service, terminated):
    batch = service.build_batch(world, terminated)
    ids = list(batch)
    maps = np.stack([batch[i]["map"] for i in ids]) if ids else np.zeros((0,), np.float32)
    feats = np.stack([batch[i]["features"] for i in ids]) if ids else np.zeros((0,), np.float32)
    summ = np.stack([summary_from_metadata(batch[i]["metadata"]) for i in ids])
    return ids, maps, feats, summ, batch


def ingest_for_obs(world, spec, terminated=None, reasons=None, max_edges=6):
    pending = set(world._ctx_reset_requests)  # peek; build_batch consumes (grid.py:932-936)
    slots = {aid: world.embedding_allocator.allocate(aid, world.tick) for aid in world.agents}
    return ingest_town(
        world, spec, terminated=terminated or {}, reasons=reasons or {}, pending_resets=pending, slots=slots, max_edges=max_edges
    )


def obs_fixture(name: str, world, config, terminated=None, note: str = "", check=None):
    terminated = terminated or {}
    spec = ObsSpec.from_config(config)
    rows = ingest_for_obs(world, spec, terminated)
    batch = rows_to_batch([rows], spec)
    batch.validate(len(spec.object_channels))
    service = WorldObservationService(config=config)
    assert tuple(service.feature_names) == spec.feature_names, name
    ids, maps, feats, summ, raw = reference_obs(world, service, terminated)
    assert ids == rows.agent_ids
    assert maps.shape[1:] == spec.map_shape, (maps.shape, spec.map_shape)
    if check is not None:
        check(raw)
    save_fixture(
        GOLDEN_DIR / f"{name}.npz",
        batch,
        spec,
        RewardSpec.from_config(config),
        capi.TL_PHASE_OBS,
        1,
        {"map": maps[None], "features": feats[None], "summary": summ[None]},
        note=note,
    )
    print(f"  {name}: {len(ids)} agents, map {maps.shape[1:]}, F={feats.shape[1]}")


# ---------------------------------------------------------------------------
# 1. the reference's own stored goldens
# ---------------------------------------------------------------------------
def baselines():
    for variant in ("hybrid", "full", "compact"):  # tests/test_observation_baselines.py:17-35
        config = make_config(variant)
        loop = SimulationLoop(config)
        world = loop.world
        world.agents.clear()
        world.agents["alice"] = AgentSnapshot(
            agent_id="alice", position=(0, 0), needs={"hunger": 0.4, "hygiene": 0.6, "energy": 0.8}, wallet=5.0
        )
        world.rebuild_spatial_index()
        stored = np.load(Path("tests/data/observations") / f"baseline_{variant}.npz", allow_pickle=True)

        def check(raw, stored=stored):
            np.testing.assert_array_equal(raw["alice"]["map"], stored["map"])
            np.testing.assert_array_equal(raw["alice"]["features"], stored["features"])
            assert raw["alice"]["metadata"] == stored["metadata"][0]

        obs_fixture(
            f"ref_baseline_{variant}",
            world,
            config,
            note=f"reference fixture tests/data/observations/baseline_{variant}.npz (bit-identical, checked at generation)",
            check=check,
        )


def social_snippet():
    config = load_config(BASE_CONFIG)  # tests/test_observations_social_snippet.py:20-41
    world = WorldState.from_config(config)
    world.agents["alice"] = AgentSnapshot(
        agent_id="alice", position=(0, 0), needs={"hunger": 0.4, "hygiene": 0.5, "energy": 0.6}, wallet=5.0
    )
    world.agents["bob"] = AgentSnapshot(
        agent_id="bob", position=(1, 0), needs={"hunger": 0.3, "hygiene": 0.2, "energy": 0.7}, wallet=3.0
    )
    world.register_rivalry_conflict("alice", "bob", intensity=1.0)  # tests/helpers/social.py:19-24
    world.update_relationship("alice", "bob", trust=0.6, familiarity=0.3)
    stored = np.load("tests/data/observations/social_snippet_gold.npz", allow_pickle=True)

    def check(raw):
        np.testing.assert_allclose(raw["alice"]["features"], stored["features"], atol=1e-6)
        assert raw["alice"]["metadata"]["social_context"]["relation_source"] == "relationships"

    obs_fixture(
        "ref_social_snippet",
        world,
        config,
        note="reference fixture tests/data/observations/social_snippet_gold.npz (atol 1e-6, checked at generation)",
        check=check,
    )


# ---------------------------------------------------------------------------
# 2. behavioural cases of the reference tests, as vectors
# ---------------------------------------------------------------------------
def builder_cases():
    # tests/test_observation_builder.py:13-38 (targets on, two agents, stove, job loop)
    config = make_config("hybrid", observations_config__hybrid__include_targets=True)
    config.conflict.rivalry.avoid_threshold = 0.1
    loop = SimulationLoop(config)
    world = loop.world
    world.agents.clear()
    world.register_object(object_id="stove_test", object_type="stove", position=(2, 0))
    world.agents["alice"] = AgentSnapshot(
        agent_id="alice",
        position=(0, 0),
        needs={"hunger": 0.4, "hygiene": 0.5, "energy": 0.6},
        wallet=2.0,
        last_action_id="wait",
        last_action_success=True,
        last_action_duration=4,
    )
    world.agents["bob"] = AgentSnapshot(
        agent_id="bob", position=(1, 0), needs={"hunger": 0.7, "hygiene": 0.8, "energy": 0.9}, wallet=3.0
    )
    world.assign_jobs_to_agents()
    obs_fixture("builder_hybrid_targets", world, config, note="tests/test_observation_builder.py:41-99")
    world.register_rivalry_conflict("alice", "bob")  # :115-124
    world.request_ctx_reset("bob") if hasattr(world, "request_ctx_reset") else world._ctx_reset_requests.add("bob")
    obs_fixture(
        "builder_hybrid_rivalry_ctxreset",
        world,
        config,
        terminated={"alice": True},
        note="rivalry fallback + pending ctx reset + terminated (tests/test_observation_builder.py:102-124)",
    )


def crowded_cases():
    # Appendix A3/A5: stacked agents and objects, reservations, all variants; unindexed agents (A1)
    for variant, extra in (
        ("hybrid", {}),
        ("full", {"observations_config__hybrid__include_targets": True}),
        (
            "compact",
            {
                "observations_config__compact__object_channels": ["stove", "Fridge ", "stove"],
                "observations_config__compact__normalize_counts": False,
                "observations_config__compact__include_targets": True,
            },
        ),
        ("compact", {}),
    ):
        config = make_config(variant, **extra)
        world = WorldState.from_config(config)
        world.tick = 777
        for oid, otype, pos in (
            ("stove_a", "stove", (2, 1)),
            ("stove_b", "stove", (2, 1)),
            ("fridge_a", "fridge", (2, 1)),
            ("bed_a", "bed", (-3, 2)),
            ("shower_a", "shower", (0, -4)),
            ("Stove_caps", "Stove", (1, 1)),
            ("fridge_far", "fridge", (9, 9)),
        ):
            world.register_object(object_id=oid, object_type=otype, position=pos)
        spots = {
            "a": (0, 0),
            "b": (0, 0),
            "c": (2, 0),
            "d": (2, 1),
            "e": (2, 1),
            "f": (-5, -5),
            "g": (3, 3),
            "h": (-6, 0),
        }
        for k, (aid, pos) in enumerate(spots.items()):
            world.agents[aid] = AgentSnapshot(
                agent_id=aid,
                position=pos,
                needs={"hunger": 0.1 * (k + 1), "hygiene": 0.9 - 0.1 * k, "energy": 0.5},
                wallet=float(k),
                shift_state=("pre_shift", "on_time", "late", "absent", "post_shift", "idle", "await_start", "on_time")[k],
                on_shift=bool(k % 2),
                lateness_counter=k,
                attendance_ratio=k / 8.0,
                wages_withheld=0.01 * k,
                last_action_id=("", "move", "request", "start", "release", "chat", "blocked", "noop")[k],
                last_action_success=bool(k % 3 == 0),
                last_action_duration=k,
                episode_tick=200 * k,
            )
        world.rebuild_spatial_index()
        world.agents["h"].position = (-6, 1)  # snapshot.position differs from the indexed tile (cache.py:30)
        # A1: an agent assigned without touching the spatial index
        world.agents["ghost"] = AgentSnapshot(
            agent_id="ghost", position=(1, 0), needs={"hunger": 0.5, "hygiene": 0.5, "energy": 0.5}
        )
        # reservations: via active reservations -> tiles
        world._active_reservations["stove_a"] = "d"
        world._active_reservations["bed_a"] = "f"
        world._spatial_index.set_reservation((2, 1), True)
        world._spatial_index.set_reservation((-3, 2), True)
        world._spatial_index.set_reservation((4, 4), True)  # a reserved tile without an object
        world.queue_manager.request_access("shower_a", "g", world.tick)
        world.queue_manager.request_access("shower_a", "h", world.tick)
        # ties: more than top_friends + top_rivals, with ties in the sort keys
        rel = world._relationships
        rel.get_relationship_ledger("a").inject(
            {
                "b": {"trust": 0.2, "familiarity": 0.1, "rivalry": 0.05},
                "c": {"trust": 0.1, "familiarity": 0.2, "rivalry": 0.30},
                "d": {"trust": 0.5, "familiarity": 0.0, "rivalry": 0.30},
                "e": {"trust": -0.2, "familiarity": 0.1, "rivalry": 0.00},
                "zed": {"trust": 0.0, "familiarity": 0.0, "rivalry": 0.9},
            }
        )
        rel.get_relationship_ledger("b").inject({"a": {"trust": 0.3, "familiarity": 0.3, "rivalry": 0.0}})
        rel.get_rivalry_ledger("c").inject([("a", 0.4), ("d", 0.8), ("e", 0.8), ("g", 0.05)])
        rel.get_rivalry_ledger("a").inject([("c", 0.75), ("b", 0.2)])
        tag = variant if not extra else f"{variant}_opts"
        obs_fixture(f"crowded_{tag}", world, config, terminated={"c": True}, note="SURVEY Appendix A1/A3/A5/A6 behaviours")


def reservation_cases():
    """TL_PHASE_RESERVATIONS: the reference's ``refresh_reservations``
    (world/systems/queues.py:55-86, driven by ``queues.step`` :14-31) against the device phase.  Each world is ingested
    in a STALE state — grants made and withdrawn through the queue manager without a refresh, as between two ticks
    (Appendix A6) — then the reference refreshes and observes; the fixture stores the stale batch, the post-refresh
    observation and the post-refresh entity words / flags of a second ingest."""
    for variant in ("hybrid", "compact"):
        config = make_config(variant)
        spec = ObsSpec.from_config(config)
        world = WorldState.from_config(config)
        world.tick = 4321
        for oid, otype, pos in (
            ("stove_a", "stove", (2, 1)),
            ("fridge_a", "fridge", (2, 1)),   # shares the stove's tile: the later object decides the tile
            ("bed_a", "bed", (-3, 2)),
            ("bed_b", "bed", (-3, 2)),        # shares bed_a's tile
            ("shower_a", "shower", (0, -4)),
            ("stove_far", "stove", (6, 6)),
        ):
            world.register_object(object_id=oid, object_type=otype, position=pos)
        spots = {"a": (0, 0), "b": (1, 0), "c": (2, 0), "d": (2, 1), "e": (-3, 1), "f": (0, -3)}
        for k, (aid, pos) in enumerate(spots.items()):
            world.agents[aid] = AgentSnapshot(agent_id=aid, position=pos, needs={"hunger": 0.3 + 0.1 * k, "hygiene": 0.6, "energy": 0.7},
                                              wallet=float(k), episode_tick=10 * k)
        world.rebuild_spatial_index()
        qm = world.queue_manager
        # state of the previous tick, refreshed: stove_a held by d, shower_a held by f with a waiting queue
        qm.request_access("stove_a", "d", world.tick)
        qm.request_access("shower_a", "f", world.tick)
        qm.request_access("shower_a", "a", world.tick)  # a waits (in_queue, not a holder)
        world.refresh_reservations()
        world._spatial_index.set_reservation((4, 4), True)  # a reserved tile no object accounts for: untouched by the refresh
        # between the ticks, through the queue manager only (no sync): d lets go of the stove, b is granted the fridge on
        # the same tile, e is granted bed_a while bed_b (later, same tile) stays free, c is granted the far stove
        qm.release("stove_a", "d", world.tick, success=True)
        qm.request_access("fridge_a", "b", world.tick)
        qm.request_access("bed_a", "e", world.tick)
        qm.request_access("stove_far", "c", world.tick)
        world.tick += 1
        pending = set(world._ctx_reset_requests)
        slots = {aid: world.embedding_allocator.allocate(aid, world.tick) for aid in world.agents}
        rows = ingest_town(world, spec, terminated={}, reasons={}, pending_resets=pending, slots=slots, device_reservations=True)
        batch = rows_to_batch([rows], spec)
        batch.validate(len(spec.object_channels))
        stale_kinds = (batch.ent >> 20) & 3
        world.refresh_reservations()  # what queues.step does at the start of the tick
        service = WorldObservationService(config=config)
        ids, maps, feats, summ, _ = reference_obs(world, service, {})
        assert ids == rows.agent_ids
        after = ingest_town(world, spec, terminated={}, reasons={}, pending_resets=set(), slots=slots, device_reservations=True,
                            origin=rows.origin)
        after_batch = rows_to_batch([after], spec, grid=(batch.grid_w, batch.grid_h))
        assert np.array_equal(after_batch.ent_holder, batch.ent_holder)
        assert not np.array_equal((after_batch.ent >> 20) & 3, stale_kinds), "the refresh must change some slot"
        assert world.reservation_tiles() == frozenset({(2, 1), (0, -4), (6, 6), (4, 4)}), world.reservation_tiles()
        save_fixture(
            GOLDEN_DIR / f"reservations_{variant}.npz", batch, spec, RewardSpec.from_config(config),
            capi.TL_PHASE_RESERVATIONS | capi.TL_PHASE_OBS, 1,
            {"map": maps[None], "features": feats[None], "summary": summ[None], "ent_after": after_batch.ent[None],
             "flags_after": after_batch.flags[None]},
            note="stale reservation slots + holders; the reference's refresh_reservations and observation",
        )
        print(f"  reservations_{variant}: {len(ids)} agents, {int((batch.ent_holder >= -1).sum())} slots, tiles {sorted(world.reservation_tiles())}")


def employment_case():
    """``EmploymentEngine.apply_job_state`` (world/employment.py:63-142) run by the reference for two
    and a half days on a scripted town — punctual, late, absent, early-leaving and jobless agents, coworkers on the same
    shift, an attendance window that rolls — with every context key, the snapshot fields the hot path reads and the events
    the reference emits recorded after every tick."""
    from townlet_b200.employment import EVENT_BITS, SHIFT_FLAG_CODE, EmploymentArrays, EmploymentSpec

    config = make_config(
        "hybrid",
        employment__exit_review_window=100, employment__grace_ticks=3, employment__absent_cutoff=12, employment__absence_slack=6,
        employment__attendance_window=2, employment__enforce_job_loop=True, behavior__job_arrival_buffer=5,
        jobs={"grocer": {"start_tick": 20, "end_tick": 60, "wage_rate": 0.02, "lateness_penalty": 0.1, "location": [0, 0]},
              "barista": {"start_tick": 30, "end_tick": 50, "wage_rate": 0.025, "lateness_penalty": 0.12, "location": [1, 0]},
              "courier": {"start_tick": 70, "end_tick": 0, "wage_rate": 0.0, "lateness_penalty": 0.0, "location": None}},
    )
    spec = EmploymentSpec.from_config(config)
    world = WorldState.from_config(config)
    jobs_of = {"a": "grocer", "b": "grocer", "c": "grocer", "d": "barista", "e": None, "f": "barista", "g": "courier"}
    for k, (aid, job) in enumerate(jobs_of.items()):
        world.agents[aid] = AgentSnapshot(agent_id=aid, position=(5, 5), needs={"hunger": 0.9, "hygiene": 0.9, "energy": 0.9},
                                          wallet=0.15 + 0.05 * k, job_id=job)
    agent_ids = list(world.agents)
    home, grocer, barista = (5, 5), (0, 0), (1, 0)

    def position(aid: str, tick: int) -> tuple[int, int]:
        day, t = divmod(tick, 100)
        if aid == "a":  # punctual every day
            return grocer if 15 <= t <= 65 else home
        if aid == "b":  # day 0: arrives after the grace period; day 1: on time; day 2: leaves early for good
            if day == 0:
                return grocer if 30 <= t <= 65 else home
            if day == 1:
                return grocer if 18 <= t <= 65 else home
            return grocer if 20 <= t <= 35 else home
        if aid == "c":  # never shows up on day 0 and 2 (absent, "took my shift"); late within grace then present on day 1
            if day == 1:
                return grocer if 22 <= t <= 65 else home
            return home
        if aid == "d":  # barista: on time, steps out for longer than the absence slack, comes back
            return barista if (30 <= t <= 36 or 46 <= t <= 50) else home
        if aid == "e":  # no job: takes the default one (grocer) and arrives late
            return grocer if 27 <= t <= 65 else home
        if aid == "f":  # barista: steps out briefly (shorter than the slack)
            return barista if (29 <= t <= 38 or 42 <= t <= 50) else home
        return home  # g: a job without a location is always "at the required location"

    emitted: list[tuple[str, dict]] = []
    engine = world.employment.engine  # the coordinator is a façade (world/employment_service.py:15-40)
    original_emit = engine._emit_event
    engine._emit_event = lambda name, payload: (emitted.append((name, dict(payload))), original_emit(name, payload))[1]
    n_ticks = 250
    N = len(agent_ids)
    names = [name for name, _, _ in capi.EMPLOYMENT_STATE_FIELDS]
    ref: dict[str, np.ndarray] = {f"emp_{name}": [] for name in names}
    extra = {k: [] for k in ("wallet", "wages_withheld", "lateness", "attendance", "wages_paid", "punctuality", "on_shift", "shift_code", "events", "pos")}
    for tick in range(n_ticks):
        world.tick = tick
        for aid in agent_ids:
            world.agents[aid].position = position(aid, tick)
        emitted.clear()
        engine.apply_job_state(world)
        arrays = EmploymentArrays(N)
        for row, aid in enumerate(agent_ids):
            arrays.load_context(row, engine._state[aid], world.agents[aid], spec)
        for name in names:
            ref[f"emp_{name}"].append(arrays.arrays[name].copy())
        snaps = [world.agents[aid] for aid in agent_ids]
        extra["wallet"].append([float(s.wallet) for s in snaps])
        extra["wages_withheld"].append([float(s.wages_withheld) for s in snaps])
        extra["lateness"].append([int(s.lateness_counter) for s in snaps])
        extra["attendance"].append([float(s.attendance_ratio) for s in snaps])
        extra["wages_paid"].append([float(engine.employment_context_wages(aid)) for aid in agent_ids])
        extra["punctuality"].append([float(engine.employment_context_punctuality(aid)) for aid in agent_ids])
        extra["on_shift"].append([int(bool(s.on_shift)) for s in snaps])
        extra["shift_code"].append([SHIFT_FLAG_CODE.get(s.shift_state, 0) for s in snaps])
        words = [0] * N
        for name, payload in emitted:
            if name in EVENT_BITS:
                words[agent_ids.index(payload["agent_id"])] |= EVENT_BITS[name]
        extra["events"].append(words)
        extra["pos"].append([list(position(aid, tick)) for aid in agent_ids])
    out = {k: np.asarray(v) for k, v in ref.items()}
    out.update({k: np.asarray(v) for k, v in extra.items()})
    out["initial_wallet"] = np.asarray([0.15 + 0.05 * k for k in range(N)])
    out["initial_job"] = np.asarray([spec.job_ids.index(j) if j else -1 for j in jobs_of.values()], dtype=np.int32)
    import json as _json
    from dataclasses import asdict

    out["spec"] = np.array(_json.dumps(asdict(spec)))
    (GOLDEN_DIR / "employment").mkdir(exist_ok=True)
    np.savez_compressed(GOLDEN_DIR / "employment" / "ref_employment.npz", **out)
    seen = sorted({name for words in extra["events"] for name, bit in EVENT_BITS.items() if any(w & bit for w in words)})
    states = sorted({int(v) for v in np.unique(out["emp_state"])})
    print(f"  ref_employment: {N} agents x {n_ticks} ticks, events {seen}, states {states}, attendance {np.unique(out['attendance']).round(3)}")


# ---------------------------------------------------------------------------
# 3. synthetic towns of the benchmark generator, materialised in the reference
# ---------------------------------------------------------------------------
def materialise(town: TownDescription, config) -> WorldState:
    world = WorldState.from_config(config)
    world.tick = town.tick
    for oid, otype, pos in zip(town.object_ids, town.object_types, town.object_positions):
        world.register_object(object_id=oid, object_type=otype, position=(int(pos[0]), int(pos[1])))
    for i, aid in enumerate(town.agent_ids):
        snap = AgentSnapshot(
            agent_id=aid,
            position=(int(town.positions[i, 0]), int(town.positions[i, 1])),
            needs={"hunger": float(town.needs[i, 0]), "hygiene": float(town.needs[i, 1]), "energy": float(town.needs[i, 2])},
            wallet=float(town.wallet[i]),
            personality_profile=town.profile[i],
            on_shift=bool(town.on_shift[i]),
            lateness_counter=int(town.lateness[i]),
            shift_state=town.shift_state[i],
            attendance_ratio=float(town.attendance[i]),
            wages_withheld=float(town.wages_withheld[i]),
            last_action_id=town.last_action_id[i],
            last_action_success=bool(town.last_action_success[i]),
            last_action_duration=int(town.last_action_duration[i]),
            episode_tick=int(town.episode_tick[i]),
        )
        snap.personality = Personality(
            extroversion=float(town.personality[i, 0]),
            forgiveness=float(town.personality[i, 1]),
            ambition=float(town.personality[i, 2]),
        )
        world.agents[aid] = snap
    for oid, holder in town.reservations.items():
        world._active_reservations[oid] = holder
    world.rebuild_spatial_index()
    from townlet.world.queue.manager import QueueEntry

    for i, aid in enumerate(town.agent_ids):  # waiting lists read by features.py:221-229
        if town.in_queue[i]:
            oid = town.object_ids[i % len(town.object_ids)]
            world.queue_manager._queues.setdefault(oid, []).append(QueueEntry(agent_id=aid, joined_tick=town.tick))
    rel = world._relationships
    for i, aid in enumerate(town.agent_ids):
        if town.ties[i]:
            rel.get_relationship_ledger(aid).inject(
                {other: {"trust": t, "familiarity": f, "rivalry": r} for other, t, f, r in town.ties[i]}
            )
        if town.rivalry[i]:
            rel.get_rivalry_ledger(aid).inject(town.rivalry[i])
    return world


def patch_employment(world, town: TownDescription):
    """Seed the employment context so the reference's wage / punctuality hooks return the synthetic values."""
    for i, aid in enumerate(town.agent_ids):  # world/employment.py:209-222 reads these keys
        ctx = world.employment_service.get_context(aid)
        ctx["wages_paid"] = float(town.wages_paid[i])
        ctx["scheduled_ticks"] = 1000
        ctx["on_time_ticks"] = int(round(float(town.punctuality[i]) * 1000))


def reference_kpi_row(world, rewards: dict, terminated: dict, config) -> tuple[np.ndarray, float, float]:
    """The per-town KPI row computed by the REFERENCE's own code on one world after a tick:
    mean lateness by ``TelemetryPublisher._update_kpi_history`` (telemetry/publisher.py:2671-2694), reward mean and
    population variance by ``RewardVarianceAnalyzer.analyze`` (stability/analyzers/reward_variance.py:104-108), the
    starving count by ``StarvationDetector.analyze``'s predicate (stability/analyzers/starvation.py:132-139), hunger
    levels as ``SimulationLoop.step`` collects them (core/sim_loop.py:740-743), the reward sum as
    ``compute_sample_metrics`` takes it (policy/metrics.py:13-25)."""
    from types import SimpleNamespace

    from townlet.stability.analyzers.reward_variance import RewardVarianceAnalyzer
    from townlet.stability.analyzers.starvation import StarvationDetector
    from townlet.telemetry.publisher import TelemetryPublisher

    stub = SimpleNamespace(_latest_conflict_snapshot={}, _latest_relationship_metrics={}, _kpi_history={})
    TelemetryPublisher._update_kpi_history(stub, world)
    lateness_mean = stub._kpi_history["employment_lateness"][-1]
    variance = RewardVarianceAnalyzer(window_ticks=1).analyze(tick=world.tick, rewards=dict(rewards))
    threshold = float(config.stability.starvation.hunger_threshold)
    hunger_levels = {aid: float(snap.needs.get("hunger", 1.0)) for aid, snap in world.agents.items()}
    starving = StarvationDetector(hunger_threshold=threshold, min_duration_ticks=1, window_ticks=10).analyze(
        tick=world.tick, hunger_levels=hunger_levels, terminated={}
    )["starvation_incidents"]
    totals = np.asarray([rewards[aid] for aid in world.agents], dtype=np.float32)
    row = np.zeros(capi.TL_N_KPI, dtype=np.float64)
    row[capi.KPI_NAMES.index("reward_sum")] = float(np.sum(totals))  # policy/metrics.py:18
    row[capi.KPI_NAMES.index("reward_sumsq")] = float(np.sum(totals.astype(np.float64) ** 2))
    row[capi.KPI_NAMES.index("hunger_min")] = min(hunger_levels.values())
    row[capi.KPI_NAMES.index("hunger_mean")] = sum(hunger_levels.values()) / len(hunger_levels)
    row[capi.KPI_NAMES.index("starving")] = starving
    row[capi.KPI_NAMES.index("terminated")] = sum(1 for flag in terminated.values() if flag)
    row[capi.KPI_NAMES.index("lateness_mean")] = lateness_mean
    row[capi.KPI_NAMES.index("wallet_mean")] = sum(float(snap.wallet) for snap in world.agents.values()) / len(world.agents)
    return row, float(variance["reward_mean"]), float(variance["reward_variance"])


def synth_cases():
    cases = (
        ("synth_hybrid_48x8", "hybrid", 48, 8, 4, {}),
        ("synth_full_48x10", "full", 48, 10, 3, {}),
        ("synth_compact_48x10", "compact", 48, 10, 3, {"observations_config__compact__object_channels": ["stove", "bed"]}),
        ("synth_hybrid_128x64", "hybrid", 128, 64, 1, {"observations_config__hybrid__include_targets": True}),
        (
            "synth_hybrid_personality",
            "hybrid",
            24,
            12,
            2,
            {"features__observations__personality_channels": True, "features__behavior__reward_multipliers": True},
        ),
    )
    for name, variant, grid, agents, n_towns, extra in cases:
        config = make_config(variant, employment__enforce_job_loop=False, **extra)
        spec = ObsSpec.from_config(config)
        rew = RewardSpec.from_config(config, fuse_faint=True)
        towns = [synth_town(t, grid, agents, terminated_frac=0.1) for t in range(n_towns)]
        # make some agents starve during the sequence and some sit at the episode cap
        for town in towns:
            town.needs[0, 0] = 0.05
            town.needs[1 % agents, 0] = 0.031
        n_ticks = 6
        worlds = [materialise(t, config) for t in towns]
        for w, t in zip(worlds, towns):
            patch_employment(w, t)
        services = [WorldObservationService(config=config) for _ in towns]
        engines = [RewardEngine(config) for _ in towns]
        lifecycles = [LifecycleManager(config) for _ in towns]
        for eng, t in zip(engines, towns):
            eng._episode_totals[t.agent_ids[-1]] = 49.9
            eng._episode_totals[t.agent_ids[-2]] = -49.95
        # ---- ingest the initial state (input of the fused tick) -------------
        rows = []
        for w, t, eng in zip(worlds, towns, engines):
            term0 = {aid: bool(t.terminated[i]) for i, aid in enumerate(t.agent_ids)}
            reasons0 = {aid: "eviction" if i % 2 else "unemployment" for i, aid in enumerate(t.agent_ids) if t.terminated[i]}
            pending = set()
            slots = {aid: w.embedding_allocator.allocate(aid, w.tick) for aid in w.agents}
            rows.append(
                ingest_town(
                    w,
                    spec,
                    terminated=term0,
                    reasons=reasons0,
                    pending_resets=pending,
                    slots=slots,
                    episode_totals=dict(eng._episode_totals),
                    termination_block=dict(eng._termination_block),
                    origin=(0, 0),
                )
            )
        batch = rows_to_batch(rows, spec, grid=(grid, grid))
        batch.validate(len(spec.object_channels))
        # the synthetic assembler must agree with the ingest of the materialised world
        direct = assemble(towns, spec)
        for fname, view in batch.views().items():
            other = getattr(direct, fname)
            if fname in ("episode_total", "slot"):
                continue
            if fname == "flags":
                mask = ~np.uint32(capi.TL_F_REASON_MASK)
                assert np.array_equal(view & mask, other & mask), fname
                continue
            assert np.array_equal(view, other), f"assemble vs ingest mismatch in {fname}"
        # ---- run the reference for n_ticks ticks ------------------------------
        N = batch.n_agents
        ref = {
            "map": np.zeros((n_ticks, N, *spec.map_shape), np.float32),
            "features": np.zeros((n_ticks, N, spec.n_features), np.float32),
            "summary": np.zeros((n_ticks, N, capi.TL_N_SUMMARY), np.int32),
            "comps": np.zeros((n_ticks, N, capi.TL_N_COMPS)),
            "terminated": np.zeros((n_ticks, N), np.uint8),
            "needs": np.zeros((n_ticks, 3, N)),
            "episode_total": np.zeros((n_ticks, N)),
            # row K: the reference's own telemetry / stability code on every town after every tick
            "kpi": np.zeros((n_ticks, len(towns), capi.TL_N_KPI)),
            "kpi_reward_mean": np.zeros((n_ticks, len(towns))),
            "kpi_reward_variance": np.zeros((n_ticks, len(towns))),
        }
        for it in range(n_ticks):
            a = 0
            for ti, (w, t, svc, eng, life) in enumerate(zip(worlds, towns, services, engines, lifecycles)):
                n = len(t.agent_ids)
                w.tick += 1  # core/sim_loop.py:635,659
                w._apply_need_decay()  # world/grid.py:1439
                life_term = life.evaluate(w, w.tick)  # lifecycle/manager.py:32-46
                life_reasons = dict(life._termination_reasons)
                terminated = {}
                reasons = {}
                for i, aid in enumerate(t.agent_ids):
                    host_flag = bool(t.terminated[i])  # persistent host-side terminations (employment exits)
                    terminated[aid] = host_flag or bool(life_term.get(aid, False))
                    if host_flag:
                        reasons[aid] = "eviction" if i % 2 else "unemployment"
                    elif aid in life_reasons:
                        reasons[aid] = life_reasons[aid]
                tpd = config.observations_config.hybrid.time_ticks_per_day
                for snap in w.agents.values():  # world/core/context.py:283-289
                    snap.episode_tick = (int(snap.episode_tick) + 1) % max(1, int(tpd))
                breakdowns = eng.compute(w, terminated, reasons)
                row, r_mean, r_var = reference_kpi_row(w, {aid: bd.total for aid, bd in breakdowns.items()}, terminated, config)
                ref["kpi"][it, ti], ref["kpi_reward_mean"][it, ti], ref["kpi_reward_variance"][it, ti] = row, r_mean, r_var
                ids, maps, feats, summ, _ = reference_obs(w, svc, terminated)
                assert ids == t.agent_ids
                ref["map"][it, a : a + n] = maps
                ref["features"][it, a : a + n] = feats
                ref["summary"][it, a : a + n] = summ
                for i, aid in enumerate(ids):
                    bd = breakdowns[aid]
                    row = ref["comps"][it, a + i]
                    for c, cname in enumerate(capi.COMP_NAMES):
                        if cname == "guardrail_active":
                            row[c] = float(bd.guardrail_active)
                        elif cname == "clipped":
                            row[c] = float(bd.clipped)
                        elif cname != "reserved":
                            row[c] = float(getattr(bd, cname))
                    ref["terminated"][it, a + i] = terminated[aid]
                    snap = w.agents[aid]
                    ref["needs"][it, :, a + i] = [snap.needs["hunger"], snap.needs["hygiene"], snap.needs["energy"]]
                    ref["episode_total"][it, a + i] = eng._episode_totals[aid]
                # the reference allocator released slots of terminated agents; re-allocate deterministically
                for aid in ids:
                    w.embedding_allocator._assignments.setdefault(aid, int(rows[ti].fields["slot"][ids.index(aid)]))
                a += n
        save_fixture(
            GOLDEN_DIR / f"{name}.npz",
            batch,
            spec,
            rew,
            capi.TL_PHASE_ALL,
            n_ticks,
            ref,
            note="synthetic towns materialised as reference WorldState; decay -> lifecycle -> reward -> observe per tick",
        )
        print(f"  {name}: {n_towns} towns x {agents} agents, {n_ticks} ticks")


# ---------------------------------------------------------------------------
# 4. reward engine cases with social events (tests/test_reward_engine.py)
# ---------------------------------------------------------------------------
def reward_cases():
    config = make_config("hybrid", features__stages__social_rewards="C2")
    rew = RewardSpec.from_config(config)
    spec = ObsSpec.from_config(config)
    world = WorldState.from_config(config)
    world.tick = 100
    needs = {
        "alice": (0.5, 0.5, 0.5),
        "bob": (0.9, 0.2, 0.7),
        "carol": (0.1, 1.0, 1.0),  # deficit > 0.85 -> chat bonus suppressed (engine.py:348-353)
        "dave": (1.0, 1.0, 1.0),
    }
    for k, (aid, (h, hy, e)) in enumerate(needs.items()):
        world.agents[aid] = AgentSnapshot(agent_id=aid, position=(k, 0), needs={"hunger": h, "hygiene": hy, "energy": e}, wallet=1.0)
    world.rebuild_spatial_index()
    world.update_relationship("alice", "bob", trust=0.4, familiarity=0.2)
    engine = RewardEngine(config)
    n_ticks = 14
    N = len(needs)
    ref = {"comps": np.zeros((n_ticks, N, capi.TL_N_COMPS)), "episode_total": np.zeros((n_ticks, N))}
    rows_per_tick = []
    script = {
        0: [("chat_success", "alice", "bob", 1.0), ("chat_success", "carol", "dave", 0.5)],
        1: [("chat_failure", "bob", "alice", None)],
        2: [("avoid", "dave")],
        3: [("chat_success", "alice", "bob", 2.0), ("chat_success", "alice", "nobody", 0.25), ("avoid", "ghost")],
    }
    term_script = {4: ({"bob": True}, {"bob": "faint"}), 5: ({"dave": True}, {"dave": "eviction"}), 9: ({"alice": True}, {})}
    for it in range(n_ticks):
        world.tick += 1
        for ev in script.get(it, []):
            if ev[0] == "chat_success":
                world.record_chat_success(ev[1], ev[2], ev[3])
            elif ev[0] == "chat_failure":
                world.record_chat_failure(ev[1], ev[2])
            else:
                world.record_relationship_guard_block(agent_id=ev[1], reason="queue_rival", target_agent="x", object_id="o")
        terminated, reasons = term_script.get(it, ({}, {}))
        totals_before = dict(engine._episode_totals)
        block_before = dict(engine._termination_block)
        recorded: dict[str, list] = {}
        orig_chat, orig_avoid = world.consume_chat_events, world.consume_relationship_avoidance_events

        def rec_chat(orig=orig_chat):
            recorded["chat"] = list(orig())
            return recorded["chat"]

        def rec_avoid(orig=orig_avoid):
            recorded["avoid"] = list(orig())
            return recorded["avoid"]

        world.consume_chat_events = rec_chat
        world.consume_relationship_avoidance_events = rec_avoid
        breakdowns = engine.compute(world, dict(terminated), dict(reasons))
        world.consume_chat_events, world.consume_relationship_avoidance_events = orig_chat, orig_avoid
        sums = reduce_social_events(
            world, normalise_events(recorded.get("chat", [])), normalise_events(recorded.get("avoid", [])), rew
        )
        rows_per_tick.append(
            ingest_town(
                world,
                spec,
                terminated=terminated,
                reasons=reasons,
                episode_totals=totals_before,
                termination_block=block_before,
                social_sums=sums,
                slots={aid: k for k, aid in enumerate(world.agents)},
            )
        )
        for i, aid in enumerate(world.agents):
            bd = breakdowns[aid]
            for c, cname in enumerate(capi.COMP_NAMES):
                if cname == "guardrail_active":
                    ref["comps"][it, i, c] = float(bd.guardrail_active)
                elif cname == "clipped":
                    ref["comps"][it, i, c] = float(bd.clipped)
                elif cname != "reserved":
                    ref["comps"][it, i, c] = float(getattr(bd, cname))
            ref["episode_total"][it, i] = engine._episode_totals[aid]
    # one "town" per tick: each row set is an independent single-tick reward problem
    batch = rows_to_batch(rows_per_tick, spec, grid=(16, 16))
    flat = {
        "comps": ref["comps"].reshape(1, n_ticks * N, capi.TL_N_COMPS),
        "episode_total": ref["episode_total"].reshape(1, n_ticks * N),
    }
    save_fixture(
        GOLDEN_DIR / "reward_social_events.npz",
        batch,
        spec,
        rew,
        capi.TL_PHASE_REWARD,
        1,
        flat,
        note="RewardEngine.compute over 14 ticks with chat/avoidance events, faint/eviction, death window; one town per tick",
    )
    print(f"  reward_social_events: {n_ticks} ticks x {N} agents")


def snapshot_cases():
    """BASELINE configs[2]: parity on world states LOADED FROM SNAPSHOTS at ticks 10 / 25 / 50.

    The shipped snapshot-{10,25,50}.json hold zero agents, so the same recipe
    (tests/fixtures/baselines/capture_snapshots.py:16-57) is re-run on a populated loop: positioned objects,
    eight spawned agents, `loop.step()` to the target tick, `snapshot_from_world` -> `SnapshotManager.save` ->
    `SnapshotManager.load` -> `apply_snapshot_to_world` into a FRESH world, and the observation service runs on
    that live restored state.
    """
    from townlet.snapshots import SnapshotManager, apply_snapshot_to_world, snapshot_from_world

    objects = (
        ("fridge_1", "fridge", (2, 2)), ("stove_1", "stove", (5, 2)), ("bed_1", "bed", (8, 3)), ("shower_1", "shower", (3, 7)),
        ("fridge_2", "fridge", (10, 9)), ("stove_2", "stove", (12, 4)), ("bed_2", "bed", (6, 11)), ("shower_2", "shower", (14, 12)),
    )
    spawns = ((1, 1), (4, 4), (7, 2), (9, 6), (3, 9), (11, 11), (13, 5), (6, 7))
    for variant in ("hybrid", "full", "compact"):
        config = make_config(variant, runtime__telemetry__provider="stub")
        loop = SimulationLoop(config)
        for oid, otype, pos in objects:
            loop.world.register_object(object_id=oid, object_type=otype, position=pos)
        for i, pos in enumerate(spawns):
            loop.world.lifecycle_service.spawn_agent(
                f"agent_{i}", pos, needs={"hunger": 0.9 - 0.08 * i, "hygiene": 0.5 + 0.05 * i, "energy": 0.7}, wallet=2.0 + i
            )
        manager = SnapshotManager(WORK / f"snapshots_{variant}")
        for tick_target in (10, 25, 50):
            while loop.tick < tick_target:
                loop.step()
            snapshot = snapshot_from_world(
                config=loop.config, world=loop.world, lifecycle=loop.lifecycle, telemetry=loop.telemetry,
                perturbations=loop.perturbations, stability=loop.stability, promotion=loop.promotion,
                rng_streams={"world": loop._rng_world, "events": loop._rng_events, "policy": loop._rng_policy},
            )
            path = manager.save(snapshot)
            restored = SimulationLoop(config)
            for oid, otype, pos in objects:  # the snapshot schema drops object positions (snapshots/state.py:92-123)
                restored.world.register_object(object_id=oid, object_type=otype, position=pos)
            apply_snapshot_to_world(restored.world, manager.load(path, config), lifecycle=restored.lifecycle)
            assert len(restored.world.agents) >= 1, "snapshot lost its agents"
            obs_fixture(
                f"snapshot_{tick_target}_{variant}",
                restored.world,
                config,
                note=f"world state restored from a populated snapshot-{tick_target}.json (capture_snapshots.py recipe)",
            )


def world_dynamics_cases():
    """actions (moves + outcome bookkeeping) and ledger decay inside the tick.

    Per tick, in WorldContext.tick order (world/core/context.py:229-289): apply_actions + process_actions
    (world/actions.py:44, world/systems/affordances.py:105), need decay, RelationshipService.decay
    (world/agents/relationships_service.py:165), lifecycle, episode_tick, reward, observe.  The fixture stores the
    action words, the per-tick reference outputs and the ledgers / positions of the final world.
    """
    from townlet.world.actions import apply_actions
    from townlet.world.systems.affordances import process_actions

    cases = (
        ("world_hybrid_48x8", "hybrid", 48, 8, 3, 8, {}, {}),
        ("world_compact_24x10", "compact", 24, 10, 2, 6, {"observations_config__compact__object_channels": ["stove", "bed"]}, {}),
        ("world_full_tie_decay", "full", 16, 6, 2, 7, {}, {"trust_decay": 0.04, "familiarity_decay": 0.03}),
        ("world_hybrid_128x64", "hybrid", 128, 64, 1, 3, {}, {}),
    )
    (GOLDEN_DIR / "world").mkdir(exist_ok=True)
    for name, variant, grid, agents, n_towns, n_ticks, extra, tie_params in cases:
        config = make_config(variant, employment__enforce_job_loop=False, **extra)
        spec = ObsSpec.from_config(config)
        rew = RewardSpec.from_config(config, fuse_faint=True)
        dyn = DynamicsSpec.from_config(config)
        for key, value in tie_params.items():
            setattr(dyn, key, value)
        rng = np.random.default_rng(77 + grid + agents)
        towns = [synth_town(t, grid, agents, terminated_frac=0.1) for t in range(n_towns)]
        for town in towns:  # edges that leave the ledgers during the sequence
            town.rivalry[0] = [(other, 0.052 + 0.004 * j) for j, (other, _) in enumerate(town.rivalry[0])]
            town.rivalry[1] = [(other, 0.05) for other, _ in town.rivalry[1]]  # all evicted at once -> rivalry features drop to 0
            first = town.ties[2][0][0]
            town.ties[2] = [(first, 0.0, 0.0, 0.0)] + town.ties[2][1:]  # a zero tie is evicted by the first decay
            second = town.ties[3][1][0]
            town.ties[3] = [town.ties[3][0], (second, 0.0, 0.0, 0.015)] + town.ties[3][2:]  # gone after two ticks
            town.ties[4] = [(other, 0.0, 0.0, 0.005) for other, _, _, _ in town.ties[4]]  # falls back to the rivalry ledger
        worlds = [materialise(t, config) for t in towns]
        for w, t in zip(worlds, towns):
            patch_employment(w, t)
            for ledger in w._relationships._relationship_ledgers.values():
                for key, value in tie_params.items():  # non-default RelationshipParameters (world/relationships.py:14-21)
                    setattr(ledger.params, key, value)
        services = [WorldObservationService(config=config) for _ in towns]
        engines = [RewardEngine(config) for _ in towns]
        lifecycles = [LifecycleManager(config) for _ in towns]

        def ingest_all():
            rows = []
            for w, t, eng in zip(worlds, towns, engines):
                term0 = {aid: bool(t.terminated[i]) for i, aid in enumerate(t.agent_ids)}
                reasons0 = {aid: "eviction" if i % 2 else "unemployment" for i, aid in enumerate(t.agent_ids) if t.terminated[i]}
                slots = {aid: w.embedding_allocator.allocate(aid, w.tick) for aid in w.agents}
                rows.append(
                    ingest_town(
                        w, spec, terminated=term0, reasons=reasons0, pending_resets=set(), slots=slots,
                        episode_totals=dict(eng._episode_totals), termination_block=dict(eng._termination_block), origin=(0, 0),
                    )
                )
            return rows

        rows = ingest_all()
        batch = rows_to_batch(rows, spec, grid=(grid, grid))
        batch.validate(len(spec.object_channels))
        N = batch.n_agents
        ref = {
            "map": np.zeros((n_ticks, N, *spec.map_shape), np.float32),
            "features": np.zeros((n_ticks, N, spec.n_features), np.float32),
            "summary": np.zeros((n_ticks, N, capi.TL_N_SUMMARY), np.int32),
            "comps": np.zeros((n_ticks, N, capi.TL_N_COMPS)),
            "terminated": np.zeros((n_ticks, N), np.uint8),
            "needs": np.zeros((n_ticks, 3, N)),
            "episode_total": np.zeros((n_ticks, N)),
        }
        words = np.zeros((n_ticks, N), np.uint32)
        for it in range(n_ticks):
            a = 0
            for ti, (w, t, svc, eng, life) in enumerate(zip(worlds, towns, services, engines, lifecycles)):
                n = len(t.agent_ids)
                w.tick += 1
                # ---- this tick's actions: moves (anywhere on the grid, also onto other agents), a move without a
                # position, wait / noop with durations, and agents without an action
                actions = {}
                for i, aid in enumerate(t.agent_ids):
                    roll = rng.random()
                    if roll < 0.45:
                        actions[aid] = {"kind": "move", "position": (int(rng.integers(0, grid)), int(rng.integers(0, grid)))}
                        if rng.random() < 0.3:
                            actions[aid]["duration"] = int(rng.integers(1, 9))
                    elif roll < 0.5:
                        actions[aid] = {"kind": "move"}
                    elif roll < 0.6:
                        actions[aid] = {"kind": "wait", "duration": int(rng.integers(1, 60))}
                    elif roll < 0.7:
                        actions[aid] = {"kind": "noop"}
                    elif roll < 0.75 and i > 0:  # lands on another agent's tile
                        other = w.agents[t.agent_ids[i - 1]]
                        actions[aid] = {"kind": "move", "position": tuple(other.position)}
                words[it, a : a + n] = encode_actions(t.agent_ids, actions)
                coerced = w.context._coerce_actions(actions) if hasattr(w, "context") and w.context is not None else None
                if coerced is None:
                    from townlet.world.actions import Action

                    coerced = [Action(agent_id=k, kind=str(v.get("kind", "noop")), payload=dict(v)) for k, v in actions.items()]
                if coerced:
                    apply_actions(w, coerced)  # emits action.invalid for the move without a position
                if actions:
                    process_actions(w, actions, tick=w.tick)
                w._apply_need_decay()
                w._relationships.decay()  # the "relationships" system step (world/systems/relationships.py:13-25)
                life_term = life.evaluate(w, w.tick)
                life_reasons = dict(life._termination_reasons)
                terminated, reasons = {}, {}
                for i, aid in enumerate(t.agent_ids):
                    host_flag = bool(t.terminated[i])
                    terminated[aid] = host_flag or bool(life_term.get(aid, False))
                    if host_flag:
                        reasons[aid] = "eviction" if i % 2 else "unemployment"
                    elif aid in life_reasons:
                        reasons[aid] = life_reasons[aid]
                tpd = config.observations_config.hybrid.time_ticks_per_day
                for snap in w.agents.values():
                    snap.episode_tick = (int(snap.episode_tick) + 1) % max(1, int(tpd))
                breakdowns = eng.compute(w, terminated, reasons)
                row, r_mean, r_var = reference_kpi_row(w, {aid: bd.total for aid, bd in breakdowns.items()}, terminated, config)
                ref["kpi"][it, ti], ref["kpi_reward_mean"][it, ti], ref["kpi_reward_variance"][it, ti] = row, r_mean, r_var
                ids, maps, feats, summ, _ = reference_obs(w, svc, terminated)
                assert ids == t.agent_ids
                ref["map"][it, a : a + n] = maps
                ref["features"][it, a : a + n] = feats
                ref["summary"][it, a : a + n] = summ
                for i, aid in enumerate(ids):
                    bd = breakdowns[aid]
                    row = ref["comps"][it, a + i]
                    for c, cname in enumerate(capi.COMP_NAMES):
                        if cname == "guardrail_active":
                            row[c] = float(bd.guardrail_active)
                        elif cname == "clipped":
                            row[c] = float(bd.clipped)
                        elif cname != "reserved":
                            row[c] = float(getattr(bd, cname))
                    ref["terminated"][it, a + i] = terminated[aid]
                    snap = w.agents[aid]
                    ref["needs"][it, :, a + i] = [snap.needs["hunger"], snap.needs["hygiene"], snap.needs["energy"]]
                    ref["episode_total"][it, a + i] = eng._episode_totals[aid]
                for aid in ids:
                    w.embedding_allocator._assignments.setdefault(aid, int(rows[ti].fields["slot"][ids.index(aid)]))
                a += n
        # ---- the final world, ingested again: ledgers, positions and spatial index after the sequence
        final = rows_to_batch(ingest_all(), spec, grid=(grid, grid))
        for fname in ("pos", "ent", "tie_count", "tie_trust", "tie_fam", "tie_riv", "tie_embed", "riv_count", "riv_score",
                      "riv_embed", "last_action_duration", "last_action_hash"):
            ref[f"final_{fname}"] = np.array(getattr(final, fname))
        ref["final_last_success"] = (np.array(final.flags) & capi.TL_F_LAST_SUCCESS).astype(np.uint32)
        save_fixture(
            GOLDEN_DIR / "world" / f"{name}.npz", batch, spec, rew, capi.TL_PHASE_WORLD, n_ticks, ref,
            note="actions -> need decay -> ledger decay -> lifecycle -> reward -> observe per tick, on reference WorldStates",
            dyn=dyn, actions=words,
        )
        evicted = int(batch.riv_count.sum()) - int(final.riv_count.sum()), int(batch.tie_count.sum()) - int(final.tie_count.sum())
        print(f"  world/{name}: {n_towns} towns x {agents} agents, {n_ticks} ticks; evicted rivalry/ties {evicted}")


def replay_case():
    """frames_to_replay_sample + build_batch (policy/replay.py:550-636, 245-276) on random per-agent frames.

    The frames carry what TrajectoryService._frames_from_envelope (policy/trajectory_service.py:91-140) puts in them:
    nested-list map / features, action_id, log_prob, value_pred, rewards / dones lists.
    """
    from townlet.policy.replay import build_batch, frames_to_replay_sample

    rng = np.random.default_rng(99)
    steps, agents, shape, feat = 6, 5, (4, 11, 11), 81
    tm = {
        "map": (rng.random((steps, agents, *shape)) < 0.05).astype(np.float32),
        "features": rng.normal(size=(steps, agents, feat)).astype(np.float32),
        "actions": rng.integers(0, 9, size=(steps, agents)).astype(np.int64),
        "log_probs": rng.normal(-2.0, 0.3, size=(steps, agents)).astype(np.float32),
        "values": rng.normal(0.0, 0.5, size=(steps + 1, agents)).astype(np.float32),
        "rewards": rng.uniform(-0.2, 0.2, size=(steps, agents)).astype(np.float32),
        "dones": (rng.random((steps, agents)) < 0.2).astype(np.float32),
    }
    samples = []
    for a in range(agents):
        frames = [
            {
                "map": tm["map"][t, a].tolist(),
                "features": tm["features"][t, a].tolist(),
                "action_id": int(tm["actions"][t, a]),
                "log_prob": float(tm["log_probs"][t, a]),
                "value_pred": float(tm["values"][t, a]),
                "rewards": [float(tm["rewards"][t, a])],
                "dones": [bool(tm["dones"][t, a])],
                "metadata": {},
            }
            for t in range(steps)
        ]
        samples.append(frames_to_replay_sample(frames))
    batch = build_batch(samples)
    (GOLDEN_DIR / "replay").mkdir(exist_ok=True)
    np.savez_compressed(
        GOLDEN_DIR / "replay" / "ref_replay.npz",
        **{f"in_{k}": v for k, v in tm.items()},
        maps=batch.maps, features=batch.features, actions=batch.actions, old_log_probs=batch.old_log_probs,
        value_preds=batch.value_preds, rewards=batch.rewards, dones=batch.dones,
    )
    print("  replay/ref_replay: frames_to_replay_sample + build_batch on 5 agents x 6 steps")


def fuzz_oracle(n_worlds: int = int(os.environ.get("TL_FUZZ_WORLDS", 240)), seed: int = int(os.environ.get("TL_FUZZ_SEED", 2026))) -> None:
    """Random reference worlds through the reference and through the C oracle, compared on the spot (nothing stored).

    Each world draws its own size, variant, window, social configuration, agent and object placement (stacked agents,
    agents at negative and far coordinates, objects sharing tiles, reservations, queues), needs near the thresholds,
    relationship and rivalry ledgers of every fill level, terminated agents and pending ctx resets; the reference runs
    build_batch + _apply_need_decay + lifecycle + reward for a few ticks, the oracle the same ticks on the ingested batch.
    Maps, features and integer summaries must be bit-identical, rewards within 1e-9.
    """
    import sys as _sys

    _sys.path.insert(0, str(REPO / "tests"))
    from helpers import assert_obs_equal, assert_reward_close
    from oracle.oracle import oracle_tick

    rng = np.random.default_rng(seed)
    checked_agents = 0
    for w_index in range(n_worlds):
        variant = ("hybrid", "full", "compact")[w_index % 3]
        extra = {}
        if variant == "compact":
            extra["observations_config__compact__map_window"] = int(rng.choice([5, 7, 9]))
            if rng.random() < 0.5:
                extra["observations_config__compact__object_channels"] = list(rng.choice(["stove", "bed", "fridge", "shower"], size=int(rng.integers(1, 4)), replace=False))
            extra["observations_config__compact__normalize_counts"] = bool(rng.random() < 0.5)
        else:
            extra["observations_config__hybrid__local_window"] = int(rng.choice([7, 9, 11, 13]))
            extra["observations_config__hybrid__include_targets"] = bool(rng.random() < 0.4)
        extra["observations_config__social_snippet__top_friends"] = int(rng.integers(0, 4))
        extra["observations_config__social_snippet__top_rivals"] = int(rng.integers(0, 4))
        extra["observations_config__social_snippet__include_aggregates"] = bool(rng.random() < 0.7)
        if extra["observations_config__social_snippet__top_friends"] + extra["observations_config__social_snippet__top_rivals"] == 0:
            extra["observations_config__social_snippet__include_aggregates"] = False
        config = make_config(variant, employment__enforce_job_loop=False, **extra)
        spec = ObsSpec.from_config(config)
        rew = RewardSpec.from_config(config, fuse_faint=True)
        grid = int(rng.choice([6, 12, 24, 48]))
        n_agents = int(rng.integers(1, 13))
        town = synth_town(int(rng.integers(0, 10_000)), grid, n_agents, terminated_frac=float(rng.choice([0.0, 0.2])))
        # crowd some agents onto shared tiles and push needs to the faint / starvation thresholds
        for i in range(n_agents):
            if rng.random() < 0.3:
                town.positions[i] = town.positions[int(rng.integers(0, n_agents))]
            if rng.random() < 0.2:
                town.needs[i, 0] = float(rng.choice([0.0, 0.03, 0.031, 0.04, 0.05, 0.051]))
            k_ties = int(rng.integers(0, min(6, n_agents - 1) + 1)) if n_agents > 1 else 0
            town.ties[i] = town.ties[i][:k_ties]
            town.rivalry[i] = town.rivalry[i][: int(rng.integers(0, len(town.rivalry[i]) + 1))]
        world = materialise(town, config)
        patch_employment(world, town)
        for aid in town.agent_ids:
            if rng.random() < 0.15:
                world.request_ctx_reset(aid)
        service = WorldObservationService(config=config)
        engine = RewardEngine(config)
        life = LifecycleManager(config)
        term0 = {aid: bool(town.terminated[i]) for i, aid in enumerate(town.agent_ids)}
        reasons0 = {aid: "eviction" if i % 2 else "unemployment" for i, aid in enumerate(town.agent_ids) if town.terminated[i]}
        pending = set(world._ctx_reset_requests)
        slots = {aid: world.embedding_allocator.allocate(aid, world.tick) for aid in world.agents}
        rows = ingest_town(world, spec, terminated=term0, reasons=reasons0, pending_resets=pending, slots=slots,
                           episode_totals={}, termination_block={}, origin=(0, 0))
        batch = rows_to_batch([rows], spec, grid=(grid, grid))
        batch.validate(len(spec.object_channels))
        n_ticks = int(rng.integers(1, 4))
        N = batch.n_agents
        ref = {"map": np.zeros((n_ticks, N, *spec.map_shape), np.float32), "features": np.zeros((n_ticks, N, spec.n_features), np.float32),
               "summary": np.zeros((n_ticks, N, capi.TL_N_SUMMARY), np.int32), "comps": np.zeros((n_ticks, N, capi.TL_N_COMPS)),
               "terminated": np.zeros((n_ticks, N), np.uint8)}
        for it in range(n_ticks):
            world.tick += 1
            world._apply_need_decay()
            life_term = life.evaluate(world, world.tick)
            life_reasons = dict(life._termination_reasons)
            terminated, reasons = {}, {}
            for i, aid in enumerate(town.agent_ids):
                host_flag = bool(town.terminated[i])
                terminated[aid] = host_flag or bool(life_term.get(aid, False))
                if host_flag:
                    reasons[aid] = "eviction" if i % 2 else "unemployment"
                elif aid in life_reasons:
                    reasons[aid] = life_reasons[aid]
            tpd = config.observations_config.hybrid.time_ticks_per_day
            for snap in world.agents.values():
                snap.episode_tick = (int(snap.episode_tick) + 1) % max(1, int(tpd))
            breakdowns = engine.compute(world, terminated, reasons)
            ids, maps, feats, summ, _ = reference_obs(world, service, terminated)
            ref["map"][it], ref["features"][it], ref["summary"][it] = maps, feats, summ
            for i, aid in enumerate(ids):
                bd = breakdowns[aid]
                for c, cname in enumerate(capi.COMP_NAMES):
                    if cname == "guardrail_active":
                        ref["comps"][it, i, c] = float(bd.guardrail_active)
                    elif cname == "clipped":
                        ref["comps"][it, i, c] = float(bd.clipped)
                    elif cname != "reserved":
                        ref["comps"][it, i, c] = float(getattr(bd, cname))
                ref["terminated"][it, i] = terminated[aid]
            for aid in ids:
                world.embedding_allocator._assignments.setdefault(aid, int(rows.fields["slot"][ids.index(aid)]))
        out = oracle_tick(batch, spec, rew, capi.TL_PHASE_ALL, n_ticks)
        what = f"fuzz world {w_index} ({variant}, grid {grid}, {n_agents} agents, {n_ticks} ticks)"
        assert_obs_equal(out, ref, what)
        assert_reward_close(out, ref, what)
        checked_agents += N * n_ticks
    print(f"  fuzz: {n_worlds} random worlds, {checked_agents} agent-ticks, oracle == reference (maps / features / summaries bit-exact)")


def fuzz_world(n_worlds: int = int(os.environ.get("TL_FUZZ_WORLDS", 120)), seed: int = int(os.environ.get("TL_FUZZ_SEED", 4242))) -> None:
    """Random worlds with random actions and ledger decay through the reference and the oracle (nothing stored).

    Per tick, in WorldContext.tick order: apply_actions + process_actions, need decay, RelationshipService.decay,
    lifecycle, reward, observe (as world_dynamics_cases); afterwards the final world is ingested again and its
    positions, entity words, last-action fields and ledgers are compared with the oracle's batch byte for byte.
    """
    import sys as _sys

    _sys.path.insert(0, str(REPO / "tests"))
    from helpers import assert_obs_equal, assert_reward_close
    from oracle.oracle import oracle_tick
    from townlet.world.actions import Action, apply_actions
    from townlet.world.systems.affordances import process_actions

    rng = np.random.default_rng(seed)
    agent_ticks = 0
    for w_index in range(n_worlds):
        variant = ("hybrid", "compact", "full")[w_index % 3]
        extra = {"conflict__rivalry__decay_per_tick": float(rng.choice([0.0, 0.005, 0.02, 0.05])),
                 "conflict__rivalry__eviction_threshold": float(rng.choice([0.0, 0.05, 0.1]))}
        config = make_config(variant, employment__enforce_job_loop=False, **extra)
        spec = ObsSpec.from_config(config)
        rew = RewardSpec.from_config(config, fuse_faint=True)
        dyn = DynamicsSpec.from_config(config)
        tie_params = {"trust_decay": float(rng.choice([0.0, 0.03, 0.2])), "familiarity_decay": float(rng.choice([0.0, 0.05])),
                      "rivalry_decay": float(rng.choice([0.0, 0.01, 0.1]))}
        dyn.trust_decay, dyn.familiarity_decay, dyn.tie_rivalry_decay = tie_params["trust_decay"], tie_params["familiarity_decay"], tie_params["rivalry_decay"]
        grid = int(rng.choice([8, 16, 32]))
        n_agents = int(rng.integers(2, 11))
        town = synth_town(int(rng.integers(0, 10_000)), grid, n_agents, terminated_frac=0.1)
        for i in range(n_agents):  # values that cross the eviction / zero thresholds within a few ticks
            town.rivalry[i] = [(o, float(rng.choice([0.04, 0.051, 0.06, 0.1, 0.3]))) for o, _ in town.rivalry[i]]
            town.ties[i] = [(o, float(rng.choice([0.0, 0.02, 0.3])), float(rng.choice([0.0, 0.04, 0.2])), float(rng.choice([0.0, 0.005, 0.15])))
                            for o, _, _, _ in town.ties[i]]
        world = materialise(town, config)
        patch_employment(world, town)
        for ledger in world._relationships._relationship_ledgers.values():
            for key, value in tie_params.items():
                setattr(ledger.params, key, value)
        service, engine, life = WorldObservationService(config=config), RewardEngine(config), LifecycleManager(config)

        def ingest():
            term0 = {aid: bool(town.terminated[i]) for i, aid in enumerate(town.agent_ids)}
            reasons0 = {aid: "eviction" if i % 2 else "unemployment" for i, aid in enumerate(town.agent_ids) if town.terminated[i]}
            slots = {aid: world.embedding_allocator.allocate(aid, world.tick) for aid in world.agents}
            return ingest_town(world, spec, terminated=term0, reasons=reasons0, pending_resets=set(), slots=slots,
                               episode_totals=dict(engine._episode_totals), termination_block=dict(engine._termination_block), origin=(0, 0))

        rows = ingest()
        batch = rows_to_batch([rows], spec, grid=(grid, grid))
        n_ticks = int(rng.integers(1, 7))
        N = batch.n_agents
        ref = {"map": np.zeros((n_ticks, N, *spec.map_shape), np.float32), "features": np.zeros((n_ticks, N, spec.n_features), np.float32),
               "summary": np.zeros((n_ticks, N, capi.TL_N_SUMMARY), np.int32), "comps": np.zeros((n_ticks, N, capi.TL_N_COMPS)),
               "terminated": np.zeros((n_ticks, N), np.uint8)}
        words = np.zeros((n_ticks, N), np.uint32)
        for it in range(n_ticks):
            world.tick += 1
            actions = {}
            for i, aid in enumerate(town.agent_ids):
                roll = rng.random()
                if roll < 0.5:
                    actions[aid] = {"kind": "move", "position": (int(rng.integers(0, grid)), int(rng.integers(0, grid))), "duration": int(rng.integers(1, 20))}
                elif roll < 0.55:
                    actions[aid] = {"kind": "move"}
                elif roll < 0.7:
                    actions[aid] = {"kind": str(rng.choice(["wait", "noop"])), "duration": int(rng.integers(1, 60))}
            words[it] = encode_actions(town.agent_ids, actions)
            coerced = [Action(agent_id=k, kind=str(v.get("kind", "noop")), payload=dict(v)) for k, v in actions.items()]
            if coerced:
                apply_actions(world, coerced)
            if actions:
                process_actions(world, actions, tick=world.tick)
            world._apply_need_decay()
            world._relationships.decay()
            life_term = life.evaluate(world, world.tick)
            life_reasons = dict(life._termination_reasons)
            terminated, reasons = {}, {}
            for i, aid in enumerate(town.agent_ids):
                host_flag = bool(town.terminated[i])
                terminated[aid] = host_flag or bool(life_term.get(aid, False))
                if host_flag:
                    reasons[aid] = "eviction" if i % 2 else "unemployment"
                elif aid in life_reasons:
                    reasons[aid] = life_reasons[aid]
            tpd = config.observations_config.hybrid.time_ticks_per_day
            for snap in world.agents.values():
                snap.episode_tick = (int(snap.episode_tick) + 1) % max(1, int(tpd))
            breakdowns = engine.compute(world, terminated, reasons)
            ids, maps, feats, summ, _ = reference_obs(world, service, terminated)
            ref["map"][it], ref["features"][it], ref["summary"][it] = maps, feats, summ
            for i, aid in enumerate(ids):
                bd = breakdowns[aid]
                for c, cname in enumerate(capi.COMP_NAMES):
                    if cname == "guardrail_active":
                        ref["comps"][it, i, c] = float(bd.guardrail_active)
                    elif cname == "clipped":
                        ref["comps"][it, i, c] = float(bd.clipped)
                    elif cname != "reserved":
                        ref["comps"][it, i, c] = float(getattr(bd, cname))
                ref["terminated"][it, i] = terminated[aid]
            for aid in ids:
                world.embedding_allocator._assignments.setdefault(aid, int(rows.fields["slot"][ids.index(aid)]))
        final = rows_to_batch([ingest()], spec, grid=(grid, grid))
        out = oracle_tick(batch, spec, rew, capi.TL_PHASE_WORLD, n_ticks, dyn=dyn, actions=words)
        what = f"fuzz world {w_index} ({variant}, grid {grid}, {n_agents} agents, {n_ticks} ticks, {tie_params}, {extra})"
        assert_obs_equal(out, ref, what)
        assert_reward_close(out, ref, what)
        for fname in ("pos", "ent", "tie_count", "tie_trust", "tie_fam", "tie_riv", "tie_embed", "riv_count", "riv_score", "riv_embed",
                      "last_action_duration", "last_action_hash"):
            got, want = np.asarray(getattr(batch, fname)), np.asarray(getattr(final, fname))
            assert np.array_equal(got.view(np.uint8), want.view(np.uint8)), f"{what}: final {fname}"
        agent_ticks += N * n_ticks
    print(f"  fuzz-world: {n_worlds} random worlds, {agent_ticks} agent-ticks with actions and ledger decay, oracle == reference")


def fuzz_dropin(n_worlds: int = int(os.environ.get("TL_FUZZ_WORLDS", 150)), seed: int = int(os.environ.get("TL_FUZZ_SEED", 515))) -> None:
    """The drop-in classes against the reference classes on random worlds (host logic: the compute underneath is the
    oracle through the test-only engine): build_batch dictionaries including the metadata, RewardEngine.compute
    breakdowns and the need-decay hook over a few ticks of twin worlds built from the same description."""
    import sys as _sys

    _sys.path.insert(0, str(REPO / "tests"))
    import townlet_b200.observation_service as obs_mod
    import townlet_b200.reward_engine as rew_mod
    from oracle_engine import OracleHostEngine

    if os.environ.get("TL_DROPIN_ENGINE", "oracle") != "cuda":  # "cuda": the real HostEngine (GPU box, composed test)
        obs_mod.HostEngine = OracleHostEngine
        rew_mod.HostEngine = OracleHostEngine
    rng = np.random.default_rng(seed)
    compared = 0
    for w_index in range(n_worlds):
        variant = ("hybrid", "full", "compact")[w_index % 3]
        extra = {"observations_config__social_snippet__top_friends": int(rng.integers(0, 4)),
                 "observations_config__social_snippet__top_rivals": int(rng.integers(1, 4)),
                 "observations_config__hybrid__include_targets": bool(rng.random() < 0.4),
                 "features__observations__personality_channels": bool(rng.random() < 0.3)}
        if variant == "compact" and rng.random() < 0.5:
            extra["observations_config__compact__object_channels"] = ["stove", "bed"]
        config = make_config(variant, employment__enforce_job_loop=False, **extra)
        grid = int(rng.choice([6, 16, 48]))
        n_agents = int(rng.integers(1, 10))
        town = synth_town(int(rng.integers(0, 10_000)), grid, n_agents, terminated_frac=0.0)
        for i in range(n_agents):
            if rng.random() < 0.3:
                town.positions[i] = town.positions[int(rng.integers(0, n_agents))]
            if rng.random() < 0.2:
                town.needs[i, 0] = float(rng.choice([0.0, 0.03, 0.031, 0.05]))
        offset = (int(rng.integers(-40, 40)), int(rng.integers(-40, 40)))  # the reference has no coordinate bounds
        town.positions = town.positions + np.array(offset)
        town.object_positions = town.object_positions + np.array(offset)
        twins = []
        for _ in range(2):
            world = materialise(town, config)
            patch_employment(world, town)
            twins.append(world)
        ref_world, our_world = twins
        resets = [aid for aid in town.agent_ids if rng.random() < 0.2]
        for world in twins:
            for aid in resets:
                world.request_ctx_reset(aid)
        ref_service, our_service = WorldObservationService(config=config), obs_mod.B200ObservationService(config=config)
        ref_engine, our_engine = RewardEngine(config), rew_mod.B200RewardEngine(config)
        rew_mod.install_need_decay(our_world)
        life_a, life_b = LifecycleManager(config), LifecycleManager(config)
        for it in range(int(rng.integers(1, 4))):
            terminated = []
            for world, life in ((ref_world, life_a), (our_world, life_b)):
                world.tick += 1
                world._apply_need_decay()
                terminated.append(dict(life.evaluate(world, world.tick)))
            assert terminated[0] == terminated[1], f"fuzz-dropin world {w_index}: lifecycle after decay"
            for aid in town.agent_ids:
                assert ref_world.agents[aid].needs == our_world.agents[aid].needs, f"fuzz-dropin world {w_index}: needs of {aid}"
            reasons = {aid: "faint" for aid, flag in terminated[0].items() if flag}
            # social events of the tick (rewards/engine.py:257-317): chats that succeed or fail between random pairs
            # (also with an id the world does not know), avoidance blocks; recorded identically in both worlds
            events = []
            for _ in range(int(rng.integers(0, 5))):
                kind = str(rng.choice(["chat_success", "chat_failure", "avoid"]))
                speaker = str(rng.choice(town.agent_ids))
                listener = str(rng.choice(town.agent_ids + ["stranger"]))
                events.append((kind, speaker, listener, float(rng.choice([0.0, 0.25, 0.5, 1.0, 2.0]))))
            for world in twins:
                for kind, speaker, listener, quality in events:
                    if kind == "chat_success":
                        world.record_chat_success(speaker, listener, quality)
                    elif kind == "chat_failure":
                        world.record_chat_failure(speaker, listener)
                    else:
                        world.record_relationship_guard_block(agent_id=speaker, reason="queue_rival", target_agent=listener, object_id="o")
            bd_ref = ref_engine.compute(ref_world, terminated[0], reasons)
            bd_our = our_engine.compute(our_world, terminated[1], reasons)
            assert list(bd_ref) == list(bd_our)
            for aid in bd_ref:
                a, b = bd_ref[aid].model_dump(), bd_our[aid].model_dump()
                for key, value in a.items():
                    if isinstance(value, float):
                        assert abs(value - b[key]) <= 1e-9 + 1e-6 * abs(value), f"fuzz-dropin world {w_index}: reward {key} of {aid}: {value} vs {b[key]}"
                    else:
                        assert value == b[key], f"fuzz-dropin world {w_index}: reward {key} of {aid}"
            want = ref_service.build_batch(ref_world, terminated[0])
            got = our_service.build_batch(our_world, terminated[1])
            assert list(want) == list(got), f"fuzz-dropin world {w_index}: agent order"
            for aid in want:
                assert np.array_equal(got[aid]["map"], want[aid]["map"]), f"fuzz-dropin world {w_index} tick {it}: map of {aid}"
                assert np.array_equal(got[aid]["features"].view(np.uint32), want[aid]["features"].view(np.uint32)), (
                    f"fuzz-dropin world {w_index} tick {it} ({variant}): features of {aid}")
                assert got[aid]["metadata"] == want[aid]["metadata"], f"fuzz-dropin world {w_index} tick {it}: metadata of {aid}"
                compared += 1
    print(f"  fuzz-dropin: {n_worlds} twin worlds, {compared} observations with metadata, rewards and decayed needs equal to the reference's")


def fuzz_gae(n_cases: int = int(os.environ.get("TL_FUZZ_WORLDS", 200)), seed: int = int(os.environ.get("TL_FUZZ_SEED", 77))) -> None:
    """compute_gae (policy/backends/pytorch/ppo_utils.py:19-67) against the oracle's scan on random rollouts."""
    import torch

    from oracle.oracle import oracle_gae
    from townlet.policy.backends.pytorch.ppo_utils import compute_gae

    rng = np.random.default_rng(seed)
    for case in range(n_cases):
        batch, steps, bootstrap = int(rng.integers(1, 40)), int(rng.integers(1, 70)), bool(rng.random() < 0.5)
        gamma, lam = float(rng.choice([0.9, 0.99, 0.999, 1.0])), float(rng.choice([0.0, 0.5, 0.95, 1.0]))
        rewards = rng.uniform(-1.0, 1.0, size=(batch, steps)).astype(np.float32)
        values = rng.normal(0.0, 1.0, size=(batch, steps + int(bootstrap))).astype(np.float32)
        dones = (rng.random((batch, steps)) < rng.choice([0.0, 0.1, 0.5])).astype(np.float32)
        res = compute_gae(torch.from_numpy(rewards), torch.from_numpy(values), torch.from_numpy(dones), gamma=gamma, gae_lambda=lam)
        adv, ret = oracle_gae(rewards.T, values.T, dones.T, gamma, lam)
        assert np.array_equal(adv.T, res.advantages.numpy()) and np.array_equal(ret.T, res.returns.numpy()), (case, batch, steps, bootstrap, gamma, lam)
    print(f"  fuzz-gae: {n_cases} random rollouts, oracle == compute_gae bit for bit")


def gae_case():
    """compute_gae (policy/backends/pytorch/ppo_utils.py:19-67) on random float32 rollouts, both value layouts."""
    import torch

    from townlet.policy.backends.pytorch.ppo_utils import compute_gae

    rng = np.random.default_rng(4321)
    out = {}
    for name, (batch, steps, bootstrap) in {"boot": (37, 19, True), "noboot": (5, 64, False), "one": (3, 1, True)}.items():
        rewards = rng.uniform(-0.2, 0.2, size=(batch, steps)).astype(np.float32)
        values = rng.normal(0.0, 0.5, size=(batch, steps + (1 if bootstrap else 0))).astype(np.float32)
        dones = (rng.random((batch, steps)) < 0.15).astype(np.float32)
        res = compute_gae(torch.from_numpy(rewards), torch.from_numpy(values), torch.from_numpy(dones), gamma=0.99, gae_lambda=0.95)
        for key, arr in (("rewards", rewards), ("values", values), ("dones", dones), ("advantages", res.advantages.numpy()), ("returns", res.returns.numpy())):
            out[f"{name}_{key}"] = arr
    (GOLDEN_DIR / "gae").mkdir(exist_ok=True)
    np.savez_compressed(GOLDEN_DIR / "gae" / "ref_gae.npz", gamma=0.99, gae_lambda=0.95, **out)
    print("  gae/ref_gae: compute_gae on 3 random rollouts (batch-major arrays as the reference takes them)")


def main():
    print(f"reference: {REF_SRC} (copy at {WORK})")
    only = os.environ.get("TL_GOLDEN_ONLY")
    if only:
        {"synth": synth_cases, "reservations": reservation_cases, "employment": employment_case, "gae": gae_case, "snapshots": snapshot_cases, "world": world_dynamics_cases, "replay": replay_case, "fuzz": fuzz_oracle, "fuzz-world": fuzz_world, "fuzz-dropin": fuzz_dropin, "fuzz-gae": fuzz_gae}[only]()
        shutil.rmtree(WORK, ignore_errors=True)
        return
    baselines()
    social_snippet()
    builder_cases()
    crowded_cases()
    reservation_cases()
    employment_case()
    synth_cases()
    reward_cases()
    snapshot_cases()
    world_dynamics_cases()
    replay_case()
    gae_case()
    shutil.rmtree(WORK, ignore_errors=True)


if __name__ == "__main__":
    main()
