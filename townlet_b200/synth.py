"""Synthetic town generator (SURVEY.md §8d "Synthetic inputs").

Town ``t`` is drawn from ``numpy.random.default_rng(seed + t)`` with
``seed = 1234``: a G×G grid, A agents on distinct tiles, ``max(4, G // 4)``
objects cycling fridge/stove/bed/shower on distinct free tiles, 25 % of the
objects reserved by a random agent, three relationship ties per agent to the
next three agents (cyclically) with mirrored rivalry edges, 2 % of agents
terminated with reason ``faint``.

``synth_town`` returns a plain description (ids as strings, Python-level
values) that both ``assemble`` (→ ``TownBatch``) and the golden-vector script
(→ a reference ``WorldState``) consume, so the ingest can be cross-checked.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import capi
from .hashing import id_embedding_words, last_action_hash
from .params import LTYPE_CODE, PROFILE_CODE, SHIFT_CODE, ObsSpec, pack_actions
from .soa import BatchLayout, TownBatch, pack_entity, pack_pos

OBJECT_TYPES = ("fridge", "stove", "bed", "shower")
# the seven shift states that occur (world/employment.py:170,237-243,314,434,446)
SHIFT_STATES7 = ("pre_shift", "idle", "await_start", "on_time", "late", "absent", "post_shift")
# world/actions.py:13-16 + world/systems/affordances.py:217-229, plus "" (never acted)
ACTION_KINDS = ("move", "request", "start", "release", "chat", "blocked", "noop", "wait", "")
PROFILES = ("balanced", "socialite", "industrious", "stoic")


@dataclass
class TownDescription:
    """One town in reference vocabulary (strings, tuples, float64)."""

    grid: int
    tick: int
    agent_ids: list[str]
    positions: np.ndarray  # [A,2] int
    needs: np.ndarray  # [A,3] f64
    wallet: np.ndarray
    attendance: np.ndarray
    wages_withheld: np.ndarray
    wages_paid: np.ndarray
    punctuality: np.ndarray
    lateness: np.ndarray
    last_action_duration: np.ndarray
    last_action_success: np.ndarray
    last_action_id: list[str]
    episode_tick: np.ndarray
    on_shift: np.ndarray
    shift_state: list[str]
    in_queue: np.ndarray
    terminated: np.ndarray
    profile: list[str]
    personality: np.ndarray  # [A,3] f64
    object_ids: list[str]
    object_types: list[str]
    object_positions: np.ndarray  # [O,2]
    reservations: dict[str, str]  # object_id -> agent_id
    ties: list[list[tuple[str, float, float, float]]] = field(default_factory=list)  # per agent, insertion order
    rivalry: list[list[tuple[str, float]]] = field(default_factory=list)  # per agent, insertion order


def synth_town(town_index: int, grid: int, n_agents: int, seed: int = 1234, terminated_frac: float = 0.02) -> TownDescription:
    rng = np.random.default_rng(seed + town_index)
    n_objects = max(4, grid // 4)
    tiles = rng.choice(grid * grid, size=n_agents + n_objects, replace=False)
    positions = np.stack([tiles[:n_agents] % grid, tiles[:n_agents] // grid], axis=1).astype(np.int64)
    obj_tiles = tiles[n_agents:]
    object_positions = np.stack([obj_tiles % grid, obj_tiles // grid], axis=1).astype(np.int64)
    agent_ids = [f"agent_{i}" for i in range(n_agents)]
    object_types = [OBJECT_TYPES[i % 4] for i in range(n_objects)]
    object_ids = [f"{object_types[i]}_{i}" for i in range(n_objects)]
    reserved_mask = rng.random(n_objects) < 0.25
    holders = rng.integers(0, n_agents, size=n_objects)
    reservations = {object_ids[i]: agent_ids[int(holders[i])] for i in range(n_objects) if reserved_mask[i]}
    needs = rng.uniform(0.05, 1.0, size=(n_agents, 3))
    wallet = rng.uniform(0.0, 5.0, size=n_agents)
    tick = int(rng.integers(0, 1440))
    shift_idx = rng.integers(0, len(SHIFT_STATES7), size=n_agents)
    action_idx = rng.integers(0, len(ACTION_KINDS), size=n_agents)
    attendance = rng.uniform(0.0, 1.0, size=n_agents)
    wages_withheld = rng.uniform(0.0, 0.05, size=n_agents) * (rng.random(n_agents) < 0.3)
    wages_paid = rng.uniform(0.0, 0.03, size=n_agents) * (rng.random(n_agents) < 0.5)
    punctuality = rng.integers(0, 1001, size=n_agents) / 1000.0  # on_time_ticks / scheduled_ticks
    lateness = rng.integers(0, 4, size=n_agents)
    duration = rng.integers(0, 12, size=n_agents)
    success = rng.random(n_agents) < 0.7
    on_shift = rng.random(n_agents) < 0.3
    in_queue = rng.random(n_agents) < 0.1
    terminated = rng.random(n_agents) < terminated_frac
    profile_idx = rng.integers(0, len(PROFILES), size=n_agents)
    personality = rng.uniform(-1.0, 1.0, size=(n_agents, 3))
    n_ties = min(3, n_agents - 1)
    tie_vals = rng.uniform(0.0, 1.0, size=(n_agents, 3, 3))
    ties: list[list[tuple[str, float, float, float]]] = []
    rivalry: list[list[tuple[str, float]]] = []
    for i in range(n_agents):
        row = []
        riv_row = []
        for j in range(n_ties):
            other = agent_ids[(i + 1 + j) % n_agents]
            trust = float(tie_vals[i, j, 0] * 0.3)
            fam = float(tie_vals[i, j, 1] * 0.3)
            riv = float(tie_vals[i, j, 2] * 0.2)
            row.append((other, trust, fam, riv))
            riv_row.append((other, riv))
        ties.append(row)
        rivalry.append(riv_row)
    return TownDescription(
        grid=grid,
        tick=tick,
        agent_ids=agent_ids,
        positions=positions,
        needs=needs,
        wallet=wallet,
        attendance=attendance,
        wages_withheld=wages_withheld,
        wages_paid=wages_paid,
        punctuality=punctuality,
        lateness=lateness,
        last_action_duration=duration,
        last_action_success=success,
        last_action_id=[ACTION_KINDS[int(k)] for k in action_idx],
        episode_tick=np.full(n_agents, tick, dtype=np.int64),
        on_shift=on_shift,
        shift_state=[SHIFT_STATES7[int(k)] for k in shift_idx],
        in_queue=in_queue,
        terminated=terminated,
        profile=[PROFILES[int(k)] for k in profile_idx],
        personality=personality,
        object_ids=object_ids,
        object_types=object_types,
        object_positions=object_positions,
        reservations=reservations,
        ties=ties,
        rivalry=rivalry,
    )


def _flags(town: TownDescription, holders: set[str]) -> np.ndarray:
    n = len(town.agent_ids)
    flags = np.zeros(n, dtype=np.uint32)
    for i in range(n):
        f = 0
        if town.on_shift[i]:
            f |= capi.TL_F_ON_SHIFT
        f |= SHIFT_CODE.get(town.shift_state[i], 0) << capi.TL_F_SHIFT_SHIFT
        if town.last_action_success[i]:
            f |= capi.TL_F_LAST_SUCCESS
        if town.in_queue[i]:
            f |= capi.TL_F_IN_QUEUE
        if town.agent_ids[i] in holders:
            f |= capi.TL_F_RESERVATION
        if town.terminated[i]:
            f |= capi.TL_F_TERMINATED | (1 << capi.TL_F_REASON_SHIFT)
        f |= PROFILE_CODE.get(town.profile[i], 0) << capi.TL_F_PROFILE_SHIFT
        flags[i] = f
    return flags


def assemble(towns: list[TownDescription], spec: ObsSpec | None = None, max_edges: int = 6,
             device_reservations: bool = False) -> TownBatch:
    """Pack town descriptions into one ``TownBatch`` (what the ingest produces from reference worlds).

    ``device_reservations``: one reservation slot per object with its holder in ``ent_holder`` (the layout of
    ``ingest_town(..., device_reservations=True)``, for ``TL_PHASE_RESERVATIONS``) instead of plain reserved tiles."""
    spec = spec or ObsSpec()
    ew = spec.embed_words
    n_agents = sum(len(t.agent_ids) for t in towns)
    n_ent = sum(len(t.agent_ids) + len(t.object_ids) + (len(t.object_ids) if device_reservations else len(t.reservations)) for t in towns)
    grid = max(t.grid for t in towns)
    layout = BatchLayout.make(len(towns), n_agents, n_ent, max_edges, ew)
    b = TownBatch(layout, grid, grid)
    ctype_of = {name: i + 1 for i, name in enumerate(spec.object_channels)} if spec.variant == "compact" else {}
    a = 0
    e = 0
    b.town_agent_off[0] = 0
    b.town_ent_off[0] = 0
    for ti, t in enumerate(towns):
        n = len(t.agent_ids)
        sl = slice(a, a + n)
        b.town_tick[ti] = t.tick
        b.pos[sl] = pack_pos(t.positions[:, 0], t.positions[:, 1])
        b.needs[:, sl] = t.needs.T
        b.wallet[sl] = t.wallet
        b.attendance[sl] = t.attendance
        b.wages_withheld[sl] = t.wages_withheld
        b.wages_paid[sl] = t.wages_paid
        b.punctuality[sl] = t.punctuality
        b.lateness[sl] = t.lateness
        b.last_action_duration[sl] = t.last_action_duration
        b.episode_tick[sl] = t.episode_tick
        b.slot[sl] = np.arange(n)  # allocation order of a fresh allocator (embedding.py:130-139)
        b.flags[sl] = _flags(t, set(t.reservations.values()))
        b.last_action_hash[sl] = np.array([last_action_hash(k) for k in t.last_action_id], dtype=np.float64)
        b.personality[:, sl] = t.personality.T
        index_of = {aid: i for i, aid in enumerate(t.agent_ids)}
        for i in range(n):
            row = t.ties[i]
            b.tie_count[a + i] = len(row)
            for j, (other, trust, fam, riv) in enumerate(row):
                b.tie_embed[a + i, j] = id_embedding_words(other, ew)
                b.tie_trust[a + i, j] = trust
                b.tie_fam[a + i, j] = fam
                b.tie_riv[a + i, j] = riv
            # rivalry_top: stable descending sort of the ledger (rivalry.py:101-110)
            top = sorted(t.rivalry[i], key=lambda item: item[1], reverse=True)[:max_edges]
            b.riv_count[a + i] = len(top)
            for j, (other, score) in enumerate(top):
                b.riv_embed[a + i, j] = id_embedding_words(other, ew)
                b.riv_score[a + i, j] = score
        del index_of
        # entities: agents, objects, reserved tiles
        b.ent[e : e + n] = pack_entity(t.positions[:, 0], t.positions[:, 1], capi.TL_KIND_AGENT)
        e += n
        no = len(t.object_ids)
        if no:
            ltype = np.array([LTYPE_CODE.get(tp, 0) for tp in t.object_types])
            ctype = np.array([ctype_of.get(str(tp).strip().lower(), 0) for tp in t.object_types])
            b.ent[e : e + no] = pack_entity(
                t.object_positions[:, 0], t.object_positions[:, 1], capi.TL_KIND_OBJECT, ltype, ctype, 1, 1
            )
            e += no
        if device_reservations and no:
            held = np.array([oid in t.reservations for oid in t.object_ids])
            agent_index = {aid: i for i, aid in enumerate(t.agent_ids)}
            b.ent[e : e + no] = pack_entity(t.object_positions[:, 0], t.object_positions[:, 1],
                                            np.where(held, capi.TL_KIND_RESERVED, capi.TL_KIND_EMPTY))
            b.ent_holder[e : e + no] = [agent_index[t.reservations[oid]] if oid in t.reservations else -1 for oid in t.object_ids]
            e += no
        elif t.reservations:
            obj_index = {oid: i for i, oid in enumerate(t.object_ids)}
            idx = np.array([obj_index[oid] for oid in t.reservations])
            idx = idx[np.lexsort((t.object_positions[idx, 0], t.object_positions[idx, 1]))]  # (y, x) order
            b.ent[e : e + len(idx)] = pack_entity(
                t.object_positions[idx, 0], t.object_positions[idx, 1], capi.TL_KIND_RESERVED
            )
            e += len(idx)
        a += n
        b.town_agent_off[ti + 1] = a
        b.town_ent_off[ti + 1] = e
    b.social_in[...] = 0.0
    return b


def synth_batch(
    n_towns: int,
    grid: int,
    n_agents: int,
    spec: ObsSpec | None = None,
    seed: int = 1234,
    first_town: int = 0,
    terminated_frac: float = 0.02,
    device_reservations: bool = False,
) -> TownBatch:
    """Synthetic batch of ``n_towns`` towns (town indices ``first_town`` ...)."""
    towns = [synth_town(first_town + t, grid, n_agents, seed, terminated_frac) for t in range(n_towns)]
    return assemble(towns, spec, device_reservations=device_reservations)


def tile_batch(base: TownBatch, repeats: int) -> TownBatch:
    """Repeat the towns of ``base`` ``repeats`` times (fast construction of very large batches).

    Used for throughput runs where generating 10^5 distinct towns on the host
    would dominate start-up; each copy gets its town ticks shifted so copies
    are not byte-identical.
    """
    lay = base.layout
    layout = BatchLayout.make(lay.n_towns * repeats, lay.n_agents * repeats, lay.n_entities * repeats, lay.max_edges, lay.embed_words)
    out = TownBatch(layout, base.grid_w, base.grid_h)
    out.has_social_in = base.has_social_in
    out.has_personality = base.has_personality
    for name, src in base.views().items():
        dst = getattr(out, name)
        if name == "town_agent_off":
            dst[0] = 0
            dst[1:] = (np.tile(src[1:], repeats) + np.repeat(np.arange(repeats) * lay.n_agents, lay.n_towns)).astype(np.int32)
        elif name == "town_ent_off":
            dst[0] = 0
            dst[1:] = (np.tile(src[1:], repeats) + np.repeat(np.arange(repeats) * lay.n_entities, lay.n_towns)).astype(np.int32)
        elif name == "town_tick":
            dst[...] = (np.tile(src, repeats) + np.repeat(np.arange(repeats) * 7, lay.n_towns)) % 1440
        elif name in ("needs", "social_in", "personality"):
            dst[...] = np.tile(src, (1, repeats))
        else:
            dst[...] = np.tile(src, (repeats,) + (1,) * (src.ndim - 1))
    return out


def random_actions(batch: TownBatch, n_ticks: int, seed: int) -> np.ndarray:
    """Seeded action words: moves anywhere on the grid, moves without a position, waits, idle agents."""
    rng = np.random.default_rng(seed)
    n = batch.n_agents
    kind = rng.choice([0, 1, 1, 1, 2, 3, 7], size=(n_ticks, n))
    has_pos = (kind == 1) & (rng.random((n_ticks, n)) < 0.9)
    x = rng.integers(0, batch.grid_w, size=(n_ticks, n))
    y = rng.integers(0, batch.grid_h, size=(n_ticks, n))
    success = rng.integers(0, 2, size=(n_ticks, n))
    duration = rng.integers(0, 64, size=(n_ticks, n))
    return pack_actions(kind, success, has_pos, x, y, duration)
