"""Ingest: reference world (adapter protocol) → structure-of-arrays rows.

Reads ONLY the read-only surface the reference's own observation / reward
code reads — ``WorldRuntimeAdapterProtocol``
(src/townlet/world/observations/interfaces.py:43-107) — in the same order, so
iteration-order dependent results (agent order, object insertion order, tie
insertion order) are preserved (SURVEY.md §7 "Order-dependent semantics").

Strings become integers here: ids → row indices and blake2s digests,
object types → landmark / channel codes, shift states and termination
reasons → small codes (SURVEY.md §7 "Strings → integers").
"""

from __future__ import annotations

from collections.abc import Mapping
from dataclasses import dataclass, field
from typing import Any

import numpy as np

from . import capi
from .hashing import id_embedding_words, last_action_hash
from .params import LTYPE_CODE, NEEDS, PROFILE_CODE, REASON_CODE, SHIFT_CODE, ObsSpec
from .soa import BatchLayout, TownBatch, pack_entity, pack_pos


def _as_float(value: object, default: float = 0.0) -> float:
    """A float64 SoA cell from whatever the snapshot holds (same coercion rules as the reference's context builder,
    world/observations/context.py:12-31: numbers and numeric strings convert, anything else yields the default)."""
    if isinstance(value, (int, float)):  # bool included
        return float(value)
    try:
        return float(value)  # type: ignore[arg-type]
    except (TypeError, ValueError):
        return default


def _as_int(value: object, default: int = 0) -> int:
    """An int32 SoA cell: ints as they are, bools as 0 / 1, floats truncated, integer strings parsed; otherwise the default."""
    if isinstance(value, (bool, int, float)):
        return int(value)
    if not isinstance(value, str):
        return default
    try:
        return int(value)
    except ValueError:
        return default


def resolve_adapter(world: Any) -> Any:
    """Normalise WorldState / WorldContext / adapter like ``ensure_world_adapter``
    (world/core/runtime_adapter.py:192-213) when the reference is importable;
    otherwise accept any object that already quacks like the adapter."""
    if hasattr(world, "agent_snapshots_view") and hasattr(world, "_employment_context_wages"):
        return world  # already an adapter (or a duck-typed stand-in)
    try:
        from townlet.world.core.runtime_adapter import ensure_world_adapter  # type: ignore

    except ImportError:
        return world
    try:
        return ensure_world_adapter(world)
    except TypeError:  # neither a WorldState nor a context: a duck-typed stand-in with the accessors the ingest reads
        return world


@dataclass
class TownRows:
    """One town's rows, still on the host, origin-shifted."""

    agent_ids: list[str]
    origin: tuple[int, int]
    extent: tuple[int, int]
    tick: int
    ent: np.ndarray
    fields: dict[str, np.ndarray] = field(default_factory=dict)
    ent_holder: np.ndarray | None = None  # per entity: holder of a reservation slot, -1 free, -2 not a slot


def ingest_town(
    world: Any,
    spec: ObsSpec,
    *,
    terminated: Mapping[str, bool] | None = None,
    reasons: Mapping[str, str] | None = None,
    pending_resets: set[str] | None = None,
    slots: Mapping[str, int] | None = None,
    episode_totals: Mapping[str, float] | None = None,
    termination_block: Mapping[str, int] | None = None,
    social_sums: tuple[Mapping[str, float], Mapping[str, float], Mapping[str, float]] | None = None,
    max_edges: int = 6,
    observe: bool = True,
    origin: tuple[int, int] | None = None,
    device_reservations: bool = False,
) -> TownRows:
    """Read one reference world into SoA rows.

    ``observe=False`` skips the observation-only accessors (objects, ties,
    queues) for callers that only run decay / reward.  ``origin`` pins the
    town origin (e.g. ``(0, 0)`` for a declared grid); by default it is the
    minimum corner of everything that occupies a tile.

    ``device_reservations=True`` lays the reserved tiles out as reservation SLOTS for ``TL_PHASE_RESERVATIONS``
    (world/systems/queues.py:55-86 on the device): one slot per object that has a position, in object order, holding
    the tile's current state, plus ``ent_holder`` — the town-local index of the agent the queue manager has granted the
    object to (``QueueManager.active_agent``), -1 while it is free.  Reserved tiles no object accounts for stay plain
    reserved words.
    """
    adapter = resolve_adapter(world)
    terminated = terminated or {}
    reasons = reasons or {}
    pending_resets = pending_resets or set()
    snapshots = dict(adapter.agent_snapshots_view())  # service.py:209
    agent_ids = list(snapshots)
    n = len(agent_ids)
    ew = spec.embed_words
    K = max_edges

    centre = np.zeros((n, 2), dtype=np.int64)
    occ = np.zeros((n, 2), dtype=np.int64)
    f: dict[str, np.ndarray] = {
        "needs": np.zeros((3, n)),
        "wallet": np.zeros(n),
        "attendance": np.zeros(n),
        "wages_withheld": np.zeros(n),
        "wages_paid": np.zeros(n),
        "punctuality": np.zeros(n),
        "episode_total": np.zeros(n),
        "social_in": np.zeros((3, n)),
        "term_block_tick": np.full(n, capi.TL_TERM_BLOCK_NONE, dtype=np.int32),
        "lateness": np.zeros(n, dtype=np.int32),
        "last_action_duration": np.zeros(n, dtype=np.int32),
        "episode_tick": np.zeros(n, dtype=np.int32),
        "slot": np.zeros(n, dtype=np.int32),
        "flags": np.zeros(n, dtype=np.uint32),
        "last_action_hash": np.zeros(n, dtype=np.float32),
        "personality": np.zeros((3, n), dtype=np.float32),
        "tie_count": np.zeros(n, dtype=np.uint8),
        "tie_embed": np.zeros((n, K, ew), dtype=np.uint64),
        "tie_trust": np.zeros((n, K)),
        "tie_fam": np.zeros((n, K)),
        "tie_riv": np.zeros((n, K)),
        "riv_count": np.zeros(n, dtype=np.uint8),
        "riv_embed": np.zeros((n, K, ew), dtype=np.uint64),
        "riv_score": np.zeros((n, K)),
    }

    wages_fn = getattr(adapter, "_employment_context_wages", None)
    punctuality_fn = getattr(adapter, "_employment_context_punctuality", None)
    if wages_fn is None:  # context.py:45-50
        raise RuntimeError("World runtime adapter is missing employment context wages hook")
    if punctuality_fn is None:
        raise RuntimeError("World runtime adapter is missing employment punctuality hook")

    holders: set[str] = set()
    queued: set[str] = set()
    relationships: Mapping[str, Any] = {}
    objects_view: Mapping[str, Mapping[str, object]] = {}
    if observe:
        objects_view = adapter.objects_snapshot()
        holders = {str(v) for v in adapter.active_reservations.values() if v}  # features.py:215-219
        queue_manager = getattr(adapter, "queue_manager", None)
        if queue_manager is not None:
            for object_id in objects_view.keys():  # features.py:221-229
                for waiting in queue_manager.queue_snapshot(object_id):
                    queued.add(str(waiting))
        getter = getattr(adapter, "relationships_snapshot", None)
        if callable(getter):
            data = getter()
            if isinstance(data, dict):
                relationships = data

    for i, agent_id in enumerate(agent_ids):
        snap = snapshots[agent_id]
        position = tuple(snap.position)
        centre[i] = (int(position[0]), int(position[1]))
        indexed = adapter.agent_position(agent_id) if observe else None
        occupancy = indexed or position  # cache.py:30
        occ[i] = (int(occupancy[0]), int(occupancy[1]))
        needs = getattr(snap, "needs", {})
        for k, need in enumerate(NEEDS):
            f["needs"][k, i] = _as_float(needs.get(need, 0.0))
        f["wallet"][i] = _as_float(getattr(snap, "wallet", 0.0))
        f["attendance"][i] = _as_float(getattr(snap, "attendance_ratio", 0.0))
        f["wages_withheld"][i] = _as_float(getattr(snap, "wages_withheld", 0.0))
        f["wages_paid"][i] = _as_float(wages_fn(agent_id))
        f["punctuality"][i] = _as_float(punctuality_fn(agent_id))
        f["lateness"][i] = _as_int(getattr(snap, "lateness_counter", 0))
        f["last_action_duration"][i] = _as_int(getattr(snap, "last_action_duration", 0))
        f["episode_tick"][i] = _as_int(getattr(snap, "episode_tick", 0))
        flags = 0
        if bool(getattr(snap, "on_shift", False)):
            flags |= capi.TL_F_ON_SHIFT
        flags |= SHIFT_CODE.get(getattr(snap, "shift_state", "pre_shift"), 0) << capi.TL_F_SHIFT_SHIFT
        if bool(getattr(snap, "last_action_success", False)):
            flags |= capi.TL_F_LAST_SUCCESS
        if agent_id in queued:
            flags |= capi.TL_F_IN_QUEUE
        if agent_id in holders:
            flags |= capi.TL_F_RESERVATION
        if terminated.get(agent_id, False):
            flags |= capi.TL_F_TERMINATED
            reason = reasons.get(agent_id)
            code = REASON_CODE.get(reason, 3 if reason else 0)  # type: ignore[arg-type]
            flags |= code << capi.TL_F_REASON_SHIFT
        if agent_id in pending_resets:
            flags |= capi.TL_F_CTX_RESET
        profile_name = getattr(snap, "personality_profile", None)
        if profile_name:
            flags |= PROFILE_CODE.get(str(profile_name).strip().lower(), 0) << capi.TL_F_PROFILE_SHIFT
        f["flags"][i] = flags
        last_action_id = getattr(snap, "last_action_id", "")
        f["last_action_hash"][i] = last_action_hash(last_action_id if isinstance(last_action_id, str) else "")
        personality = getattr(snap, "personality", None)
        f["personality"][0, i] = float(getattr(personality, "extroversion", 0.0))
        f["personality"][1, i] = float(getattr(personality, "forgiveness", 0.0))
        f["personality"][2, i] = float(getattr(personality, "ambition", 0.0))
        if slots is not None:
            f["slot"][i] = int(slots[agent_id])
        if episode_totals is not None:
            f["episode_total"][i] = float(episode_totals.get(agent_id, 0.0))
        if termination_block is not None and agent_id in termination_block:
            f["term_block_tick"][i] = int(termination_block[agent_id])
        if social_sums is not None:
            for k in range(3):
                f["social_in"][k, i] = float(social_sums[k].get(agent_id, 0.0))
        if observe:
            candidate = relationships.get(agent_id, {}) or {}  # social.py:171-178
            ties = list(candidate.items()) if isinstance(candidate, dict) else []
            if len(ties) > K:
                raise ValueError(f"agent {agent_id!r} has {len(ties)} ties; max_edges={K}")
            f["tie_count"][i] = len(ties)
            for j, (other_id, metrics) in enumerate(ties):
                f["tie_embed"][i, j] = id_embedding_words(str(other_id), ew)
                f["tie_trust"][i, j] = _as_float(metrics.get("trust", 0.0))
                f["tie_fam"][i, j] = _as_float(metrics.get("familiarity", 0.0))
                f["tie_riv"][i, j] = _as_float(metrics.get("rivalry", 0.0))
            top = list(adapter.rivalry_top(agent_id, K))  # features.py:239-243 / social.py:196
            f["riv_count"][i] = len(top)
            for j, (other_id, score) in enumerate(top):
                f["riv_embed"][i, j] = id_embedding_words(str(other_id), ew)
                f["riv_score"][i, j] = float(score)

    # ---- entities ----------------------------------------------------------
    ctype_of = {name: k + 1 for k, name in enumerate(spec.object_channels)} if spec.variant == "compact" else {}
    obj_rows: list[tuple[int, int, int, int, int, int]] = []  # x, y, ltype, ctype, lm, tile
    reserved: list[tuple[int, int]] = []
    slots: list[tuple[int, int, int, int]] = []  # x, y, kind, holder (device_reservations)
    if observe:
        tile_of: dict[str, tuple[int, int]] = {}
        for position, ids in adapter.objects_by_position_view().items():  # cache.py:33-38
            for object_id in ids:
                if object_id in objects_view:
                    tile_of[object_id] = (int(position[0]), int(position[1]))
        for object_id, payload in objects_view.items():  # insertion order (observe.py:120-131)
            object_type = payload.get("object_type")
            ltype = LTYPE_CODE.get(object_type, 0) if isinstance(object_type, str) else 0
            ctype = ctype_of.get(str(object_type or "").strip().lower(), 0)  # map.py:218
            position = payload.get("position")
            lm_pos = None
            if isinstance(position, tuple) and len(position) == 2 and all(isinstance(c, int) for c in position):
                lm_pos = (int(position[0]), int(position[1]))
            tile_pos = tile_of.get(object_id)
            if tile_pos is not None and lm_pos == tile_pos:
                obj_rows.append((tile_pos[0], tile_pos[1], ltype, ctype, 1, 1))
            else:
                if lm_pos is not None:
                    obj_rows.append((lm_pos[0], lm_pos[1], ltype, ctype, 1, 0))
                if tile_pos is not None:
                    obj_rows.append((tile_pos[0], tile_pos[1], ltype, ctype, 0, 1))
        # a set in the reference (spatial.py:117-120): canonical (y, x) order here
        reserved = sorted({(int(p[0]), int(p[1])) for p in adapter.reservation_tiles()}, key=lambda p: (p[1], p[0]))
        if device_reservations:
            queue_manager = getattr(adapter, "queue_manager", None)
            index_of = {aid: k for k, aid in enumerate(agent_ids)}
            reserved_now = set(reserved)
            positions: list[tuple[int, int]] = []
            for object_id, payload in objects_view.items():  # refresh_reservations walks the objects in this order
                position = payload.get("position")
                if not (isinstance(position, tuple) and len(position) == 2):
                    continue  # sync_reservation skips objects without a position (queues.py:69-71, 82-83)
                active = queue_manager.active_agent(object_id) if queue_manager is not None else None
                holder = -1 if active is None else index_of.get(str(active), n)  # a holder outside the town: held, nobody flagged
                positions.append((int(position[0]), int(position[1])))
                slots.append((int(position[0]), int(position[1]), capi.TL_KIND_EMPTY, holder))
            for k, pos_k in enumerate(positions):  # the tile's current state sits on the last slot of the tile
                if pos_k in reserved_now and pos_k not in positions[k + 1 :]:
                    slots[k] = (pos_k[0], pos_k[1], capi.TL_KIND_RESERVED, slots[k][3])
            reserved = [p for p in reserved if p not in set(positions)]

    xs = [occ[:, 0]] if n else []
    ys = [occ[:, 1]] if n else []
    if obj_rows:
        arr = np.array(obj_rows, dtype=np.int64)
        xs.append(arr[:, 0])
        ys.append(arr[:, 1])
    if reserved:
        rarr = np.array(reserved, dtype=np.int64)
        xs.append(rarr[:, 0])
        ys.append(rarr[:, 1])
    if slots:
        sarr = np.array(slots, dtype=np.int64)
        xs.append(sarr[:, 0])
        ys.append(sarr[:, 1])
    if xs:
        allx = np.concatenate(xs)
        ally = np.concatenate(ys)
        ox, oy = (int(allx.min()), int(ally.min())) if origin is None else origin
        if allx.min() < ox or ally.min() < oy:
            raise ValueError("entity left of / above the declared town origin")
        extent = (int(allx.max()) - ox + 1, int(ally.max()) - oy + 1)
    else:
        ox, oy = origin or (0, 0)
        extent = (1, 1)
    ent_parts = []
    if n:
        ent_parts.append(pack_entity(occ[:, 0] - ox, occ[:, 1] - oy, capi.TL_KIND_AGENT))
    if obj_rows:
        ent_parts.append(
            pack_entity(arr[:, 0] - ox, arr[:, 1] - oy, capi.TL_KIND_OBJECT, arr[:, 2], arr[:, 3], arr[:, 4], arr[:, 5])
        )
    holder_parts = [np.full(part.size, -2, dtype=np.int32) for part in ent_parts]
    if slots:
        ent_parts.append(pack_entity(sarr[:, 0] - ox, sarr[:, 1] - oy, sarr[:, 2]))
        holder_parts.append(sarr[:, 3].astype(np.int32))
    if reserved:
        ent_parts.append(pack_entity(rarr[:, 0] - ox, rarr[:, 1] - oy, capi.TL_KIND_RESERVED))
        holder_parts.append(np.full(len(reserved), -2, dtype=np.int32))
    ent = np.concatenate(ent_parts) if ent_parts else np.zeros(0, dtype=np.uint32)
    ent_holder = np.concatenate(holder_parts) if holder_parts else np.zeros(0, dtype=np.int32)
    f["pos"] = pack_pos(centre[:, 0] - ox, centre[:, 1] - oy) if n else np.zeros(0, dtype=np.int32)
    return TownRows(
        agent_ids=agent_ids,
        origin=(ox, oy),
        extent=extent,
        tick=int(adapter.tick),
        ent=ent,
        fields=f,
        ent_holder=ent_holder,
    )


def _round_up(v: int, m: int) -> int:
    return (max(v, 1) + m - 1) // m * m


def rows_to_batch(towns: list[TownRows], spec: ObsSpec, max_edges: int = 6, grid: tuple[int, int] | None = None) -> TownBatch:
    """Concatenate ingested towns into one ``TownBatch``."""
    n_agents = sum(len(t.agent_ids) for t in towns)
    n_ent = sum(int(t.ent.size) for t in towns)
    if grid is None:
        gw = _round_up(max((t.extent[0] for t in towns), default=1), 16)
        gh = _round_up(max((t.extent[1] for t in towns), default=1), 16)
    else:
        gw, gh = grid
    layout = BatchLayout.make(len(towns), n_agents, n_ent, max_edges, spec.embed_words)
    b = TownBatch(layout, gw, gh)
    a = e = 0
    b.town_agent_off[0] = 0
    b.town_ent_off[0] = 0
    for ti, t in enumerate(towns):
        n = len(t.agent_ids)
        b.town_tick[ti] = t.tick
        for name, values in t.fields.items():
            dst = getattr(b, name)
            if name in ("needs", "social_in", "personality"):
                dst[:, a : a + n] = values
            else:
                dst[a : a + n] = values
        b.ent[e : e + t.ent.size] = t.ent
        b.ent_holder[e : e + t.ent.size] = t.ent_holder if t.ent_holder is not None else -2
        a += n
        e += int(t.ent.size)
        b.town_agent_off[ti + 1] = a
        b.town_ent_off[ti + 1] = e
    return b
