"""Structure-of-arrays town batch (host mirror of ``TlTownBatch``).

Replaces the per-agent Python objects of the reference — ``AgentSnapshot``
(src/townlet/world/agents/snapshot.py:19-44), ``InteractiveObject``
(world/grid.py:121-128), the relationship / rivalry ledgers
(world/relationships.py:40, world/rivalry.py:36) and the spatial index
(world/spatial.py:14) — with flat arrays, town-major.

All arrays of one batch live in ONE byte arena at 256-byte aligned offsets so
that a batch moves host→device as a single copy and the device layout is the
same bytes.  The fields the tick mutates (needs, episode totals, termination
blocks, episode ticks, flags, town ticks) come first, so reading state back is
one contiguous device→host copy of ``mutable_bytes``.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Callable

import numpy as np

from . import capi

_ALIGN = 256

# (name, dtype, shape builder(T, N, E, K, EW), mutable)
_FIELDS: tuple[tuple[str, str, Callable[[int, int, int, int, int], tuple[int, ...]], bool], ...] = (
    # ---- mutated by the tick --------------------------------------------
    ("needs", "f8", lambda T, N, E, K, EW: (3, N), True),
    ("episode_total", "f8", lambda T, N, E, K, EW: (N,), True),
    ("term_block_tick", "i4", lambda T, N, E, K, EW: (N,), True),
    ("episode_tick", "i4", lambda T, N, E, K, EW: (N,), True),
    ("flags", "u4", lambda T, N, E, K, EW: (N,), True),
    ("town_tick", "i4", lambda T, N, E, K, EW: (T,), True),
    # ---- read-only for the tick -----------------------------------------
    ("town_agent_off", "i4", lambda T, N, E, K, EW: (T + 1,), False),
    ("town_ent_off", "i4", lambda T, N, E, K, EW: (T + 1,), False),
    ("ent", "u4", lambda T, N, E, K, EW: (E,), False),
    ("pos", "i4", lambda T, N, E, K, EW: (N,), False),
    ("wallet", "f8", lambda T, N, E, K, EW: (N,), False),
    ("attendance", "f8", lambda T, N, E, K, EW: (N,), False),
    ("wages_withheld", "f8", lambda T, N, E, K, EW: (N,), False),
    ("wages_paid", "f8", lambda T, N, E, K, EW: (N,), False),
    ("punctuality", "f8", lambda T, N, E, K, EW: (N,), False),
    ("social_in", "f8", lambda T, N, E, K, EW: (3, N), False),
    ("lateness", "i4", lambda T, N, E, K, EW: (N,), False),
    ("last_action_duration", "i4", lambda T, N, E, K, EW: (N,), False),
    ("slot", "i4", lambda T, N, E, K, EW: (N,), False),
    ("last_action_hash", "f4", lambda T, N, E, K, EW: (N,), False),
    ("personality", "f4", lambda T, N, E, K, EW: (3, N), False),
    ("tie_count", "u1", lambda T, N, E, K, EW: (N,), False),
    ("tie_embed", "u8", lambda T, N, E, K, EW: (N, K, EW), False),
    ("tie_trust", "f8", lambda T, N, E, K, EW: (N, K), False),
    ("tie_fam", "f8", lambda T, N, E, K, EW: (N, K), False),
    ("tie_riv", "f8", lambda T, N, E, K, EW: (N, K), False),
    ("riv_count", "u1", lambda T, N, E, K, EW: (N,), False),
    ("riv_embed", "u8", lambda T, N, E, K, EW: (N, K, EW), False),
    ("riv_score", "f8", lambda T, N, E, K, EW: (N, K), False),
    # reservation slots (TL_PHASE_RESERVATIONS): holder of the slot's object, -1 free, -2 not a slot
    ("ent_holder", "i4", lambda T, N, E, K, EW: (E,), False),
)

FIELD_NAMES = tuple(f[0] for f in _FIELDS)
MUTABLE_FIELDS = tuple(f[0] for f in _FIELDS if f[3])


@dataclass(frozen=True)
class BatchLayout:
    """Byte offsets of every field inside the arena for given sizes."""

    n_towns: int
    n_agents: int
    n_entities: int
    max_edges: int
    embed_words: int
    offsets: dict[str, int]
    shapes: dict[str, tuple[int, ...]]
    dtypes: dict[str, np.dtype]
    total_bytes: int
    mutable_bytes: int

    @staticmethod
    def make(n_towns: int, n_agents: int, n_entities: int, max_edges: int = 6, embed_words: int = 1) -> "BatchLayout":
        if not (0 < max_edges <= capi.TL_MAX_EDGES):
            raise ValueError(f"max_edges must be in 1..{capi.TL_MAX_EDGES}")
        if not (1 <= embed_words <= 4):
            raise ValueError("embed_words must be in 1..4")
        offsets: dict[str, int] = {}
        shapes: dict[str, tuple[int, ...]] = {}
        dtypes: dict[str, np.dtype] = {}
        cursor = 0
        mutable_bytes = 0
        for name, dt, shape_fn, mutable in _FIELDS:
            shape = shape_fn(n_towns, n_agents, n_entities, max_edges, embed_words)
            dtype = np.dtype(dt)
            offsets[name] = cursor
            shapes[name] = shape
            dtypes[name] = dtype
            nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
            cursor += (nbytes + _ALIGN - 1) // _ALIGN * _ALIGN
            if mutable:
                mutable_bytes = cursor
        return BatchLayout(
            n_towns, n_agents, n_entities, max_edges, embed_words, offsets, shapes, dtypes, max(cursor, _ALIGN), mutable_bytes
        )


def pack_pos(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """Pack int16 (x, y) into the ``pos`` word."""
    x = np.asarray(x, dtype=np.int64)
    y = np.asarray(y, dtype=np.int64)
    if x.size and (x.min() < -32768 or x.max() > 32767 or y.min() < -32768 or y.max() > 32767):
        raise ValueError("window centre outside int16 range")
    return ((x & 0xFFFF) | ((y & 0xFFFF) << 16)).astype(np.uint32).view(np.int32)


def unpack_pos(pos: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    u = np.asarray(pos).view(np.uint32)
    return (u & 0xFFFF).astype(np.uint16).view(np.int16).astype(np.int32), (u >> 16).astype(np.uint16).view(
        np.int16
    ).astype(np.int32)


def pack_entity(x, y, kind, ltype=0, ctype=0, lm=0, tile=0) -> np.ndarray:
    x = np.asarray(x, dtype=np.int64)
    y = np.asarray(y, dtype=np.int64)
    if x.size and (x.min() < 0 or y.min() < 0 or x.max() >= capi.TL_MAX_GRID or y.max() >= capi.TL_MAX_GRID):
        raise ValueError("entity coordinate outside 0..1023 after town-origin shift")
    word = (
        x
        | (y << 10)
        | (np.asarray(kind, dtype=np.int64) << 20)
        | (np.asarray(ltype, dtype=np.int64) << 22)
        | (np.asarray(ctype, dtype=np.int64) << 25)
        | (np.asarray(lm, dtype=np.int64) << 29)
        | (np.asarray(tile, dtype=np.int64) << 30)
    )
    return word.astype(np.uint32)


class TownBatch:
    """Host-side SoA batch: numpy views into one byte arena."""

    def __init__(
        self,
        layout: BatchLayout,
        grid_w: int,
        grid_h: int,
        arena: np.ndarray | None = None,
    ) -> None:
        if not (0 < grid_w <= capi.TL_MAX_GRID and 0 < grid_h <= capi.TL_MAX_GRID):
            raise ValueError("grid dimensions must be in 1..1024")
        self.layout = layout
        self.grid_w = int(grid_w)
        self.grid_h = int(grid_h)
        fresh = arena is None
        if arena is None:
            arena = np.zeros(layout.total_bytes, dtype=np.uint8)
        if arena.dtype != np.uint8 or arena.size < layout.total_bytes:
            raise ValueError("arena must be uint8 with at least layout.total_bytes bytes")
        self.arena = arena
        self._views: dict[str, np.ndarray] = {}
        for name in FIELD_NAMES:
            off = layout.offsets[name]
            shape = layout.shapes[name]
            dtype = layout.dtypes[name]
            nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
            self._views[name] = arena[off : off + nbytes].view(dtype).reshape(shape)
        if fresh:
            self._views["term_block_tick"][...] = capi.TL_TERM_BLOCK_NONE
            self._views["ent_holder"][...] = -2
        self.has_social_in = True
        self.has_personality = True

    def __getattr__(self, name: str) -> np.ndarray:
        views = self.__dict__.get("_views")
        if views is not None and name in views:
            return views[name]
        raise AttributeError(name)

    # sizes ----------------------------------------------------------------
    @property
    def n_towns(self) -> int:
        return self.layout.n_towns

    @property
    def n_agents(self) -> int:
        return self.layout.n_agents

    @property
    def n_entities(self) -> int:
        return self.layout.n_entities

    def views(self) -> dict[str, np.ndarray]:
        return dict(self._views)

    def copy(self) -> "TownBatch":
        other = TownBatch(self.layout, self.grid_w, self.grid_h, self.arena.copy())
        other.has_social_in = self.has_social_in
        other.has_personality = self.has_personality
        return other

    # C view ---------------------------------------------------------------
    def struct(self, base_ptr: int | None = None) -> capi.TlTownBatch:
        """``TlTownBatch`` whose pointers address ``base_ptr`` + field offset.

        ``base_ptr`` defaults to the host arena; pass a device pointer holding
        the same bytes for the device entry points.
        """
        if base_ptr is None:
            base_ptr = self.arena.ctypes.data
        lay = self.layout
        s = capi.TlTownBatch()
        s.n_towns = lay.n_towns
        s.n_agents = lay.n_agents
        s.n_entities = lay.n_entities
        s.max_edges = lay.max_edges
        s.embed_words = lay.embed_words
        s.grid_w = self.grid_w
        s.grid_h = self.grid_h
        for name in FIELD_NAMES:
            setattr(s, name, C.c_void_p(base_ptr + lay.offsets[name]))
        # the arrays share one block: the host-buffer entry points may move it with a single copy
        s.arena_base = C.c_void_p(base_ptr)
        s.arena_bytes = lay.total_bytes
        if not self.has_social_in:
            s.social_in = None
        if not self.has_personality:
            s.personality = None
        return s

    def validate(self, n_ctypes: int = 0) -> None:
        """Check the invariants the kernels rely on (raises ``ValueError``)."""
        lay = self.layout
        ao = self.town_agent_off
        eo = self.town_ent_off
        if ao[0] != 0 or ao[-1] != lay.n_agents or np.any(np.diff(ao) < 0):
            raise ValueError("town_agent_off must be a monotone prefix over n_agents")
        if eo[0] != 0 or eo[-1] != lay.n_entities or np.any(np.diff(eo) < 0):
            raise ValueError("town_ent_off must be a monotone prefix over n_entities")
        if np.any(np.diff(eo) < np.diff(ao)):
            raise ValueError("every agent needs an occupancy entity")
        ent = self.ent
        x = ent & 1023
        y = (ent >> 10) & 1023
        if ent.size and (x.max() >= self.grid_w or y.max() >= self.grid_h):
            raise ValueError("entity outside the declared town grid")
        if np.any(self.tie_count > lay.max_edges) or np.any(self.riv_count > lay.max_edges):
            raise ValueError("edge count exceeds max_edges")
        # per-tile counters are 8 bit (agents) / 7 bit (objects) / 4 bit (types) on device
        town_of_ent = np.repeat(np.arange(lay.n_towns, dtype=np.int64), np.diff(eo))
        key = (town_of_ent * self.grid_h + y.astype(np.int64)) * self.grid_w + x.astype(np.int64)
        kind = (ent >> 20) & 3
        for mask, limit, what in (
            (kind == capi.TL_KIND_AGENT, 255, "agents"),
            ((kind == capi.TL_KIND_OBJECT) & (((ent >> 30) & 1) == 1), 127, "objects"),
        ):
            if mask.any():
                _, counts = np.unique(key[mask], return_counts=True)
                if counts.max() > limit:
                    raise ValueError(f"more than {limit} {what} on one tile")
        if n_ctypes:
            ct = (ent >> 25) & 15
            mask = (kind == capi.TL_KIND_OBJECT) & (ct > 0) & (((ent >> 30) & 1) == 1)
            if mask.any():
                _, counts = np.unique(key[mask] * 16 + ct[mask], return_counts=True)
                if counts.max() > 15:
                    raise ValueError("more than 15 objects of one channel type on one tile")
        # agent entities come first in each town, in agent order
        n_per_town = np.diff(ao)
        first = eo[:-1]
        for t in np.nonzero(n_per_town)[0][:64]:  # spot-check a prefix; full check is O(E) below
            seg = kind[first[t] : first[t] + n_per_town[t]]
            if np.any(seg != capi.TL_KIND_AGENT):
                raise ValueError("agent occupancy entities must lead each town's entity list")

    def slice_towns(self, t0: int, t1: int) -> "TownBatch":
        """Contiguous town shard [t0, t1) as a new batch (multi-GPU partition)."""
        lay = self.layout
        a0, a1 = int(self.town_agent_off[t0]), int(self.town_agent_off[t1])
        e0, e1 = int(self.town_ent_off[t0]), int(self.town_ent_off[t1])
        sub_layout = BatchLayout.make(t1 - t0, a1 - a0, e1 - e0, lay.max_edges, lay.embed_words)
        out = TownBatch(sub_layout, self.grid_w, self.grid_h)
        out.has_social_in = self.has_social_in
        out.has_personality = self.has_personality
        for name in FIELD_NAMES:
            src = self._views[name]
            dst = out._views[name]
            if name == "town_agent_off":
                dst[...] = self.town_agent_off[t0 : t1 + 1] - a0
            elif name == "town_ent_off":
                dst[...] = self.town_ent_off[t0 : t1 + 1] - e0
            elif name == "town_tick":
                dst[...] = src[t0:t1]
            elif name in ("ent", "ent_holder"):
                dst[...] = src[e0:e1]
            elif src.ndim >= 1 and src.shape[0] == 3 and name in ("needs", "social_in", "personality"):
                dst[...] = src[:, a0:a1]
            else:
                dst[...] = src[a0:a1]
        return out
