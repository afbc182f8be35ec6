"""Drop-in for the reference's ``WorldObservationService``.

Implements ``ObservationServiceProtocol``
(src/townlet/world/observations/interfaces.py:111-140): properties
``variant``, ``feature_names``, ``social_vector_length``; methods
``build_batch(world, terminated)`` and ``build_single(world, agent_id, *, slot)``
with the reference's argument meaning, return shape, side effects and errors
(src/townlet/world/observations/service.py:27-517).  The arithmetic runs in the
CUDA library through ``tl_host_tick``; this class only ingests the world into
the SoA batch and scatters the tensors back into the per-agent dict layout.

Side effects preserved (SURVEY.md §7 "Destructive reads"): ctx-reset requests
are consumed exactly once per call, embedding slots are allocated in agent
iteration order and released for terminated agents.
"""

from __future__ import annotations

from collections.abc import Mapping
from typing import Any

import numpy as np

from . import capi
from .engine import HostEngine
from .ingest import ingest_town, resolve_adapter, rows_to_batch
from .params import ObsSpec

_RELATION_SOURCE = {0: "empty", 1: "relationships", 2: "rivalry_fallback"}

# agents/models.py:61-85 — metadata only (encoders/features.py:383-404)
_PROFILE_TABLES: dict[str, dict[str, dict[str, float]]] = {
    "balanced": {"needs": {}, "rewards": {}, "behaviour": {}},
    "socialite": {"needs": {"hygiene": 1.1}, "rewards": {"social": 1.2}, "behaviour": {"chat_preference": 1.3}},
    "industrious": {"needs": {"energy": 0.9}, "rewards": {"employment": 1.25}, "behaviour": {"work_affinity": 1.4}},
    "stoic": {"needs": {"hunger": 0.95}, "rewards": {"survival": 1.1}, "behaviour": {"conflict_tolerance": 0.6}},
}


class B200ObservationService:
    """``ObservationServiceProtocol`` backed by the sm_100a observation kernel."""

    MAP_CHANNELS = ("self", "agents", "objects", "reservations")
    SHIFT_STATES = ("pre_shift", "on_time", "late", "absent", "post_shift")

    def __init__(self, *, config: Any, device: int = 0, max_edges: int | None = None) -> None:
        self.config = config
        self.spec = ObsSpec.from_config(config)  # raises ValueError when relationships stage is OFF
        self._variant = config.observation_variant
        self._feature_names = list(self.spec.feature_names)
        self._feature_index = {name: idx for idx, name in enumerate(self._feature_names)}
        self._max_edges = int(max_edges or max(6, int(config.conflict.rivalry.max_edges)))
        if self._max_edges > capi.TL_MAX_EDGES:
            raise ValueError(f"max_edges above {capi.TL_MAX_EDGES} is not supported")
        self._engine = HostEngine(self.spec, None, device)
        # the configuration sections the reference service exposes (service.py:36-38, 62-63); its tests read them
        self.hybrid_cfg = config.observations_config.hybrid
        self.compact_cfg = config.observations_config.compact
        self.social_cfg = config.observations_config.social_snippet
        self.full_channels = ("self", "agents", "objects", "reservations", "path_dx", "path_dy")
        self.compact_map_channels = (
            "self", "agents", "objects", "reservations", *(f"object:{name}" for name in self.spec.object_channels), "walkable",
        )
        # float64 nearest distance by squared distance, np.hypot as in map.py:95
        radius = self.spec.radius
        self._hypot_by_d2: dict[int, float] = {}
        for dy in range(-radius, radius + 1):
            for dx in range(-radius, radius + 1):
                self._hypot_by_d2.setdefault(dx * dx + dy * dy, float(np.hypot(dx, dy)))

    def _aggregate_names(self) -> list[str]:
        """service.py:505-514"""
        if self.social_cfg.include_aggregates:
            return ["social_trust_mean", "social_trust_max", "social_rivalry_mean", "social_rivalry_max"]
        return []

    # ---- protocol properties ---------------------------------------------
    @property
    def variant(self) -> object:
        return self._variant

    @property
    def feature_names(self) -> tuple[str, ...]:
        return tuple(self._feature_names)

    @property
    def social_vector_length(self) -> int:
        return self.spec.social_vector_length

    # ---- protocol methods ------------------------------------------------
    def build_batch(self, world: Any, terminated: Mapping[str, bool]) -> Mapping[str, Mapping[str, Any]]:
        adapter = resolve_adapter(world)
        snapshots = dict(adapter.agent_snapshots_view())
        pending_resets = set(adapter.consume_ctx_reset_requests())  # destructive, once (service.py:210)
        tick = adapter.tick
        slots: dict[str, int] = {}
        allocator = adapter.embedding_allocator
        for agent_id in snapshots:
            # one loop in agent order, as the reference interleaves them (service.py:216-229): allocate, and release
            # at once when the agent is terminated — so a later agent without a slot can take the one just freed
            # (forced reuse when the pool is full).  The slot value is already captured for the kernel.
            slots[agent_id] = allocator.allocate(agent_id, tick)
            if terminated.get(agent_id, False):
                allocator.release(agent_id, tick)
        observations: dict[str, dict[str, Any]] = {}
        if snapshots:
            rows = ingest_town(
                adapter,
                self.spec,
                terminated=terminated,
                pending_resets=pending_resets,
                slots=slots,
                max_edges=self._max_edges,
            )
            batch = rows_to_batch([rows], self.spec, self._max_edges)
            batch.validate(len(self.spec.object_channels))
            out = self._engine.alloc_outputs(batch, capi.TL_PHASE_OBS, 1, summary=True)
            self._engine.tick(batch, out, capi.TL_PHASE_OBS, 1)
            for i, agent_id in enumerate(rows.agent_ids):
                snapshot = snapshots[agent_id]
                observations[agent_id] = {
                    "map": np.array(out["map"][0, i], dtype=np.float32),  # caller owns fresh arrays
                    "features": np.array(out["features"][0, i], dtype=np.float32),
                    "metadata": self._build_metadata(slots[agent_id], out["summary"][0, i], snapshot),
                }
        return observations

    def build_single(self, world: Any, agent_id: str, *, slot: int | None = None) -> Mapping[str, Any]:
        batch = self.build_batch(world, {})
        if agent_id not in batch:
            raise KeyError(f"agent '{agent_id}' not present in observation batch")
        entry = batch[agent_id]
        if slot is None:
            return entry
        payload = dict(entry)
        metadata_raw = payload.get("metadata", {})
        metadata = dict(metadata_raw) if isinstance(metadata_raw, dict) else {}
        metadata["embedding_slot"] = slot
        payload["metadata"] = metadata
        return payload

    # ---- metadata (service.py:474-517, features.py:320-330, 383-404) ------
    def _local_summary(self, row: np.ndarray) -> dict[str, float]:
        spec = self.spec
        window = spec.window
        radius = spec.radius
        total_tiles = window * window
        other_agents = int(row[0])
        object_total = int(row[1])
        reserved_tiles = int(row[2])
        d2 = int(row[3])
        max_tiles = max(1, total_tiles - 1)
        agent_ratio = min(1.0, other_agents / max_tiles)
        object_ratio = min(1.0, object_total / max(1, total_tiles))
        reserved_ratio = min(1.0, reserved_tiles / max(1, total_tiles))
        if d2 == 0:
            nearest = 0.0
            nearest_norm = 0.0
        else:
            nearest = self._hypot_by_d2[d2]
            nearest_norm = max(0.0, min(1.0, nearest / max(1, radius)))
        return {
            "agent_count": float(other_agents),
            "object_count": float(object_total),
            "reserved_tiles": float(reserved_tiles),
            "radius": float(radius),
            "agent_ratio": float(agent_ratio),
            "object_ratio": float(object_ratio),
            "reserved_ratio": float(reserved_ratio),
            "nearest_agent_distance": float(nearest),
            "nearest_agent_distance_norm": float(nearest_norm),
        }

    def _personality_context(self, snapshot: Any) -> dict[str, object]:
        personality = getattr(snapshot, "personality", None)
        profile_name = str(getattr(snapshot, "personality_profile", "") or "balanced")
        metadata: dict[str, object] = {
            "profile": profile_name,
            "traits": {
                "extroversion": float(getattr(personality, "extroversion", 0.0)),
                "forgiveness": float(getattr(personality, "forgiveness", 0.0)),
                "ambition": float(getattr(personality, "ambition", 0.0)),
            },
        }
        tables = _PROFILE_TABLES.get(profile_name.lower())
        if tables is not None:
            metadata["multipliers"] = {key: dict(value) for key, value in tables.items()}
        return metadata

    def _build_metadata(self, slot: int, summary_row: np.ndarray, snapshot: Any) -> dict[str, object]:
        spec = self.spec
        metadata: dict[str, object] = {
            "variant": self._variant,
            "map_shape": tuple(int(v) for v in spec.map_shape),
            "map_channels": list(spec.map_channels),
            "feature_names": list(self._feature_names),
            "embedding_slot": slot,
        }
        if spec.landmark_slices:
            metadata["landmark_slices"] = dict(spec.landmark_slices)
        if spec.social_slots:
            filled = int(summary_row[capi.SUMMARY_NAMES.index("filled_slots")])
            source = _RELATION_SOURCE[int(summary_row[capi.SUMMARY_NAMES.index("relation_source")])]
            metadata["social_context"] = {
                "configured_slots": spec.social_slots,
                "slot_dim": spec.social_slot_dim,
                "aggregates": spec.aggregate_names,
                "filled_slots": filled,
                "relation_source": source,
                "has_data": bool(filled),
            }
        metadata["local_summary"] = self._local_summary(summary_row)
        if spec.personality_channels:
            metadata["personality"] = self._personality_context(snapshot)
        if spec.variant == "compact":
            cfg = self.config.observations_config.compact
            metadata["compact"] = {
                "map_window": spec.window,
                "object_channels": list(spec.object_channels),
                "normalize_counts": bool(cfg.normalize_counts),
                "include_targets": bool(cfg.include_targets),
            }
        return metadata


__all__ = ["B200ObservationService"]
