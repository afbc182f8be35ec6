"""Kernel parameters derived from the reference ``SimulationConfig``.

``ObsSpec`` / ``RewardSpec`` are plain value objects: ``from_config`` reads a
reference config (duck-typed, src/townlet/config/loader.py:205-296); the
defaults equal ``configs/examples/poc_hybrid.yaml`` so synthetic runs need no
reference install.  ``ObsSpec.luts()`` evaluates the reference's float64
expressions on the host once and stores the float32 results, so the device
reproduces them bit-for-bit (SURVEY.md §7 "Float width").
"""

from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Any

import numpy as np

from . import capi

LANDMARKS = ("fridge", "stove", "bed", "shower")  # service.py:63
LTYPE_CODE = {name: i + 1 for i, name in enumerate(LANDMARKS)}
SHIFT_STATES = ("pre_shift", "on_time", "late", "absent", "post_shift")  # service.py:31
SHIFT_CODE = {name: i for i, name in enumerate(SHIFT_STATES)}
PROFILE_CODE = {"balanced": 1, "socialite": 2, "industrious": 3, "stoic": 4}  # agents/models.py:60-86
REASON_CODE = {"faint": 1, "eviction": 2}
NEEDS = ("hunger", "hygiene", "energy")

# agents/models.py:61-85 (need_multipliers / reward_bias of the built-in profiles)
_PROFILE_NEED_MULT = {
    "socialite": {"hygiene": 1.1},
    "industrious": {"energy": 0.9},
    "stoic": {"hunger": 0.95},
}
_PROFILE_REWARD_BIAS = {
    "socialite": {"social": 1.2},
    "industrious": {"employment": 1.25},
    "stoic": {"survival": 1.1},
}
_BIAS_GROUPS = ("social", "employment", "survival")

_BASE_FEATURES = (
    "need_hunger",
    "need_hygiene",
    "need_energy",
    "wallet",
    "lateness_counter",
    "on_shift",
    "time_sin",
    "time_cos",
    "attendance_ratio",
    "wages_withheld",
    "shift_pre",
    "shift_on_time",
    "shift_late",
    "shift_absent",
    "shift_post",
    "embedding_slot_norm",
    "ctx_reset_flag",
    "episode_progress",
    "rivalry_max",
    "rivalry_avoid_count",
    "last_action_id_hash",
    "last_action_success",
    "last_action_duration",
    "reservation_active",
    "in_queue",
    "path_hint_north",
    "path_hint_south",
    "path_hint_east",
    "path_hint_west",
    "neighbor_agent_ratio",
    "neighbor_object_ratio",
    "reserved_tile_ratio",
    "nearest_agent_distance",
)
assert len(_BASE_FEATURES) == capi.TL_N_BASE_FEATURES


@dataclass
class ObsSpec:
    """Observation layout (config/observations.py; service.py:33-81)."""

    variant: str = "hybrid"
    local_window: int = 11  # hybrid.local_window (hybrid + full)
    compact_window: int = 7  # compact.map_window
    include_targets: bool = False  # of the active variant (service.py:66-70)
    time_ticks_per_day: int = 1440
    object_channels: tuple[str, ...] = ()  # compact, sanitised (service.py:170-180)
    normalize_counts: bool = True
    top_friends: int = 2
    top_rivals: int = 2
    embed_dim: int = 8
    include_aggregates: bool = True
    personality_channels: bool = False
    max_slots: int = 64
    rivalry_max_edges: int = 6
    avoid_threshold: float = 0.7
    relationships_stage: str = "B"

    def __post_init__(self) -> None:
        if self.variant not in capi.TL_VARIANT:
            raise ValueError(f"Unsupported observation variant: {self.variant}")
        if self.relationships_stage == "OFF":
            # service.py:40-44
            raise ValueError("WorldObservationService requires relationships stage to be enabled")
        if self.window % 2 == 0 or not (3 <= self.window <= capi.TL_MAX_WINDOW):
            raise ValueError(f"window must be odd and in 3..{capi.TL_MAX_WINDOW}")
        if len(self.object_channels) > capi.TL_MAX_CTYPES:
            raise ValueError(f"at most {capi.TL_MAX_CTYPES} compact object channels are supported")
        if self.top_friends + self.top_rivals > capi.TL_MAX_SOCIAL_SLOTS:
            raise ValueError("Sum of top_friends and top_rivals must be <= 8 for tensor budget")
        if self.top_friends == 0 and self.top_rivals == 0:
            self.include_aggregates = False  # config/observations.py:41-42
        if not (1 <= self.embed_dim <= 32):
            raise ValueError("embed_dim must be in 1..32")

    # ---- construction from the reference config --------------------------
    @classmethod
    def from_config(cls, config: Any, variant: str | None = None) -> "ObsSpec":
        oc = config.observations_config
        variant = str(variant or config.observation_variant)
        channels: list[str] = []
        for name in oc.compact.object_channels:  # service.py:170-180
            sanitized = str(name).strip().lower()
            if sanitized and sanitized not in channels:
                channels.append(sanitized)
        personality = False
        if hasattr(config, "personality_channels_enabled"):
            try:
                personality = bool(config.personality_channels_enabled())
            except Exception:
                personality = False
        include_targets = oc.compact.include_targets if variant == "compact" else oc.hybrid.include_targets
        return cls(
            variant=variant,
            local_window=int(oc.hybrid.local_window),
            compact_window=int(oc.compact.map_window),
            include_targets=bool(include_targets),
            time_ticks_per_day=int(oc.hybrid.time_ticks_per_day),
            object_channels=tuple(channels),
            normalize_counts=bool(oc.compact.normalize_counts),
            top_friends=int(oc.social_snippet.top_friends),
            top_rivals=int(oc.social_snippet.top_rivals),
            embed_dim=int(oc.social_snippet.embed_dim),
            include_aggregates=bool(oc.social_snippet.include_aggregates),
            personality_channels=personality,
            max_slots=int(config.embedding_allocator.max_slots),
            rivalry_max_edges=int(config.conflict.rivalry.max_edges),
            avoid_threshold=float(config.conflict.rivalry.avoid_threshold),
            relationships_stage=str(config.features.stages.relationships),
        )

    # ---- derived layout ----------------------------------------------------
    @property
    def window(self) -> int:
        return self.compact_window if self.variant == "compact" else self.local_window

    @property
    def radius(self) -> int:
        return self.window // 2

    @property
    def map_channels(self) -> tuple[str, ...]:
        base = ("self", "agents", "objects", "reservations")  # service.py:30
        if self.variant == "hybrid":
            return base
        if self.variant == "full":
            return (*base, "path_dx", "path_dy")
        return (*base, *(f"object:{n}" for n in self.object_channels), "walkable")

    @property
    def n_channels(self) -> int:
        return len(self.map_channels)

    @property
    def map_shape(self) -> tuple[int, int, int]:
        return (self.n_channels, self.window, self.window)

    @property
    def social_slots(self) -> int:
        return max(0, self.top_friends + self.top_rivals)

    @property
    def social_slot_dim(self) -> int:
        return self.embed_dim + 3

    @property
    def social_vector_length(self) -> int:  # service.py:182-187
        length = self.social_slots * self.social_slot_dim
        if self.include_aggregates:
            length += 4
        return length

    @property
    def aggregate_names(self) -> list[str]:
        if self.include_aggregates:
            return ["social_trust_mean", "social_trust_max", "social_rivalry_mean", "social_rivalry_max"]
        return []

    @property
    def feature_names(self) -> tuple[str, ...]:  # service.py:83-168
        names = list(_BASE_FEATURES)
        if self.include_targets:
            for landmark in LANDMARKS:
                names.extend([f"{landmark}_dx", f"{landmark}_dy", f"{landmark}_dist"])
        if self.personality_channels:
            names.extend(["personality_extroversion", "personality_forgiveness", "personality_ambition"])
        for s in range(self.social_slots):
            for d in range(self.social_slot_dim):
                names.append(f"social_slot{s}_d{d}")
        names.extend(self.aggregate_names)
        return tuple(names)

    @property
    def n_features(self) -> int:
        return len(self.feature_names)

    @property
    def landmark_base(self) -> int:
        return capi.TL_N_BASE_FEATURES if self.include_targets else -1

    @property
    def personality_base(self) -> int:
        if not self.personality_channels:
            return -1
        return capi.TL_N_BASE_FEATURES + (12 if self.include_targets else 0)

    @property
    def social_base(self) -> int:
        return self.n_features - self.social_vector_length if self.social_vector_length else -1

    @property
    def landmark_slices(self) -> dict[str, tuple[int, int]]:  # service.py:62-81
        if not self.include_targets:
            return {}
        base = self.landmark_base
        return {name: (base + 3 * i, base + 3 * i + 3) for i, name in enumerate(LANDMARKS)}

    @property
    def embed_words(self) -> int:
        return (self.embed_dim + 7) // 8

    @property
    def map_elems(self) -> int:
        return self.n_channels * self.window * self.window

    # ---- look-up tables ------------------------------------------------------
    def luts(self) -> dict[str, np.ndarray]:
        """float32 tables of the reference's float64 expressions (names = TlObsParams fields)."""
        tpd = self.time_ticks_per_day
        time_lut = np.zeros((tpd, 2), dtype=np.float32)
        for m in range(tpd):  # features.py:142-148
            phase = m / tpd
            time_lut[m, 0] = math.sin(math.tau * phase)
            time_lut[m, 1] = math.cos(math.tau * phase)
        episode_length = max(1, tpd)  # features.py:180-188
        progress = np.zeros(tpd + 1, dtype=np.float32)
        for e in range(tpd + 1):
            progress[e] = max(0.0, min(1.0, float(e) / float(episode_length)))
        slot_lut = np.zeros(self.max_slots, dtype=np.float32)
        for s in range(self.max_slots):  # features.py:175
            slot_lut[s] = float(s) / float(self.max_slots)
        window = self.window
        radius = self.radius
        total_tiles = window * window
        max_tiles = max(1, total_tiles - 1)  # map.py:127
        agent_ratio = np.zeros(total_tiles, dtype=np.float32)
        for n in range(total_tiles):
            agent_ratio[n] = min(1.0, n / max_tiles)
        tile_ratio = np.zeros(total_tiles + 1, dtype=np.float32)
        for n in range(total_tiles + 1):  # map.py:129-130
            tile_ratio[n] = min(1.0, n / max(1, total_tiles))
        nearest = np.zeros(2 * radius * radius + 1, dtype=np.float32)
        seen: dict[int, float] = {}
        for dy in range(-radius, radius + 1):  # scan order of map.py:80-81
            for dx in range(-radius, radius + 1):
                d2 = dx * dx + dy * dy
                distance = float(np.hypot(dx, dy))
                norm = max(0.0, min(1.0, distance / max(1, radius))) if distance > 0 else 0.0
                value = float(np.float32(norm))
                if d2 in seen and seen[d2] != value:
                    raise AssertionError("np.hypot disagrees for equal squared distances")
                seen[d2] = value
                nearest[d2] = value
        path = np.zeros((2, total_tiles), dtype=np.float32)
        for dy in range(-radius, radius + 1):  # map.py:115-125
            for dx in range(-radius, radius + 1):
                distance = float(np.hypot(dx, dy))
                if distance > 0:
                    tile = (dy + radius) * window + (dx + radius)
                    path[0, tile] = float(dx) / distance
                    path[1, tile] = float(dy) / distance
        return {
            "time_lut": time_lut,
            "progress_lut": progress,
            "slot_lut": slot_lut,
            "agent_ratio_lut": agent_ratio,
            "tile_ratio_lut": tile_ratio,
            "nearest_lut": nearest,
            "path_lut": path,
            # values.astype(np.float32) / 255.0 stays float32 (social.py:238-239)
            "embed_lut": (np.arange(256, dtype=np.uint8).astype(np.float32) / 255.0).astype(np.float32),
        }

    def struct(self, lut_ptrs: dict[str, int] | None = None) -> capi.TlObsParams:
        p = capi.TlObsParams()
        p.variant = capi.TL_VARIANT[self.variant]
        p.window = self.window
        p.n_channels = self.n_channels
        p.n_features = self.n_features
        p.n_ctypes = len(self.object_channels) if self.variant == "compact" else 0
        p.normalize_counts = int(self.normalize_counts)
        p.ticks_per_day = self.time_ticks_per_day
        p.max_slots = self.max_slots
        p.rivalry_max_edges = self.rivalry_max_edges
        p.top_friends = self.top_friends
        p.top_rivals = self.top_rivals
        p.embed_dim = self.embed_dim
        p.include_aggregates = int(self.include_aggregates)
        p.landmark_base = self.landmark_base
        p.personality_base = self.personality_base
        p.social_base = self.social_base
        p.path_hint_ltype = LTYPE_CODE["stove"]  # features.py:264
        p.avoid_threshold = self.avoid_threshold
        for name, ptr in (lut_ptrs or {}).items():
            setattr(p, name, C.c_void_p(ptr))
        return p


@dataclass
class RewardSpec:
    """Reward / decay parameters (config/rewards.py:33-69)."""

    decay_rates: dict[str, float] = field(default_factory=lambda: {"hunger": 0.01, "hygiene": 0.005, "energy": 0.008})
    needs_weights: dict[str, float] = field(default_factory=lambda: {"hunger": 1.0, "hygiene": 0.6, "energy": 0.8})
    survival_tick: float = 0.002
    wage_rate: float = 0.01
    punctuality_bonus: float = 0.05
    faint_penalty: float = -1.0
    eviction_penalty: float = -2.0
    clip_per_tick: float = 0.2
    clip_per_episode: float = 50.0
    no_positive_within_death_ticks: int = 10
    personality_scaling: bool = False  # features.behavior.reward_multipliers
    fuse_faint: bool = False
    faint_threshold: float = 0.03  # lifecycle/manager.py:41
    starvation_threshold: float = 0.05  # config/rewards.py:96
    # social reward weights are applied on the host when events are reduced (engine.py:257-317)
    chat_base: float = 0.01
    coeff_trust: float = 0.3
    coeff_fam: float = 0.2
    avoid_conflict: float = 0.005
    social_stage: str = "C1"

    @classmethod
    def from_config(cls, config: Any, fuse_faint: bool = False) -> "RewardSpec":
        r = config.rewards
        scaling = False
        if hasattr(config, "reward_personality_scaling_enabled"):
            scaling = bool(config.reward_personality_scaling_enabled())
        starvation = 0.05
        stability = getattr(config, "stability", None)
        if stability is not None and hasattr(stability, "starvation"):
            starvation = float(stability.starvation.hunger_threshold)
        return cls(
            decay_rates={str(k): float(v) for k, v in r.decay_rates.items()},
            needs_weights={n: float(getattr(r.needs_weights, n, 0.0)) for n in NEEDS},
            survival_tick=float(r.survival_tick),
            wage_rate=float(r.wage_rate),
            punctuality_bonus=float(r.punctuality_bonus),
            faint_penalty=float(r.faint_penalty),
            eviction_penalty=float(r.eviction_penalty),
            clip_per_tick=float(r.clip.clip_per_tick),
            clip_per_episode=float(r.clip.clip_per_episode),
            no_positive_within_death_ticks=int(r.clip.no_positive_within_death_ticks),
            personality_scaling=scaling,
            fuse_faint=fuse_faint,
            starvation_threshold=starvation,
            chat_base=float(r.social.C1_chat_base),
            coeff_trust=float(r.social.C1_coeff_trust),
            coeff_fam=float(r.social.C1_coeff_fam),
            avoid_conflict=float(r.social.C2_avoid_conflict),
            social_stage=str(getattr(config.features.stages, "social_rewards", "OFF")),
        )

    @property
    def social_stage_value(self) -> int:  # engine.py:201-204
        return {"OFF": 0, "C1": 1, "C2": 2, "C3": 3}.get(self.social_stage, 0)

    def struct(self) -> capi.TlRewardParams:
        p = capi.TlRewardParams()
        for k, need in enumerate(NEEDS):
            p.decay_rate[k] = float(self.decay_rates.get(need, 0.0))
            p.need_weight[k] = float(self.needs_weights.get(need, 0.0))
        p.survival_tick = self.survival_tick
        p.wage_rate = self.wage_rate
        p.punctuality_bonus = self.punctuality_bonus
        p.faint_penalty = self.faint_penalty
        p.eviction_penalty = self.eviction_penalty
        p.clip_per_tick = self.clip_per_tick
        p.clip_per_episode = self.clip_per_episode
        p.faint_threshold = self.faint_threshold
        p.starvation_threshold = self.starvation_threshold
        for code in range(capi.TL_MAX_PROFILES):
            for k in range(3):
                p.need_mult[code][k] = 1.0
                p.reward_bias[code][k] = 1.0
        for name, code in PROFILE_CODE.items():
            for k, need in enumerate(NEEDS):
                p.need_mult[code][k] = float(_PROFILE_NEED_MULT.get(name, {}).get(need, 1.0))
            for g, group in enumerate(_BIAS_GROUPS):
                p.reward_bias[code][g] = float(_PROFILE_REWARD_BIAS.get(name, {}).get(group, 1.0))
        p.block_window = int(self.no_positive_within_death_ticks)
        p.decay_personality = int(self.personality_scaling)
        p.reward_personality = int(self.personality_scaling)
        p.fuse_faint = int(self.fuse_faint)
        return p


@dataclass
class DynamicsSpec:
    """World dynamics around the hot path.

    Relationship ties decay with ``RelationshipParameters`` (world/relationships.py:14-21; the service
    builds them with defaults, relationships_service.py:237-238), the rivalry ledger with
    ``conflict.rivalry`` (config/conflict.py:17-24, relationships_service.py:240-249).
    """

    trust_decay: float = 0.0
    familiarity_decay: float = 0.0
    tie_rivalry_decay: float = 0.01
    rivalry_decay_per_tick: float = 0.005
    rivalry_min: float = 0.0
    rivalry_max: float = 1.0
    rivalry_eviction_threshold: float = 0.05

    @classmethod
    def from_config(cls, config: Any) -> "DynamicsSpec":
        cfg = config.conflict.rivalry
        return cls(
            rivalry_decay_per_tick=float(cfg.decay_per_tick),
            rivalry_min=float(cfg.min_value),
            rivalry_max=float(cfg.max_value),
            rivalry_eviction_threshold=float(cfg.eviction_threshold),
        )

    def struct(self) -> capi.TlDynamicsParams:
        from .hashing import last_action_hash

        d = capi.TlDynamicsParams()
        d.trust_decay = self.trust_decay
        d.familiarity_decay = self.familiarity_decay
        d.tie_rivalry_decay = self.tie_rivalry_decay
        d.rivalry_decay_per_tick = self.rivalry_decay_per_tick
        d.rivalry_min = self.rivalry_min
        d.rivalry_max = self.rivalry_max
        d.rivalry_eviction_threshold = self.rivalry_eviction_threshold
        for code, kind in enumerate(capi.ACT_KINDS):
            d.action_hash[code] = last_action_hash(kind)
        return d


def pack_actions(kind, success=0, has_pos=0, x=0, y=0, duration=1):
    """Action words (include/townlet_b200.h: TL_ACT_PACK) from integer arrays or scalars."""
    kind, success, has_pos, x, y, duration = (np.asarray(v, dtype=np.int64) for v in (kind, success, has_pos, x, y, duration))
    if np.any((kind < 0) | (kind >= capi.TL_N_ACT_KINDS)):
        raise ValueError("action kind code out of range")
    if np.any((duration < 0) | (duration > 63)):
        raise ValueError("action duration must lie in 0..63")
    if np.any(has_pos.astype(bool) & ((x < 0) | (x >= capi.TL_MAX_GRID) | (y < 0) | (y >= capi.TL_MAX_GRID))):
        raise ValueError("move target outside 0..1023")
    word = (kind & 15) | ((success & 1) << 4) | ((has_pos & 1) << 5) | ((x & 1023) << 6) | ((y & 1023) << 16) | ((duration & 63) << 26)
    return word.astype(np.uint32)
