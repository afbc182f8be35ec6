"""Duck-typed stand-ins for the reference world / config (no reference install needed).

``FakeWorld`` offers the read-only surface the hot path reads —
``WorldRuntimeAdapterProtocol`` (src/townlet/world/observations/interfaces.py:43-107)
plus ``agents`` / ``tick`` / event consumers of ``WorldState`` — built from a
synthetic ``TownDescription``.
"""

from __future__ import annotations

from types import MappingProxyType, SimpleNamespace
from typing import Any

from townlet_b200.synth import TownDescription


def fake_config(variant: str = "hybrid", **obs_overrides: Any) -> SimpleNamespace:
    hybrid = SimpleNamespace(local_window=11, include_targets=False, time_ticks_per_day=1440)
    compact = SimpleNamespace(map_window=7, include_targets=False, object_channels=[], normalize_counts=True)
    social = SimpleNamespace(top_friends=2, top_rivals=2, embed_dim=8, include_aggregates=True)
    for key, value in obs_overrides.items():
        section, name = key.split("__")
        setattr({"hybrid": hybrid, "compact": compact, "social": social}[section], name, value)
    rewards = SimpleNamespace(
        needs_weights=SimpleNamespace(hunger=1.0, hygiene=0.6, energy=0.8),
        decay_rates={"hunger": 0.01, "hygiene": 0.005, "energy": 0.008},
        punctuality_bonus=0.05,
        wage_rate=0.01,
        survival_tick=0.002,
        faint_penalty=-1.0,
        eviction_penalty=-2.0,
        social=SimpleNamespace(C1_chat_base=0.01, C1_coeff_trust=0.3, C1_coeff_fam=0.2, C2_avoid_conflict=0.005),
        clip=SimpleNamespace(clip_per_tick=0.2, clip_per_episode=50, no_positive_within_death_ticks=10),
    )
    config = SimpleNamespace(
        observation_variant=variant,
        observations_config=SimpleNamespace(hybrid=hybrid, compact=compact, social_snippet=social),
        embedding_allocator=SimpleNamespace(max_slots=64),
        conflict=SimpleNamespace(rivalry=SimpleNamespace(max_edges=6, avoid_threshold=0.7)),
        features=SimpleNamespace(stages=SimpleNamespace(relationships="B", social_rewards="C2")),
        rewards=rewards,
        stability=SimpleNamespace(starvation=SimpleNamespace(hunger_threshold=0.05)),
    )
    config.personality_channels_enabled = lambda: False
    config.reward_personality_scaling_enabled = lambda: False
    return config


class _Allocator:
    """Slot allocation in first-come order (observations/embedding.py:130-139, fresh allocator)."""

    def __init__(self) -> None:
        self.assignments: dict[str, int] = {}
        self.released: list[str] = []

    def allocate(self, agent_id: str, tick: int) -> int:
        return self.assignments.setdefault(agent_id, len(self.assignments))

    def release(self, agent_id: str, tick: int) -> None:
        self.released.append(agent_id)

    def has_assignment(self, agent_id: str) -> bool:
        return agent_id in self.assignments and agent_id not in self.released


class _QueueManager:
    def __init__(self, queues: dict[str, list[str]]) -> None:
        self._queues = queues

    def queue_snapshot(self, object_id: str) -> list[str]:
        return list(self._queues.get(object_id, []))


class FakeWorld:
    def __init__(self, town: TownDescription, config: Any) -> None:
        self.config = config
        self.tick = town.tick
        self.town = town
        self.agents: dict[str, SimpleNamespace] = {}
        for i, aid in enumerate(town.agent_ids):
            self.agents[aid] = SimpleNamespace(
                agent_id=aid,
                position=(int(town.positions[i, 0]), int(town.positions[i, 1])),
                needs={"hunger": float(town.needs[i, 0]), "hygiene": float(town.needs[i, 1]), "energy": float(town.needs[i, 2])},
                wallet=float(town.wallet[i]),
                personality=SimpleNamespace(
                    extroversion=float(town.personality[i, 0]),
                    forgiveness=float(town.personality[i, 1]),
                    ambition=float(town.personality[i, 2]),
                ),
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
        self._wages = {aid: float(town.wages_paid[i]) for i, aid in enumerate(town.agent_ids)}
        self._punct = {aid: float(town.punctuality[i]) for i, aid in enumerate(town.agent_ids)}
        queues: dict[str, list[str]] = {}
        for i, aid in enumerate(town.agent_ids):
            if town.in_queue[i]:
                queues.setdefault(town.object_ids[i % len(town.object_ids)], []).append(aid)
        self.queue_manager = _QueueManager(queues)
        self.embedding_allocator = _Allocator()
        self._ctx_resets: set[str] = set()
        self.ctx_reset_consumed = 0
        self.chat_events: list[dict[str, object]] = []
        self.avoidance_events: list[dict[str, object]] = []
        self.events_consumed = 0
        self._relationships = object()
        self._employment_service = object()
        self._economy_service = object()

    # ---- WorldRuntimeAdapterProtocol -------------------------------------
    def agent_snapshots_view(self):
        return MappingProxyType(self.agents)

    def agent_position(self, agent_id: str):
        return self.agents[agent_id].position

    def consume_ctx_reset_requests(self) -> set[str]:
        self.ctx_reset_consumed += 1
        pending, self._ctx_resets = set(self._ctx_resets), set()
        return pending

    def request_ctx_reset(self, agent_id: str) -> None:
        self._ctx_resets.add(agent_id)

    def objects_snapshot(self):
        t = self.town
        return MappingProxyType(
            {
                oid: {"object_type": t.object_types[i], "position": (int(t.object_positions[i, 0]), int(t.object_positions[i, 1]))}
                for i, oid in enumerate(t.object_ids)
            }
        )

    def objects_by_position_view(self):
        out: dict[tuple[int, int], tuple[str, ...]] = {}
        t = self.town
        for i, oid in enumerate(t.object_ids):
            pos = (int(t.object_positions[i, 0]), int(t.object_positions[i, 1]))
            out[pos] = (*out.get(pos, ()), oid)
        return MappingProxyType(out)

    def reservation_tiles(self):
        t = self.town
        index = {oid: i for i, oid in enumerate(t.object_ids)}
        return frozenset((int(t.object_positions[index[o], 0]), int(t.object_positions[index[o], 1])) for o in t.reservations)

    @property
    def active_reservations(self):
        return MappingProxyType(dict(self.town.reservations))

    def relationships_snapshot(self):
        t = self.town
        return {
            aid: {other: {"trust": tr, "familiarity": fa, "rivalry": rv} for other, tr, fa, rv in t.ties[i]}
            for i, aid in enumerate(t.agent_ids)
            if t.ties[i]
        }

    def rivalry_top(self, agent_id: str, limit: int):
        i = self.town.agent_ids.index(agent_id)
        return sorted(self.town.rivalry[i], key=lambda item: item[1], reverse=True)[:limit]

    def relationship_tie(self, subject: str, target: str):
        i = self.town.agent_ids.index(subject)
        for other, tr, fa, rv in self.town.ties[i]:
            if other == target:
                return SimpleNamespace(trust=tr, familiarity=fa, rivalry=rv)
        return None

    def _employment_context_wages(self, agent_id: str) -> float:
        return self._wages.get(agent_id, 0.0)

    def _employment_context_punctuality(self, agent_id: str) -> float:
        return self._punct.get(agent_id, 0.0)

    # ---- WorldState event buffers (grid.py:1073-1081) ---------------------
    def consume_chat_events(self):
        self.events_consumed += 1
        events, self.chat_events = self.chat_events, []
        return events

    def consume_relationship_avoidance_events(self):
        events, self.avoidance_events = self.avoidance_events, []
        return events
