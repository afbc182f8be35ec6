"""Drop-in for the reference's ``RewardEngine`` and the world's need-decay hook.

``B200RewardEngine`` keeps the reference's surface —
``compute(world, terminated, reasons) -> dict[str, RewardBreakdown]``,
``latest_reward_breakdown()``, ``latest_social_events()`` and the two pieces of
per-agent state ``_termination_block`` / ``_episode_totals``
(src/townlet/rewards/engine.py:25-47, 426-433) — and runs the per-agent
arithmetic in the CUDA library.  Sparse chat / avoidance events are folded to
per-agent sums on the host (``reward_events``), exactly once per call, like the
reference's destructive reads (engine.py:52-53).

``install_need_decay`` swaps ``WorldState._apply_need_decay``
(src/townlet/world/grid.py:1439-1465) for the device path while keeping the
non-decay tail of that method (employment / economy bookkeeping).
"""

from __future__ import annotations

import sys
from collections.abc import Mapping
from dataclasses import dataclass, field
from typing import Any

from . import capi
from .engine import HostEngine
from .ingest import ingest_town, rows_to_batch
from .params import NEEDS, ObsSpec, RewardSpec
from .reward_events import normalise_events, reduce_social_events


@dataclass
class _PlainBreakdown:
    """Used only when ``townlet.dto.rewards`` is not importable (stand-alone tests).

    Field names follow ``RewardBreakdown`` (src/townlet/dto/rewards.py:17-60).
    """

    agent_id: str
    tick: int
    total: float
    homeostasis: float = 0.0
    work: float = 0.0
    social: float = 0.0
    shaping: float = 0.0
    survival: float = 0.0
    needs_penalty: float = 0.0
    wage: float = 0.0
    punctuality: float = 0.0
    social_bonus: float = 0.0
    social_penalty: float = 0.0
    social_avoidance: float = 0.0
    terminal_penalty: float = 0.0
    clip_adjustment: float = 0.0
    needs: dict[str, float] = field(default_factory=dict)
    guardrail_active: bool = False
    clipped: bool = False


def _breakdown_cls() -> Any:
    try:
        from townlet.dto.rewards import RewardBreakdown  # type: ignore

        return RewardBreakdown
    except ImportError:
        return _PlainBreakdown


_DUMMY_OBS = ObsSpec()  # the reward path needs no observation parameters; embed_words of the default layout


class B200RewardEngine:
    """``RewardEngine`` whose per-agent loop (engine.py:71-172) runs on the GPU."""

    def __init__(self, config: Any, device: int = 0) -> None:
        self.config = config
        self.spec = RewardSpec.from_config(config)
        self._termination_block: dict[str, int] = {}
        self._episode_totals: dict[str, float] = {}
        self._latest_breakdown: dict[str, dict[str, float]] = {}
        self._latest_social_events: list[dict[str, object]] = []
        self._reward_scaling_enabled = self.spec.personality_scaling
        self._engine = HostEngine(None, self.spec, device)
        self._breakdown = _breakdown_cls()

    # ------------------------------------------------------------------
    def compute(
        self,
        world: Any,
        terminated: dict[str, bool],
        reasons: Mapping[str, str] | None = None,
    ) -> dict[str, Any]:
        current_tick = int(getattr(world, "tick", 0))
        reason_map = dict(reasons or {})
        block_window = int(self.spec.no_positive_within_death_ticks)
        self._prune_termination_blocks(current_tick, block_window)
        self._reset_episode_totals(terminated, world)

        chat_events = normalise_events(self._consume(world, "consume_chat_events"))
        avoidance_events = normalise_events(self._consume(world, "consume_relationship_avoidance_events"))
        sums = reduce_social_events(world, chat_events, avoidance_events, self.spec)

        breakdowns: dict[str, Any] = {}
        legacy: dict[str, dict[str, float]] = {}
        agent_ids = list(world.agents)
        if agent_ids:
            rows = ingest_town(
                _RewardView(world),
                _DUMMY_OBS,
                terminated=terminated,
                reasons=reason_map,
                episode_totals=self._episode_totals,
                termination_block=self._termination_block,
                social_sums=sums,
                observe=False,
            )
            batch = rows_to_batch([rows], _DUMMY_OBS, grid=(16, 16))
            out = self._engine.alloc_outputs(batch, capi.TL_PHASE_REWARD, 1, comps=True)
            self._engine.tick(batch, out, capi.TL_PHASE_REWARD, 1)
            comps = out["comps"][0]
            idx = {name: i for i, name in enumerate(capi.COMP_NAMES)}
            for i, agent_id in enumerate(rows.agent_ids):
                row = comps[i]
                snapshot = world.agents[agent_id]
                self._episode_totals[agent_id] = float(batch.episode_total[i])
                block = int(batch.term_block_tick[i])
                if block == capi.TL_TERM_BLOCK_NONE:
                    self._termination_block.pop(agent_id, None)
                else:
                    self._termination_block[agent_id] = block
                components = {
                    name: float(row[idx[name]])
                    for name in (
                        "survival",
                        "needs_penalty",
                        "social_bonus",
                        "social_penalty",
                        "social_avoidance",
                        "social",
                        "wage",
                        "punctuality",
                        "terminal_penalty",
                        "clip_adjustment",
                        "total",
                    )
                }
                breakdowns[agent_id] = self._breakdown(
                    agent_id=agent_id,
                    tick=current_tick,
                    total=components["total"],
                    homeostasis=float(row[idx["homeostasis"]]),
                    work=float(row[idx["work"]]),
                    social=components["social"],
                    shaping=0.0,
                    survival=components["survival"],
                    needs_penalty=components["needs_penalty"],
                    wage=components["wage"],
                    punctuality=components["punctuality"],
                    social_bonus=components["social_bonus"],
                    social_penalty=components["social_penalty"],
                    social_avoidance=components["social_avoidance"],
                    terminal_penalty=components["terminal_penalty"],
                    clip_adjustment=components["clip_adjustment"],
                    needs=dict(snapshot.needs),
                    guardrail_active=bool(row[idx["guardrail_active"]] != 0.0),
                    clipped=bool(row[idx["clipped"]] != 0.0),
                )
                legacy[agent_id] = components
        self._latest_breakdown = legacy
        self._latest_social_events = self._prepare_social_event_log(chat_events, avoidance_events)
        return breakdowns

    # ---- bookkeeping that stays on the host (engine.py:363-373, 415-424) ----
    @staticmethod
    def _consume(world: Any, name: str) -> object:
        consumer = getattr(world, name, None)
        return consumer() if callable(consumer) else ()

    def _prune_termination_blocks(self, tick: int, window: int) -> None:
        if window < 0:
            self._termination_block.clear()
            return
        for agent_id in [a for a, start in self._termination_block.items() if tick - start > window]:
            self._termination_block.pop(agent_id, None)

    def _reset_episode_totals(self, terminated: Mapping[str, bool], world: Any) -> None:
        for agent_id, flag in terminated.items():
            if flag:
                self._episode_totals.pop(agent_id, None)
        for agent_id in set(self._episode_totals) - set(world.agents):
            self._episode_totals.pop(agent_id, None)

    @staticmethod
    def _prepare_social_event_log(chat_events, avoidance_events) -> list[dict[str, object]]:
        log: list[dict[str, object]] = []
        for event in chat_events:  # engine.py:319-346
            etype = str(event.get("event", "chat"))
            if etype not in {"chat_success", "chat_failure"}:
                continue
            log.append(
                {"type": etype, "speaker": event.get("speaker"), "listener": event.get("listener"), "quality": event.get("quality")}
            )
        for event in avoidance_events:
            log.append(
                {"type": "rivalry_avoidance", "agent": event.get("agent"), "object": event.get("object_id"), "reason": event.get("reason")}
            )
        return log

    def latest_reward_breakdown(self) -> dict[str, dict[str, float]]:
        return {agent: dict(components) for agent, components in self._latest_breakdown.items()}

    def latest_social_events(self) -> list[dict[str, object]]:
        return [dict(event) for event in self._latest_social_events]


class _RewardView:
    """Minimal adapter view over a ``WorldState`` for the reward-only ingest.

    ``RewardEngine.compute`` reads ``world.agents``, ``world.tick`` and the
    employment context hooks (engine.py:375-398 via observations/context.py:34-70).
    """

    def __init__(self, world: Any) -> None:
        self._world = world

    @property
    def tick(self) -> int:
        return int(getattr(self._world, "tick", 0))

    def agent_snapshots_view(self) -> Mapping[str, Any]:
        return self._world.agents

    def _reference_context(self, agent_id: str) -> Mapping[str, Any] | None:
        """The reference engine's own seam: ``observation_agent_context`` as bound in ``townlet.rewards.engine``
        (engine.py:12, read at :383 and :396; the reference's tests monkeypatch exactly this name)."""
        module = sys.modules.get("townlet.rewards.engine")
        hook = getattr(module, "observation_agent_context", None) if module is not None else None
        if hook is None:
            return None
        return hook(self._world, agent_id)

    def _employment_context_wages(self, agent_id: str) -> float:
        ctx = self._reference_context(agent_id)
        if ctx is not None:
            return _coerce_float(ctx.get("wages_paid", 0.0))
        service = getattr(self._world, "employment_service", None)
        if service is not None:
            return service.context_wages(agent_id)
        return self._world._employment_context_wages(agent_id)

    def _employment_context_punctuality(self, agent_id: str) -> float:
        ctx = self._reference_context(agent_id)
        if ctx is not None:
            return _coerce_float(ctx.get("punctuality_bonus", 0.0))
        service = getattr(self._world, "employment_service", None)
        if service is not None:
            return service.context_punctuality(agent_id)
        return self._world._employment_context_punctuality(agent_id)


def _coerce_float(value: Any, default: float = 0.0) -> float:
    """utils/coerce.py: coerce_float — anything that is not a number counts as the default."""
    try:
        return float(value)
    except (TypeError, ValueError):
        return default


def install_need_decay(world: Any, device: int = 0) -> Any:
    """Replace ``world._apply_need_decay`` (grid.py:1439) by the device path.

    Returns the new zero-argument callable.  The affordance environment captured
    the bound method at construction (world/affordances/core.py:43, grid.py:386),
    so its field is re-pointed as well when present.
    """
    spec = RewardSpec.from_config(world.config)
    engine = HostEngine(None, spec, device)

    def apply_need_decay() -> None:
        if world.agents:
            rows = ingest_town(_RewardView(world), _DUMMY_OBS, observe=False)
            batch = rows_to_batch([rows], _DUMMY_OBS, grid=(16, 16))
            out = engine.alloc_outputs(batch, capi.TL_PHASE_DECAY, 1, comps=False)
            engine.tick(batch, out, capi.TL_PHASE_DECAY, 1)
            for i, agent_id in enumerate(rows.agent_ids):
                needs = world.agents[agent_id].needs
                for k, need in enumerate(NEEDS):
                    if need in needs and need in spec.decay_rates:
                        needs[need] = float(batch.needs[k, i])
        # the tail of the reference method (grid.py:1456-1465) is not need decay; keep it
        if getattr(world, "_relationships", None) is None:
            raise RuntimeError("Relationship service missing; modular decay requires RelationshipService")
        if getattr(world, "_employment_service", None) is None:
            coordinator = getattr(world, "employment", None)
            if coordinator is not None and hasattr(coordinator, "apply_job_state"):
                coordinator.apply_job_state(world)
        if getattr(world, "_economy_service", None) is None and hasattr(world, "_update_basket_metrics"):
            world._update_basket_metrics()

    world._apply_need_decay = apply_need_decay
    # WorldState keeps the environment on its affordance service (grid.py:411-416, affordances/service.py:29)
    service = getattr(world, "_affordance_service", None)
    env = getattr(service, "environment", None)
    if env is not None and hasattr(env, "apply_need_decay"):
        try:
            env.apply_need_decay = apply_need_decay
        except Exception:  # pragma: no cover - frozen dataclass
            pass
    return apply_need_decay


__all__ = ["B200RewardEngine", "install_need_decay"]
