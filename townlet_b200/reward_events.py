"""Host pre-reduction of sparse social events to dense per-agent sums.

Chat / avoidance events are a short per-town list; the reference folds them
into per-agent ``social_bonus`` / ``social_penalty`` / ``social_avoidance``
before its per-agent loop (src/townlet/rewards/engine.py:52-63).  The same
fold runs here on the host and its result feeds the reward kernel through
``TlTownBatch.social_in``.
"""

from __future__ import annotations

from collections.abc import Iterable, Mapping
from typing import Any

from .params import RewardSpec


def _coerce_float(value: object, default: float) -> float:
    # utils/coerce.py:29-50
    if isinstance(value, bool):
        return float(value)
    if isinstance(value, (int, float)):
        return float(value)
    try:
        return float(value)  # type: ignore[arg-type]
    except (TypeError, ValueError):
        return default


def normalise_events(events: object) -> list[dict[str, object]]:
    """engine.py:435-444"""
    if not isinstance(events, Iterable):
        return []
    out: list[dict[str, object]] = []
    for event in events:
        if isinstance(event, Mapping):
            out.append({str(key): value for key, value in event.items()})
    return out


def needs_override(snapshot: Any) -> bool:
    """engine.py:348-353: no chat bonus while any need deficit exceeds 0.85."""
    for value in snapshot.needs.values():
        deficit = 1.0 - _coerce_float(value, 1.0)
        if deficit > 0.85:
            return True
    return False


def chat_rewards(
    world: Any, events: Iterable[dict[str, object]], spec: RewardSpec
) -> tuple[dict[str, float], dict[str, float]]:
    """engine.py:257-301 (``_compute_chat_rewards``)."""
    bonus: dict[str, float] = {}
    penalties: dict[str, float] = {}
    base = float(spec.chat_base)
    coeff_trust = float(spec.coeff_trust)
    coeff_fam = float(spec.coeff_fam)
    for event in events:
        kind = event.get("event")
        speaker = str(event.get("speaker")) if event.get("speaker") else None
        listener = str(event.get("listener")) if event.get("listener") else None
        if kind != "chat_success":
            if kind == "chat_failure":
                if not speaker or not listener:
                    continue
                penalty = -0.5 * base
                penalties[speaker] = penalties.get(speaker, 0.0) + penalty
                penalties[listener] = penalties.get(listener, 0.0) + penalty * 0.5
            continue
        if not speaker or not listener:
            continue
        quality = _coerce_float(event.get("quality", 1.0), 1.0)
        quality_factor = max(0.0, min(1.0, quality))
        for subject, target in ((speaker, listener), (listener, speaker)):
            snapshot = world.agents.get(subject)
            if snapshot is None:
                continue
            if needs_override(snapshot):
                continue
            tie = world.relationship_tie(subject, target)
            trust = tie.trust if tie is not None else 0.0
            familiarity = tie.familiarity if tie is not None else 0.0
            reward = base + coeff_trust * trust + coeff_fam * familiarity
            if quality_factor != 1.0:
                reward *= quality_factor
            bonus[subject] = bonus.get(subject, 0.0) + reward
    return bonus, penalties


def avoidance_rewards(world: Any, events: Iterable[dict[str, object]], spec: RewardSpec) -> dict[str, float]:
    """engine.py:303-317 (``_compute_avoidance_rewards``)."""
    rewards: dict[str, float] = {}
    value = float(spec.avoid_conflict)
    if value == 0.0:
        return rewards
    for event in events:
        agent_id = str(event.get("agent")) if event.get("agent") else None
        if not agent_id or agent_id not in world.agents:
            continue
        rewards[agent_id] = rewards.get(agent_id, 0.0) + value
    return rewards


def reduce_social_events(
    world: Any,
    chat_events: list[dict[str, object]],
    avoidance_events: list[dict[str, object]],
    spec: RewardSpec,
) -> tuple[dict[str, float], dict[str, float], dict[str, float]]:
    """Stage gating of engine.py:58-63."""
    bonus: dict[str, float] = {}
    penalty: dict[str, float] = {}
    avoid: dict[str, float] = {}
    if spec.social_stage_value >= 1:
        bonus, penalty = chat_rewards(world, chat_events, spec)
    if spec.social_stage_value >= 2:
        avoid = avoidance_rewards(world, avoidance_events, spec)
    return bonus, penalty, avoid
