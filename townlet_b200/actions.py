"""Action payloads of the reference -> action words of the device tick (``TL_PHASE_ACTIONS``).

The reference applies a tick's actions in ``WorldContext.tick``
(src/townlet/world/core/context.py:246-257): ``apply_actions`` (world/actions.py:44-105) and then
``process_actions`` (world/systems/affordances.py:105-229), whose bookkeeping is what the observation
reads afterwards (``last_action_id`` / ``last_action_success`` / ``last_action_duration``,
world/affordances/core.py:198-206, and the position of a move).  Moves are resolved on the device;
for the kinds that go through host services (queue requests, affordance start / release, chat) the
caller passes the success those services returned.
"""

from __future__ import annotations

from collections.abc import Mapping, Sequence
from typing import Any

import numpy as np

from . import capi
from .params import pack_actions

KIND_CODE = {kind: code for code, kind in enumerate(capi.ACT_KINDS) if kind}


def encode_actions(
    agent_ids: Sequence[str],
    actions: Mapping[str, Any],
    origin: tuple[int, int] = (0, 0),
    success: Mapping[str, bool] | None = None,
) -> np.ndarray:
    """One action word per agent of a town, in ``agent_ids`` order (0 = no action this tick).

    ``actions`` maps agent id -> payload mapping as ``process_actions`` receives it; ``origin`` is the
    town origin chosen at ingest (positions are stored town-relative); ``success`` carries the outcome
    of host-resolved kinds.  Raises ``ValueError`` for a kind the action word cannot name.
    """
    kinds = np.zeros(len(agent_ids), dtype=np.int64)
    ok = np.zeros_like(kinds)
    has_pos = np.zeros_like(kinds)
    xs = np.zeros_like(kinds)
    ys = np.zeros_like(kinds)
    durations = np.ones_like(kinds)
    for i, agent_id in enumerate(agent_ids):
        action = actions.get(agent_id)
        if not isinstance(action, Mapping):  # affordances.py:129-130: silently skipped
            continue
        kind = action.get("kind")
        if not kind:  # affordances.py:212: nothing is recorded without a kind
            continue
        if kind not in KIND_CODE:
            raise ValueError(f"action kind {kind!r} has no action-word code")
        kinds[i] = KIND_CODE[kind]
        durations[i] = int(action.get("duration", 1))
        if kind == "move":
            position = action.get("position")
            if position:  # affordances.py:157
                has_pos[i] = 1
                xs[i] = int(position[0]) - origin[0]
                ys[i] = int(position[1]) - origin[1]
        elif success is not None:
            ok[i] = 1 if success.get(agent_id, False) else 0
    return pack_actions(kinds, ok, has_pos, xs, ys, durations)
