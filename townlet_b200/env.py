"""ParallelEnv-shaped wrapper over the batched tensors.

The reference declares a PettingZoo dependency but ships no environment class
(SURVEY.md §0 fact 2); its closest surface is ``PolicyRuntime``
(src/townlet/policy/runner.py:93-100) and ``WorldRuntime.observe/tick``
(src/townlet/ports/world.py:63-84).  This class offers the ParallelEnv calling
convention (``reset`` / ``step`` returning observations, rewards, terminations,
truncations, infos) over all towns of a device batch at once, duck-typed so
that ``pettingzoo`` itself is not required.

Actions: with a ``DynamicsSpec`` the tick applies device-resident action words itself (moves and
outcome bookkeeping, ``TL_PHASE_ACTIONS``) and decays the relationship / rivalry ledgers
(``TL_PHASE_SOCIAL_DECAY``) in ``WorldContext.tick`` order (world/core/context.py:229-289); pass the words
to ``step`` (``townlet_b200.actions.encode_actions`` / ``params.pack_actions``).  Kinds that go through host
services (queues, affordances, chat) are resolved by the caller, who may also pass ``apply_actions`` to
mutate the device batch before the tick.
"""

from __future__ import annotations

from collections.abc import Callable
from typing import Any

import torch

from . import capi
from .engine import DeviceTownBatch, TickOutputs
from .params import DynamicsSpec, ObsSpec, RewardSpec
from .soa import TownBatch


class TownletParallelEnv:
    """All agents of all towns step together; tensors stay on the device."""

    metadata = {"name": "townlet_b200_v0", "is_parallelizable": True}

    def __init__(
        self,
        batch: TownBatch,
        obs: ObsSpec,
        rew: RewardSpec,
        device: str | torch.device = "cuda:0",
        apply_actions: Callable[[DeviceTownBatch, Any], None] | None = None,
        max_ticks: int | None = None,
        dyn: DynamicsSpec | None = None,
    ) -> None:
        self._initial = batch.copy()
        self.obs_spec = obs
        self.rew_spec = rew
        self.device = torch.device(device)
        self._apply_actions = apply_actions
        self._max_ticks = max_ticks
        self._dev = DeviceTownBatch(batch.copy(), obs, rew, self.device, dyn=dyn)
        self._world_phases = dyn is not None
        self._no_actions = torch.zeros(batch.n_agents, dtype=torch.int32, device=self.device) if dyn is not None else None
        self._out: TickOutputs = self._dev.alloc_outputs(capi.TL_PHASE_ALL, 1)
        self._ticks = 0
        self.num_agents = batch.n_agents
        self.possible_agents = list(range(batch.n_agents))  # row index = agent handle
        self.agents = list(self.possible_agents)

    # ---- spaces (shapes only; gymnasium is optional) ----------------------
    @property
    def observation_shapes(self) -> dict[str, tuple[int, ...]]:
        return {"map": self.obs_spec.map_shape, "features": (self.obs_spec.n_features,)}

    def _observations(self) -> dict[str, torch.Tensor]:
        assert self._out.map is not None and self._out.features is not None
        return {"map": self._out.map[0], "features": self._out.features[0]}

    def reset(self, seed: int | None = None, options: dict | None = None):
        del seed, options
        self._dev.upload(self._initial)
        self._ticks = 0
        self._dev.tick(self._out, capi.TL_PHASE_OBS, 1)
        return self._observations(), {}

    def step(self, actions: Any = None):
        """Advance every town one tick.  ``actions``: int32 device tensor ``[n_agents]`` of action words when
        the env was built with a ``DynamicsSpec`` (``None`` = nobody acts); otherwise whatever the
        ``apply_actions`` callback expects."""
        if self._apply_actions is not None:
            self._apply_actions(self._dev, actions)
        if self._world_phases:
            words = actions if isinstance(actions, torch.Tensor) else self._no_actions
            self._dev.tick(self._out, capi.TL_PHASE_WORLD, 1, actions=words)
        else:
            self._dev.tick(self._out, capi.TL_PHASE_ALL, 1)
        self._ticks += 1
        assert self._out.reward is not None and self._out.terminated is not None and self._out.kpi is not None
        terminations = self._out.terminated[0].bool()
        truncated = self._max_ticks is not None and self._ticks >= self._max_ticks
        truncations = torch.full_like(terminations, bool(truncated))
        infos = {"kpi": self._out.kpi[0], "tick": self._ticks}
        return self._observations(), self._out.reward[0], terminations, truncations, infos

    def state(self) -> DeviceTownBatch:
        return self._dev

    def close(self) -> None:
        pass
