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

The two other device-resident pieces of a tick sit in the same step loop when asked for:
``employment=(EmploymentSpec, EmploymentArrays)`` runs the shift state machine (``tl_employment_step``,
``EmploymentEngine.apply_job_state``) for the tick about to be computed before the tick kernel reads wallet, wages and
shift state — its per-agent event bits come back in ``infos["employment_events"]`` for the host services that own
the two coworker ledger updates; a batch built with reservation slots (``ingest_town(..., device_reservations=True)``,
``ent_holder``) gets ``TL_PHASE_RESERVATIONS``, the device form of ``refresh_reservations`` (world/systems/queues.py:14-31).
"""

from __future__ import annotations

from collections.abc import Callable
from typing import Any

import torch

from . import capi
from .employment import DeviceEmployment, EmploymentArrays, EmploymentSpec
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
        employment: tuple[EmploymentSpec, EmploymentArrays | None] | None = None,
        device_reservations: bool | None = None,
        town_origin: Any = None,
    ) -> None:
        """``device_reservations``: refresh the reservation slots on the device every tick (default: whenever the batch
        carries holders, i.e. was ingested with ``device_reservations=True``).  ``town_origin``: ``[towns, 2]`` world
        coordinates of every town's origin (job locations of the employment spec are world coordinates)."""
        self._initial = batch.copy()
        self.obs_spec = obs
        self.rew_spec = rew
        self.device = torch.device(device)
        self._apply_actions = apply_actions
        self._max_ticks = max_ticks
        self._dev = DeviceTownBatch(batch.copy(), obs, rew, self.device, dyn=dyn)
        self._world_phases = dyn is not None
        if device_reservations is None:
            device_reservations = bool((batch.ent_holder != -2).any())  # -2 = the batch has no reservation slots
        self._extra_phases = capi.TL_PHASE_RESERVATIONS if device_reservations else 0
        self._employment_spec = employment
        self._town_origin = town_origin
        self._employment: DeviceEmployment | None = None
        self._make_employment()
        self._no_actions = torch.zeros(batch.n_agents, dtype=torch.int32, device=self.device) if dyn is not None else None
        self._out: TickOutputs = self._dev.alloc_outputs(capi.TL_PHASE_ALL, 1)
        self._ticks = 0
        self.num_agents = batch.n_agents
        self.possible_agents = list(range(batch.n_agents))  # row index = agent handle
        self.agents = list(self.possible_agents)

    def _make_employment(self) -> None:
        if self._employment_spec is None:
            return
        spec, arrays = self._employment_spec
        initial = arrays.copy() if arrays is not None else EmploymentArrays(self._dev.n_agents)
        self._employment = DeviceEmployment(self._dev, spec, initial, town_origin=self._town_origin)

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
        self._make_employment()
        self._ticks = 0
        self._dev.tick(self._out, capi.TL_PHASE_OBS, 1)
        return self._observations(), {}

    def step(self, actions: Any = None):
        """Advance every town one tick.  ``actions``: int32 device tensor ``[n_agents]`` of action words when
        the env was built with a ``DynamicsSpec`` (``None`` = nobody acts); otherwise whatever the
        ``apply_actions`` callback expects."""
        if self._apply_actions is not None:
            self._apply_actions(self._dev, actions)
        events = None
        if self._employment is not None:  # the shift state machine of the tick about to be computed (town tick + 1)
            events = self._employment.step(1, tick_offset=1)
        if self._world_phases:
            words = actions if isinstance(actions, torch.Tensor) else self._no_actions
            self._dev.tick(self._out, capi.TL_PHASE_WORLD | self._extra_phases, 1, actions=words)
        else:
            self._dev.tick(self._out, capi.TL_PHASE_ALL | self._extra_phases, 1)
        self._ticks += 1
        assert self._out.reward is not None and self._out.terminated is not None and self._out.kpi is not None
        terminations = self._out.terminated[0].bool()
        truncated = self._max_ticks is not None and self._ticks >= self._max_ticks
        truncations = torch.full_like(terminations, bool(truncated))
        infos = {"kpi": self._out.kpi[0], "tick": self._ticks}
        if events is not None:
            infos["employment_events"] = events[0]  # townlet_b200.employment.EVENT_BITS per agent
        return self._observations(), self._out.reward[0], terminations, truncations, infos

    def state(self) -> DeviceTownBatch:
        return self._dev

    def employment_state(self) -> DeviceEmployment | None:
        return self._employment

    def close(self) -> None:
        pass
