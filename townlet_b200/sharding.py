"""Multi-GPU partition of the town axis and the end-of-rollout KPI gather.

Towns never interact (every accessor of the path is scoped to one
``WorldState``), so a shard is a contiguous block of towns and
the tick needs NO collective.  The only communication is the optional
end-of-rollout statistics exchange: an ``all_reduce(sum)`` of a small per-rank
accumulator and / or an ``all_gather`` of per-town KPI rows — NCCL over NVLink
on GPUs, gloo in the CPU tests.  The reference itself has nothing distributed (one process, one
``SimulationLoop``, src/townlet/core/sim_loop.py:632); its per-town KPIs are the telemetry publisher's
(src/townlet/telemetry/publisher.py:2671-2694), which is what the gathered rows feed.
"""

from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .soa import TownBatch


def town_range(n_towns: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous block [t0, t1) of towns owned by ``rank`` (sizes differ by at most one)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(n_towns, world_size)
    t0 = rank * base + min(rank, extra)
    t1 = t0 + base + (1 if rank < extra else 0)
    return t0, t1


def shard_batch(batch: TownBatch, rank: int, world_size: int) -> TownBatch:
    """The rank's block of towns as its own batch (offsets rebased to zero)."""
    t0, t1 = town_range(batch.n_towns, rank, world_size)
    return batch.slice_towns(t0, t1)


def shard_actions(actions: "np.ndarray | torch.Tensor", batch: TownBatch, rank: int, world_size: int):
    """The rank's columns of action words ``[..., n_agents]`` (agents are town-major, so a shard is a slice)."""
    t0, t1 = town_range(batch.n_towns, rank, world_size)
    a0, a1 = int(batch.town_agent_off[t0]), int(batch.town_agent_off[t1])
    return actions[..., a0:a1]


def reduce_kpi_totals(kpi: torch.Tensor, group: dist.ProcessGroup | None = None) -> torch.Tensor:
    """Sum per-town KPI rows ``[..., T_local, K]`` over towns and ranks -> ``[K]`` float64.

    Additive columns only make sense summed (reward_sum, reward_sumsq, starving,
    terminated); means / minima should be gathered with ``gather_kpi_rows``.
    """
    total = kpi.reshape(-1, kpi.shape[-1]).sum(dim=0).to(torch.float64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    return total


def gather_kpi_rows(kpi: torch.Tensor, n_towns_total: int, group: dist.ProcessGroup | None = None) -> torch.Tensor:
    """All-gather per-town KPI rows ``[T_local, K]`` into the global ``[T, K]`` order.

    Shards may differ in size by one town, so rows are padded to the largest
    shard for the collective and trimmed afterwards.
    """
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return kpi
    world = dist.get_world_size(group)
    sizes = [town_range(n_towns_total, r, world) for r in range(world)]
    pad = max(t1 - t0 for t0, t1 in sizes)
    local = torch.zeros((pad, kpi.shape[-1]), dtype=kpi.dtype, device=kpi.device)
    local[: kpi.shape[0]] = kpi
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local, group=group)
    return torch.cat([g[: t1 - t0] for g, (t0, t1) in zip(gathered, sizes)], dim=0)
