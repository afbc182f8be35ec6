"""CPU, world_size 2 over gloo: town partition and the end-of-rollout KPI exchange."""

from __future__ import annotations

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from townlet_b200 import capi
from townlet_b200.params import ObsSpec, RewardSpec
from townlet_b200.sharding import gather_kpi_rows, reduce_kpi_totals, shard_actions, shard_batch, town_range
from townlet_b200.synth import synth_batch

N_TOWNS = 7


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_town_ranges_cover_every_town_once() -> None:
    for n in (1, 7, 1024, 4097):
        for world in (1, 2, 3, 8):
            spans = [town_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(t1 - t0 for t0, t1 in spans) - min(t1 - t0 for t0, t1 in spans) <= 1


def _worker(rank: int, world: int, port: int, queue) -> None:
    from oracle.oracle import oracle_tick  # the shard's compute stands in for the GPU tick

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    obs, rew = ObsSpec(), RewardSpec(fuse_faint=True)
    full = synth_batch(N_TOWNS, 48, 8, obs)
    shard = shard_batch(full, rank, world)
    out = oracle_tick(shard, obs, rew, capi.TL_PHASE_ALL, 2)
    kpi = torch.from_numpy(out["kpi"][-1])
    rows = gather_kpi_rows(kpi, N_TOWNS)
    totals = reduce_kpi_totals(torch.from_numpy(out["kpi"]))
    if rank == 0:
        queue.put((rows.numpy(), totals.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_ticks_equal_the_unsharded_run() -> None:
    from oracle.oracle import oracle_tick

    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, queue)) for r in range(2)]
    for p in procs:
        p.start()
    rows, totals = queue.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    obs, rew = ObsSpec(), RewardSpec(fuse_faint=True)
    full = synth_batch(N_TOWNS, 48, 8, obs)
    ref = oracle_tick(full, obs, rew, capi.TL_PHASE_ALL, 2)
    assert np.array_equal(rows, ref["kpi"][-1])
    np.testing.assert_allclose(totals, ref["kpi"].reshape(-1, capi.TL_N_KPI).sum(axis=0), rtol=1e-6)


def test_shards_partition_the_state() -> None:
    obs = ObsSpec()
    full = synth_batch(N_TOWNS, 48, 8, obs)
    parts = [shard_batch(full, r, 3) for r in range(3)]
    assert sum(p.n_towns for p in parts) == N_TOWNS and sum(p.n_agents for p in parts) == full.n_agents
    assert np.array_equal(np.concatenate([p.needs for p in parts], axis=1), full.needs)
    assert np.array_equal(np.concatenate([p.ent for p in parts]), full.ent)
    for p in parts:
        p.validate()


def test_sharded_world_phases_equal_the_unsharded_run() -> None:
    """Action words shard with the agents; moves and ledger decay never cross a town, so shards stay independent."""
    from oracle.oracle import oracle_tick
    from townlet_b200.params import DynamicsSpec
    from townlet_b200.synth import random_actions

    obs, rew, dyn = ObsSpec(), RewardSpec(fuse_faint=True), DynamicsSpec()
    full = synth_batch(N_TOWNS, 32, 6, obs)
    words = random_actions(full, 4, seed=11)
    parts = [shard_batch(full, r, 3) for r in range(3)]
    outs = [
        oracle_tick(part, obs, rew, capi.TL_PHASE_WORLD, 4, dyn=dyn, actions=np.ascontiguousarray(shard_actions(words, full, r, 3)))
        for r, part in enumerate(parts)
    ]
    ref = oracle_tick(full, obs, rew, capi.TL_PHASE_WORLD, 4, dyn=dyn, actions=words)
    for key in ("map", "features", "reward", "terminated"):
        assert np.array_equal(np.concatenate([o[key] for o in outs], axis=1), ref[key], equal_nan=True), key
    assert np.array_equal(np.concatenate([p.pos for p in parts]), full.pos)
    assert np.array_equal(np.concatenate([p.riv_score for p in parts]), full.riv_score)
    assert np.array_equal(np.concatenate([p.tie_count for p in parts]), full.tie_count)
