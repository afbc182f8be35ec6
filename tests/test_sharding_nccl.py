"""GPU, world_size 2 over NCCL (skipped on a single-GPU box): the sharded CUDA ticks and the end-of-rollout KPI
exchange against the unsharded CUDA run and the oracle."""

from __future__ import annotations

import os
import socket

import numpy as np
import pytest

N_TOWNS = 37


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, queue) -> None:
    import torch
    import torch.distributed as dist

    from townlet_b200 import capi
    from townlet_b200.engine import DeviceTownBatch
    from townlet_b200.params import ObsSpec, RewardSpec
    from townlet_b200.sharding import gather_kpi_rows, reduce_kpi_totals, shard_batch
    from townlet_b200.synth import synth_batch

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    obs, rew = ObsSpec(), RewardSpec(fuse_faint=True)
    full = synth_batch(N_TOWNS, 48, 8, obs)
    shard = shard_batch(full, rank, world)
    dev = DeviceTownBatch(shard, obs, rew, f"cuda:{rank}")
    out = dev.alloc_outputs(capi.TL_PHASE_ALL, 3)
    dev.tick(out, capi.TL_PHASE_ALL, 3)
    rows = gather_kpi_rows(out.kpi[-1], N_TOWNS)  # all-gather over NVLink, uneven shards (19 + 18 towns)
    totals = reduce_kpi_totals(out.kpi)
    torch.cuda.synchronize()
    if rank == 0:
        queue.put((rows.cpu().numpy(), totals.cpu().numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_nccl_kpi_exchange_equals_the_unsharded_run() -> None:
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from oracle.oracle import oracle_tick
    from townlet_b200 import capi
    from townlet_b200.engine import DeviceTownBatch
    from townlet_b200.params import ObsSpec, RewardSpec
    from townlet_b200.synth import synth_batch

    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, queue)) for r in range(2)]
    for p in procs:
        p.start()
    rows, totals = queue.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    obs, rew = ObsSpec(), RewardSpec(fuse_faint=True)
    full = synth_batch(N_TOWNS, 48, 8, obs)
    dev = DeviceTownBatch(full.copy(), obs, rew, "cuda:0")
    out = dev.alloc_outputs(capi.TL_PHASE_ALL, 3)
    dev.tick(out, capi.TL_PHASE_ALL, 3)
    torch.cuda.synchronize()
    whole = out.kpi.cpu().numpy()
    assert np.array_equal(rows, whole[-1])  # the same kernels on the same towns: bit-identical rows
    np.testing.assert_allclose(totals, whole.reshape(-1, capi.TL_N_KPI).astype(np.float64).sum(axis=0), rtol=1e-6)
    ref = oracle_tick(full, obs, rew, capi.TL_PHASE_ALL, 3)
    np.testing.assert_allclose(rows, ref["kpi"][-1], rtol=1e-5, atol=1e-6)
