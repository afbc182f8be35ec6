"""GPU: the map delivered as bfloat16 only (``TlObsOut.map == NULL`` with ``map_bf16`` rows) — exact for the hybrid map,
whose cells are 0 or 1 — and the rollout that keeps those rows as its map observation (``RolloutCapture(map_dtype="bf16")``).
"""

from __future__ import annotations

import numpy as np
import pytest

from townlet_b200 import capi
from townlet_b200.params import DynamicsSpec, ObsSpec, RewardSpec
from townlet_b200.synth import random_actions, synth_batch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [False, True], ids=["tick", "world"])
def test_bf16_only_map_equals_the_oracle_map(world) -> None:
    import torch

    from oracle.oracle import oracle_tick
    from townlet_b200.engine import DeviceTownBatch

    obs, rew = ObsSpec(), RewardSpec(fuse_faint=True)
    dyn = DynamicsSpec() if world else None
    phases = capi.TL_PHASE_WORLD if world else capi.TL_PHASE_ALL
    batch = synth_batch(40, 48, 8, obs)
    ref_batch = batch.copy()
    n_ticks = 3
    words = random_actions(batch, n_ticks, seed=11) if world else None
    dev = DeviceTownBatch(batch, obs, rew, dyn=dyn)
    out = dev.alloc_outputs(phases, n_ticks, summary=True, map_f32=False)
    assert out.map is None
    rows = torch.full((n_ticks, batch.n_agents, 640), 7.0, dtype=torch.bfloat16, device="cuda")  # stale data must be overwritten
    out.map_bf16, out.features_bf16 = rows[:, :, :512], rows[:, :, 512:]
    actions = torch.from_numpy(words.view(np.int32)).cuda() if world else None
    dev.tick(out, phases, n_ticks, actions=actions)
    torch.cuda.synchronize()
    ref = oracle_tick(ref_batch, obs, rew, phases, n_ticks, dyn=dyn, actions=words)
    got = rows.float().cpu().numpy()
    flat = ref["map"].reshape(n_ticks, batch.n_agents, -1)
    assert np.array_equal(got[:, :, : flat.shape[-1]], flat)  # every cell, exactly
    assert not got[:, :, flat.shape[-1] : 512].any()  # zero padding up to the row's map width
    assert np.array_equal(out.features.cpu().numpy().view(np.uint32), ref["features"].view(np.uint32))
    feat_bf16 = torch.from_numpy(ref["features"]).to(torch.bfloat16).float().numpy()
    assert np.array_equal(got[:, :, 512 : 512 + feat_bf16.shape[-1]], feat_bf16)
    np.testing.assert_allclose(out.reward.cpu().numpy(), ref["reward"], rtol=1e-6, atol=1e-9)
    assert np.array_equal(out.summary.cpu().numpy(), ref["summary"])


def test_bf16_only_map_is_refused_where_the_kernel_does_not_write_the_rows(kernel_form) -> None:
    import torch

    from townlet_b200.engine import DeviceTownBatch

    rew = RewardSpec(fuse_faint=True)
    for variant, setup in (("compact", None), ("hybrid", ("TL_MODE", "grid")), ("hybrid", ("TL_GENERIC", "1"))):
        if setup:
            kernel_form.setenv(*setup)
        obs = ObsSpec(variant=variant)
        batch = synth_batch(4, 48, 8, obs)
        dev = DeviceTownBatch(batch, obs, rew)
        out = dev.alloc_outputs(capi.TL_PHASE_ALL, 1, map_f32=False)
        width = (int(np.prod(obs.map_shape)) + 63) // 64 * 64
        out.map_bf16 = torch.zeros((batch.n_agents, width), dtype=torch.bfloat16, device="cuda")
        with pytest.raises(capi.TownletB200Error, match="map may be NULL only"):
            dev.tick(out, capi.TL_PHASE_ALL, 1)
    # and without any map output at all
    obs = ObsSpec()
    dev = DeviceTownBatch(synth_batch(4, 48, 8, obs), obs, rew)
    out = dev.alloc_outputs(capi.TL_PHASE_ALL, 1, map_f32=False)
    with pytest.raises(capi.TownletB200Error, match="observation outputs are NULL"):
        dev.tick(out, capi.TL_PHASE_ALL, 1)


@pytest.mark.parametrize("act_on_world", [False, True], ids=["observe", "act"])
@pytest.mark.parametrize("graphed", [False, True], ids=["eager", "graph"])
def test_rollout_with_bf16_map_rows_equals_the_float32_rollout(act_on_world, graphed) -> None:
    import torch

    from townlet_b200.engine import DeviceTownBatch
    from townlet_b200.rollout import PolicyMLP, RolloutCapture

    obs, rew = ObsSpec(), RewardSpec(fuse_faint=True)
    dyn = DynamicsSpec() if act_on_world else None
    batch = synth_batch(20, 48, 8, obs)
    torch.manual_seed(2)
    policy = PolicyMLP(obs.n_features, obs.map_shape, action_dim=9)
    devs = [DeviceTownBatch(batch.copy(), obs, rew, dyn=dyn) for _ in range(2)]
    caps = [RolloutCapture(devs[0], policy, act_on_world=act_on_world, map_dtype="f32"),
            RolloutCapture(devs[1], policy, act_on_world=act_on_world, map_dtype="bf16")]
    for cap in caps:
        cap.seed = 9
    steps = 5
    if graphed:
        runs = [cap.build_graph(steps) for cap in caps]
        step = lambda k: runs[k].replay()
    else:
        step = lambda k: caps[k].capture(steps)
    for _ in range(2):  # two consecutive captures: the carried observation rows open the second one
        a, b = step(0), step(1)
        torch.cuda.synchronize()
        assert b.map.dtype == torch.bfloat16 and b.map.shape == (steps, batch.n_agents, int(np.prod(obs.map_shape)))
        assert torch.equal(a.map.reshape(steps, batch.n_agents, -1), b.map.float())
        assert torch.equal(a.features, b.features)
        assert torch.equal(a.actions, b.actions) and torch.equal(a.log_probs, b.log_probs) and torch.equal(a.values, b.values)
        assert torch.equal(a.rewards, b.rewards) and torch.equal(a.dones, b.dones) and torch.equal(a.kpi, b.kpi)
        assert torch.equal(a.advantages, b.advantages)
    assert torch.equal(devs[0].arena, devs[1].arena)
    ra, rb = a.replay_batch(), b.replay_batch()
    assert np.array_equal(ra["maps"].reshape(rb["maps"].shape), rb["maps"]) and np.array_equal(ra["actions"], rb["actions"])
