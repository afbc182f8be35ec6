"""GAE scan: the C oracle against the reference's compute_gae golden (CPU), the kernel against the oracle (GPU)."""

from __future__ import annotations

from pathlib import Path

import numpy as np
import pytest

from oracle.oracle import oracle_gae

GOLD = Path(__file__).resolve().parent / "golden" / "gae" / "ref_gae.npz"


@pytest.mark.parametrize("name", ["boot", "noboot", "one"])
def test_oracle_gae_matches_reference_compute_gae(name) -> None:
    d = np.load(GOLD)
    adv, ret = oracle_gae(d[f"{name}_rewards"].T, d[f"{name}_values"].T, d[f"{name}_dones"].T, float(d["gamma"]), float(d["gae_lambda"]))
    assert np.array_equal(adv.T, d[f"{name}_advantages"])  # bit-exact float32
    assert np.array_equal(ret.T, d[f"{name}_returns"])


@pytest.mark.gpu
@pytest.mark.parametrize("steps,agents,bootstrap", [(19, 37, True), (64, 5, False), (1, 3, True), (128, 8192, True)])
def test_gae_kernel_matches_oracle(steps, agents, bootstrap) -> None:
    import torch

    from townlet_b200.rollout import gae

    rng = np.random.default_rng(steps * 1000 + agents)
    rewards = rng.uniform(-0.2, 0.2, size=(steps, agents)).astype(np.float32)
    values = rng.normal(0.0, 0.5, size=(steps + int(bootstrap), agents)).astype(np.float32)
    dones = (rng.random((steps, agents)) < 0.1).astype(np.float32)
    adv, ret = gae(torch.from_numpy(rewards).cuda(), torch.from_numpy(values).cuda(), torch.from_numpy(dones).cuda(), 0.99, 0.95)
    ref_adv, ref_ret = oracle_gae(rewards, values, dones, 0.99, 0.95)
    assert np.array_equal(adv.cpu().numpy(), ref_adv)
    assert np.array_equal(ret.cpu().numpy(), ref_ret)


@pytest.mark.gpu
def test_reference_gae_golden_through_the_kernel() -> None:
    import torch

    from townlet_b200.rollout import gae

    d = np.load(GOLD)
    for name in ("boot", "noboot", "one"):
        t = lambda k: torch.from_numpy(np.ascontiguousarray(d[f"{name}_{k}"].T)).cuda()  # noqa: E731
        adv, ret = gae(t("rewards"), t("values"), t("dones"), float(d["gamma"]), float(d["gae_lambda"]))
        assert np.array_equal(adv.cpu().numpy().T, d[f"{name}_advantages"])
        assert np.array_equal(ret.cpu().numpy().T, d[f"{name}_returns"])


@pytest.mark.gpu
def test_rollout_capture_matches_tick_by_tick_oracle() -> None:
    import torch

    from oracle.oracle import oracle_tick
    from townlet_b200 import capi
    from townlet_b200.engine import DeviceTownBatch
    from townlet_b200.params import ObsSpec, RewardSpec
    from townlet_b200.rollout import PolicyMLP, RolloutCapture
    from townlet_b200.synth import synth_batch

    obs, rew = ObsSpec(), RewardSpec(fuse_faint=True)
    batch = synth_batch(24, 48, 8, obs)
    ref_batch = batch.copy()
    dev = DeviceTownBatch(batch, obs, rew)
    torch.manual_seed(0)
    policy = PolicyMLP(obs.n_features, obs.map_shape, action_dim=9)
    cap = RolloutCapture(dev, policy, autocast_dtype=None)
    ro = cap.capture(5)
    # step t = observation t, the action chosen on it, and the reward / done of the tick that applied it (tick t + 1;
    # policy/trajectory_service.py:127-138): six ticks for five steps, the first one only opens the trajectory
    ref = oracle_tick(ref_batch, obs, rew, capi.TL_PHASE_ALL, 6)
    assert np.array_equal(ro.map.cpu().numpy(), ref["map"][:5])
    assert np.array_equal(ro.features.cpu().numpy().view(np.uint32), ref["features"][:5].view(np.uint32))
    np.testing.assert_allclose(ro.rewards.cpu().numpy(), ref["reward"][1:], rtol=1e-6, atol=1e-9)
    assert np.array_equal(ro.dones.cpu().numpy(), ref["terminated"][1:].astype(np.float32))
    np.testing.assert_allclose(ro.kpi.cpu().numpy(), ref["kpi"][1:], rtol=1e-5, atol=1e-6)
    assert ro.values.shape == (6, batch.n_agents) and ro.actions.shape == (5, batch.n_agents)
    # the bootstrap row is the policy's value of the observation after the last tick, not a repeat of the last step
    last_map = torch.from_numpy(ref["map"][5]).cuda()
    last_feat = torch.from_numpy(ref["features"][5]).cuda()
    with torch.no_grad():
        boot = cap.policy(last_map, last_feat)[1].float()
    assert torch.allclose(ro.values[5], boot, rtol=1e-5, atol=1e-6) and not torch.equal(ro.values[5], ro.values[4])
    ref_adv, ref_ret = oracle_gae(ro.rewards.cpu().numpy(), ro.values.cpu().numpy(), ro.dones.cpu().numpy(), 0.99, 0.95)
    assert np.array_equal(ro.advantages.cpu().numpy(), ref_adv) and np.array_equal(ro.returns.cpu().numpy(), ref_ret)
    # a second capture continues the trajectory: its first observation is the last one of the first capture
    ro2 = cap.capture(3)
    ref2 = oracle_tick(ref_batch, obs, rew, capi.TL_PHASE_ALL, 3)
    assert np.array_equal(ro2.map[0].cpu().numpy(), ref["map"][5]) and np.array_equal(ro2.map[1:].cpu().numpy(), ref2["map"][:2])
    np.testing.assert_allclose(ro2.rewards.cpu().numpy(), ref2["reward"], rtol=1e-6, atol=1e-9)
    assert torch.equal(ro2.values[0], ro.values[5])


@pytest.mark.gpu
@pytest.mark.parametrize("act_on_world", [False, True])
def test_graphed_rollout_equals_eager_rollout(act_on_world) -> None:
    import torch

    from townlet_b200 import capi
    from townlet_b200.engine import DeviceTownBatch
    from townlet_b200.params import DynamicsSpec, ObsSpec, RewardSpec
    from townlet_b200.rollout import PolicyMLP, RolloutCapture
    from townlet_b200.synth import synth_batch

    obs, rew = ObsSpec(variant="hybrid" if act_on_world else "compact"), RewardSpec(fuse_faint=True)
    dyn = DynamicsSpec() if act_on_world else None
    batch = synth_batch(16, 48, 10, obs)
    policy = PolicyMLP(obs.n_features, obs.map_shape, action_dim=9)
    eager_dev = DeviceTownBatch(batch.copy(), obs, rew, dyn=dyn)
    graph_dev = DeviceTownBatch(batch.copy(), obs, rew, dyn=dyn)
    eager = RolloutCapture(eager_dev, policy, autocast_dtype=None, act_on_world=act_on_world)
    graph_cap = RolloutCapture(graph_dev, policy, autocast_dtype=None, act_on_world=act_on_world)
    eager.seed = graph_cap.seed = 5
    graphed = graph_cap.build_graph(4)  # its warm-up must leave towns, action words and draw counter untouched
    assert graphed.library_launches >= 4
    for _ in range(2):  # two consecutive rollouts: the graph continues from the device state
        a = eager.capture(4)
        b = graphed.replay()
        torch.cuda.synchronize()
        assert torch.equal(a.map, b.map) and torch.equal(a.features, b.features)
        assert torch.equal(a.rewards, b.rewards) and torch.equal(a.dones, b.dones)
        assert torch.equal(a.actions, b.actions) and torch.allclose(a.values, b.values)
    assert torch.equal(eager_dev.arena, graph_dev.arena)
    assert capi.load_library().tl_launch_count() > 0


def test_rollout_replay_batch_matches_reference_frames_to_replay_sample() -> None:
    """Time-major rollout tensors -> the reference's ReplayBatch arrays (policy/replay.py:550-636, 245-276).

    CPU tensors suffice: the export is a transpose of the rollout buffers, whatever device they live on.
    """
    import torch

    from townlet_b200.rollout import Rollout

    d = np.load(Path(__file__).resolve().parent / "golden" / "replay" / "ref_replay.npz")
    t = {k[3:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("in_")}
    steps, agents = t["rewards"].shape
    rollout = Rollout(
        map=t["map"], features=t["features"], rewards=t["rewards"], dones=t["dones"], values=t["values"],
        actions=t["actions"], log_probs=t["log_probs"], advantages=torch.zeros(steps, agents),
        returns=torch.zeros(steps, agents), kpi=torch.zeros(steps, 1, 8),
    )
    got = rollout.replay_batch(bootstrap="repeat")
    for key in ("maps", "features", "actions", "old_log_probs", "value_preds", "rewards", "dones"):
        assert got[key].dtype == d[key].dtype, key
        assert np.array_equal(got[key], d[key]), key
    # the device rollout also carries the real bootstrap value of the final observation
    boot = rollout.replay_batch(agents=slice(1, 4))
    assert boot["value_preds"].shape == (3, steps + 1)
    assert np.array_equal(boot["value_preds"][:, -1], d["in_values"][-1, 1:4])
    assert np.array_equal(boot["maps"], d["maps"][1:4])


@pytest.mark.gpu
def test_rollout_capture_with_the_policy_moving_the_agents() -> None:
    """act_on_world: the sampled actions become action words, the next tick applies them (SURVEY §8f ranks 2 + 3).

    The oracle replays the same rollout from the recorded action indices: word t is built from action t - 1 and the
    positions before tick t, exactly as the capture does on the device.
    """
    import torch

    from oracle.oracle import oracle_tick
    from townlet_b200 import capi
    from townlet_b200.engine import DeviceTownBatch
    from townlet_b200.params import DynamicsSpec, ObsSpec, RewardSpec, pack_actions
    from townlet_b200.rollout import PolicyMLP, RolloutCapture
    from townlet_b200.synth import synth_batch

    obs, rew, dyn = ObsSpec(), RewardSpec(fuse_faint=True), DynamicsSpec()
    batch = synth_batch(12, 24, 6, obs)
    ref_batch = batch.copy()
    dev = DeviceTownBatch(batch, obs, rew, dyn=dyn)
    torch.manual_seed(3)
    policy = PolicyMLP(obs.n_features, obs.map_shape, action_dim=9)
    steps = 6
    ro = RolloutCapture(dev, policy, autocast_dtype=None, act_on_world=True).capture(steps)
    torch.cuda.synchronize()
    actions = ro.actions.cpu().numpy()
    kinds = np.array(RolloutCapture.ACTION_KIND + (4,) * 11)[:16]
    dxs = np.array(RolloutCapture.ACTION_DX + (0,) * 11)[:16]
    dys = np.array(RolloutCapture.ACTION_DY + (0,) * 11)[:16]
    moved = False
    for t in range(steps + 1):  # tick 0 opens the trajectory; tick t applies action t - 1 and yields reward t - 1
        if t == 0:
            words = np.zeros(batch.n_agents, np.uint32)
        else:
            idx = np.minimum(actions[t - 1], 15)
            x = ref_batch.pos.view(np.uint32) & 0xFFFF
            y = (ref_batch.pos.view(np.uint32) >> 16) & 0xFFFF
            nx = np.clip(x.astype(np.int64) + dxs[idx], 0, ref_batch.grid_w - 1)
            ny = np.clip(y.astype(np.int64) + dys[idx], 0, ref_batch.grid_h - 1)
            words = pack_actions(kinds[idx], 0, kinds[idx] == 1, nx, ny, 1)
            moved = moved or bool(np.any((kinds[idx] == 1) & ((nx != x) | (ny != y))))
        ref = oracle_tick(ref_batch, obs, rew, capi.TL_PHASE_WORLD, 1, dyn=dyn, actions=words)
        if t < steps:
            assert np.array_equal(ro.map[t].cpu().numpy(), ref["map"][0]), t
            assert np.array_equal(ro.features[t].cpu().numpy().view(np.uint32), ref["features"][0].view(np.uint32)), t
        if t > 0:
            np.testing.assert_allclose(ro.rewards[t - 1].cpu().numpy(), ref["reward"][0], rtol=1e-6, atol=1e-9)
    assert moved, "the sampled actions never moved an agent"
    state = dev.download_state()
    assert np.array_equal(state.pos, ref_batch.pos) and np.array_equal(state.ent, ref_batch.ent)


@pytest.mark.gpu
def test_cast_pad_bf16_and_the_padded_policy_forward() -> None:
    """tl_cast_pad_bf16 equals torch's float32 -> bfloat16 cast plus zero padding, and the policy's padded bfloat16
    inference path agrees with plain autocast to bfloat16 rounding."""
    import torch

    from townlet_b200 import capi
    from townlet_b200.rollout import PolicyMLP

    lib = capi.load_library()
    torch.manual_seed(1)
    for rows, width, padded in ((1000, 484, 512), (77, 81, 128), (5, 7, 8), (3, 64, 64)):
        src = torch.randn(rows, width, device="cuda") * 3
        dst = torch.full((rows, padded), 7.0, dtype=torch.bfloat16, device="cuda")
        capi.check(lib.tl_cast_pad_bf16(src.data_ptr(), dst.data_ptr(), rows, width, padded, torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        assert torch.equal(dst[:, :width], src.to(torch.bfloat16)) and not dst[:, width:].any()
    assert lib.tl_cast_pad_bf16(None, None, 1, 4, 4, None) == -1 and lib.tl_cast_pad_bf16(1, 1, 1, 5, 4, None) == -1
    # destination rows inside a wider buffer (the policy's combined input rows)
    for rows, width, padded, stride, col in ((300, 484, 512, 640, 0), (300, 81, 128, 640, 512), (9, 7, 8, 26, 2)):
        src = torch.randn(rows, width, device="cuda")
        buf = torch.full((rows, stride), 7.0, dtype=torch.bfloat16, device="cuda")
        capi.check(lib.tl_cast_pad_bf16_strided(src.data_ptr(), buf[:, col:].data_ptr(), rows, width, padded, stride,
                                                torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        assert torch.equal(buf[:, col : col + width], src.to(torch.bfloat16)) and not buf[:, col + width : col + padded].any()
        assert (buf[:, :col] == 7).all() and (buf[:, col + padded :] == 7).all()
    assert lib.tl_cast_pad_bf16_strided(1, 1, 1, 4, 8, 6, None) == -1

    model = PolicyMLP(81, (4, 11, 11), 9).cuda().eval()
    maps = (torch.rand(4096, 4, 11, 11, device="cuda") < 0.05).float()
    feats = torch.randn(4096, 81, device="cuda")
    with torch.no_grad():
        exact_logits, exact_value = model(maps, feats)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits, value = model(maps, feats)  # padded path
            hidden = torch.cat([model.map_encoder(maps), model.feature_encoder(feats)], dim=-1)
            plain_logits = model.policy_head(model.shared(hidden))
    assert hasattr(model, "_pad_in")
    err_padded = (logits.float() - exact_logits).abs().max().item()
    err_plain = (plain_logits.float() - exact_logits).abs().max().item()
    assert err_padded < 2 * err_plain + 1e-3, (err_padded, err_plain)
    assert (value.float() - exact_value).abs().max().item() < 0.02


def _mirror_case(variant: str, world: bool, kernel_form, env: dict[str, str], three_dim: bool, n_ticks: int = 3,
                 combined: bool = False) -> None:
    import torch

    from townlet_b200 import capi
    from townlet_b200.engine import DeviceTownBatch
    from townlet_b200.params import DynamicsSpec, ObsSpec, RewardSpec, pack_actions
    from townlet_b200.synth import synth_batch

    for key, value in env.items():
        kernel_form.setenv(key, value)
    obs, rew = ObsSpec(variant=variant), RewardSpec(fuse_faint=True)
    # stacked agents, objects and reserved tiles inside the windows: every patch kind is written
    batch = synth_batch(37, 16, 9, obs)
    dev = DeviceTownBatch(batch, obs, rew, dyn=DynamicsSpec() if world else None)
    out = dev.alloc_outputs(capi.TL_PHASE_ALL, n_ticks)
    flat = int(np.prod(obs.map_shape))
    row = (flat + 63) // 64 * 64
    shape = (n_ticks, batch.n_agents, row) if three_dim else (batch.n_agents, row)
    frow = (obs.n_features + 63) // 64 * 64
    if combined:  # both mirrors as column slices of one buffer of combined rows (what PolicyMLP reads)
        both = torch.full((*shape[:-1], row + frow + 8), 7.0, dtype=torch.bfloat16, device="cuda")
        out.map_bf16, out.features_bf16 = both[..., :row], both[..., row : row + frow]
    else:
        out.map_bf16 = torch.full(shape, 7.0, dtype=torch.bfloat16, device="cuda")
        out.features_bf16 = torch.full((*shape[:-1], frow), 7.0, dtype=torch.bfloat16, device="cuda")
    if world:
        rng = np.random.default_rng(5)
        n = batch.n_agents
        words = pack_actions(np.full(n, 1), 1, np.ones(n, bool), rng.integers(0, 16, n), rng.integers(0, 16, n), 1)
        actions = torch.from_numpy(words.view(np.int32)).cuda()
        dev.tick(out, capi.TL_PHASE_WORLD, n_ticks, actions=actions)
    else:
        dev.tick(out, capi.TL_PHASE_ALL, n_ticks)
    torch.cuda.synchronize()
    want = out.map.reshape(n_ticks, batch.n_agents, flat).to(torch.bfloat16)
    got = out.map_bf16 if three_dim else out.map_bf16[None]
    want = want if three_dim else want[-1:]
    assert want.float().sum().item() > 2 * batch.n_agents  # more than the self cells
    assert torch.equal(got[..., :flat].view(torch.int16), want.view(torch.int16))
    assert not got[..., flat:].any()
    want_f = out.features.to(torch.bfloat16) if three_dim else out.features[-1:].to(torch.bfloat16)
    got_f = out.features_bf16 if three_dim else out.features_bf16[None]
    assert torch.equal(got_f[..., : obs.n_features].view(torch.int16), want_f.view(torch.int16))
    assert not got_f[..., obs.n_features :].any()
    if combined:
        assert (both[..., row + frow :] == 7).all()  # nothing is written past the two widths


@pytest.mark.gpu
@pytest.mark.parametrize(
    "variant,world,env,three_dim",
    [
        ("hybrid", False, {}, True),  # written by the tiles role
        ("hybrid", False, {}, False),  # one buffer overwritten every tick
        ("hybrid", True, {}, True),  # world-dynamics instantiation, agents moved by action words
        ("hybrid", False, {"TL_NO_MIRROR_FUSE": "1"}, True),  # cast pass after the launch
        ("hybrid", False, {"TL_MODE": "grid"}, True),
        ("hybrid", False, {"TL_GENERIC": "1"}, False),
        ("full", False, {}, True),  # path channels: round to nearest even
        ("compact", False, {}, True),
    ],
)
def test_bf16_mirror_of_the_map_equals_the_cast_of_the_float32_map(variant, world, env, three_dim, kernel_form) -> None:
    """TlObsOut.map_bf16 (ABI v3): every kernel form leaves the padded bfloat16 rows torch's cast of the float32 map gives."""
    _mirror_case(variant, world, kernel_form, env, three_dim)
    _mirror_case(variant, world, kernel_form, env, three_dim, combined=True)


@pytest.mark.gpu
def test_bf16_mirror_arguments_are_validated() -> None:
    import torch

    from townlet_b200 import capi
    from townlet_b200.engine import DeviceTownBatch
    from townlet_b200.params import ObsSpec, RewardSpec
    from townlet_b200.synth import synth_batch

    obs = ObsSpec()
    batch = synth_batch(3, 16, 4, obs)
    dev = DeviceTownBatch(batch, obs, RewardSpec())
    out = dev.alloc_outputs(capi.TL_PHASE_ALL, 1)
    for shape in ((batch.n_agents, 480), (batch.n_agents, 490)):  # narrower than the tile / not a multiple of 8
        out.map_bf16 = torch.zeros(shape, dtype=torch.bfloat16, device="cuda")
        with pytest.raises(capi.TownletB200Error, match="map_bf16"):
            dev.tick(out, capi.TL_PHASE_ALL, 1)
    out.map_bf16 = torch.zeros((batch.n_agents, 512), dtype=torch.float16, device="cuda")
    with pytest.raises(capi.TownletB200Error, match="map_bf16"):
        dev.tick(out, capi.TL_PHASE_ALL, 1)
    out.map_bf16 = None
    out.features_bf16 = torch.zeros((batch.n_agents, 80), dtype=torch.bfloat16, device="cuda")  # narrower than F = 81
    with pytest.raises(capi.TownletB200Error, match="features_bf16"):
        dev.tick(out, capi.TL_PHASE_ALL, 1)
    # either mirror alone is served
    out.features_bf16 = torch.full((batch.n_agents, 88), 7.0, dtype=torch.bfloat16, device="cuda")
    dev.tick(out, capi.TL_PHASE_OBS, 1)
    torch.cuda.synchronize()
    assert torch.equal(out.features_bf16[:, :81], out.features[0].to(torch.bfloat16)) and not out.features_bf16[:, 81:].any()


@pytest.mark.gpu
@pytest.mark.parametrize("act_on_world", [False, True])
def test_rollout_with_the_kernel_written_bf16_map_equals_the_cast_pass(act_on_world) -> None:
    """The bfloat16 policy step reads the rows the tick kernel wrote instead of casting the float32 map itself: same
    rollout bit for bit (same seeds), eager and graphed."""
    import torch

    from townlet_b200.engine import DeviceTownBatch
    from townlet_b200.params import DynamicsSpec, ObsSpec, RewardSpec
    from townlet_b200.rollout import PolicyMLP, RolloutCapture
    from townlet_b200.synth import synth_batch

    obs, rew = ObsSpec(), RewardSpec(fuse_faint=True)
    torch.manual_seed(11)
    policy = PolicyMLP(obs.n_features, obs.map_shape, action_dim=9)
    rollouts = []
    for fuse in (False, True):
        dev = DeviceTownBatch(synth_batch(20, 24, 8, obs), obs, rew, dyn=DynamicsSpec() if act_on_world else None)
        cap = RolloutCapture(dev, policy, act_on_world=act_on_world)
        cap.fuse_policy = False  # both arms on the library products: what differs is who writes the bfloat16 rows
        cap.fuse_map_cast = fuse
        cap.seed = 23  # Philox key of the sampling kernel (otherwise torch.initial_seed() at construction)
        torch.manual_seed(23)
        ro = cap.capture(6)
        torch.cuda.synchronize()
        assert (cap._map_bf16 is not None) == fuse
        rollouts.append(ro)
    a, b = rollouts
    for name in ("map", "features", "rewards", "values", "actions", "log_probs", "advantages", "returns"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    assert len(torch.unique(b.actions)) > 1


def test_philox_restatement_matches_the_random123_known_answers() -> None:
    """The three Philox4x32-10 known-answer vectors of the Random123 library (kat_vectors) pin the oracle's generator."""
    from oracle.oracle import philox4x32_10

    m = 0xFFFFFFFF
    cases = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((m, m, m, m), (m, m), (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0), (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for counter, key, want in cases:
        assert tuple(int(w) for w in philox4x32_10(*counter, *key)) == want


def test_oracle_sampler_draws_the_softmax_distribution() -> None:
    from oracle.oracle import oracle_sample_actions

    logits = np.tile(np.array([0.5, -1.0, 2.0, 0.0, 1.0, -3.0, 0.25, 0.0, 1.5, 9.0], np.float32), (200_000, 1))
    actions, logp_act, logp, u = oracle_sample_actions(logits, 9, seed=99, counter=3, offset=2)
    p = np.exp(logp[0].astype(np.float64))
    assert abs(p.sum() - 1.0) < 1e-6 and 0.0 < u.min() and u.max() < 1.0
    freq = np.bincount(actions, minlength=9) / len(actions)
    assert np.all(np.abs(freq - p) < 5 * np.sqrt(p * (1 - p) / len(actions)) + 1e-4), (freq, p)
    assert np.array_equal(logp_act, logp[np.arange(len(actions)), actions])


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,row,n_actions,with_value", [("float32", 9, 9, False), ("bfloat16", 16, 9, True), ("float32", 20, 16, True), ("bfloat16", 8, 1, False)])
def test_sample_actions_kernel_matches_the_oracle(dtype, row, n_actions, with_value) -> None:
    """tl_sample_actions against the numpy restatement: log-probabilities to float32 rounding, the same uniforms, hence
    the same actions (a last-bit difference of logf between libm and CUDA may flip a near-tie: at most 1 row in 10 000),
    and torch's own log_softmax as a second opinion."""
    import torch

    from oracle.oracle import oracle_sample_actions
    from townlet_b200.rollout import sample_actions

    torch.manual_seed(4)
    n = 50_000
    heads = (torch.randn(n, row, device="cuda") * 2).to(getattr(torch, dtype))
    counter = torch.tensor([5], dtype=torch.int64, device="cuda")
    got = sample_actions(heads, n_actions, 1234567890123, counter, 7, with_value=with_value)
    torch.cuda.synchronize()
    ref_act, ref_logp_act, ref_logp, _ = oracle_sample_actions(heads.float().cpu().numpy(), n_actions, 1234567890123, 5, 7)
    act, logp_act = got[0].cpu().numpy(), got[1].cpu().numpy()
    assert (act != ref_act).mean() <= 1e-4
    np.testing.assert_allclose(logp_act, ref_logp[np.arange(n), act], rtol=0, atol=2e-6)
    torch_logp = torch.log_softmax(heads[:, :n_actions].float(), dim=-1).gather(1, got[0].unsqueeze(1)).squeeze(1)
    assert torch.allclose(got[1], torch_logp, rtol=0, atol=2e-6)
    if with_value:
        assert torch.equal(got[2], heads[:, n_actions].float())
    # a pure function of (seed, counter + offset, row): the same draw again, another one a step later
    again = sample_actions(heads, n_actions, 1234567890123, None, 12)
    later = sample_actions(heads, n_actions, 1234567890123, counter, 8)
    assert torch.equal(again[0], got[0])
    assert n_actions == 1 or (later[0] != got[0]).float().mean().item() > 0.2


@pytest.mark.gpu
def test_sample_actions_rejects_bad_arguments_and_graph_replays_draw_fresh_numbers() -> None:
    import torch

    from townlet_b200 import capi
    from townlet_b200.engine import DeviceTownBatch
    from townlet_b200.params import ObsSpec, RewardSpec
    from townlet_b200.rollout import PolicyMLP, RolloutCapture, sample_actions
    from townlet_b200.synth import synth_batch

    heads = torch.zeros(4, 17, device="cuda")
    with pytest.raises(capi.TownletB200Error):
        sample_actions(heads, 17, 0)  # more than TL_MAX_ACTIONS
    with pytest.raises(capi.TownletB200Error):
        sample_actions(heads[:, :9].contiguous(), 9, 0, with_value=True)  # no value column
    with pytest.raises(capi.TownletB200Error):
        sample_actions(heads.cpu(), 9, 0)

    obs, rew = ObsSpec(), RewardSpec(fuse_faint=True)
    torch.manual_seed(2)
    policy = PolicyMLP(obs.n_features, obs.map_shape, action_dim=9)
    cap = RolloutCapture(DeviceTownBatch(synth_batch(64, 24, 8, obs), obs, rew), policy)
    graphed = cap.build_graph(3)
    first = graphed.replay().actions.clone()
    second = graphed.replay().actions.clone()
    torch.cuda.synchronize()
    assert (first != second).float().mean().item() > 0.3  # the device counter moved on between the replays
    assert first.min().item() >= 0 and first.max().item() < 9
    freq = torch.bincount(torch.cat([first, second]).flatten(), minlength=9).float()
    assert (freq > 0).all()
