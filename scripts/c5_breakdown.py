"""configs[4] per GPU: where a tick of the rollout goes — graphs of the tick kernel alone, the policy step alone, and both."""
import sys
import torch
sys.path.insert(0, '.')
from townlet_b200 import capi
from townlet_b200.engine import DeviceTownBatch
from townlet_b200.params import ObsSpec, RewardSpec
from townlet_b200.rollout import PolicyMLP, RolloutCapture, policy_step
from townlet_b200.synth import synth_batch, tile_batch

obs, rew = ObsSpec(), RewardSpec(fuse_faint=True)
batch = tile_batch(synth_batch(256, 48, 8, obs), 32)  # 8 192 towns x 8 agents
T = 64
for mode in ("f32", "bf16"):
    dev = DeviceTownBatch(batch.copy(), obs, rew)
    torch.manual_seed(0)
    cap = RolloutCapture(dev, PolicyMLP(obs.n_features, obs.map_shape, action_dim=9), map_dtype=mode)
    out = cap.alloc_outputs(T, summary=False, comps=False, kpi=True)
    rows = cap._new_rows(T)
    cap.prime(out, T, rows)
    mirror, saved = cap._with_mirrors(out)
    fused = cap._fused_policy()
    stream = torch.cuda.Stream()

    def timed(fn, label):
        with torch.cuda.stream(stream):
            fn()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                fn()
            for _ in range(2):
                g.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
        print(f"map {mode}: {label}: {e0.elapsed_time(e1) * 1e3 / 5 / T:.1f} us per tick", flush=True)

    def ticks():
        for t in range(T):
            cap._tick(out, t + 1)

    def policies():
        for t in range(T):
            cap._policy_step(out, t + 1, t, rows, mirror)

    def both():
        for t in range(T):
            cap._tick(out, t + 1)
            cap._policy_step(out, t + 1, t, rows, mirror)

    timed(ticks, "tick kernel alone")
    timed(policies, "policy step alone")
    timed(both, "tick + policy step")
    out.map_bf16, out.features_bf16 = saved
