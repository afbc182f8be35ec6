#!/usr/bin/env python
"""bench.py — agent-steps/s of the per-tick hot path (obs build + need decay + reward).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3-compact|c3-full|c4|c5|c2-world|c5-world]
    python bench.py --impl reference ...      # the CPU arm (oracle port, all host threads)

A "step" is ONE TICK of the whole synthetic batch: need decay -> faint predicate -> reward -> observation build
for every agent.  One launch advances a rollout segment of ticks (128 for configs[1]; the kernel re-reads state and
recomputes every output every tick; a multi-tick launch only amortises launch overhead).  The launch is PLANNED once
(``tl_plan_create``) and then repeated back to back from one C call — ``reps`` launches, at least ``--min-launches``
(200), at least ``--min-seconds`` (0.1 s) of device time and at least the ``--steps K`` ticks asked for — so the timed
region holds nothing but kernels: no Python, no ctypes struct building, no attribute queries between the two CUDA events.  Every tick writes its observation tensors to
its own slot of an output ring of at least 1 GiB (8x the 126 MiB L2), so every output byte drains to HBM.

Rank 0 prints ONE JSON line (README / DESIGN.md §5): the headline workload plus, under ``workloads``, the other
single-GPU configurations of BASELINE.json measured the same way in the same process.
"""

from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

# BASELINE.json configs; bytes per agent-step from SURVEY.md §8(d)
WORKLOADS = {
    "c2": dict(variant="hybrid", towns=1024, grid=48, agents=8, bytes_per_agent_step=2712,
               name="configs[1]: hybrid, 1,024 towns x 48x48 x 8 agents"),
    "c3-compact": dict(variant="compact", towns=16384, grid=48, agents=10, bytes_per_agent_step=1756,
                       name="configs[2]: compact, 16,384 towns x 48x48 x 10 agents"),
    "c3-full": dict(variant="full", towns=16384, grid=48, agents=10, bytes_per_agent_step=3680,
                    name="configs[2]: full, 16,384 towns x 48x48 x 10 agents"),
    "c4": dict(variant="hybrid", towns=4096, grid=128, agents=64, bytes_per_agent_step=2712,
               name="configs[3]: hybrid, 4,096 towns x 128x128 x 64 agents"),
    # configs[1] with the world-dynamics phases: random action words every tick, ledger decay;
    # +24 B per agent-step (action word read, pos / entity word / flags / last-action duration and hash write-back)
    "c2-world": dict(variant="hybrid", towns=1024, grid=48, agents=8, bytes_per_agent_step=2712 + 24, world=True,
                     name="configs[1] + actions and ledger decay in the tick: hybrid, 1,024 towns x 48x48 x 8 agents"),
    # configs[4] per GPU: 65,536 towns over 8 GPUs = 8,192 towns per rank, policy forward in the loop
    "c5": dict(variant="hybrid", towns=8192, grid=48, agents=8, bytes_per_agent_step=2712, policy=True,
               name="configs[4]: PPO rollout capture, 8,192 towns per GPU x 48x48 x 8 agents, policy forward in-loop"),
    # configs[4] with the map kept ONCE, as the bfloat16 rows the policy reads (exact: its cells are 0 or 1) instead of a float32
    # tensor plus those rows: 2 712 - 1 936 (float32 map) + 1 280 (the padded bfloat16 row of map and features) bytes per agent-step
    "c5-bf16": dict(variant="hybrid", towns=8192, grid=48, agents=8, bytes_per_agent_step=2712 - 1936 + 1280, policy=True, map_bf16=True,
                    name="configs[4], map observation stored as bfloat16 rows only: 8,192 towns per GPU x 48x48 x 8 agents, policy forward in-loop"),
    # configs[4] with the sampled actions applied by the next tick (moves) and ledger decay: the policy drives the world
    "c5-world": dict(variant="hybrid", towns=8192, grid=48, agents=8, bytes_per_agent_step=2712 + 24, policy=True, world=True,
                     name="configs[4] + the policy's actions applied on the device: 8,192 towns per GPU x 48x48 x 8 agents"),
}
EXTRA_WORKLOADS = ("c3-compact", "c3-full", "c4", "c5", "c5-bf16")  # measured beside the headline workload in the default run
L2_BYTES = 126 * 1024 * 1024
RING_BYTES = 8 * L2_BYTES  # output ring: >= 8 x L2 (>= 1 GiB)
DISTINCT_TOWNS = 256  # towns generated on the host; larger batches tile them
METRIC = "agent-steps/sec (obs build+needs+reward)"


def load_peaks() -> tuple[float, str]:
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        try:
            return float(json.loads(path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def make_batch(wl: dict, rank: int, towns: int | None = None):
    from townlet_b200.params import ObsSpec
    from townlet_b200.synth import synth_batch, tile_batch

    towns = towns or wl["towns"]
    obs = ObsSpec(variant=wl["variant"])
    distinct = min(towns, DISTINCT_TOWNS)
    base = synth_batch(distinct, wl["grid"], wl["agents"], obs, first_town=rank * wl["towns"])
    reps = towns // distinct
    return (tile_batch(base, reps) if reps > 1 else base), obs


def bench_config(wl: dict) -> dict:
    """The ``config`` object of the JSON line — the same keys and values in the GPU arm and the reference arm."""
    return {
        "workload": wl["name"],
        "variant": wl["variant"],
        "towns_per_gpu": wl["towns"],
        "agents_per_town": wl["agents"],
        "grid": wl["grid"],
        "l2": "outputs larger than L2: every tick writes its own slot of an output ring of >= 1 GiB (8 x the 126 MiB L2); "
        "the SoA state is re-read from global memory every tick",
        "partition": "contiguous block of towns per rank, no collective in the tick; KPI all-reduce after the rollout",
    }


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons of one GPU every ~10 ms from start() to stop(); the launching thread is
    inside one C call (GIL released) while the timed launches are issued, so the sampler never competes with it."""

    def __init__(self, index: int) -> None:
        super().__init__(daemon=True)
        self.index = index
        self.samples: list[tuple[float, int, int]] = []  # (perf_counter, sm MHz, throttle mask)
        self.max_mhz = 0
        self.error: str | None = None
        self._halt = threading.Event()
        self.ready = threading.Event()

    def run(self) -> None:
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = int(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.names = {
                nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            }
            self.ready.set()
            while not self._halt.is_set():
                mhz = int(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.samples.append((time.perf_counter(), mhz, mask))
                self._halt.wait(0.01)
        except Exception as exc:  # pragma: no cover - NVML missing
            self.error = f"nvml_unavailable:{type(exc).__name__}"
            self.ready.set()

    def stop(self) -> None:
        self._halt.set()
        self.join(timeout=2)

    def window(self, t0: float, t1: float) -> dict:
        """Clocks line of the samples taken between two perf_counter stamps (the timed region)."""
        inside = [(m, k) for (t, m, k) in self.samples if t0 <= t <= t1] or [(m, k) for (_, m, k) in self.samples[-1:]]
        mhz = sorted(m for m, _ in inside)
        reasons = sorted({name for _, k in inside for bit, name in getattr(self, "names", {}).items() if k & bit})
        if self.error:
            reasons.append(self.error)
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None, "sm_max_mhz": self.max_mhz or None, "reasons": reasons,
                "samples_in_timed_region": len(inside)}


def cpu_oracle_rate(wl: dict, budget_s: float, threads: int) -> tuple[float, str]:
    """agent-steps/s of the C oracle (CPU restatement of the reference) on the workload's full batch, bounded in time."""
    from oracle.oracle import oracle_tick
    from townlet_b200 import capi
    from townlet_b200.params import DynamicsSpec, RewardSpec

    batch, obs = make_batch(wl, 0)
    rew = RewardSpec(fuse_faint=True)
    phases, extra = capi.TL_PHASE_ALL, {}
    per_call = 2
    if wl.get("world"):
        from townlet_b200.synth import random_actions

        phases = capi.TL_PHASE_WORLD
        extra = dict(dyn=DynamicsSpec(), actions=random_actions(batch, per_call, seed=4321))
    oracle_tick(batch, obs, rew, phases, per_call, n_threads=threads, want_comps=False, **extra)  # warm-up
    ticks_done = 0
    t0 = time.perf_counter()
    while True:
        oracle_tick(batch, obs, rew, phases, per_call, n_threads=threads, want_comps=False, **extra)
        ticks_done += per_call
        if time.perf_counter() - t0 >= budget_s:
            break
    dt = time.perf_counter() - t0
    rate = batch.n_agents * ticks_done / dt
    return rate, (f"{batch.n_towns} towns x {wl['agents']} agents x {ticks_done} ticks in {dt:.1f} s, "
                  f"C oracle (oracle/townlet_oracle.c, gcc -O3 -mavx2, OpenMP over towns, {threads} threads)")


def python_reference_baseline(wl: dict, timeout_s: float = 240.0) -> dict:
    """The reference's OWN Python on this box's host cores (oracle/reference_timing.py in a subprocess)."""
    script = ROOT / "oracle" / "reference_timing.py"
    if not (ROOT / "oracle" / "_ref" / "src" / "townlet").is_dir():
        return {"unavailable": "oracle/_ref is missing (oracle/make_ref.sh copies the reference where /root/reference exists)"}
    try:
        proc = subprocess.run(
            [sys.executable, str(script), "--variant", wl["variant"], "--grid", str(wl["grid"]), "--agents", str(wl["agents"])],
            capture_output=True, text=True, timeout=timeout_s, cwd=ROOT,
        )
    except subprocess.TimeoutExpired:
        return {"unavailable": f"oracle/reference_timing.py exceeded {timeout_s:.0f} s"}
    for line in reversed(proc.stdout.splitlines()):
        if line.startswith("{") and '"hot_path"' in line:
            try:
                return json.loads(line)
            except json.JSONDecodeError:
                break
    return {"unavailable": f"oracle/reference_timing.py rc={proc.returncode}: {proc.stderr.strip()[-200:]}"}


def run_reference_arm(args, wl: dict) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # bounded sample per step: the whole run stays within about a minute whatever K and W are
    per_step_budget = min(1.0, 60.0 / max(1, args.warmup + args.steps))
    rates = []
    sample = ""
    for i in range(args.warmup + args.steps):
        rate, sample = cpu_oracle_rate(wl, per_step_budget, threads)
        if i >= args.warmup:
            rates.append(rate)
    value = sum(rates) / len(rates)
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": "agent-steps/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * wl["towns"] * wl["agents"] / value,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": bench_config(wl),
        "cpu_baseline": {
            "value": value,
            "unit": "agent-steps/s",
            "cores": threads,
            "kind": "port",
            "sample": "per step: " + sample + "; the Python reference itself is timed by the GPU arm (cpu_baseline.python_reference)",
        },
        "e2e": {"value": value, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


class Ranks:
    """The process group (or none) with the three things the bench needs from it."""

    def __init__(self, torch, dist, world: int, device) -> None:
        self.torch, self.dist, self.world, self.device = torch, dist, world, device

    def barrier(self) -> None:
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce_max(self, value: float) -> float:
        if self.world == 1:
            return float(value)
        t = self.torch.tensor([value], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(self, tensor):
        if self.world > 1:
            self.dist.all_reduce(tensor, op=self.dist.ReduceOp.SUM)
        return tensor

    def gather(self, tensor) -> list:
        if self.world == 1:
            return [tensor]
        parts = [self.torch.empty_like(tensor) for _ in range(self.world)]
        self.dist.all_gather(parts, tensor)
        return parts


def ticks_per_launch_for(wl: dict, steps: int, override: int | None) -> int:
    """Ticks one launch advances.  A launch is a rollout segment: 128 ticks for the configs[1]-sized batches (the rollout
    length of configs[4]; configs[1] moves only ~22 MB per tick, so SURVEY's caveat asks for a multi-tick
    launch), 8 for the large batches (0.3-0.6 GB of output per tick), 128-tick captured rollouts with the policy in the
    loop (configs[4]: one KPI exchange per 128-tick rollout).  `--steps K` is the number of ticks the caller asks to have timed: the timed region is at least that long (it
    is in fact `reps` launches, >= 200 launches / >= 0.1 s); `--ticks-per-launch` overrides the segment length."""
    del steps
    if override:
        return max(1, override)
    if wl.get("policy"):
        return 128
    return 8 if wl["towns"] * wl["agents"] > 65536 else 128


def measure_tick_workload(key: str, args, ranks: Ranks, rank: int, sampler: ClockSampler | None, tpl: int, warm_ticks: int) -> dict:
    """Device-timed throughput of one tick-only workload through a planned launch repeated back to back."""
    import torch

    from townlet_b200 import capi
    from townlet_b200.engine import DeviceTownBatch
    from townlet_b200.params import DynamicsSpec, RewardSpec

    wl = WORKLOADS[key]
    lib = capi.load_library()
    batch, obs = make_batch(wl, rank)
    rew = RewardSpec(fuse_faint=True)
    dyn = DynamicsSpec() if wl.get("world") else None
    dev = DeviceTownBatch(batch, obs, rew, ranks.device, dyn=dyn)
    n_agents = batch.n_agents
    out_bytes_per_tick = n_agents * (obs.map_elems + obs.n_features + 1) * 4
    ring = max(tpl, -(-RING_BYTES // out_bytes_per_tick))
    ring = -(-ring // tpl) * tpl
    phases = capi.TL_PHASE_WORLD if wl.get("world") else capi.TL_PHASE_ALL
    if os.environ.get("TL_BENCH_PHASES"):  # experiments only: override the phase mask
        phases = int(os.environ["TL_BENCH_PHASES"], 0)
    out = dev.alloc_outputs(phases, ring, summary=False, comps=False, kpi=True)
    actions = None
    if wl.get("world"):
        # seeded action words for every tick of a launch (60 % moves to a random tile, waits, idle agents)
        from townlet_b200.synth import random_actions

        actions = torch.from_numpy(random_actions(batch, tpl, seed=4321 + rank).view("int32")).to(ranks.device)
    plan = dev.plan(out, phases, tpl, actions=actions)

    plan.launch_many(max(1, -(-warm_ticks // tpl)))  # W untimed warm-up ticks (at least one launch)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    plan.launch_many(8)
    e1.record()
    torch.cuda.synchronize()
    est_s = max(e0.elapsed_time(e1) * 1e-3 / 8, 1e-6)
    reps = int(ranks.reduce_max(max(args.min_launches, math.ceil(1.1 * args.min_seconds / est_s), -(-args.steps // tpl))))

    ranks.barrier()
    launches0 = lib.tl_launch_count()
    t0 = time.perf_counter()
    e0.record()
    plan.launch_many(reps)  # one C call: `reps` cudaLaunchKernel, nothing else
    e1.record()
    ranks.barrier()
    t1 = time.perf_counter()
    launches = lib.tl_launch_count() - launches0
    elapsed_ms = ranks.reduce_max(e0.elapsed_time(e1))
    clocks = sampler.window(t0, t1) if sampler is not None else None
    ticks = reps * tpl
    value = ranks.world * n_agents * ticks / (elapsed_ms * 1e-3)

    # per-launch distribution: the same launches with a CUDA event after each one (a separate, untimed-for-value pass)
    n_dist = min(reps, 256)
    plan.launch_many(n_dist, timed=True)
    torch.cuda.synchronize()
    per_launch = sorted(float(x) for x in plan.launch_ms(n_dist))

    kpi_total = ranks.reduce_sum(out.kpi[:tpl].sum(dim=(0, 1)).double())
    peak, peak_src = load_peaks()
    avg_launch_s = elapsed_ms * 1e-3 / reps  # includes the gaps between launches
    algo_bytes = wl["bytes_per_agent_step"] * n_agents * tpl
    achieved = algo_bytes / avg_launch_s / 1e9
    grid, threads, smem = plan.grid
    roofline = {
        "bound": "hbm",
        "achieved": achieved,
        "peak": peak,
        "unit": "GB/s",
        "frac": achieved / peak,
        "traffic": None,
        "kernel": f"tick_kernel_roles<{wl['variant']}{', world' if wl.get('world') else ''}>",
        "algorithmic_bytes_per_launch": algo_bytes,
        "avg_launch_ms": avg_launch_s * 1e3,
        "launch_ms_median_min_max": [per_launch[len(per_launch) // 2], per_launch[0], per_launch[-1]],
        "launch_grid": {"ctas": grid, "threads": threads, "dynamic_smem_bytes": smem},
        "peak_source": peak_src,
    }
    traffic_file = ROOT / "profiles" / "traffic.json"
    if traffic_file.exists():
        try:
            t = json.loads(traffic_file.read_text()).get(key)
            if t:  # ncu dram__bytes_read.sum + dram__bytes_write.sum per agent-step x agent-steps of one launch
                roofline["traffic"] = t["dram_bytes_per_agent_step"] * n_agents * tpl
                roofline["traffic_source"] = "profiles/traffic.json (ncu --set full, bytes per agent-step x agent-steps per launch)"
                # the same fraction by measured DRAM bytes: SoA reads that stay in L2 count in `achieved` but not here
                roofline["frac_by_traffic"] = roofline["traffic"] / avg_launch_s / 1e9 / peak
        except Exception:
            pass
    plan.close()
    return {
        "wl": wl, "batch": batch, "obs": obs, "rew": rew, "dev": dev, "n_agents": n_agents,
        "value": value, "elapsed_ms": elapsed_ms, "ticks": ticks, "reps": reps, "tpl": tpl, "ring": ring,
        "ring_bytes": ring * out_bytes_per_tick, "launches": int(launches), "clocks": clocks, "roofline": roofline,
        "kpi_total": kpi_total, "kpi_first_launch": out.kpi[:tpl],
    }


def measure_policy_workload(key: str, args, ranks: Ranks, rank: int, sampler: ClockSampler | None, tpl: int) -> dict:
    """configs[4] per GPU: a captured rollout (tick + batched policy step per tick, GAE at the end) replayed back to back,
    with the KPI all-reduce after every rollout."""
    import torch

    from townlet_b200 import capi
    from townlet_b200.engine import DeviceTownBatch
    from townlet_b200.params import DynamicsSpec, RewardSpec
    from townlet_b200.rollout import PolicyMLP, RolloutCapture

    wl = WORKLOADS[key]
    lib = capi.load_library()
    batch, obs = make_batch(wl, rank)
    rew = RewardSpec(fuse_faint=True)
    dyn = DynamicsSpec() if wl.get("world") else None
    dev = DeviceTownBatch(batch, obs, rew, ranks.device, dyn=dyn)
    n_agents = batch.n_agents
    torch.manual_seed(1234 + rank)
    capture = RolloutCapture(dev, PolicyMLP(obs.n_features, obs.map_shape, action_dim=9), act_on_world=bool(wl.get("world")),
                             map_dtype="bf16" if wl.get("map_bf16") else "f32")
    out = capture.alloc_outputs(tpl, summary=False, comps=False, kpi=True)
    before = lib.tl_launch_count()
    graphed = capture.build_graph(tpl, out)
    graph_launches = graphed.library_launches  # kernels of this library inside one replay (counted while capturing)
    del before

    def rollout() -> None:
        ro = graphed.replay()
        ranks.reduce_sum(ro.kpi.sum(dim=(0, 1)).double())  # end-of-rollout KPI exchange (the path's only communication)

    for _ in range(2):
        rollout()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2):
        rollout()
    e1.record()
    torch.cuda.synchronize()
    est_s = max(e0.elapsed_time(e1) * 1e-3 / 2, 1e-6)
    reps = int(ranks.reduce_max(max(8, math.ceil(1.1 * args.min_seconds / est_s), -(-args.steps // tpl))))
    ranks.barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        rollout()
    e1.record()
    ranks.barrier()
    t1 = time.perf_counter()
    elapsed_ms = ranks.reduce_max(e0.elapsed_time(e1))
    clocks = sampler.window(t0, t1) if sampler is not None else None
    ticks = reps * tpl
    value = ranks.world * n_agents * ticks / (elapsed_ms * 1e-3)
    peak, peak_src = load_peaks()
    avg_s = elapsed_ms * 1e-3 / reps
    algo_bytes = wl["bytes_per_agent_step"] * n_agents * tpl
    roofline = {
        "bound": "hbm", "achieved": algo_bytes / avg_s / 1e9, "peak": peak, "unit": "GB/s", "frac": algo_bytes / avg_s / 1e9 / peak,
        "traffic": None,
        "kernel": "one rollout = per tick: tick_kernel_roles<hybrid, mirror> + batched policy step + sample_actions_kernel; gae_kernel at the end "
        "(bytes counted: the tick kernel's only)",
        "algorithmic_bytes_per_launch": algo_bytes, "avg_launch_ms": avg_s * 1e3, "peak_source": peak_src,
    }
    return {
        "wl": wl, "batch": batch, "obs": obs, "rew": rew, "dev": dev, "n_agents": n_agents,
        "value": value, "elapsed_ms": elapsed_ms, "ticks": ticks, "reps": reps, "tpl": tpl, "ring": capture.slots_for(tpl),
        "ring_bytes": 0, "launches": int(reps * graph_launches), "clocks": clocks, "roofline": roofline,
        "kpi_total": ranks.reduce_sum(graphed.rollout.kpi.sum(dim=(0, 1)).double()), "kpi_first_launch": graphed.rollout.kpi,
        "policy": capture.describe(),
    }


def measure_e2e(key: str, res: dict, args, ranks: Ranks, rank: int, local_rank: int) -> dict:
    """The same tick through the host-buffer C ABI (``tl_host_submit`` / ``tl_host_wait``): every step uploads a whole
    batch from pinned HOST memory and brings its observations, rewards, flags, KPI rows and mutated state back to
    pinned HOST buffers — all copies inside the timed region (wall clock, max over ranks).  `depth` independent
    batches of the workload's size are in flight, one per slot, so the upload of one step overlaps the download of
    another; a step is still one full batch up, one tick, one full set of outputs down.  The map crosses the bus as
    float32, uint8 or bits and is delivered as the reference's float32 tensor (expanded by the
    context's host threads) or as the wire bytes."""
    import torch

    from townlet_b200 import capi
    from townlet_b200.engine import HostEngine

    wl, obs, rew = WORKLOADS[key], res["obs"], res["rew"]
    cores = os.cpu_count() or 1
    threads = max(1, min(int(os.environ.get("TL_E2E_THREADS", "12")), cores // ranks.world - 1))
    depth = max(1, min(int(os.environ.get("TL_E2E_DEPTH", "3")), capi.TL_HOST_SLOTS))
    if os.environ.get("TL_E2E_NT"):
        capi.set_option("stream_stores", int(os.environ["TL_E2E_NT"]))
    host = HostEngine(obs, rew, local_rank, threads=threads)
    batches = [host.pinned_batch(make_batch(wl, rank + k * ranks.world)[0]) for k in range(depth)]
    n_agents = batches[0].n_agents
    phases = capi.TL_PHASE_ALL
    steps = args.e2e_steps

    # the relationship / rivalry ledgers (65 % of the batch bytes) change only on chat / conflict events: a caller that
    # keeps them resident uploads the other arrays every step
    ledgers = {"tie_count", "tie_embed", "tie_trust", "tie_fam", "tie_riv", "riv_count", "riv_embed", "riv_score"}
    not_ledgers = tuple(name for name in capi.FIELD_INDEX if name not in ledgers)

    def run(wire: str, deliver: str, pipelined: bool, dirty=None) -> dict:
        outs = [host.alloc_outputs(b, phases, 1, summary=False, comps=False, pin=True, map_wire=wire if deliver == "wire" else None)
                for b in batches]
        fmt = "wire" if deliver == "wire" else "f32"
        prepared = [host.make_step(b, o, phases, 1, wire=wire, map_format=fmt, dirty=dirty) for b, o in zip(batches, outs)]
        spent = {"submit": 0.0, "wait": 0.0}

        def loop(n: int) -> None:
            if not pipelined:
                for i in range(n):
                    host.submit(0, step=prepared[0])
                    host.wait(0)
                return
            for i in range(n + depth):
                slot = i % depth
                ta = time.perf_counter()
                if i >= depth:
                    host.wait(slot)
                tb = time.perf_counter()
                if i < n:
                    host.submit(slot, step=prepared[slot])
                spent["wait"] += tb - ta
                spent["submit"] += time.perf_counter() - tb

        loop(2 * depth)
        spent["submit"] = spent["wait"] = 0.0
        ranks.barrier()
        t0 = time.perf_counter()
        loop(steps)
        dt = ranks.reduce_max(time.perf_counter() - t0)
        return {"value": ranks.world * n_agents * steps / dt, "ms_per_step": 1e3 * dt / steps,
                "h2d_bytes_per_step": host.last_h2d_bytes, "d2h_bytes_per_step": host.last_d2h_bytes,
                "host_ms_per_step": {k: 1e3 * v / steps for k, v in spent.items()} if pipelined else None}

    binary = wl["variant"] != "compact"  # compact cells are counts only when normalize_counts is off (default: on -> 0 / 1)
    formats = {
        "f32 wire, synchronous (one step at a time)": run("f32", "f32", False),
        "f32 wire": run("f32", "f32", True),
        "u8 wire -> float32 on the host": run("u8", "f32", True),
        "u8 wire, delivered as uint8": run("u8", "wire", True),
    }
    formats["bits wire -> float32 on the host"] = run("bits", "f32", True)
    formats["bits wire, delivered as bits"] = run("bits", "wire", True)
    resident = {
        "bits wire -> float32 on the host, ledgers resident": run("bits", "f32", True, not_ledgers),
        "bits wire, delivered as bits, ledgers resident": run("bits", "wire", True, not_ledgers),
    }
    del binary
    host.close()
    f32_names = [k for k in formats if "delivered as" not in k]
    best = max(f32_names, key=lambda k: formats[k]["value"])
    return {
        "value": formats[best]["value"],
        "unit": "agent-steps/s",
        "h2d_bytes_per_step": formats[best]["h2d_bytes_per_step"],
        "d2h_bytes_per_step": formats[best]["d2h_bytes_per_step"],
        "steps": steps,
        "api": f"tl_host_submit / tl_host_wait, pinned host buffers, {depth} batches in flight, {threads} host threads; "
        f"headline = the fastest format that delivers the reference's float32 tensors: {best}",
        "formats": formats,
        "resident_ledgers": resident,
    }


def measure_e2e_world(res: dict, args, ranks: Ranks) -> dict:
    """World phases run on device-resident batches: per step the action words come from pinned host memory and the
    observations / rewards go back to pinned host memory through the public DeviceTownBatch API."""
    import torch

    from townlet_b200 import capi
    from townlet_b200.synth import random_actions

    dev, batch = res["dev"], res["batch"]
    n_agents = batch.n_agents
    phases = capi.TL_PHASE_WORLD
    eout = dev.alloc_outputs(phases, 1, summary=False, comps=False, kpi=True)
    h_act = torch.zeros(n_agents, dtype=torch.int32, pin_memory=True)
    h_act.copy_(torch.from_numpy(random_actions(batch, 1, seed=99).view("int32"))[0])
    d_act = torch.empty(n_agents, dtype=torch.int32, device=ranks.device)
    h_out = {k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True) for k, v in vars(eout).items() if v is not None}

    def step() -> None:
        d_act.copy_(h_act, non_blocking=True)
        dev.tick(eout, phases, 1, actions=d_act)
        for k, v in h_out.items():
            v.copy_(getattr(eout, k), non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(3):
        step()
    ranks.barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        step()
    dt = ranks.reduce_max(time.perf_counter() - t0)
    return {
        "value": ranks.world * n_agents * args.e2e_steps / dt,
        "unit": "agent-steps/s",
        "h2d_bytes_per_step": int(h_act.numel() * 4),
        "d2h_bytes_per_step": int(sum(v.numel() * v.element_size() for v in h_out.values())),
        "steps": args.e2e_steps,
        "api": "DeviceTownBatch.tick (device-resident towns; action words in, observations out through pinned host buffers)",
    }


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20, help="K: ticks to time at least (the timed region is `reps` launches of a rollout segment, never shorter)")
    ap.add_argument("--warmup", type=int, default=5, help="W: untimed warm-up ticks (at least 3)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--ticks-per-launch", type=int, default=None, help="override: ticks per launch (default 128; 8 for the large batches)")
    ap.add_argument("--min-launches", type=int, default=200, help="the timed region repeats the K-tick launch at least this often")
    ap.add_argument("--min-seconds", type=float, default=0.1, help="... and for at least this much device time")
    ap.add_argument("--e2e-steps", type=int, default=100)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the other configurations (`workloads`)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (kernel A/B runs)")
    ap.add_argument("--option", action="append", default=[], metavar="NAME=VALUE",
                    help="a library switch for an A/B run (tl_set_option; defaults are the product configuration)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]

    if args.impl == "reference":
        run_reference_arm(args, wl)
        return
    if args.option:
        from townlet_b200 import capi

        for item in args.option:
            name, _, value = item.partition("=")
            capi.set_option(name, int(value))

    import torch
    import torch.distributed as dist

    from townlet_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    if args.gpus != world and rank == 0 and world > 1:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE {world}", file=sys.stderr)
    device = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(device)
    capi.load_library()
    ranks = Ranks(torch, dist, world, device)
    warm = max(3, args.warmup)

    sampler = ClockSampler(local_rank)  # running before, during and after every timed region
    sampler.start()
    sampler.ready.wait(timeout=10)

    def measure(key: str) -> dict:
        w = WORKLOADS[key]
        tpl = ticks_per_launch_for(w, args.steps, args.ticks_per_launch)
        if w.get("policy"):
            return measure_policy_workload(key, args, ranks, rank, sampler, tpl)
        return measure_tick_workload(key, args, ranks, rank, sampler, tpl, warm)

    res = measure(args.workload)

    # ---- e2e: host buffers through the C ABI (H2D + kernel + D2H per step) ----
    if args.no_e2e or (wl.get("world") and wl.get("policy")):
        e2e = None  # (world + policy: the rollout never leaves the device; see c5 for the host-buffer figure of the same batch)
    elif wl.get("world"):
        e2e = measure_e2e_world(res, args, ranks)
    else:
        e2e = measure_e2e(args.workload, res, args, ranks, rank, local_rank)

    # ---- the other configurations of BASELINE.json, same method, same process ----
    workloads = {}
    n_agents_main, launches_main = res["n_agents"], res["launches"]
    res.pop("dev", None)
    if not args.no_extra:
        for key in EXTRA_WORKLOADS:
            if key == args.workload:
                continue
            torch.cuda.empty_cache()
            r = measure(key)
            workloads[key] = {
                "workload": r["wl"]["name"],
                "value": r["value"],
                "unit": "agent-steps/s",
                "ms_per_tick": r["elapsed_ms"] / r["ticks"],
                "roofline_frac": r["roofline"]["frac"],
                "achieved_gbs": r["roofline"]["achieved"],
                "frac_by_traffic": r["roofline"].get("frac_by_traffic"),
                "bytes_per_agent_step": r["wl"]["bytes_per_agent_step"],
                "ticks_per_launch": r["tpl"],
                "reps": r["reps"],
                "timed_region_s": r["elapsed_ms"] * 1e-3,
                "gpu_launches": r["launches"],
                "sm_mhz": (r["clocks"] or {}).get("sm_mhz"),
                "reasons": (r["clocks"] or {}).get("reasons"),
            }
            if "policy" in r:
                workloads[key]["policy"] = r["policy"]
            del r
    sampler.stop()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        rate, sample = cpu_oracle_rate(wl, args.cpu_seconds, threads)
        cpu_baseline = {"value": rate, "unit": "agent-steps/s", "cores": threads, "kind": "port", "sample": sample,
                        "python_reference": python_reference_baseline(wl)}

    # the KPI all-reduce checked against the per-rank sums (gathered separately): the total every rank holds must be
    # the sum of the shares
    own = res["kpi_first_launch"].sum(dim=(0, 1)).double()
    shares = torch.stack(ranks.gather(own)).sum(dim=0)
    kpi_ok = bool(torch.isfinite(res["kpi_total"]).all().item()) and bool(torch.allclose(shares, res["kpi_total"], rtol=1e-9, atol=1e-9))

    if rank == 0:
        line = {
            "metric": METRIC,
            "value": res["value"],
            "unit": "agent-steps/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": warm,
            "ms_per_step": res["elapsed_ms"] / res["ticks"],
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f64",
            "data": "synthetic",
            "config": bench_config(wl),
            "timing": {
                "ticks_per_launch": res["tpl"],
                "reps": res["reps"],
                "timed_ticks": res["ticks"],
                "timed_region_s": res["elapsed_ms"] * 1e-3,
                "how": "one planned launch of `ticks_per_launch` ticks repeated `reps` times back to back by one C call between two "
                "CUDA events on the launching stream (timed_ticks >= the K steps asked for); barrier + synchronize on both sides; "
                "max over ranks",
                "output_ring_ticks": res["ring"],
                "output_ring_mib": res["ring_bytes"] / 2**20,
                "soa_state_mib": res["batch"].layout.total_bytes / 2**20,
            },
            "clocks": res["clocks"],
            "e2e": e2e,
            "gpu_launches": int(launches_main),
            "roofline": res["roofline"],
            "cpu_baseline": cpu_baseline,
            "workloads": workloads,
            "kpi_check": {"reward_sum": float(res["kpi_total"][0].item()), "terminated": float(res["kpi_total"][5].item()),
                          "consistent": kpi_ok},
        }
        if "policy" in res:
            line["policy"] = res["policy"]
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
