#!/usr/bin/env python
"""bench.py — agent-steps/s of the per-tick hot path (obs build + need decay + reward).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3-compact|c3-full|c4|c5|c2-world|c5-world]
    python bench.py --impl reference ...      # the CPU arm (oracle port, all host threads)

A "step" is ONE TICK of the whole synthetic batch: need decay -> faint
predicate -> reward -> observation build for every agent.  One launch advances
``--ticks-per-launch`` ticks (the kernel re-reads state and re-stages the town
grids every tick; only launch overhead is amortised).  Each tick writes its
observation tensors to its own slot of a rollout ring that is larger than L2,
so every output byte is drained to HBM.

One JSON line is printed by rank 0 (see README / DESIGN.md §Measurement).
"""

from __future__ import annotations

import argparse
import json
import os
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
    # configs[4] with the sampled actions applied by the next tick (moves) and ledger decay: the policy drives the world
    "c5-world": dict(variant="hybrid", towns=8192, grid=48, agents=8, bytes_per_agent_step=2712 + 24, policy=True, world=True,
                     name="configs[4] + the policy's actions applied on the device: 8,192 towns per GPU x 48x48 x 8 agents"),
}
L2_BYTES = 126 * 1024 * 1024
DISTINCT_TOWNS = 256  # towns generated on the host; larger batches tile them


def load_peaks() -> tuple[float, str]:
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        try:
            return float(json.loads(path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def make_batch(wl: dict, rank: int):
    from townlet_b200.params import ObsSpec
    from townlet_b200.synth import synth_batch, tile_batch

    obs = ObsSpec(variant=wl["variant"])
    distinct = min(wl["towns"], DISTINCT_TOWNS)
    base = synth_batch(distinct, wl["grid"], wl["agents"], obs, first_town=rank * wl["towns"])
    reps = wl["towns"] // distinct
    return (tile_batch(base, reps) if reps > 1 else base), obs


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons of one GPU while the timed region runs."""

    def __init__(self, index: int) -> None:
        super().__init__(daemon=True)
        self.index = index
        self.samples: list[int] = []
        self.reasons: set[str] = set()
        self.max_mhz = 0
        self._halt = threading.Event()
        self.ready = threading.Event()
        self.go = threading.Event()

    def run(self) -> None:
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = int(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            names = {
                nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            }
            self.ready.set()
            self.go.wait()
            while True:  # at least one sample, taken while the timed launches are queued / running
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
                if self._halt.wait(0.005):
                    break
        except Exception as exc:  # pragma: no cover - NVML missing
            self.reasons.add(f"nvml_unavailable:{type(exc).__name__}")
            self.ready.set()

    def stop(self) -> dict:
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz or None, "reasons": sorted(self.reasons)}


def cpu_oracle_rate(wl: dict, budget_s: float, threads: int) -> tuple[float, str]:
    """agent-steps/s of the C oracle (CPU restatement of the reference) on a bounded sample."""
    from oracle.oracle import oracle_tick
    from townlet_b200 import capi
    from townlet_b200.params import DynamicsSpec, RewardSpec

    sample_towns = min(wl["towns"], DISTINCT_TOWNS)
    small = dict(wl, towns=sample_towns)
    batch, obs = make_batch(small, 0)
    rew = RewardSpec(fuse_faint=True)
    phases, extra = capi.TL_PHASE_ALL, {}
    if wl.get("world"):
        from townlet_b200.synth import random_actions

        phases = capi.TL_PHASE_WORLD
        extra = dict(dyn=DynamicsSpec(), actions=random_actions(batch, 4, seed=4321))
    oracle_tick(batch, obs, rew, phases, 4, n_threads=threads, want_comps=False, **extra)  # warm-up
    ticks_done = 0
    t0 = time.perf_counter()
    while True:
        oracle_tick(batch, obs, rew, phases, 4, n_threads=threads, want_comps=False, **extra)
        ticks_done += 4
        if time.perf_counter() - t0 >= budget_s:
            break
    dt = time.perf_counter() - t0
    rate = batch.n_agents * ticks_done / dt
    return rate, f"{sample_towns} towns x {wl['agents']} agents x {ticks_done} ticks, {dt:.1f} s, C oracle (OpenMP over towns)"


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
        "metric": "agent-steps/sec (obs build+needs+reward)",
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
        "config": {"workload": wl["name"], "variant": wl["variant"], "towns": wl["towns"], "agents_per_town": wl["agents"], "grid": wl["grid"]},
        "cpu_baseline": {
            "value": value,
            "unit": "agent-steps/s",
            "cores": threads,
            "kind": "port",
            "sample": sample + "; the Python reference itself measured 1.1-3.7 k agent-steps/s per core (BASELINE.md §2)",
        },
        "e2e": {"value": value, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8192)
    ap.add_argument("--warmup", type=int, default=256)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--ticks-per-launch", type=int, default=None,
                    help="ticks advanced by one launch (default: 128 for the configs[1]-sized workloads, a rollout as in "
                    "configs[4]; 32 for the larger ones, whose 128-tick output ring would not fit)")
    ap.add_argument("--e2e-steps", type=int, default=50)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]

    if args.impl == "reference":
        run_reference_arm(args, wl)
        return

    import torch
    import torch.distributed as dist

    from townlet_b200 import capi
    from townlet_b200.engine import DeviceTownBatch, HostEngine
    from townlet_b200.params import DynamicsSpec, RewardSpec

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
    lib = capi.load_library()

    # every rank owns its own block of towns (towns are independent: no collective in the tick)
    batch, obs = make_batch(wl, rank)
    rew = RewardSpec(fuse_faint=True)
    dyn = DynamicsSpec() if wl.get("world") else None
    dev = DeviceTownBatch(batch, obs, rew, device, dyn=dyn)
    n_agents = batch.n_agents
    if args.ticks_per_launch is None:
        args.ticks_per_launch = 128 if wl["towns"] * wl["agents"] <= 16384 and not wl.get("policy") else 32
    tpl = max(1, min(args.ticks_per_launch, args.steps))
    steps = (args.steps // tpl) * tpl
    warm = max(3, args.warmup)
    out_bytes_per_tick = n_agents * (obs.map_elems + obs.n_features + 1) * 4
    ring = max(tpl, -(-2 * L2_BYTES // out_bytes_per_tick))  # ring of tick slots > 2 x L2
    ring = -(-ring // tpl) * tpl
    if wl.get("policy"):
        ring = tpl  # a rollout buffer of tpl ticks, consumed by the policy every tick
    out = dev.alloc_outputs(capi.TL_PHASE_ALL, ring, summary=False, comps=False, kpi=True)

    phases = capi.TL_PHASE_WORLD if wl.get("world") else capi.TL_PHASE_ALL
    if os.environ.get("TL_BENCH_PHASES"):  # experiments only: override the phase mask
        phases = int(os.environ["TL_BENCH_PHASES"], 0)
    action_words = pristine = None
    if wl.get("world") and not wl.get("policy"):
        # seeded action words for every tick of a launch (60 % moves to a random tile, waits, idle agents); every
        # launch starts from the pristine state (an episode reset), so the ledgers are live in every launch
        from townlet_b200.synth import random_actions

        action_words = torch.from_numpy(random_actions(batch, tpl, seed=4321 + rank).view("int32")).to(device)
        pristine = dev.arena.clone()

    capture = None
    if wl.get("policy"):
        # rollout capture: one tick launch + one batched policy forward per tick, GAE scan and KPI
        # all-reduce at the end of every rollout of `tpl` ticks (townlet_b200/rollout.py)
        from townlet_b200.rollout import PolicyMLP, RolloutCapture

        torch.manual_seed(1234 + rank)
        capture = RolloutCapture(dev, PolicyMLP(obs.n_features, obs.map_shape, action_dim=9), act_on_world=bool(wl.get("world")))
        graphed = None
        graph_launches = 0
        if not os.environ.get("TL_NO_GRAPH"):
            before = lib.tl_launch_count()
            graphed = capture.build_graph(tpl, out)
            graph_launches = 3 * tpl + 1  # kernels of this library in the captured rollout: per tick 1 tick + 2 cast/pad, + 1 GAE scan
            del before

    def launch(i: int) -> None:
        if capture is not None:
            ro = graphed.replay() if graphed is not None else capture.capture(tpl, out)
            totals = ro.kpi.sum(dim=(0, 1)).double()
            if world > 1:
                dist.all_reduce(totals, op=dist.ReduceOp.SUM)
            return
        if pristine is not None:
            dev.arena.copy_(pristine)
        dev.tick(out, phases, tpl, tick0=(i * tpl) % ring, actions=action_words)

    for i in range(-(-warm // tpl)):
        launch(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

    n_launch = steps // tpl
    events = [torch.cuda.Event(enable_timing=True) for _ in range(n_launch + 1)]
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.ready.wait(timeout=10)
    launches0 = lib.tl_launch_count()
    events[0].record()
    for i in range(n_launch):
        launch(i)
        events[i + 1].record()
        if i == 0:
            sampler.go.set()  # the first sample falls inside the timed region
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches = lib.tl_launch_count() - launches0
    if wl.get("policy") and graphed is not None:
        launches += n_launch * graph_launches  # graph replays re-issue the captured launches
    clocks = sampler.stop()
    elapsed_ms = events[0].elapsed_time(events[-1])
    per_launch_ms = [events[i].elapsed_time(events[i + 1]) for i in range(n_launch)]
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    value = world * n_agents * steps / (elapsed_ms * 1e-3)

    # end-of-rollout KPI gather (the only communication of the path; off the tick)
    kpi_total = out.kpi[:tpl].sum(dim=(0, 1)).double()
    if world > 1:
        dist.all_reduce(kpi_total, op=dist.ReduceOp.SUM)

    # ---- roofline of the dominant kernel (tick_kernel) ------------------------
    peak, peak_src = load_peaks()
    avg_launch_s = (sum(per_launch_ms) / len(per_launch_ms)) * 1e-3
    algo_bytes = wl["bytes_per_agent_step"] * n_agents * tpl
    achieved = algo_bytes / avg_launch_s / 1e9
    roofline = {
        "bound": "hbm",
        "achieved": achieved,
        "peak": peak,
        "unit": "GB/s",
        "frac": achieved / peak,
        "traffic": None,
        "kernel": f"tick_kernel_roles<{wl['variant']}{', world' if wl.get('world') else ''}>"
        + (" + batched policy forward + gae_kernel" if wl.get("policy") else ""),
        "algorithmic_bytes_per_launch": algo_bytes,
        "avg_launch_ms": avg_launch_s * 1e3,
        "peak_source": peak_src,
    }
    traffic_file = ROOT / "profiles" / "traffic.json"
    if traffic_file.exists():
        try:
            t = json.loads(traffic_file.read_text()).get(args.workload)
            if t:  # ncu dram__bytes_read.sum + dram__bytes_write.sum per agent-step x agent-steps of one launch
                roofline["traffic"] = t["dram_bytes_per_agent_step"] * n_agents * tpl
                roofline["traffic_source"] = "profiles/traffic.json (ncu --set full, bytes per agent-step x agent-steps per launch)"
                # the same fraction by measured DRAM bytes: SoA reads that stay in L2 count in `achieved` but not here
                roofline["frac_by_traffic"] = roofline["traffic"] / avg_launch_s / 1e9 / peak
        except Exception:
            pass

    # ---- e2e: host buffers through tl_host_tick (H2D + kernel + D2H per step) --
    e2e = None
    if wl.get("world") and wl.get("policy"):
        e2e = None  # the rollout never leaves the device; see c5 for the host-buffer figure of the same batch
    elif wl.get("world"):
        # the world phases run on device-resident batches: per step the action words come from pinned host memory
        # and the observations / rewards go back to pinned host memory through the public DeviceTownBatch API
        eout = dev.alloc_outputs(phases, 1, summary=False, comps=False, kpi=True)
        h_act = torch.zeros(n_agents, dtype=torch.int32, pin_memory=True)
        h_act.copy_(action_words[0].cpu())
        d_act = torch.empty(n_agents, dtype=torch.int32, device=device)
        h_out = {k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True) for k, v in vars(eout).items() if v is not None}

        def e2e_step() -> None:
            d_act.copy_(h_act, non_blocking=True)
            dev.tick(eout, phases, 1, actions=d_act)
            for k, v in h_out.items():
                v.copy_(getattr(eout, k), non_blocking=True)
            torch.cuda.synchronize()

        for _ in range(3):
            e2e_step()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {
            "value": world * n_agents * args.e2e_steps / float(te.item()),
            "unit": "agent-steps/s",
            "h2d_bytes_per_step": int(h_act.numel() * 4),
            "d2h_bytes_per_step": int(sum(v.numel() * v.element_size() for v in h_out.values())),
            "api": "DeviceTownBatch.tick (device-resident towns; action words in, observations out through pinned host buffers)",
        }
    elif rank == 0 or world > 1:
        host = HostEngine(obs, rew, local_rank)
        pinned = torch.empty(batch.layout.total_bytes, dtype=torch.uint8, pin_memory=True)
        pinned.numpy()[:] = batch.arena[: batch.layout.total_bytes]
        from townlet_b200.soa import TownBatch

        hbatch = TownBatch(batch.layout, batch.grid_w, batch.grid_h, pinned.numpy())
        hout = host.alloc_outputs(hbatch, capi.TL_PHASE_ALL, 1, summary=False, comps=False, pin=True)
        for _ in range(3):
            host.tick(hbatch, hout, capi.TL_PHASE_ALL, 1)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            host.tick(hbatch, hout, capi.TL_PHASE_ALL, 1)
        dt = time.perf_counter() - t0
        te = torch.tensor([dt], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {
            "value": world * n_agents * args.e2e_steps / float(te.item()),
            "unit": "agent-steps/s",
            "h2d_bytes_per_step": host.last_h2d_bytes,
            "d2h_bytes_per_step": host.last_d2h_bytes,
            "api": "tl_host_tick (pinned host buffers, synchronous)",
        }
        host.close()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        rate, sample = cpu_oracle_rate(wl, args.cpu_seconds, threads)
        cpu_baseline = {"value": rate, "unit": "agent-steps/s", "cores": threads, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": "agent-steps/sec (obs build+needs+reward)",
            "value": value,
            "unit": "agent-steps/s",
            "n_gpus": world,
            "steps": steps,
            "warmup": warm,
            "ms_per_step": elapsed_ms / steps,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f64",
            "data": "synthetic",
            "config": {
                "workload": wl["name"],
                "variant": wl["variant"],
                "towns_per_gpu": wl["towns"],
                "agents_per_town": wl["agents"],
                "grid": wl["grid"],
                "ticks_per_launch": tpl,
                "l2": f"outputs go to a {ring}-tick rollout ring ({ring * out_bytes_per_tick / 2**20:.0f} MiB > L2 126 MiB); "
                f"SoA state ({batch.layout.total_bytes / 2**20:.1f} MiB) is re-read every tick",
                "partition": "contiguous block of towns per rank, no collective in the tick; KPI all-reduce after the rollout",
            },
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "kpi_check": {"reward_sum": float(kpi_total[0].item()), "terminated": float(kpi_total[5].item())},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
