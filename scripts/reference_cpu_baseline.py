#!/usr/bin/env python
"""Time the REFERENCE's own Python implementation of the hot path.

Runs only where the reference checkout exists (the build container); the GPU box uses the C oracle
instead (`bench.py --impl reference`).  Writes profiles/<tag>_reference_python_baseline.json.

    python scripts/reference_cpu_baseline.py r1
"""
from __future__ import annotations

import json
import multiprocessing as mp
import os
import platform
import subprocess
import sys
import time
from pathlib import Path

os.environ.setdefault("OMP_NUM_THREADS", "1")  # one worker per core: no BLAS / OpenMP oversubscription
os.environ.setdefault("MKL_NUM_THREADS", "1")
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r1"


def hot_path_rate(variant: str, grid: int, agents: int, iters: int = 30, town_index: int = 0) -> dict:
    """build_batch -> _apply_need_decay -> RewardEngine.compute on one seeded town."""
    import make_golden as mg  # copies the reference to a temp dir and puts it on sys.path
    from townlet.rewards.engine import RewardEngine
    from townlet.world.observations.service import WorldObservationService

    from townlet_b200.synth import synth_town

    config = mg.make_config(variant, employment__enforce_job_loop=False)
    town = synth_town(town_index, grid, agents)
    world = mg.materialise(town, config)
    mg.patch_employment(world, town)
    service = WorldObservationService(config=config)
    engine = RewardEngine(config)
    for _ in range(3):
        service.build_batch(world, {})
    parts = {"obs": 0.0, "decay": 0.0, "reward": 0.0}
    for _ in range(iters):
        world.tick += 1
        t0 = time.perf_counter()
        service.build_batch(world, {})
        t1 = time.perf_counter()
        world._apply_need_decay()
        t2 = time.perf_counter()
        engine.compute(world, {})
        t3 = time.perf_counter()
        parts["obs"] += t1 - t0
        parts["decay"] += t2 - t1
        parts["reward"] += t3 - t2
    total = sum(parts.values())
    return {
        "variant": variant, "grid": grid, "agents": agents, "iterations": iters,
        "agent_steps_per_s": agents * iters / total,
        "ms_per_batch": {k: 1e3 * v / iters for k, v in parts.items()},
    }


def _worker(args):
    return hot_path_rate(*args)["agent_steps_per_s"]


def main() -> None:
    import make_golden as mg

    out: dict = {
        "where": "build container (no GPU), single run, noisy",
        "cpu": platform.processor() or platform.machine(),
        "cores": os.cpu_count(),
        "python": platform.python_version(),
        "hot_path_single_core": [],
    }
    for variant, grid, agents in (("hybrid", 48, 8), ("compact", 48, 10), ("full", 48, 10), ("hybrid", 128, 64)):
        res = hot_path_rate(variant, grid, agents, iters=20 if agents > 16 else 40)
        out["hot_path_single_core"].append(res)
        print(f"{variant} {grid}x{grid} x {agents}: {res['agent_steps_per_s']:.0f} agent-steps/s per core", flush=True)
    cores = os.cpu_count() or 1
    with mp.get_context("fork").Pool(cores) as pool:
        rates = pool.map(_worker, [("hybrid", 48, 8, 40, i) for i in range(cores)])
    out["hot_path_all_cores_hybrid_48x8"] = {"workers": cores, "sum_agent_steps_per_s": sum(rates)}
    print(f"hybrid 48x48 x 8 on {cores} workers: {sum(rates):.0f} agent-steps/s summed", flush=True)

    # the reference tick loop as shipped (0 agents)
    ref = Path.cwd()  # make_golden chdir'ed into the writable reference copy
    proc = subprocess.run(
        [sys.executable, "scripts/benchmark_tick.py", "configs/examples/poc_hybrid.yaml", "--ticks", "100"],
        cwd=ref, env={**os.environ, "PYTHONPATH": str(ref / "src")}, capture_output=True, text=True,
    )
    summary = [l for l in proc.stdout.splitlines() if l.startswith("average_tick") or l.startswith("ticks")]
    out["benchmark_tick_as_shipped"] = {"stdout_tail": summary[:3], "returncode": proc.returncode}
    print("benchmark_tick.py:", out["benchmark_tick_as_shipped"]["stdout_tail"], flush=True)

    # the same loop, populated with 8 agents, telemetry stubbed
    from townlet.core.sim_loop import SimulationLoop

    config = mg.make_config("hybrid", runtime__telemetry__provider="stub")
    loop = SimulationLoop(config)
    for k, otype in enumerate(("fridge", "stove", "bed", "shower", "fridge", "stove", "bed", "shower")):
        loop.world.register_object(object_id=f"{otype}_{k}", object_type=otype, position=(2 + 3 * k, 2 + (k % 3) * 4))
    for i in range(8):
        loop.world.lifecycle_service.spawn_agent(f"agent_{i}", (1 + 2 * i, 10), needs={"hunger": 0.9, "hygiene": 0.8, "energy": 0.7}, wallet=3.0)
    for _ in range(5):
        loop.step()
    t0 = time.perf_counter()
    ticks = 100
    for _ in range(ticks):
        loop.step()
    dt = time.perf_counter() - t0
    out["populated_tick_loop"] = {"agents": len(loop.world.agents), "ticks": ticks, "ms_per_tick": 1e3 * dt / ticks,
                                  "agent_steps_per_s": len(loop.world.agents) * ticks / dt, "telemetry": "stub"}
    print("populated loop:", out["populated_tick_loop"], flush=True)
    (ROOT / "profiles" / f"{TAG}_reference_python_baseline.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
