#!/usr/bin/env python
"""Time the REFERENCE's own Python implementation of the hot path — TEST / MEASUREMENT INFRASTRUCTURE ONLY.

Runs the unmodified reference from ``oracle/_ref`` (``oracle/make_ref.sh``; ``TOWNLET_REFERENCE`` overrides) on
this machine's host cores and prints ONE JSON object (SURVEY.md §8d "CPU baseline, timed in the same run"):

1. ``hot_path``: one process per host core, each owning one seeded synthetic town of the benchmark's generator,
   timing ``WorldObservationService.build_batch`` (world/observations/service.py:201) +
   ``WorldState._apply_need_decay`` (world/grid.py:1439) + ``RewardEngine.compute`` (rewards/engine.py:34);
   per-core and summed agent-steps/s;
2. ``benchmark_tick``: ``scripts/benchmark_tick.py configs/examples/poc_hybrid.yaml --ticks 100`` exactly as
   shipped (0 agents, ticks/s only);
3. ``populated_loop``: the same ``SimulationLoop`` with 8 agents spawned through
   ``world.lifecycle_service.spawn_agent`` (the script has no seeding hook, scripts/benchmark_tick.py:20-29),
   telemetry stubbed.

``bench.py`` runs this file in a subprocess for its ``cpu_baseline.python_reference`` record; never imported by
the product package.

    python oracle/reference_timing.py [--variant hybrid --grid 48 --agents 8 --iters 50 --workers N]
"""
from __future__ import annotations

import argparse
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
if "TOWNLET_REFERENCE" not in os.environ:
    os.environ["TOWNLET_REFERENCE"] = str(ROOT / "oracle" / "_ref")
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))


def cpu_model() -> str:
    try:
        for line in Path("/proc/cpuinfo").read_text().splitlines():
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return platform.processor() or platform.machine()


def hot_path_rate(variant: str, grid: int, agents: int, iters: int, town_index: int) -> dict:
    """build_batch -> _apply_need_decay -> RewardEngine.compute on one seeded town, ``iters`` ticks."""
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
    return {"agent_steps_per_s": agents * iters / total, "ms_per_tick": {k: 1e3 * v / iters for k, v in parts.items()}}


def _worker(args):
    return hot_path_rate(*args)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", default="hybrid")
    ap.add_argument("--grid", type=int, default=48)
    ap.add_argument("--agents", type=int, default=8)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--workers", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--loop-ticks", type=int, default=100)
    args = ap.parse_args()
    ref = Path(os.environ["TOWNLET_REFERENCE"])
    if not (ref / "src" / "townlet").is_dir():
        print(json.dumps({"unavailable": f"no reference tree at {ref} (run oracle/make_ref.sh where /root/reference exists)"}))
        return
    import make_golden as mg  # noqa: F401  (imports the reference once; the workers fork after this)

    out: dict = {"cpu": cpu_model(), "cores": os.cpu_count(), "python": platform.python_version(),
                 "reference_revision": (ref / "REVISION").read_text().strip() if (ref / "REVISION").exists() else "unknown"}
    jobs = [(args.variant, args.grid, args.agents, args.iters, i) for i in range(args.workers)]
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(args.workers) as pool:
        results = pool.map(_worker, jobs)
    rates = [r["agent_steps_per_s"] for r in results]
    out["hot_path"] = {
        "what": "build_batch + _apply_need_decay + RewardEngine.compute, one seeded town per worker process",
        "variant": args.variant, "grid": args.grid, "agents_per_town": args.agents, "ticks": args.iters, "workers": args.workers,
        "per_core_agent_steps_per_s": sum(rates) / len(rates),
        "per_core_min_max": [min(rates), max(rates)],
        "summed_agent_steps_per_s": sum(rates),
        "ms_per_tick_worker0": results[0]["ms_per_tick"],
        "wall_s": time.perf_counter() - t0,
    }

    # the reference tick loop exactly as shipped (0 agents -> ticks/s only)
    cwd = Path.cwd()  # make_golden chdir'ed into its writable copy of the reference
    proc = subprocess.run(
        [sys.executable, "scripts/benchmark_tick.py", "configs/examples/poc_hybrid.yaml", "--ticks", str(args.loop_ticks)],
        cwd=cwd, env={**os.environ, "PYTHONPATH": str(cwd / "src")}, capture_output=True, text=True,
    )
    shipped: dict = {"command": f"scripts/benchmark_tick.py configs/examples/poc_hybrid.yaml --ticks {args.loop_ticks}",
                     "returncode": proc.returncode, "agents": 0}
    for line in proc.stdout.splitlines():  # the script prints `average_tick_seconds=<float>` after the telemetry stream
        if line.startswith("average_tick_seconds="):
            shipped["average_tick_seconds"] = float(line.split("=", 1)[1])
            shipped["ticks_per_second"] = 1.0 / shipped["average_tick_seconds"] if shipped["average_tick_seconds"] > 0 else None
    if proc.returncode != 0:
        shipped["stderr_tail"] = proc.stderr[-300:]
    out["benchmark_tick"] = shipped

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
    for _ in range(args.loop_ticks):
        loop.step()
    dt = time.perf_counter() - t0
    out["populated_loop"] = {"what": "SimulationLoop.step with 8 spawned agents, telemetry stub", "agents": len(loop.world.agents),
                             "ticks": args.loop_ticks, "ms_per_tick": 1e3 * dt / args.loop_ticks,
                             "agent_steps_per_s": len(loop.world.agents) * args.loop_ticks / dt}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
