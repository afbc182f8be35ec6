#!/usr/bin/env python
"""Turn the raw ncu reports / bench lines in gpurun_out/ into the tracked summaries under profiles/.

    python scripts/make_profile_summary.py r1
"""
import csv
import io
import json
import shutil
import subprocess
import sys
from pathlib import Path

TAG = sys.argv[1] if len(sys.argv) > 1 else "r1"
ROOT = Path(__file__).resolve().parents[1]
OUT = ROOT / "gpurun_out"
PROF = ROOT / "profiles"
PROF.mkdir(exist_ok=True)
UNITS = {"c2": 32 * 8192, "c3-compact": 4 * 163840, "c3-full": 4 * 163840, "c4": 4 * 262144}
ALGO = {"c2": 2712, "c3-compact": 1756, "c3-full": 3680, "c4": 2712}
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}

traffic = {}
table = []
for w, units in UNITS.items():
    rep = OUT / f"{TAG}_prof_{w}.ncu-rep"
    if not rep.exists():
        continue
    out = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, unit_row, r = rows[0], rows[1], rows[2]

    def val(key, scaled=False):
        i = hdr.index(key)
        v = float(r[i].replace(",", ""))
        return v * SCALE.get(unit_row[i], 1.0) if scaled else v

    rd, wr = val("dram__bytes_read.sum", True), val("dram__bytes_write.sum", True)
    dur = val("gpu__time_duration.sum", True)
    traffic[w] = {
        "dram_bytes_per_agent_step": (rd + wr) / units,
        "algorithmic_bytes_per_agent_step": ALGO[w],
        "ncu_launch_agent_steps": units,
    }
    kernel = r[hdr.index("Kernel Name")].split("(")[0].replace("void <unnamed>::", "")
    table.append(
        (w, kernel, units, dur * 1e6, rd / 1e6, wr / 1e6, (rd + wr) / units, ALGO[w], (rd + wr) / dur / 1e9,
         val("smsp__inst_executed.sum") / units, val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
         val("sm__warps_active.avg.per_cycle_active"), int(val("launch__registers_per_thread")),
         val("launch__shared_mem_per_block_dynamic"))
    )
(PROF / "traffic.json").write_text(json.dumps(traffic, indent=1))

lines = [f"# Round {TAG[1:]} — ncu `--set full` captures of the tick kernel (one launch each, `--clock-control none`)", ""]
lines.append(f"Captured with `scripts/collect_profiles.sh {TAG}` on a B200; the raw `.ncu-rep` files stay in `gpurun_out/` (scratch).")
lines.append("ncu durations are cold-cache and serialised; the steady-state numbers are the bench lines in this directory.")
lines.append("")
lines.append("| workload | kernel | agent-steps in launch | duration µs | DRAM read MB | DRAM write MB | DRAM B / agent-step | algorithmic B / agent-step | DRAM GB/s in capture | warp instr / agent-step | issue active % | warps / SM | regs | dyn smem KB |")
lines.append("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for row in table:
    lines.append("| %s | `%s` | %d | %.1f | %.1f | %.1f | %.0f | %d | %.0f | %.0f | %.1f | %.1f | %d | %.1f |" % row)
lines.append("")
lines.append("## Bench lines (`python bench.py --gpus 1 --steps 20 --warmup 5`: the headline workload and, under `workloads`, the others)")
lines.append("")
lines.append("| workload | agent-steps/s | roofline frac | ms per tick | e2e agent-steps/s (host buffers) | CPU oracle agent-steps/s (cores) | SM MHz |")
lines.append("|---|---|---|---|---|---|---|")
src = OUT / f"{TAG}_bench_c2.json"
if src.exists():
    shutil.copy(src, PROF / src.name)
    d = json.loads(src.read_text().strip().splitlines()[-1])
    cpu = d.get("cpu_baseline") or {}
    lines.append(
        "| c2 (headline) | %.4g | %.3f | %.5f | %.4g | %s | %s |"
        % (d["value"], d["roofline"]["frac"], d["ms_per_step"], d["e2e"]["value"],
           ("%.4g (%d)" % (cpu["value"], cpu["cores"])) if cpu else "-", d["clocks"]["sm_mhz"])
    )
    for w, v in (d.get("workloads") or {}).items():
        lines.append("| %s | %.4g | %.3f | %.5f | - | - | %s |" % (w, v["value"], v["roofline_frac"], v["ms_per_tick"], v.get("sm_mhz")))
ref = OUT / f"{TAG}_bench_reference_c2.json"
if ref.exists():
    shutil.copy(ref, PROF / ref.name)
launches = OUT / f"{TAG}_launches_c2.csv"
if launches.exists():
    shutil.copy(launches, PROF / launches.name)
    rows = [r for r in csv.reader(launches.open()) if len(r) > 10 and r[-3] == "gpu__time_duration.sum"]
    total = sum(float(r[-1].replace(",", "")) for r in rows)
    ours = sum(float(r[-1].replace(",", "")) for r in rows if "tick_kernel" in r[4])
    lines += ["", f"## Launch list (`{launches.name}`): {len(rows)} launches, tick kernels are {100 * ours / max(total, 1):.1f} % of the summed device time",
              "(the rest: torch fills / copies of the bench setup, outside the timed region)."]
for w in ("c2", "c4"):
    rep = OUT / f"{TAG}_prof_{w}.ncu-rep"
    if not rep.exists():
        continue
    txt = subprocess.run([sys.executable, str(ROOT / "scripts" / "ncu_lines.py"), str(rep), "--units", str(UNITS[w]), "--top", "25"],
                         capture_output=True, text=True).stdout
    lines += ["", f"## {w}: top source lines by warp instructions per agent-step", "```"] + [l[:150] for l in txt.splitlines()] + ["```"]
(PROF / f"{TAG}_ncu_summary.md").write_text("\n".join(lines) + "\n")
print("\n".join(lines[:24]))
