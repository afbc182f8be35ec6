#!/usr/bin/env python
"""Summarise an ncu report by source line: instructions executed and stall samples.

    python scripts/ncu_lines.py gpurun_out/prof.ncu-rep [--units N] [--top 40]

``--units`` divides the counts (e.g. agent-steps in the profiled launch) so the
table reads as warp instructions per unit.
"""
import argparse
import csv
import io
import subprocess
import sys


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--units", type=float, default=1.0)
    ap.add_argument("--top", type=int, default=40)
    args = ap.parse_args()
    out = subprocess.run(
        ["ncu", "-i", args.report, "--page", "source", "--csv", "--print-source", "cuda,sass"],
        capture_output=True, text=True, check=True,
    ).stdout
    rows = list(csv.reader(io.StringIO(out)))
    sections = []  # (file, header, rows)
    cur_file = None
    i = 0
    kernel_idx = -1
    while i < len(rows):
        r = rows[i]
        if r and r[0] in ("File Path", "File Name"):
            cur_file = r[1]
        if r and r[0] == "Line No":
            hdr = r
            body = []
            i += 1
            while i < len(rows) and rows[i] and rows[i][0] not in ("Function Name", "File Path", "File Name", "Line No"):
                body.append(rows[i])
                i += 1
            sections.append((kernel_idx, cur_file, hdr, body))
            continue
        i += 1
    lines = []
    total_inst = 0
    total_samples = 0
    for kidx, fname, hdr, body in sections:
        ia = hdr.index("Instructions Executed")
        isamp = hdr.index("# Samples")
        for r in body:
            if r[0] == "":
                continue  # SASS row
            inst = int(r[ia]) if r[ia].isdigit() else 0
            samp = int(r[isamp]) if r[isamp].isdigit() else 0
            total_inst += inst
            total_samples += samp
            lines.append((inst, samp, fname.split("/")[-1], r[0], r[1].strip()))
    lines.sort(reverse=True)
    print(f"total warp instructions {total_inst} ({total_inst / args.units:.1f} per unit), samples {total_samples}")
    print(f"{'inst/unit':>10} {'inst%':>6} {'smp%':>6}  file:line  source")
    for inst, samp, fname, ln, src in lines[: args.top]:
        print(f"{inst / args.units:10.2f} {100 * inst / max(total_inst, 1):6.1f} {100 * samp / max(total_samples, 1):6.1f}  {fname}:{ln}  {src[:110]}")


if __name__ == "__main__":
    main()
