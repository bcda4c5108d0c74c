#!/usr/bin/env python
"""Per-source-line totals (executed warp instructions, stall samples) from an ncu report (needs -lineinfo).
Usage: ncu_lines.py report.ncu-rep kernel_regex [top]"""
import csv
import io
import subprocess
import sys

rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
out, fpath, seen_fn, seen_lines = [], "", 0, set()
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fpath = r[1].split("/")[-1]
    if len(r) >= 2 and r[0] == "Function Name":
        seen_fn += 1
    if len(r) > 8 and r[0].isdigit():
        try:
            if (fpath, int(r[0])) in seen_lines:
                continue
            seen_lines.add((fpath, int(r[0])))
            out.append((int(r[7]), int(r[6]), fpath, int(r[0]), r[1].strip()))
        except ValueError:
            pass
tot = sum(o[0] for o in out) or 1
stot = sum(o[1] for o in out) or 1
print("total warp instructions %d, samples %d (first launch matching; %d function sections)" % (tot, stot, seen_fn))
key = (lambda o: o[1]) if (len(sys.argv) > 4 and sys.argv[4] == "samples") else (lambda o: o[0])
for n, s, f, ln, src in sorted(out, key=key, reverse=True)[:top]:
    print("%5.1f%% instr %5.1f%% samples  %s:%d  %s" % (100.0 * n / tot, 100.0 * s / stot, f, ln, src[:90]))
