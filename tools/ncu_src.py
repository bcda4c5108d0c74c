#!/usr/bin/env python
"""Summarise `ncu --page source --csv` for one kernel: executed warp instructions and stall samples by opcode,
plus the hottest instructions.  Usage: ncu_src.py report.ncu-rep kernel_regex [top]"""
import collections
import csv
import io
import re
import subprocess
import sys


def main():
    rep, kre = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    # the report may contain several launches; use the first
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    end = hdr[1] - 1 if len(hdr) > 1 else len(rows)
    H = rows[hdr[0]]
    ix = {h: i for i, h in enumerate(H)}
    ops, samples, tot, stot = collections.Counter(), collections.Counter(), 0, 0
    hot = []
    for r in rows[hdr[0] + 1:end]:
        if len(r) < len(H) - 2:
            continue
        try:
            n = int(r[ix["Instructions Executed"]])
            s = int(r[ix["# Samples"]])
        except ValueError:
            continue
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ix["Source"]])
        if not m:
            continue
        op = m.group(1)
        ops[op] += n
        samples[op] += s
        tot += n
        stot += s
        hot.append((s, n, r[ix["Source"]].strip()))
    print("kernel %s: %d warp instructions, %d stall samples" % (rows[hdr[0] - 1][1][:80] if hdr[0] else kre, tot, stot))
    for o, n in ops.most_common(top):
        print("  %-12s %12d %5.1f%%   samples %5.1f%%" % (o, n, 100.0 * n / tot, 100.0 * samples[o] / max(stot, 1)))
    print("hottest instructions by stall samples:")
    for s, n, src in sorted(hot, reverse=True)[:top]:
        print("  %6d samples %10d exec  %s" % (s, n, src[:100]))


if __name__ == "__main__":
    main()
