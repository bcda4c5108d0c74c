#!/usr/bin/env python
"""profiles/traffic.json from ncu --set full reports: DRAM bytes (read + write) per launch of each stage's kernel.
Usage: ncu_traffic.py report1.ncu-rep [report2.ncu-rep ...]   (later reports override earlier ones)"""
import csv, io, json, os, subprocess, sys

STAGE_OF = [("knn_kernel", "knn"), ("sort_kernel", "sort"), ("conv_in_kernel", "conv_in"), ("proxy_block_kernel", "proxy_block"),
            ("tc_gemm_bres_kernel<__nv_bfloat16, 256", "conv5"), ("tc_gemm_bres_kernel<__nv_bfloat16, 64", "assign_gemm"),
            ("tc_gemm_kernel<__nv_bfloat16, 64", "vlad_gemm"), ("tc_gemm_kernel<float, 256", "hidden_gemm"),
            ("vlad_residual_kernel", "vlad_finalize")]
out = {}
for rep in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h, units = rows[0], rows[1]
    ik, ir, iw, it = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("gpu__time_duration.sum")
    ii = h.index("smsp__inst_executed.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    acc = {}
    for r in rows[2:]:
        for pat, st in STAGE_OF:
            if pat in r[ik]:
                b = float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
                a = acc.setdefault(st, [0.0, 0, r[ik].split("(")[0], 0.0])
                a[0] += b
                a[1] += 1
                a[3] += float(r[ii].replace(",", ""))
                break
    for st, (b, n, name, inst) in acc.items():
        out[st] = {"dram_bytes_per_launch": b / n, "warp_instructions_per_launch": inst / n, "launches_profiled": n, "kernel": name,
                   "report": os.path.basename(rep),
                   "note": "ncu replays each launch with cold caches; batch = the bench's 128 clouds per call"}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
