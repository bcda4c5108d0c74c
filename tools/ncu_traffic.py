#!/usr/bin/env python
"""profiles/traffic.json from ncu --set full reports: DRAM bytes (read + write) per launch of each stage's kernel.
Usage: ncu_traffic.py report1.ncu-rep [report2.ncu-rep ...]   (later reports override earlier ones)"""
import csv, io, json, os, subprocess, sys

STAGE_OF = [("knn_bound_kernel", "knn"), ("knn_collect_kernel", "knn"), ("knn_finalize_kernel", "knn"), ("knn_slow_kernel", "knn"),
            ("sort_kernel", "sort"), ("conv_in_kernel", "conv_in"), ("proxy_block_kernel", "proxy_block"),
            ("tc_gemm_bres_kernel<__nv_bfloat16, 256", "conv5"), ("tc_gemm_bres_kernel<float, 256, 2", "colmax"), ("tc_gemm_bres_kernel<__half, 256, 2", "colmax"),
            ("tc_gemm_bres_kernel<__nv_bfloat16, 64", "assign_gemm"),
            ("assign_vlad_kernel", "assign_vlad"), ("assign_vlad_fp8_kernel", "assign_vlad"), ("tc_gemm_kernel<__nv_bfloat16, 64", "vlad_gemm"), ("tc_gemm_kernel<float, 256", "hidden_gemm"),
            ("vlad_residual_kernel", "vlad_finalize"), ("retr_score_kernel<1>", "retrieve_emit"), ("retr_score_kernel<0>", "retrieve_sample"),
            ("select_kernel", "retrieve_select"), ("rerank_kernel", "retrieve_rerank"), ("sample_threshold_kernel", "retrieve_threshold")]
MULTI = {"knn"}      # stages made of several different kernels per call: per-kernel averages are summed
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
                key = (st, pat if st in MULTI else "")
                a = acc.setdefault(key, [0.0, 0, r[ik].split("(")[0], 0.0, 0.0])
                a[0] += b
                a[1] += 1
                a[3] += float(r[ii].replace(",", ""))
                a[4] += float(r[it].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(units[it], 1.0)
                break
    stages = {}
    for (st, pat), (b, n, name, inst, us) in acc.items():
        e = stages.setdefault(st, {"dram_bytes_per_launch": 0.0, "warp_instructions_per_launch": 0.0, "ncu_us_per_launch": 0.0,
                                   "launches_profiled": 0, "kernel": []})
        e["dram_bytes_per_launch"] += b / n
        e["warp_instructions_per_launch"] += inst / n
        e["ncu_us_per_launch"] += us / n
        e["launches_profiled"] += n
        e["kernel"].append(name)
    for st, e in stages.items():
        e["kernel"] = ", ".join(sorted(set(e["kernel"])))
        e["report"] = os.path.basename(rep)
        e["note"] = ("ncu replays each launch with cold caches; embedding kernels: 128 clouds per call; a stage made of several kernels "
                     "(kNN) is the sum of its kernels' per-launch averages")
        out[st] = e
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
