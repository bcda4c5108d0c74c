#!/usr/bin/env python
"""bench.py -- EPC-Net clouds/sec (4096 points) on N B200s, batch-sharded (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one pass of the embedding hot path (kNN graph -> ProxyConv backbone -> conv5 -> G_VLAD -> 256-d
descriptor) over one batch of `--clouds` synthetic 4096-point clouds per GPU (weak scaling, no collective on the
data path).  Rank 0 prints ONE JSON line:
  value            clouds/s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e              the same metric through evaluate.get_latent_vectors with HOST buffers (pinned H2D + D2H timed)
  roofline         the dominant kernel's algorithmic FLOP/s (or B/s) vs MEASURED_PEAKS.json, timed live with events
  cpu_baseline     the oracle (dense-as-written numpy restatement of the TF graph) on the host cores, bounded sample
`--impl reference` times that CPU restatement alone (TF 1.12 itself cannot run here).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _data  # noqa: E402

METRIC = "EPC-Net clouds/sec (4096 pts)"
UNIT = "clouds/s"
ARCH = "epc-net"
N_POINTS = 4096
# Algorithmic work per cloud (N=4096, k=20): FLOPs from SURVEY.md Appendix C; BYTES = compulsory HBM traffic of the stage as
# the data is laid out here: inputs read once + outputs written once, weights amortised over the batch.
MiB = 1024.0 * 1024.0
STAGE_MODEL = {
    # stage: (FLOP/cloud, HBM bytes/cloud, pipe that does the FLOPs)
    "sort": (0.0, 48 * 1024 + 64 * 1024 + 16 * 1024 + 4 * 1024, "alu"),
    "knn": (8.0 * N_POINTS * N_POINTS, 64 * 1024 + 16 * 1024 + 4 * 1024 + 160 * 1024 + 32 * 1024, "alu"),   # dense-equivalent pairs
    "conv_in": (2.0 * N_POINTS * 3 * 64, 64 * 1024 + 0.5 * MiB, "alu"),
    # 4 launches per cloud: x (fp16) in, neighbour lists + counts in, concat slice (bf16) out, next x (fp16) out
    "proxy_block": (4 * (3 * 2.0 * N_POINTS * 64 * 64 + N_POINTS * 20 * 64), 4 * (0.5 * MiB + 176 * 1024 + 0.5 * MiB) + 3 * 0.5 * MiB, "tensor"),
    "conv5": (2.0 * N_POINTS * 256 * 1024, 2 * MiB + 4 * MiB, "tensor"),           # concat16 in, H' (fp8 e4m3) out
    "assign_gemm": (2.0 * N_POINTS * 1024 * 64, 8 * MiB + 0.5 * MiB, "tensor"),    # H in, S' out
    "vlad_gemm": (2.0 * 64 * N_POINTS * 1024, 8 * MiB + 0.5 * MiB + 0.5 * MiB, "tensor"),
    # assignment + VLAD in one launch on the fp8 H': read by both roles (2 x 4 MiB), S'' (fp8, 128 B rows) written, V slabs written
    "assign_vlad": (2 * 2.0 * N_POINTS * 1024 * 64, 2 * 4 * MiB + 0.5 * MiB + 0.5 * MiB, "tensor"),
    "vlad_finalize": (0.0, 3 * 0.25 * MiB + 2 * 0.25 * MiB, "alu"),
    "hidden_gemm": (2.0 * 4 * 16384 * 256, 0.25 * MiB + 16.8e6 / 128.0, "tensor"),   # 16.8 MB of weights per 128-cloud call
}
FP32_ALU_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12     # non-tensor FFMA peak of a B200 at its 1965 MHz boost clock (74.4)


def stage_roofline(stage, ms, clouds, peaks):
    """Achieved algorithmic rate of one stage against the two rooflines that can bound it; `frac` is the larger."""
    flop, byt, pipe = STAGE_MODEL[stage]
    sec = ms * 1e-3
    hbm = byt * clouds / sec / 1e9
    out = {"hbm_gbs": hbm, "hbm_frac": hbm / peaks["hbm_gbs"]}
    if flop:
        tf = flop * clouds / sec / 1e12
        peak = peaks["bf16_tflops_sustained"] if pipe == "tensor" else peaks.get("fp32_tflops", FP32_ALU_TFLOPS)
        out.update({"tflops": tf, "flop_peak": peak, "flop_frac": tf / peak, "pipe": pipe})
    return out


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--clouds", type=int, default=8192, help="clouds per GPU per step (a multiple of --batch)")
    ap.add_argument("--batch", type=int, default=512, help="clouds per embed() call (split over --streams library calls)")
    ap.add_argument("--chunk", type=int, default=256, help="clouds per library call")
    ap.add_argument("--e2e-clouds", type=int, default=4096, help="clouds per step of the end-to-end (host buffer) measurement")
    ap.add_argument("--no-epc-net-l", action="store_true", help="skip the EPC-Net-L extra key")
    ap.add_argument("--streams", type=int, default=2, help="CUDA streams the calls of one step alternate over")
    ap.add_argument("--cpu-sample", type=int, default=24, help="clouds in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-retrieval", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the post-run oracle check of the timed outputs")
    ap.add_argument("--arch", default=ARCH, choices=["epc-net", "epc-net-l"],
                    help="epc-net = BASELINE configs[1] (the headline); epc-net-l = configs[2] (lightweight variant)")
    return ap.parse_args()


def make_clouds(n, seed):
    rng = np.random.default_rng(seed)
    return rng.uniform(-1.0, 1.0, (n, N_POINTS, 3)).astype(np.float32)       # SURVEY.md 8d C1/C2 inputs


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "bf16_tflops": d.get("bf16_tflops", 1590.0),
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", 1400.0), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def ncu_warp_instructions(stage):
    """smsp__inst_executed.sum per launch of the stage's kernel at 128 clouds per call, from the same committed ncu capture."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p)).get(stage, {}).get("warp_instructions_per_launch")
    except Exception:
        return None


def ncu_traffic(stage):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the stage's kernel, from the committed ncu capture
    (profiles/traffic.json, written by tools/ncu_traffic.py); None when no capture is on file."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    try:
        return json.load(open(p)).get(stage, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


class ClockSampler(object):
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline(n_sample, V, params, arch=ARCH):
    """The reference's CPU path: dense-as-written restatement (oracle), ONE cloud per call like evaluate.py:355."""
    from oracle import epc_oracle
    clouds = make_clouds(n_sample, 4242)
    epc_oracle.forward(arch, clouds[:1][None], V, params)          # warm-up (BLAS thread pools)
    ts = []
    for i in range(n_sample):
        t0 = time.perf_counter()
        epc_oracle.forward(arch, clouds[i:i + 1][None], V, params)
        ts.append(time.perf_counter() - t0)
    return n_sample / float(np.sum(ts)), float(np.median(ts))


def parity_check(got, clouds, rows, V, arch):
    """max-abs / min cosine of the given descriptor rows against the oracle's forward on the same clouds."""
    from oracle import epc_oracle
    params = _data.default_params(arch)
    worst_abs, worst_cos = 0.0, 1.0
    for r in rows:
        ref = epc_oracle.forward(arch, clouds[r:r + 1][None], V, params).reshape(-1)
        g = got[r]
        worst_abs = max(worst_abs, float(np.abs(g - ref).max()))
        worst_cos = min(worst_cos, float((g * ref).sum() / max(np.linalg.norm(g) * np.linalg.norm(ref), 1e-30)))
    return worst_abs, worst_cos


def run_reference(args):
    """--impl reference: the CPU restatement, every host thread the BLAS takes, same metric/config keys."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    variables = importlib.import_module("epc-net_b200.variables")
    from oracle import epc_oracle
    V = variables.synthetic_variables(ARCH, 1)
    params = _data.default_params(ARCH)
    per_step = 2                                                    # bounded sample: 2 clouds per step
    clouds = make_clouds(per_step * (args.steps + args.warmup), 4242)
    k = 0
    for _ in range(args.warmup):
        for j in range(per_step):
            epc_oracle.forward(ARCH, clouds[k:k + 1][None], V, params)
            k += 1
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for j in range(per_step):
            epc_oracle.forward(ARCH, clouds[k:k + 1][None], V, params)
            k += 1
    dt = time.perf_counter() - t0
    val = per_step * args.steps / dt
    cores = os.cpu_count()
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "EPC-Net (configs/epc-net.yaml) embedding of synthetic 4096-point clouds, 1 cloud per call "
                               "(evaluate.py:355), seeded random-init weights", "clouds_per_step": per_step},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d clouds/step x %d steps, numpy/BLAS dense-as-written restatement of the TF-1.12 "
                                   "graph (TensorFlow not runnable here)" % (per_step, args.steps)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def measure_arch(arch, args, mods, rank, world, local, dist, full):
    """Device-resident throughput, end-to-end throughput, per-stage table and parity of one architecture.
    full=False (the EPC-Net-L extra key): fewer steps, no clock sampling."""
    import torch
    lib_mod, variables, engine_mod, evaluate, models = mods
    V = variables.synthetic_variables(arch, 1)
    store = variables.VariableStore(V)
    params = dict(_data.default_params(arch), EMBED_CHUNK=args.chunk, EMBED_STREAMS=args.streams, VARIABLES=store)
    eng = engine_mod.get_engine(arch, params, store=store)
    eng_serial = engine_mod.get_engine(arch, dict(params, EMBED_STREAMS=1), store=store)      # per-stage timing pass only
    K, W = (args.steps, args.warmup) if full else (min(args.steps, 6), 3)
    CB = args.batch                                                # clouds per embed() call (args.chunk per library call)
    calls = max(1, (args.clouds + CB - 1) // CB)
    B = calls * CB                                                 # clouds per GPU per step
    nbatch = 4                                                     # distinct batches the calls rotate over
    host_batches = [make_clouds(CB, 1000 + 97 * rank + i) for i in range(nbatch)]
    dev_batches = [torch.from_numpy(b).cuda() for b in host_batches]
    outs = [torch.empty((CB, 256), dtype=torch.float32, device="cuda") for _ in range(2)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i, e):
        for j in range(calls):
            c = i * calls + j
            e.embed(dev_batches[c % nbatch], out=outs[c & 1])
        return (i * calls + calls - 1)                             # index of the step's last call

    # ---- device-resident throughput -------------------------------------------------------------------
    for i in range(W):
        step(i, eng)
    barrier()
    sampler = ClockSampler(local) if (full and rank == 0) else None
    if sampler:
        sampler.start()
    lib_mod.launch_count_reset()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(K):
        last_call = step(W + i, eng)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = lib_mod.launch_count()
    last_outs = {c: outs[c & 1].cpu().numpy() for c in (last_call - 1, last_call)} if calls > 1 else {last_call: outs[last_call & 1].cpu().numpy()}
    # a few steps once more with the per-stage event brackets on, on ONE stream (brackets of concurrent streams would time
    # each other's kernels); they stay out of `value`
    Kp = min(K, 3)
    lib_mod.profile_reset()
    lib_mod.profile_enable(True)
    for i in range(Kp):
        step(W + i, eng_serial)
    torch.cuda.synchronize()
    stages = lib_mod.profile_read()
    lib_mod.profile_enable(False)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * B * K / (ms * 1e-3)

    # ---- end to end: host arrays in, host arrays out, through the reference-facing call ---------------------
    # One get_latent_vectors call over the K steps' clouds, the way evaluate.py:283-293 calls it on a whole database set:
    # every 128-cloud call is staged through pinned memory, copied H2D, embedded, and the descriptors copied back -- all
    # inside the timed region; the engine overlaps call i+1's copy with call i's compute.
    ops = {"MODEL": models.load(arch), "params": params}
    Be = min(B, args.e2e_clouds)                                   # clouds per e2e step (bounds the host array: K x Be x 48 KB)
    reps = (K * Be + nbatch * CB - 1) // (nbatch * CB)
    big = np.concatenate(host_batches * reps, 0)[:K * Be]
    names = {i: {} for i in range(len(big))}
    warm = np.concatenate(host_batches[:2], 0)
    evaluate.get_latent_vectors(None, ops, {i: {} for i in range(len(warm))}, warm)
    evaluate.get_latent_vectors(None, ops, names, big)                  # sizes the staging buffers for the timed call
    barrier()
    t0 = time.perf_counter()
    desc = evaluate.get_latent_vectors(None, ops, names, big)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = world * Be * K / e2e_s
    clocks = sampler.stop() if sampler else None          # sampled across both timed regions
    assert desc.shape == (Be * K, 256) and np.isfinite(desc).all()
    # parity of what was just timed: first/last descriptor of each library call of the last device-resident calls and of the
    # end-to-end result against the CPU oracle (the checker, never the thing measured)
    parity = None
    if rank == 0 and not args.no_parity:
        worst_abs, worst_cos, n_rows = 0.0, 1.0, 0
        for c, got in last_outs.items():
            rows = sorted({0, min(args.chunk, CB) - 1, min(args.chunk, CB - 1), CB - 1}) if full else [0, CB - 1]
            a, b = parity_check(got, host_batches[c % nbatch], rows, V, arch)
            worst_abs, worst_cos, n_rows = max(worst_abs, a), min(worst_cos, b), n_rows + len(rows)
        rows2 = sorted({0, CB - 1, len(big) - CB, len(big) - 1}) if full else [0, len(big) - 1]
        a, b = parity_check(desc, big, rows2, V, arch)
        worst_abs, worst_cos, n_rows = max(worst_abs, a), min(worst_cos, b), n_rows + len(rows2)
        parity = {"parity_checked": True, "max_abs": worst_abs, "min_cos": worst_cos, "rows_checked": n_rows,
                  "tolerance": {"max_abs": 1e-3, "min_cos": 0.9999},
                  "against": "oracle/epc_oracle.forward (dense-as-written restatement of models/%s.py)" % arch}
        assert worst_abs <= 1e-3 and worst_cos >= 0.9999, parity
    return {"arch": arch, "V": V, "value": value, "ms": ms, "K": K, "W": W, "B": B, "CB": CB, "calls": calls, "nbatch": nbatch,
            "launches": launches, "stages": stages, "stage_clouds": B * Kp, "e2e": e2e, "e2e_s": e2e_s, "Be": Be,
            "clocks": clocks, "parity": parity}


def ffma_peak(lib_mod, torch):
    """The FP32 pipe's peak measured now, on this GPU, at the clocks of the moment (epc_microbench_ffma)."""
    import ctypes
    lib = lib_mod.load()
    fl = ctypes.c_double(0.0)
    st = torch.cuda.current_stream().cuda_stream
    lib_mod.check(lib.epc_microbench_ffma(4096, ctypes.byref(fl), st))
    best = 0.0
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib_mod.check(lib.epc_microbench_ffma(65536, ctypes.byref(fl), st))
        e1.record()
        torch.cuda.synchronize()
        best = max(best, fl.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


def stage_table_of(stages, clouds, peaks):
    total = sum(v[0] for v in stages.values()) or 1.0
    table = {}
    for k, v in stages.items():
        e = {"us_per_cloud": v[0] / clouds * 1e3, "share": v[0] / total, "launches": v[1]}
        if k in STAGE_MODEL and v[0] > 0:
            e.update(stage_roofline(k, v[0], clouds, peaks))
        table[k] = e
    return table, total


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    lib_mod = importlib.import_module("epc-net_b200._lib")
    variables = importlib.import_module("epc-net_b200.variables")
    engine_mod = importlib.import_module("epc-net_b200.engine")
    evaluate = importlib.import_module("epc-net_b200.evaluate")
    models = importlib.import_module("epc-net_b200.models")
    mods = (lib_mod, variables, engine_mod, evaluate, models)

    arch = args.arch
    r = measure_arch(arch, args, mods, rank, world, local, dist, full=True)
    other = None
    if arch == "epc-net" and not args.no_epc_net_l:
        # BASELINE.json configs[2]: the lightweight variant at its own batch shape (256 clouds per library call)
        largs = argparse.Namespace(**dict(vars(args), chunk=256, batch=512))
        other = measure_arch("epc-net-l", largs, mods, rank, world, local, dist, full=False)
    retr = None
    if not args.no_retrieval:
        retr = bench_retrieval(evaluate, lib_mod, torch, dist if world > 1 else None, rank, world)      # collective when world > 1
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (stage with the largest share of device time) ----------------
    peaks = measured_peaks()
    fp32_peak = ffma_peak(lib_mod, torch)
    peaks["fp32_tflops"] = fp32_peak
    stages, B, K, W = r["stages"], r["B"], r["K"], r["W"]
    stage_table, total_stage_ms = stage_table_of(stages, r["stage_clouds"], peaks)
    top = max(stages.items(), key=lambda kv: kv[1][0])[0]
    top_ms, top_n = stages[top]
    clouds_per_launch = r["stage_clouds"] / float(top_n)
    traffic = ncu_traffic(top)
    roof = {"kernel": top, "share_of_step": top_ms / total_stage_ms, "launches": top_n,
            "avg_launch_ms": top_ms / top_n, "clouds_per_launch": clouds_per_launch, "peak_source": peaks["source"],
            "traffic": traffic * 1.0 if traffic is not None else None,
            "timing": "CUDA events bracketing the stage's kernels on their stream, in a separate single-stream pass of %d steps "
                      "right after the timed region (stage sum there: %.2f us/cloud; timed two-stream pass: %.2f us/cloud)"
                      % (min(K, 3), total_stage_ms / r["stage_clouds"] * 1e3, r["ms"] / (B * K) * 1e3)}
    sr = stage_roofline(top, top_ms, r["stage_clouds"], peaks) if top in STAGE_MODEL else None
    if sr is not None and sr.get("pipe") == "tensor" and sr["flop_frac"] >= sr["hbm_frac"]:
        roof.update({"bound": "tensor", "achieved": sr["tflops"], "peak": sr["flop_peak"], "unit": "TFLOP/s", "frac": sr["flop_frac"],
                     "note": "algorithmic FLOP of the stage / event-timed duration; peak = measured sustained bf16 cuBLAS"})
    elif sr is not None and sr.get("pipe") == "alu" and sr.get("flop_frac", 0.0) >= sr["hbm_frac"]:
        roof.update({"bound": "fp32", "achieved": sr["tflops"], "peak": fp32_peak, "unit": "TFLOP/s", "frac": sr["flop_frac"],
                     "peak_source": "measured in this run: epc_microbench_ffma (register-only FFMA chains on every SM), %.1f TFLOP/s; "
                                    "theoretical 148 SM x 128 lanes x 2 x 1.965 GHz = %.1f" % (fp32_peak, FP32_ALU_TFLOPS),
                     "hbm_frac": sr["hbm_frac"],
                     "note": "ALU-bound: the whole cloud sits in shared memory, HBM traffic is negligible (hbm_frac).  achieved = "
                             "dense-equivalent 8 N^2 FLOP per cloud (SURVEY 8d) / event-timed duration; exact AABB pruning skips most "
                             "pairs, the selection networks are the remaining work",
                     "l1_pipe_note": "ncu (profiles/r2_front_kernels_ncu.txt): l1tex__data_pipe_lsu_wavefronts 66 % (knn_bound) / 72 % "
                                     "(knn_collect) of peak -- one float4 candidate per lane and distance is 4 shared-memory wavefronts "
                                     "per warp step beside ~4 arithmetic instructions, so the kernel sits against the L1 data pipe and "
                                     "the FP32 pipe together (DESIGN.md section 8)"})
        inst = ncu_warp_instructions(top)
        if inst and r["clocks"] and r["clocks"].get("sm_mhz"):
            ach = inst / 128.0 * clouds_per_launch / (top_ms / top_n * 1e-3) / 1e9
            peak = torch.cuda.get_device_properties(local).multi_processor_count * 4 * r["clocks"]["sm_mhz"] * 1e6 / 1e9
            roof.update({"issue_achieved": ach, "issue_peak": peak, "issue_unit": "G warp-instr/s", "issue_frac": ach / peak,
                         "issue_note": "warp instructions of the stage's kernels per 128 clouds from the committed ncu capture "
                                       "(profiles/traffic.json), duration event-timed here"})
    elif sr is not None:
        roof.update({"bound": "hbm", "achieved": sr["hbm_gbs"], "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": sr["hbm_frac"],
                     "algorithmic_bytes_per_launch": STAGE_MODEL[top][1] * clouds_per_launch})

    line = {
        "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": r["ms"] / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "tensor_operands": "fp16 (ProxyConv 64x64 layers), bf16 (conv5), fp8 e4m3 with exact power-of-two scales and stochastic rounding "
                           "(per-point features H for assignment/VLAD), tf32 (hidden FC), fp16 (EPC-Net-L conv5); fp32 accumulation",
        "data": "synthetic",
        "config": {"workload": ("EPC-Net (configs/epc-net.yaml: 4 ProxyConv blocks + G_VLAD, 256-d)" if arch == "epc-net" else
                                "EPC-Net-L (configs/epc-net-l.yaml: 2 ProxyConv blocks + max-pool + FC, 256-d)") +
                               " batch embedding of synthetic uniform(-1,1) 4096-point clouds, seeded random-init weights, "
                               "batch-sharded",
                   "clouds_per_gpu_per_step": B, "clouds_per_call": args.chunk, "calls_per_step": r["calls"] * max(1, r["CB"] // args.chunk),
                   "streams": args.streams, "knn_arith": "muladd", "timed_region_s": r["ms"] * 1e-3,
                   "l2": "the calls of a step rotate over %d distinct %d-cloud batches; every call streams >0.5 GB of intermediates "
                         "(>> 126 MB L2)" % (r["nbatch"], r["CB"])},
        "e2e": {"value": r["e2e"], "unit": UNIT, "h2d_bytes_per_step": r["Be"] * N_POINTS * 3 * 4, "d2h_bytes_per_step": r["Be"] * 256 * 4,
                "clouds_per_step": r["Be"], "timed_region_s": r["e2e_s"],
                "api": "one evaluate.get_latent_vectors(host ndarray of steps x clouds) -> host ndarray call; per-call pinned staging, H2D and D2H inside"},
        "gpu_launches": int(r["launches"]), "clocks": r["clocks"], "roofline": roof, "stages": stage_table,
    }
    if r["parity"] is not None:
        line.update(r["parity"])
    if other is not None:
        ltable, ltotal = stage_table_of(other["stages"], other["stage_clouds"], peaks)
        ltop = max(other["stages"].items(), key=lambda kv: kv[1][0])[0]
        line["epc_net_l"] = {
            "workload": "EPC-Net-L (configs/epc-net-l.yaml; BASELINE configs[2]) large-batch embedding, %d clouds per step, %d per library call"
                        % (other["B"], 256),
            "value": other["value"], "unit": UNIT, "steps": other["K"], "warmup": other["W"], "timed_region_s": other["ms"] * 1e-3,
            "e2e": {"value": other["e2e"], "unit": UNIT, "clouds_per_step": other["Be"]},
            "gpu_launches": int(other["launches"]), "dominant_kernel": ltop, "dominant_share": other["stages"][ltop][0] / ltotal,
            "stages": {k: {"us_per_cloud": v["us_per_cloud"], "share": v["share"]} for k, v in ltable.items() if v["share"] >= 0.02},
            "parity": other["parity"]}
    if not args.no_retrieval:
        line["retrieval"] = retr
    if not args.no_cpu_baseline:
        cps, med = cpu_baseline(args.cpu_sample, r["V"], _data.default_params(arch), arch)
        line["cpu_baseline"] = {"value": cps, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                "sample": "%d clouds, 1 cloud per call (evaluate.py:355), median %.3f s/cloud; numpy/BLAS "
                                          "dense-as-written restatement of the TF-1.12 graph" % (args.cpu_sample, med)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def bench_retrieval(evaluate, lib_mod, torch, dist, rank, world):
    """Secondary metrics of BASELINE.json: retrieval queries/s and recall@1 on SURVEY 8d C5 (D=20k, Q=3k, k=25).
    The database is prepared once (RetrievalIndex = the reference's KDTree(database_output) object, evaluate.py:463) and
    every timed call is one query() of all Q queries.  With N ranks the database rows are sharded N ways (queries
    replicated); every rank finds its local top-25 with global row ids, ONE NCCL all-gather of the packed (distance | index)
    [2,Q,25] buffer per rank, then the (distance, index) merge (epc_merge_topk_strided)."""
    D, Q, k, dim = 20000, 3000, 25, 256
    db, q, src = _data.retrieval_problem(D=D, Q=Q, seed=7)
    qt = torch.from_numpy(q).cuda()
    if dist is None:
        index = evaluate.RetrievalIndex(torch.from_numpy(db).cuda())
    else:
        dmod = importlib.import_module("epc-net_b200.dist")
        index = dmod.ShardedRetrieval(db)

    def timed(qq, reps):
        for _ in range(3):
            d, i = index.query(qq, k)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            d, i = index.query(qq, k)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / reps, i

    ms, i = timed(qt, 50)
    # per-stage device time of the local (per-shard) work, single stream
    lib_mod.profile_reset()
    lib_mod.profile_enable(True)
    for _ in range(5):
        index.query(qt, k)
    torch.cuda.synchronize()
    st = {n: v[0] / 5 for n, v in lib_mod.profile_read().items() if n.startswith("retrieve")}
    lib_mod.profile_enable(False)
    # a large query batch (16 x Q): the per-call latencies (launches, the all-gather) amortise, the sharded work dominates
    QL = 16 * Q
    ql = torch.from_numpy(np.tile(q, (16, 1)) + np.random.default_rng(5).normal(0, 0.01, (QL, dim)).astype(np.float32)).cuda()
    ms_l, _ = timed(ql, 5)
    # world >= 4: also the (2 database shards x world/2 query groups) layout -- the per-query work (selection, re-rank, merge)
    # that pure database sharding repeats on every rank is split as well; same results
    grid2 = None
    if dist is not None and world >= 4 and world % 2 == 0:
        index = dmod.ShardedRetrieval(db, db_shards=2)
        g_ms, g_i = timed(qt, 50)
        g_ms_l, _ = timed(ql, 5)
        grid2 = {"db_shards": 2, "query_groups": world // 2, "queries_per_s": Q / (g_ms * 1e-3), "ms_per_call": g_ms,
                 "large_batch_queries_per_s": QL / (g_ms_l * 1e-3), "same_indices": bool(torch.equal(g_i, i)),
                 "collective": "all_gather of the packed lists inside each 2-rank shard group + all_gather of the merged query "
                               "slices across the groups"}
    if rank != 0:
        return None
    peaks = measured_peaks()
    idx = i.cpu().numpy()
    n_cpu = 200
    from sklearn.neighbors import KDTree
    tree = KDTree(db)
    tree.query(q[:1], k=k)
    t0 = time.perf_counter()
    ref = np.stack([tree.query(q[j:j + 1], k=k)[1][0] for j in range(n_cpu)], 0)     # evaluate.py:481, one query per call
    cpu_qps = n_cpu / (time.perf_counter() - t0)
    d_local = (D + world - 1) // world
    sample_rows = 2048 if d_local > 4096 else 0
    score_flop = 3 * 2.0 * Q * (d_local + sample_rows) * dim          # three bf16 products per fp32-accurate dot product
    roof = {}
    if st.get("retrieve_score"):
        tf = score_flop / (st["retrieve_score"] * 1e-3) / 1e12
        roof["score"] = {"ms": st["retrieve_score"], "bound": "tensor", "achieved": tf, "peak": peaks["bf16_tflops_sustained"],
                         "unit": "TFLOP/s", "frac": tf / peaks["bf16_tflops_sustained"],
                         "note": "bf16-pair split + threshold sample + tcgen05 scoring with the candidate filter in the epilogue; "
                                 "FLOP = 3 products x 2 Q (D_local + sample rows) dim; no Q x D matrix is written"}
    if st.get("retrieve_select"):
        roof["select"] = {"ms": st["retrieve_select"], "bound": "latency/issue",
                          "note": "one warp per query over the emitted (score, row) lists (~2 % of the rows)"}
    if st.get("retrieve_rerank"):
        gb = Q * 32.0 * dim * 4 / (st["retrieve_rerank"] * 1e-3) / 1e9
        roof["rerank"] = {"ms": st["retrieve_rerank"], "bound": "fp64 pipe / gather", "achieved": gb, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                          "frac": gb / peaks["hbm_gbs"],
                          "note": "float64 re-rank of 32 candidates per query (sequential sums, sklearn order) + the exact-fallback launch"}
    return {"queries_per_s": Q / (ms * 1e-3), "ms_per_call": ms, "recall_at_1": float((idx[:, 0] == src).mean()), "D": D, "Q": Q, "k": k,
            "db_shards": world,
            "collective": "one nccl all_gather of the packed (fp64 dist | int64 idx)[2,Q,25] buffer per rank + merge" if world > 1 else None,
            "top25_identical_to_kdtree": bool(np.array_equal(idx[:n_cpu], ref)),
            "stages_ms": st, "roofline": roof,
            "large_batch": {"Q": QL, "queries_per_s": QL / (ms_l * 1e-3), "ms_per_call": ms_l}, "grid_2d": grid2,
            "cpu_kdtree_queries_per_s": cpu_qps, "cpu_sample": "%d queries, sklearn KDTree, 1 query per call" % n_cpu}


if __name__ == "__main__":
    main()
