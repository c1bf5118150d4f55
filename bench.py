#!/usr/bin/env python3
"""bench.py -- headline benchmark of the hot path (BASELINE.json: portfolio-scenario valuations/s,
plus codec ints/s, with roofline fractions and a CPU baseline).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c5]

One "step" = one pass of the whole hot path (valuation -> moments/VaR/CVaR -> breach lists -> JavaFastPFOR
encoding) over one batch of synthetic portfolios whose inputs are already resident in HBM.
Workload at N=1 (BASELINE configs[1], "c2"): 10 000 portfolios x 100 000 trials x 32 factors, FP64, alpha 0.99.
N>1: portfolio-sharded, weak scaling -- every rank values its own 10 000-portfolio shard against the same
scenario matrix; there is no data-path collective (torch.distributed/NCCL is used only for the barrier and the
max-over-ranks of the device time).

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle (C restatement of the path; the Java
reference cannot run here: no JVM) on the box's host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (portfolios per GPU, trials, factors, alpha, seed)
    "c2": (10_000, 100_000, 32, 0.99, 0x5EED0002),
    "c3": (125_000, 10_000, 64, 0.99, 0x5EED0003),
    # trial-sharded (BASELINE config C5): ALL 100 000 portfolios on every GPU, 1 250 000 trials PER GPU (10 M on 8),
    # 99.9 % VaR/CVaR, one NCCL all-gather of tail candidates per step
    "c5": (100_000, 1_250_000, 32, 0.999, 0x5EED0005),
}
E_SCALE, S_DRIFT = 100.0, 0.05
METRIC, UNIT = "portfolio_scenario_valuations_per_sec", "valuations/s"


_JSON_OUT = None


def keep_stdout_for_the_json_line():
    """Libraries print to fd 1 (NCCL's version banner, for one): point fd 1 at stderr for the run and keep the real
    stdout for the ONE JSON line."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: str):
    out = _JSON_OUT or sys.stdout
    out.write(line + "\n")
    out.flush()


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        # under load = the upper half of the samples (idle samples before/after the region drag the median down)
        load = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def pinned_array(lib, shape, dtype):
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = ctypes.c_void_p()
    if lib.mcp_host_alloc_pinned(max(n, 16), ctypes.byref(p)) != 0:
        raise MemoryError("pinned allocation failed")
    buf = (ctypes.c_char * max(n, 16)).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


def cpu_baseline_risk(P, N, F, alpha, seed, target_s=12.0):
    """The oracle (C restatement; OpenMP over portfolios) on a bounded sample: all N trials, a few portfolios."""
    from oracle import cref
    cores = host_cores()  # torchrun exports OMP_NUM_THREADS=1: ask for every core explicitly
    E = cref.gen_normal(P, F, seed, 1, E_SCALE, 0.0, nthreads=cores)
    S = cref.gen_normal(N, F, seed, 2, 1.0, S_DRIFT, nthreads=cores)
    calib = min(P, 16 * cores)
    t0 = time.perf_counter()
    cref.risk(E, S, alpha, 0, calib, nthreads=cores)
    dt = time.perf_counter() - t0
    ps = int(min(P, max(calib, calib * target_s / max(dt, 1e-3))))  # about target_s seconds of all-core CPU work
    for _ in range(2):
        t0 = time.perf_counter()
        r = cref.risk(E, S, alpha, 0, ps, nthreads=cores)
        off = np.arange(ps + 1, dtype=np.int64) * r["tail"]
        enc, _ = cref.batch_compress(cref.CODEC_FASTPFOR256_VB | cref.FLAG_DELTA, 1, r["breach"].reshape(-1), off,
                                     4 * (r["tail"] + 64), nthreads=cores)
        dt = time.perf_counter() - t0
        if dt > target_s / 2 or ps >= P:
            break
        ps = int(min(P, ps * target_s / max(dt, 1e-3)))  # the calibration run under-estimated the rate: grow the sample once
    return {"value": ps * N / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{ps} of {P} portfolios x all {N} trials x {F} factors (valuation+moments+select+sort+encode), {dt:.1f} s"}, len(enc)


def cpu_baseline_codec(n_lists=200_000, n_trials=10_000, rate=0.01, seed=0x5EED0004, fastpfor=False):
    from oracle import cref
    cores = host_cores()
    rng = np.random.default_rng(seed)
    # gaps ~ geometric(rate): same distribution as the Bernoulli-mask lists the GPU generates
    lens = rng.binomial(n_trials, rate, size=n_lists)
    off = np.zeros(n_lists + 1, np.int64)
    np.cumsum(lens, out=off[1:])
    gaps = rng.geometric(rate, size=int(off[-1])).astype(np.int64)
    vals = np.cumsum(gaps)
    vals -= np.repeat(vals[off[:-1]] - gaps[off[:-1]], lens)  # restart each list near 0
    vals = vals.astype(np.int32)
    cid = (cref.CODEC_FASTPFOR256_VB if fastpfor else cref.CODEC_BP32_VB) | cref.FLAG_DELTA
    stride = 4 * (200 if n_trials <= 10_000 else 40 + n_trials // 50)
    t0 = time.perf_counter()
    buf, blens = cref.batch_compress(cid, 1, vals, off, stride, nthreads=cores)
    t1 = time.perf_counter()
    cref.batch_uncompress(cid, 1, buf, stride, blens, off, nthreads=cores)
    t2 = time.perf_counter()
    n = int(off[-1])
    return {"encode_ints_per_s": n / (t1 - t0), "decode_ints_per_s": n / (t2 - t1), "cores": cores, "kind": "port",
            "sample": f"{n_lists} lists x ~{n // n_lists} ints (1% of {n_trials} trials), "
                      f"delta+{'FastPFOR256' if fastpfor else 'BP32'}+VB app framing"}


def run_reference(args, rank, world):
    """--impl reference: the CPU oracle on the host cores (rank 0 only)."""
    if rank != 0:
        return
    P, N, F, alpha, seed = WORKLOADS[args.workload]
    vals = []
    cb = None
    for i in range(args.warmup + args.steps):
        cb, _ = cpu_baseline_risk(P, N, F, alpha, seed, target_s=8.0)
        if i >= args.warmup:
            vals.append(cb["value"])
    v = float(np.mean(vals))
    cb["value"] = v
    ms = 1e3 * P * N / v
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {P} portfolios x {N} trials x {F} factors, alpha {alpha}; CPU oracle "
                               "(C restatement of the path; the Java reference has no valuation code and no JVM exists here) "
                               "on a bounded sample, rate extrapolated per valuation"},
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def run_c5(args, rank, local_rank, world):
    """--workload c5: the trial-sharded mode.  Every rank holds all P exposure rows and its own 1.25 M-trial shard of
    the scenarios (generated on device from the global row index); one step = local candidates -> ONE NCCL all-gather
    -> merge -> breach lists / CVaR / moments / encoding of the rank's portfolio slice."""
    import mcp_b200
    from mcp_b200 import native as N
    import torch
    import torch.distributed as dist
    P, Nl, F, alpha, seed = WORKLOADS["c5"]
    if args.c5_portfolios:
        P = args.c5_portfolios
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = mcp_b200.Context(local_rank)
    uid = [ctx.comm_unique_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(uid[0], rank, world)
    n_global, base = Nl * world, rank * Nl
    ctx.generate_exposures(P, F, seed, E_SCALE, 0.0)
    ctx.generate_scenarios(Nl, F, seed, 1.0, S_DRIFT, row0=base)
    cid, framing = N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA, N.FRAME_APP
    dfma_peak, dmma_peak = ctx.fp64_peak(), ctx.fp64_mma_peak()
    peak = max(dfma_peak, dmma_peak)

    def barrier():
        ctx.sync()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        ctx.run_trial_sharded(alpha, n_global, base, cid, framing)
    ctx.profile(True)
    ctx.profile_reset()
    barrier()
    ms = []
    for _ in range(args.steps):
        ctx.flush_l2()
        barrier()
        ctx.timer_start()
        ctx.run_trial_sharded(alpha, n_global, base, cid, framing)
        ms.append(ctx.timer_stop())
    barrier()
    clocks = sampler.stop()
    prof = ctx.profile_get()
    dev_ms = float(np.sum(ms))
    if world > 1:
        t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
    ms_per_step = dev_ms / args.steps
    sizes, stats = ctx.run_sizes(), ctx.run_stats()
    m = int(N.lib().mcp_ts_candidates(n_global, alpha, world))
    block = 64 + P * m * 12 + P * 16
    v_ms = prof["value"]["ms"] / args.steps
    flops = 2.0 * F * P * Nl  # per rank
    if rank == 0:
        emit(json.dumps({
            "metric": METRIC, "value": P * n_global / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"c5: {P} portfolios (all on every GPU) x {Nl} trials per GPU ({n_global} in total) x {F} factors, alpha {alpha}, "
                                   f"tail {sizes['tail']}, trial-sharded: {m} tail candidates per portfolio and rank, one NCCL all-gather of "
                                   f"{block / 1e6:.0f} MB per rank, merge, Delta+FastPFOR256+VariableByte of the rank's portfolio slice",
                       "inputs": "Philox4x32-10 synthetic E and S generated on device (S rows by global trial index), resident in HBM",
                       "l2": "L2 flushed before every timed step; CUDA events on the library's stream, max over ranks",
                       "second_round_rows": stats["fallback_rows"]},
            "roofline": {"bound": "tensor", "kernel": "k_value_mma", "achieved": flops / (v_ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                         "frac": flops / (v_ms * 1e-3) / 1e12 / peak, "traffic": None,
                         "peak_source": f"live probes: DMMA {dmma_peak:.1f}, DFMA {dfma_peak:.1f} TF/s (of measured)",
                         "whole_step_frac": flops / (ms_per_step * 1e-3) / 1e12 / peak},
            "kernel_breakdown": {k: {"ms_per_step": v["ms"] / args.steps, "launches_per_step": v["launches"] / args.steps}
                                 for k, v in prof.items() if v["launches"]},
            "cpu_baseline": None, "e2e": None, "gpu_launches": int(sum(v["launches"] for v in prof.values())), "clocks": clocks,
        }))
    barrier()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-codec", action="store_true")
    ap.add_argument("--codec-lists", type=int, default=1_000_000)
    ap.add_argument("--c5-portfolios", type=int, default=0, help="override the c5 portfolio count (smaller smoke runs)")
    args = ap.parse_args()
    keep_stdout_for_the_json_line()
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline must use every host core, and libgomp reads
    # the variable when it is first loaded (with the oracle, later)
    os.environ["OMP_NUM_THREADS"] = str(host_cores())
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)
    if args.workload == "c5":
        run_c5(args, rank, local_rank, world)
        return

    import mcp_b200
    from mcp_b200 import native as N

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    ctx = mcp_b200.Context(local_rank)  # raises without libmcp.so / GPU: no fallback
    L = N.lib()
    P, Nn, F, alpha, seed = WORKLOADS[args.workload]
    cid, framing = N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA, N.FRAME_APP

    def barrier():
        ctx.sync()
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    # ---- inputs resident in HBM: this rank's portfolio shard + the replicated scenario matrix
    ctx.generate_exposures(P, F, seed, E_SCALE, 0.0, row0=rank * P)
    ctx.generate_scenarios(Nn, F, seed, 1.0, S_DRIFT)
    dfma_peak, dmma_peak = ctx.fp64_peak(), ctx.fp64_mma_peak()
    fp64_peak = max(dfma_peak, dmma_peak)  # the kernel issues DMMA: hold it to the higher of the two live probes
    sampler = ClockSampler(local_rank)
    sampler.start()  # sampled from the warm-up on: the timed region itself is only ~0.1 s
    for _ in range(args.warmup):
        ctx.run(alpha, cid, framing)
    ctx.profile(True)
    ctx.profile_reset()
    barrier()
    step_ms = []
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.flush_l2()          # not timed: evicts E/S/scratch from the 126 MB L2 between steps
        ctx.sync()
        ctx.timer_start()
        ctx.run(alpha, cid, framing)
        step_ms.append(ctx.timer_stop())
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    prof = ctx.profile_get()
    ctx.profile(False)
    dev_ms = float(np.sum(step_ms))
    if dist is not None:
        import torch
        t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
    sizes = ctx.run_sizes()
    ms_per_step = dev_ms / args.steps
    value = world * P * Nn / (ms_per_step * 1e-3)
    launches = int(sum(v["launches"] for v in prof.values()))

    # ---- roofline of the dominant kernel class (valuation): algorithmic FLOPs = 2*F per valuation
    kv = prof["value"]
    v_ms = kv["ms"] / max(args.steps, 1)
    flops = 2.0 * F * P * Nn
    achieved = flops / (v_ms * 1e-3) / 1e12 if v_ms > 0 else 0.0
    traffic = None
    try:  # DRAM bytes per launch of the dominant (sparse-step) valuation launch from the committed ncu --set full capture
        traffic = json.load(open(os.path.join(ROOT, "profiles", "value_kernel_traffic.json")))[args.workload]["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "tensor", "kernel": "k_value_mma (FP64 tensor-core valuation + fused tail filter)", "achieved": achieved,
                "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak if fp64_peak else None, "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json has no FP64 entry (it holds HBM GB/s and bf16 TF/s): the peak is the higher of two "
                               "live probes on this GPU, run just before the timed region -- dependent-free mma.sync.m8n8k4.f64 (DMMA, "
                               f"{dmma_peak:.1f} TF/s) and dependent-free scalar DFMA ({dfma_peak:.1f} TF/s); 'of measured'",
                "pipe": "FP64 tensor pipe (DMMA.8x8x4); algorithmic FLOPs = 2*F per valuation",
                "launches_per_step": kv["launches"] / max(args.steps, 1), "ms_per_step": v_ms,
                "whole_step_frac": (flops / (ms_per_step * 1e-3) / 1e12) / fp64_peak if fp64_peak else None}
    breakdown = {k: {"ms_per_step": v["ms"] / args.steps, "launches_per_step": v["launches"] / args.steps} for k, v in prof.items()
                 if v["launches"]}

    # ---- e2e: the same step through the one-call C ABI with HOST buffers (pinned), H2D + D2H inside the timed region
    e2e = None
    if rank == 0 or world > 1:
        Eh = pinned_array(L, (P, F), np.float64)
        Sh = pinned_array(L, (Nn, F), np.float64)
        Eh[:] = ctx.download_exposures(P, F)
        Sh[:] = ctx.download_scenarios(Nn, F)
        tl = sizes["tail"]
        cap = int(P * L.mcp_codec_max_bytes(cid, framing, tl))
        outs = {"mean": pinned_array(L, (P,), np.float64), "variance": pinned_array(L, (P,), np.float64),
                "var_loss": pinned_array(L, (P,), np.float64), "var_trial": pinned_array(L, (P,), np.int32),
                "cvar": pinned_array(L, (P,), np.float64), "breach": pinned_array(L, (P, tl), np.int32),
                "enc_off": pinned_array(L, (P + 1,), np.int64), "enc": pinned_array(L, (cap,), np.uint8)}
        n_enc = ctypes.c_int64(0)

        def ptr(a):
            return a.ctypes.data_as(ctypes.c_void_p)

        def one(raw):
            rc = L.mcp_simulate(ctx._h, ptr(Eh), P, ptr(Sh), Nn, F, alpha, cid, framing, ptr(outs["mean"]), ptr(outs["variance"]),
                                ptr(outs["var_loss"]), ptr(outs["var_trial"]), ptr(outs["cvar"]), ptr(outs["breach"]) if raw else None,
                                ptr(outs["enc_off"]), ptr(outs["enc"]), cap, ctypes.byref(n_enc))
            if rc != 0:
                raise RuntimeError(f"mcp_simulate failed: {rc} {L.mcp_last_error(ctx._h)}")

        def timed(raw):
            one(raw)
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                one(raw)
            barrier()
            e_s = (time.perf_counter() - t0) / args.steps
            if dist is not None:
                import torch
                t = torch.tensor([e_s], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                e_s = float(t.item())
            return e_s
        e_s, e_raw = timed(False), timed(True)
        h2d = Eh.nbytes + Sh.nbytes
        d2h = sum(outs[k].nbytes for k in ("mean", "variance", "var_loss", "var_trial", "cvar", "enc_off")) + int(n_enc.value)
        e2e = {"value": world * P * Nn / e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": e_s * 1e3,
               "api": "mcp_simulate, one call per step, pinned host buffers: uploads E and S, runs the path, downloads what the sink "
                      "document holds (mean, variance, VaR loss + trial, CVaR, CSR offsets and the encoded breach lists); wall clock "
                      "around the calls, max over ranks",
               "with_raw_breach_lists": {"value": world * P * Nn / e_raw, "ms_per_step": e_raw * 1e3,
                                         "d2h_bytes_per_step": int(d2h + outs["breach"].nbytes),
                                         "note": "same call, additionally downloading the uncompressed int32 breach lists"}}

    # ---- codec throughput (BASELINE config C4), device-resident.  Lists are independent, so N GPUs are N shards with no
    # collective (weak scaling, like the valuation): EVERY rank encodes and decodes its own lists (first_list = rank * nl),
    # the time is the max over ranks, ints/s is the whole job's; the roofline fraction is per GPU.
    codec = None
    pk = peaks()
    hbm_peak = pk["hbm_gbs"] if pk else 6650.0
    if not args.no_codec:
        src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if pk else "6650 GB/s (of fallback)"

        def over_ranks(x, op):
            if dist is None:
                return x
            import torch
            t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local_rank}")
            dist.all_reduce(t, op=op)
            return float(t.item())

        def codec_leg(cid, what, nl, ntrials):
            """encode + decode of nl breach lists (1 % of ntrials trials each), device-resident, L2 flushed before every
            timed call; algorithmic bytes = the ints (4 B each) + the encoded bytes + the CSR offsets in and out"""
            dv, do, total = ctx.generate_breach_lists(nl, ntrials, 0.01, 0x5EED0004, first_list=rank * nl)
            capb = 5 * total + 16 * nl
            barrier()
            d_oo, d_out, d_back = ctx.dev_alloc(8 * (nl + 1)), ctx.dev_alloc(capb), ctx.dev_alloc(4 * total + 16)
            enc_ms, dec_ms, enc_bytes = [], [], 0
            for it in range(args.warmup + args.steps):
                ctx.flush_l2()
                ctx.sync()
                ctx.timer_start()
                enc_bytes = ctx.codec_encode_dev(cid, N.FRAME_APP, dv, do, nl, total, d_oo, d_out, capb)
                t_e = ctx.timer_stop()
                ctx.flush_l2()
                ctx.sync()
                ctx.timer_start()
                ctx.codec_decode_dev(cid, N.FRAME_APP, d_out, d_oo, nl, d_back, do)
                t_d = ctx.timer_stop()
                if it >= args.warmup:
                    enc_ms.append(t_e)
                    dec_ms.append(t_d)
            for p in (d_oo, d_out, d_back):  # (dv / do belong to the context)
                ctx.dev_free(p)
            MAXOP, SUMOP = (dist.ReduceOp.MAX, dist.ReduceOp.SUM) if dist is not None else (None, None)
            e_ms, d_ms = over_ranks(float(np.mean(enc_ms)), MAXOP), over_ranks(float(np.mean(dec_ms)), MAXOP)
            all_ints, all_bytes = over_ranks(float(total), SUMOP), over_ranks(float(enc_bytes), SUMOP)
            algo_bytes = (4.0 * all_ints + all_bytes + 16.0 * nl * world) / world  # per GPU
            return {"workload": f"{nl} lists per GPU, 1% of {ntrials} trials (~{total // nl} ints each), {what}, app framing",
                    "n_gpus": world, "ints": int(all_ints), "encoded_bytes": int(all_bytes), "bytes_per_int": all_bytes / all_ints,
                    "encode_ints_per_s": all_ints / (e_ms * 1e-3), "decode_ints_per_s": all_ints / (d_ms * 1e-3),
                    "encode_ms": e_ms, "decode_ms": d_ms,
                    "roofline_encode": {"bound": "hbm", "achieved": algo_bytes / (e_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s per GPU",
                                        "frac": algo_bytes / (e_ms * 1e-3) / 1e9 / hbm_peak},
                    "roofline_decode": {"bound": "hbm", "achieved": algo_bytes / (d_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s per GPU",
                                        "frac": algo_bytes / (d_ms * 1e-3) / 1e9 / hbm_peak},
                    "peak_source": src}

        nl = args.codec_lists
        # BASELINE config C4 under the codec the north star names.  A list of ~100 ints is shorter than one FastPFOR block
        # (256), so Composition writes the 0 marker and the whole gap-coded list goes to VariableByte (FastPFOR.java:311-313)
        codec = codec_leg(N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA, "Delta + Composition(FastPFOR, VariableByte) [c4]", nl, 10_000)
        # the same lists under the codec the reference's loader really uses (LoadPortfolios.java:725-726)
        codec["loader_codec"] = codec_leg(N.CODEC_BP32_VB | N.FLAG_DELTA, "Delta + Composition(BinaryPacking, VariableByte) "
                                          "[c4, the loader's codec]", nl, 10_000)
        # lists long enough for FastPFOR pages (the risk step's own breach lists look like this)
        codec["fastpfor_pages"] = codec_leg(N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA, "Delta + Composition(FastPFOR, VariableByte), "
                                            "3 FastPFOR blocks + a VariableByte tail per list", max(nl // 10, 1), 100_000)
        st = ctx.codec_stats()
        codec["kernels"] = {"calls_served_by_thread_per_list_kernels": st["short_calls"], "handed_on": st["short_fallbacks"],
                            "calls_served_by_batched_kernels": st["fast_encodes"], "handed_back": st["fast_fallbacks"]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, _ = cpu_baseline_risk(P, Nn, F, alpha, seed)
        if codec is not None:
            codec["cpu_baseline"] = cpu_baseline_codec(fastpfor=True)
            codec["loader_codec"]["cpu_baseline"] = cpu_baseline_codec()
            codec["fastpfor_pages"]["cpu_baseline"] = cpu_baseline_codec(n_lists=20_000, n_trials=100_000, fastpfor=True)

    if rank == 0:
        emit(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"{args.workload}: {P} portfolios per GPU x {Nn} trials x {F} factors, alpha {alpha}, "
                                   f"tail {sizes['tail']} trials/portfolio, breach lists -> Delta+FastPFOR256+VariableByte, app framing",
                       "inputs": "Philox4x32-10 Irwin-Hall(12) synthetic E (scale 100) and S (drift 0.05), generated on device, resident in HBM",
                       "sharding": f"portfolio-sharded x{world}, no data-path collective",
                       "l2": "L2 flushed (256 MiB memset) before every timed step; steps timed individually with CUDA events on the library's stream",
                       "encoded_bytes_per_step": sizes["enc_bytes"]},
            "roofline": roofline, "kernel_breakdown": breakdown, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
            "clocks": clocks, "codec": codec, "wall_s_timed_region": wall,
        }))
    barrier()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
