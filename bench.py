#!/usr/bin/env python3
"""bench.py -- headline benchmark of the hot path (BASELINE.json: portfolio-scenario valuations/s,
plus codec ints/s, with roofline fractions and a CPU baseline).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload auto|c2|c3|c5]

One "step" = one pass of the whole hot path (valuation -> moments/VaR/CVaR -> breach lists -> JavaFastPFOR
encoding) over one batch of synthetic portfolios whose inputs are already resident in HBM.

Which configuration is the headline (`--workload auto`, the default) follows BASELINE.json:
  N = 1   configs[1] "c2": 10 000 portfolios x 100 000 trials x 32 factors, FP64, alpha 0.99, one B200.
          Sub-object `c3`: the whole of configs[2] (1 M portfolios x 10 k trials x 64 factors) on this one GPU -- the N = 1
          point of the strong-scaling curve below.
  N > 1   configs[2] "c3": 1 000 000 portfolios x 10 000 trials x 64 factors, PORTFOLIO-SHARDED across the N GPUs
          (rank g owns rows [g P/N, (g+1) P/N); the 5 MB scenario matrix is replicated; no data-path collective), strong
          scaling.  Sub-objects: `c3_one_gpu` (rank 0 alone runs the whole 1 M in the same job: the curve's N = 1 point),
          `c2_weak` (every rank runs a c2 shard: last round's headline), `c5` (configs[4]: 100 000 portfolios x 1.25 M
          trials PER GPU, trial-sharded, ONE NCCL all-gather of tail candidates per step, 99.9 % VaR/CVaR) and `codec`
          (configs[3]).  Four portfolios of every configuration are checked against the CPU oracle after the timed region
          (`parity_rows_checked`; the device generators are bit-identical to the oracle's, so no inputs are downloaded).
torch.distributed/NCCL carries the barrier and the max-over-ranks of the device time (and nothing else, except in c5).

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU baseline (oracle/risk_fast.c: the C restatement of the
path written for speed -- SIMD across trials, blocked, prefiltered select; bit-identical to the literal oracle; the Java
reference cannot run here: no JVM, and has no valuation code) on the box's host cores on a bounded sample of the same
workload.
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
    # c3 is 1 000 000 portfolios IN TOTAL (strong scaling): the per-GPU count is filled in from the world size
    "c3": (1_000_000, 10_000, 64, 0.99, 0x5EED0003),
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


def bind_to_gpu_numa_node(local_rank: int) -> dict:
    """Pin this rank to the CPUs of the NUMA node its GPU hangs off BEFORE any pinned allocation: cudaHostAlloc pages are
    first-touch, so the rank's pinned buffers then live on the socket whose PCIe root the GPU is on (8 ranks sharing one
    socket's memory and its UPI link is what stretched e2e from 3.24 to 3.47 ms at N = 8 last round)."""
    try:
        q = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local_rank)],
                           capture_output=True, text=True, timeout=20).stdout.strip().lower()
        bus = q[-12:] if len(q) >= 12 else q  # 00000000:1B:00.0 -> 0000:1b:00.0
        base = f"/sys/bus/pci/devices/{bus}"
        node = int(open(base + "/numa_node").read())
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= set(os.sched_getaffinity(0))
        if node < 0 or not cpus:
            return {"numa_node": node, "bound": False}
        os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "bound": True, "cpus": len(cpus)}
    except Exception as e:  # containers without sysfs / nvidia-smi: run unbound
        return {"bound": False, "why": repr(e)[:80]}


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


def workload_string(name: str, world: int) -> str:
    """The same `config.workload` text from both arms (ours and --impl reference)."""
    P, Nn, F, alpha, _ = WORKLOADS[name]
    if name == "c2":
        return (f"c2: {P} portfolios per GPU x {Nn} trials x {F} factors, alpha {alpha}, FP64; breach lists -> "
                f"Delta+FastPFOR256+VariableByte, app framing; portfolio-sharded x{world} (weak: one c2 shard per GPU)")
    if name == "c3":
        return (f"c3: {P} portfolios in total x {Nn} trials x {F} factors, alpha {alpha}, FP64; breach lists -> "
                f"Delta+FastPFOR256+VariableByte, app framing; portfolio-sharded x{world} (strong: {P // world} portfolios per GPU), "
                "no data-path collective")
    return (f"c5: {P} portfolios (all on every GPU) x {Nn} trials per GPU ({Nn * world} in total) x {F} factors, alpha {alpha}, FP64; "
            f"trial-sharded x{world}, one NCCL all-gather of tail candidates, merge, Delta+FastPFOR256+VariableByte of the rank's slice")


def headline_workload(args, world: int) -> str:
    if args.workload != "auto":
        return args.workload
    return "c2" if world == 1 else "c3"


def cpu_baseline_risk(P, N, F, alpha, seed, target_s=12.0):
    """The CPU baseline (oracle/risk_fast.c; OpenMP over portfolio groups) on a bounded sample: all N trials, as many
    portfolios as ~target_s seconds of all-core work hold."""
    from oracle import cref
    cores = host_cores()  # torchrun exports OMP_NUM_THREADS=1: ask for every core explicitly
    Ps = min(P, 200_000)
    E = cref.gen_normal(Ps, F, seed, 1, E_SCALE, 0.0, nthreads=cores)
    S = cref.gen_normal(N, F, seed, 2, 1.0, S_DRIFT, nthreads=cores)
    calib = min(Ps, 64 * cores)
    t0 = time.perf_counter()
    cref.risk(E, S, alpha, 0, calib, nthreads=cores, fast=True)
    dt = time.perf_counter() - t0
    ps = int(min(Ps, max(calib, calib * target_s / max(dt, 1e-3))))  # about target_s seconds of all-core CPU work
    for _ in range(2):
        t0 = time.perf_counter()
        r = cref.risk(E, S, alpha, 0, ps, nthreads=cores, fast=True)
        off = np.arange(ps + 1, dtype=np.int64) * r["tail"]
        enc, _ = cref.batch_compress(cref.CODEC_FASTPFOR256_VB | cref.FLAG_DELTA, 1, r["breach"].reshape(-1), off,
                                     4 * (r["tail"] + 64), nthreads=cores)
        dt = time.perf_counter() - t0
        if dt > target_s / 2 or ps >= Ps:
            break
        ps = int(min(Ps, ps * target_s / max(dt, 1e-3)))  # the calibration run under-estimated the rate: grow the sample once
    return {"value": ps * N / dt, "unit": UNIT, "cores": cores, "kind": "port", "isa": cref.fast_isa(),
            "sample": f"{ps} of {P} portfolios x all {N} trials x {F} factors (valuation+moments+select+sort+encode, "
                      f"oracle/risk_fast.c: SIMD across trials, bit-identical to the scalar oracle), {dt:.1f} s"}, len(enc)


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
    """--impl reference: the CPU baseline on the host cores (rank 0 only), on the configuration our arm headlines."""
    if rank != 0:
        return
    name = headline_workload(args, world)
    P, N, F, alpha, seed = WORKLOADS[name]
    if name == "c5":
        N = min(N, 1_250_000)  # one GPU's trial shard: the sample is per valuation anyway
    vals = []
    cb = None
    for i in range(args.warmup + args.steps):
        cb, _ = cpu_baseline_risk(P, N, F, alpha, seed, target_s=8.0)
        if i >= args.warmup:
            vals.append(cb["value"])
    v = float(np.mean(vals))
    cb["value"] = v
    total = P * N * (world if name in ("c2", "c5") else 1)
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / v, "higher_is_better": True,
        "scaling": "strong" if name == "c3" else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(name, world),
                   "arm": "CPU baseline = oracle/risk_fast.c (C restatement of the path written for speed; the Java reference has no "
                          "valuation code and no JVM exists here) on ONE host, all cores, on a bounded sample; rate per valuation"},
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------ our arm
class Job:
    """What every leg of the product arm shares: the context, the ranks, the barrier and the max-over-ranks rule."""

    def __init__(self, args, rank, local_rank, world):
        import mcp_b200
        from mcp_b200 import native as N
        self.args, self.rank, self.local_rank, self.world = args, rank, local_rank, world
        self.numa = bind_to_gpu_numa_node(local_rank) if world > 1 and not args.no_numa_bind else {"bound": False}
        self.N = N
        self.dist = None
        self.torch = None
        if world > 1:
            import torch
            import torch.distributed as dist
            torch.cuda.set_device(local_rank)
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            self.dist, self.torch = dist, torch
        self.ctx = mcp_b200.Context(local_rank)  # raises without libmcp.so / GPU: no fallback
        self.L = N.lib()
        self.cid, self.framing = N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA, N.FRAME_APP
        self.dfma_peak, self.dmma_peak = self.ctx.fp64_peak(), self.ctx.fp64_mma_peak()
        self.fp64_peak = max(self.dfma_peak, self.dmma_peak)  # the kernel issues DMMA: hold it to the higher live probe
        self.peak_source = ("MEASURED_PEAKS.json has no FP64 entry (it holds HBM GB/s and bf16 TF/s): the peak is the higher of two live "
                            "probes on this GPU, run just before the timed region -- dependent-free mma.sync.m8n8k4.f64 (DMMA, "
                            f"{self.dmma_peak:.1f} TF/s) and dependent-free scalar DFMA ({self.dfma_peak:.1f} TF/s); 'of measured'")
        self.launches = 0

    def barrier(self):
        self.ctx.sync()
        if self.dist is not None:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def over_ranks(self, x, op="max"):
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=f"cuda:{self.local_rank}")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return float(t.item())

    def traffic(self, name, P_local=None):
        """DRAM bytes per launch of the dominant valuation launch from the committed ncu capture of this workload shape
        (profiles/tools/capture_value_traffic.sh -> profiles/value_kernel_traffic.json; ncu cannot run inside the bench)"""
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", "value_kernel_traffic.json")))
            for key in (f"{name}_{P_local}", name):
                if key in t:
                    return t[key]["dram_bytes_per_launch"]
        except Exception:
            pass
        return None

    def close(self):
        self.barrier()
        self.ctx.close()
        if self.dist is not None:
            self.dist.destroy_process_group()


def parity_rows(job, name, P_local, row0, Nn, F, alpha, seed, n_global=None, trial_base=0, rows=4):
    """Spot check after the timed region: `rows` portfolios of the last run against the CPU oracle (the literal scalar one)
    on inputs regenerated on the host -- the device generators are bit-identical to orc_gen_normal.  VaR loss + trial index,
    breach list and encoded bytes bit-exact; mean / variance / CVaR to 1e-9.  Returns the number of rows that passed."""
    from oracle import cref
    ctx, cores = job.ctx, host_cores()
    r = ctx.fetch()
    Pl = r["mean"].shape[0]
    if Pl == 0:
        return 0
    pick = sorted({0, Pl // 3, (2 * Pl) // 3, Pl - 1})[:rows]
    Ntot = n_global if n_global is not None else Nn
    S = cref.gen_normal(Ntot, F, seed, 2, 1.0, S_DRIFT, nthreads=cores)
    ok = 0
    for p in pick:
        E1 = cref.gen_normal(1, F, seed, 1, E_SCALE, 0.0, row0=row0 + p)
        o = cref.risk(E1, S, alpha, nthreads=1)
        good = (r["var_trial"][p] == o["var_trial"][0] and r["var_loss"][p] == o["var_loss"][0]
                and np.array_equal(r["breach"][p], o["breach"][0])
                and bytes(r["enc"][r["enc_off"][p]: r["enc_off"][p + 1]]) == cref.get_compressed_instruments(job.cid, o["breach"][0])
                and np.allclose(r["cvar"][p], o["cvar"][0], rtol=1e-9, atol=0)
                and np.allclose(r["mean"][p], o["mean"][0], rtol=1e-9, atol=1e-300)
                and np.allclose(r["variance"][p], o["variance"][0], rtol=1e-9, atol=0))
        if not good:
            raise RuntimeError(f"{name}: portfolio {row0 + p} differs from the oracle")
        ok += 1
    return ok


def risk_leg(job, name, P_local, row0, ranks_active=None, check=True):
    """Device-timed steps of one portfolio-sharded configuration with the inputs resident in HBM.
    ranks_active = None: every rank runs its shard (time = max over ranks); = [0]: rank 0 alone (the others wait)."""
    args, ctx = job.args, job.ctx
    _, Nn, F, alpha, seed = WORKLOADS[name]
    active = ranks_active is None or job.rank in ranks_active
    step_ms, prof, sizes, clocks, wall = [], None, None, None, 0.0
    job.barrier()
    if active:
        ctx.generate_exposures(P_local, F, seed, E_SCALE, 0.0, row0=row0)
        ctx.generate_scenarios(Nn, F, seed, 1.0, S_DRIFT)
        sampler = ClockSampler(job.local_rank)
        sampler.start()  # sampled from the warm-up on: the timed region itself may be only ~0.1 s
        for _ in range(args.warmup):
            ctx.run(alpha, job.cid, job.framing)
        ctx.profile(True)
        ctx.profile_reset()
    job.barrier()
    if active:
        wall0 = time.perf_counter()
        for _ in range(args.steps):
            ctx.flush_l2()          # not timed: evicts E/S/scratch from the 126 MB L2 between steps
            ctx.sync()
            ctx.timer_start()
            ctx.run(alpha, job.cid, job.framing)
            step_ms.append(ctx.timer_stop())
        ctx.sync()
        wall = time.perf_counter() - wall0
    job.barrier()
    if active:
        clocks = sampler.stop()
        prof = ctx.profile_get()
        ctx.profile(False)
        sizes = ctx.run_sizes()
    dev_ms = job.over_ranks(float(np.sum(step_ms)) if active else 0.0)
    ms_per_step = dev_ms / args.steps
    out = {"ms_per_step": ms_per_step, "P_local": P_local, "N": Nn, "F": F, "alpha": alpha, "seed": seed}
    if active and job.rank == 0:
        kv = prof["value"]
        v_ms = kv["ms"] / max(args.steps, 1)
        flops = 2.0 * F * P_local * Nn
        achieved = flops / (v_ms * 1e-3) / 1e12 if v_ms > 0 else 0.0
        pk = job.fp64_peak
        out["roofline"] = {"bound": "tensor", "kernel": "k_value_mma (FP64 tensor-core valuation + fused tail filter)", "achieved": achieved,
                           "peak": pk, "unit": "TFLOP/s", "frac": achieved / pk if pk else None, "traffic": job.traffic(name, P_local),
                           "peak_source": job.peak_source, "pipe": "FP64 tensor pipe (DMMA.8x8x4); algorithmic FLOPs = 2*F per valuation",
                           "launches_per_step": kv["launches"] / max(args.steps, 1), "ms_per_step": v_ms,
                           "whole_step_frac": (flops / (ms_per_step * 1e-3) / 1e12) / pk if pk else None}
        out["kernel_breakdown"] = {k: {"ms_per_step": v["ms"] / args.steps, "launches_per_step": v["launches"] / args.steps}
                                   for k, v in prof.items() if v["launches"]}
        out["launches"] = int(sum(v["launches"] for v in prof.values()))
        out["clocks"], out["sizes"], out["wall_s_timed_region"] = clocks, sizes, wall
        out["fallback_rows"] = ctx.run_stats()["fallback_rows"]
        if check:
            out["parity_rows_checked"] = parity_rows(job, name, P_local, row0, Nn, F, alpha, seed)
    return out


def e2e_leg(job, name, P_local, row0):
    """The same step through the one-call C ABI with HOST buffers (pinned): H2D + D2H inside the timed region, wall clock."""
    args, ctx, L, N = job.args, job.ctx, job.L, job.N
    _, Nn, F, alpha, seed = WORKLOADS[name]
    cid, framing = job.cid, job.framing
    Eh = pinned_array(L, (P_local, F), np.float64)
    Sh = pinned_array(L, (Nn, F), np.float64)
    ctx.generate_exposures(P_local, F, seed, E_SCALE, 0.0, row0=row0)
    ctx.generate_scenarios(Nn, F, seed, 1.0, S_DRIFT)
    Eh[:] = ctx.download_exposures(P_local, F)
    Sh[:] = ctx.download_scenarios(Nn, F)
    tl = int(L.mcp_tail_size(Nn, alpha))
    # the encoded lists of this workload are far below the worst-case bound: size the pinned buffer from a first run
    ctx.run(alpha, cid, framing)
    cap = int(ctx.run_sizes()["enc_bytes"] * 1.25) + 4096
    outs = {"mean": pinned_array(L, (P_local,), np.float64), "variance": pinned_array(L, (P_local,), np.float64),
            "var_loss": pinned_array(L, (P_local,), np.float64), "var_trial": pinned_array(L, (P_local,), np.int32),
            "cvar": pinned_array(L, (P_local,), np.float64), "breach": pinned_array(L, (P_local, tl), np.int32),
            "enc_off": pinned_array(L, (P_local + 1,), np.int64), "enc": pinned_array(L, (cap,), np.uint8)}
    n_enc = ctypes.c_int64(0)

    def ptr(a):
        return a.ctypes.data_as(ctypes.c_void_p)

    def one(raw):
        rc = L.mcp_simulate(ctx._h, ptr(Eh), P_local, ptr(Sh), Nn, F, alpha, cid, framing, ptr(outs["mean"]), ptr(outs["variance"]),
                            ptr(outs["var_loss"]), ptr(outs["var_trial"]), ptr(outs["cvar"]), ptr(outs["breach"]) if raw else None,
                            ptr(outs["enc_off"]), ptr(outs["enc"]), cap, ctypes.byref(n_enc))
        if rc != 0:
            raise RuntimeError(f"mcp_simulate failed: {rc} {L.mcp_last_error(ctx._h)}")

    def timed(raw):
        one(raw)
        job.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            one(raw)
        job.barrier()
        return job.over_ranks((time.perf_counter() - t0) / args.steps)
    e_s, e_raw = timed(False), timed(True)

    # ---- the same calls, two in flight: a second context on the same GPU and one host thread per context (contexts are
    # independent; a context is used by one thread at a time), so that the uploads and downloads of one step ride under the
    # kernels of the other.  Every step still uploads its own inputs and downloads its own results inside the timed region.
    import threading
    import mcp_b200
    depth = 2
    ctx2 = mcp_b200.Context(job.local_rank)
    if getattr(args, "simulate_turns", -1) >= 0:  # experiments: the library decides by upload size otherwise
        ctx.set_option("simulate_turns", args.simulate_turns)
        ctx2.set_option("simulate_turns", args.simulate_turns)
    outs2 = {k: pinned_array(L, v.shape, v.dtype) for k, v in outs.items()}
    n_enc2 = ctypes.c_int64(0)
    lanes = [(ctx, outs, n_enc), (ctx2, outs2, n_enc2)]

    def one_on(lane):
        c, o, n = lane
        rc = L.mcp_simulate(c._h, ptr(Eh), P_local, ptr(Sh), Nn, F, alpha, cid, framing, ptr(o["mean"]), ptr(o["variance"]),
                            ptr(o["var_loss"]), ptr(o["var_trial"]), ptr(o["cvar"]), None, ptr(o["enc_off"]), ptr(o["enc"]), cap,
                            ctypes.byref(n))
        if rc != 0:
            raise RuntimeError(f"mcp_simulate failed: {rc} {L.mcp_last_error(c._h)}")
    per = (args.steps + depth - 1) // depth
    errs = []

    def worker(lane, count):
        try:
            for _ in range(count):
                one_on(lane)
        except Exception as ex:  # surfaced after the join
            errs.append(ex)
    for lane in lanes:
        one_on(lane)  # warm-up: allocations of the second context
    job.barrier()
    ctx2.sync()
    t0 = time.perf_counter()
    th = [threading.Thread(target=worker, args=(lane, per)) for lane in lanes]
    for t_ in th:
        t_.start()
    for t_ in th:
        t_.join()
    ctx2.sync()
    job.barrier()
    e_pipe = job.over_ranks((time.perf_counter() - t0) / (per * depth))
    if errs:
        raise errs[0]
    same = all(np.array_equal(outs[k], outs2[k]) for k in ("mean", "variance", "var_loss", "var_trial", "cvar", "enc_off")) \
        and n_enc.value == n_enc2.value and np.array_equal(outs["enc"][: n_enc.value], outs2["enc"][: n_enc2.value])
    if not same:
        raise RuntimeError("e2e: the two pipelined contexts disagree")
    for a in outs2.values():
        L.mcp_host_free_pinned(ctypes.c_void_p(a.ctypes.data))
    ctx2.close()
    h2d = Eh.nbytes + Sh.nbytes
    d2h = sum(outs[k].nbytes for k in ("mean", "variance", "var_loss", "var_trial", "cvar", "enc_off")) + int(n_enc.value)
    units = P_local * Nn * (job.world if name == "c2" else 1) if name == "c2" else WORKLOADS[name][0] * Nn
    for a in list(outs.values()) + [Eh, Sh]:
        L.mcp_host_free_pinned(ctypes.c_void_p(a.ctypes.data))
    return {"value": units / e_pipe, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "ms_per_step": e_pipe * 1e3,
            "api": "mcp_simulate, one call per step, pinned host buffers: every call uploads E (the rank's shard) and S, runs the path and "
                   "downloads what the sink document holds (mean, variance, VaR loss + trial, CVaR, CSR offsets and the encoded breach "
                   "lists).  Two calls in flight per rank (two contexts on the GPU, one host thread each) so that one step's copies ride "
                   "under the other's kernels; wall clock around all the calls / steps, max over ranks; bytes are per rank and step",
            "calls_in_flight": depth,
            "one_call_at_a_time": {"value": units / e_s, "ms_per_step": e_s * 1e3,
                                   "note": "the same calls back to back on one context (the latency of a step from host buffers)"},
            "with_raw_breach_lists": {"value": units / e_raw, "ms_per_step": e_raw * 1e3,
                                      "d2h_bytes_per_step": int(d2h + outs["breach"].nbytes),
                                      "note": "same call, additionally downloading the uncompressed int32 breach lists"}}


def c5_leg(job, steps, warmup):
    """configs[4]: the trial-sharded mode.  Every rank holds all P exposure rows and its own 1.25 M-trial shard of the
    scenarios (generated on device from the global row index); one step = local candidates -> ONE NCCL all-gather ->
    merge -> breach lists / CVaR / moments / encoding of the rank's portfolio slice."""
    args, ctx, N, world, rank = job.args, job.ctx, job.N, job.world, job.rank
    P, Nl, F, alpha, seed = WORKLOADS["c5"]
    if args.c5_portfolios:
        P = args.c5_portfolios
    uid = [ctx.comm_unique_id() if rank == 0 else None]
    if world > 1:
        job.dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(uid[0], rank, world)
    ctx.set_option("ts_exchange", 0 if args.c5_allgather else 1)
    n_global, base = Nl * world, rank * Nl
    ctx.generate_exposures(P, F, seed, E_SCALE, 0.0)
    ctx.generate_scenarios(Nl, F, seed, 1.0, S_DRIFT, row0=base)
    cid, framing = job.cid, job.framing
    sampler = ClockSampler(job.local_rank)
    sampler.start()
    for _ in range(warmup):
        ctx.run_trial_sharded(alpha, n_global, base, cid, framing)
    ctx.profile(True)
    ctx.profile_reset()
    job.barrier()
    ms = []
    for _ in range(steps):
        ctx.flush_l2()
        job.barrier()
        ctx.timer_start()
        ctx.run_trial_sharded(alpha, n_global, base, cid, framing)
        ms.append(ctx.timer_stop())
    job.barrier()
    clocks = sampler.stop()
    prof = ctx.profile_get()
    ctx.profile(False)
    ms_per_step = job.over_ranks(float(np.sum(ms))) / steps
    sizes, stats = ctx.run_sizes(), ctx.run_stats()
    m = int(N.lib().mcp_ts_candidates(n_global, alpha, world))
    block = 64 + P * m * 12 + P * 16
    v_ms = prof["value"]["ms"] / steps
    comm_ms = prof["comm"]["ms"] / steps
    flops = 2.0 * F * P * Nl  # per rank
    pk = job.fp64_peak
    out = None
    checked = None
    if rank == 0 and not args.no_parity:
        # rank 0's slice starts at portfolio 0; the oracle needs the whole scenario matrix on the host (bounded: <= 2.6 GB)
        checked = parity_rows(job, "c5", sizes["P"], 0, Nl, F, alpha, seed, n_global=n_global, rows=2 if n_global > 4_000_000 else 4) \
            if n_global * F * 8 <= (3 << 30) else 0
    if rank == 0:
        gathered = block * world
        # a rank merges only its own portfolio slice, so it is sent only that slice of every other rank's block (grouped
        # ncclSend / ncclRecv, option ts_exchange = 1, the default); ts_exchange = 0 is the all-gather of whole blocks
        slices = not args.c5_allgather
        moved = (block - 64) * (world - 1) // world + 64 * (world - 1) if slices else gathered - block
        out = {"metric": METRIC, "value": P * n_global / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
               "ms_per_step": ms_per_step, "scaling": "weak (trials per GPU fixed)",
               "config": {"workload": workload_string("c5", world) if not args.c5_portfolios else f"c5 with {P} portfolios",
                          "tail": sizes["tail"], "candidates_per_portfolio_and_rank": m,
                          "inputs": "Philox4x32-10 synthetic E and S generated on device (S rows by global trial index), resident in HBM",
                          "l2": "L2 flushed before every timed step; CUDA events on the library's stream, max over ranks",
                          "second_round_rows": stats["fallback_rows"]},
               "collective": {"name": ("slice exchange: one group of ncclSend / ncclRecv per step -- every rank receives its own portfolio slice of "
                                       "every other rank's candidate block (1 / world of what an all-gather of the blocks moves; --c5-allgather "
                                       "runs that instead)" if slices else "ncclAllGather of the ranks' candidate blocks (one per step)")
                                      + "; + an ncclAllGather of one byte per portfolio of exactness flags",
                              "bytes_sent_per_rank": moved if slices else block, "bytes_received_per_rank": moved, "ms_per_step": comm_ms,
                              "all_gather_bytes_received_per_rank": gathered - block,
                              "achieved_gbs_received": moved / (comm_ms * 1e-3) / 1e9 if comm_ms > 0 else None,
                              "peer_copy_peak_gbs": 770.0, "peak_source": "SURVEY.md 8(e): measured peer-copy bandwidth per direction"},
               "roofline": {"bound": "tensor", "kernel": "k_value_mma", "achieved": flops / (v_ms * 1e-3) / 1e12, "peak": pk, "unit": "TFLOP/s",
                            "frac": flops / (v_ms * 1e-3) / 1e12 / pk, "traffic": job.traffic("c5"), "peak_source": job.peak_source,
                            "whole_step_frac": flops / (ms_per_step * 1e-3) / 1e12 / pk},
               "kernel_breakdown": {k: {"ms_per_step": v["ms"] / steps, "launches_per_step": v["launches"] / steps}
                                    for k, v in prof.items() if v["launches"]},
               "parity_rows_checked": checked, "gpu_launches": int(sum(v["launches"] for v in prof.values())), "clocks": clocks}
    job.barrier()
    ctx.comm_destroy()
    return out


def weights_leg(ctx, args, nl=100_000, per_list=320):
    """Row f3: the loader's weights column (LoadPortfolios.getCompressedWeights, raw Snappy over the little-endian doubles;
    C1's portfolios hold ~320 positions each).  mcp_weights_encode / _decode take HOST buffers, so the wall time holds the copies;
    the kernels' own time is the library's event time around its launches."""
    rng = np.random.default_rng(0x5EED0F3)
    lens = rng.integers(per_list // 2, per_list * 3 // 2, nl)
    off = np.zeros(nl + 1, np.int64)
    np.cumsum(lens, out=off[1:])
    # weights as the generator writes them: a few decimal digits, many repeats inside a portfolio (that is what Snappy finds)
    w = np.round(rng.random(int(off[-1])) * 100.0, 1)
    reps = max(min(args.steps, 5), 1)
    oo, blob = ctx.weights_encode(w, off)
    ctx.profile(True)
    ctx.profile_reset()
    t0 = time.perf_counter()
    for _ in range(reps):
        oo, blob = ctx.weights_encode(w, off)
    t_e = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for _ in range(reps):
        back = ctx.weights_decode(blob, oo, off)
    t_d = (time.perf_counter() - t0) / reps
    prof = ctx.profile_get()
    ctx.profile(False)
    assert np.array_equal(back, w), "weights column round trip"
    return {"workload": f"{nl} portfolios x ~{per_list} weights (doubles with one decimal), raw-Snappy format, host buffers",
            "doubles": int(off[-1]), "encoded_bytes": int(blob.size), "ratio": float(blob.size) / (8.0 * float(off[-1])),
            "encode_doubles_per_s_host_call": float(off[-1]) / t_e, "decode_doubles_per_s_host_call": float(off[-1]) / t_d,
            "encode_kernels_ms": prof["encode"]["ms"] / reps, "decode_kernels_ms": prof["decode"]["ms"] / reps,
            "round_trip_checked": True}


def codec_legs(job):
    """BASELINE config C4, device-resident.  Lists are independent, so N GPUs are N shards with no collective (weak scaling):
    EVERY rank encodes and decodes its own lists (first_list = rank * nl), the time is the max over ranks, ints/s is the whole
    job's; the roofline fraction is per GPU."""
    args, ctx, N, world, rank = job.args, job.ctx, job.N, job.world, job.rank
    pk = peaks()
    hbm_peak = pk["hbm_gbs"] if pk else 6650.0
    src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if pk else "6650 GB/s (of fallback)"

    def codec_leg(cid, what, nl, ntrials):
        """encode + decode of nl breach lists (1 % of ntrials trials each), device-resident, L2 flushed before every
        timed call; algorithmic bytes = the ints (4 B each) + the encoded bytes + the CSR offsets in and out"""
        dv, do, total = ctx.generate_breach_lists(nl, ntrials, 0.01, 0x5EED0004, first_list=rank * nl)
        capb = 5 * total + 16 * nl
        job.barrier()
        d_oo, d_out, d_back = ctx.dev_alloc(8 * (nl + 1)), ctx.dev_alloc(capb), ctx.dev_alloc(4 * total + 16)
        enc_ms, dec_ms, enc_bytes = [], [], 0
        ctx.profile(True)
        for it in range(args.warmup + args.steps):
            if it == args.warmup:
                ctx.profile_reset()
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
        prof = ctx.profile_get()
        ctx.profile(False)
        for p in (d_oo, d_out, d_back):  # (dv / do belong to the context)
            ctx.dev_free(p)
        e_ms, d_ms = job.over_ranks(float(np.mean(enc_ms))), job.over_ranks(float(np.mean(dec_ms)))
        # the same calls by the library's own events around its kernel launches: without the host round trip that hands the
        # encoded size back (the call-level figures above include it: ~25 us of idle GPU per encode call)
        ek_ms = job.over_ranks(prof["encode"]["ms"] / max(args.steps, 1))
        dk_ms = job.over_ranks(prof["decode"]["ms"] / max(args.steps, 1))
        all_ints, all_bytes = job.over_ranks(float(total), "sum"), job.over_ranks(float(enc_bytes), "sum")
        algo_bytes = (4.0 * all_ints + all_bytes + 16.0 * nl * world) / world  # per GPU
        return {"workload": f"{nl} lists per GPU, 1% of {ntrials} trials (~{total // nl} ints each), {what}, app framing",
                "n_gpus": world, "ints": int(all_ints), "encoded_bytes": int(all_bytes), "bytes_per_int": all_bytes / all_ints,
                "encode_ints_per_s": all_ints / (e_ms * 1e-3), "decode_ints_per_s": all_ints / (d_ms * 1e-3),
                "encode_ms": e_ms, "decode_ms": d_ms, "encode_kernels_ms": ek_ms, "decode_kernels_ms": dk_ms,
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
    # the other wire formats of SURVEY.md section 8 rows a8 / f4 on the C4 lists (same timing rules, fewer repetitions)
    if not args.no_extras:
        keep = (args.steps, args.warmup)
        args.steps, args.warmup = max(min(args.steps, 5), 1), min(args.warmup, 3)
        others = {}
        for key, cid, what in (("ibp32_ivb", N.CODEC_IBP32_IVB, "IntegratedComposition(IntegratedBinaryPacking, IntegratedVariableByte)"),
                               ("sk_ibp32_ivb", N.CODEC_SK_IBP32_IVB, "IntegratedIntCompressor framing of the same"),
                               ("fastpfor128_vb", N.CODEC_FASTPFOR128_VB | N.FLAG_DELTA, "Delta + Composition(FastPFOR128, VariableByte)"),
                               ("xor128_ivb", N.CODEC_XOR128_IVB, "IntegratedComposition(XorBinaryPacking, IntegratedVariableByte)"),
                               ("optpfd_vb", N.CODEC_OPTPFD_VB | N.FLAG_DELTA, "Delta + Composition(OptPFD, VariableByte)")):
            leg = codec_leg(cid, what + " [c4 lists]", nl, 10_000)
            others[key] = {k: leg[k] for k in ("workload", "bytes_per_int", "encode_ints_per_s", "decode_ints_per_s", "encode_ms", "decode_ms")}
            others[key]["roofline_frac"] = [leg["roofline_encode"]["frac"], leg["roofline_decode"]["frac"]]
        codec["other_codecs"] = others
        args.steps, args.warmup = keep
        if rank == 0:
            codec["weights_column"] = weights_leg(ctx, args)
    st = ctx.codec_stats()
    codec["kernels"] = {"calls_served_by_thread_per_list_kernels": st["short_calls"], "handed_on": st["short_fallbacks"],
                        "calls_served_by_batched_kernels": st["fast_encodes"], "handed_back": st["fast_fallbacks"]}
    # the drop-in single-list call on the shipped sample's lists (17 .. 1 110 ints): a GPU round trip per list
    if rank == 0:
        try:
            c1 = json.load(open(os.path.join(ROOT, "tests", "golden", "c1_sample.json")))
            lists = [np.asarray(p["instruments"], np.int32) for p in c1["portfolios"]]
            for v in lists:
                ctx.get_compressed_instruments(N.CODEC_BP32_VB, v)
            t0 = time.perf_counter()
            reps = 20
            for _ in range(reps):
                for v in lists:
                    ctx.get_compressed_instruments(N.CODEC_BP32_VB, v)
            per_call = (time.perf_counter() - t0) / (reps * len(lists))
            vals = np.concatenate(lists)
            off = np.zeros(len(lists) + 1, np.int64)
            np.cumsum([len(v) for v in lists], out=off[1:])
            ctx.codec_encode(N.CODEC_BP32_VB, N.FRAME_APP, vals, off)
            t0 = time.perf_counter()
            for _ in range(reps):
                ctx.codec_encode(N.CODEC_BP32_VB, N.FRAME_APP, vals, off)
            batched = (time.perf_counter() - t0) / reps
            codec["single_list_latency"] = {"what": "mcp_get_compressed_instruments on each of the 10 instrument lists of portfolios.json (C1), host "
                                                    "buffers, one call = one GPU round trip", "us_per_call": per_call * 1e6,
                                            "us_for_all_ten_in_one_batched_call": batched * 1e6}
        except Exception as e:  # the fixture is not on this box
            codec["single_list_latency"] = {"unavailable": repr(e)}
    return codec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto"] + sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-codec", action="store_true")
    ap.add_argument("--c5-allgather", action="store_true", help="c5: exchange whole candidate blocks by ncclAllGather (ts_exchange = 0)")
    ap.add_argument("--simulate-turns", type=int, default=-1, help="experiments: force mcp_set_option('simulate_turns') in the e2e leg")
    ap.add_argument("--no-extras", action="store_true", help="only the headline leg (profiling runs under ncu)")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle spot check after the timed region")
    ap.add_argument("--no-numa-bind", action="store_true", help="multi-rank runs: do not pin the rank to its GPU's NUMA node")
    ap.add_argument("--codec-lists", type=int, default=1_000_000)
    ap.add_argument("--c3-portfolios", type=int, default=0, help="override the c3 TOTAL portfolio count (smaller smoke runs)")
    ap.add_argument("--c5-portfolios", type=int, default=0, help="override the c5 portfolio count (smaller smoke runs)")
    args = ap.parse_args()
    keep_stdout_for_the_json_line()
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline must use every host core, and libgomp reads
    # the variable when it is first loaded (with the oracle, later)
    os.environ["OMP_NUM_THREADS"] = str(host_cores())
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.c3_portfolios:
        WORKLOADS["c3"] = (args.c3_portfolios,) + WORKLOADS["c3"][1:]
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)
    from mcp_b200 import sharding
    job = Job(args, rank, local_rank, world)
    name = headline_workload(args, world)
    extras = not args.no_extras
    check = not args.no_parity
    line = None

    if name == "c5":
        line = c5_leg(job, args.steps, args.warmup)
        if rank == 0:
            line.update({"higher_is_better": True, "vs_baseline": None, "dtype": "f64", "data": "synthetic", "scaling": "weak",
                         "cpu_baseline": None, "e2e": None})
    else:
        Ptot, Nn, F, alpha, seed = WORKLOADS[name]
        if name == "c2":
            P_local, row0, units, scaling = Ptot, rank * Ptot, world * Ptot * Nn, "weak"
        else:
            p0, p1 = sharding.portfolio_shard(Ptot, rank, world)
            P_local, row0, units, scaling = p1 - p0, p0, Ptot * Nn, "strong"
        head = risk_leg(job, name, P_local, row0, check=check)
        e2e = e2e_leg(job, name, P_local, row0)
        sub = {}
        if extras:
            if world == 1 and name == "c2":
                c3P = WORKLOADS["c3"][0]
                c3 = risk_leg(job, "c3", c3P, 0, check=check)
                if rank == 0:
                    c3["value"] = c3P * WORKLOADS["c3"][1] / (c3["ms_per_step"] * 1e-3)
                    c3["config"] = {"workload": workload_string("c3", 1)}
                    c3["scaling_note"] = "the N = 1 point of the c3 strong-scaling curve (bench.py --gpus N headlines c3 for N > 1)"
                sub["c3"] = c3
            if world > 1 and name == "c3":
                one = risk_leg(job, "c3", Ptot, 0, ranks_active=[0], check=False)
                if rank == 0:
                    one["value"] = Ptot * Nn / (one["ms_per_step"] * 1e-3)
                    one["note"] = "rank 0 alone runs the whole 1 M portfolios in the same job (the other ranks wait): the N = 1 point of this curve"
                sub["c3_one_gpu"] = one
                c2P, c2N = WORKLOADS["c2"][0], WORKLOADS["c2"][1]
                c2 = risk_leg(job, "c2", c2P, rank * c2P, check=check)
                c2e = e2e_leg(job, "c2", c2P, rank * c2P)
                if rank == 0:
                    c2["value"] = world * c2P * c2N / (c2["ms_per_step"] * 1e-3)
                    c2["config"] = {"workload": workload_string("c2", world)}
                    c2["scaling"] = "weak"
                    c2["e2e"] = c2e
                sub["c2_weak"] = c2
                sub["c5"] = c5_leg(job, max(3, min(args.steps, 5)), 3)
            if not args.no_codec:
                sub["codec"] = codec_legs(job)
        cpu = None
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            cpu, _ = cpu_baseline_risk(Ptot, Nn, F, alpha, seed)
            if sub.get("codec") is not None:
                sub["codec"]["cpu_baseline"] = cpu_baseline_codec(fastpfor=True)
                sub["codec"]["loader_codec"]["cpu_baseline"] = cpu_baseline_codec()
                sub["codec"]["fastpfor_pages"]["cpu_baseline"] = cpu_baseline_codec(n_lists=20_000, n_trials=100_000, fastpfor=True)
        if rank == 0:
            launches = head["launches"] + sum(int(v.get("launches", 0) or 0) + int(v.get("gpu_launches", 0) or 0)
                                              for v in sub.values() if isinstance(v, dict))
            line = {
                "metric": METRIC, "value": units / (head["ms_per_step"] * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": scaling,
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_string(name, world),
                           "tail": head["sizes"]["tail"],
                           "inputs": "Philox4x32-10 Irwin-Hall(12) synthetic E (scale 100) and S (drift 0.05), generated on device, resident in HBM",
                           "sharding": f"portfolio-sharded x{world}, no data-path collective",
                           "numa": job.numa,
                           "l2": "L2 flushed (256 MiB memset) before every timed step; steps timed individually with CUDA events on the library's stream",
                           "encoded_bytes_per_step": head["sizes"]["enc_bytes"], "fallback_rows": head["fallback_rows"]},
                "roofline": head["roofline"], "kernel_breakdown": head["kernel_breakdown"], "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": launches, "gpu_launches_headline_leg": head["launches"], "clocks": head["clocks"],
                "parity_rows_checked": head.get("parity_rows_checked"), "wall_s_timed_region": head["wall_s_timed_region"],
            }
            if world > 1 and name == "c3":
                line["scaling_note"] = ("strong scaling of BASELINE configs[2]; the N = 1 point of THIS workload is `c3_one_gpu.value` here and "
                                        "`c3.value` in the --gpus 1 line, whose headline is configs[1] (c2) as BASELINE.json names it for one GPU")
            line.update(sub)
    if rank == 0 and line is not None:
        emit(json.dumps(line))
    job.close()


if __name__ == "__main__":
    main()
