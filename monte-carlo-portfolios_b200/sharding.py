"""Portfolio sharding across the GPUs of one box.

The hot path partitions by portfolio: rank g owns exposure rows [g*P/G, (g+1)*P/G) and the whole (replicated)
scenario matrix, so there is NO data-path collective.  torch.distributed is used only to assemble results on
rank 0 and to agree on the max-over-ranks time; one process per GPU, NCCL on a GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np


def portfolio_shard(P: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced: the first P % world ranks get one extra portfolio."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(P, world)
    p0 = rank * base + min(rank, extra)
    return p0, p0 + base + (1 if rank < extra else 0)


def trial_shard(n_trials: int, rank: int, world: int) -> tuple[int, int]:
    """Balanced contiguous trial ranges for the trial-sharded mode (every shard <= ceil(N / world) rows)."""
    return portfolio_shard(n_trials, rank, world)


def ts_candidates(n_global: int, tail: int, world: int) -> int:
    """Candidates per rank and portfolio of the first exchange round: the shard's expected share of the tail plus
    six binomial standard deviations (mirrors ts_candidates_per_rank in csrc/risk_kernels.cu)."""
    import math
    nmax = -(-n_global // world)
    share = tail * nmax / n_global
    m = math.ceil(share + 6.0 * math.sqrt(share * (1.0 - 1.0 / world)) + 8.0)
    return max(1, min(m, tail, nmax))


def ts_merge_rows(cands, tail: int, sent_all):
    """Host statement of the merge rule.  cands[g] = (loss[rows][m], trial[rows][m]) from rank g (any order inside a
    row); sent_all[g] = rank g had no more than m trials.  Returns (loss[rows][tail], trial[rows][tail] sorted by
    (loss, trial) descending, exact[rows]): a row is exact unless a rank that held candidates back has its smallest
    sent candidate above the merged threshold (same rule as k_ts_check)."""
    loss = np.concatenate([c[0] for c in cands], axis=1)
    trial = np.concatenate([c[1] for c in cands], axis=1)
    order = np.lexsort((trial, loss), axis=1)[:, ::-1][:, :tail]
    top_l = np.take_along_axis(loss, order, axis=1)
    top_t = np.take_along_axis(trial, order, axis=1)
    thr_l, thr_t = top_l[:, -1], top_t[:, -1]
    exact = np.ones(loss.shape[0], bool)
    for (l, tr), everything in zip(cands, sent_all):
        if everything:
            continue
        o = np.lexsort((tr, l), axis=1)[:, 0]
        ml = np.take_along_axis(l, o[:, None], axis=1)[:, 0]
        mt = np.take_along_axis(tr, o[:, None], axis=1)[:, 0]
        exact &= ~((ml > thr_l) | ((ml == thr_l) & (mt > thr_t)))
    return top_l, top_t, exact


def run_trial_sharded_gpu(ctx, E, S_local, alpha, n_global, trial_base, rank, world, codec_id, framing, uid: bytes):
    """One rank of a trial-sharded run on its own GPU: NCCL inside libmcp.so does the one exchange."""
    ctx.upload_exposures(E)
    ctx.upload_scenarios(S_local)
    ctx.comm_init(uid, rank, world)
    ctx.run_trial_sharded(alpha, n_global, trial_base, codec_id, framing)
    r = ctx.fetch()
    r["range"] = portfolio_shard(E.shape[0], rank, world)
    r["fallback_rows"] = ctx.run_stats()["fallback_rows"]
    return r


def gpu_compute(device: int):
    """Default shard compute: the CUDA path through the C ABI on `device` (no fallback)."""
    from .engine import Context

    def run(E_shard, S, alpha, codec_id, framing):
        with Context(device) as ctx:
            return ctx.simulate(E_shard, S, alpha, codec_id, framing)
    return run


def run_sharded(E, S, alpha, codec_id, framing, compute, dist=None):
    """Every rank values its own portfolio shard with `compute`; rank 0 returns the assembled result dict
    (other ranks return None).  `dist` = torch.distributed (initialised) or None for a single process."""
    rank = dist.get_rank() if dist is not None else 0
    world = dist.get_world_size() if dist is not None else 1
    P = E.shape[0]
    p0, p1 = portfolio_shard(P, rank, world)
    local = compute(np.ascontiguousarray(E[p0:p1]), S, alpha, codec_id, framing)
    local = {k: v for k, v in local.items()}
    local["range"] = (p0, p1)
    if dist is None:
        return assemble([local], P)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(local, gathered, dst=0)
    return assemble(gathered, P) if rank == 0 else None


def assemble(parts, P: int) -> dict:
    """Concatenate per-rank results in portfolio order; encoded blobs are re-based into one CSR."""
    parts = sorted(parts, key=lambda d: d["range"][0])
    assert parts[0]["range"][0] == 0 and parts[-1]["range"][1] == P
    for a, b in zip(parts, parts[1:]):
        assert a["range"][1] == b["range"][0], "shards must tile the portfolio range"
    out = {"tail": parts[0]["tail"]}
    for k in ("mean", "variance", "var_loss", "var_trial", "cvar", "breach"):
        out[k] = np.concatenate([d[k] for d in parts])
    offs, blobs, base = [np.zeros(1, np.int64)], [], 0
    for d in parts:
        eo = np.asarray(d["enc_off"], np.int64)
        offs.append(eo[1:] + base)
        blobs.append(np.asarray(d["enc"], np.uint8)[: eo[-1]])
        base += int(eo[-1])
    out["enc_off"] = np.concatenate(offs)
    out["enc"] = np.concatenate(blobs) if blobs else np.zeros(0, np.uint8)
    return out


def max_over_ranks(value: float, dist=None, device=None) -> float:
    """The multi-GPU timing rule: report the slowest rank."""
    if dist is None:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
