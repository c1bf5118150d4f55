"""N > 1 host logic on CPU: world_size 2, gloo.  The shard compute is the oracle here (this is a test of the
sharding/assembly plumbing, not of the CUDA path; the GPU box runs the same code with mcp_b200.sharding.gpu_compute)."""
import os
import socket

import numpy as np
import pytest

import riskcases
from mcp_b200 import sharding
from oracle import cref

CID = cref.CODEC_FASTPFOR256_VB | cref.FLAG_DELTA


def oracle_compute(E_shard, S, alpha, codec_id, framing):
    r = cref.risk(E_shard, S, alpha)
    blobs = [cref.get_compressed_instruments(codec_id, r["breach"][i]) for i in range(E_shard.shape[0])]
    off = np.zeros(len(blobs) + 1, np.int64)
    np.cumsum([len(b) for b in blobs], out=off[1:])
    r["enc_off"] = off
    r["enc"] = np.frombuffer(b"".join(blobs), np.uint8) if blobs else np.zeros(0, np.uint8)
    return r


def test_portfolio_shard_tiles_the_range():
    for P in (0, 1, 7, 10, 10000, 1000003):
        for world in (1, 2, 3, 4, 8):
            ranges = [sharding.portfolio_shard(P, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == P
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.portfolio_shard(10, 2, 2)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    E, S = riskcases.synthetic(37, 2000, 32, 5)
    out = sharding.run_sharded(E, S, 0.99, CID, 1, oracle_compute, dist)
    slow = sharding.max_over_ranks(1.0 + rank, dist)
    dist.barrier()
    if rank == 0:
        q.put((out, slow))
    dist.destroy_process_group()


def test_world_size_2_gloo_matches_single_process():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out, slow = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert slow == 2.0  # max over ranks
    E, S = riskcases.synthetic(37, 2000, 32, 5)
    ref = sharding.run_sharded(E, S, 0.99, CID, 1, oracle_compute, None)
    for k in ("mean", "variance", "var_loss", "var_trial", "cvar", "breach", "enc_off", "enc"):
        assert np.array_equal(out[k], ref[k]), k
    assert out["enc_off"][-1] == out["enc"].size and out["breach"].shape == (37, 21)


# ---- trial-sharded mode: host statement of the exchange + merge rule, world_size 2 over gloo -----------------------
def _ts_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    E, S = riskcases.synthetic(23, 6001, 16, 9)
    S[1000:1200] *= 3.0            # fat region on rank 0: uneven tail shares
    alpha, Nn = 0.98, S.shape[0]
    t = int(cref.tail_size(Nn, alpha))
    out = {}
    for label, m in (("auto", sharding.ts_candidates(Nn, t, world)), ("tight", -(-t // world))):
        n0, n1 = sharding.trial_shard(Nn, rank, world)
        loss = 0.0 - cref.value(E, S[n0:n1])                      # this rank's shard of the losses
        trial = np.broadcast_to(np.arange(n0, n1, dtype=np.int64), loss.shape)
        o = np.lexsort((trial, loss), axis=1)[:, ::-1][:, :m]     # local top-m by (loss, trial)
        mine = torch.from_numpy(np.stack([np.take_along_axis(loss, o, 1), np.take_along_axis(trial, o, 1).astype(np.float64)]))
        got = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(got, mine)                                # the one exchange
        cands = [(g[0].numpy(), g[1].numpy().astype(np.int64)) for g in got]
        sent_all = [m >= sharding.trial_shard(Nn, g, world)[1] - sharding.trial_shard(Nn, g, world)[0] for g in range(world)]
        out[label] = sharding.ts_merge_rows(cands, t, sent_all)
    dist.barrier()
    if rank == 0:
        q.put(out)
    dist.destroy_process_group()


def test_trial_sharded_merge_rule_world_size_2_gloo():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ts_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    E, S = riskcases.synthetic(23, 6001, 16, 9)
    S[1000:1200] *= 3.0
    ref = cref.risk(E, S, 0.98)
    for label in ("auto", "tight"):
        top_l, top_t, exact = out[label]
        # rows the rule declares exact must BE exact (the rule may only err on the safe side)
        for p in np.flatnonzero(exact):
            assert np.array_equal(np.sort(top_t[p]), ref["breach"][p]), (label, p)
            assert top_l[p, -1] == ref["var_loss"][p] and top_t[p, -1] == ref["var_trial"][p]
    assert out["auto"][2].all()            # the automatic slack is enough here
    assert not out["tight"][2].all()       # m = ceil(t / world) is not: those rows would take the second round
