"""Trial-sharded mode (BASELINE config C5) on ONE GPU: the ranks are emulated as separate contexts whose exchange
blocks are concatenated by the test (the phase calls mcp_ts_* exist exactly for callers with their own transport),
so every kernel of the mode runs; the NCCL all-gather itself is covered by tests/test_gpu_nccl.py on 2 GPUs.
Results must equal the oracle's on the UNsharded problem: trial indices, breach lists and bytes bit-exact."""
import numpy as np
import pytest

import mcp_b200
import riskcases
from mcp_b200 import native as N
from mcp_b200 import sharding
from oracle import cref

pytestmark = pytest.mark.gpu
RTOL = 1e-9
CID = N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA


def emulate(E, S, alpha, world, cid=CID, framing=N.FRAME_APP, slack=None):
    P, Nn = E.shape[0], S.shape[0]
    ctxs = [mcp_b200.Context(0) for _ in range(world)]
    bufs = []
    try:
        blocks = []
        for g, c in enumerate(ctxs):
            if slack is not None:
                c.set_option("ts_slack", slack)
            n0, n1 = sharding.trial_shard(Nn, g, world)
            c.upload_exposures(E)
            c.upload_scenarios(S[n0:n1])
            ptr, nb = c.ts_local(alpha, Nn, n0, world)
            b = np.empty(nb, np.uint8)
            c.d2h(b, ptr)
            blocks.append(b)
        assert len({b.size for b in blocks}) == 1
        gathered = np.concatenate(blocks)
        fblocks = []
        for g, c in enumerate(ctxs):
            d = c.dev_alloc(gathered.size)
            bufs.append((c, d))
            c.h2d(d, gathered)
            ptr, nb = c.ts_merge(d, g)                    # each rank merges and tests its own portfolio slice
            b = np.empty(nb, np.uint8)
            c.d2h(b, ptr)
            fblocks.append(b)
        allflags = np.concatenate(fblocks)                # the second (tiny) exchange: one byte per portfolio
        flagged = []
        for c in ctxs:
            d = c.dev_alloc(allflags.size)
            bufs.append((c, d))
            c.h2d(d, allflags)
            flagged.append(c.ts_flags(d))
        assert len(set(flagged)) == 1, flagged            # every rank sees the same list
        if flagged[0] > 0:
            blocks2 = []
            for c in ctxs:
                ptr, nb = c.ts_local2()
                b = np.empty(nb, np.uint8)
                c.d2h(b, ptr)
                blocks2.append(b)
            g2 = np.concatenate(blocks2)
            for c in ctxs:
                d = c.dev_alloc(g2.size)
                bufs.append((c, d))
                c.h2d(d, g2)
                c.ts_merge2(d)
        parts = []
        for g, c in enumerate(ctxs):
            c.ts_finish(g, world, cid, framing)
            r = c.fetch()
            r["range"] = sharding.portfolio_shard(P, g, world)
            assert r["mean"].shape[0] == r["range"][1] - r["range"][0]
            parts.append(r)
        return sharding.assemble(parts, P), flagged[0]
    finally:
        for c, d in bufs:
            c.dev_free(d)
        for c in ctxs:
            c.close()


def check(E, S, alpha, world, cid=CID, framing=N.FRAME_APP, slack=None):
    got, flagged = emulate(E, S, alpha, world, cid, framing, slack)
    o = cref.risk(E, S, alpha)
    assert got["tail"] == o["tail"]
    assert np.array_equal(got["var_trial"], o["var_trial"])
    assert np.array_equal(got["var_loss"], o["var_loss"])
    assert np.array_equal(got["breach"], o["breach"])
    for p in range(E.shape[0]):
        blob = bytes(got["enc"][got["enc_off"][p]: got["enc_off"][p + 1]])
        ref = cref.get_compressed_instruments(cid, o["breach"][p]) if framing == N.FRAME_APP else cref.compress(cid, o["breach"][p]).tobytes()
        assert blob == ref, p
    np.testing.assert_allclose(got["mean"], o["mean"], rtol=RTOL, atol=1e-9 * np.abs(o["mean"]).max())
    np.testing.assert_allclose(got["variance"], o["variance"], rtol=RTOL, atol=0)
    np.testing.assert_allclose(got["cvar"], o["cvar"], rtol=RTOL, atol=0)
    return flagged


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_trial_sharded_matches_unsharded_oracle(world):
    E, S = riskcases.synthetic(67, 30001, 32, 50 + world)      # ragged: 30001 trials, 67 portfolios
    assert check(E, S, 0.99, world) == 0                       # exchangeable trials: one round is enough


def test_second_round_when_the_first_is_not_enough():
    """m forced down to ceil(t / world): most rows cannot be proven exact from the first round and take the second."""
    E, S = riskcases.synthetic(40, 24000, 32, 61)
    flagged = check(E, S, 0.99, 4, slack=0)
    assert flagged > 0


def test_adversarial_orders_put_the_whole_tail_on_one_rank():
    F, Nn, P = 32, 40000, 24
    S = np.zeros((Nn, F))
    S[:, 0] = -np.linspace(0.0, 1.0, Nn)           # losses grow with the trial index: the tail is the last rank's
    S[:, 1] = 1e-3 * np.cos(np.arange(Nn))
    E = np.abs(riskcases.synthetic(P, 1, F, 3)[0]) + 1.0
    E[::2] *= -1.0                                 # ... or the first rank's
    assert check(E, S, 0.99, 4) == P               # every row needs the second round, and comes out exact
    assert check(E, S, 0.99, 2) == P
    assert check(E, S, 0.999, 2) == 0              # tail of 41: every rank sends its full local tail, one round is exact


def test_ties_and_other_codecs_and_factor_counts():
    E, S = riskcases.with_ties(12, 20000, 32, 14)
    check(E, S, 0.9, 3)
    E, S = riskcases.synthetic(20, 16000, 64, 15)
    check(E, S, 0.98, 2, cid=N.CODEC_BP32_VB | N.FLAG_DELTA, framing=N.FRAME_WORDS)
    E, S = riskcases.synthetic(9, 9000, 16, 16)
    check(E, S, 0.995, 5)


def test_argument_checks(ctx):
    E, S = riskcases.synthetic(4, 1000, 32, 1)
    ctx.upload_exposures(E)
    ctx.upload_scenarios(S)
    with pytest.raises(mcp_b200.McpError):
        ctx.ts_local(0.99, 1500, 0, 2)       # 1000 local rows > ceil(1500 / 2): shards must be balanced
    with pytest.raises(mcp_b200.McpError):
        ctx.ts_local(0.99, 2000, 1500, 2)    # range past the end
    with pytest.raises(mcp_b200.McpError):
        ctx.run_trial_sharded(0.99, 2000, 0)  # no communicator
    assert N.lib().mcp_ts_candidates(10_000_000, 0.999, 8) == sharding.ts_candidates(10_000_000, 10001, 8)
