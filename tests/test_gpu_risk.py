"""GPU parity for kernels (1)-(3) + the fused encode: CUDA path through the C ABI against the oracle.
Bars (BASELINE.json north_star): breach lists and encoded byte streams bit-exact, VaR trial index bit-exact,
portfolio values bit-identical (single fma chain on both sides), moments / CVaR within 1e-9 relative."""
import numpy as np
import pytest

import cases
import mcp_b200
import riskcases
from mcp_b200 import native as N
from oracle import cref

pytestmark = pytest.mark.gpu
RTOL = 1e-9  # north_star tolerance for FP64 moments


def check_against_oracle(ctx, E, S, alpha, cid=N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA, framing=N.FRAME_APP, rows=None):
    ctx.upload_exposures(E)
    ctx.upload_scenarios(S)
    ctx.run(alpha, cid, framing)
    r = ctx.fetch()
    P = E.shape[0]
    rows = range(P) if rows is None else rows
    o = cref.risk(E, S, alpha)
    assert r["tail"] == o["tail"]
    for p in rows:
        assert r["var_trial"][p] == o["var_trial"][p], p
        assert r["var_loss"][p] == o["var_loss"][p], p
        assert np.array_equal(r["breach"][p], o["breach"][p]), p
        blob = bytes(r["enc"][r["enc_off"][p]: r["enc_off"][p + 1]])
        ref = cref.get_compressed_instruments(cid, o["breach"][p]) if framing == N.FRAME_APP else cref.compress(cid, o["breach"][p]).tobytes()
        assert blob == ref, p
    rr = list(rows)
    np.testing.assert_allclose(r["mean"][rr], o["mean"][rr], rtol=RTOL, atol=0)
    np.testing.assert_allclose(r["variance"][rr], o["variance"][rr], rtol=RTOL, atol=0)
    np.testing.assert_allclose(r["cvar"][rr], o["cvar"][rr], rtol=RTOL, atol=0)
    return r


@pytest.mark.parametrize("shape", [(10, 10, 10), (70, 1000, 32), (65, 999, 64), (33, 257, 5), (3, 4097, 33), (129, 130, 96)])
def test_value_bit_identical(ctx, shape):
    P, Nn, F = shape
    E, S = riskcases.synthetic(P, Nn, F, 7)
    ctx.upload_exposures(E)
    ctx.upload_scenarios(S)
    V = ctx.value(0, P, Nn)
    assert np.array_equal(V.view(np.uint64), cref.value(E, S).view(np.uint64))
    V2 = ctx.value(P // 2, P, Nn)
    assert np.array_equal(V2, V[P // 2:])


@pytest.mark.parametrize("F", [16, 32, 48, 64])
def test_tensor_core_valuation_is_the_fma_chain(ctx, F):
    """The FP64 tensor-core kernel (mma.sync m8n8k4 / DMMA) must reproduce the k-ascending fma chain bit for bit.
    Inputs span 40 binades so that any other accumulation order or an unfused multiply would show up."""
    rng = np.random.default_rng(F)
    P, Nn = 300, 1203
    E = (rng.random((P, F)) - 0.5) * np.exp2(rng.integers(-20, 20, (P, F)).astype(np.float64))
    S = (rng.random((Nn, F)) - 0.5) * np.exp2(rng.integers(-20, 20, (Nn, F)).astype(np.float64))
    ctx.upload_exposures(E)
    ctx.upload_scenarios(S)
    ref = cref.value(E, S).view(np.uint64)
    ctx.set_option("value_mma", 1)
    assert np.array_equal(ctx.value(0, P, Nn).view(np.uint64), ref)
    assert np.array_equal(ctx.value(7, 270, Nn).view(np.uint64), ref[7:270])
    ctx.set_option("value_mma", 0)
    try:
        assert np.array_equal(ctx.value(0, P, Nn).view(np.uint64), ref)   # scalar-fma kernel
    finally:
        ctx.set_option("value_mma", 1)
    check_against_oracle(ctx, E, S, 0.97, rows=[0, 1, 150, 299])


@pytest.mark.parametrize("shape", [(10, 10, 10, 0.9), (40, 1000, 32, 0.99), (30, 999, 64, 0.95), (9, 257, 5, 0.5),
                                   (5, 64, 3, 1.0), (12, 20000, 32, 0.99), (3, 100000, 32, 0.99), (2, 50000, 7, 0.6)])
def test_run_matches_oracle(ctx, shape):
    P, Nn, F, alpha = shape
    E, S = riskcases.synthetic(P, Nn, F, 1234 + P)
    check_against_oracle(ctx, E, S, alpha)


def test_tile_overlap_on_two_streams_gives_identical_results(ctx):
    """overlap_tiles > 1 alternates portfolio tiles between two streams with separate candidate buffers."""
    E, S = riskcases.synthetic(2100, 9000, 32, 77)
    rows = list(range(0, 2100, 97)) + [2099]
    r1 = check_against_oracle(ctx, E, S, 0.99, rows=rows)
    ctx.set_option("overlap_tiles", 4)
    try:
        r4 = check_against_oracle(ctx, E, S, 0.99, rows=rows)
        ctx.set_option("optimistic", 0)
        r4g = check_against_oracle(ctx, E, S, 0.99, rows=rows)
    finally:
        ctx.set_option("overlap_tiles", 1)
        ctx.set_option("optimistic", 1)
    for k in ("var_loss", "var_trial", "breach", "enc_off", "enc", "cvar", "mean", "variance"):
        assert np.array_equal(r1[k], r4[k]) and np.array_equal(r1[k], r4g[k]), k


@pytest.mark.parametrize("opt", ["force_dense", "force_sort_finalize", "fused_moments", "reg_select", "optimistic", "small_select", "warp_select"])
def test_alternate_kernel_paths_agree(ctx, opt):
    """F = 32 normally takes the streaming kernels (two-step optimistic schedule, dense-step selection from
    registers) + bitmap compaction; the dense tile + single selection, the sort-based breach list, the fused
    per-trial moments, the shared-memory selection and the guaranteed schedule must give identical results."""
    default = 1 if opt in ("reg_select", "optimistic", "small_select", "warp_select") else 0
    ctx.set_option(opt, 1 - default)
    try:
        E, S = riskcases.synthetic(40, 9000, 32, 21)
        check_against_oracle(ctx, E, S, 0.99)
        E, S = riskcases.with_ties(8, 5000, 32, 22)
        check_against_oracle(ctx, E, S, 0.9)
        E, S = riskcases.synthetic(20, 60000, 32, 23)
        check_against_oracle(ctx, E, S, 0.995)
        E, S = riskcases.synthetic(300, 10000, 64, 24)           # C3 shape: tail of 101 -> the 64-thread selection kernels
        check_against_oracle(ctx, E, S, 0.99, rows=list(range(0, 300, 13)))
        E, S = riskcases.with_ties(6, 12000, 32, 25)             # ... with thousands of tied losses
        check_against_oracle(ctx, E, S, 0.995)
    finally:
        ctx.set_option(opt, default)


@pytest.mark.parametrize("F", [1, 3, 10, 17, 33, 47, 63])
def test_odd_factor_counts_take_the_padded_tensor_core_path(ctx, F):
    """F without a tensor-core kernel is padded with zero columns to 16/32/48/64 (exact: fma(0,0,acc) = acc); results
    must equal the oracle's and the dense scalar-fma path's, and the stored matrices stay unpadded."""
    E, S = riskcases.synthetic(70, 6000, F, 200 + F)
    r1 = check_against_oracle(ctx, E, S, 0.99)
    assert np.array_equal(ctx.download_exposures(70, F), E) and np.array_equal(ctx.download_scenarios(6000, F), S)
    assert np.array_equal(ctx.value(0, 3, 6000).view(np.uint64), cref.value(E[:3], S).view(np.uint64))
    ctx.set_option("force_dense", 1)
    try:
        r0 = check_against_oracle(ctx, E, S, 0.99)
    finally:
        ctx.set_option("force_dense", 0)
    for k in ("var_loss", "var_trial", "breach", "enc"):
        assert np.array_equal(r0[k], r1[k]), k


def test_streaming_path_many_portfolios_ragged(ctx):
    """F = 32 streaming path with a ragged last portfolio group and trial counts that are not multiples of 4/8/32."""
    for P, Nn, alpha, F in ((513, 4099, 0.99, 32), (1030, 2049, 0.9, 32), (70, 33333, 0.999, 32), (3, 7, 0.5, 32),
                            (600, 12345, 0.95, 32), (300, 9001, 0.99, 64), (257, 5003, 0.98, 16), (100, 20011, 0.995, 48)):
        E, S = riskcases.synthetic(P, Nn, F, 100 + P)
        check_against_oracle(ctx, E, S, alpha, rows=list(range(0, P, max(1, P // 37))) + [P - 1])


def test_large_tails_take_the_wide_selection_kernels(ctx):
    """Confidence levels with tails of thousands of trials: the optimistic sample is capped at 4 096 values, the sparse
    selection runs with 1 024-thread CTAs (one per SM), very large tails overflow into the global-memory passes."""
    E, S = riskcases.synthetic(24, 100000, 32, 91)
    for alpha in (0.95, 0.9, 0.6):
        check_against_oracle(ctx, E, S, alpha, rows=[0, 11, 23])
        assert ctx.run_stats()["fallback_rows"] == 0
    ctx.set_option("force_sort_finalize", 1)        # no bitmap over the trial range: the counting finalize (tails >= 1 024) ...
    try:
        r1 = check_against_oracle(ctx, E, S, 0.93, rows=[0, 11, 23])
        ctx.set_option("radix_finalize", 0)         # ... and the bitonic sort with 1 024-thread CTAs (tails >= 4 097)
        r0 = check_against_oracle(ctx, E, S, 0.93, rows=[0, 11, 23])
        for k in ("breach", "cvar", "enc"):
            assert np.array_equal(r0[k], r1[k]), k   # (CVaR bit for bit: same summation order)
    finally:
        ctx.set_option("force_sort_finalize", 0)
        ctx.set_option("radix_finalize", 1)


def test_counting_finalize_on_clustered_tails(ctx):
    """k_finalize_radix: buckets of trial >> shift, then a rank inside each bucket -- by shuffles up to 32 entries, through a
    warp-private bitmap beyond.  Losses that grow with the trial index put a row's whole tail into consecutive trials, i.e.
    into a few FULL buckets (300 000 trials: buckets of 128); a second shape has the tail in clusters of ~40."""
    Nn, F, P = 300_000, 32, 6
    S = np.zeros((Nn, F))
    S[:, 0] = -np.linspace(0.0, 1.0, Nn)                       # loss ascending in the trial index: the tail is the last t trials
    S[:, 1] = 1e-4 * np.cos(np.arange(Nn))
    E = np.abs(riskcases.synthetic(P, 1, F, 5)[0]) + 1.0
    rng = np.random.default_rng(17)
    S2 = rng.standard_normal((Nn, F)) * 0.01
    centres = rng.choice(Nn - 64, 200, replace=False)
    for c in centres:                                           # 200 clusters of 40 consecutive bad trials
        S2[c:c + 40, 0] -= 5.0 + rng.random(40)
    ctx.set_option("force_sort_finalize", 1)
    try:
        for SS, alpha in ((S, 0.99), (S2, 0.975), (S, 0.95)):   # tails of 3 000, 7 500 and 15 000 (the last one: shared memory for 15 000
            r1 = check_against_oracle(ctx, E, SS, alpha)        #  entries exceeds the kernel's budget -> bitonic sort in global memory)
            ctx.set_option("radix_finalize", 0)
            r0 = check_against_oracle(ctx, E, SS, alpha, rows=[0, P - 1])
            ctx.set_option("radix_finalize", 1)
            for k in ("breach", "enc"):
                assert np.array_equal(r0[k], r1[k]), (k, alpha)
            np.testing.assert_allclose(r0["cvar"], r1["cvar"], rtol=1e-13, atol=0)  # (the 256-thread sort kernel sums in another tree)
    finally:
        ctx.set_option("force_sort_finalize", 0)
        ctx.set_option("radix_finalize", 1)


def test_thin_samples_take_a_middle_step(ctx):
    """Many trials per tail element (t n0 / N < 16): the optimistic schedule inserts a middle step that tightens the
    bound before the long final step; adversarial orders still fall back exactly."""
    E, S = riskcases.synthetic(10, 1_000_000, 32, 93)
    check_against_oracle(ctx, E, S, 0.999)                      # t = 1001 of 10^6: ts(4096) = 4.1 -> middle step at 64 000
    assert ctx.run_stats()["fallback_rows"] == 0
    Nn = 600_000
    S2 = np.zeros((Nn, 32))
    S2[:, 0] = np.linspace(0.0, 1.0, Nn)                         # descending losses: every row loses the bet at the middle step
    S2[:, 1] = 1e-3 * np.cos(np.arange(Nn))
    E2 = np.abs(riskcases.synthetic(6, 1, 32, 3)[0]) + 1.0
    check_against_oracle(ctx, E2, S2, 0.999)
    assert ctx.run_stats()["fallback_rows"] == 6


def test_ties_break_by_trial_index(ctx):
    E, S = riskcases.with_ties(20, 3000, 4, 9)
    check_against_oracle(ctx, E, S, 0.9)
    E, S = riskcases.with_ties(4, 40000, 3, 10)   # thousands of identical losses: deep radix passes
    check_against_oracle(ctx, E, S, 0.75)
    E, S = riskcases.with_ties(70, 40000, 32, 11)  # same through the F = 32 streaming path
    check_against_oracle(ctx, E, S, 0.75, rows=[0, 1, 35, 69])


@pytest.mark.parametrize("warp", [1, 0], ids=["warp-per-row", "cta-per-row"])
def test_small_tails_on_hostile_rows(ctx, warp):
    """Tails of ~100 trials (the C3 shape) take the warp-per-row selection of the sparse step.  Rows it cannot hold -- trial
    orders that make every trial a candidate -- are handed back in a list and redone by the CTA-per-row kernel; rows whose
    optimistic bound was too tight are reported and redone with the guaranteed schedule; heavy ties are broken by trial
    index.  All in one run, next to ordinary rows, against the oracle; identical with the warp kernel switched off."""
    F, Nn, P = 64, 10000, 96
    E, S = riskcases.synthetic(P, Nn, F, 77)
    S[:, 0] = np.linspace(0.0, 1.0, Nn)            # a trend in factor 0 ...
    E[:, 0] = 0.0
    E[0::4, 0] = 5.0e4                             # ... that dominates these rows: losses descend with the trial index (the bet fails)
    E[1::4, 0] = -5.0e4                            # ... losses ascend: every trial beats the sample's bound (the row does not fit a warp)
    Et, St = riskcases.with_ties(8, Nn, F, 78)     # heavily tied losses (integer inputs), in a run of their own
    ctx.set_option("warp_select", warp)
    try:
        check_against_oracle(ctx, E, S, 0.99, rows=list(range(0, P, 5)) + [1, 2, 3, P - 1])
        assert ctx.run_stats()["fallback_rows"] == len(range(0, P, 4))
        check_against_oracle(ctx, Et, St, 0.99)
    finally:
        ctx.set_option("warp_select", 1)


def test_degenerate_rows(ctx):
    E, S = riskcases.synthetic(6, 500, 8, 5)
    E[0] = 0.0            # all losses +0.0: the tail is the last t trials
    E[1] = -E[2]          # mirrored
    S[100] = S[200]       # identical scenarios -> identical losses
    check_against_oracle(ctx, E, S, 0.97)


def test_adversarial_trial_order(ctx):
    """Losses that increase with the trial index (the worst case for any streaming threshold filter)."""
    for F in (4, 32):
        _adversarial(ctx, F)


def _adversarial(ctx, F):
    Nn = 30000
    S = np.zeros((Nn, F))
    S[:, 0] = -np.linspace(0.0, 1.0, Nn)
    S[:, 1] = 1e-3 * np.cos(np.arange(Nn))
    E = np.abs(riskcases.synthetic(8, 1, F, 3)[0]) + 1.0
    check_against_oracle(ctx, E, S, 0.99)
    check_against_oracle(ctx, E, S[::-1].copy(), 0.99)


def test_optimistic_bound_falls_back_exactly(ctx):
    """The two-step schedule bets on an optimistic filter bound (r-th largest loss of the first ~4t trials).  Trial
    orders that break the bet must be detected per row and redone exactly: losses DEscending in the trial index make
    the bound too tight for every row with positive exposure, ascending ones make it uselessly loose (every trial a
    candidate); rows with mirrored exposure see the opposite order in the same run."""
    F, Nn, P = 32, 60000, 300
    S = np.zeros((Nn, F))
    S[:, 0] = np.linspace(0.0, 1.0, Nn)           # V grows with the trial index for E[:,0] > 0: losses descend
    S[:, 1] = 1e-3 * np.cos(np.arange(Nn))
    E = np.abs(riskcases.synthetic(P, 1, F, 3)[0]) + 1.0
    E[::3] *= -1.0                                 # every third row: losses ascend instead
    rows = list(range(0, P, 7)) + [P - 1]
    check_against_oracle(ctx, E, S, 0.99, rows=rows)
    fb = ctx.run_stats()["fallback_rows"]
    assert fb == P - len(range(0, P, 3)), fb       # exactly the descending-loss rows were redone
    # exchangeable order: the bet holds, nothing is redone, and the guaranteed schedule gives the same bytes
    E2, S2 = riskcases.synthetic(P, Nn, F, 41)
    r1 = check_against_oracle(ctx, E2, S2, 0.99, rows=rows)
    assert ctx.run_stats()["fallback_rows"] == 0
    ctx.set_option("optimistic", 0)
    try:
        r0 = check_against_oracle(ctx, E2, S2, 0.99, rows=rows)
        assert ctx.run_stats()["fallback_rows"] == 0
    finally:
        ctx.set_option("optimistic", 1)
    for k in ("var_loss", "var_trial", "breach", "enc_off", "enc", "cvar"):
        assert np.array_equal(r0[k], r1[k]), k


def test_ingest_builds_the_exposure_matrix_on_the_device(ctx):
    """Positions (instrument, weight) x loading table -> E on the GPU (mcp_build_exposures), bit-identical to the
    oracle's file-order fma chain: the shipped sample (3 204 positions, duplicates, unsorted ids) and a synthetic
    batch with empty portfolios, F up to 70, and portfolios longer than one warp round."""
    d, E, S = riskcases.c1_inputs()
    inst, w, off = riskcases.c1_positions(d)
    B = cref.gen_normal(10000, 10, 0x5EED0001, 3, 1e-2, 0.0)
    ctx.build_exposures(inst, w, off, B)
    assert np.array_equal(ctx.download_exposures(10, 10).view(np.uint64), E.view(np.uint64))
    ctx.upload_scenarios(S)
    ctx.run(0.9)                                    # ... and straight into the path
    o = cref.risk(E, S, 0.9)
    r = ctx.fetch()
    assert np.array_equal(r["breach"], o["breach"]) and np.array_equal(r["var_loss"], o["var_loss"])
    rng = np.random.default_rng(5)
    for F in (1, 32, 64, 70):
        lens = rng.integers(0, 200, 300)
        lens[[3, 77]] = 0
        off = np.zeros(301, np.int64)
        np.cumsum(lens, out=off[1:])
        inst = rng.integers(0, 500, off[-1]).astype(np.int32)
        w = rng.normal(size=off[-1]) * np.exp2(rng.integers(-10, 10, off[-1]))
        B = rng.normal(size=(500, F))
        ctx.build_exposures(inst, w, off, B)
        assert np.array_equal(ctx.download_exposures(300, F).view(np.uint64), cref.build_exposures(inst, w, off, B).view(np.uint64)), F
    with pytest.raises(Exception):
        ctx.build_exposures(np.array([500], np.int32), np.ones(1), np.array([0, 1], np.int64), np.ones((500, 4)))


def test_simulate_portfolios_entry_point_on_the_shipped_sample(ctx, tmp_path):
    """portfolios.json x macro.json -> SimulatePortfolios -> one sink document per portfolio (the reference creates the
    /apps/simulation_results table, ManageMapRDBTable.java:15, and never writes it)."""
    import json
    import mcp_b200
    d, E, S = riskcases.c1_inputs()
    pf, mf = tmp_path / "portfolios.json", tmp_path / "macro.json"
    with open(pf, "w") as fh:
        for p in d["portfolios"]:
            fh.write(json.dumps({"id": p["id"], "instruments": [{"instrument": i, "weight": w} for i, w in zip(p["instruments"], p["weights"])]}) + "\n")
    with open(mf, "w") as fh:
        for k, row in enumerate(d["macro"]):
            fh.write(json.dumps({"date_offset": k, **{f"m{j}": v for j, v in enumerate(row)}}) + "\n")
    B = cref.gen_normal(10000, 10, 0x5EED0001, 3, 1e-2, 0.0)
    docs = mcp_b200.SimulatePortfolios(ctx).run(str(pf), str(mf), B, 0.9)
    o = cref.risk(E, S, 0.9)
    cid = N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA
    assert len(docs) == 10
    for p, doc in enumerate(docs):
        assert doc["var"]["loss"] == o["var_loss"][p] and doc["var"]["trial"] == o["var_trial"][p]
        b = doc["breaches"]
        lst = o["breach"][p]
        if b["algo"] == "none":
            assert b["compressed"] == mcp_b200.LoadPortfolios.integersToBytes(lst)
        else:
            assert np.array_equal(cref.get_decompressed_instruments(cid, b["compressed"], o["tail"]), lst)


def test_c1_shipped_sample(ctx):
    """BASELINE config C1: portfolios.json x macro.json (fixture tests/golden/c1_sample.json)."""
    d, E, S = riskcases.c1_inputs()
    for cid in (N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA, N.CODEC_BP32_VB):
        check_against_oracle(ctx, E, S, 0.9, cid)


@pytest.mark.parametrize("cid", [N.CODEC_BP32_VB, N.CODEC_BP32_VB | N.FLAG_DELTA, N.CODEC_FASTPFOR128_VB | N.FLAG_DELTA,
                                 N.CODEC_IBP32_IVB, N.CODEC_SK_IBP32_IVB])
def test_every_breach_codec(ctx, cid):
    E, S = riskcases.synthetic(16, 30000, 32, 77)
    check_against_oracle(ctx, E, S, 0.99, cid, N.FRAME_APP)
    check_against_oracle(ctx, E, S, 0.99, cid, N.FRAME_WORDS, rows=[0, 15])


def test_simulate_one_call(ctx):
    E, S = riskcases.synthetic(25, 5000, 32, 8)
    r = ctx.simulate(E, S, 0.99)
    o = cref.risk(E, S, 0.99)
    assert np.array_equal(r["breach"], o["breach"]) and np.array_equal(r["var_trial"], o["var_trial"])
    cid = N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA
    for p in (0, 24):
        assert bytes(r["enc"][r["enc_off"][p]: r["enc_off"][p + 1]]) == cref.get_compressed_instruments(cid, o["breach"][p])


@pytest.mark.parametrize("turns", [-1, 0, 1], ids=["by-size", "free", "turns"])
def test_simulate_pipeline_two_calls_in_flight(turns):
    """SimulatePipeline: two contexts on one GPU, one host thread each -- steps submitted back to back come back as the
    steps run one at a time would have (and as the oracle says), whatever the interleaving of their copies and kernels:
    left free, or with the calls taking turns on the SMs (option simulate_turns; by upload size unless set)."""
    steps = [riskcases.synthetic(20 + i, 40000, 32, 100 + i) for i in range(5)]
    with mcp_b200.SimulatePipeline(device=0, depth=2) as pipe:
        for c in pipe.contexts():
            c.set_option("simulate_turns", turns)
        futures = [pipe.submit(E, S, 0.99) for E, S in steps]
        results = [f.result() for f in futures]
    cid = N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA
    for (E, S), r in zip(steps, results):
        o = cref.risk(E, S, 0.99)
        assert np.array_equal(r["breach"], o["breach"]) and np.array_equal(r["var_trial"], o["var_trial"])
        assert np.array_equal(r["var_loss"], o["var_loss"])
        np.testing.assert_allclose(r["cvar"], o["cvar"], rtol=RTOL)
        for p in (0, E.shape[0] - 1):
            assert bytes(r["enc"][r["enc_off"][p]: r["enc_off"][p + 1]]) == cref.get_compressed_instruments(cid, o["breach"][p])


def test_simulate_with_the_scenario_tail_uploaded_in_two_parts(ctx):
    """mcp_simulate uploads the scenario tail in two parts and launches the sparse valuation in two halves, the first as
    soon as its rows are there; results are those of the single upload and of the oracle."""
    E, S = riskcases.synthetic(300, 40000, 32, 12)
    o = cref.risk(E, S, 0.99)
    got = {}
    try:
        for split in (1, 0):
            ctx.set_option("upload_split", split)
            r = ctx.simulate(E, S, 0.99)
            assert np.array_equal(r["breach"], o["breach"]) and np.array_equal(r["var_trial"], o["var_trial"]), split
            assert np.array_equal(r["var_loss"], o["var_loss"]), split
            np.testing.assert_allclose(r["cvar"], o["cvar"], rtol=RTOL, atol=0)
            got[split] = r
    finally:
        ctx.set_option("upload_split", 1)
    for k in ("enc", "enc_off", "cvar", "mean", "variance"):
        assert np.array_equal(got[0][k], got[1][k]), k


def test_simulate_when_the_optimistic_bet_fails(ctx):
    """mcp_simulate sends the statistics and the raw lists to the host under the (speculative) encode; rows that lost the
    optimistic bet are redone afterwards, and what the caller finally holds must be the repaired values."""
    F, Nn, P = 32, 60000, 60
    S = np.zeros((Nn, F))
    S[:, 0] = np.linspace(0.0, 1.0, Nn)           # losses descend with the trial index for positive exposures
    S[:, 1] = 1e-3 * np.cos(np.arange(Nn))
    E = np.abs(riskcases.synthetic(P, 1, F, 3)[0]) + 1.0
    E[::3] *= -1.0
    r = ctx.simulate(E, S, 0.99)
    assert ctx.run_stats()["fallback_rows"] == P - len(range(0, P, 3))
    o = cref.risk(E, S, 0.99)
    assert np.array_equal(r["var_trial"], o["var_trial"]) and np.array_equal(r["var_loss"], o["var_loss"])
    assert np.array_equal(r["breach"], o["breach"])
    np.testing.assert_allclose(r["cvar"], o["cvar"], rtol=RTOL, atol=0)
    np.testing.assert_allclose(r["mean"], o["mean"], rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(r["variance"], o["variance"], rtol=RTOL, atol=0)
    cid = N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA
    for p in (0, 1, P - 1):
        assert bytes(r["enc"][r["enc_off"][p]: r["enc_off"][p + 1]]) == cref.get_compressed_instruments(cid, o["breach"][p])


def test_device_generators_match_the_oracle(ctx):
    ctx.generate_exposures(100, 32, 0x5EED0002, 100.0, 0.0)
    ctx.generate_scenarios(1000, 32, 0x5EED0002, 1.0, 0.05)
    assert np.array_equal(ctx.download_exposures(100, 32), cref.gen_normal(100, 32, 0x5EED0002, 1, 100.0, 0.0))
    assert np.array_equal(ctx.download_scenarios(1000, 32), cref.gen_normal(1000, 32, 0x5EED0002, 2, 1.0, 0.05))
    ctx.generate_scenarios(10, 32, 0x5EED0002, 1.0, 0.05, row0=500)
    assert np.array_equal(ctx.download_scenarios(10, 32), cref.gen_normal(10, 32, 0x5EED0002, 2, 1.0, 0.05, row0=500))


def test_full_size_c2_properties(ctx):
    """BASELINE config C2 at full size (10 000 x 100 000 x 32) on device-generated inputs: spot rows against the
    oracle plus size-independent properties on every row."""
    P, Nn, F, alpha = 10000, 100000, 32, 0.99
    ctx.generate_exposures(P, F, 0x5EED0002, 100.0, 0.0)
    ctx.generate_scenarios(Nn, F, 0x5EED0002, 1.0, 0.05)
    cid = N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA
    ctx.run(alpha, cid, N.FRAME_APP)
    r = ctx.fetch()
    t = r["tail"]
    assert t == 1001
    assert ctx.run_stats()["fallback_rows"] == 0  # the optimistic bound held for every portfolio
    b = r["breach"]
    assert np.all(np.diff(b, axis=1) > 0) and b.min() >= 0 and b.max() < Nn          # sorted, distinct, in range
    assert np.all((b == r["var_trial"][:, None]).sum(axis=1) == 1)                     # the VaR trial is in its list
    assert np.all(r["cvar"] >= r["var_loss"]) and np.all(r["variance"] > 0)
    # decode the encoded lists back on the GPU: encode -> decode round trip over all 10 000 lists
    off = np.arange(P + 1, dtype=np.int64) * t
    back = ctx.codec_decode(cid, N.FRAME_APP, r["enc"], r["enc_off"], off)
    assert np.array_equal(back.reshape(P, t), b)
    # oracle on a handful of rows (host-generated inputs are bit-identical to the device generator)
    rows = [0, 1, 4999, 9999]
    E = cref.gen_normal(P, F, 0x5EED0002, 1, 100.0, 0.0)
    S = cref.gen_normal(Nn, F, 0x5EED0002, 2, 1.0, 0.05)
    for p in rows:
        o = cref.risk(E, S, alpha, p, p + 1)
        assert r["var_trial"][p] == o["var_trial"][0] and r["var_loss"][p] == o["var_loss"][0]
        assert np.array_equal(b[p], o["breach"][0])
        assert bytes(r["enc"][r["enc_off"][p]: r["enc_off"][p + 1]]) == cref.get_compressed_instruments(cid, o["breach"][0])
        np.testing.assert_allclose([r["mean"][p], r["variance"][p], r["cvar"][p]],
                                   [o["mean"][0], o["variance"][0], o["cvar"][0]], rtol=RTOL, atol=0)


def test_sharded_assembly_on_gpu(ctx):
    """The portfolio-sharding host logic with the real CUDA compute (one process; two simulated shards)."""
    from mcp_b200 import sharding
    E, S = riskcases.synthetic(70, 6000, 32, 31)
    cid = N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA
    parts = []
    for rank in range(2):
        p0, p1 = sharding.portfolio_shard(70, rank, 2)
        r = ctx.simulate(E[p0:p1], S, 0.99, cid, N.FRAME_APP)
        r["range"] = (p0, p1)
        parts.append(r)
    whole = sharding.assemble(parts, 70)
    one = ctx.simulate(E, S, 0.99, cid, N.FRAME_APP)
    for k in ("var_loss", "var_trial", "breach", "enc_off", "enc", "cvar"):
        assert np.array_equal(whole[k], one[k]), k
    o = cref.risk(E, S, 0.99)
    assert np.array_equal(whole["breach"], o["breach"])


def test_full_size_c3_properties(ctx):
    """BASELINE config C3 at full size on ONE GPU (1 000 000 portfolios x 10 000 trials x 64 factors, the N = 1 point of the
    strong-scaling curve; the 125 000-row k_select_t<64> path runs eight portfolio tiles): spot rows against the oracle,
    size-independent properties on every row, and the flat codec kernels' encode -> decode round trip over all 10^6 lists."""
    P, Nn, F, alpha, seed = 1_000_000, 10_000, 64, 0.99, 0x5EED0003
    ctx.generate_exposures(P, F, seed, 100.0, 0.0)
    ctx.generate_scenarios(Nn, F, seed, 1.0, 0.05)
    cid = N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA
    ctx.run(alpha, cid, N.FRAME_APP)
    r = ctx.fetch()
    t = r["tail"]
    assert t == 101
    assert ctx.run_stats()["fallback_rows"] == 0
    b = r["breach"]
    assert b.shape == (P, t)
    assert np.all(np.diff(b, axis=1) > 0) and b.min() >= 0 and b.max() < Nn
    assert np.all((b == r["var_trial"][:, None]).sum(axis=1) == 1)
    assert np.all(r["cvar"] >= r["var_loss"]) and np.all(r["variance"] > 0)
    assert r["enc_off"][0] == 0 and r["enc_off"][-1] == r["enc"].size and np.all(np.diff(r["enc_off"]) > 0)
    off = np.arange(P + 1, dtype=np.int64) * t
    back = ctx.codec_decode(cid, N.FRAME_APP, r["enc"], r["enc_off"], off)
    assert np.array_equal(back.reshape(P, t), b)
    S = cref.gen_normal(Nn, F, seed, 2, 1.0, 0.05)
    for p in (0, 124_999, 125_000, 500_001, 999_999):  # first / last rows of portfolio tiles and of the run
        E1 = cref.gen_normal(1, F, seed, 1, 100.0, 0.0, row0=p)
        o = cref.risk(E1, S, alpha)
        assert r["var_trial"][p] == o["var_trial"][0] and r["var_loss"][p] == o["var_loss"][0]
        assert np.array_equal(b[p], o["breach"][0])
        assert bytes(r["enc"][r["enc_off"][p]: r["enc_off"][p + 1]]) == cref.get_compressed_instruments(cid, o["breach"][0])
        np.testing.assert_allclose([r["mean"][p], r["variance"][p], r["cvar"][p]],
                                   [o["mean"][0], o["variance"][0], o["cvar"][0]], rtol=RTOL, atol=0)
