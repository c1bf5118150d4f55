"""Property tests of the CPU oracle (hypothesis): the checker has to be right for ANY list, not only for the shapes the
reference's JUnit tests use.  Each property is one the Java library guarantees by construction."""
import numpy as np
from hypothesis import given, settings, strategies as st

import cases
from oracle import cref

INTS = st.lists(st.integers(-2**31, 2**31 - 1), min_size=0, max_size=700)
SMALL = st.lists(st.integers(0, 1 << 14), min_size=0, max_size=1500)
CODECS = st.sampled_from(cases.ALL_CODECS)
DELTA = st.sampled_from([0, cases.DELTA])


def _ok(codec, delta):
    return not (delta and codec in (cref.CODEC_IBP32_IVB, cref.CODEC_SK_IBP32_IVB))


@settings(max_examples=300, deadline=None)
@given(vals=st.one_of(INTS, SMALL), codec=CODECS, delta=DELTA)
def test_round_trip_any_list(vals, codec, delta):
    """uncompress(compress(x)) == x for every codec id, with Java's 32-bit wrap-around for Delta."""
    if not _ok(codec, delta):
        return
    v = np.array(vals, np.int64).astype(np.int32)
    w = cref.compress(codec | delta, v)
    assert np.array_equal(cref.uncompress(codec | delta, w, v.size), v)
    # the loader's framing: big-endian bytes, one trailing zero int (LoadPortfolios.java:622)
    b = cref.get_compressed_instruments(codec | delta, v)
    assert len(b) == 4 * (w.size + 1) and b[-4:] == b"\0\0\0\0"
    assert bytes(b[:-4]) == w.astype(">i4").tobytes()


@settings(max_examples=200, deadline=None)
@given(vals=st.one_of(INTS, SMALL))
def test_skippable_streams_differ_from_compositions_in_word_zero_only(vals):
    v = np.array(vals, np.int64).astype(np.int32)
    for sk, comp in ((cref.CODEC_SK_BP32_VB, cref.CODEC_BP32_VB), (cref.CODEC_SK_FASTPFOR256_VB, cref.CODEC_FASTPFOR256_VB),
                     (cref.CODEC_SK_FASTPFOR128_VB, cref.CODEC_FASTPFOR128_VB)):
        a, b = cref.compress(sk, v), cref.compress(comp, v)
        assert a[0] == v.size
        if v.size:
            assert np.array_equal(a[1:], b[1:])
        else:
            assert a.size == 1 and b.size == 0


@settings(max_examples=200, deadline=None)
@given(vals=SMALL)
def test_delta_flag_is_delta_then_codec(vals):
    """MCP_FLAG_DELTA = Delta.delta on a copy, then the codec (Delta.java:24-28): same words as encoding the gaps."""
    v = np.array(vals, np.int64).astype(np.int32)
    gaps = v.copy()
    if v.size:
        gaps[1:] = (v[1:].astype(np.int64) - v[:-1].astype(np.int64)).astype(np.int32)
    for codec in (cref.CODEC_BP32_VB, cref.CODEC_FASTPFOR256_VB, cref.CODEC_VB, cref.CODEC_SK_BP32_VB):
        assert np.array_equal(cref.compress(codec | cases.DELTA, v), cref.compress(codec, gaps)), codec


@settings(max_examples=150, deadline=None)
@given(vals=st.lists(st.integers(0, 2**31 - 1), min_size=0, max_size=600))
def test_integrated_vbyte_is_vbyte_of_the_gaps_for_sorted_lists(vals):
    """IntegratedComposition's blocks carry differences, its VariableByte tail restarts at 0 (IntegratedVariableByte.java:40):
    for a list shorter than one block the stream is Composition's 0 marker + VariableByte of the gaps from 0."""
    v = np.sort(np.array(vals, np.int64)).astype(np.int32)[:31]
    a = cref.compress(cref.CODEC_IBP32_IVB, v)
    b = cref.compress(cref.CODEC_VB | cases.DELTA, v)
    if v.size:
        assert a[0] == 0 and np.array_equal(a[1:], b)


# ---------------------------------------------------------------- risk oracle
from test_oracle_risk import brute  # noqa: E402  (the lexsort restatement of the specification)


@settings(max_examples=120, deadline=None)
@given(P=st.integers(1, 4), N=st.integers(1, 300), F=st.integers(1, 9), levels=st.integers(1, 5), seed=st.integers(0, 2**31 - 1),
       alpha=st.sampled_from([0.5, 0.9, 0.95, 0.99, 1.0]))
def test_risk_oracle_against_lexsort_with_heavy_ties(P, N, F, levels, seed, alpha):
    """Small integer exposures and scenario values make most losses collide exactly: the VaR order statistic, its trial
    index and the breach list must still be the lexicographic (loss, trial) ones of the specification."""
    rng = np.random.default_rng(seed)
    E = rng.integers(-levels, levels + 1, (P, F)).astype(np.float64)
    S = rng.integers(-levels, levels + 1, (N, F)).astype(np.float64)
    V, ref, t = brute(E, S, alpha)
    r = cref.risk(E, S, alpha)
    assert r["tail"] == t and 1 <= t <= N
    assert np.array_equal(r["var_trial"], ref["var_trial"]) and np.array_equal(r["var_loss"], ref["var_loss"])
    for p in range(P):
        b = r["breach"][p]
        assert np.array_equal(b, ref["breach"][p]) and np.all(np.diff(b) > 0) and b.size == t
        assert r["var_trial"][p] in b
    assert np.allclose(r["cvar"], ref["cvar"], rtol=1e-12, atol=1e-12)
    assert np.all(np.asarray(r["cvar"]) >= np.asarray(r["var_loss"]) - 1e-12)   # the tail mean is not below its smallest element
