"""The oracle against streams EXECUTED from the reference's Java (tests/golden/codec_streams_golden.json).

This is what pins oracle/codec_oracle.c -- and through it the CUDA kernels -- to the reference: every stream in the
fixture was written by the reference's own FastPFOR / BinaryPacking / VariableByte / Composition / Delta /
Integrated* / IntCompressor classes, run from /root/reference by the javamini transpiler (no JVM in this image)."""
import base64
import hashlib
import json
import os

import numpy as np
import pytest

import cases
import goldenstreams as G
from oracle import cref

N_ORACLE_CODECS = 11


def test_fixture_covers_every_codec_and_the_multipage_inputs():
    g = G.gold()
    assert "javamini" in g["source"]
    ids = {int(k) for d in g["streams"].values() for k in d}
    assert {c | f for c in range(11) for f in (0, 0x100)} <= ids
    assert "unsorted1333333" in g["digests"] and "sorted2342351" in g["digests"]
    assert len(g["streams"]) > 200 and len(g["loader"]) == 10


def test_oracle_equals_the_reference_on_every_golden_stream():
    n = 0
    for label, vals, cid, want, dig in G.entries(N_ORACLE_CODECS, big=True):
        G.check(cref.compress(cid, vals), want, dig, f"{label} codec {cid:#x}")
        n += 1
    assert n > 4000


def test_oracle_decodes_the_reference_streams():
    """the other direction: streams the reference wrote, decoded by the oracle"""
    n = 0
    for label, vals, cid, want, _ in G.entries(N_ORACLE_CODECS, big=False):
        if want is None:
            continue
        back = cref.uncompress(cid, want, len(vals))
        assert np.array_equal(back, cases.to_i32(vals)), (label, cid)
        n += 1
    assert n > 3000


def test_pairs_on_which_the_reference_throws():
    """IntCompressor / IntegratedIntCompressor allocate input.length + 1024 output ints (IntCompressor.java:39): on
    incompressible input the reference itself dies with ArrayIndexOutOfBoundsException.  Recorded, not compared."""
    th = G.gold()["throws"]
    for label, d in th.items():
        for key, kind in d.items():
            assert (int(key) & 0xFF) in (4, 6, 7, 8) and kind == "ArrayIndexOutOfBoundsException", (label, key, kind)


@pytest.mark.parametrize("block", [256, 128])
def test_one_long_lived_fastpfor_object(block):
    """FastPFOR.java:143 clears dataPointers but never dataTobePacked: the padding of an exception stream's last
    packed group carries values of earlier calls and pages.  Observed from the reference here, not emulated."""
    entry = [e for e in G.gold()["reused"] if e["codec"] == (1 if block == 256 else 2)][0]
    seq = G.reused_inputs(entry["seed"])
    fp = cref.FastPFOR(block)
    fresh_differs = 0
    for v, rec in zip(seq, entry["streams"]):
        n = len(v)
        m = n - n % block
        # Composition(FastPFOR, VariableByte) on ONE FastPFOR object: FastPFOR part from the long-lived oracle object ...
        got = fp.compress(v[:m]) if m else np.zeros(1, np.int32)
        tail = cref.compress(cref.CODEC_VB, v[m:]) if n > m else np.zeros(0, np.int32)
        got = np.concatenate([got, tail])
        if isinstance(rec, str):
            G.check(got, G.words(rec), None, f"reused FastPFOR{block}, n={n}")
        else:
            G.check(got, None, tuple(rec), f"reused FastPFOR{block}, n={n}")
        fresh = cref.compress(1 if block == 256 else 2, v)
        fresh_differs += int(not np.array_equal(fresh, got))
    assert fresh_differs > 0, "the sequence was meant to expose the stale-buffer padding"


def test_loader_bytes_of_the_shipped_portfolios():
    """LoadPortfolios.getCompressedInstruments (executed from LoadPortfolios.java:608-623) on portfolios.json's lists."""
    g = G.gold()
    c1 = json.load(open(os.path.join(os.path.dirname(G.PATH), "c1_sample.json")))
    for p, rec in zip(c1["portfolios"], g["loader"]):
        blob = cref.get_compressed_instruments(cref.CODEC_BP32_VB, p["instruments"])
        assert len(p["instruments"]) == rec["n"] and len(blob) == rec["bytes"]
        assert hashlib.sha256(blob).hexdigest() == rec["sha256"]
        if rec["b64"]:
            assert blob == base64.b64decode(rec["b64"])
