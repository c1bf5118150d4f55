"""Reader of tests/golden/codec_streams_golden.json -- streams produced by EXECUTING the reference's Java sources
(tests/golden/make_codec_streams_golden.py, javamini).  Shared by the oracle test and the GPU parity test."""
from __future__ import annotations

import base64
import hashlib
import json
import os
import zlib

import numpy as np

import cases

PATH = os.path.join(os.path.dirname(__file__), "golden", "codec_streams_golden.json")
_gold = None


def gold() -> dict:
    global _gold
    if _gold is None:
        _gold = json.load(open(PATH))
    return _gold


def words(b64: str) -> np.ndarray:
    return np.frombuffer(zlib.decompress(base64.b64decode(b64)), dtype="<u4").view(np.int32)


def entries(max_codec: int, big: bool = True):
    """-> (label, input values, codec id with flag, expected int32 words | None, (nwords, sha256) | None)"""
    g = gold()
    for label, vals in cases.reference_matrix(big=big):
        for src, full in ((g["streams"], True), (g["digests"], False)):
            for key, v in src.get(label, {}).items():
                cid = int(key)
                if (cid & 0xFF) >= max_codec:
                    continue
                yield (label, vals, cid, words(v), None) if full else (label, vals, cid, None, tuple(v))


def check(got_words: np.ndarray, want, dig, what: str):
    got = np.ascontiguousarray(got_words).view(np.int32)
    if want is not None:
        assert got.size == want.size and np.array_equal(got, want), f"{what}: stream differs from the reference's"
    else:
        assert got.size == dig[0], f"{what}: {got.size} words, the reference wrote {dig[0]}"
        assert hashlib.sha256(got.astype("<u4").tobytes()).hexdigest() == dig[1], f"{what}: stream differs from the reference's"


def reused_inputs(seed: int):
    """The list sequence make_codec_streams_golden.py feeds ONE long-lived FastPFOR object."""
    rng = np.random.default_rng(seed)
    seq = []
    for n, hi_rate in ((1024, 0.2), (512, 0.05), (2048, 0.3), (256, 0.5), (70000, 0.1), (768, 0.02)):
        v = rng.integers(0, 32, size=n)
        m = rng.random(n) < hi_rate
        v = np.where(m, rng.integers(0, 1 << rng.integers(6, 31), size=n), v)
        seq.append(v)
    return seq
