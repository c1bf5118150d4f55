#!/usr/bin/env python3
"""Golden codec streams produced by EXECUTING the reference's Java sources (javamini; no JVM here).

For every input of tests/cases.reference_matrix(big=True) and every codec id of include/mcp.h
(with and without the Delta flag) this runs the reference's own classes -- Composition, BinaryPacking,
VariableByte, FastPFOR, FastPFOR128, Delta, IntegratedComposition, IntegratedBinaryPacking,
IntegratedVariableByte, IntCompressor, IntegratedIntCompressor, SkippableComposition,
SkippableIntegratedComposition, XorBinaryPacking, OptPFD (+ S16) -- on a fresh codec object, checks that the
reference decodes its own stream back to the input, and records the stream: in full for lists of up to
FULL_MAX ints, as (length, SHA-256) above that.  Also recorded:

* `reused`: ONE FastPFOR / FastPFOR128 object compressing a sequence of lists, so that the stale
  exception-buffer padding (FastPFOR.java:143 clears the pointers, never the buffers) is observed from the
  reference rather than emulated;
* `loader`: LoadPortfolios.getCompressedInstruments (LoadPortfolios.java:608-623, lifted verbatim out of the
  file) on the instrument lists of the shipped portfolios.json, i.e. the bytes the reference stores.

Run in the build container only (needs /root/reference):  python tests/golden/make_codec_streams_golden.py
The tests (tests/test_golden_streams.py, tests/test_gpu_codec.py) compare oracle/ and the CUDA path to this file.
"""
from __future__ import annotations

import base64
import hashlib
import json
import os
import sys
import time
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import cases  # noqa: E402
import refjava as R  # noqa: E402

FULL_MAX = 1100
CODECS = list(range(11))  # 0..8 as in mcp.h; 9 = IntegratedComposition(XorBinaryPacking, IVB); 10 = Composition(OptPFD, VB)


def pack_words(words) -> str:
    a = np.asarray([w & 0xFFFFFFFF for w in words], dtype="<u4")
    return base64.b64encode(zlib.compress(a.tobytes(), 9)).decode()


def digest(words) -> list:
    a = np.asarray([w & 0xFFFFFFFF for w in words], dtype="<u4")
    return [int(a.size), hashlib.sha256(a.tobytes()).hexdigest()]


def main():
    t0 = time.time()
    gold = {
        "source": "JavaFastPFOR 0.1.8-SNAPSHOT + maprdb-portfolio sources under /root/reference, executed by tests/golden/javamini.py",
        "full_max": FULL_MAX, "streams": {}, "digests": {}, "throws": {}, "reused": [], "loader": [],
    }
    for label, vals in cases.reference_matrix(big=True):
        want = [R.w32(v) for v in vals]
        for base in CODECS:
            for d in (0, R.DELTA):
                cid = base | d
                if len(vals) > 200000 and base in (9, 10):
                    continue  # the multi-page inputs are about FastPFOR; OptPFD's cost search is slow interpreted
                key = f"{cid}"
                try:
                    words = R.compress(cid, vals)
                except (R.J.JavaException, IndexError) as ex:
                    # the reference itself throws (IntCompressor's n + 1024 output array is too small for incompressible
                    # input): recorded, so that the tests know this pair has no stream
                    gold["throws"].setdefault(label, {})[key] = getattr(ex, "kind", "ArrayIndexOutOfBoundsException")
                    continue
                if len(vals) <= 70000:
                    back = R.uncompress(cid, words, len(vals))
                    assert back == want, ("the reference does not round-trip?", label, cid)
                if len(vals) <= FULL_MAX:
                    gold["streams"].setdefault(label, {})[key] = pack_words(words)
                else:
                    gold["digests"].setdefault(label, {})[key] = digest(words)
        print(f"{label}: n={len(vals)}  ({time.time() - t0:.0f}s)", flush=True)

    # one long-lived object, several lists (stale exception buffers carry over between calls and pages)
    rng = np.random.default_rng(0x57A1E)
    seq = []
    for n, hi_rate in ((1024, 0.2), (512, 0.05), (2048, 0.3), (256, 0.5), (70000, 0.1), (768, 0.02)):
        v = rng.integers(0, 32, size=n)
        m = rng.random(n) < hi_rate
        v = np.where(m, rng.integers(0, 1 << rng.integers(6, 31), size=n), v)
        seq.append(v)
    for base in (1, 2):
        codec = R.new_codec(base)
        entry = {"codec": base, "seed": 0x57A1E, "streams": []}
        for v in seq:
            words = R.compress_with(codec, base, v)
            entry["streams"].append(pack_words(words) if len(v) <= 4096 else digest(words))
        gold["reused"].append(entry)

    # the bytes the reference's loader stores for the shipped portfolios
    c1 = json.load(open(os.path.join(HERE, "c1_sample.json")))
    for p in c1["portfolios"]:
        inst = p["instruments"]
        blob = R.get_compressed_instruments(inst)
        assert R.get_decompressed_instruments(blob, len(inst)) == inst
        gold["loader"].append({"n": len(inst), "sha256": hashlib.sha256(blob).hexdigest(), "bytes": len(blob),
                               "b64": base64.b64encode(blob).decode() if len(inst) <= 64 else None})
    out = os.path.join(HERE, "codec_streams_golden.json")
    with open(out, "w") as f:
        json.dump(gold, f, separators=(",", ":"))
    print("wrote", out, os.path.getsize(out), "bytes in", f"{time.time() - t0:.0f}s")


if __name__ == "__main__":
    main()
