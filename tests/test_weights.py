"""The loader's weights column: Snappy.compress(double[]) = the raw Snappy format over the doubles'
bytes (LoadPortfolios.java:563-595).  snappy-java is un-vendored, so the compressed BYTES are parity-unpinned; what is
checked: the format (hand-assembled streams with every element kind decode), round trips, the oracle = the GPU, and the
loader's store-iff-smaller rule."""
import json
import os

import numpy as np
import pytest

import mcp_b200
from oracle import cref

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def handmade_streams():
    """(stream, expected bytes): assembled by hand from the Snappy format description -- not by our encoder."""
    out = []
    # 16 literal bytes
    lit = bytes(range(16))
    out.append((bytes([16, (16 - 1) << 2]) + lit, lit))
    # literal "abcdefgh", copy with 1-byte offset (tag 01): length 8, offset 8 -> 16 bytes
    s = bytes([16, (8 - 1) << 2]) + b"abcdefgh" + bytes([1 | ((8 - 4) << 2) | (0 << 5), 8])
    out.append((s, b"abcdefgh" * 2))
    # copy with 2-byte offset (tag 10), overlapping: literal 8 bytes, copy length 24 offset 8 -> pattern repeats
    s = bytes([32, (8 - 1) << 2]) + b"01234567" + bytes([2 | ((24 - 1) << 2), 8, 0])
    out.append((s, b"01234567" * 4))
    # copy with 4-byte offset (tag 11)
    s = bytes([24, (8 - 1) << 2]) + b"ABCDEFGH" + bytes([3 | ((16 - 1) << 2), 8, 0, 0, 0])
    out.append((s, b"ABCDEFGH" * 3))
    # long literal: 64 bytes -> tag 60 << 2 with one extra length byte; varint length 64
    lit = bytes((7 * i) & 255 for i in range(64))
    out.append((bytes([64, 60 << 2, 63]) + lit, lit))
    # 2-byte varint length (200 = 0xC8 0x01), literal of 200 bytes with a one-byte extra length
    lit = bytes((3 * i + 1) & 255 for i in range(200))
    out.append((bytes([0xC8, 0x01, 60 << 2, 199]) + lit, lit))
    return out


def test_oracle_decodes_every_element_kind_of_the_snappy_format():
    for stream, want in handmade_streams():
        assert cref.snappy_uncompress(stream, len(want)) == want
    for bad in (b"", bytes([8, 0]), bytes([8, (8 - 1) << 2]) + b"abc", bytes([16, 1 | (4 << 2), 8])):
        with pytest.raises(ValueError):
            cref.snappy_uncompress(bad, 64)


def test_oracle_encoder_writes_valid_snappy_and_round_trips():
    rng = np.random.default_rng(3)
    cases_ = [np.zeros(0), np.array([1.5]), rng.random(1110) * 500 + 1, np.repeat([1.25, 2.5, 1.25], 400), np.full(5000, 7.0),
              np.round(rng.random(3000) * 8) / 8]
    for w in cases_:
        blob = cref.get_compressed_weights(w)
        assert np.array_equal(cref.get_decompressed_weights(blob, w.size), w)
        assert len(blob) <= 32 + 8 * w.size + (8 * w.size) // 6
    # repetitive weights compress, random mantissas do not (the loader then stores them raw under "none")
    assert len(cref.get_compressed_weights(np.full(5000, 7.0))) < 8 * 5000 // 10
    assert len(cref.get_compressed_weights(rng.random(1110))) >= 8 * 1110


def test_double_byte_order_of_the_none_fallback():
    w = np.array([1.0, -2.5, 539.2])
    assert cref.doubles_to_bytes(w) == w.astype(">f8").tobytes() == mcp_b200.LoadPortfolios.doublesToBytes(w)


def test_weight_document_rule():
    """LoadPortfolios.java:270-278: the blob is stored only if smaller than 8 n bytes."""
    lp = mcp_b200.LoadPortfolios.__new__(mcp_b200.LoadPortfolios)
    w = [1.5, 2.5, 3.5]
    d = lp.getWeightDocument(w, compressed=b"x" * 23)
    assert d == {"uncompressedDoubleSize": 3, "compressed": b"x" * 23, "compressedByteSize": 23, "algo": "snappy"}
    d = lp.getWeightDocument(w, compressed=b"x" * 24)
    assert d["algo"] == "none" and d["compressed"] == np.asarray(w).astype(">f8").tobytes() and d["compressedByteSize"] == 24


@pytest.mark.gpu
def test_gpu_weights_codec_equals_the_oracle_and_reads_foreign_streams(ctx):
    rng = np.random.default_rng(9)
    lists = [np.zeros(0), np.array([3.25]), rng.random(17) * 500, rng.random(1110) * 539.2 + 1, np.repeat([1.25, 2.5, 1.25], 400),
             np.full(9000, 7.0), np.round(rng.random(3000) * 8) / 8, np.tile(rng.random(7), 300)]
    off = np.zeros(len(lists) + 1, np.int64)
    np.cumsum([len(x) for x in lists], out=off[1:])
    vals = np.concatenate(lists)
    out_off, out = ctx.weights_encode(vals, off)
    for i, w in enumerate(lists):
        assert bytes(out[out_off[i]: out_off[i + 1]]) == cref.get_compressed_weights(w), i
    assert np.array_equal(ctx.weights_decode(out, out_off, off), vals)
    # single-list calls, like the loader makes them
    blob = ctx.get_compressed_weights(lists[3])
    assert blob == cref.get_compressed_weights(lists[3])
    assert np.array_equal(ctx.get_decompressed_weights(blob, lists[3].size), lists[3])
    # streams our encoder did not write: every element kind of the format, as doubles
    for stream, want in handmade_streams():
        if len(want) % 8:
            continue
        got = ctx.get_decompressed_weights(stream, len(want) // 8)
        assert got.tobytes() == want
    with pytest.raises(mcp_b200.McpError) as e:
        ctx.get_decompressed_weights(bytes([16, 1 | (4 << 2), 8]), 2)   # a copy before any output
    assert e.value.code == mcp_b200.native.MCP_E_FORMAT
    with pytest.raises(mcp_b200.CapacityError):
        ctx.weights_encode(vals, off, cap=int(out.size) - 1)


@pytest.mark.gpu
def test_gpu_weights_codec_on_many_lists(ctx):
    """Many lists of every length up to the loader's 8 192-position cap (LoadPortfolios.java:705), bytes against the oracle:
    decimal weights give a copy every few bytes, random mantissas none."""
    rng = np.random.default_rng(77)
    lens = list(rng.integers(0, 400, 600)) + [1279, 1280, 1281, 1023, 1024, 1025, 2047, 2049, 640, 8191]
    lists = []
    for i, n in enumerate(lens):
        kind = i % 4
        if kind == 0:
            w = np.round(rng.random(n) * 100.0, 1)
        elif kind == 1:
            w = rng.random(n)
        elif kind == 2:
            w = rng.integers(0, 5, n).astype(np.float64) * 0.25
        else:
            w = np.where(rng.random(n) < 0.5, 1.0, np.round(rng.random(n), 2))
        lists.append(w)
    off = np.zeros(len(lists) + 1, np.int64)
    np.cumsum([len(x) for x in lists], out=off[1:])
    vals = np.concatenate(lists)
    out_off, out = ctx.weights_encode(vals, off)
    for i, w in enumerate(lists):
        assert bytes(out[out_off[i]: out_off[i + 1]]) == cref.get_compressed_weights(w), (i, len(w))
    assert np.array_equal(ctx.weights_decode(out, out_off, off), vals)


@pytest.mark.gpu
def test_weight_documents_of_the_shipped_portfolios(ctx):
    c1 = json.load(open(os.path.join(GOLD, "c1_sample.json")))
    lp = mcp_b200.LoadPortfolios(ctx)
    for p in c1["portfolios"]:
        w = p["weights"]
        d = lp.getWeightDocument(w)
        assert d["uncompressedDoubleSize"] == len(w) and d["compressedByteSize"] == len(d["compressed"])
        back = lp.getDecompressedWeights(d["compressed"], len(w)) if d["algo"] == "snappy" else np.frombuffer(d["compressed"], ">f8")
        assert np.array_equal(np.asarray(back, np.float64), np.asarray(w, np.float64))
