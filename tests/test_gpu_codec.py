"""GPU parity: the CUDA codec path (through the C ABI) against the oracle, byte for byte."""
import base64
import json
import os

import numpy as np
import pytest

import cases
import mcp_b200
from mcp_b200 import native as N
from oracle import cref

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
KAT_INTS = [719, 5, 3990, 4112, 3328, 78, 1163, 135, 263, 67, 67, 1860, 75, 1342, 1145, 38, 1216]
KAT_B64 = "AAAAABaFhU8AoBCfiQvOmoIHgQeORMPDeYo+y4lApogAAAAA"


def batch_of(matrix):
    lens = [len(v) for _, v in matrix]
    off = np.zeros(len(matrix) + 1, np.int64)
    np.cumsum(lens, out=off[1:])
    vals = np.concatenate([cases.to_i32(v) for _, v in matrix])
    return vals, off


FLAT_DEFAULT = 1  # the library's default for option flat_codec


def test_kat_through_the_c_abi(ctx):
    blob = base64.b64decode(KAT_B64)
    assert ctx.get_compressed_instruments(N.CODEC_BP32_VB, KAT_INTS) == blob
    assert list(ctx.get_decompressed_instruments(N.CODEC_BP32_VB, blob, 17)) == KAT_INTS
    lp = mcp_b200.LoadPortfolios(ctx)
    assert lp.getCompressedInstruments(KAT_INTS) == blob
    doc = lp.getInstrumentDocument(KAT_INTS)
    assert doc["algo"] == "bpacking+variablebyte" and doc["compressedByteSize"] == 36 and doc["uncompressedIntegerSize"] == 17


@pytest.fixture(params=[(1, 2, 0), (1, 1, 0), (1, 0, 0), (0, 0, 0), (1, 2, 1)],
                ids=["shortlist-everywhere", "default", "batched", "warp-per-list", "flat"])
def codec_ctx(ctx, request):
    """The kernel families behind the codec ABI: the thread-per-list kernels (lists averaging <= 144 values), the
    batched single-pass kernels (thread per block), the warp-per-list size/scan/emit kernels that handle everything
    else, and the flat (thread-per-16-values) kernels for short lists that are all VariableByte."""
    ctx.set_option("fast_codec", request.param[0])
    ctx.set_option("short_list_codec", request.param[1])
    ctx.set_option("flat_codec", request.param[2])
    yield ctx
    ctx.set_option("fast_codec", 1)
    ctx.set_option("short_list_codec", 1)
    ctx.set_option("flat_codec", FLAT_DEFAULT)


@pytest.mark.parametrize("codec", cases.ALL_CODECS, ids=cases.NAMES)
@pytest.mark.parametrize("delta", [0, cases.DELTA], ids=["plain", "delta"])
@pytest.mark.parametrize("framing", [N.FRAME_WORDS, N.FRAME_APP], ids=["words", "app"])
def test_encode_bit_exact_and_decode_roundtrip(codec_ctx, codec, delta, framing):
    ctx = codec_ctx
    matrix = cases.reference_matrix(big=False)
    vals, off = batch_of(matrix)
    cid = codec | delta
    out_off, out = ctx.codec_encode(cid, framing, vals, off)
    assert out_off[0] == 0 and out_off[-1] == out.size
    for i, (label, v) in enumerate(matrix):
        got = bytes(out[out_off[i]: out_off[i + 1]])
        ref = cref.get_compressed_instruments(cid, v) if framing == N.FRAME_APP else cref.compress(cid, v).tobytes()
        assert got == ref, f"{label}: GPU bytes differ from the oracle ({len(got)} vs {len(ref)} bytes)"
    back = ctx.codec_decode(cid, framing, out, out_off, off)
    assert np.array_equal(back, vals)


def test_golden_bitpacking_through_gpu(ctx):
    """A 32-int list whose maxbits is b makes BinaryPacking emit [32][b][fastpackwithoutmask_b(...)]:
    compare the packed words with the vectors interpreted from the reference's BitPacking.java."""
    g = json.load(open(os.path.join(GOLD, "bitpacking_golden.json")))
    cs = [c for c in g["fastpackwithoutmask"] if c["b"] > 0 and max(c["in"]).bit_length() == c["b"] and max(c["in"]) < 2 ** c["b"]]
    assert len({c["b"] for c in cs}) == 32
    vals = np.concatenate([np.array(c["in"], np.uint32).view(np.int32) for c in cs])
    off = np.arange(len(cs) + 1, dtype=np.int64) * 32
    out_off, out = ctx.codec_encode(N.CODEC_BP32_VB, N.FRAME_WORDS, vals, off)
    for i, c in enumerate(cs):
        w = out[out_off[i]: out_off[i + 1]].view(np.uint32)
        assert w[0] == 32 and w[1] == c["b"] and list(w[2:]) == c["out"], c["b"]
    # and the integrated packer: sorted lists with init 0
    cs = [c for c in g["integratedpack"] if c["init"] == 0]
    # (generator uses random init; the in-list chaining from 0 is covered by the oracle parity test above)
    un = g["fastunpack"]
    # decode side: feed [32][b][words] and expect fastunpack_b's output
    blobs, lens = [], []
    for c in un:
        words = np.array([32, c["b"]] + c["in"], np.uint32)
        blobs.append(words.view(np.uint8))
        lens.append(words.size * 4)
    in_off = np.zeros(len(un) + 1, np.int64)
    np.cumsum(lens, out=in_off[1:])
    out_off2 = np.arange(len(un) + 1, dtype=np.int64) * 32
    dec = ctx.codec_decode(N.CODEC_BP32_VB, N.FRAME_WORDS, np.concatenate(blobs), in_off, out_off2).view(np.uint32)
    for i, c in enumerate(un):
        assert list(dec[32 * i: 32 * i + 32]) == c["out"], c["b"]


def test_breach_like_batch_all_codecs(codec_ctx):
    ctx = codec_ctx
    rng = np.random.default_rng(11)
    vals, off = cases.breach_like_lists(rng, 3000, 10000, 0.01)       # config C4 shape: ~100 ints per list
    vals2, off2 = cases.breach_like_lists(rng, 300, 100000, 0.01)     # ~1000 ints per list: exercises FastPFOR blocks
    for v, o in ((vals, off), (vals2, off2)):
        for cid in (N.CODEC_BP32_VB, N.CODEC_BP32_VB | N.FLAG_DELTA, N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA,
                    N.CODEC_FASTPFOR128_VB | N.FLAG_DELTA, N.CODEC_IBP32_IVB, N.CODEC_SK_IBP32_IVB, N.CODEC_VB | N.FLAG_DELTA):
            out_off, out = ctx.codec_encode(cid, N.FRAME_APP, v, o)
            nl = o.size - 1
            stride = int(np.max(np.diff(out_off)))
            ref, lens = cref.batch_compress(cid, 1, v, o, stride)
            assert np.array_equal(np.diff(out_off), lens)
            for i in range(nl):
                assert bytes(out[out_off[i]: out_off[i + 1]]) == bytes(ref[i * stride: i * stride + lens[i]]), (cid, i)
            assert np.array_equal(ctx.codec_decode(cid, N.FRAME_APP, out, out_off, o), v)


@pytest.mark.parametrize("cid", [N.CODEC_BP32_VB, N.CODEC_BP32_VB | N.FLAG_DELTA, N.CODEC_SK_BP32_VB | N.FLAG_DELTA, N.CODEC_IBP32_IVB,
                                 N.CODEC_SK_IBP32_IVB], ids=["bp", "bp+delta", "skbp+delta", "ibp", "skibp"])
@pytest.mark.parametrize("framing", [N.FRAME_WORDS, N.FRAME_APP], ids=["words", "app"])
def test_fast_path_batch_edges(ctx, cid, framing):
    """Shapes that stress the batching of the fast path: batches whose values or compressed words overflow the
    shared-memory staging (long lists among short ones, incompressible 32-bit data), empty lists, lists that are all
    tail, exact multiples of 32 and 128, and negative (top-bit) values."""
    rng = np.random.default_rng(cid * 7 + framing)
    lists = []
    for _ in range(400):
        kind = rng.integers(0, 8)
        n = int(rng.choice([0, 1, 5, 31, 32, 33, 64, 96, 100, 127, 128, 129, 160, 255, 256, 300, 512, 640, 1000]))
        if kind == 0:
            v = rng.integers(-2**31, 2**31 - 1, n)                              # incompressible: width 32
        elif kind == 1:
            v = np.sort(rng.integers(0, 2**31 - 1, n))                          # sorted, big gaps
        elif kind == 2:
            v = np.cumsum(rng.geometric(0.01, n))                               # breach-like
        elif kind == 3:
            v = np.full(n, int(rng.integers(0, 1000)))                          # constant: width 0 under delta
        elif kind == 4:
            v = rng.integers(0, 2 ** int(rng.integers(1, 31)), n)               # uniform width
        elif kind == 5:
            v = np.cumsum(rng.integers(0, 3, n)) + 2**31 - 50                   # crosses the sign bit
        else:
            v = rng.integers(0, 200, n)
        lists.append(np.asarray(v, np.int64))
    for j in (17, 211):
        lists[j] = np.cumsum(rng.geometric(0.02, 6000))                         # longer than the value staging area
    vals = np.concatenate([cases.to_i32(v) for v in lists])
    off = np.zeros(len(lists) + 1, np.int64)
    np.cumsum([len(v) for v in lists], out=off[1:])
    out_off, out = ctx.codec_encode(cid, framing, vals, off)
    assert out_off[0] == 0 and out_off[-1] == out.size
    for i, v in enumerate(lists):
        ref = cref.get_compressed_instruments(cid, v) if framing == N.FRAME_APP else cref.compress(cid, v).tobytes()
        assert bytes(out[out_off[i]: out_off[i + 1]]) == ref, (i, len(v))
    assert np.array_equal(ctx.codec_decode(cid, framing, out, out_off, off), vals)
    # a malformed width anywhere in the batch is reported, not decoded
    if framing == N.FRAME_WORDS and cid == N.CODEC_BP32_VB:
        bad = out.copy()
        i = next(k for k, v in enumerate(lists) if 32 <= len(v) < 128)
        bad.view(np.uint32)[out_off[i] // 4 + 1] = 99
        with pytest.raises(mcp_b200.McpError) as e:
            ctx.codec_decode(cid, framing, bad, out_off, off)
        assert e.value.code == N.MCP_E_FORMAT


def test_fast_kernels_are_the_ones_that_run(ctx):
    """The batched single-pass kernels must actually serve the shapes they were built for (C4's ~100-int lists under
    the loader's codec, ~1 000-int breach lists under FastPFOR) instead of silently handing them back."""
    rng = np.random.default_rng(2)
    for cid, (nl, ntr) in ((N.CODEC_BP32_VB | N.FLAG_DELTA, (2000, 10000)), (N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA, (400, 100000)),
                           (N.CODEC_FASTPFOR128_VB | N.FLAG_DELTA, (400, 100000)), (N.CODEC_FASTPFOR256_VB, (2000, 10000))):
        vals, off = cases.breach_like_lists(rng, nl, ntr, 0.01)
        before = ctx.codec_stats()
        out_off, out = ctx.codec_encode(cid, N.FRAME_APP, vals, off, cap=int(8 * vals.size + 64 * nl))
        after = ctx.codec_stats()
        assert after["fast_encodes"] == before["fast_encodes"] + 1 and after["fast_fallbacks"] == before["fast_fallbacks"], cid
        for i in (0, nl // 2, nl - 1):
            assert bytes(out[out_off[i]: out_off[i + 1]]) == cref.get_compressed_instruments(cid, vals[off[i]: off[i + 1]]), (cid, i)
    # a list far beyond the staging area among short ones: handed back, still right
    lists = [np.cumsum(rng.geometric(0.01, 900)) for _ in range(40)]
    lists[7] = np.cumsum(rng.geometric(0.01, 30000))
    vals = np.concatenate(lists).astype(np.int32)
    off = np.zeros(41, np.int64)
    np.cumsum([len(x) for x in lists], out=off[1:])
    cid = N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA
    before = ctx.codec_stats()
    out_off, out = ctx.codec_encode(cid, N.FRAME_WORDS, vals, off, cap=int(8 * vals.size))
    assert ctx.codec_stats()["fast_fallbacks"] == before["fast_fallbacks"] + 1
    for i in (0, 7, 39):
        assert bytes(out[out_off[i]: out_off[i + 1]]) == cref.compress(cid, lists[i]).tobytes(), i


def test_fastpfor_fast_decode_rejects_malformed_pages(ctx):
    rng = np.random.default_rng(8)
    vals, off = cases.breach_like_lists(rng, 50, 100000, 0.01)      # ~1 000 ints per list: FastPFOR pages
    cid = N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA
    out_off, out = ctx.codec_encode(cid, N.FRAME_WORDS, vals, off)
    assert np.array_equal(ctx.codec_decode(cid, N.FRAME_WORDS, out, out_off, off), vals)
    w = out.view(np.uint32)
    for what in ("where_meta", "block width", "header"):
        bad = w.copy()
        p0 = out_off[7] // 4
        if what == "where_meta":
            bad[p0 + 1] = 0x7fffffff                      # page[0] points outside the list
        elif what == "block width":
            wm = int(w[p0 + 1])
            bad[p0 + 1 + wm + 1] = (bad[p0 + 1 + wm + 1] & 0xffffff00) | 77  # first metadata byte: b = 77 > 32
        else:
            bad[p0] = 256 * 17                            # Composition header disagrees with the list length
        with pytest.raises(mcp_b200.McpError) as e:
            ctx.codec_decode(cid, N.FRAME_WORDS, bad.view(np.uint8), out_off, off)
        assert e.value.code == N.MCP_E_FORMAT, what


@pytest.mark.parametrize("framing", [N.FRAME_WORDS, N.FRAME_APP], ids=["words", "app"])
@pytest.mark.parametrize("codec", [N.CODEC_FASTPFOR128_VB | N.FLAG_DELTA, N.CODEC_XOR128_IVB, N.CODEC_OPTPFD_VB | N.FLAG_DELTA],
                         ids=["fastpfor128", "xor128", "optpfd"])
def test_flat_decoder_rejects_malformed_blocks_in_page_tiles(ctx, codec, framing):
    """Lists of ~110 values under the 128-int codecs: the flat decoder leaves the tiles that hold a block to
    k_vb4_decode_pages; a damaged block there must end in MCP_E_FORMAT (never in a read or write outside the buffers),
    and the undamaged stream must still decode."""
    rng = np.random.default_rng(31)
    vals, off = _awkward("near128", rng)
    out_off, out = ctx.codec_encode(codec, framing, vals, off)
    assert np.array_equal(ctx.codec_decode(codec, framing, out, out_off, off), vals)
    lens = np.diff(off)
    victims = [int(i) for i in np.flatnonzero(lens >= 128)[:6]]
    assert victims
    swap = (lambda x: int(np.uint32(x).byteswap())) if framing == N.FRAME_APP else (lambda x: int(x))
    for i in victims:
        for what in ("count word", "first block word", "truncated"):
            bad, boff = out.copy(), out_off.copy()
            w = bad.view(np.uint32)
            p0 = int(out_off[i] // 4)
            if what == "count word":
                w[p0] = swap(128 * 5)                      # the first stage claims more blocks than the list has
            elif what == "first block word":
                w[p0 + 1] = swap(0x7fffff4d)               # FastPFOR: where_meta outside the list; Xor / OptPFD: widths > 32 / S16 words beyond it
            else:
                boff = out_off.copy()
                boff[i + 1:] -= 16                         # the list loses its last four words
                bad = np.concatenate([out[: out_off[i + 1] - 16], out[out_off[i + 1]:]])
            with pytest.raises(mcp_b200.McpError) as e:
                ctx.codec_decode(codec, framing, bad, boff, off)
            assert e.value.code == N.MCP_E_FORMAT, (i, what)


def test_multipage_fastpfor(ctx):
    """> 65 536 ints: FastPFOR pages (FastPFOR.java:98-102), including the stale exception-buffer padding that a
    Java instance carries from page to page (FastPFOR.java:143): bytes identical to the oracle's emulation."""
    label, v = cases.reference_matrix(big=True)[-2]
    assert len(v) == 1333333
    vi = cases.to_i32(v)
    off = np.array([0, len(v)], np.int64)
    for cid in (N.CODEC_FASTPFOR256_VB, N.CODEC_FASTPFOR128_VB):
        w = cref.compress(cid, v)
        dec = ctx.codec_decode(cid, N.FRAME_WORDS, w.view(np.uint8), np.array([0, w.size * 4], np.int64), off)
        assert np.array_equal(dec, vi)
        out_off, out = ctx.codec_encode(cid, N.FRAME_WORDS, vi, off)
        assert bytes(out) == w.tobytes()
        assert np.array_equal(cref.uncompress(cid, out.view(np.int32), len(v)), vi)
    label, s = cases.reference_matrix(big=True)[-1]
    si = cases.to_i32(s)
    off = np.array([0, len(s)], np.int64)
    for cid in (N.CODEC_SK_IBP32_IVB, N.CODEC_BP32_VB | N.FLAG_DELTA):
        out_off, out = ctx.codec_encode(cid, N.FRAME_WORDS, si, off)
        assert bytes(out) == cref.compress(cid, s).tobytes()
        assert np.array_equal(ctx.codec_decode(cid, N.FRAME_WORDS, out, out_off, off), si)


def test_integercodec_interface_like_the_reference_tests(ctx):
    """BasicTest.saulTest (BasicTest.java:47-70): compress {2,3,4,5} at output offset x, then uncompress."""
    for codec in (mcp_b200.Composition(mcp_b200.BinaryPacking(), mcp_b200.VariableByte(), ctx),
                  mcp_b200.Composition(mcp_b200.FastPFOR(), mcp_b200.VariableByte(), ctx),
                  mcp_b200.Composition(mcp_b200.FastPFOR128(), mcp_b200.VariableByte(), ctx),
                  mcp_b200.IntegratedComposition(mcp_b200.IntegratedBinaryPacking(), mcp_b200.IntegratedVariableByte(), ctx),
                  mcp_b200.VariableByteCodec(ctx), mcp_b200.IntegratedVariableByteCodec(ctx)):
        for x in (0, 1, 7, 49):
            a = np.array([2, 3, 4, 5], np.int32)
            b = np.zeros(90, np.int32)
            c = np.zeros(a.size, np.int32)
            aOffset, bOffset = mcp_b200.IntWrapper(0), mcp_b200.IntWrapper(x)
            codec.compress(a, aOffset, a.size, b, bOffset)
            length = bOffset.get() - x
            assert aOffset.get() == 4
            bOffset.set(x)
            cOffset = mcp_b200.IntWrapper(0)
            codec.uncompress(b, bOffset, length, c, cOffset)
            assert list(c) == [2, 3, 4, 5] and cOffset.get() == 4 and bOffset.get() == x + length


def test_intcompressor_interfaces(ctx):
    # IntCompressorTest.java:32-76
    for N_ in (1, 10, 100, 1000, 10000):
        orig = 3 * np.arange(N_, dtype=np.int32) + 5
        for ic in (mcp_b200.IntCompressor(ctx), mcp_b200.IntegratedIntCompressor(ctx),
                   mcp_b200.IntCompressor(mcp_b200.SkippableComposition(mcp_b200.FastPFOR(), mcp_b200.VariableByte(), ctx)),
                   mcp_b200.IntCompressor(mcp_b200.SkippableComposition(mcp_b200.FastPFOR128(), mcp_b200.VariableByte(), ctx))):
            comp = ic.compress(orig)
            assert comp[0] == N_
            assert np.array_equal(ic.uncompress(comp), orig)


def test_out_too_small_is_an_index_error(ctx):
    """ExampleTest.java:55-57: an `out` array that is too small raises ArrayIndexOutOfBoundsException."""
    codec = mcp_b200.Composition(mcp_b200.BinaryPacking(), mcp_b200.VariableByte(), ctx)
    rng = np.random.default_rng(3)
    a = rng.integers(0, 2**31 - 1, 1000).astype(np.int32)
    with pytest.raises(IndexError):
        codec.compress(a, mcp_b200.IntWrapper(0), a.size, np.zeros(10, np.int32), mcp_b200.IntWrapper(0))
    with pytest.raises(mcp_b200.CapacityError):
        ctx.codec_encode(N.CODEC_BP32_VB, N.FRAME_APP, a, np.array([0, 1000], np.int64), cap=16)
    w = ctx.compress_ints(N.CODEC_BP32_VB, a)
    with pytest.raises(IndexError):
        ctx.uncompress_ints(N.CODEC_BP32_VB, w, out_cap=999)


def test_malformed_streams_are_rejected(ctx):
    good = ctx.compress_ints(N.CODEC_BP32_VB, np.arange(100, dtype=np.int32))
    bad = good.copy()
    bad[1] = 77  # block width > 32
    with pytest.raises(mcp_b200.McpError) as e:
        ctx.uncompress_ints(N.CODEC_BP32_VB, bad, out_cap=200)
    assert e.value.code == N.MCP_E_FORMAT
    with pytest.raises(mcp_b200.McpError):
        ctx.codec_decode(N.CODEC_BP32_VB, N.FRAME_WORDS, good.view(np.uint8)[:-4].copy(),
                         np.array([0, good.size * 4 - 4], np.int64), np.array([0, 100], np.int64))
    with pytest.raises(mcp_b200.McpError):
        ctx.codec_encode(99, 0, np.zeros(4, np.int32), np.array([0, 4], np.int64))


def test_empty_batch_and_empty_lists(ctx):
    out_off, out = ctx.codec_encode(N.CODEC_BP32_VB, N.FRAME_APP, np.zeros(0, np.int32), np.zeros(1, np.int64))
    assert out.size == 0 and list(out_off) == [0]
    off = np.array([0, 0, 3, 3], np.int64)
    out_off, out = ctx.codec_encode(N.CODEC_BP32_VB, N.FRAME_APP, np.array([7, 8, 9], np.int32), off)
    assert bytes(out[out_off[0]:out_off[1]]) == b"\0\0\0\0" and bytes(out[out_off[2]:out_off[3]]) == b"\0\0\0\0"
    assert np.array_equal(ctx.codec_decode(N.CODEC_BP32_VB, N.FRAME_APP, out, out_off, off), [7, 8, 9])


@pytest.mark.parametrize("cid", [N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA, N.CODEC_BP32_VB | N.FLAG_DELTA], ids=["fastpfor256+delta", "bp+delta"])
def test_full_size_c4_properties(ctx, cid):
    """BASELINE config C4 at full size (1 M lists at 1 % of 10 000 trials), device-resident, under the codec the north
    star names and under the loader's: a bit-exact encode -> decode round trip, plus a spot check of 64 lists against
    the oracle generator + oracle codec."""
    nl = 1_000_000
    dv, do, total = ctx.generate_breach_lists(nl, 10000, 0.01, 0x5EED0004)
    assert 0.95e8 < total < 1.05e8
    d_out_off = ctx.dev_alloc(8 * (nl + 1))
    cap = 5 * total + 16 * nl
    d_out = ctx.dev_alloc(cap)
    d_back = ctx.dev_alloc(4 * total + 16)
    try:
        enc = ctx.codec_encode_dev(cid, N.FRAME_APP, dv, do, nl, total, d_out_off, d_out, cap)
        assert enc < 2.2 * total  # well under 2.2 bytes/int at gap ~100
        ctx.codec_decode_dev(cid, N.FRAME_APP, d_out, d_out_off, nl, d_back, do)
        a = np.empty(total, np.int32)
        b = np.empty(total, np.int32)
        ctx.d2h(a, dv)
        ctx.d2h(b, d_back)
        assert np.array_equal(a, b)
        off = np.empty(nl + 1, np.int64)
        ctx.d2h(off, do)
        ooff = np.empty(nl + 1, np.int64)
        ctx.d2h(ooff, d_out_off)
        blob = np.empty(enc, np.uint8)
        ctx.d2h(blob, d_out)
        for li in list(range(32)) + list(range(nl - 32, nl)):
            ref_list = cref.gen_breach_list(li, 10000, 0.01, 0x5EED0004)
            assert np.array_equal(a[off[li]: off[li + 1]], ref_list), li
            assert bytes(blob[ooff[li]: ooff[li + 1]]) == cref.get_compressed_instruments(cid, ref_list), li
    finally:
        for p in (d_out_off, d_out, d_back):
            ctx.dev_free(p)


@pytest.fixture
def shortlist_everywhere(ctx):
    """short_list_codec = 2: every codec with short lists goes to the thread-per-list kernels; the default is restored
    whatever the test does."""
    ctx.set_option("short_list_codec", 2)
    try:
        yield ctx
    finally:
        ctx.set_option("fast_codec", 1)
        ctx.set_option("short_list_codec", 1)


def _short_lists(rng, count, lengths):
    lists = []
    for _ in range(count):
        kind = rng.integers(0, 8)
        n = int(rng.choice(lengths))
        if kind == 0:
            v = rng.integers(-2**31, 2**31 - 1, n)                              # incompressible: width 32, 5-byte VByte
        elif kind == 1:
            v = np.sort(rng.integers(0, 2**31 - 1, n))                          # sorted, big gaps
        elif kind == 3:
            v = np.full(n, int(rng.integers(0, 1000)))                          # constant: width 0 under delta
        elif kind == 4:
            v = rng.integers(0, 2 ** int(rng.integers(1, 31)), n)               # uniform width
        elif kind == 5:
            v = np.cumsum(rng.integers(0, 3, n)) + 2**31 - 50                   # crosses the sign bit
        else:
            v = np.cumsum(rng.geometric(0.01, n))                               # breach-like
        lists.append(np.asarray(v, np.int64))
    return lists


@pytest.mark.parametrize("codec", cases.FAST_CODECS, ids=cases.FAST_NAMES)
@pytest.mark.parametrize("delta", [0, cases.DELTA], ids=["plain", "delta"])
@pytest.mark.parametrize("framing", [N.FRAME_WORDS, N.FRAME_APP], ids=["words", "app"])
def test_short_list_kernels_bit_exact(shortlist_everywhere, codec, delta, framing):
    """The thread-per-list kernels on the lengths around every boundary they have (empty, all tail, 31/32/33,
    127/128/129, a FastPFOR128 block, columns that overflow and send the warp batch to the warp-per-list routines or
    the call to the batched kernels), every codec id, both framings: bytes equal to the oracle's, decode round trip."""
    if codec in (N.CODEC_IBP32_IVB, N.CODEC_SK_IBP32_IVB) and delta:
        pytest.skip("Delta before an integrated codec is not a combination the reference uses")
    rng = np.random.default_rng(1000 + 16 * codec + 2 * (1 if delta else 0) + framing)
    cid = codec | delta
    ctx = shortlist_everywhere
    for lengths in ([0, 1, 5, 31, 32, 33, 64, 96, 100, 127], [0, 1, 31, 32, 33, 64, 96, 100, 127, 128, 129, 160, 200, 255, 256, 300]):
        lists = _short_lists(rng, 333, lengths)
        vals = np.concatenate([cases.to_i32(v) for v in lists])
        off = np.zeros(len(lists) + 1, np.int64)
        np.cumsum([len(v) for v in lists], out=off[1:])
        assert vals.size <= 144 * len(lists)
        # (lists that average a FastPFOR block or more are not offered to these kernels at all: flat_pages_common)
        tried = 0 if codec in (N.CODEC_FASTPFOR128_VB, N.CODEC_SK_FASTPFOR128_VB) and vals.size >= 0.95 * 128 * len(lists) else 1
        before = ctx.codec_stats()
        out_off, out = ctx.codec_encode(cid, framing, vals, off, cap=int(8 * vals.size + 64 * len(lists)))
        mid = ctx.codec_stats()
        assert mid["short_calls"] + mid["short_fallbacks"] == before["short_calls"] + before["short_fallbacks"] + tried
        assert out_off[0] == 0 and out_off[-1] == out.size
        for i, v in enumerate(lists):
            ref = cref.get_compressed_instruments(cid, v) if framing == N.FRAME_APP else cref.compress(cid, v).tobytes()
            assert bytes(out[out_off[i]: out_off[i + 1]]) == ref, (i, len(v))
        assert np.array_equal(ctx.codec_decode(cid, framing, out, out_off, off), vals)
        after = ctx.codec_stats()
        assert after["short_calls"] + after["short_fallbacks"] == mid["short_calls"] + mid["short_fallbacks"] + tried


@pytest.mark.parametrize("cid", [N.CODEC_BP32_VB | N.FLAG_DELTA, N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA, N.CODEC_FASTPFOR128_VB | N.FLAG_DELTA,
                                 N.CODEC_VB | N.FLAG_DELTA, N.CODEC_SK_IBP32_IVB, N.CODEC_IBP32_IVB, N.CODEC_SK_BP32_VB | N.FLAG_DELTA],
                         ids=["bp+delta", "fastpfor256+delta", "fastpfor128+delta", "vb+delta", "skibp", "ibp", "skbp+delta"])
def test_short_list_kernels_serve_c4(shortlist_everywhere, cid):
    """BASELINE config C4's shape (1 % of 10 000 trials, ~100 ints per list, not a multiple of 32 lists) is served by the
    thread-per-list kernels themselves -- no hand-over -- in both directions, and the bytes are the ones the batched
    and the warp-per-list kernels produce."""
    rng = np.random.default_rng(77)
    vals, off = cases.breach_like_lists(rng, 4099, 10000, 0.01)
    nl = off.size - 1
    ctx = shortlist_everywhere
    before = ctx.codec_stats()
    out_off, out = ctx.codec_encode(cid, N.FRAME_APP, vals, off, cap=int(8 * vals.size + 64 * nl))
    back = ctx.codec_decode(cid, N.FRAME_APP, out, out_off, off)
    after = ctx.codec_stats()
    if (cid & 0xff) == N.CODEC_FASTPFOR128_VB and np.max(np.diff(off)) >= 128 and not FLAT_DEFAULT:
        # a list of a whole FastPFOR128 block is not the thread-per-list kernels': the call moves on to the batched kernels
        # (the flat kernels, the default, leave the tiles that hold such a list to k_vb4_pages and serve the call)
        assert after["short_calls"] == before["short_calls"] and after["short_fallbacks"] == before["short_fallbacks"] + 2
    else:
        assert after["short_calls"] == before["short_calls"] + 2 and after["short_fallbacks"] == before["short_fallbacks"]
    assert np.array_equal(back, vals)
    stride = int(np.max(np.diff(out_off)))
    ref, lens = cref.batch_compress(cid, 1, vals, off, stride)
    assert np.array_equal(np.diff(out_off), lens)
    for i in range(nl):
        assert bytes(out[out_off[i]: out_off[i + 1]]) == bytes(ref[i * stride: i * stride + lens[i]]), i
    try:
        for opts in (("short_list_codec", 0), ("fast_codec", 0)):
            ctx.set_option(*opts)
            o2, b2 = ctx.codec_encode(cid, N.FRAME_APP, vals, off)
            assert np.array_equal(o2, out_off) and np.array_equal(b2, out)
            assert np.array_equal(ctx.codec_decode(cid, N.FRAME_APP, out, out_off, off), vals)
    finally:
        ctx.set_option("fast_codec", 1)
        ctx.set_option("short_list_codec", 1)


def test_short_list_decode_rejects_malformed_streams(shortlist_everywhere):
    rng = np.random.default_rng(5)
    vals, off = cases.breach_like_lists(rng, 500, 10000, 0.01)
    cid = N.CODEC_BP32_VB | N.FLAG_DELTA
    ctx = shortlist_everywhere
    out_off, out = ctx.codec_encode(cid, N.FRAME_WORDS, vals, off)
    w = out.view(np.uint32)
    i = 123
    assert off[i + 1] - off[i] >= 64
    for what, (pos, val) in {"count header": (0, 31), "block width": (1, 40), "truncated": (None, None)}.items():
        bad, boff = out.copy(), out_off.copy()
        if pos is None:
            boff = out_off.copy()
            boff[i + 1:] -= 8                                   # the list loses its last two words
            bad = np.concatenate([out[: out_off[i + 1] - 8], out[out_off[i + 1]:]])
        else:
            bad.view(np.uint32)[out_off[i] // 4 + pos] = val
        with pytest.raises(mcp_b200.McpError) as e:
            ctx.codec_decode(cid, N.FRAME_WORDS, bad, boff, off)
        assert e.value.code == N.MCP_E_FORMAT, what
    assert w[out_off[i] // 4] == 32 * ((off[i + 1] - off[i]) // 32)


def test_short_list_encode_reports_capacity(shortlist_everywhere):
    rng = np.random.default_rng(6)
    vals, off = cases.breach_like_lists(rng, 300, 10000, 0.01)
    ctx = shortlist_everywhere
    out_off, out = ctx.codec_encode(N.CODEC_BP32_VB | N.FLAG_DELTA, N.FRAME_APP, vals, off)
    with pytest.raises(mcp_b200.CapacityError):
        ctx.codec_encode(N.CODEC_BP32_VB | N.FLAG_DELTA, N.FRAME_APP, vals, off, cap=int(out.size) - 4)
    o2, b2 = ctx.codec_encode(N.CODEC_BP32_VB | N.FLAG_DELTA, N.FRAME_APP, vals, off, cap=int(out.size))
    assert np.array_equal(b2, out) and np.array_equal(o2, out_off)


@pytest.mark.parametrize("cid", [N.CODEC_BP32_VB | N.FLAG_DELTA, N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA, N.CODEC_VB, N.CODEC_SK_IBP32_IVB],
                         ids=["bp+delta", "fastpfor256+delta", "vb", "skibp"])
@pytest.mark.parametrize("short", [2, 0], ids=["thread-per-list", "batched"])
def test_device_codec_stays_inside_its_buffers(ctx, cid, short):
    """No compute-sanitizer on this pool: the device-resident entry points run between guard words instead.  Output
    offsets, the encoded stream and the decoded values each sit in a buffer of exactly the size the call is entitled to,
    followed by 4 KiB of a known pattern that must survive (mixed lengths, so that the in-kernel warp-per-list route
    and the last, partial warp batch are exercised too)."""
    rng = np.random.default_rng(31 + cid)
    lists = _short_lists(rng, 1000 + 13, [0, 1, 31, 32, 33, 64, 96, 100, 127, 128, 129, 160, 200])
    vals = np.concatenate([cases.to_i32(v) for v in lists])
    off = np.zeros(len(lists) + 1, np.int64)
    np.cumsum([len(v) for v in lists], out=off[1:])
    nl, total = len(lists), int(vals.size)
    ref_off, ref = ctx.codec_encode(cid, N.FRAME_APP, vals, off)
    enc_bytes = int(ref.size)
    guard = np.full(1024, 0x5AC3A55A, np.uint32)
    sizes = {"vals": 4 * total, "off": 8 * (nl + 1), "ooff": 8 * (nl + 1), "enc": enc_bytes, "back": 4 * total}
    bufs = {k: ctx.dev_alloc(v + guard.nbytes) for k, v in sizes.items()}
    try:
        ctx.set_option("short_list_codec", short)
        for k, v in sizes.items():
            ctx.h2d(bufs[k] + v, guard)
        ctx.h2d(bufs["vals"], vals)
        ctx.h2d(bufs["off"], off)
        got = ctx.codec_encode_dev(cid, N.FRAME_APP, bufs["vals"], bufs["off"], nl, total, bufs["ooff"], bufs["enc"], enc_bytes)
        assert got == enc_bytes
        ctx.codec_decode_dev(cid, N.FRAME_APP, bufs["enc"], bufs["ooff"], nl, bufs["back"], bufs["off"])
        out = np.empty(enc_bytes, np.uint8)
        ooff = np.empty(nl + 1, np.int64)
        back = np.empty(total, np.int32)
        ctx.d2h(out, bufs["enc"])
        ctx.d2h(ooff, bufs["ooff"])
        ctx.d2h(back, bufs["back"])
        assert np.array_equal(ooff, ref_off) and np.array_equal(out, ref) and np.array_equal(back, vals)
        for k, v in sizes.items():
            g = np.empty_like(guard)
            ctx.d2h(g, bufs[k] + v)
            assert np.array_equal(g, guard), f"guard after `{k}` overwritten"
        # one byte less than needed: reported, and still nothing outside
        with pytest.raises(mcp_b200.CapacityError):
            ctx.codec_encode_dev(cid, N.FRAME_APP, bufs["vals"], bufs["off"], nl, total, bufs["ooff"], bufs["enc"], enc_bytes - 4)
        g = np.empty_like(guard)
        ctx.d2h(g, bufs["enc"] + enc_bytes)
        assert np.array_equal(g, guard)
    finally:
        ctx.set_option("short_list_codec", 1)
        for p in bufs.values():
            ctx.dev_free(p)


SKIPPABLE = [(mcp_b200.BinaryPacking, mcp_b200.VariableByte), (mcp_b200.FastPFOR, mcp_b200.VariableByte),
             (mcp_b200.FastPFOR128, mcp_b200.VariableByte), (mcp_b200.IntegratedBinaryPacking, mcp_b200.IntegratedVariableByte)]


@pytest.mark.parametrize("stages", SKIPPABLE, ids=["bp+vb", "fastpfor+vb", "fastpfor128+vb", "ibp+ivb"])
def test_skippable_headless_api_like_the_reference_tests(ctx, stages):
    """SkippableBasicTest.java:34-66 consistentTest: headlessCompress then headlessUncompress with the count supplied by
    the caller, asserting that the decoder consumed exactly what the encoder produced (:57-58) -- lengths around every
    block boundary instead of all of 0..4096 (each call is a GPU round trip).  The stream equals IntCompressor's without
    compressed[0] (IntCompressor.java:38-46) and, for the non-integrated compositions, the oracle's."""
    c = mcp_b200.SkippableComposition(stages[0](), stages[1](), ctx)
    for n in (0, 1, 31, 32, 33, 127, 128, 129, 255, 256, 257, 511, 512, 1000, 4096):
        data = np.array([k % 128 for k in range(n)], np.int32)
        if stages[0] is mcp_b200.IntegratedBinaryPacking:
            data = np.cumsum(data).astype(np.int32)
        out = np.zeros(n + 1024, np.int32)
        inpos, outpos = mcp_b200.IntWrapper(0), mcp_b200.IntWrapper(7)
        c.headlessCompress(data, inpos, n, out, outpos)
        assert inpos.get() == n
        produced = outpos.get() - 7
        assert np.array_equal(out[7: outpos.get()], cref.compress(c.codec_id, data)[1:]), n
        back = np.zeros(n + 3, np.int32)
        ipos, opos = mcp_b200.IntWrapper(7), mcp_b200.IntWrapper(3)
        c.headlessUncompress(out, ipos, len(out) - 7, back, opos, n)   # the whole padded buffer, like the Java test
        assert opos.get() == 3 + n and ipos.get() - 7 == produced, n   # SkippableBasicTest.java:57-58
        assert np.array_equal(back[3:], data), n


def test_two_lists_back_to_back_headless(ctx):
    """ExampleTest.java's headless example: two arrays compressed one after the other into the same output, decoded
    from their known lengths."""
    c = mcp_b200.SkippableComposition(mcp_b200.FastPFOR(), mcp_b200.VariableByte(), ctx)
    rng = np.random.default_rng(21)
    a = np.cumsum(rng.geometric(0.01, 2342)).astype(np.int32)
    b = rng.integers(0, 1 << 20, 777).astype(np.int32)
    out = np.zeros(a.size + b.size + 2048, np.int32)
    inpos, outpos = mcp_b200.IntWrapper(0), mcp_b200.IntWrapper(0)
    c.headlessCompress(a, inpos, a.size, out, outpos)
    mid = outpos.get()
    inpos.set(0)
    c.headlessCompress(b, inpos, b.size, out, outpos)
    end = outpos.get()
    ra, rb = np.zeros(a.size, np.int32), np.zeros(b.size, np.int32)
    ipos = mcp_b200.IntWrapper(0)
    c.headlessUncompress(out, ipos, end, ra, mcp_b200.IntWrapper(0), a.size)
    assert ipos.get() == mid
    c.headlessUncompress(out, ipos, end - mid, rb, mcp_b200.IntWrapper(0), b.size)
    assert ipos.get() == end and np.array_equal(ra, a) and np.array_equal(rb, b)
    ic = mcp_b200.IntCompressor(c)
    w = ic.compress(a)
    assert w[0] == a.size and np.array_equal(w[1:], out[:mid]) and np.array_equal(ic.uncompress(w), a)
    with pytest.raises(mcp_b200.McpError):
        ctx.headless_compress_ints(N.CODEC_FASTPFOR256_VB, a)  # not a skippable id


@pytest.mark.parametrize("big", [False, True], ids=["matrix", "multipage"])
def test_gpu_equals_streams_executed_from_the_reference_java(codec_ctx, big):
    """tests/golden/codec_streams_golden.json holds what the reference's own Java classes write (executed from
    /root/reference by the javamini transpiler): the CUDA encoders must produce those words, the CUDA decoders must read
    them -- for every codec id, with and without Delta, in every kernel family."""
    import goldenstreams as G
    ctx = codec_ctx
    by_cid = {}
    for label, vals, cid, want, dig in G.entries(N.CODEC_COUNT, big=big):
        if big != (len(vals) > 1000000):
            continue
        by_cid.setdefault(cid, []).append((label, vals, want, dig))
    assert len(by_cid) == 2 * (9 if big else N.CODEC_COUNT)  # (the multi-page inputs are about FastPFOR: ids 0..8 in the fixture)
    for cid, items in sorted(by_cid.items()):
        vals, off = batch_of([(lb, v) for lb, v, _, _ in items])
        out_off, out = ctx.codec_encode(cid, N.FRAME_WORDS, vals, off)
        for i, (label, v, want, dig) in enumerate(items):
            G.check(out[out_off[i]: out_off[i + 1]].view(np.int32), want, dig, f"{label} codec {cid:#x}")
        # decode the REFERENCE's streams (where the fixture holds them in full), not our own
        full = [(lb, v, w) for lb, v, w, _ in items if w is not None]
        if full:
            blobs = [w.view(np.uint8) for _, _, w in full]
            in_off = np.zeros(len(full) + 1, np.int64)
            np.cumsum([b.size for b in blobs], out=in_off[1:])
            o_off = np.zeros(len(full) + 1, np.int64)
            np.cumsum([len(v) for _, v, _ in full], out=o_off[1:])
            back = ctx.codec_decode(cid, N.FRAME_WORDS, np.concatenate(blobs) if in_off[-1] else np.zeros(0, np.uint8), in_off, o_off)
            assert np.array_equal(back, np.concatenate([cases.to_i32(v) for _, v, _ in full]))


def test_loader_bytes_of_the_shipped_portfolios_on_the_gpu(ctx):
    """The bytes LoadPortfolios.getCompressedInstruments (executed from the reference's Java) stores for portfolios.json."""
    import hashlib
    import goldenstreams as G
    c1 = json.load(open(os.path.join(GOLD, "c1_sample.json")))
    lp = mcp_b200.LoadPortfolios(ctx)
    for p, rec in zip(c1["portfolios"], G.gold()["loader"]):
        blob = lp.getCompressedInstruments(p["instruments"])
        assert len(blob) == rec["bytes"] and hashlib.sha256(blob).hexdigest() == rec["sha256"]
        assert list(lp.getDecompressedInstruments(blob, rec["n"])) == p["instruments"]


def test_variablebyte_worst_case_fits_the_documented_bound(ctx):
    """mcp_codec_max_bytes must bound VariableByte's 5 bytes per int (values >= 2^28; wrapped deltas)."""
    v = np.full(10000, -1, np.int64)
    for cid in (N.CODEC_VB, N.CODEC_VB | N.FLAG_DELTA):
        words = ctx.compress_ints(cid, v)
        assert np.array_equal(np.asarray(words).view(np.int32), cref.compress(cid, v))


def _awkward(kind, rng):
    """list shapes that stress the flat (thread-per-32-values) kernels: tile boundaries, empty lists, long lists, wide values"""
    if kind == "breach":
        return cases.breach_like_lists(rng, 3000, 10000, 0.01)
    if kind == "tiny":        # many empty / one-value lists: up to 127 lists per tile
        lens = rng.integers(0, 4, 5000)
    elif kind == "mixed":     # 0..250 values: a tile holds 4 000 positions, lists straddle every thread boundary
        lens = rng.integers(0, 250, 800)
    elif kind == "wide":      # values that need 3-5 bytes (negative ints, unsorted data: wrapped deltas)
        lens = rng.integers(0, 120, 1500)
    elif kind == "long":      # lists longer than a tile next to short ones (VariableByte alone: no block limit)
        lens = np.concatenate([rng.integers(0, 60, 300), [5000, 0, 4096, 4095, 1, 9000], rng.integers(0, 60, 100)])
    elif kind == "near128":   # FastPFOR128: a third of the lists reach one block (a page + a VariableByte tail), the rest are all tail
        lens = np.concatenate([rng.integers(60, 160, 900), [128, 127, 129, 255, 256, 257, 0, 1]])
    elif kind == "exact":     # lists of exactly 32 / 64 / 128 values: list starts fall on thread boundaries
        lens = rng.choice([32, 64, 128, 0, 96], 700)
    off = np.zeros(lens.size + 1, np.int64)
    off[1:] = np.cumsum(lens)
    if kind == "wide":
        vals = rng.integers(-2**31, 2**31 - 1, off[-1]).astype(np.int32)
    else:
        vals = np.concatenate([np.sort(rng.integers(0, 50000, int(l))) for l in lens] + [np.zeros(0, np.int64)]).astype(np.int32)
    return vals, off


@pytest.mark.parametrize("kind", ["breach", "tiny", "mixed", "wide", "long", "exact", "near128"])
@pytest.mark.parametrize("shift", [0, 1, 3], ids=["aligned", "shift1", "shift3"])
def test_flat_kernels_on_awkward_shapes(ctx, kind, shift):
    """The flat VariableByte kernels (size pass + tile scan + emit pass; tile-per-CTA decoder) against the oracle, list by
    list, on shapes chosen to hit their edges; device buffers start `shift` ints off a 16-byte boundary (the kernels align
    their windows themselves) and are followed by guard words."""
    rng = np.random.default_rng(1000 + shift)
    vals, off = _awkward(kind, rng)
    nl, total = off.size - 1, int(vals.size)
    codecs = [N.CODEC_VB | N.FLAG_DELTA, N.CODEC_VB]
    if kind not in ("long",):
        codecs += [N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA, N.CODEC_SK_FASTPFOR128_VB | N.FLAG_DELTA, N.CODEC_FASTPFOR256_VB]
    if kind in ("breach", "near128"):  # tiles with a 128-int block among all-VariableByte lists: the k_vb4_pages kernels
        codecs += [N.CODEC_FASTPFOR128_VB | N.FLAG_DELTA, N.CODEC_FASTPFOR128_VB, N.CODEC_SK_FASTPFOR128_VB,
                   N.CODEC_XOR128_IVB, N.CODEC_OPTPFD_VB | N.FLAG_DELTA, N.CODEC_OPTPFD_VB]
    if kind in ("tiny", "wide"):       # (no list reaches a block: the flat kernels alone; `wide` sends tiles to the warp-per-list routines)
        codecs += [N.CODEC_XOR128_IVB, N.CODEC_OPTPFD_VB | N.FLAG_DELTA]
    guard = np.full(256, 0x5AC3A55A, np.uint32)
    ctx.set_option("short_list_codec", 2)
    ctx.set_option("flat_codec", 1)
    try:
        for cid in codecs:
            before = ctx.codec_stats()
            for fr in (N.FRAME_APP, N.FRAME_WORDS):
                want = [cref.get_compressed_instruments(cid, vals[off[i]:off[i + 1]]) if fr == N.FRAME_APP
                        else cref.compress(cid, vals[off[i]:off[i + 1]]).tobytes() for i in range(nl)]
                enc_bytes = sum(len(w) for w in want)
                sizes = {"vals": 4 * total, "off": 8 * (nl + 1), "ooff": 8 * (nl + 1), "enc": enc_bytes, "back": 4 * total}
                bufs = {k: ctx.dev_alloc(v + guard.nbytes + 64) for k, v in sizes.items()}
                sh = {k: (4 * shift if k in ("vals", "enc", "back") else 0) for k in sizes}
                try:
                    for k, v in sizes.items():
                        ctx.h2d(bufs[k] + sh[k] + v, guard)
                    if total:
                        ctx.h2d(bufs["vals"] + sh["vals"], vals)
                    ctx.h2d(bufs["off"], off)
                    got = ctx.codec_encode_dev(cid, fr, bufs["vals"] + sh["vals"], bufs["off"], nl, total, bufs["ooff"],
                                               bufs["enc"] + sh["enc"], enc_bytes)
                    assert got == enc_bytes, (kind, hex(cid), fr)
                    ooff = np.empty(nl + 1, np.int64)
                    ctx.d2h(ooff, bufs["ooff"])
                    out = np.empty(enc_bytes, np.uint8)
                    if enc_bytes:
                        ctx.d2h(out, bufs["enc"] + sh["enc"])
                    assert ooff[0] == 0 and ooff[-1] == enc_bytes
                    for i in range(nl):
                        assert bytes(out[ooff[i]:ooff[i + 1]]) == want[i], (kind, hex(cid), fr, i)
                    ctx.codec_decode_dev(cid, fr, bufs["enc"] + sh["enc"], bufs["ooff"], nl, bufs["back"] + sh["back"], bufs["off"])
                    back = np.empty(total, np.int32)
                    if total:
                        ctx.d2h(back, bufs["back"] + sh["back"])
                    assert np.array_equal(back, vals), (kind, hex(cid), fr)
                    for k, v in sizes.items():
                        g = np.empty_like(guard)
                        ctx.d2h(g, bufs[k] + sh[k] + v)
                        assert np.array_equal(g, guard), f"guard after `{k}` overwritten ({kind}, {hex(cid)}, {fr})"
                finally:
                    for q in bufs.values():
                        ctx.dev_free(q)
            if kind in ("breach", "near128") and (cid & 0xff) in (N.CODEC_FASTPFOR128_VB, N.CODEC_SK_FASTPFOR128_VB, N.CODEC_XOR128_IVB,
                                                                  N.CODEC_OPTPFD_VB):
                after = ctx.codec_stats()  # ... and nothing was handed on to the batched / warp-per-list kernels
                assert after["short_fallbacks"] == before["short_fallbacks"], (hex(cid), before, after)
                assert after["short_calls"] == before["short_calls"] + 4, (hex(cid), before, after)  # 2 framings x (encode + decode)
    finally:
        ctx.set_option("short_list_codec", 1)
        ctx.set_option("flat_codec", FLAT_DEFAULT)
