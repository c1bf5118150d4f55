"""Pins the CPU oracle (oracle/codec_oracle.c) to the reference.

1. the one byte-level known-answer vector the reference holds for this path
   (setup.txt:47-48 <-> portfolios.json:1);
2. bit-packing golden vectors produced by interpreting the reference's own
   BitPacking.java / IntegratedBitPacking.java (tests/golden/make_bitpacking_golden.py);
3. the reference's JUnit round-trip / size-bound / edge-case matrix.
"""
import base64
import json
import os

import numpy as np
import pytest

import cases
from oracle import cref

GOLD = os.path.join(os.path.dirname(__file__), "golden")

# setup.txt:47-48 (document dump of portfolio _id "0") and portfolios.json:1
KAT_INTS = [719, 5, 3990, 4112, 3328, 78, 1163, 135, 263, 67, 67, 1860, 75, 1342, 1145, 38, 1216]
KAT_B64 = "AAAAABaFhU8AoBCfiQvOmoIHgQeORMPDeYo+y4lApogAAAAA"


def test_kat_setup_txt_encode_and_decode():
    blob = base64.b64decode(KAT_B64)
    assert len(blob) == 36
    assert cref.get_compressed_instruments(cref.CODEC_BP32_VB, KAT_INTS) == blob
    assert list(cref.get_decompressed_instruments(cref.CODEC_BP32_VB, blob, 17)) == KAT_INTS
    assert len(blob) < 4 * 17  # LoadPortfolios.java:254: the compressed form is the one stored


def test_kat_words():
    w = cref.compress(cref.CODEC_BP32_VB, KAT_INTS).view(np.uint32)
    assert [hex(x) for x in w] == ["0x0", "0x1685854f", "0xa0109f", "0x890bce9a", "0x82078107", "0x8e44c3c3",
                                   "0x798a3ecb", "0x8940a688"]


@pytest.fixture(scope="module")
def golden():
    return json.load(open(os.path.join(GOLD, "bitpacking_golden.json")))


def test_golden_fastpack(golden):
    for c in golden["fastpack"]:
        assert list(cref.fastpack(np.array(c["in"], np.uint32), c["b"], True)) == c["out"], c["b"]


def test_golden_fastpackwithoutmask(golden):
    for c in golden["fastpackwithoutmask"]:
        assert list(cref.fastpack(np.array(c["in"], np.uint32), c["b"], False)) == c["out"], c["b"]


def test_golden_fastunpack(golden):
    for c in golden["fastunpack"]:
        assert list(cref.fastunpack(c["in"], c["b"])) == c["out"], c["b"]


def test_golden_integratedpack(golden):
    for c in golden["integratedpack"]:
        assert list(cref.integratedpack(c["init"], np.array(c["in"], np.uint32), c["b"])) == c["out"], c["b"]


def test_golden_integratedunpack(golden):
    for c in golden["integratedunpack"]:
        assert list(cref.integratedunpack(c["init"], c["in"], c["b"])) == c["out"], c["b"]


def test_bits():
    assert cref.bits(0) == 0 and cref.bits(1) == 1 and cref.bits(255) == 8 and cref.bits(256) == 9
    assert cref.bits(-1) == 32  # negative Java ints need 32 bits (Util.java:101-103)


@pytest.mark.parametrize("codec", cases.ALL_CODECS, ids=cases.NAMES)
@pytest.mark.parametrize("delta", [0, cases.DELTA], ids=["plain", "delta"])
def test_roundtrip_reference_matrix(codec, delta):
    for label, v in cases.reference_matrix(big=False):
        w = cref.compress(codec | delta, v)
        d = cref.uncompress(codec | delta, w, len(v))
        assert np.array_equal(d, cases.to_i32(v)), label


@pytest.mark.parametrize("codec", [cref.CODEC_FASTPFOR256_VB, cref.CODEC_FASTPFOR128_VB, cref.CODEC_SK_IBP32_IVB])
def test_roundtrip_multipage(codec):
    for label, v in cases.reference_matrix(big=True)[-2:]:  # 1 333 333 ints (BasicTest.java:505), 2 342 351 sorted (IntCompressorTest.java:66)
        w = cref.compress(codec, v)
        assert np.array_equal(cref.uncompress(codec, w, len(v)), cases.to_i32(v)), label


def test_saul_offsets():
    # BasicTest.java:47-70: {2,3,4,5} written at output offset x must decode; our C API is offset-free
    # (the caller passes out+x), so the property is that the words do not depend on x
    w0 = cref.compress(cref.CODEC_BP32_VB, [2, 3, 4, 5])
    assert list(cref.uncompress(cref.CODEC_BP32_VB, w0, 4)) == [2, 3, 4, 5]
    assert w0[0] == 0  # Composition marker: BinaryPacking wrote nothing for n < 32 (Composition.java:45-48)


def test_empty_and_spurious_output():
    # BasicTest.java:325-396 zeroinzerouttest / spuriousouttest
    for codec in cases.ALL_CODECS:
        w = cref.compress(codec, [])
        if codec in (cref.CODEC_SK_IBP32_IVB, cref.CODEC_SK_BP32_VB, cref.CODEC_SK_FASTPFOR256_VB, cref.CODEC_SK_FASTPFOR128_VB):
            assert list(w) == [0]  # IntCompressor: compressed[0] = input.length
        else:
            assert w.size == 0
    # block codecs emit nothing for inlength < block: the bare FastPFOR instance
    f = cref.FastPFOR(256)
    assert f.compress(np.arange(255)).size == 0
    f128 = cref.FastPFOR(128)
    assert f128.compress(np.arange(127)).size == 0


def test_boundary_sizes():
    # BoundaryTest.java:19-90: compressed length <= input length for these; AdhocTest / TestUtils.java:97-124: fits in n+1
    for n in (31, 32, 33, 127, 128, 129, 255, 256, 257):
        v = np.arange(n)
        for codec in (cref.CODEC_BP32_VB, cref.CODEC_IBP32_IVB):
            assert cref.compress(codec, v).size <= n
    for v in ([0, 16384], [65535, 65535]):
        for codec in (cref.CODEC_FASTPFOR256_VB, cref.CODEC_FASTPFOR128_VB):
            assert cref.compress(codec, v).size <= len(v) + 1
    assert cref.compress(cref.CODEC_VB, [-1]).size <= 2


def test_binarypacking_layout_by_hand():
    # 32 values of 5 bits then one VByte value: header 32, width word 5, 5 packed words, 1 VByte word
    v = list(range(32)) + [300]
    w = cref.compress(cref.CODEC_BP32_VB, v).view(np.uint32)
    assert w[0] == 32 and w[1] == 5 and w.size == 2 + 5 + 1
    stream = sum(int(x) << (5 * i) for i, x in enumerate(range(32)))
    assert [int(x) for x in w[2:7]] == [(stream >> (32 * j)) & 0xFFFFFFFF for j in range(5)]
    assert int(w[7]) == (300 & 127) | (((300 >> 7) | 128) << 8)  # low group first, high bit on the LAST byte
    # 128 values: one header word holds four widths, MSB first (BinaryPacking.java:64-65)
    v = [1] * 32 + [3] * 32 + [7] * 32 + [255] * 32
    w = cref.compress(cref.CODEC_BP32_VB, v).view(np.uint32)
    assert w[0] == 128 and w[1] == (1 << 24) | (2 << 16) | (3 << 8) | 8


def test_fastpfor_layout_by_hand():
    # one block, all zeros but data[126] = -1: b = 0, one 32-bit exception (BasicTest.java:627-664)
    d = np.zeros(256, np.int64)
    d[126] = -1
    w = cref.compress(cref.CODEC_FASTPFOR256_VB, d).view(np.uint32)
    # [256][where_meta=1][bytesize=4][b=0,c=1,maxb=32,pos=126][bitmap bit31][count=1][0xffffffff]
    assert list(w) == [256, 1, 4, 0 | (1 << 8) | (32 << 16) | (126 << 24), 1 << 31, 1, 0xFFFFFFFF]


def test_fastpfor_cost_ties_keep_larger_b():
    # 255 zeros and a single 1: maxb=1.  b=0 costs 1*8+1*1+0+8-1=16 > 256?  no: 16 < 256 -> b=0 wins
    d = np.zeros(256, np.int64)
    d[3] = 1
    w = cref.compress(cref.CODEC_FASTPFOR256_VB, d).view(np.uint32)
    assert w[3] & 0xFF == 0 and (w[3] >> 8) & 0xFF == 1 and (w[3] >> 16) & 0xFF == 1
    assert w[2] == 4 and w[4] == 0  # index-1 exceptions are implicit: empty bitmap, nothing stored


def test_fastpfor_stale_exception_buffer():
    """FastPFOR never clears dataTobePacked (FastPFOR.java:143): a reused instance leaks the previous
    call's exception values into the padding bits of the next call's last packed group."""
    a = np.zeros(256, np.int64)
    a[:40] = (1 << 20) - 1           # 40 exceptions of 20 bits: stream 20 buffer holds 40 stale values
    b = np.zeros(256, np.int64)
    b[:3] = (1 << 20) - 7            # 3 exceptions: the group is padded with stale entries 3..31
    fresh = cref.compress(cref.CODEC_FASTPFOR256_VB, b)
    inst = cref.FastPFOR(256)
    inst.compress(a)
    reused = inst.compress(b)
    assert reused.size == fresh.size - 0 and not np.array_equal(reused, fresh[:reused.size])
    # both decode to the same integers: decoders ignore those bits
    out, _ = cref.FastPFOR(256).uncompress(reused, 256)
    assert np.array_equal(out, cases.to_i32(b))


def test_integrated_tail_restart_quirk():
    # IntegratedComposition restarts the VByte tail's delta base at 0 (IntegratedVariableByte.java:40);
    # the skippable composition chains it (SkippableIntegratedComposition.java:52,58)
    v = np.arange(1000, 1000 + 33)
    w = cref.compress(cref.CODEC_IBP32_IVB, v).view(np.uint32)
    sk = cref.compress(cref.CODEC_SK_IBP32_IVB, v).view(np.uint32)
    assert w[-1] == (1032 & 127) | (((1032 >> 7) | 128) << 8)   # raw 1032, two bytes
    assert sk[-1] == (1 | 128)                                    # gap 1, one byte
    assert sk[0] == 33


def test_integrated_b32_stores_raw():
    v = np.array([0, -1] * 16)  # deltas need 32 bits -> block stored raw (IntegratedBitPacking.java:1446-1451)
    w = cref.compress(cref.CODEC_IBP32_IVB, v).view(np.uint32)
    assert w[0] == 32 and w[1] == 32 and np.array_equal(w[2:34], cases.to_i32(v).view(np.uint32))


def test_app_framing_and_none_fallback():
    # LoadPortfolios.java:622: outpos+1 ints, big-endian; empty list -> 4 zero bytes, not smaller than 0 -> "none"
    assert cref.get_compressed_instruments(cref.CODEC_BP32_VB, []) == b"\0\0\0\0"
    assert cref.integers_to_bytes([1, -2]) == b"\x00\x00\x00\x01\xff\xff\xff\xfe"
    rng = np.random.default_rng(1)
    v = rng.integers(-2**31, 2**31, 64)
    blob = cref.get_compressed_instruments(cref.CODEC_BP32_VB, v)
    assert len(blob) >= 4 * 64  # incompressible: the loader would store integersToBytes() with algo "none"
    assert np.array_equal(cref.get_decompressed_instruments(cref.CODEC_BP32_VB, blob, 64), cases.to_i32(v))


def test_batch_matches_single():
    rng = np.random.default_rng(5)
    vals, off = cases.breach_like_lists(rng, 200, 3000, 0.02)
    for codec in (cref.CODEC_BP32_VB, cref.CODEC_FASTPFOR128_VB | cases.DELTA, cref.CODEC_SK_IBP32_IVB):
        for framing in (0, 1):
            buf, lens = cref.batch_compress(codec, framing, vals, off, 4 * 400)
            for i in (0, 17, 199):
                one = vals[off[i]:off[i + 1]]
                ref = cref.get_compressed_instruments(codec, one) if framing else cref.compress(codec, one).tobytes()
                assert bytes(buf[i * 1600: i * 1600 + lens[i]]) == ref
            back = cref.batch_uncompress(codec, framing, buf, 1600, lens, off)
            assert np.array_equal(back, vals)


@pytest.mark.parametrize("sk,comp", [(cref.CODEC_SK_FASTPFOR256_VB, cref.CODEC_FASTPFOR256_VB), (cref.CODEC_SK_FASTPFOR128_VB, cref.CODEC_FASTPFOR128_VB),
                                     (cref.CODEC_SK_BP32_VB, cref.CODEC_BP32_VB)])
def test_skippable_stream_is_the_composition_stream_without_its_header(sk, comp):
    """IntCompressor over SkippableComposition(F1, VariableByte) writes [n][headless F1][headless VB]
    (IntCompressor.java:38-46, SkippableComposition.java:37-45); Composition(F1, VariableByte) writes [F1's count word or the
    0 marker][the same two headless streams] (Composition.java:37-51, FastPFOR.java:309-317, BinaryPacking.java:43-51).  So
    the two differ in word 0 only -- which ties the skippable framing to the KAT-pinned composition."""
    rng = np.random.default_rng(3)
    for n in (0, 1, 31, 32, 100, 127, 128, 129, 255, 256, 257, 1000, 4096, 70000):
        v = np.cumsum(rng.geometric(0.02, n)).astype(np.int32)
        if n > 300:
            v[rng.integers(0, n, 5)] = -7  # 32-bit exceptions
        a, b = cref.compress(sk, v), cref.compress(comp, v)
        assert a[0] == n
        if n == 0:
            assert a.size == 1 and b.size == 0
        else:
            assert np.array_equal(a[1:], b[1:]), n
        assert np.array_equal(cref.uncompress(sk, a, n), v)
