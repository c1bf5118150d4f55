#!/usr/bin/env python3
"""The reference's codecs, executed from its own Java sources through javamini (no JVM here).

Runs only where /root/reference exists (the build container).  Everything below calls the
*reference's* classes: this file only spells out, per codec id of include/mcp.h, the Java
expression a user of the reference would write.
"""
from __future__ import annotations

import os

import javamini as J

REF = "/root/reference"
LIB = REF + "/JavaFastPFOR/src/main/java/me/lemire/integercompression/"
APP = REF + "/maprdb-portfolio/src/main/java/com/mapr/db/data/"
LIB_FILES = [
    "IntWrapper.java", "IntegerCODEC.java", "SkippableIntegerCODEC.java", "ByteIntegerCODEC.java", "Util.java",
    "BitPacking.java", "BinaryPacking.java", "VariableByte.java", "Composition.java", "SkippableComposition.java",
    "FastPFOR.java", "FastPFOR128.java", "IntCompressor.java", "S16.java", "OptPFD.java",
    "differential/IntegratedIntegerCODEC.java", "differential/SkippableIntegratedIntegerCODEC.java",
    "differential/IntegratedByteIntegerCODEC.java", "differential/Delta.java", "differential/IntegratedBitPacking.java",
    "differential/IntegratedBinaryPacking.java", "differential/IntegratedVariableByte.java",
    "differential/IntegratedComposition.java", "differential/SkippableIntegratedComposition.java",
    "differential/IntegratedIntCompressor.java", "differential/XorBinaryPacking.java",
]
CODEC_SIG = ("int[]", "IntWrapper", "int", "int[]", "IntWrapper")
DELTA = 0x100

_prog = None


def program() -> J.Program:
    global _prog
    if _prog is None:
        if not os.path.isdir(REF):
            raise RuntimeError("needs /root/reference (build container only)")
        # the loader's own two methods + its codec field, lifted verbatim out of the MapR-DB glue around them
        loader = J.extract_members(APP + "LoadPortfolios.java",
                                   ["integerCodec", "getCompressedInstruments", "getDecompressedInstruments"], "LoadPortfolios")
        _prog = J.Program([LIB + f for f in LIB_FILES] + [APP + "TypeSize.java", APP + "BitManipulationHelper.java"],
                          [("LoadPortfolios.java (extract)", loader)])
    return _prog


def w32(v):
    return ((int(v) + 0x80000000) & 0xFFFFFFFF) - 0x80000000


def new_codec(cid: int):
    """The Java object behind codec id `cid` (include/mcp.h) -- always a FRESH instance, like the parity contract says."""
    P = program()
    n = P.new
    if cid == 0:
        return n("Composition", n("BinaryPacking"), n("VariableByte"))
    if cid == 1:
        return n("Composition", n("FastPFOR"), n("VariableByte"))
    if cid == 2:
        return n("Composition", n("FastPFOR128"), n("VariableByte"))
    if cid == 3:
        return n("IntegratedComposition", n("IntegratedBinaryPacking"), n("IntegratedVariableByte"))
    if cid == 4:
        return n("IntegratedIntCompressor")
    if cid == 5:
        return n("VariableByte")
    if cid == 6:
        return n("IntCompressor")
    if cid == 7:
        return n("IntCompressor", n("SkippableComposition", n("FastPFOR"), n("VariableByte")))
    if cid == 8:
        return n("IntCompressor", n("SkippableComposition", n("FastPFOR128"), n("VariableByte")))
    if cid == 9:
        return n("IntegratedComposition", n("XorBinaryPacking"), n("IntegratedVariableByte"))
    if cid == 10:
        return n("Composition", n("OptPFD"), n("VariableByte"))
    raise ValueError(cid)


def compress_with(codec, cid: int, vals) -> list:
    """codec.compress(in, 0, n, out, 0) (IntegerCODEC) or codec.compress(int[]) (the IntCompressor family)."""
    P = program()
    data = [w32(v) for v in vals]
    if cid & DELTA:
        P.Delta.method = None
        P.method(P.Delta, "delta", "int[]")(data)
    base = cid & 0xFF
    if base in (4, 6, 7, 8):
        return list(P.method(codec, "compress", "int[]")(data))
    out = [0] * (4 * len(data) + 1024)
    inpos, outpos = P.new("IntWrapper", 0), P.new("IntWrapper", 0)
    P.method(codec, "compress", *CODEC_SIG)(data, inpos, len(data), out, outpos)
    assert inpos.get() == len(data) or len(data) == 0, (inpos.get(), len(data))
    return out[:outpos.get()]


def compress(cid: int, vals) -> list:
    return compress_with(new_codec(cid & 0xFF), cid, vals)


def uncompress(cid: int, words, n: int) -> list:
    P = program()
    base = cid & 0xFF
    codec = new_codec(base)
    if base in (4, 6, 7, 8):
        out = list(P.method(codec, "uncompress", "int[]")(list(words)))
    else:
        out = [0] * n
        inpos, outpos = P.new("IntWrapper", 0), P.new("IntWrapper", 0)
        P.method(codec, "uncompress", *CODEC_SIG)(list(words), inpos, len(words), out, outpos)
        out = out[:outpos.get()]
    if cid & DELTA:
        P.method(P.Delta, "inverseDelta", "int[]")(out)
    return out


def get_compressed_instruments(vals) -> bytes:
    """LoadPortfolios.getCompressedInstruments (LoadPortfolios.java:608-623) on a fresh loader object."""
    P = program()
    ld = P.new("LoadPortfolios")
    lst = J.ArrayList(w32(v) for v in vals)
    b = getattr(ld, J.mangle("getCompressedInstruments", ["List<int>"]))(lst)
    return bytes(v & 0xFF for v in b)


def get_decompressed_instruments(blob: bytes, n: int) -> list:
    P = program()
    ld = P.new("LoadPortfolios")
    sb = [v - 256 if v > 127 else v for v in blob]
    return list(getattr(ld, J.mangle("getDecompressedInstruments", ["byte[]", "int"]))(sb, n))
