"""Host-side mirror of the reference's codec interfaces, forwarding to the C ABI.

Same class names, argument order and position-wrapper semantics as
JavaFastPFOR/src/main/java/me/lemire/integercompression/{IntegerCODEC,IntWrapper,Composition,
BinaryPacking,VariableByte,FastPFOR,FastPFOR128,IntCompressor}.java and
differential/{IntegratedComposition,IntegratedBinaryPacking,IntegratedVariableByte,
IntegratedIntCompressor,Delta}.java, so the parity tests read like the reference's JUnit
tests.  Every compress/uncompress call runs the sm_100a kernels through libmcp.so; a
composition this build has no kernel for raises NotImplementedError instead of falling
back to anything.
"""
from __future__ import annotations

import numpy as np

from . import _native as N
from .engine import Context, as_i32

_default_ctx: Context | None = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


class IntWrapper:
    """IntWrapper.java:14-91 -- a mutable int used for in/out positions."""

    def __init__(self, v: int = 0):
        self.value = int(v)

    def get(self) -> int:
        return self.value

    def set(self, v: int):
        self.value = int(v)

    def add(self, v: int):
        self.value += int(v)

    def increment(self):
        self.value += 1

    def intValue(self) -> int:
        return self.value

    def __int__(self):
        return self.value

    def __repr__(self):
        return f"IntWrapper({self.value})"


class _Stage:
    """Marker for a codec building block (only meaningful inside a composition)."""
    name = "?"

    def __repr__(self):
        return self.name


class BinaryPacking(_Stage):
    name = "BinaryPacking"


class VariableByte(_Stage):
    name = "VariableByte"


class FastPFOR(_Stage):
    name = "FastPFOR"


class FastPFOR128(_Stage):
    name = "FastPFOR128"


class IntegratedBinaryPacking(_Stage):
    name = "IntegratedBinaryPacking"


class XorBinaryPacking(_Stage):
    """differential/XorBinaryPacking.java:26 (an IntegratedIntegerCODEC: use it inside IntegratedComposition)."""
    name = "XorBinaryPacking"


class OptPFD(_Stage):
    """OptPFD.java:35."""
    name = "OptPFD"


class IntegratedVariableByte(_Stage):
    name = "IntegratedVariableByte"


class IntegerCODEC:
    """IntegerCODEC.java:36-58 over one codec_id of include/mcp.h."""

    def __init__(self, codec_id: int, ctx: Context | None = None, name: str | None = None):
        self.codec_id = codec_id
        self._ctx = ctx
        self._name = name or f"codec{codec_id}"

    @property
    def ctx(self) -> Context:
        return self._ctx or default_context()

    def compress(self, in_, inpos: IntWrapper, inlength: int, out, outpos: IntWrapper):
        """compress(int[] in, IntWrapper inpos, int inlength, int[] out, IntWrapper outpos).
        Raises CapacityError (an IndexError, like Java's ArrayIndexOutOfBoundsException) if `out` is too small."""
        src = as_i32(in_)[inpos.get(): inpos.get() + inlength]
        room = len(out) - outpos.get()
        words = self.ctx.compress_ints(self.codec_id, src, out_cap=max(room, 0))
        out[outpos.get(): outpos.get() + words.size] = words
        inpos.add(inlength)
        outpos.add(words.size)

    def uncompress(self, in_, inpos: IntWrapper, inlength: int, out, outpos: IntWrapper):
        src = as_i32(in_)[inpos.get(): inpos.get() + inlength]
        room = len(out) - outpos.get()
        vals = self.ctx.uncompress_ints(self.codec_id, src, out_cap=max(room, 0))
        out[outpos.get(): outpos.get() + vals.size] = vals
        inpos.add(inlength)
        outpos.add(vals.size)

    def toString(self) -> str:
        return self._name

    __repr__ = toString


_COMPOSITIONS = {
    ("BinaryPacking", "VariableByte"): N.CODEC_BP32_VB,
    ("FastPFOR", "VariableByte"): N.CODEC_FASTPFOR256_VB,
    ("FastPFOR128", "VariableByte"): N.CODEC_FASTPFOR128_VB,
    ("OptPFD", "VariableByte"): N.CODEC_OPTPFD_VB,
}
_INTEGRATED_COMPOSITIONS = {
    ("IntegratedBinaryPacking", "IntegratedVariableByte"): N.CODEC_IBP32_IVB,
    ("XorBinaryPacking", "IntegratedVariableByte"): N.CODEC_XOR128_IVB,
}


class Composition(IntegerCODEC):
    """Composition.java:36-62.  `delta=True` adds Delta.delta / inverseDelta around it (MCP_FLAG_DELTA)."""

    def __init__(self, f1: _Stage, f2: _Stage, ctx: Context | None = None, delta: bool = False):
        key = (f1.name, f2.name)
        if key not in _COMPOSITIONS:
            raise NotImplementedError(f"no sm_100a kernel for Composition({f1}, {f2}); supported: {sorted(_COMPOSITIONS)}")
        super().__init__(_COMPOSITIONS[key] | (N.FLAG_DELTA if delta else 0), ctx, f"{f1} + {f2}")


class IntegratedComposition(IntegerCODEC):
    """differential/IntegratedComposition.java:41-66."""

    def __init__(self, f1: _Stage, f2: _Stage, ctx: Context | None = None):
        key = (f1.name, f2.name)
        if key not in _INTEGRATED_COMPOSITIONS:
            raise NotImplementedError(f"no sm_100a kernel for IntegratedComposition({f1}, {f2})")
        super().__init__(_INTEGRATED_COMPOSITIONS[key], ctx, f"{f1} + {f2}")


class VariableByteCodec(IntegerCODEC):
    """A bare `new VariableByte()` used as an IntegerCODEC (VariableByte.java:32-36)."""

    def __init__(self, ctx: Context | None = None):
        super().__init__(N.CODEC_VB, ctx, "VariableByte")


_SKIPPABLE = {
    ("BinaryPacking", "VariableByte"): N.CODEC_SK_BP32_VB,
    ("FastPFOR", "VariableByte"): N.CODEC_SK_FASTPFOR256_VB,
    ("FastPFOR128", "VariableByte"): N.CODEC_SK_FASTPFOR128_VB,
    ("IntegratedBinaryPacking", "IntegratedVariableByte"): N.CODEC_SK_IBP32_IVB,
}


class SkippableComposition:
    """SkippableComposition.java:37-54 (and differential/SkippableIntegratedComposition.java:46-71 with initvalue 0) behind
    the SkippableIntegerCODEC interface (SkippableIntegerCODEC.java:43-66): the stream carries no count word, the caller
    supplies `num` when decoding."""

    def __init__(self, f1: _Stage, f2: _Stage, ctx: Context | None = None):
        key = (f1.name, f2.name)
        if key not in _SKIPPABLE:
            raise NotImplementedError(f"no sm_100a kernel for SkippableComposition({f1}, {f2}); supported: {sorted(_SKIPPABLE)}")
        self.codec_id = _SKIPPABLE[key]
        self._ctx = ctx
        self._name = f"{f1}+{f2}"

    @property
    def ctx(self) -> Context:
        return self._ctx or default_context()

    def headlessCompress(self, in_, inpos: IntWrapper, inlength: int, out, outpos: IntWrapper):
        src = as_i32(in_)[inpos.get(): inpos.get() + inlength]
        room = len(out) - outpos.get()
        words = self.ctx.headless_compress_ints(self.codec_id, src, out_cap=max(room, 0))
        out[outpos.get(): outpos.get() + words.size] = words
        inpos.add(inlength)
        outpos.add(words.size)

    def headlessUncompress(self, in_, inpos: IntWrapper, inlength: int, out, outpos: IntWrapper, num: int):
        src = as_i32(in_)[inpos.get(): inpos.get() + inlength]
        if len(out) - outpos.get() < num:
            raise IndexError("out too small")  # Java: ArrayIndexOutOfBoundsException
        vals, used = self.ctx.headless_uncompress_ints(self.codec_id, src, num)
        out[outpos.get(): outpos.get() + num] = vals
        inpos.add(used)
        outpos.add(num)

    def toString(self) -> str:
        return self._name

    __repr__ = toString


class IntegratedVariableByteCodec(IntegerCODEC):
    """A bare `new IntegratedVariableByte()` used as an IntegerCODEC (differential/IntegratedVariableByte.java:35-76,114-141):
    VariableByte over the gaps from 0 -- byte for byte what VariableByte writes after Delta.delta."""

    def __init__(self, ctx: Context | None = None):
        super().__init__(N.CODEC_VB | N.FLAG_DELTA, ctx, "IntegratedVariableByte")


class IntCompressor:
    """IntCompressor.java:27-61: default codec SkippableComposition(BinaryPacking, VariableByte), or the one given
    (IntCompressor(SkippableIntegerCODEC c), :27-29)."""
    _codec = N.CODEC_SK_BP32_VB

    def __init__(self, codec: SkippableComposition | None = None, ctx: Context | None = None):
        if isinstance(codec, Context):  # (older call sites passed the context first)
            codec, ctx = None, codec
        if codec is not None:
            if codec.codec_id == N.CODEC_SK_IBP32_IVB:
                raise NotImplementedError("integrated compositions belong to IntegratedIntCompressor")
            self._codec = codec.codec_id
            ctx = ctx or codec._ctx
        self._ctx = ctx

    @property
    def ctx(self) -> Context:
        return self._ctx or default_context()

    def compress(self, input_) -> np.ndarray:
        return self.ctx.compress_ints(self._codec, input_)

    def uncompress(self, compressed) -> np.ndarray:
        c = as_i32(compressed)
        if c.size == 0:
            raise IndexError("compressed[0] missing")
        return self.ctx.uncompress_ints(self._codec, c, out_cap=int(c[0]))


class IntegratedIntCompressor(IntCompressor):
    """differential/IntegratedIntCompressor.java:38-62 (sorted lists)."""
    _codec = N.CODEC_SK_IBP32_IVB
