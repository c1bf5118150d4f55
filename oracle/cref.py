"""ctypes binding of the C oracle (oracle/_build/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

CODEC_BP32_VB = 0
CODEC_FASTPFOR256_VB = 1
CODEC_FASTPFOR128_VB = 2
CODEC_IBP32_IVB = 3
CODEC_SK_IBP32_IVB = 4
CODEC_VB = 5
CODEC_SK_BP32_VB = 6
CODEC_SK_FASTPFOR256_VB = 7
CODEC_SK_FASTPFOR128_VB = 8
CODEC_XOR128_IVB = 9
CODEC_OPTPFD_VB = 10
FLAG_DELTA = 0x100


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc, a few seconds)."""
    srcs = [os.path.join(_HERE, f) for f in ("codec_oracle.c", "risk_oracle.c", "risk_fast.c", "weights_oracle.c", "oracle.h", "Makefile")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        i32p, u32p, i64p, u8p, f64p = (C.POINTER(t) for t in (C.c_int32, C.c_uint32, C.c_int64, C.c_uint8, C.c_double))
        L.orc_bits.restype = C.c_int
        L.orc_bits.argtypes = [C.c_uint32]
        for name in ("orc_fastpack", "orc_fastpackwithoutmask", "orc_fastunpack"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [u32p, u32p, C.c_int]
        for name in ("orc_integratedpack", "orc_integratedunpack"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [C.c_uint32, u32p, u32p, C.c_int]
        L.orc_compress.restype = C.c_int64
        L.orc_compress.argtypes = [C.c_int, i32p, C.c_int64, i32p]
        L.orc_uncompress.restype = C.c_int64
        L.orc_uncompress.argtypes = [C.c_int, i32p, C.c_int64, i32p, C.c_int64]
        L.orc_max_compressed_words.restype = C.c_int64
        L.orc_max_compressed_words.argtypes = [C.c_int64]
        L.orc_fastpfor_new.restype = C.c_void_p
        L.orc_fastpfor_new.argtypes = [C.c_int]
        L.orc_fastpfor_free.restype = None
        L.orc_fastpfor_free.argtypes = [C.c_void_p]
        L.orc_fastpfor_compress.restype = C.c_int64
        L.orc_fastpfor_compress.argtypes = [C.c_void_p, i32p, C.c_int64, i32p]
        L.orc_fastpfor_uncompress.restype = C.c_int64
        L.orc_fastpfor_uncompress.argtypes = [C.c_void_p, i32p, C.c_int64, i32p, i64p]
        L.orc_get_compressed_instruments.restype = C.c_int64
        L.orc_get_compressed_instruments.argtypes = [C.c_int, i32p, C.c_int64, u8p]
        L.orc_get_decompressed_instruments.restype = C.c_int64
        L.orc_get_decompressed_instruments.argtypes = [C.c_int, u8p, C.c_int64, i32p, C.c_int64]
        L.orc_integers_to_bytes.restype = None
        L.orc_integers_to_bytes.argtypes = [i32p, C.c_int64, u8p]
        L.orc_batch_compress.restype = C.c_int
        L.orc_batch_compress.argtypes = [C.c_int, C.c_int, i32p, i64p, C.c_int64, u8p, C.c_int64, i64p, C.c_int]
        L.orc_batch_uncompress.restype = C.c_int
        L.orc_batch_uncompress.argtypes = [C.c_int, C.c_int, u8p, C.c_int64, i64p, C.c_int64, i32p, i64p, C.c_int]
        L.orc_value.restype = None
        L.orc_build_exposures.argtypes = [i32p, f64p, i64p, C.c_int64, f64p, C.c_int64, C.c_int32, f64p]
        L.orc_build_exposures.restype = C.c_int
        L.orc_value.argtypes = [f64p, f64p, C.c_int64, C.c_int64, C.c_int32, f64p, C.c_int]
        L.orc_risk.restype = C.c_int64
        L.orc_risk.argtypes = [f64p, f64p, C.c_int64, C.c_int64, C.c_int32, C.c_double, C.c_int64, C.c_int64,
                               f64p, f64p, f64p, i32p, f64p, i32p, C.c_int]
        L.orc_risk_fast.restype = C.c_int64
        L.orc_risk_fast.argtypes = L.orc_risk.argtypes
        L.orc_value_fast.restype = None
        L.orc_value_fast.argtypes = L.orc_value.argtypes
        L.orc_fast_isa.restype = C.c_char_p
        L.orc_tail_size.restype = C.c_int64
        L.orc_tail_size.argtypes = [C.c_int64, C.c_double]
        L.orc_philox4x32_10.restype = None
        L.orc_philox4x32_10.argtypes = [u32p, u32p, u32p]
        L.orc_gen_normal.restype = None
        L.orc_gen_normal.argtypes = [f64p, C.c_int64, C.c_int32, C.c_uint64, C.c_uint32, C.c_double, C.c_double,
                                     C.c_int64, C.c_int]
        L.orc_gen_breach_list.restype = C.c_int64
        L.orc_gen_breach_list.argtypes = [i32p, C.c_int64, C.c_int32, C.c_double, C.c_uint64]
        L.orc_num_threads.restype = C.c_int
        L.orc_snappy_compress.restype = C.c_int64
        L.orc_snappy_compress.argtypes = [u8p, C.c_int64, u8p]
        L.orc_snappy_uncompress.restype = C.c_int64
        L.orc_snappy_uncompress.argtypes = [u8p, C.c_int64, u8p, C.c_int64]
        L.orc_doubles_to_bytes.restype = None
        L.orc_doubles_to_bytes.argtypes = [f64p, C.c_int64, u8p]
        _lib = L
    return _lib


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))


def _i32(a) -> np.ndarray:
    """Java int[] semantics: accept signed or unsigned 32-bit values."""
    a = np.asarray(a)
    if a.dtype == np.int32:
        return np.ascontiguousarray(a)
    return np.ascontiguousarray(a.astype(np.int64) & 0xFFFFFFFF).astype(np.uint32).view(np.int32)


# ----------------------------------------------------------------- primitives
def bits(v: int) -> int:
    return lib().orc_bits(v & 0xFFFFFFFF)


def fastpack(vals, b: int, masked: bool = True) -> np.ndarray:
    v = _i32(vals).view(np.uint32)
    assert v.size == 32
    out = np.zeros(max(b, 1), np.uint32)
    (lib().orc_fastpack if masked else lib().orc_fastpackwithoutmask)(_p(v, C.c_uint32), _p(out, C.c_uint32), b)
    return out[:b]


def fastunpack(words, b: int) -> np.ndarray:
    w = np.ascontiguousarray(np.asarray(words, np.uint32))
    if w.size == 0:
        w = np.zeros(1, np.uint32)
    out = np.zeros(32, np.uint32)
    lib().orc_fastunpack(_p(w, C.c_uint32), _p(out, C.c_uint32), b)
    return out


def integratedpack(init: int, vals, b: int) -> np.ndarray:
    v = _i32(vals).view(np.uint32)
    out = np.zeros(max(b, 1), np.uint32)
    lib().orc_integratedpack(init & 0xFFFFFFFF, _p(v, C.c_uint32), _p(out, C.c_uint32), b)
    return out[:b]


def integratedunpack(init: int, words, b: int) -> np.ndarray:
    w = np.ascontiguousarray(np.asarray(words, np.uint32))
    if w.size == 0:
        w = np.zeros(1, np.uint32)
    out = np.zeros(32, np.uint32)
    lib().orc_integratedunpack(init & 0xFFFFFFFF, _p(w, C.c_uint32), _p(out, C.c_uint32), b)
    return out


# --------------------------------------------------------------------- codecs
def compress(codec: int, vals) -> np.ndarray:
    """codec.compress(in, 0, n, out, 0) on a fresh instance -> int32 words."""
    v = _i32(vals)
    out = np.zeros(lib().orc_max_compressed_words(v.size), np.int32)
    n = lib().orc_compress(codec, _p(v, C.c_int32), v.size, _p(out, C.c_int32))
    if n < 0:
        raise ValueError("bad codec")
    return out[:n].copy()


def uncompress(codec: int, words, out_cap: int) -> np.ndarray:
    w = _i32(words)
    out = np.zeros(max(out_cap, 1), np.int32)
    n = lib().orc_uncompress(codec, _p(w, C.c_int32), w.size, _p(out, C.c_int32), out_cap)
    if n < 0:
        raise ValueError("decode error")
    return out[:n].copy()


def get_compressed_instruments(codec: int, vals) -> bytes:
    v = _i32(vals)
    buf = np.zeros(4 * (lib().orc_max_compressed_words(v.size) + 1), np.uint8)
    n = lib().orc_get_compressed_instruments(codec, _p(v, C.c_int32), v.size, _p(buf, C.c_uint8))
    return bytes(buf[:n])


def get_decompressed_instruments(codec: int, blob: bytes, n: int) -> np.ndarray:
    b = np.frombuffer(blob, np.uint8).copy() if len(blob) else np.zeros(1, np.uint8)
    out = np.zeros(max(n, 1), np.int32)
    r = lib().orc_get_decompressed_instruments(codec, _p(b, C.c_uint8), len(blob), _p(out, C.c_int32), n)
    if r < 0:
        raise ValueError("decode error")
    return out[:r].copy()


def integers_to_bytes(vals) -> bytes:
    v = _i32(vals)
    buf = np.zeros(4 * max(v.size, 1), np.uint8)
    lib().orc_integers_to_bytes(_p(v, C.c_int32), v.size, _p(buf, C.c_uint8))
    return bytes(buf[: 4 * v.size])


class FastPFOR:
    """A long-lived FastPFOR instance (stale exception buffers persist across calls)."""

    def __init__(self, block: int = 256):
        self._h = lib().orc_fastpfor_new(block)
        self.block = block

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_fastpfor_free(self._h)
            self._h = None

    def compress(self, vals) -> np.ndarray:
        v = _i32(vals)
        out = np.zeros(lib().orc_max_compressed_words(v.size), np.int32)
        n = lib().orc_fastpfor_compress(self._h, _p(v, C.c_int32), v.size, _p(out, C.c_int32))
        return out[:n].copy()

    def uncompress(self, words, n: int):
        w = _i32(words)
        out = np.zeros(max(n, 1), np.int32)
        consumed = C.c_int64(0)
        r = lib().orc_fastpfor_uncompress(self._h, _p(w, C.c_int32), w.size, _p(out, C.c_int32), C.byref(consumed))
        return out[:r].copy(), consumed.value


def batch_compress(codec: int, framing: int, vals: np.ndarray, list_off: np.ndarray, stride: int, nthreads: int = 0):
    v = _i32(vals)
    off = np.ascontiguousarray(list_off, np.int64)
    nl = off.size - 1
    out = np.zeros(nl * stride, np.uint8)
    lens = np.zeros(nl, np.int64)
    nt = nthreads or lib().orc_num_threads()
    rc = lib().orc_batch_compress(codec, framing, _p(v, C.c_int32), _p(off, C.c_int64), nl, _p(out, C.c_uint8), stride,
                                  _p(lens, C.c_int64), nt)
    if rc != 0:
        raise ValueError("batch compress failed (stride too small?)")
    return out, lens


def batch_uncompress(codec: int, framing: int, buf: np.ndarray, stride: int, lens: np.ndarray, out_off: np.ndarray,
                     nthreads: int = 0) -> np.ndarray:
    off = np.ascontiguousarray(out_off, np.int64)
    lens = np.ascontiguousarray(lens, np.int64)
    out = np.zeros(max(int(off[-1]), 1), np.int32)
    nt = nthreads or lib().orc_num_threads()
    rc = lib().orc_batch_uncompress(codec, framing, _p(buf, C.c_uint8), stride, _p(lens, C.c_int64), lens.size,
                                    _p(out, C.c_int32), _p(off, C.c_int64), nt)
    if rc != 0:
        raise ValueError("batch uncompress failed")
    return out[: int(off[-1])]


# ----------------------------------------------------------------------- risk
def build_exposures(instruments, weights, pos_off, loadings) -> np.ndarray:
    """E[p][f] = fma chain over the portfolio's positions in file order of weight * loadings[instrument][f]."""
    inst = np.ascontiguousarray(instruments, np.int32)
    w = np.ascontiguousarray(weights, np.float64)
    off = np.ascontiguousarray(pos_off, np.int64)
    B = np.ascontiguousarray(loadings, np.float64)
    P, F = off.size - 1, B.shape[1]
    E = np.empty((P, F), np.float64)
    if lib().orc_build_exposures(_p(inst, C.c_int32), _p(w, C.c_double), _p(off, C.c_int64), P, _p(B, C.c_double), B.shape[0], F,
                                 _p(E, C.c_double)) != 0:
        raise ValueError("instrument id outside the loading table")
    return E


def value(E: np.ndarray, S: np.ndarray, nthreads: int = 0, fast: bool = False) -> np.ndarray:
    """fast=True: the SIMD baseline (risk_fast.c) instead of the literal scalar chain -- bit-identical by construction."""
    E = np.ascontiguousarray(E, np.float64)
    S = np.ascontiguousarray(S, np.float64)
    P, F = E.shape
    N = S.shape[0]
    assert S.shape[1] == F
    V = np.empty((P, N), np.float64)
    (lib().orc_value_fast if fast else lib().orc_value)(_p(E, C.c_double), _p(S, C.c_double), P, N, F, _p(V, C.c_double), nthreads or lib().orc_num_threads())
    return V


def tail_size(N: int, alpha: float) -> int:
    return lib().orc_tail_size(N, alpha)


def fast_isa() -> str:
    return lib().orc_fast_isa().decode()


def risk(E: np.ndarray, S: np.ndarray, alpha: float, p0: int = 0, p1: int | None = None, nthreads: int = 0,
         fast: bool = False) -> dict:
    E = np.ascontiguousarray(E, np.float64)
    S = np.ascontiguousarray(S, np.float64)
    P, F = E.shape
    N = S.shape[0]
    p1 = P if p1 is None else p1
    q = p1 - p0
    t = tail_size(N, alpha)
    res = {
        "mean": np.empty(q), "variance": np.empty(q), "var_loss": np.empty(q),
        "var_trial": np.empty(q, np.int32), "cvar": np.empty(q), "breach": np.empty((q, t), np.int32),
    }
    (lib().orc_risk_fast if fast else lib().orc_risk)(_p(E, C.c_double), _p(S, C.c_double), P, N, F, alpha, p0, p1,
                   _p(res["mean"], C.c_double), _p(res["variance"], C.c_double), _p(res["var_loss"], C.c_double),
                   _p(res["var_trial"], C.c_int32), _p(res["cvar"], C.c_double), _p(res["breach"], C.c_int32),
                   nthreads or lib().orc_num_threads())
    res["tail"] = t
    return res


def philox(ctr, key) -> np.ndarray:
    c = np.asarray(ctr, np.uint32)
    k = np.asarray(key, np.uint32)
    o = np.zeros(4, np.uint32)
    lib().orc_philox4x32_10(_p(c, C.c_uint32), _p(k, C.c_uint32), _p(o, C.c_uint32))
    return o


def gen_normal(rows: int, cols: int, seed: int, tensor_id: int, scale: float = 1.0, shift: float = 0.0,
               row0: int = 0, nthreads: int = 0) -> np.ndarray:
    out = np.empty((rows, cols), np.float64)
    lib().orc_gen_normal(_p(out, C.c_double), rows, cols, seed, tensor_id, scale, shift, row0,
                         nthreads or lib().orc_num_threads())
    return out


def gen_breach_list(list_id: int, N: int, rate: float, seed: int) -> np.ndarray:
    out = np.empty(N, np.int32)
    n = lib().orc_gen_breach_list(_p(out, C.c_int32), list_id, N, rate, seed)
    return out[:n].copy()


# -------------------------------------------------------------------- weights
def snappy_compress(data: bytes) -> bytes:
    """Raw Snappy stream by the repository's own greedy matcher (weights_oracle.c): valid Snappy, not snappy-java's bytes."""
    b = np.frombuffer(bytes(data), np.uint8).copy() if len(data) else np.zeros(1, np.uint8)
    out = np.zeros(32 + len(data) + len(data) // 6, np.uint8)
    n = lib().orc_snappy_compress(_p(b, C.c_uint8), len(data), _p(out, C.c_uint8))
    return bytes(out[:n])


def snappy_uncompress(blob: bytes, out_cap: int) -> bytes:
    b = np.frombuffer(bytes(blob), np.uint8).copy() if len(blob) else np.zeros(1, np.uint8)
    out = np.zeros(max(out_cap, 1), np.uint8)
    n = lib().orc_snappy_uncompress(_p(b, C.c_uint8), len(blob), _p(out, C.c_uint8), out_cap)
    if n < 0:
        raise ValueError("malformed Snappy stream")
    return bytes(out[:n])


def get_compressed_weights(weights) -> bytes:
    """LoadPortfolios.getCompressedWeights: Snappy.compress(double[]) (LoadPortfolios.java:563-595)."""
    w = np.ascontiguousarray(weights, np.float64)
    return snappy_compress(w.tobytes())


def get_decompressed_weights(blob: bytes, n: int) -> np.ndarray:
    return np.frombuffer(snappy_uncompress(blob, 8 * n), np.float64).copy()


def doubles_to_bytes(weights) -> bytes:
    w = np.ascontiguousarray(weights, np.float64)
    out = np.zeros(max(8 * w.size, 1), np.uint8)
    lib().orc_doubles_to_bytes(_p(w, C.c_double), w.size, _p(out, C.c_uint8))
    return bytes(out[: 8 * w.size])
