"""Context object over the C ABI: one per GPU / per thread (include/mcp.h threading rule).

Thin: numpy arrays in, numpy arrays out, every call is one C-ABI call.  No arithmetic of
the hot path happens in Python.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N


def _ptr(a):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def as_i32(a) -> np.ndarray:
    """Java int[] semantics: signed or unsigned 32-bit inputs are accepted."""
    a = np.asarray(a)
    if a.dtype == np.int32:
        return np.ascontiguousarray(a)
    if a.dtype == np.uint32:
        return np.ascontiguousarray(a).view(np.int32)
    return np.ascontiguousarray(a.astype(np.int64) & 0xFFFFFFFF).astype(np.uint32).view(np.int32)


class Context:
    """mcp_ctx wrapper.  Raises McpError (CapacityError for MCP_E_CAP) on any failure."""

    def __init__(self, device: int = 0):
        self._L = N.lib()
        h = C.c_void_p()
        rc = self._L.mcp_create(device, C.byref(h))
        if rc != N.MCP_OK:
            raise N.McpError(rc, "mcp_create", self._L.mcp_strerror(rc).decode() +
                             " -- a B200 (sm_100a) is required; there is no CPU fallback")
        self._h = h
        self.device = device

    # -- plumbing
    def close(self):
        if getattr(self, "_h", None):
            self._L.mcp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int, what: str):
        if rc == N.MCP_OK:
            return
        detail = self._L.mcp_last_error(self._h).decode(errors="replace")
        cls = N.CapacityError if rc == N.MCP_E_CAP else N.McpError
        raise cls(rc, what, f"{self._L.mcp_strerror(rc).decode()}: {detail}")

    def device_info(self) -> dict:
        sm, ma, mi, hb = C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
        self._check(self._L.mcp_device_info(self._h, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(hb)), "device_info")
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "hbm_bytes": hb.value}

    def sync(self):
        self._check(self._L.mcp_sync(self._h), "sync")

    def set_option(self, key: str, value: int):
        self._check(self._L.mcp_set_option(self._h, key.encode(), value), "set_option")

    # -- batched codec (host buffers)
    def codec_encode(self, codec_id: int, framing: int, vals, list_off, cap: int | None = None):
        """-> (out_off int64[nlists+1], out uint8[total])"""
        v = as_i32(vals)
        off = np.ascontiguousarray(list_off, np.int64)
        nl = off.size - 1
        out_off = np.zeros(nl + 1, np.int64)
        total = C.c_int64(0)
        if cap is None:
            # ask for the size first (MCP_E_CAP reports the required size), then encode
            rc = self._L.mcp_codec_encode(self._h, codec_id, framing, _ptr(v), _ptr(off), nl, _ptr(out_off), None, 0,
                                          C.byref(total))
            if rc not in (N.MCP_OK, N.MCP_E_CAP):
                self._check(rc, "codec_encode")
            cap = total.value
        out = np.zeros(max(cap, 1), np.uint8)
        rc = self._L.mcp_codec_encode(self._h, codec_id, framing, _ptr(v), _ptr(off), nl, _ptr(out_off), _ptr(out), cap,
                                      C.byref(total))
        self._check(rc, "codec_encode")
        return out_off, out[: total.value]

    def codec_decode(self, codec_id: int, framing: int, buf, in_off, out_off) -> np.ndarray:
        b = np.ascontiguousarray(buf, np.uint8)
        io = np.ascontiguousarray(in_off, np.int64)
        oo = np.ascontiguousarray(out_off, np.int64)
        out = np.zeros(max(int(oo[-1]), 1), np.int32)
        rc = self._L.mcp_codec_decode(self._h, codec_id, framing, _ptr(b), _ptr(io), io.size - 1, _ptr(out), _ptr(oo))
        self._check(rc, "codec_decode")
        return out[: int(oo[-1])]

    def codec_stats(self) -> dict:
        a, b = C.c_int64(), C.c_int64()
        self._check(self._L.mcp_codec_stats(self._h, C.byref(a), C.byref(b)), "codec_stats")
        c, d = C.c_int64(), C.c_int64()
        self._check(self._L.mcp_codec_stats_short(self._h, C.byref(c), C.byref(d)), "codec_stats_short")
        return {"fast_encodes": a.value, "fast_fallbacks": b.value, "short_calls": c.value, "short_fallbacks": d.value}

    # -- single-list drop-ins
    def compress_ints(self, codec_id: int, vals, out_cap: int | None = None) -> np.ndarray:
        v = as_i32(vals)
        cap = int(N.lib().mcp_codec_max_bytes(codec_id, N.FRAME_WORDS, v.size) // 4) if out_cap is None else out_cap
        out = np.zeros(max(cap, 1), np.int32)
        n = C.c_int32(0)
        self._check(self._L.mcp_codec_compress(self._h, codec_id, _ptr(v), v.size, _ptr(out), cap, C.byref(n)), "compress")
        return out[: n.value]

    def uncompress_ints(self, codec_id: int, words, out_cap: int) -> np.ndarray:
        w = as_i32(words)
        out = np.zeros(max(out_cap, 1), np.int32)
        n = C.c_int32(0)
        self._check(self._L.mcp_codec_uncompress(self._h, codec_id, _ptr(w), w.size, _ptr(out), out_cap, C.byref(n)),
                    "uncompress")
        return out[: n.value]

    def headless_compress_ints(self, codec_id: int, vals, out_cap: int | None = None) -> np.ndarray:
        """SkippableIntegerCODEC.headlessCompress of the composition behind a CODEC_SK_* id: no count word."""
        v = as_i32(vals)
        cap = int(N.lib().mcp_codec_max_bytes(codec_id, N.FRAME_WORDS, v.size) // 4) if out_cap is None else out_cap
        out = np.zeros(max(cap, 1), np.int32)
        n = C.c_int32(0)
        self._check(self._L.mcp_codec_headless_compress(self._h, codec_id, _ptr(v), v.size, _ptr(out), cap, C.byref(n)),
                    "headlessCompress")
        return out[: n.value]

    def headless_uncompress_ints(self, codec_id: int, words, num: int):
        """SkippableIntegerCODEC.headlessUncompress: -> (the `num` ints, words consumed)."""
        w = as_i32(words)
        out = np.zeros(max(num, 1), np.int32)
        used = C.c_int32(0)
        self._check(self._L.mcp_codec_headless_uncompress(self._h, codec_id, _ptr(w), w.size, _ptr(out), num, C.byref(used)),
                    "headlessUncompress")
        return out[:num], used.value

    def get_compressed_instruments(self, codec_id: int, instruments) -> bytes:
        v = as_i32(instruments)
        cap = int(self._L.mcp_codec_max_bytes(codec_id, N.FRAME_APP, v.size))
        out = np.zeros(cap, np.uint8)
        n = C.c_int64(0)
        self._check(self._L.mcp_get_compressed_instruments(self._h, codec_id, _ptr(v), v.size, _ptr(out), cap, C.byref(n)),
                    "getCompressedInstruments")
        return bytes(out[: n.value])

    def get_decompressed_instruments(self, codec_id: int, blob: bytes, number_of_integers: int) -> np.ndarray:
        b = np.frombuffer(blob, np.uint8).copy() if len(blob) else np.zeros(1, np.uint8)
        out = np.zeros(max(number_of_integers, 1), np.int32)
        self._check(self._L.mcp_get_decompressed_instruments(self._h, codec_id, _ptr(b), len(blob), number_of_integers,
                                                             _ptr(out)), "getDecompressedInstruments")
        return out[:number_of_integers]

    # -- weights column (raw Snappy format over the doubles' bytes; include/mcp.h)
    def get_compressed_weights(self, weights) -> bytes:
        w = np.ascontiguousarray(weights, np.float64)
        cap = int(self._L.mcp_weights_max_bytes(w.size))
        out = np.zeros(cap, np.uint8)
        n = C.c_int64(0)
        self._check(self._L.mcp_get_compressed_weights(self._h, _ptr(w), w.size, _ptr(out), cap, C.byref(n)), "getCompressedWeights")
        return bytes(out[: n.value])

    def get_decompressed_weights(self, blob: bytes, number_of_doubles: int) -> np.ndarray:
        b = np.frombuffer(blob, np.uint8).copy() if len(blob) else np.zeros(1, np.uint8)
        out = np.zeros(max(number_of_doubles, 1), np.float64)
        self._check(self._L.mcp_get_decompressed_weights(self._h, _ptr(b), len(blob), number_of_doubles, _ptr(out)),
                    "getDecompressedWeights")
        return out[:number_of_doubles]

    def weights_encode(self, weights, list_off, cap: int | None = None):
        """Batched: list i = weights[list_off[i] .. list_off[i+1]) -> (byte offsets [nlists + 1], bytes)."""
        w = np.ascontiguousarray(weights, np.float64)
        off = np.ascontiguousarray(list_off, np.int64)
        nl = off.size - 1
        cap = int(cap if cap is not None else 8 * w.size + (8 * w.size) // 6 + 40 * max(nl, 1))
        out_off = np.zeros(nl + 1, np.int64)
        out = np.zeros(max(cap, 1), np.uint8)
        n = C.c_int64(0)
        self._check(self._L.mcp_weights_encode(self._h, _ptr(w), _ptr(off), nl, _ptr(out_off), _ptr(out), cap, C.byref(n)), "weights_encode")
        return out_off, out[: n.value]

    def weights_decode(self, buf, in_off, out_off) -> np.ndarray:
        b = np.ascontiguousarray(buf, np.uint8)
        if b.size == 0:
            b = np.zeros(1, np.uint8)
        io, oo = np.ascontiguousarray(in_off, np.int64), np.ascontiguousarray(out_off, np.int64)
        out = np.zeros(max(int(oo[-1]), 1), np.float64)
        self._check(self._L.mcp_weights_decode(self._h, _ptr(b), _ptr(io), io.size - 1, _ptr(out), _ptr(oo)), "weights_decode")
        return out[: int(oo[-1])]

    # -- risk path
    def upload_exposures(self, E):
        E = np.ascontiguousarray(E, np.float64)
        self._check(self._L.mcp_upload_exposures(self._h, _ptr(E), E.shape[0], E.shape[1]), "upload_exposures")

    def build_exposures(self, instruments, weights, pos_off, loadings):
        """Ingest: (instrument, weight) positions in CSR form x the instrument -> factor loading table -> E on the device."""
        inst = np.ascontiguousarray(instruments, np.int32)
        w = np.ascontiguousarray(weights, np.float64)
        off = np.ascontiguousarray(pos_off, np.int64)
        B = np.ascontiguousarray(loadings, np.float64)
        self._check(self._L.mcp_build_exposures(self._h, _ptr(inst), _ptr(w), _ptr(off), off.size - 1, _ptr(B), B.shape[0], B.shape[1]),
                    "build_exposures")

    def upload_scenarios(self, S):
        S = np.ascontiguousarray(S, np.float64)
        self._check(self._L.mcp_upload_scenarios(self._h, _ptr(S), S.shape[0], S.shape[1]), "upload_scenarios")

    def generate_exposures(self, P, F, seed, scale=1.0, shift=0.0, row0=0):
        self._check(self._L.mcp_generate_exposures(self._h, P, F, seed, scale, shift, row0), "generate_exposures")
        self._PF = (P, F)

    def generate_scenarios(self, Nn, F, seed, scale=1.0, shift=0.0, row0=0):
        self._check(self._L.mcp_generate_scenarios(self._h, Nn, F, seed, scale, shift, row0), "generate_scenarios")
        self._NF = (Nn, F)

    def download_exposures(self, P, F) -> np.ndarray:
        E = np.empty((P, F), np.float64)
        self._check(self._L.mcp_download_exposures(self._h, _ptr(E)), "download_exposures")
        return E

    def download_scenarios(self, Nn, F) -> np.ndarray:
        S = np.empty((Nn, F), np.float64)
        self._check(self._L.mcp_download_scenarios(self._h, _ptr(S)), "download_scenarios")
        return S

    def value(self, p0: int, p1: int, n_trials: int) -> np.ndarray:
        V = np.empty((p1 - p0, n_trials), np.float64)
        self._check(self._L.mcp_value(self._h, p0, p1, _ptr(V)), "value")
        return V

    def run(self, alpha: float, codec_id: int = N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA, framing: int = N.FRAME_APP):
        self._check(self._L.mcp_run(self._h, alpha, codec_id, framing), "run")

    def run_sizes(self) -> dict:
        P, Nn, t, eb = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        self._check(self._L.mcp_run_sizes(self._h, C.byref(P), C.byref(Nn), C.byref(t), C.byref(eb)), "run_sizes")
        return {"P": P.value, "N": Nn.value, "tail": t.value, "enc_bytes": eb.value}

    # -- trial-sharded mode (BASELINE config C5)
    @staticmethod
    def comm_unique_id() -> bytes:
        """128-byte NCCL unique id (rank 0 creates it and hands it to the other ranks)."""
        buf = (C.c_uint8 * 128)()
        rc = N.lib().mcp_comm_unique_id(buf)
        if rc != N.MCP_OK:
            raise N.McpError(rc, "comm_unique_id", "NCCL (libnccl.so.2) is not available")
        return bytes(buf)

    def comm_init(self, uid: bytes, rank: int, world: int):
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        self._check(self._L.mcp_comm_init(self._h, buf, rank, world), "comm_init")

    def comm_destroy(self):
        self._check(self._L.mcp_comm_destroy(self._h), "comm_destroy")

    def run_trial_sharded(self, alpha: float, n_global: int, trial_base: int,
                          codec_id: int = N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA, framing: int = N.FRAME_APP):
        """Local candidates -> ONE NCCL all-gather -> merge -> this rank's portfolio slice (fetch() returns it)."""
        self._check(self._L.mcp_run_trial_sharded(self._h, alpha, codec_id, framing, n_global, trial_base), "run_trial_sharded")

    def ts_local(self, alpha: float, n_global: int, trial_base: int, world: int):
        """-> (device pointer, bytes) of this rank's exchange block (bring-your-own-transport variant)."""
        p, nb = C.c_void_p(), C.c_int64()
        self._check(self._L.mcp_ts_local(self._h, alpha, n_global, trial_base, world, C.byref(p), C.byref(nb)), "ts_local")
        return p.value, nb.value

    def ts_merge(self, d_gathered: int, rank: int):
        """Merge this rank's portfolio slice -> (device pointer, bytes) of its flag block."""
        p, nb = C.c_void_p(), C.c_int64()
        self._check(self._L.mcp_ts_merge(self._h, C.c_void_p(d_gathered), rank, C.byref(p), C.byref(nb)), "ts_merge")
        return p.value, nb.value

    def ts_flags(self, d_gathered_flags: int) -> int:
        fl = C.c_int64()
        self._check(self._L.mcp_ts_flags(self._h, C.c_void_p(d_gathered_flags), C.byref(fl)), "ts_flags")
        return fl.value

    def ts_local2(self):
        p, nb = C.c_void_p(), C.c_int64()
        self._check(self._L.mcp_ts_local2(self._h, C.byref(p), C.byref(nb)), "ts_local2")
        return p.value, nb.value

    def ts_merge2(self, d_gathered2: int):
        self._check(self._L.mcp_ts_merge2(self._h, C.c_void_p(d_gathered2)), "ts_merge2")

    def ts_finish(self, rank: int, world: int, codec_id: int = N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA, framing: int = N.FRAME_APP):
        self._check(self._L.mcp_ts_finish(self._h, rank, world, codec_id, framing), "ts_finish")

    def run_stats(self) -> dict:
        fb = C.c_int64()
        self._check(self._L.mcp_run_stats(self._h, C.byref(fb)), "run_stats")
        return {"fallback_rows": fb.value}

    def fetch(self, breach: bool = True, encoded: bool = True) -> dict:
        sz = self.run_sizes()
        P, t, eb = sz["P"], sz["tail"], sz["enc_bytes"]
        r = {
            "mean": np.empty(P), "variance": np.empty(P), "var_loss": np.empty(P), "var_trial": np.empty(P, np.int32),
            "cvar": np.empty(P), "tail": t,
            "breach": np.empty((P, t), np.int32) if breach else None,
            "enc_off": np.empty(P + 1, np.int64) if encoded else None,
            "enc": np.empty(max(eb, 1), np.uint8) if encoded else None,
        }
        self._check(self._L.mcp_fetch(self._h, _ptr(r["mean"]), _ptr(r["variance"]), _ptr(r["var_loss"]), _ptr(r["var_trial"]),
                                      _ptr(r["cvar"]), _ptr(r["breach"]), _ptr(r["enc_off"]), _ptr(r["enc"]), eb), "fetch")
        if encoded:
            r["enc"] = r["enc"][:eb]
        return r

    def simulate(self, E, S, alpha: float, codec_id: int = N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA,
                 framing: int = N.FRAME_APP) -> dict:
        """Upload + run + fetch in one C-ABI call (the end-to-end path)."""
        E = np.ascontiguousarray(E, np.float64)
        S = np.ascontiguousarray(S, np.float64)
        P, F = E.shape
        Nn = S.shape[0]
        t = int(self._L.mcp_tail_size(Nn, alpha))
        cap = int(P * self._L.mcp_codec_max_bytes(codec_id, framing, t))
        r = {
            "mean": np.empty(P), "variance": np.empty(P), "var_loss": np.empty(P), "var_trial": np.empty(P, np.int32),
            "cvar": np.empty(P), "tail": t, "breach": np.empty((P, t), np.int32),
            "enc_off": np.empty(P + 1, np.int64), "enc": np.empty(max(cap, 1), np.uint8),
        }
        n = C.c_int64(0)
        self._check(self._L.mcp_simulate(self._h, _ptr(E), P, _ptr(S), Nn, F, alpha, codec_id, framing, _ptr(r["mean"]),
                                         _ptr(r["variance"]), _ptr(r["var_loss"]), _ptr(r["var_trial"]), _ptr(r["cvar"]),
                                         _ptr(r["breach"]), _ptr(r["enc_off"]), _ptr(r["enc"]), cap, C.byref(n)), "simulate")
        r["enc"] = r["enc"][: n.value]
        return r

    # -- measurement
    def profile(self, on: bool):
        self._check(self._L.mcp_profile_enable(self._h, int(on)), "profile_enable")

    def profile_reset(self):
        self._check(self._L.mcp_profile_reset(self._h), "profile_reset")

    def profile_get(self) -> dict:
        ln = (C.c_int64 * N.K_COUNT)()
        ms = (C.c_double * N.K_COUNT)()
        self._check(self._L.mcp_profile_get(self._h, ln, ms), "profile_get")
        return {N.K_NAMES[i]: {"launches": ln[i], "ms": ms[i]} for i in range(N.K_COUNT)}

    def timer_start(self):
        self._check(self._L.mcp_timer_start(self._h), "timer_start")

    def timer_stop(self) -> float:
        ms = C.c_double(0)
        self._check(self._L.mcp_timer_stop(self._h, C.byref(ms)), "timer_stop")
        return ms.value

    def fp64_peak(self) -> float:
        tf = C.c_double(0)
        self._check(self._L.mcp_fp64_peak(self._h, C.byref(tf)), "fp64_peak")
        return tf.value

    def fp64_mma_peak(self) -> float:
        tf = C.c_double(0)
        self._check(self._L.mcp_fp64_mma_peak(self._h, C.byref(tf)), "fp64_mma_peak")
        return tf.value

    def flush_l2(self):
        self._check(self._L.mcp_flush_l2(self._h), "flush_l2")

    # -- device-resident codec path (benchmarks)
    def generate_breach_lists(self, nlists: int, n_trials: int, rate: float, seed: int, first_list: int = 0):
        dv, do, tot = C.c_void_p(), C.c_void_p(), C.c_int64()
        self._check(self._L.mcp_generate_breach_lists(self._h, nlists, n_trials, rate, seed, first_list, C.byref(dv),
                                                      C.byref(do), C.byref(tot)), "generate_breach_lists")
        return dv.value, do.value, tot.value

    def dev_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._check(self._L.mcp_dev_alloc(self._h, nbytes, C.byref(p)), "dev_alloc")
        return p.value

    def dev_free(self, dptr: int):
        self._check(self._L.mcp_dev_free(self._h, C.c_void_p(dptr)), "dev_free")

    def h2d(self, dptr: int, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        self._check(self._L.mcp_memcpy_h2d(self._h, C.c_void_p(dptr), _ptr(arr), arr.nbytes), "h2d")

    def d2h(self, arr: np.ndarray, dptr: int, nbytes: int | None = None):
        self._check(self._L.mcp_memcpy_d2h(self._h, _ptr(arr), C.c_void_p(dptr), arr.nbytes if nbytes is None else nbytes), "d2h")

    def codec_encode_dev(self, codec_id, framing, d_vals, d_list_off, nlists, total_vals, d_out_off, d_out, cap) -> int:
        tot = C.c_int64(0)
        self._check(self._L.mcp_codec_encode_dev(self._h, codec_id, framing, C.c_void_p(d_vals), C.c_void_p(d_list_off),
                                                 nlists, total_vals, C.c_void_p(d_out_off), C.c_void_p(d_out), cap,
                                                 C.byref(tot)), "codec_encode_dev")
        return tot.value

    def codec_decode_dev(self, codec_id, framing, d_in, d_in_off, nlists, d_out, d_out_off):
        self._check(self._L.mcp_codec_decode_dev(self._h, codec_id, framing, C.c_void_p(d_in), C.c_void_p(d_in_off), nlists,
                                                 C.c_void_p(d_out), C.c_void_p(d_out_off)), "codec_decode_dev")


class SimulatePipeline:
    """Several `mcp_simulate` calls in flight on one GPU: `depth` contexts (contexts are independent; a context is used by
    one thread at a time) and one worker thread per context, so that the uploads and downloads of one step ride under the
    kernels of another.  Results are those of `Context.simulate`, step for step.

        with SimulatePipeline(device=0, depth=2) as pipe:
            futures = [pipe.submit(E, S, 0.99) for E, S in steps]
            results = [f.result() for f in futures]

    (`bench.py`'s e2e leg does the same through the raw C ABI with pinned buffers: 2.85 ms per C2 step against 3.23 ms
    for the calls back to back and 2.81 ms for the kernels alone.)"""

    def __init__(self, device: int = 0, depth: int = 2):
        import queue
        from concurrent.futures import ThreadPoolExecutor
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self._ctxs = [Context(device) for _ in range(depth)]
        self._free = queue.Queue()
        for c in self._ctxs:
            self._free.put(c)
        self._pool = ThreadPoolExecutor(max_workers=depth, thread_name_prefix="mcp-simulate")

    def contexts(self):
        return list(self._ctxs)

    def submit(self, E, S, alpha: float, codec_id: int = N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA, framing: int = N.FRAME_APP):
        """Queue one step; returns a concurrent.futures.Future whose result is the dict `Context.simulate` returns."""
        def job():
            ctx = self._free.get()
            try:
                return ctx.simulate(E, S, alpha, codec_id, framing)
            finally:
                self._free.put(ctx)
        return self._pool.submit(job)

    def close(self):
        self._pool.shutdown(wait=True)
        for c in self._ctxs:
            c.close()
        self._ctxs = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
