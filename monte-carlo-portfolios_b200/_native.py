"""ctypes binding of libmcp.so (the C ABI in include/mcp.h).

The shared library is the product; this file only declares its signatures.  There is
no Python or CPU implementation behind any of these calls: if libmcp.so is missing or
no B200 is visible, the calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MCP_LIB_PATH") or os.path.join(_HERE, "libmcp.so")  # override: experiments only
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "mcp.h")

MCP_OK, MCP_E_ARG, MCP_E_CAP, MCP_E_CUDA, MCP_E_NOMEM, MCP_E_FORMAT, MCP_E_STATE, MCP_E_UNSUPPORTED, MCP_E_NCCL = (
    0, -1, -2, -3, -4, -5, -6, -7, -8)

(CODEC_BP32_VB, CODEC_FASTPFOR256_VB, CODEC_FASTPFOR128_VB, CODEC_IBP32_IVB, CODEC_SK_IBP32_IVB, CODEC_VB, CODEC_SK_BP32_VB,
 CODEC_SK_FASTPFOR256_VB, CODEC_SK_FASTPFOR128_VB, CODEC_XOR128_IVB, CODEC_OPTPFD_VB) = range(11)
CODEC_COUNT = 11
FLAG_DELTA = 0x100
FRAME_WORDS, FRAME_APP = 0, 1
K_VALUE, K_SELECT, K_ENCODE, K_DECODE, K_SCAN, K_GEN, K_MOMENTS, K_COMM, K_COUNT = 0, 1, 2, 3, 4, 5, 6, 7, 8
K_NAMES = ["value", "select", "encode", "decode", "scan", "gen", "moments", "comm"]


class McpError(RuntimeError):
    def __init__(self, code: int, what: str, detail: str = ""):
        self.code = code
        super().__init__(f"{what}: {code} ({detail})" if detail else f"{what}: {code}")


class CapacityError(McpError, IndexError):
    """MCP_E_CAP -- the analogue of Java's ArrayIndexOutOfBoundsException when `out` is too small."""


def build(verbose: bool = False) -> str:
    """Compile libmcp.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-j8", "-C", os.path.join(_HERE, "csrc")], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("libmcp.so build failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """Load libmcp.so.  Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the CUDA extension is the product; there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u64, f64, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double, C.c_size_t
    P = C.POINTER
    sig = {
        "mcp_strerror": (C.c_char_p, [C.c_int]),
        "mcp_last_error": (C.c_char_p, [vp]),
        "mcp_create": (C.c_int, [C.c_int, P(vp)]),
        "mcp_destroy": (None, [vp]),
        "mcp_device_info": (C.c_int, [vp, P(C.c_int), P(C.c_int), P(C.c_int), P(sz)]),
        "mcp_sync": (C.c_int, [vp]),
        "mcp_set_option": (C.c_int, [vp, C.c_char_p, i64]),
        "mcp_codec_max_bytes": (i64, [C.c_int, C.c_int, i64]),
        "mcp_codec_compress": (C.c_int, [vp, C.c_int, vp, i32, vp, i32, P(i32)]),
        "mcp_codec_uncompress": (C.c_int, [vp, C.c_int, vp, i32, vp, i32, P(i32)]),
        "mcp_codec_headless_compress": (C.c_int, [vp, C.c_int, vp, i32, vp, i32, P(i32)]),
        "mcp_codec_headless_uncompress": (C.c_int, [vp, C.c_int, vp, i32, vp, i32, P(i32)]),
        "mcp_get_compressed_instruments": (C.c_int, [vp, C.c_int, vp, i32, vp, i64, P(i64)]),
        "mcp_get_decompressed_instruments": (C.c_int, [vp, C.c_int, vp, i64, i32, vp]),
        "mcp_weights_max_bytes": (i64, [i64]),
        "mcp_weights_encode": (C.c_int, [vp, vp, vp, i64, vp, vp, i64, P(i64)]),
        "mcp_weights_decode": (C.c_int, [vp, vp, vp, i64, vp, vp]),
        "mcp_get_compressed_weights": (C.c_int, [vp, vp, i32, vp, i64, P(i64)]),
        "mcp_get_decompressed_weights": (C.c_int, [vp, vp, i64, i32, vp]),
        "mcp_codec_encode": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, i64, vp, vp, i64, P(i64)]),
        "mcp_codec_decode": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, i64, vp, vp]),
        "mcp_codec_stats": (C.c_int, [vp, P(i64), P(i64)]),
        "mcp_codec_stats_short": (C.c_int, [vp, P(i64), P(i64)]),
        "mcp_codec_encode_dev": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, i64, i64, vp, vp, i64, P(i64)]),
        "mcp_codec_decode_dev": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, i64, vp, vp]),
        "mcp_upload_exposures": (C.c_int, [vp, vp, i64, i32]),
        "mcp_upload_scenarios": (C.c_int, [vp, vp, i64, i32]),
        "mcp_build_exposures": (C.c_int, [vp, vp, vp, vp, i64, vp, i64, i32]),
        "mcp_generate_exposures": (C.c_int, [vp, i64, i32, u64, f64, f64, i64]),
        "mcp_generate_scenarios": (C.c_int, [vp, i64, i32, u64, f64, f64, i64]),
        "mcp_download_exposures": (C.c_int, [vp, vp]),
        "mcp_download_scenarios": (C.c_int, [vp, vp]),
        "mcp_value": (C.c_int, [vp, i64, i64, vp]),
        "mcp_tail_size": (i64, [i64, f64]),
        "mcp_run": (C.c_int, [vp, f64, C.c_int, C.c_int]),
        "mcp_run_sizes": (C.c_int, [vp, P(i64), P(i64), P(i64), P(i64)]),
        "mcp_run_stats": (C.c_int, [vp, P(i64)]),
        "mcp_comm_unique_id": (C.c_int, [vp]),
        "mcp_comm_init": (C.c_int, [vp, vp, C.c_int, C.c_int]),
        "mcp_comm_destroy": (C.c_int, [vp]),
        "mcp_run_trial_sharded": (C.c_int, [vp, f64, C.c_int, C.c_int, i64, i64]),
        "mcp_ts_local": (C.c_int, [vp, f64, i64, i64, C.c_int, P(vp), P(i64)]),
        "mcp_ts_merge": (C.c_int, [vp, vp, C.c_int, P(vp), P(i64)]),
        "mcp_ts_flags": (C.c_int, [vp, vp, P(i64)]),
        "mcp_ts_local2": (C.c_int, [vp, P(vp), P(i64)]),
        "mcp_ts_merge2": (C.c_int, [vp, vp]),
        "mcp_ts_finish": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int]),
        "mcp_ts_candidates": (i64, [i64, f64, C.c_int]),
        "mcp_fetch": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, vp, i64]),
        "mcp_simulate": (C.c_int, [vp, vp, i64, vp, i64, i32, f64, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, i64, P(i64)]),
        "mcp_profile_enable": (C.c_int, [vp, C.c_int]),
        "mcp_profile_reset": (C.c_int, [vp]),
        "mcp_profile_get": (C.c_int, [vp, P(i64), P(f64)]),
        "mcp_timer_start": (C.c_int, [vp]),
        "mcp_timer_stop": (C.c_int, [vp, P(f64)]),
        "mcp_fp64_peak": (C.c_int, [vp, P(f64)]),
        "mcp_fp64_mma_peak": (C.c_int, [vp, P(f64)]),
        "mcp_flush_l2": (C.c_int, [vp]),
        "mcp_generate_breach_lists": (C.c_int, [vp, i64, i32, f64, u64, i64, P(vp), P(vp), P(i64)]),
        "mcp_dev_alloc": (C.c_int, [vp, sz, P(vp)]),
        "mcp_dev_free": (C.c_int, [vp, vp]),
        "mcp_memcpy_h2d": (C.c_int, [vp, vp, vp, sz]),
        "mcp_memcpy_d2h": (C.c_int, [vp, vp, vp, sz]),
        "mcp_host_alloc_pinned": (C.c_int, [sz, P(vp)]),
        "mcp_host_free_pinned": (C.c_int, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError if the .so does not export what mcp.h declares
        fn.restype = res
        fn.argtypes = args
    L._mcp_signatures = sig
    _lib = L
    return L


def declared_symbols() -> list[str]:
    """Every function name include/mcp.h declares (used by the no-GPU export test)."""
    import re
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mcp_[a-z0-9_]+)\s*\(", text)))
