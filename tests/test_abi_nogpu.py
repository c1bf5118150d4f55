"""No-GPU checks of the C-ABI library and the host-side mirror of the reference interfaces."""
import ctypes
import os

import numpy as np
import pytest

import mcp_b200
from mcp_b200 import native as N


def test_library_loads_and_exports_every_declared_symbol():
    L = N.lib()
    declared = N.declared_symbols()
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(L, name), f"include/mcp.h declares {name} but libmcp.so does not export it"
    assert set(declared) == set(L._mcp_signatures), "binding table out of sync with include/mcp.h"


def test_library_is_sm100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", N.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not any(a in out for a in ("sm_90", "sm_80", "sm_89"))


def test_strerror_and_pure_helpers():
    L = N.lib()
    assert L.mcp_strerror(0) == b"ok"
    assert b"small" in L.mcp_strerror(N.MCP_E_CAP)
    for n in (0, 1, 17, 1000, 70000):
        for codec in range(N.CODEC_COUNT):
            assert L.mcp_codec_max_bytes(codec, N.FRAME_APP, n) == L.mcp_codec_max_bytes(codec, N.FRAME_WORDS, n) + 4
            assert L.mcp_codec_max_bytes(codec, N.FRAME_WORDS, n) >= 4 * n  # never smaller than raw ints
    assert L.mcp_codec_max_bytes(99, 0, 10) == -1
    # tail size: DESIGN.md 3.4
    assert L.mcp_tail_size(100000, 0.99) == 1001
    assert L.mcp_tail_size(10000, 0.99) == 101
    assert L.mcp_tail_size(10000000, 0.999) == 10001
    assert L.mcp_tail_size(10, 0.9) == 2
    assert L.mcp_tail_size(10, 1.0) == 1
    assert L.mcp_tail_size(0, 0.5) == 0
    from oracle import cref
    for n in (1, 7, 10, 999, 100000):
        for a in (0.5, 0.9, 0.95, 0.99, 0.999, 1.0):
            assert L.mcp_tail_size(n, a) == cref.tail_size(n, a)


def test_null_context_is_an_error_not_a_crash():
    L = N.lib()
    assert L.mcp_sync(None) == N.MCP_E_ARG
    assert L.mcp_run(None, 0.99, 0, 0) == N.MCP_E_ARG
    assert L.mcp_last_error(None) == b"null context"
    L.mcp_destroy(None)


def test_no_cpu_fallback_without_gpu():
    """Without a GPU mcp_create must fail with MCP_E_CUDA -- the product never computes on the CPU."""
    L = N.lib()
    h = ctypes.c_void_p()
    rc = L.mcp_create(0, ctypes.byref(h))
    if rc == N.MCP_OK:  # running on the GPU box
        L.mcp_destroy(h)
        pytest.skip("a GPU is present")
    assert rc == N.MCP_E_CUDA and not h.value
    with pytest.raises(N.McpError):
        mcp_b200.Context(0)


def test_product_does_not_import_the_oracle():
    pkg = os.path.dirname(N.__file__)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f


def test_intwrapper_semantics():
    w = mcp_b200.IntWrapper(3)
    w.add(4)
    w.increment()
    assert w.get() == 8 and w.intValue() == 8 and int(w) == 8
    w.set(1)
    assert w.get() == 1


def test_composition_mapping():
    c = mcp_b200.Composition(mcp_b200.BinaryPacking(), mcp_b200.VariableByte())
    assert c.codec_id == N.CODEC_BP32_VB and c.toString() == "BinaryPacking + VariableByte"
    assert mcp_b200.Composition(mcp_b200.FastPFOR(), mcp_b200.VariableByte(), delta=True).codec_id == N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA
    assert mcp_b200.Composition(mcp_b200.FastPFOR128(), mcp_b200.VariableByte()).codec_id == N.CODEC_FASTPFOR128_VB
    assert mcp_b200.IntegratedComposition(mcp_b200.IntegratedBinaryPacking(), mcp_b200.IntegratedVariableByte()).codec_id == N.CODEC_IBP32_IVB
    with pytest.raises(NotImplementedError):  # no silent fallback for compositions without a kernel
        mcp_b200.Composition(mcp_b200.VariableByte(), mcp_b200.BinaryPacking())


def test_instrument_document_fallback_rule():
    """LoadPortfolios.java:254-260: the compressed blob is stored only if smaller than 4n bytes."""
    lp = mcp_b200.LoadPortfolios.__new__(mcp_b200.LoadPortfolios)  # no context needed: blob injected
    doc = lp.getInstrumentDocument([1, 2, 3, 4], compressed=b"\0" * 12)
    assert doc == {"uncompressedIntegerSize": 4, "compressed": b"\0" * 12, "compressedByteSize": 12, "algo": "bpacking+variablebyte"}
    doc = lp.getInstrumentDocument([1, -2], compressed=b"\0" * 8)
    assert doc["algo"] == "none" and doc["compressed"] == b"\x00\x00\x00\x01\xff\xff\xff\xfe" and doc["compressedByteSize"] == 8
    doc = lp.getInstrumentDocument([], compressed=b"\0\0\0\0")
    assert doc["algo"] == "none" and doc["compressed"] == b""


def test_as_i32_java_int_semantics():
    from mcp_b200.engine import as_i32
    assert list(as_i32([-1, 2**31 - 1, 2**32 - 1, 2**31])) == [-1, 2**31 - 1, -1, -2**31]
    assert as_i32(np.array([1, 2], np.uint32)).dtype == np.int32


def test_max_bytes_bounds_the_worst_case_of_every_codec():
    """mcp_codec_max_bytes is documented as an upper bound: VariableByte alone needs 5 bytes per int for values
    >= 2^28 (negative ints, wrapped deltas of unsorted input) -- the oracle's real sizes must fit for every codec."""
    from oracle import cref
    L = N.lib()
    rng = np.random.default_rng(7)
    inputs = [np.full(10000, -1), np.full(2351, 1 << 28), rng.integers(-2**31, 2**31, size=70001),
              np.where(rng.random(66000) < 0.5, rng.integers(0, 1 << 31, size=66000), 0), np.arange(5000)[::-1] * 100003]
    for codec in range(N.CODEC_COUNT):
        for flag in (0, N.FLAG_DELTA):
            for v in inputs:
                words = cref.compress(codec | flag, v).size
                assert 4 * words <= L.mcp_codec_max_bytes(codec | flag, N.FRAME_WORDS, len(v)), (codec, flag, len(v), words)


def test_missing_nccl_is_an_error_code_not_a_crash():
    """libmcp.so binds NCCL with dlopen at first use; without the library mcp_comm_unique_id must return MCP_E_NCCL
    (not segfault: dlerror() may be read only once).  Run in a child process: the binding is cached per process."""
    import subprocess
    import sys
    code = ("import ctypes, sys; sys.path.insert(0, %r); from mcp_b200 import native as N; L = N.lib(); "
            "buf = (ctypes.c_uint8 * 128)(); rc = L.mcp_comm_unique_id(buf); print('rc', rc); "
            "sys.exit(0 if rc == N.MCP_E_NCCL else 3)") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MCP_NCCL_LIB="/nonexistent/libnccl.so.2")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
