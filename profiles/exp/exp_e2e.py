"""e2e through mcp_simulate (pinned host buffers), C2: scenario tail uploaded in one part vs two (option upload_split)."""
import os, sys, time, ctypes
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mcp_b200
from mcp_b200 import native as N
import bench
ctx = mcp_b200.Context(0)
L = N.lib()
P, Nn, F, alpha, seed = 10000, 100000, 32, 0.99, 0x5EED0002
from oracle import cref
E0 = cref.gen_normal(P, F, seed, 1, 100.0, 0.0); S0 = cref.gen_normal(Nn, F, seed, 2, 1.0, 0.05)
hE = bench.pinned_array(L, (P, F), np.float64); hS = bench.pinned_array(L, (Nn, F), np.float64)
hE[:] = E0; hS[:] = S0
t = int(L.mcp_tail_size(Nn, ctypes.c_double(alpha)))
cid = N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA
cap = int(P * L.mcp_codec_max_bytes(cid, N.FRAME_APP, t))
outs = {k: bench.pinned_array(L, (P,), np.float64) for k in ("mean", "variance", "var_loss", "cvar")}
vt = bench.pinned_array(L, (P,), np.int32); eo = bench.pinned_array(L, (P + 1,), np.int64); enc = bench.pinned_array(L, (cap,), np.uint8)
vp = ctypes.c_void_p
def ptr(a): return a.ctypes.data_as(vp)
def once():
    n = ctypes.c_int64(0)
    rc = L.mcp_simulate(ctx._h, ptr(hE), P, ptr(hS), Nn, F, ctypes.c_double(alpha), cid, N.FRAME_APP, ptr(outs["mean"]), ptr(outs["variance"]),
                        ptr(outs["var_loss"]), ptr(vt), ptr(outs["cvar"]), None, ptr(eo), ptr(enc), cap, ctypes.byref(n))
    assert rc == 0, rc
ref = None
for rep in range(3):
    for split in (0, 1):
        ctx.set_option("upload_split", split)
        for _ in range(3): once()
        ts = []
        for _ in range(30):
            t0 = time.perf_counter(); once(); ts.append(time.perf_counter() - t0)
        blob = bytes(enc[: int(eo[P])])
        if ref is None: ref = blob
        assert blob == ref
        print(f"upload_split={split}: median {1e3*np.median(ts):.3f} ms  mean {1e3*np.mean(ts):.3f}  min {1e3*np.min(ts):.3f}", flush=True)
