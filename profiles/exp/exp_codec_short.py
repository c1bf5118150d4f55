"""C4 codec timing (1M lists x ~100 ints, app framing), device-resident: thread-per-list kernels vs batched kernels."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mcp_b200
from mcp_b200 import native as N
ctx = mcp_b200.Context(0)
nl = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ntr = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
dv, do, total = ctx.generate_breach_lists(nl, ntr, 0.01, 0x5EED0004)
capb = 5 * total + 16 * nl
d_oo, d_out, d_back = ctx.dev_alloc(8 * (nl + 1)), ctx.dev_alloc(capb), ctx.dev_alloc(4 * total + 16)
for name, cid in (("bp32+vb|delta", N.CODEC_BP32_VB | N.FLAG_DELTA), ("fastpfor256+vb|delta", N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA),
                  ("ibp32+ivb", N.CODEC_IBP32_IVB)):
    for short in (2, 0, 2):
        ctx.set_option("short_list_codec", short)
        e, d = [], []
        for it in range(8):
            ctx.flush_l2(); ctx.sync(); ctx.timer_start()
            nb = ctx.codec_encode_dev(cid, N.FRAME_APP, dv, do, nl, total, d_oo, d_out, capb)
            te = ctx.timer_stop()
            ctx.flush_l2(); ctx.sync(); ctx.timer_start()
            ctx.codec_decode_dev(cid, N.FRAME_APP, d_out, d_oo, nl, d_back, do)
            td = ctx.timer_stop()
            if it >= 3: e.append(te); d.append(td)
        e, d = np.mean(e), np.mean(d)
        ab = 4.0 * total + nb + 16.0 * nl
        print(f"{name} short={short} ints={total} bytes={nb} enc {e:.3f} ms {total/e/1e6:.1f} Gint/s {ab/e/1e6:.0f} GB/s ({ab/e/1e6/6551.4:.3f}) | dec {d:.3f} ms {total/d/1e6:.1f} Gint/s {ab/d/1e6:.0f} GB/s ({ab/d/1e6/6551.4:.3f})  {ctx.codec_stats()}", flush=True)
