"""C4 codec timing (1M lists x ~100 ints, app framing), device-resident: flat VariableByte kernels vs thread-per-list kernels."""
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
ref = None
for name, cid in (("fastpfor256+vb|delta", N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA), ("vb|delta", N.CODEC_VB | N.FLAG_DELTA),
                  ("bp32+vb|delta", N.CODEC_BP32_VB | N.FLAG_DELTA)):
    for flat in (1, 0):
        ctx.set_option("flat_codec", flat)
        e, d = [], []
        for it in range(8):
            ctx.flush_l2(); ctx.sync(); ctx.timer_start()
            nb = ctx.codec_encode_dev(cid, N.FRAME_APP, dv, do, nl, total, d_oo, d_out, capb)
            te = ctx.timer_stop()
            ctx.flush_l2(); ctx.sync(); ctx.timer_start()
            ctx.codec_decode_dev(cid, N.FRAME_APP, d_out, d_oo, nl, d_back, do)
            td = ctx.timer_stop()
            if it >= 3: e.append(te); d.append(td)
        back = np.empty(total, np.int32); ctx.d2h(back, d_back)
        orig = np.empty(total, np.int32); ctx.d2h(orig, dv)
        enc = np.empty(nb, np.uint8); ctx.d2h(enc, d_out)
        import hashlib
        h = hashlib.sha256(enc.tobytes()).hexdigest()[:12]
        e, d = np.mean(e), np.mean(d)
        ab = 4.0 * total + nb + 16.0 * nl
        print(f"{name} flat={flat} ints={total} bytes={nb} sha={h} roundtrip={np.array_equal(back, orig)} enc {e:.3f} ms {total/e/1e6:.1f} Gint/s ({ab/e/1e6/6551.4:.3f}) | dec {d:.3f} ms {total/d/1e6:.1f} Gint/s ({ab/d/1e6/6551.4:.3f})  {ctx.codec_stats()}", flush=True)
