import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import mcp_b200, cases
from mcp_b200 import native as N
from oracle import cref
ctx = mcp_b200.Context(0)
rng = np.random.default_rng(11)
vals, off = cases.breach_like_lists(rng, 3000, 10000, 0.01)
for cid in (N.CODEC_BP32_VB, N.CODEC_BP32_VB | N.FLAG_DELTA, N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA,
            N.CODEC_FASTPFOR128_VB | N.FLAG_DELTA, N.CODEC_IBP32_IVB, N.CODEC_SK_IBP32_IVB, N.CODEC_VB | N.FLAG_DELTA):
    out_off, out = ctx.codec_encode(cid, N.FRAME_APP, vals, off)
    try:
        back = ctx.codec_decode(cid, N.FRAME_APP, out, out_off, off)
        print(hex(cid), "ok", np.array_equal(back, vals), ctx.codec_stats())
    except Exception as e:
        print(hex(cid), "FAIL", e)
        # bisect: decode batches of 32 lists
        for b in range(0, 3000, 32):
            e_ = min(b + 32, 3000)
            oo = out_off[b:e_ + 1] - out_off[b]; vo = off[b:e_ + 1] - off[b]
            try:
                bk = ctx.codec_decode(cid, N.FRAME_APP, out[out_off[b]:out_off[e_]], oo, vo)
            except Exception as e2:
                lens = np.diff(vo); wl = np.diff(oo) // 4
                print("  batch", b, "lens", lens.tolist(), "words", wl.tolist())
                break
