import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, cases, mcp_b200
from mcp_b200 import native as N
from oracle import cref
ctx = mcp_b200.Context(0)
ctx.set_option("short_list_codec", 2)
ctx.set_option("flat_codec", 1)
rng = np.random.default_rng(11)
vals, off = cases.breach_like_lists(rng, 3000, 10000, 0.01)
bad = 0
for cid in (N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA, N.CODEC_FASTPFOR128_VB | N.FLAG_DELTA, N.CODEC_VB | N.FLAG_DELTA, N.CODEC_VB, N.CODEC_SK_FASTPFOR256_VB|N.FLAG_DELTA, N.CODEC_FASTPFOR256_VB):
    for fr in (N.FRAME_APP, N.FRAME_WORDS):
        try:
            out_off, out = ctx.codec_encode(cid, fr, vals, off)
            ok_bytes = all(bytes(out[out_off[i]: out_off[i + 1]]) == (cref.get_compressed_instruments(cid, vals[off[i]:off[i+1]]) if fr == N.FRAME_APP else cref.compress(cid, vals[off[i]:off[i+1]]).tobytes()) for i in range(off.size - 1))
            back = ctx.codec_decode(cid, fr, out, out_off, off)
            print(hex(cid), fr, 'bytes', ok_bytes, 'roundtrip', np.array_equal(back, vals), ctx.codec_stats())
            bad += (not ok_bytes) or (not np.array_equal(back, vals))
        except Exception as e:
            print(hex(cid), fr, 'FAIL', e); bad += 1
print("BAD", bad)
