import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, mcp_b200
from mcp_b200 import native as N
from oracle import cref
ctx = mcp_b200.Context(0)
ctx.set_option("short_list_codec", 2)
ctx.set_option("flat_codec", 1)
rng = np.random.default_rng(1)
lens = np.array([5, 0, 3, 20, 7])
off = np.zeros(lens.size + 1, np.int64); off[1:] = np.cumsum(lens)
vals = np.concatenate([np.sort(rng.integers(0, 5000, l)) for l in lens]).astype(np.int32)
cid = N.CODEC_VB | N.FLAG_DELTA
out_off, out = ctx.codec_encode(cid, N.FRAME_WORDS, vals, off)
print(out_off, out.size, ctx.codec_stats())
for i in range(lens.size):
    print(i, bytes(out[out_off[i]:out_off[i+1]]).hex(), cref.compress(cid, vals[off[i]:off[i+1]]).tobytes().hex())
