"""Bit-exactness of the flat VariableByte kernels against the oracle on awkward shapes (debugging aid)."""
import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, cases, mcp_b200
from mcp_b200 import native as N
from oracle import cref
ctx = mcp_b200.Context(0)
ctx.set_option("short_list_codec", 2)
ctx.set_option("flat_codec", 1)
rng = np.random.default_rng(11)
def mk(kind):
    if kind == "breach":
        return cases.breach_like_lists(rng, 3000, 10000, 0.01)
    if kind == "tiny":  # many empty / one-value lists
        lens = rng.integers(0, 4, 5000)
    elif kind == "mixed":
        lens = rng.integers(0, 250, 800)
    elif kind == "big":  # values that need 3-5 bytes
        lens = rng.integers(0, 120, 1500)
    off = np.zeros(lens.size + 1, np.int64); off[1:] = np.cumsum(lens)
    if kind == "big":
        vals = rng.integers(-2**31, 2**31 - 1, off[-1]).astype(np.int32)
    else:
        vals = np.concatenate([np.sort(rng.integers(0, 50000, l)) for l in lens] + [np.zeros(0, np.int64)]).astype(np.int32)
    return vals, off
bad = 0
for kind in ("breach", "tiny", "mixed", "big"):
    vals, off = mk(kind)
    for cid in (N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA, N.CODEC_FASTPFOR128_VB | N.FLAG_DELTA, N.CODEC_VB | N.FLAG_DELTA, N.CODEC_VB, N.CODEC_SK_FASTPFOR256_VB|N.FLAG_DELTA, N.CODEC_FASTPFOR256_VB,
                N.CODEC_BP32_VB | N.FLAG_DELTA, N.CODEC_BP32_VB, N.CODEC_SK_BP32_VB | N.FLAG_DELTA, N.CODEC_SK_BP32_VB):
        for fr in (N.FRAME_APP, N.FRAME_WORDS):
            try:
                out_off, out = ctx.codec_encode(cid, fr, vals, off)
                ok_bytes = all(bytes(out[out_off[i]: out_off[i + 1]]) == (cref.get_compressed_instruments(cid, vals[off[i]:off[i+1]]) if fr == N.FRAME_APP else cref.compress(cid, vals[off[i]:off[i+1]]).tobytes()) for i in range(off.size - 1))
                if not ok_bytes:
                    for i in range(off.size - 1):
                        got = bytes(out[out_off[i]: out_off[i + 1]])
                        want = (cref.get_compressed_instruments(cid, vals[off[i]:off[i+1]]) if fr == N.FRAME_APP else cref.compress(cid, vals[off[i]:off[i+1]]).tobytes())
                        if got != want:
                            print('  first bad list', i, 'n', off[i+1]-off[i], 'got', len(got), 'want', len(want), got[:24].hex(), want[:24].hex(), 'vals', vals[off[i]:off[i]+4]); break
                back = ctx.codec_decode(cid, fr, out, out_off, off)
                print(kind, hex(cid), fr, 'bytes', ok_bytes, 'roundtrip', np.array_equal(back, vals), ctx.codec_stats_short() if hasattr(ctx, 'codec_stats_short') else '')
                bad += (not ok_bytes) or (not np.array_equal(back, vals))
            except Exception as e:
                print(kind, hex(cid), fr, 'FAIL', e); bad += 1
print("BAD", bad)
