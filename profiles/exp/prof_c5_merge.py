"""What a C5 rank does AFTER the exchange (merge of its portfolio slice, exactness check, finalize, encode), on one GPU with
the eight ranks' blocks produced one after the other (12 500 portfolios instead of 100 000: the slice is 1 562 rows)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mcp_b200
ctx = mcp_b200.Context(0)
P, Nl, F, alpha, seed, world = 12_500, 1_250_000, 32, 0.999, 0x5EED0005, 8
ctx.generate_exposures(P, F, seed, 100.0, 0.0)
blocks = []
for g in range(world):
    ctx.generate_scenarios(Nl, F, seed, 1.0, 0.05, row0=g * Nl)
    ptr, nb = ctx.ts_local(alpha, Nl * world, g * Nl, world)
    b = np.empty(nb, np.uint8)
    ctx.d2h(b, ptr)
    blocks.append(b)
gathered = np.concatenate(blocks)
d = ctx.dev_alloc(gathered.size)
ctx.h2d(d, gathered)
# rank 0 again (ts_merge / ts_finish belong to the context's last ts_local: redo shard 0)
ctx.generate_scenarios(Nl, F, seed, 1.0, 0.05, row0=0)
ctx.ts_local(alpha, Nl * world, 0, world)
ctx.sync()
ctx.profile(True); ctx.profile_reset()
ctx.timer_start()
fp, fb = ctx.ts_merge(d, 0)
t_merge = ctx.timer_stop()
flags = np.zeros(fb * world, np.uint8)
df = ctx.dev_alloc(flags.size)
ctx.h2d(df, flags)
print("flagged", ctx.ts_flags(df))
ctx.timer_start()
ctx.ts_finish(0, world)
t_fin = ctx.timer_stop()
print(f"slice of {P // world} rows: merge + check {t_merge:.3f} ms, finish {t_fin:.3f} ms  -> x8 rows: {8 * t_merge:.1f} / {8 * t_fin:.1f} ms", ctx.profile_get())
