"""mcp_simulate, two calls in flight: kernels of the two calls free to interleave (turns 0) vs taking turns on the SMs (1).
C2 (10 000 x 100 000 x 32) and the C3 shards of 8 / 2 / 1 GPUs (125 000 / 500 000 / 1 000 000 x 10 000 x 64) on one GPU."""
import os, sys, argparse
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
shapes = [("c2", 10_000), ("c3", 125_000), ("c3", 500_000)]
if len(sys.argv) > 1:
    shapes = [(a.split(":")[0], int(a.split(":")[1])) for a in sys.argv[1:]]
for turns in (0, 1):
    args = argparse.Namespace(steps=12, warmup=3, no_numa_bind=True, simulate_turns=turns)
    job = bench.Job(args, 0, 0, 1)
    for name, P in shapes:
        r = bench.e2e_leg(job, name, P, 0)
        print(f"turns {turns} {name} P {P}: two in flight {r['ms_per_step']:.3f} ms, one at a time {r['one_call_at_a_time']['ms_per_step']:.3f} ms", flush=True)
    job.close()
