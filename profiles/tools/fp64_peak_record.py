"""Record the FP64 peak probes (the denominator of the valuation kernel's roofline) with the clocks they ran at.
   python profiles/tools/fp64_peak_record.py > gpurun_out/fp64_peak.json   (on the GPU box; copied to profiles/r2/)"""
import json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mcp_b200

def smi():
    q = "clocks.sm,clocks.max.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active"
    try:
        out = subprocess.run(["nvidia-smi", "-i", "0", f"--query-gpu={q}", "--format=csv,noheader"], capture_output=True, text=True, timeout=10).stdout.strip()
        return dict(zip(q.split(","), [x.strip() for x in out.split(",")]))
    except Exception as e:
        return {"error": repr(e)}

ctx = mcp_b200.Context(0)
info = ctx.device_info()
for _ in range(3):
    ctx.fp64_mma_peak(); ctx.fp64_peak()
runs = []
for i in range(5):
    before = smi()
    dmma = ctx.fp64_mma_peak()
    dfma = ctx.fp64_peak()
    runs.append({"dmma_tflops": dmma, "dfma_tflops": dfma, "nvidia_smi_before": before, "nvidia_smi_after": smi()})
sm, ghz = info["sm_count"], 1.965
print(json.dumps({
    "what": "dependent-free mma.sync.m8n8k4.f64 loop (k_fp64_mma_peak, 1024 threads/CTA) and dependent-free DFMA loop (k_fp64_peak), "
            "CUDA-event timed inside libmcp.so; bench.py runs the same two probes just before its timed region and holds the "
            "valuation kernel to the higher one",
    "device": info,
    "nominal": {"formula": "SMs x 128 FP64 FLOP/clk x SM clock", "tflops_at_1965_mhz": sm * 128 * ghz * 1e9 / 1e12},
    "runs": runs,
    "max_dmma_tflops": max(r["dmma_tflops"] for r in runs), "max_dfma_tflops": max(r["dfma_tflops"] for r in runs),
}, indent=1))
