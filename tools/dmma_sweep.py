"""DMMA issue-rate sweep: TFLOP/s vs (CTAs/SM, warps/CTA, independent accumulators/warp)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fxb200  # noqa: E402

ctx = fxb200.default_context()
res = []
iters = 4000
for ctas, warps in [(1, 4), (1, 8), (2, 4), (2, 8), (1, 16), (4, 8), (8, 8)]:
    for ilp in (1, 2, 4, 8, 16, 32):
        ms = fxb200.bench_kernel(7, n=ctas, d=ilp, b=warps, m=iters, reps=3)
        tf = ctx.sm_count * ctas * warps * iters * ilp * 512.0 / float(np.min(ms)) / 1e9
        res.append(dict(ctas_per_sm=ctas, warps_per_cta=warps, ilp=ilp, tflops=round(tf, 2)))
        print(res[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/dmma_sweep.json", "w"), indent=1)
