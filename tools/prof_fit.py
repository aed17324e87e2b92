"""One KPCA fit on synthetic data -- the short command ncu wraps to capture kernels that only
run inside a fit (e.g. the fused centring + first K*V pass: the first symm_dmma_kernel launch).
Usage: python tools/prof_fit.py <n> <d> <k>"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fxb200  # noqa: E402
from fxb200 import synth  # noqa: E402

n, d, k = [int(x) for x in sys.argv[1:4]]
X, _ = synth.mix(n, d, 17, 3.0, 3002)
V, L, st = fxb200.kpca(X, fxb200.kernels.GaussianKernel(1.0 / d), k=k, return_stats=True)
print("fit ms", st["total_ms"], "center_ms", st["center_ms"], "symm_ms", st["symm_ms"], "steps", st["steps"])
