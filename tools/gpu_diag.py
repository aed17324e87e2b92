"""GPU diagnostics: fp64 peak microbenchmarks (the roofline denominators BASELINE.md asks the
builder to measure), per-kernel timings on device-resident inputs and a few fits with their
phase breakdown.  Writes gpurun_out/diag.json.  Usage: python tools/gpu_diag.py [quick|full]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fxb200  # noqa: E402
from fxb200.synth import mix, CONFIGS  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "quick"
out = {}
ctx = fxb200.Context(0)
sm = ctx.sm_count


def best(ms):
    return float(np.min(ms)), float(np.median(ms))


def section(name, fn):
    t0 = time.time()
    try:
        out[name] = fn()
    except Exception as e:      # keep going: one broken kernel must not hide the others
        out[name] = {"error": "%s: %s" % (type(e).__name__, e)}
    print(name, json.dumps(out[name]), "(%.1fs)" % (time.time() - t0), flush=True)


def peaks():
    r = {}
    ms = fxb200.bench_kernel(4, reps=5, ctx=ctx)
    flops = sm * 8 * 8 * 20000 * 8 * 512.0
    r["dmma_tflops_best"] = flops / best(ms)[0] / 1e9
    r["dmma_tflops_median"] = flops / best(ms)[1] / 1e9
    ms = fxb200.bench_kernel(5, reps=5, ctx=ctx)
    flops = sm * 8 * 256 * 20000 * 16 * 2.0
    r["dfma_tflops_best"] = flops / best(ms)[0] / 1e9
    ms = fxb200.bench_kernel(6, n=1 << 29, reps=5, ctx=ctx)
    r["copy_GBs_best"] = 2 * (1 << 29) * 8 / best(ms)[0] / 1e6
    try:
        import torch
        a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
        b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
        torch.matmul(a, b)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        r["cublas_dgemm_8192_tflops_best"] = 2 * 8192.0 ** 3 / min(ts) / 1e9
        del a, b
        torch.cuda.empty_cache()
    except Exception as e:
        r["cublas_error"] = str(e)
    return r


def kernels():
    r = {}
    for (n, d) in ([(16384, 128), (32768, 256)] if mode == "quick" else [(16384, 128), (32768, 256), (65536, 256)]):
        for kid, kn in ((0, "rbf"), (1, "poly3")):
            ms = fxb200.bench_kernel(0, n=n, d=d, kernel_id=kid, reps=3, ctx=ctx)
            r["gram_%s_n%d_d%d" % (kn, n, d)] = {"ms": best(ms)[0], "tflops": 2.0 * n * n * d / best(ms)[0] / 1e9}
    for n in ([32768] if mode == "quick" else [32768, 65536]):
        ms = fxb200.bench_kernel(1, n=n, d=16, reps=5, ctx=ctx)
        r["center_n%d" % n] = {"ms": best(ms)[0], "GBs": 16.0 * n * n / best(ms)[0] / 1e6}
        for b in (8, 16, 24, 32, 64):
            ms = fxb200.bench_kernel(2, n=n, b=b, reps=5, ctx=ctx)
            r["symm_n%d_b%d" % (n, b)] = {"ms": best(ms)[0], "GBs": 8.0 * n * n / best(ms)[0] / 1e6,
                                          "tflops": 2.0 * n * n * b / best(ms)[0] / 1e9}
    for (m, n, d, b) in [(65536, 32768, 64, 32), (65536, 32768, 256, 16)]:
        ms = fxb200.bench_kernel(3, n=n, d=d, b=b, m=m, reps=3, ctx=ctx)
        r["transform_m%d_n%d_d%d_k%d" % (m, n, d, b)] = {
            "ms": best(ms)[0], "tflops": 2.0 * m * n * (d + b) / best(ms)[0] / 1e9}
    return r


def fit(name):
    def run():
        cfg = CONFIGS[name]
        X, _ = mix(cfg["n"], cfg["d"], cfg["c"], cfg["sep"], cfg["seed"])
        kern = fxb200.from_spec(cfg["kernel"], cfg["params"])
        res = {}
        for rep in range(2):
            t0 = time.time()
            V, L, st = fxb200.kpca(X, kern, k=cfg["k"], ctx=ctx, return_stats=True)
            res["wall_s_%d" % rep] = time.time() - t0
        res["stats"] = st
        res["lambda"] = L.tolist()
        res["orth_err"] = float(np.abs(V.T @ V - np.eye(cfg["k"])).max())
        return res
    return run


section("peaks", peaks)
section("kernels", kernels)
section("fit_C1", fit("C1"))
section("fit_C2", fit("C2"))
if mode == "full":
    section("fit_HEADLINE", fit("HEADLINE"))
    section("fit_C3", fit("C3"))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "diag.json"), "w") as f:
    json.dump(out, f, indent=1)
