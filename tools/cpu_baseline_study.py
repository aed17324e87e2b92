"""CPU baseline study (SURVEY.md 8d / BASELINE.md 3): times the oracle's dense reference path
(Gram + map, centring, LAPACK dsyevd 'V','U') at several n on this box's host cores, fits
t = a n^2 + c n^3, extrapolates to the headline n, and times a CPU iterative top-k solve
(scipy eigsh on the CPU-built centred K) as the fairer secondary baseline.
Writes gpurun_out/cpu_baseline_study.json."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
import fxb200  # noqa: E402
from fxb200.synth import mix, CONFIGS  # noqa: E402
from threadpoolctl import threadpool_info  # noqa: E402

cfg = CONFIGS["HEADLINE"]
sizes = [int(a) for a in sys.argv[1:]] or [1000, 2048, 4096, 8192]
kern = O.GaussianKernel(*cfg["params"])
rows = []
for n in sizes:
    X, _ = mix(n, cfg["d"], cfg["c"], cfg["sep"], cfg["seed"])
    t0 = time.perf_counter(); K = O.kernelmatrix(kern, X)
    t1 = time.perf_counter(); Kc, _, _ = O.centerkernelmatrix(K)
    t2 = time.perf_counter(); V, L = O.components_eig(Kc)
    t3 = time.perf_counter()
    rows.append(dict(n=n, gram_s=t1 - t0, center_s=t2 - t1, dsyevd_s=t3 - t2, total_s=t3 - t0))
    print(rows[-1], flush=True)
    if n == max(sizes):
        from scipy.sparse.linalg import eigsh
        t4 = time.perf_counter()
        w, v = eigsh(Kc, k=cfg["k"], which="LA", tol=1e-12)
        t5 = time.perf_counter()
        iterative = dict(n=n, eigsh_s=t5 - t4, rel_err_vs_dsyevd=float(np.abs(np.sort(w)[::-1] / (L[:cfg["k"]]) - 1).max()))
        print(iterative, flush=True)
ns = np.array([r["n"] for r in rows], float)
c3 = float(np.mean([r["dsyevd_s"] / r["n"] ** 3 for r in rows[-2:]]))
a2 = float(np.mean([(r["gram_s"] + r["center_s"]) / r["n"] ** 2 for r in rows[-2:]]))
N = cfg["n"]
out = dict(config="HEADLINE n=%d d=%d k=%d" % (N, cfg["d"], cfg["k"]), cores=os.cpu_count(),
           blas=[dict(api=p.get("internal_api"), threads=p.get("num_threads"), version=p.get("version"))
                 for p in threadpool_info()],
           julia="N/A: Julia is not installed in this image; the reference's dense path is timed through "
                 "the NumPy/SciPy-LAPACK oracle (same dsyevd routine)",
           measurements=rows, iterative=iterative,
           fit=dict(a_n2=a2, c_n3=c3),
           extrapolated_headline_s=a2 * N ** 2 + c3 * N ** 3,
           extrapolated_headline_eigsh_s=(a2 * N ** 2 + iterative["eigsh_s"] * (N / iterative["n"]) ** 2),
           memory_note="dense dsyevd at n=65536 needs ~4 n^2 doubles = 137 GB")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "cpu_baseline_study.json"), "w"), indent=1)
print(json.dumps(out["fit"]), out["extrapolated_headline_s"], out["extrapolated_headline_eigsh_s"])
