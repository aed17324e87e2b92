import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fxb200
rng = np.random.default_rng(0)
for m in (16, 32, 48, 64, 80, 112, 128, 160):
    for kind in ("random", "graded", "neardiag"):
        if kind == "random":
            A = rng.standard_normal((m, m)); A = A + A.T
        elif kind == "graded":
            Q, _ = np.linalg.qr(rng.standard_normal((m, m)))
            A = (Q * np.geomspace(1e3, 1e-13, m)) @ Q.T; A = 0.5 * (A + A.T)
        else:
            A = np.diag(np.geomspace(1e3, 1e-6, m)) + 1e-6 * rng.standard_normal((m, m)); A = 0.5 * (A + A.T)
        fxb200.syevd_small(A)
        t0 = time.perf_counter()
        w, V, sw = fxb200.syevd_small(A)
        dt = time.perf_counter() - t0
        wo = np.linalg.eigvalsh(A)[::-1]
        print("m=%3d %-8s sweeps=%2d wall=%.0f us  err=%.1e resid=%.1e" % (m, kind, sw, dt * 1e6, np.abs(w - wo).max() / np.abs(wo).max(), np.abs(A @ V - V * w).max() / np.abs(wo).max()))
