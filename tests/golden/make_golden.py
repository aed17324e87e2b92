"""Mints the golden vectors under tests/golden/ from the CPU oracle (the reference has
none: all its test inputs are unseeded rand()).  Run from the repo root:
    python tests/golden/make_golden.py
Inputs are regenerated from (generator, seed), outputs are stored."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O                     # noqa: E402
import fxb200                          # noqa: E402
from fxb200.synth import mix           # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (n, d, c, sep, seed, kernel name, params, k, m, zseed)
    "kpca_gaussian": (256, 8, 5, 3.0, 7001, "gaussian", (1.0 / 8,), 4, 32, 7002),
    "kpca_polynomial": (256, 8, 5, 3.0, 7003, "polynomial", (1.0 / 8, 1.0, 3.0), 4, 32, 7004),
    "kpca_sigmoid": (200, 6, 4, 3.0, 7005, "sigmoid", (1.0 / 32, 0.25), 3, 17, 7006),
    "kpca_ragged": (1237, 13, 4, 3.0, 7007, "gaussian", (1.0 / 13,), 3, 45, 7008),
}


def okernel(name, params):
    return {"gaussian": O.GaussianKernel, "polynomial": O.PolynomialKernel,
            "sigmoid": O.SigmoidKernel}[name](*params)


def main():
    for name, (n, d, c, sep, seed, kname, params, k, m, zseed) in CASES.items():
        X, _ = mix(n, d, c, sep, seed)
        Z, _ = mix(m, d, c, sep, zseed)
        kern = okernel(kname, params)
        K = O.kernelmatrix(kern, X)
        Kc, r, mu = O.centerkernelmatrix(K)
        V, L, st = O.kpca_topk(X, kern, k)
        T = O.transform_kpca(X, kern, V, Z)
        idx = np.arange(0, n, max(1, n // 16))
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            meta=np.array([n, d, c, sep, seed, k, m, zseed], dtype=np.float64),
            kernel=np.array(kname), params=np.array(params, dtype=np.float64),
            K_rows=K[idx], Kc_rows=Kc[idx], rows=idx, colmean=r, grandmean=np.array(mu),
            L=L, V_abs=np.abs(V), T_abs=np.abs(T), x_checksum=np.array([X.sum(), (X * X).sum()]))
        print(name, "L =", L)


if __name__ == "__main__":
    main()
