"""CPU: the oracle reproduces the committed golden vectors (guards against silent drift
of the oracle or of the synthetic generator)."""
import glob
import os

import numpy as np
import pytest

import oracle as O
from fxb200.synth import mix

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "kpca_*.npz")))


def load_case(path):
    g = np.load(path)
    n, d, c, sep, seed, k, m, zseed = g["meta"]
    X, _ = mix(int(n), int(d), int(c), float(sep), int(seed))
    Z, _ = mix(int(m), int(d), int(c), float(sep), int(zseed))
    kname = str(g["kernel"])
    kern = {"gaussian": O.GaussianKernel, "polynomial": O.PolynomialKernel,
            "sigmoid": O.SigmoidKernel}[kname](*g["params"])
    return g, X, Z, kern, int(k)


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_oracle_matches_golden(path):
    g, X, Z, kern, k = load_case(path)
    np.testing.assert_allclose([X.sum(), (X * X).sum()], g["x_checksum"], rtol=1e-13)
    K = O.kernelmatrix(kern, X)
    np.testing.assert_allclose(K[g["rows"]], g["K_rows"], rtol=1e-13, atol=1e-15)
    V, L, _ = O.kpca_topk(X, kern, k)
    np.testing.assert_allclose(L, g["L"], rtol=1e-11)
    np.testing.assert_allclose(np.abs(V), g["V_abs"], atol=1e-9)
    np.testing.assert_allclose(np.abs(O.transform_kpca(X, kern, V, Z)), g["T_abs"],
                               rtol=1e-8, atol=1e-10 * np.abs(g["T_abs"]).max())
