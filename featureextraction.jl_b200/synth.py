"""Deterministic synthetic inputs for the KPCA / KLDA hot path.

``mix(n, d, c, sep, seed)`` is a Gaussian mixture with ``c`` centres whose
kernel matrices have a clear spectral gap after the (c-1)-th eigenvalue, so
the top-k parity criterion (subspace angle) is well posed for k = c-1.
The same generator feeds the oracle, the CUDA path and bench.py.
"""
import numpy as np


def mix(n, d, c, sep=3.0, seed=0):
    """Returns (X, y): X is n x d float64 **column-major** (Julia Matrix
    layout, rows = observations), y the 0-based mixture component per row."""
    rng = np.random.default_rng(seed)
    centres = sep * rng.standard_normal((c, d))
    y = rng.integers(0, c, n)
    X = centres[y] + rng.standard_normal((n, d))
    return np.asfortranarray(X), y


# Named configurations of BASELINE.json.  gamma = 1/d.
CONFIGS = {
    # name: (n, d, k, c, sep, seed, kernel, kernel params)
    "C1": dict(n=1000, d=10, k=2, c=3, sep=3.0, seed=1001, kernel="gaussian", params=(0.1,)),
    "C2": dict(n=16384, d=128, k=8, c=9, sep=3.0, seed=2001, kernel="gaussian", params=(1.0 / 128,)),
    "C3": dict(n=65536, d=256, k=16, c=17, sep=3.0, seed=3001, kernel="polynomial",
               params=(1.0 / 256, 1.0, 3.0)),
    "HEADLINE": dict(n=65536, d=256, k=16, c=17, sep=3.0, seed=3002, kernel="gaussian",
                     params=(1.0 / 256,)),
    "C4": dict(n=32768, d=512, k=9, c=10, sep=2.0, seed=4001, kernel="gaussian", params=(1.0 / 512,)),
    "C5": dict(n=131072, d=64, k=32, c=33, sep=3.0, seed=5001, kernel="gaussian", params=(1.0 / 64,)),
}
