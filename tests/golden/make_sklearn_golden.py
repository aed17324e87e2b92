"""Mints INDEPENDENT golden vectors from scikit-learn's KernelPCA (not from oracle/): the kernels the
reference's own tests leave unpinned (Gaussian, polynomial degree 3, sigmoid -- SURVEY.md 8c) and the
out-of-sample transform, at n = 512 and n = 4096.

    python tests/golden/make_sklearn_golden.py        # scikit-learn >= 1.0, a few seconds

Conventions bridged here (and nowhere else):
  sklearn rbf      exp(-gamma |x-y|^2)            == GaussianKernel(gamma)
  sklearn poly     (gamma x'y + coef0)^degree     == PolynomialKernel(gamma, coef0, degree)   (test_kpca.jl:49 order)
  sklearn sigmoid  tanh(gamma x'y + coef0)        == SigmoidKernel(gamma, coef0)
  eigenvalues_  = eigenvalues of the double-centred K   -> reference returns them / (n-1)   (kpca.jl:56)
  eigenvectors_ = unit-norm eigenvectors (sign free)      == the reference's V              (common.jl:164)
  transform(Z)  = K~_z V diag(eigenvalues_)^-1/2          -> reference-shaped transform = K~_z V
Inputs are regenerated from (generator, seed) by the tests; outputs are stored."""
import os
import sys

import numpy as np
from sklearn.decomposition import KernelPCA

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fxb200.synth import mix           # noqa: E402  (generator only; no library call)

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (n, d, c, sep, seed, kernel, (gamma, coef0, degree), k, m, zseed)
    "sk_rbf_512": (512, 12, 5, 3.0, 8101, "rbf", (1.0 / 12, 0.0, 0.0), 4, 48, 8102),
    "sk_poly3_512": (512, 12, 5, 3.0, 8103, "poly", (1.0 / 12, 1.0, 3.0), 4, 48, 8104),
    "sk_sigmoid_512": (512, 12, 5, 3.0, 8105, "sigmoid", (1.0 / 48, 0.25, 0.0), 4, 48, 8106),
    "sk_rbf_4096": (4096, 32, 9, 3.0, 8107, "rbf", (1.0 / 32, 0.0, 0.0), 8, 64, 8108),
    "sk_poly3_4096": (4096, 32, 9, 3.0, 8109, "poly", (1.0 / 32, 1.0, 3.0), 8, 64, 8110),
    "sk_sigmoid_4096": (4096, 32, 9, 3.0, 8111, "sigmoid", (1.0 / 128, 0.25, 0.0), 8, 64, 8112),
}


def main():
    for name, (n, d, c, sep, seed, kname, (gamma, coef0, degree), k, m, zseed) in CASES.items():
        X, _ = mix(n, d, c, sep, seed)
        Z, _ = mix(m, d, c, sep, zseed)
        kw = dict(gamma=gamma)
        if kname in ("poly", "sigmoid"):
            kw["coef0"] = coef0
        if kname == "poly":
            kw["degree"] = int(degree)
        kp = KernelPCA(n_components=k + 4, kernel=kname, eigen_solver="dense", **kw)
        kp.fit(np.ascontiguousarray(X))
        ev = kp.eigenvalues_.copy()
        V = kp.eigenvectors_[:, :k].copy()
        T = kp.transform(np.ascontiguousarray(Z))[:, :k] * np.sqrt(ev[:k])[None, :]
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            meta=np.array([n, d, c, sep, seed, k, m, zseed], dtype=np.float64),
            kernel=np.array(kname), params=np.array([gamma, coef0, degree]),
            L=ev / (n - 1), V=V, T=T, x_checksum=np.array([X.sum(), (X * X).sum()]),
            how=np.array("sklearn %s KernelPCA(eigen_solver='dense')" % __import__("sklearn").__version__))
        print(name, "L =", ev[:k + 2] / (n - 1))


if __name__ == "__main__":
    main()
