"""Oracle cross-checks beyond the reference's own tests: scikit-learn KernelPCA
(independent implementation, dense solver) and the identities of SURVEY.md 4/7.3
that the CUDA design relies on."""
import numpy as np
import pytest

import oracle as O
import fxb200
from fxb200.synth import mix


def _sk_kernels():
    return [
        (O.GaussianKernel(0.1), dict(kernel="rbf", gamma=0.1)),
        (O.PolynomialKernel(0.1, 1.0, 3.0), dict(kernel="poly", gamma=0.1, coef0=1.0, degree=3)),
        (O.SigmoidKernel(0.01, 0.5), dict(kernel="sigmoid", gamma=0.01, coef0=0.5)),
    ]


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_kpca_matches_sklearn(idx):
    from sklearn.decomposition import KernelPCA
    kern, skargs = _sk_kernels()[idx]
    X, _ = mix(300, 10, 3, 3.0, 11)
    k = 2
    V, L, st = O.kpca_topk(X, kern, k)
    sk = KernelPCA(n_components=k, eigen_solver="dense", **skargs).fit(X)
    np.testing.assert_allclose(sk.eigenvalues_ / (len(X) - 1), L, rtol=1e-10)
    # sklearn's transform = full centring formula then 1/sqrt(lambda) scaling
    Z, _ = mix(50, 10, 3, 3.0, 12)
    T_ours = O.transform_kpca(X, kern, V, Z)
    T_sk = sk.transform(Z) * np.sqrt(sk.eigenvalues_)[None, :]
    np.testing.assert_allclose(np.abs(T_ours), np.abs(T_sk), rtol=1e-8, atol=1e-10)


def test_gaussian_direct_vs_gemm():
    X, _ = mix(200, 10, 3, 3.0, 13)
    kern = O.GaussianKernel(0.1)
    np.testing.assert_allclose(O.kernelmatrix(kern, X, method="direct"),
                               O.kernelmatrix(kern, X), rtol=0, atol=1e-13)


def test_transform_identities():
    X, _ = mix(300, 10, 3, 3.0, 14)
    kern = O.GaussianKernel(0.1)
    n = len(X); k = 2
    V, L, st = O.kpca_topk(X, kern, k)
    # in-sample: transform(X) == V * diag(lambda_unscaled)
    T = O.transform_kpca(X, kern, V, X)
    np.testing.assert_allclose(T, V * (L * (n - 1))[None, :], atol=1e-12)
    # 1'W ~ 0 for components with non-zero eigenvalue
    assert np.abs(V.sum(0)).max() < 1e-12
    # so the transform only needs a k-vector bias:  K_zx W - 1 (r'W)
    Z, _ = mix(40, 10, 3, 3.0, 15)
    Kzx = O.kernelmatrix(kern, Z, X)
    np.testing.assert_allclose(O.transform_kpca(X, kern, V, Z),
                               Kzx @ V - (st["r"] @ V)[None, :], atol=1e-12)


def test_alpha_eps_fold_into_eigenvalues():
    """regularize!/perturb! (kpca.jl:53-54) only remap eigenvalues:
    lambda' = (1-a) lambda + a tr(K~)/n + eps (n-1)."""
    X, _ = mix(120, 6, 3, 3.0, 16)
    kern = O.GaussianKernel(0.2)
    n = len(X)
    V0, L0 = O.kpca(X, kern)
    Kc, _, _ = O.centerkernelmatrix(O.kernelmatrix(kern, X))
    for alpha, eps in [(0.3, 0.0), (0.0, 0.7), (0.5, 1.0)]:
        V1, L1 = O.kpca(X, kern, alpha, eps)
        lam = (1 - alpha) * L0 * (n - 1) + alpha * np.trace(Kc) / n + eps * (n - 1)
        np.testing.assert_allclose(L1, lam / (n - 1), rtol=1e-11, atol=1e-13)
        np.testing.assert_allclose(np.abs(V1[:, :3]), np.abs(V0[:, :3]), atol=1e-9)


def test_klda_reduced_form():
    """the non-zero generalized eigenpairs of (Sigma_b, Sigma_w + eps I)
    come from the c x c matrix S = H_b (Sigma_w+eps I)^-1 H_b'."""
    X, y0 = mix(150, 5, 4, 2.0, 17)
    y = y0 + 1; c = 4
    freq = np.bincount(y0, minlength=c) / len(y)
    kern = O.GaussianKernel(0.2)
    K = O.kernelmatrix(kern, X)
    H_b, H_w, _ = O.lda_matrices(K, y, c, freq)
    S_b, S_w = O.scatter_matrices(H_b, 0, 0, H_w, 0, 0)
    eps_w = 1e-6 * np.trace(S_w) / len(X)
    V, L = O.klda(X, kern, y, c, freq, eps_w=eps_w)
    Sw = S_w + eps_w * np.eye(len(X))
    G = np.linalg.solve(Sw, H_b.T)                 # n x c
    S = H_b @ G
    lam, U = np.linalg.eigh(0.5 * (S + S.T))
    lam, U = lam[::-1], U[:, ::-1]
    np.testing.assert_allclose(lam[:c - 1], L, rtol=1e-8)
    A = G @ U[:, :c - 1] / np.sqrt(lam[:c - 1])[None, :]
    np.testing.assert_allclose(np.abs(A), np.abs(V), rtol=1e-5, atol=1e-7 * np.abs(V).max())
    # H_w == (I - P) K / (n - c) with P the class-averaging projector
    P = np.zeros((len(X), len(X)))
    for cl in range(1, c + 1):
        idx = np.where(y == cl)[0]
        P[np.ix_(idx, idx)] = 1.0 / len(idx)
    np.testing.assert_allclose(H_w, (np.eye(len(X)) - P) @ K / (len(X) - c), atol=1e-14)
