"""Every test under /root/reference/test restated against the CPU oracle.

This is what pins the oracle (no Julia in the image, the reference's inputs are
unseeded ``rand`` so there are no golden vectors upstream).  Tolerance mirrors
Julia 0.4's ``@test_approx_eq``: 1e4 * length * eps(max|.|).
"""
import numpy as np
import pytest

import oracle as O


def approx_eq(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    tol = 1e4 * a.size * max(np.spacing(np.abs(a).max()), np.spacing(np.abs(b).max()))
    assert np.abs(a - b).max() <= tol, (np.abs(a - b).max(), tol)


A32 = np.array([[1., 2.], [3., 4.], [5., 6.]])


def test_translate():                               # test_common.jl:1-18
    b = np.array([1., 2.]); c = np.array([1., 2., 3.])
    approx_eq(O.translate(A32.copy(), 1.0), A32 + 1.0)
    approx_eq(O.translate(1.0, A32.copy()), A32 + 1.0)
    approx_eq(O.translate(A32.copy(), b), A32 + b[None, :])
    approx_eq(O.translate(c, A32.copy()), A32 + c[:, None])
    with pytest.raises(O.common.DimensionMismatch):
        O.translate(A32, c)
    with pytest.raises(O.common.DimensionMismatch):
        O.translate(b, A32)


A33 = np.array([[1., 2., 3.], [4., 5., 6.], [7., 8., 9.]])


def test_perturb():                                 # test_common.jl:20-32
    approx_eq(O.perturb(A33, 1.0), A33 + np.eye(3))
    with pytest.raises(O.common.DimensionMismatch):
        O.perturb(A32, 1.0)


def test_regularize():                              # test_common.jl:34-50
    a = np.trace(A33) / 3
    approx_eq(O.regularize(A33, 0.0, a), A33)
    approx_eq(O.regularize(A33, 0.5, a), 0.5 * (A33 + a * np.eye(3)))
    approx_eq(O.regularize(A33, 1.0, a), a * np.eye(3))
    with pytest.raises(O.common.DimensionMismatch):
        O.regularize(A32, 1.0, 1.0)


def test_syml_symu():                               # test_common.jl:53-75
    AL = np.array([[1., 2., 3.], [2., 5., 6.], [3., 6., 9.]])
    AU = np.array([[1., 4., 7.], [4., 5., 8.], [7., 8., 9.]])
    for T in (np.float32, np.float64):
        B = O.syml(A33.astype(T)); assert B.dtype == T; approx_eq(B, AL)
        B = O.symu(A33.astype(T)); assert B.dtype == T; approx_eq(B, AU)


def test_components_svd():                          # test_common.jl:77-86
    X = np.random.default_rng(1).random((20, 5))
    V1, D1 = O.components_svd(X.copy())
    _u, s, vt = np.linalg.svd(X, full_matrices=False)
    approx_eq(np.abs(V1), np.abs(vt.T))
    approx_eq(D1, s ** 2)


def test_components_eig():                          # test_common.jl:88-99
    X = np.random.default_rng(2).random((20, 5))
    S = X.T @ X
    V1, D1 = O.components_eig(S.copy())
    D2, V2 = np.linalg.eigh(S)
    approx_eq(np.abs(V1), np.abs(V2[:, ::-1]))
    approx_eq(np.abs(D1), np.abs(D2[::-1]))


def test_components_geig():                         # test_common.jl:101-115
    import scipy.linalg as sl
    rng = np.random.default_rng(3)
    X1, X2 = rng.random((20, 5)), rng.random((20, 5))
    S1, S2 = X1.T @ X1, X2.T @ X2
    V1, D1 = O.components_geig(S1.copy(), S2.copy())
    D2, V2 = sl.eigh(S1, S2)                        # B-normalised like sygvd / Julia eig(A,B)
    approx_eq(np.abs(V1), np.abs(V2[:, ::-1]))
    approx_eq(np.abs(D1), np.abs(D2[::-1]))


n, p = 20, 5                                        # test_kpca.jl:1-2


def test_pca_components_eig():                      # test_kpca.jl:4-16
    X = np.random.default_rng(4).random((n, p))
    H = X - X.mean(0)
    S = H.T @ H
    for alpha in (0.0, 0.5, 1.0):
        for eps in (0.0, 1.0):
            V1, D1 = O.pca_components_eig(H.copy(), alpha, eps)
            V2, D2 = O.components_eig(O.perturb(O.regularize(S, alpha, np.trace(S) / p), eps))
            if alpha != 1.0:                        # alpha = 1 is fully degenerate: any basis is valid
                approx_eq(np.abs(V1), np.abs(V2))
            approx_eq(D1, D2)


def test_pca_components_svd():                      # test_kpca.jl:18-27
    X = np.random.default_rng(5).random((n, p))
    H = X - X.mean(0)
    V1, D1 = O.pca_components_svd(H.copy())
    V2, D2 = O.components_svd(H.copy())
    approx_eq(V1, V2); approx_eq(D1, D2)


def test_pca():                                     # test_kpca.jl:29-42
    X = np.random.default_rng(6).random((n, p))
    H = X - X.mean(0)
    S = H.T @ H / (n - 1)
    for alpha in (0.0, 0.5):
        for eps in (0.0, 1.0):
            V1, D1 = O.pca(X.copy(), alpha, eps)
            D2, V2 = np.linalg.eigh(O.perturb(O.regularize(S, alpha, np.trace(S) / p), eps))
            approx_eq(np.abs(V1), np.abs(V2[:, ::-1]))
            approx_eq(np.abs(D1), np.abs(D2[::-1]))
            approx_eq(np.abs(O.transform_pca(V1, X)), np.abs(X @ V2[:, ::-1]))
    # alpha = 1: eigenvalues only (degenerate eigenspace)
    for eps in (0.0, 1.0):
        _V1, D1 = O.pca(X.copy(), 1.0, eps)
        D2 = np.linalg.eigvalsh(O.perturb(O.regularize(S, 1.0, np.trace(S) / p), eps))
        approx_eq(np.abs(D1), np.abs(D2[::-1]))


def test_kpca_equals_pca_for_linear_kernel():       # test_kpca.jl:44-56  (THE hot-path pin)
    X = np.random.default_rng(7).random((n, p))
    Xc = X - X.mean(0)
    S = Xc.T @ Xc / (n - 1)
    kappa = O.PolynomialKernel(1.0, 0.0, 1.0)       # equivalent to dot product
    for eps in (0.0, 1.0):
        V1, D1 = O.kpca(X.copy(), kappa, 0.0, eps)
        V2, D2 = O.pca(X.copy(), 0.0, eps)
        approx_eq((V1.T @ Xc @ Xc.T @ V1 / (n - 1))[:p, :p], V2.T @ S @ V2)
        approx_eq(np.abs(D1)[:p], np.abs(D2))


def test_lda_orphan_runs():                         # test_klda.jl:1-9 (no assertions upstream)
    rng = np.random.default_rng(8)
    nc, pp, k = 3, 3, 4
    y = np.repeat(np.arange(1, k + 1), nc)
    X = np.vstack([rng.random((nc, pp)) + 10 * rng.random((1, 3)) for _ in range(k)])
    freq = np.full(k, 1.0 / k)
    # upstream call (klda.jl:42 pencil (Sigma_b, Sigma_b)): with p = 3 <= k-1 Sigma_b is SPD and
    # every generalized eigenvalue is exactly 1 -- the bug is silent at the orphan test's sizes ...
    _V, L = O.lda(X.copy(), y, k, freq, fixed_pencil=False)
    approx_eq(L, np.ones(pp))
    # ... and fails inside LAPACK (B not SPD, info > n) as soon as p > k-1
    X5 = np.hstack([X, rng.random((nc * k, 2))])
    with pytest.raises(O.common.LAPACKException):
        O.lda(X5.copy(), y, k, freq, fixed_pencil=False)
    V, L = O.lda(X.copy(), y, k, freq, eps_w=1e-8)
    assert V.shape == (pp, pp) and np.all(np.diff(L) <= 1e-12)
    with pytest.raises(IndexError):
        O.class_counts(np.array([1, 5]), 4)         # klda.jl:8
