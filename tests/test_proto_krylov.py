"""Validates the eigensolver algorithm (NumPy prototype of csrc/solver.cu)
against LAPACK dsyevd on the oracle's centred kernel matrices."""
import numpy as np
import pytest

import oracle as O
from fxb200.synth import mix
from proto_krylov import topk, jacobi_eigh


def subspace_angle(A, B):
    Qa, _ = np.linalg.qr(A); Qb, _ = np.linalg.qr(B)
    s = np.linalg.svd(Qa.T @ Qb, compute_uv=False)
    # sin of the largest principal angle, computed stably
    R = Qb - Qa @ (Qa.T @ Qb)
    return np.linalg.norm(R, 2)


@pytest.mark.parametrize("kern,k,c", [
    (O.GaussianKernel(1 / 32), 8, 9),
    (O.PolynomialKernel(1 / 32, 1.0, 3.0), 4, 5),
    (O.SigmoidKernel(1 / 64, 0.1), 3, 4),
])
def test_proto_matches_dense(kern, k, c):
    X, _ = mix(1500, 32, c, 3.0, 21)
    K = O.kernelmatrix(kern, X)
    Kc, _, _ = O.centerkernelmatrix(K)
    V, th, info = topk(lambda Vf: Kc @ Vf, len(X), 0, k)
    assert info["converged"], info["hist"]
    Vd, Ld = O.components_eig(Kc)
    np.testing.assert_allclose(th, Ld[:k], rtol=1e-10)
    assert subspace_angle(V, Vd[:, :k]) < 1e-8
    assert np.abs(np.abs((V * Vd[:, :k]).sum(0)) - 1).max() < 1e-8


def test_proto_slow_decay_restarts():
    """No spectral gap (iid Gaussian data): must still converge through restarts."""
    rng = np.random.default_rng(5)
    X = rng.standard_normal((1200, 16))
    Kc, _, _ = O.centerkernelmatrix(O.kernelmatrix(O.GaussianKernel(1 / 16), X))
    k = 6
    V, th, info = topk(lambda Vf: Kc @ Vf, len(X), 0, k, max_steps=400)
    assert info["converged"]
    Vd, Ld = O.components_eig(Kc)
    np.testing.assert_allclose(th, Ld[:k], rtol=1e-10)


def test_jacobi_parallel_order():
    rng = np.random.default_rng(6)
    for m in (5, 16, 33, 64):
        A = rng.standard_normal((m, m)); A = A + A.T
        d, V, sweeps = jacobi_eigh(A)
        w = np.linalg.eigvalsh(A)[::-1]
        np.testing.assert_allclose(d, w, rtol=0, atol=1e-13 * np.abs(w).max())
        assert np.abs(V.T @ V - np.eye(m)).max() < 1e-13
        assert np.abs(A @ V - V * d[None, :]).max() < 1e-12 * np.abs(w).max()
        assert sweeps <= 12


def test_block_orthonormalisation_cholqr_and_svqb_fallback():
    """Both branches of the block orthonormalisation (smalldense.cu chol_factors_kernel /
    svqb_*): a well conditioned block takes CholQR, a rank-deficient one falls back to SVQB; in both
    cases W = W_orth Cfac, and two passes give orthonormal columns wherever the block has rank."""
    import proto_krylov as P
    rng = np.random.default_rng(5)
    ops = P.LocalOps()
    W = rng.standard_normal((500, 16)) * np.logspace(0, 3, 16)[None, :]       # badly scaled, full rank
    assert P.chol_factors((W.T @ W) / np.outer(np.linalg.norm(W, axis=0), np.linalg.norm(W, axis=0)), 1e-8) is not None
    Q1, C1, nd1 = P.svqb(W, ops)
    Q2, C2, nd2 = P.svqb(Q1, ops)
    assert nd1 == 0 and nd2 == 0
    np.testing.assert_allclose(Q2.T @ Q2, np.eye(16), atol=1e-14)
    np.testing.assert_allclose(Q2 @ (C2 @ C1), W, rtol=0, atol=1e-10 * np.abs(W).max())
    # same block through the SVQB branch only: same column space
    Q1s, _, _ = P.svqb(W, ops, chol_thr=None)
    Q2s, _, _ = P.svqb(Q1s, ops, chol_thr=None)
    np.testing.assert_allclose(Q2 @ (Q2.T @ Q2s), Q2s, atol=1e-12)
    # rank 10 < 16: Cholesky must refuse, SVQB must report the deficiency
    Wd = rng.standard_normal((500, 10)) @ rng.standard_normal((10, 16))
    G = Wd.T @ Wd
    d = np.sqrt(np.diag(G))
    assert P.chol_factors(G / np.outer(d, d), 1e-8) is None
    Qd, Cd, ndd = P.svqb(Wd, ops)
    assert ndd == 6
    np.testing.assert_allclose(Qd @ Cd, Wd, rtol=0, atol=1e-9 * np.abs(Wd).max())


def test_merged_reprojection_and_block_gram_downdate():
    """solver.cu::reproject_and_orthonormalise reduces Hs = [Q W]'W ONCE and takes the Gram matrix of the
    re-projected block as G = W'W - H2'H2 (smalldense.cu gram_downdate_kernel) instead of forming
    W - Q H2 first and reducing a second time.  That is safe because H2 is a SECOND-pass coefficient block
    (~1e-8 of W after the first Gram-Schmidt pass), so the subtraction cannot cancel: the downdated Gram
    equals the explicit one to rounding and CholQR from it gives an orthonormal block orthogonal to Q."""
    import proto_krylov as P
    rng = np.random.default_rng(17)
    n, m, b = 4000, 48, 16
    Q, _ = np.linalg.qr(rng.standard_normal((n, m)))
    W0 = rng.standard_normal((n, b)) * np.logspace(0, 2, b)[None, :] + Q @ rng.standard_normal((m, b)) * 50.0
    W = W0 - Q @ (Q.T @ W0)                              # first pass (project_out)
    H2 = Q.T @ W                                         # second-pass coefficients: tiny
    assert np.abs(H2).max() <= 1e-11 * np.abs(W).max() * n ** 0.5
    G_merged = W.T @ W - H2.T @ H2
    Wp = W - Q @ H2
    G_explicit = Wp.T @ Wp
    assert np.abs(G_merged - G_explicit).max() <= 1e-14 * np.abs(G_explicit).max()
    d = np.sqrt(np.diag(G_merged))
    fac = P.chol_factors(G_merged / np.outer(d, d), 1e-8)
    assert fac is not None
    L = np.linalg.cholesky(G_merged)
    Wo = np.linalg.solve(L, Wp.T).T                      # (W - Q H2) L^-T
    assert np.abs(Wo.T @ Wo - np.eye(b)).max() <= 1e-12
    assert np.abs(Q.T @ Wo).max() <= 1e-13
    # the same identity would NOT be safe with FIRST-pass coefficients (a block with a large component
    # inside span(Q)): there it cancels, which is why only the second pass is merged
    Wb = rng.standard_normal((n, b)) + Q @ rng.standard_normal((m, b)) * 1e6
    H1 = Q.T @ Wb
    Wb_p = Wb - Q @ H1
    G_bad = Wb.T @ Wb - H1.T @ H1
    assert np.abs(G_bad - Wb_p.T @ Wb_p).max() > 1e-6 * np.abs(Wb_p.T @ Wb_p).max()
