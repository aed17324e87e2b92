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
