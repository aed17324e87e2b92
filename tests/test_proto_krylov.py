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
