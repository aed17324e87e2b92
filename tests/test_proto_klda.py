"""CPU tests of the NumPy model of the iterative KLDA solve (tests/proto_klda_lanczos.py): the Woodbury
collapse, the pentadiagonal Galerkin solve with its residual formula, and the refinement reach the
accuracy of a dense LAPACK solve of the explicitly formed pencil."""
import numpy as np
import pytest

from fxb200.synth import mix
from proto_klda_lanczos import klda_solve, solve_A


def _pencil(n, d, c, rel, seed=71):
    X, y0 = mix(n, d, c, 2.0, seed)
    cnt = np.bincount(y0, minlength=c).astype(float)
    freq = cnt / n
    sq = (X * X).sum(1)
    K = np.exp(-(1.0 / d) * np.maximum(sq[:, None] + sq[None, :] - 2.0 * X @ X.T, 0.0))
    E = np.zeros((n, c)); E[np.arange(n), y0] = 1.0
    M = (E.T @ K) / cnt[:, None]
    Hw = (K - M[y0]) / (n - c)                                      # klda.jl:49-50
    Sw = Hw.T @ Hw
    shift = rel * np.trace(Sw) / n
    return K, y0, freq, Sw + shift * np.eye(n), shift


def test_galerkin_residual_formula_is_exact():
    K, y0, freq, Swr, shift = _pencil(500, 16, 4, 1e-6)
    n = K.shape[0]; cw = 1.0 / (n - 4) ** 2
    E = np.zeros((n, 4)); E[np.arange(n), y0] = 1.0
    Y, passes, ests = solve_A(K, E, cw, shift, 1e-9)
    true = np.linalg.norm(E - (cw * K @ (K @ Y) + shift * Y), axis=0) / np.linalg.norm(E, axis=0)
    assert passes < 200
    assert np.all(true <= 1e-9 * 1.05) and np.allclose(true, ests, rtol=0.05)


@pytest.mark.parametrize("n,d,c,rel", [(600, 8, 3, 1e-3), (900, 32, 5, 1e-6)])
def test_klda_solve_matches_dense_solve(n, d, c, rel):
    K, y0, freq, Swr, shift = _pencil(n, d, c, rel)
    cw = 1.0 / (n - c) ** 2
    x, h, passes, hist = klda_solve(K, y0, c, freq, cw, shift, cg_tol=1e-11)
    xd = np.linalg.solve(Swr, h)
    assert np.linalg.norm(x - xd) <= 1e-8 * np.linalg.norm(xd)
    # the true residual reaches the level of the dense LAPACK solve (within 4x)
    res = np.linalg.norm(Swr @ x - h, axis=0) / np.linalg.norm(h, axis=0)
    resd = np.linalg.norm(Swr @ xd - h, axis=0) / np.linalg.norm(h, axis=0)
    assert res.max() <= max(4 * resd.max(), 1e-11), (res.max(), resd.max(), hist)
    S = h.T @ x
    lam = np.linalg.eigvalsh(0.5 * (S + S.T))[::-1][:c - 1]
    Sd = h.T @ xd
    lamd = np.linalg.eigvalsh(0.5 * (Sd + Sd.T))[::-1][:c - 1]
    assert np.abs(lam - lamd).max() <= 1e-9 * lamd[0]
