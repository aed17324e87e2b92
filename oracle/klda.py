"""Oracle: /root/reference/src/klda.jl restated (test infrastructure only).

Labels are 1-based like the reference (klda.jl:8).  The reference's
``lda_components!`` passes (Sigma_b, Sigma_b) to the generalized solver
(klda.jl:42) and ``klda!`` is exported but undefined; ``fixed_pencil=True``
(default) solves the intended (Sigma_b, Sigma_w) pencil -- SURVEY.md A.1/A.3.
ALL of this file is PARITY UNPINNED by the reference's own tests
(test/test_klda.jl is an orphan with no assertions).
"""
import numpy as np
from scipy.linalg import blas

from .common import (translate, perturb, regularize, syml, components_geig,
                     DimensionMismatch)
from .kernels import kernelmatrix


def class_counts(y, k):
    """klda.jl:5-12."""
    y = np.asarray(y)
    counts = np.zeros(k, dtype=np.int64)
    for i, yi in enumerate(y):
        if not (1 <= yi <= k):
            raise IndexError("Index %d out of range." % (i + 1))   # klda.jl:8
        counts[yi - 1] += 1
    return counts


def class_totals(X, y, k):
    """klda.jl:14-22: k x p class sums."""
    X = np.asarray(X, dtype=np.float64); y = np.asarray(y)
    n, p = X.shape
    if len(y) != n:
        raise DimensionMismatch("X and y must have the same number of rows.")
    M = np.zeros((k, p))
    np.add.at(M, y - 1, X)
    return M


def center_rows(X, M, y):
    """klda.jl:24-30: X[i,:] -= M[y[i],:]."""
    return np.asarray(X, dtype=np.float64) - np.asarray(M)[np.asarray(y) - 1]


def lda_matrices(X, y, k, frequency):
    """klda.jl:45-55 ``lda_matrices!`` -> (H_b, H_w, mu)."""
    X = np.asarray(X, dtype=np.float64)
    n, p = X.shape
    frequency = np.asarray(frequency, dtype=np.float64)
    M = class_totals(X, y, k)                               # klda.jl:47
    M = M * (1.0 / class_counts(y, k))[:, None]             # klda.jl:48
    H_w = center_rows(X, M, y)                              # klda.jl:49
    H_w = H_w * (1.0 / (n - k))                             # klda.jl:50
    mu = frequency @ M                                      # klda.jl:51
    M = translate(M, -mu)                                   # klda.jl:52
    H_b = np.sqrt(frequency)[:, None] * M                   # klda.jl:53
    return H_b, H_w, mu


def scatter_matrices(H_b, alpha_b, eps_b, H_w, alpha_w, eps_w):
    """klda.jl:35-41: Sigma_b, Sigma_w with shrink / perturb applied."""
    S_b = syml(blas.dsyrk(1.0, np.asarray(H_b, float), trans=1, lower=0))   # klda.jl:35
    if alpha_b != 0:
        S_b = regularize(S_b, alpha_b, np.trace(S_b) / S_b.shape[0])
    if eps_b != 0:
        S_b = perturb(S_b, eps_b)
    S_w = syml(blas.dsyrk(1.0, np.asarray(H_w, float), trans=1, lower=0))   # klda.jl:39
    if alpha_w != 0:
        S_w = regularize(S_w, alpha_w, np.trace(S_w) / S_w.shape[0])
    if eps_w != 0:
        S_w = perturb(S_w, eps_w)
    return S_b, S_w


def lda_components(H_b, alpha_b, eps_b, H_w, alpha_w, eps_w, fixed_pencil=True):
    """klda.jl:34-43 ``lda_components!``.  ``fixed_pencil=False`` reproduces the
    upstream call ``components_geig!(Sigma_b, Sigma_b)`` (klda.jl:42), which
    fails inside LAPACK for rank-deficient Sigma_b."""
    S_b, S_w = scatter_matrices(H_b, alpha_b, eps_b, H_w, alpha_w, eps_w)
    if fixed_pencil:
        return components_geig(S_b, S_w)
    return components_geig(S_b, S_b.copy())                 # klda.jl:42 as written


def lda(X, y, k, frequency, alpha_b=0.0, eps_b=0.0, alpha_w=0.0, eps_w=0.0,
        fixed_pencil=True):
    """klda.jl:57-61 ``lda!`` (out of place)."""
    H_b, H_w, _mu = lda_matrices(X, y, k, frequency)
    return lda_components(H_b, alpha_b, eps_b, H_w, alpha_w, eps_w, fixed_pencil)


def transform_lda(W, Z):
    """klda.jl:65."""
    return np.asarray(Z) @ np.asarray(W)


def klda(X, kernel, y, k, frequency, alpha_b=0.0, eps_b=0.0, alpha_w=0.0,
         eps_w=0.0, ncomp=None):
    """Defined semantics for the exported-but-undefined ``klda!``
    (FeatureExtraction.jl:11): feed the uncentred kernel matrix
    K = k(X,X) through the LDA pipeline as the data matrix with the pencil
    fixed to (Sigma_b, Sigma_w).  eps_w > 0 is required (Sigma_w is singular).
    Returns the top ``ncomp`` (default k-1) pairs, descending."""
    K = kernelmatrix(kernel, np.asarray(X, dtype=np.float64))
    V, L = lda(K, y, k, frequency, alpha_b, eps_b, alpha_w, eps_w, True)
    ncomp = (k - 1) if ncomp is None else ncomp
    return V[:, :ncomp].copy(), L[:ncomp].copy()


def transform_klda(X, kernel, W, Z):
    """klda.jl:71 with the missing kernel argument restored: k(Z,X) W, no centring."""
    return kernelmatrix(kernel, np.asarray(Z, float), np.asarray(X, float)) @ W
