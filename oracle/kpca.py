"""Oracle: /root/reference/src/kpca.jl restated (test infrastructure only)."""
import numpy as np
from scipy.linalg import blas

from .common import (translate, perturb, regularize, syml, components_eig,
                     components_svd)
from .kernels import kernelmatrix, centerkernelmatrix


def pca_components_eig(H, alpha, eps):
    """kpca.jl:8-13 ``pca_components_eig!``: Sigma = H'H (syrk 'U','T' + syml!),
    shrink towards tr/p * I, perturb, full eigendecomposition (descending)."""
    H = np.asarray(H, dtype=np.float64)
    S = syml(blas.dsyrk(1.0, H, trans=1, lower=0))          # kpca.jl:9
    if alpha != 0:
        S = regularize(S, alpha, np.trace(S) / S.shape[0])  # kpca.jl:10
    if eps != 0:
        S = perturb(S, eps)                                 # kpca.jl:11
    return components_eig(S)                                # kpca.jl:12


def pca_components_svd(H):
    """kpca.jl:15 alias of components_svd!."""
    return components_svd(H)


def pca(X, alpha=0.0, eps=0.0):
    """kpca.jl:17-31 ``pca`` / ``pca!``: centre columns, SVD path when
    alpha == eps == 0 else covariance-eig path; eigenvalues / (n-1)."""
    X = np.array(X, dtype=np.float64, copy=True)
    n, p = X.shape
    mu = X.mean(axis=0)                                     # kpca.jl:19
    H = translate(X, -mu)                                   # kpca.jl:20
    if alpha == 0 and eps == 0:
        V, L = pca_components_svd(H)                        # kpca.jl:21
    else:
        V, L = pca_components_eig(H, alpha, eps * (n - 1))
    return V, L / (n - 1)                                   # kpca.jl:22


def transform_pca(W, Z):
    """kpca.jl:37."""
    return np.asarray(Z) @ np.asarray(W)


def kpca(X, kernel, alpha=0.0, eps=0.0, center=True, return_state=False):
    """kpca.jl:49-57 ``kpca``: K = kernelmatrix; centerkernelmatrix!;
    regularize!(K, alpha, tr(K)/n); perturb!(K, eps*(n-1)); components_eig!
    (dsyevd 'V','U', reversed to descending); eigenvalues / (n-1).
    Returns ALL n eigenpairs like the reference.  ``center=False`` is this
    engine's addition (the reference always centres, kpca.jl:52)."""
    X = np.asarray(X, dtype=np.float64)
    n, p = X.shape                                          # kpca.jl:50
    K = kernelmatrix(kernel, X)                             # kpca.jl:51
    r = np.zeros(n); mu = 0.0
    if center:
        K, r, mu = centerkernelmatrix(K)                    # kpca.jl:52
    if alpha != 0:
        K = regularize(K, alpha, np.trace(K) / n)           # kpca.jl:53
    if eps != 0:
        K = perturb(K, eps * (n - 1))                       # kpca.jl:54
    V, L = components_eig(K)                                # kpca.jl:55
    L = L / (n - 1)                                         # kpca.jl:56
    if return_state:
        return V, L, dict(r=r, mu=mu)
    return V, L


def kpca_topk(X, kernel, k, alpha=0.0, eps=0.0, center=True):
    """First k columns / values of :func:`kpca` (the reference has no k;
    SURVEY.md A.2) plus the centring state the transform needs."""
    V, L, st = kpca(X, kernel, alpha, eps, center, return_state=True)
    return V[:, :k].copy(), L[:k].copy(), st


def transform_kpca(X, kernel, W, Z, center=True):
    """Defined semantics for kpca.jl:64-68 (the reference body throws
    UndefVarError at :67):
        (k(Z,X) - 1 r' - r_z 1' + mu) W
    with r, mu the training column means / grand mean and r_z the per-query
    row means.  PARITY UNPINNED by the reference."""
    X = np.asarray(X, dtype=np.float64); Z = np.asarray(Z, dtype=np.float64)
    Kzx = kernelmatrix(kernel, Z, X)                        # kpca.jl:65
    if not center:
        return Kzx @ W
    K = kernelmatrix(kernel, X)
    n = K.shape[0]
    r = K.sum(axis=0) / n
    mu = r.sum() / n
    rz = Kzx.mean(axis=1)                                   # kpca.jl:66
    return (Kzx - r[None, :] - rz[:, None] + mu) @ W


def transform_kpca_reference_intent(X, kernel, W, Z):
    """What kpca.jl:66-67 was reaching for: subtract only the per-query row
    mean, then project (documentation / tests only)."""
    Kzx = kernelmatrix(kernel, np.asarray(Z, float), np.asarray(X, float))
    return (Kzx - Kzx.mean(axis=1)[:, None]) @ W
