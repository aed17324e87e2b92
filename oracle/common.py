"""Oracle: /root/reference/src/common.jl restated (test infrastructure only).

All functions are out-of-place on NumPy float arrays unless named ``*_``.
"""
import numpy as np
from scipy.linalg import lapack


class DimensionMismatch(ValueError):
    """Mirror of Julia's DimensionMismatch (common.jl:16,24,33,42)."""


def translate(A, b):
    """common.jl:7-29 ``translate!``: A += b for scalar b, a length-size(A,1)
    vector given FIRST (``translate!(b, A)``: adds b[i] to row i) or a
    length-size(A,2) vector given SECOND (``translate!(A, b)``: adds b[j] to
    column j).  Call as translate(A, b) or translate(b, A) like the reference."""
    if np.isscalar(A) or (isinstance(A, np.ndarray) and A.ndim == 0):
        A, b = b, A                              # common.jl:13 translate!(b::T, A)
    if np.isscalar(b) or (isinstance(b, np.ndarray) and b.ndim == 0):
        return np.asarray(A) + b                 # common.jl:7-12
    A = np.asarray(A); b = np.asarray(b)
    if A.ndim == 1 and b.ndim == 2:              # translate!(b::Vector, A::Matrix) common.jl:15-21
        vec, M = A, b
        if M.shape[0] != vec.shape[0]:
            raise DimensionMismatch("first dimension of A does not match length of b")
        return M + vec[:, None]
    if A.ndim == 2 and b.ndim == 1:              # translate!(A::Matrix, b::Vector) common.jl:23-29
        if A.shape[1] != b.shape[0]:
            raise DimensionMismatch("second dimension of A does not match length of b")
        return A + b[None, :]
    raise TypeError("translate: unsupported argument combination")


def _chksquare(A, msg="Matrix A must be square."):
    if A.ndim != 2 or A.shape[0] != A.shape[1]:
        raise DimensionMismatch(msg)
    return A.shape[0]


def perturb(A, eps):
    """common.jl:32-38 ``perturb!``: A := A + eps*I."""
    A = np.array(A, copy=True)
    n = _chksquare(A)
    A[np.arange(n), np.arange(n)] += eps
    return A


def regularize(A, alpha, beta):
    """common.jl:41-50 ``regularize!``: A := (1-alpha)*A + alpha*beta*I."""
    A = np.array(A, copy=True)
    n = _chksquare(A)
    A *= (1 - alpha)                             # common.jl:45-47
    A[np.arange(n), np.arange(n)] += alpha * beta  # common.jl:48
    return A


def syml(S):
    """common.jl:53-60 ``syml!``: copy the upper triangle onto the lower."""
    S = np.array(S, copy=True)
    if S.ndim != 2 or S.shape[0] != S.shape[1]:
        raise ValueError("S must be square")      # ArgumentError, common.jl:54
    iu = np.triu_indices(S.shape[0], 1)
    S[(iu[1], iu[0])] = S[iu]
    return S


def symu(S):
    """common.jl:63-70 ``symu!``: copy the lower triangle onto the upper."""
    S = np.array(S, copy=True)
    if S.ndim != 2 or S.shape[0] != S.shape[1]:
        raise ValueError("S must be square")
    il = np.tril_indices(S.shape[0], -1)
    S[(il[1], il[0])] = S[il]
    return S


class LAPACKException(RuntimeError):
    """common.jl:90-92 ``@lapackerror`` (info > 0)."""
    def __init__(self, info):
        super().__init__("LAPACK info=%d" % info)
        self.info = info


def _lapackerror(info):
    if info < 0:                                  # common.jl:87-89
        raise ValueError("invalid argument #%d to LAPACK call" % (-info))
    if info > 0:
        raise LAPACKException(info)


def syevd(jobz, uplo, A):
    """common.jl:107-143: ``?syevd`` with jobz/uplo, workspace query + solve.
    Returns (W ascending, V) for jobz='V' or W for jobz='N'."""
    A = np.asarray(A)
    if uplo not in ("U", "L"):
        raise ValueError("invalid uplo")         # common.jl:101-105
    _chksquare(A)
    f = lapack.dsyevd if A.dtype == np.float64 else lapack.ssyevd
    w, v, info = f(A, compute_v=1 if jobz == "V" else 0, lower=1 if uplo == "L" else 0)
    _lapackerror(info)                            # common.jl:132
    return (w, v) if jobz == "V" else w


def components_svd(X):
    """common.jl:152-158 ``components_svd!``: thin SVD via gesdd, returns
    (V, sigma**2), singular values descending."""
    X = np.asarray(X)
    f = lapack.dgesdd if X.dtype == np.float64 else lapack.sgesdd
    _u, s, vt, info = f(X, compute_uv=1, full_matrices=0)
    _lapackerror(info)
    return vt.T.copy(), s ** 2


def components_eig(S):
    """common.jl:161-165 ``components_eig!``: syevd('V','U'), then reverse to
    descending order (fresh copies)."""
    S = np.asarray(S)
    _chksquare(S)
    D, V = syevd("V", "U", S)
    return V[:, ::-1].copy(), D[::-1].copy()


def components_geig(S_m, S_x):
    """common.jl:168-174 ``components_geig!``: sygvd(itype=1,'V','U') for
    S_m v = lambda S_x v, descending; vectors normalised v' S_x v = 1."""
    S_m = np.asarray(S_m); S_x = np.asarray(S_x)
    p = _chksquare(S_x, "Covariance matrix for X must be square.")
    if S_m.shape[1] != p:                         # common.jl:171
        raise DimensionMismatch("Covariance matrices for X and M must be of the same order.")
    f = lapack.dsygvd if S_m.dtype == np.float64 else lapack.ssygvd
    w, v, info = f(S_m, S_x, itype=1, jobz="V", uplo="U")
    _lapackerror(info)
    return v[:, ::-1].copy(), w[::-1].copy()
