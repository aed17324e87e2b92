"""Oracle: the MLKernels calls on the hot path, restated (test infrastructure only).

MLKernels is an UN-VENDORED dependency of the reference (REQUIRE:2 ``MLKernels
0.1`` -- an old-style lower bound, not an exact pin; ``importall MLKernels`` at
src/FeatureExtraction.jl:5).  Its source is not under /root/reference and there
is no network, so its published definitions are restated here and anchored on
the reference's own call sites:

  kernelmatrix(k, X)        kpca.jl:51   rows of X are observations (kpca.jl:50,
                                          test_kpca.jl:47-53), full symmetric n x n
  kernelmatrix(k, Z, X)     kpca.jl:65   m x n cross matrix K_zx[i,j] = k(z_i, x_j)
  centerkernelmatrix!(K)    kpca.jl:52   symmetric double centring, pinned by
                                          test_kpca.jl:44-56 (identity only holds
                                          if K~ = Xc Xc')
  PolynomialKernel(a,c,d)   test_kpca.jl:49  (a x'y + c)^d, argument order pinned
                                          by the same test ((1,0,1) == dot product)
  GaussianKernel(alpha)     exp(-alpha ||x-y||^2)      PARITY UNPINNED
  SigmoidKernel(a,c)        tanh(a x'y + c)            PARITY UNPINNED
"""
import numpy as np


class Kernel:
    """Stand-in for MLKernels' ``Kernel{T}`` (kpca.jl:49,64; klda.jl:71)."""
    kernel_id = -1

    def params(self):
        raise NotImplementedError

    def of_dot(self, G, nx=None, ny=None):
        """Map a matrix of dot products G[i,j] = x_i . y_j (plus squared norms
        for distance kernels) to kernel values."""
        raise NotImplementedError


class GaussianKernel(Kernel):
    """k(x,y) = exp(-alpha * ||x-y||^2), alpha > 0 a direct scale parameter."""
    kernel_id = 0

    def __init__(self, alpha=1.0):
        if not alpha > 0:
            raise ValueError("alpha must be positive")
        self.alpha = float(alpha)

    def params(self):
        return (self.alpha, 0.0, 0.0)

    def of_dot(self, G, nx=None, ny=None):
        d2 = nx[:, None] + ny[None, :] - 2.0 * G
        np.maximum(d2, 0.0, out=d2)              # cancellation guard
        return np.exp(-self.alpha * d2)


class PolynomialKernel(Kernel):
    """k(x,y) = (a * x'y + c)^d -- constructor order (a, c, d) as test_kpca.jl:49."""
    kernel_id = 1

    def __init__(self, a=1.0, c=1.0, d=2.0):
        self.a, self.c, self.d = float(a), float(c), float(d)

    def params(self):
        return (self.a, self.c, self.d)

    def of_dot(self, G, nx=None, ny=None):
        base = self.a * G + self.c
        if self.d == 1.0:
            return base
        if float(self.d).is_integer() and 0 < self.d <= 16:
            out = base.copy()
            for _ in range(int(self.d) - 1):     # repeated product: exact op order the GPU uses
                out *= base
            return out
        return np.power(base, self.d)


class SigmoidKernel(Kernel):
    """k(x,y) = tanh(a * x'y + c)."""
    kernel_id = 2

    def __init__(self, a=1.0, c=0.0):
        self.a, self.c = float(a), float(c)

    def params(self):
        return (self.a, self.c, 0.0)

    def of_dot(self, G, nx=None, ny=None):
        return np.tanh(self.a * G + self.c)


def kernelmatrix(kernel, A, B=None, method="gemm"):
    """kernelmatrix(k, X) (kpca.jl:51) / kernelmatrix(k, Z, X) (kpca.jl:65).

    ``method='gemm'`` follows MLKernels' route (Gram by BLAS, then the scalar
    map with ||x||^2 + ||y||^2 - 2x'y for distance kernels); the symmetric
    case forces the diagonal of the squared distance to exactly 0.
    ``method='direct'`` evaluates sum((x-y)^2) explicitly (small inputs only)
    and serves as an independent check of the cancellation handling."""
    A = np.asarray(A, dtype=np.float64)
    sym = B is None
    Bm = A if sym else np.asarray(B, dtype=np.float64)
    if A.shape[1] != Bm.shape[1]:
        raise ValueError("dimension mismatch between A and B")
    if method == "direct" and isinstance(kernel, GaussianKernel):
        d2 = ((A[:, None, :] - Bm[None, :, :]) ** 2).sum(-1)
        return np.exp(-kernel.alpha * d2)
    G = A @ Bm.T
    nx = np.einsum("ij,ij->i", A, A)
    ny = nx if sym else np.einsum("ij,ij->i", Bm, Bm)
    if sym and isinstance(kernel, GaussianKernel):
        # exact zero distance on the diagonal, independent of summation order
        G = G.copy()
        G[np.arange(len(nx)), np.arange(len(nx))] = nx
    K = kernel.of_dot(G, nx, ny)
    if sym:
        K = 0.5 * (K + K.T)                      # the BLAS Gram is symmetric up to rounding
    return K


def centerkernelmatrix(K):
    """centerkernelmatrix!(K) (kpca.jl:52): K~ = K - 1 r' - r 1' + mu with r the
    column means and mu the grand mean.  Returns (K~, r, mu)."""
    K = np.asarray(K, dtype=np.float64)
    n = K.shape[0]
    r = K.sum(axis=0) / n
    mu = r.sum() / n
    Kc = K - r[None, :] - r[:, None] + mu
    return Kc, r, mu
