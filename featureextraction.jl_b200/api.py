"""Host-side mirror of the reference's function API for the KPCA / KLDA path
(/root/reference/src/FeatureExtraction.jl:7-16 export list): same names,
positional order and error behaviour, every call going through the C ABI of
include/fxb200.h into hand-written sm_100a CUDA.  No CPU arithmetic here."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import SolverOpts, Stats, check
from .kernels import Kernel

_default_ctx = None


class Context:
    """One GPU (one rank).  ``Context(device)`` or ``Context.distributed(...)``."""

    def __init__(self, device=0, _handle=None):
        self._lib = _lib.load()
        if _handle is not None:
            self._h = _handle
        else:
            h = C.c_void_p()
            check(self._lib.fx_init(int(device), C.byref(h)), "fx_init")
            self._h = h
        d, r, w, s = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._lib.fx_ctx_info(self._h, C.byref(d), C.byref(r), C.byref(w), C.byref(s))
        self.device, self.rank, self.world, self.sm_count = d.value, r.value, w.value, s.value

    @classmethod
    def distributed(cls, device, rank, world, unique_id):
        lib = _lib.load()
        h = C.c_void_p()
        buf = C.create_string_buffer(bytes(unique_id), 128)
        check(lib.fx_init_distributed(int(device), int(rank), int(world), buf, C.byref(h)),
              "fx_init_distributed")
        return cls(_handle=h)

    @classmethod
    def multi(cls, devices):
        """One process, several GPUs (fx_init_multi): K row-sharded over ``devices`` by the library
        itself, one host thread per device, NCCL and peer access set up internally."""
        lib = _lib.load()
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        check(lib.fx_init_multi(devs, len(devices), C.byref(h)), "fx_init_multi")
        return cls(_handle=h)

    @staticmethod
    def unique_id():
        lib = _lib.load()
        buf = C.create_string_buffer(128)
        check(lib.fx_comm_unique_id(buf), "fx_comm_unique_id")
        return bytes(buf.raw)

    @property
    def peer_access(self):
        """True if the last collective fit shared the symmetric blocks of K between ranks over
        NVLink peer mappings (fx_ctx_peer_access)."""
        return bool(self._lib.fx_ctx_peer_access(self._h))

    def set_stream(self, cuda_stream_ptr):
        check(self._lib.fx_set_stream(self._h, C.c_void_p(cuda_stream_ptr)), "fx_set_stream")

    def close(self):
        if getattr(self, "_h", None):
            self._lib.fx_shutdown(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


def _f64_colmajor(A, name):
    A = np.asarray(A)
    if A.ndim != 2:
        raise _lib.DimensionMismatch("%s must be a matrix" % name)
    if A.dtype != np.float64 or not A.flags.f_contiguous:
        A = np.asfortranarray(A, dtype=np.float64)
    return A


def _kparams(kernel):
    if not isinstance(kernel, Kernel) and not hasattr(kernel, "kernel_id"):
        raise TypeError("kernel must be a GaussianKernel / PolynomialKernel / SigmoidKernel")
    p = (C.c_double * 3)(*kernel.params())
    return int(kernel.kernel_id), p


def _out_dtype(*arrays):
    """The reference's functions are generic over T in (Float32, Float64) (test/runtests.jl:7).  This
    engine computes in float64 only: Float32 inputs are widened on the way in and results are
    narrowed on the way out, so callers get their element type back (like test_common.jl:66)."""
    for A in arrays:
        if isinstance(A, np.ndarray) and A.dtype == np.float32:
            return np.float32
    return np.float64


def _narrow(dt, *outs):
    if dt == np.float64:
        return outs if len(outs) > 1 else outs[0]
    conv = tuple(o.astype(dt) if isinstance(o, np.ndarray) and o.dtype == np.float64 else o for o in outs)
    return conv if len(conv) > 1 else conv[0]


def make_opts(**kw):
    o = SolverOpts()
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Model:
    """Fitted state in HBM (fx_model).  The reference has no model object -- its
    transform takes (X, kernel, W) again (kpca.jl:59-64); ``transform_kpca`` below
    keeps that signature on top of this handle."""

    def __init__(self, ctx, handle):
        self.ctx, self._h = ctx, handle
        n, d, k, kid = C.c_int64(), C.c_int64(), C.c_int(), C.c_int()
        ctx._lib.fx_model_info(handle, C.byref(n), C.byref(d), C.byref(k), C.byref(kid))
        self.n, self.d, self.k, self.kernel_id = n.value, d.value, k.value, kid.value

    def transform(self, Z, return_stats=False, _naive=False):
        Z = _f64_colmajor(Z, "Z")
        m, d = Z.shape
        if d != self.d:
            raise _lib.DimensionMismatch("Z has %d columns, the model was fitted on %d" % (d, self.d))
        out = np.empty((m, self.k), dtype=np.float64, order="F")
        st = Stats()
        fn = self.ctx._lib.fx_kpca_transform_naive_f64 if _naive else self.ctx._lib.fx_kpca_transform_f64
        check(fn(self.ctx._h, self._h, _ptr(Z), m, m, _ptr(out), m, C.byref(st)), "transform")
        return (out, st.as_dict()) if return_stats else out

    def free(self):
        if self._h:
            self.ctx._lib.fx_model_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def kpca(X, kernel, alpha=0.0, eps=0.0, *, k=None, center=True, ctx=None, opts=None,
         return_model=False, return_stats=False):
    """kpca(X, k, alpha, eps) -> (V, L)   [reference: src/kpca.jl:49-57]

    X is n x p with rows = observations.  Additions the reference lacks (SURVEY.md
    8b): ``k`` components (default: all n, feasible for n <= 1024 like the dense
    reference) and ``center`` (the reference always centres, kpca.jl:52).
    V has unit-norm columns (sign free), L is descending and already / (n-1)."""
    ctx = ctx or default_context()
    dt = _out_dtype(X)
    X = _f64_colmajor(X, "X")
    n, p = X.shape
    if k is None:
        k = n
    kid, kp = _kparams(kernel)
    V = np.empty((n, k), dtype=np.float64, order="F")
    L = np.empty(k, dtype=np.float64)
    st = Stats()
    mh = C.c_void_p()
    rc = ctx._lib.fx_kpca_fit_f64(ctx._h, _ptr(X), n, p, n, kid, kp, int(k), 1 if center else 0,
                                  float(alpha), float(eps), C.byref(opts) if opts is not None else None,
                                  _ptr(V), n, _ptr(L), C.byref(mh) if return_model else None,
                                  C.byref(st))
    check(rc, "kpca")
    out = list(_narrow(dt, V, L))
    if return_model:
        out.append(Model(ctx, mh))
    if return_stats:
        out.append(st.as_dict())
    return tuple(out)


def model_from_components(X, kernel, W, *, center=True, ctx=None):
    ctx = ctx or default_context()
    X = _f64_colmajor(X, "X")
    W = _f64_colmajor(W, "W")
    n, p = X.shape
    if W.shape[0] != n:
        raise _lib.DimensionMismatch("W must have as many rows as X")
    kid, kp = _kparams(kernel)
    mh = C.c_void_p()
    st = Stats()
    check(ctx._lib.fx_model_create_f64(ctx._h, _ptr(X), n, p, n, kid, kp, _ptr(W), W.shape[1], n,
                                       1 if center else 0, C.byref(mh), C.byref(st)), "model_create")
    return Model(ctx, mh)


def transform_kpca(X, kernel, W, Z, *, center=True, ctx=None):
    """transform_kpca(X, k, W, Z) -> m x k   [reference: src/kpca.jl:64-68, broken upstream]
    = (k(Z,X) - 1 r' - r_z 1' + mu) W, evaluated as k(Z,X) W - 1 (r'W)."""
    dt = _out_dtype(X, W, Z)
    m = model_from_components(X, kernel, W, center=center, ctx=ctx)
    try:
        return _narrow(dt, m.transform(Z))
    finally:
        m.free()


def transform_klda(X, kernel, W, Z, *, ctx=None):
    """transform_klda(X, k, W, Z) = kernelmatrix(k, Z, X) * W   [reference: src/klda.jl:71]"""
    return transform_kpca(X, kernel, W, Z, center=False, ctx=ctx)


def klda(X, kernel, y, nclass, frequency, alpha_b=0.0, eps_b=0.0, alpha_w=0.0, eps_w=0.0, *,
         ncomp=None, ctx=None, opts=None, return_model=False, return_stats=False):
    """Kernel LDA: the LDA pipeline of src/klda.jl:34-61 applied to K = k(X,X) with the pencil
    (Sigma_b, Sigma_w) (``klda!`` is exported at FeatureExtraction.jl:11 but undefined).
    y: 1-based labels (klda.jl:8).  eps_w < 0 means |eps_w| * tr(Sigma_w)/n."""
    ctx = ctx or default_context()
    X = _f64_colmajor(X, "X")
    n, p = X.shape
    y = np.ascontiguousarray(y, dtype=np.int64)
    if y.shape != (n,):
        raise _lib.DimensionMismatch("X and y must have the same number of rows.")   # klda.jl:16
    freq = np.ascontiguousarray(frequency, dtype=np.float64)
    if freq.shape != (nclass,):
        raise _lib.DimensionMismatch("frequency must have one entry per class")
    ncomp = (nclass - 1) if ncomp is None else int(ncomp)
    kid, kp = _kparams(kernel)
    A = np.empty((n, ncomp), dtype=np.float64, order="F")
    L = np.empty(ncomp, dtype=np.float64)
    st = Stats()
    mh = C.c_void_p()
    rc = ctx._lib.fx_klda_fit_f64(ctx._h, _ptr(X), n, p, n, kid, kp, _ptr(y), int(nclass), _ptr(freq),
                                  float(alpha_b), float(eps_b), float(alpha_w), float(eps_w), ncomp,
                                  C.byref(opts) if opts is not None else None, _ptr(A), n, _ptr(L),
                                  C.byref(mh) if return_model else None, C.byref(st))
    check(rc, "klda")
    out = [A, L]
    if return_model:
        out.append(Model(ctx, mh))
    if return_stats:
        out.append(st.as_dict())
    return tuple(out)


# ---- linear solvers of the export list (p <= 512) ----------------------------------------------
def pca(X, alpha=0.0, eps=0.0, *, ctx=None, return_mean=False):
    """pca(X, alpha, eps) -> (V, L)   [reference: src/kpca.jl:17-31]"""
    ctx = ctx or default_context()
    dt = _out_dtype(X)
    X = _f64_colmajor(X, "X")
    n, p = X.shape
    V = np.empty((p, p), dtype=np.float64, order="F")
    L = np.empty(p)
    mu = np.empty(p)
    st = Stats()
    check(ctx._lib.fx_pca_f64(ctx._h, _ptr(X), n, p, n, float(alpha), float(eps), _ptr(V), p, _ptr(L),
                              _ptr(mu), C.byref(st)), "pca")
    return _narrow(dt, V, L, mu) if return_mean else _narrow(dt, V, L)


def _project(W, Z, ctx, what):
    ctx = ctx or default_context()
    W = _f64_colmajor(W, "W"); Z = _f64_colmajor(Z, "Z")
    p, k = W.shape
    m = Z.shape[0]
    if Z.shape[1] != p:
        raise _lib.DimensionMismatch("%s: Z has %d columns, W has %d rows" % (what, Z.shape[1], p))
    out = np.empty((m, k), dtype=np.float64, order="F")
    st = Stats()
    check(ctx._lib.fx_project_f64(ctx._h, _ptr(W), p, k, p, _ptr(Z), m, m, _ptr(out), m, C.byref(st)), what)
    return out


def transform_pca(W, Z, *, ctx=None):
    """transform_pca(W, Z) = Z * W   [reference: src/kpca.jl:37]"""
    return _project(W, Z, ctx, "transform_pca")


def transform_lda(W, Z, *, ctx=None):
    """transform_lda(W, Z) = Z * W   [reference: src/klda.jl:65]"""
    return _project(W, Z, ctx, "transform_lda")


def lda(X, y, nclass, frequency, alpha_b=0.0, eps_b=0.0, alpha_w=0.0, eps_w=0.0, *, ctx=None):
    """lda!(X, y, k, frequency, ...) -> (V, L) out of place   [reference: src/klda.jl:57-61, with the
    pencil (Sigma_b, Sigma_w) instead of upstream's (Sigma_b, Sigma_b) at klda.jl:42]"""
    ctx = ctx or default_context()
    X = _f64_colmajor(X, "X")
    n, p = X.shape
    y = np.ascontiguousarray(y, dtype=np.int64)
    if y.shape != (n,):
        raise _lib.DimensionMismatch("X and y must have the same number of rows.")   # klda.jl:16
    freq = np.ascontiguousarray(frequency, dtype=np.float64)
    if freq.shape != (nclass,):
        raise _lib.DimensionMismatch("frequency must have one entry per class")
    V = np.empty((p, p), dtype=np.float64, order="F")
    L = np.empty(p)
    st = Stats()
    check(ctx._lib.fx_lda_f64(ctx._h, _ptr(X), n, p, n, _ptr(y), int(nclass), _ptr(freq), float(alpha_b),
                              float(eps_b), float(alpha_w), float(eps_w), _ptr(V), p, _ptr(L), C.byref(st)),
          "lda")
    return V, L


# ---- building blocks (parity tests, per-kernel benches) --------------------------------------
def kernelmatrix(kernel, X, Z=None, *, ctx=None, variant=0, return_stats=False):
    """kernelmatrix(k, X) -> n x n  /  kernelmatrix(k, Z, X) -> m x n (pass Z=...)  [kpca.jl:51,65]"""
    ctx = ctx or default_context()
    X = _f64_colmajor(X, "X")
    n, d = X.shape
    kid, kp = _kparams(kernel)
    st = Stats()
    if Z is None:
        K = np.empty((n, n), dtype=np.float64, order="F")
        check(ctx._lib.fx_gram_f64(ctx._h, None, n, n, _ptr(X), n, n, d, kid, kp, variant, _ptr(K), n,
                                   C.byref(st)), "kernelmatrix")
    else:
        Z = _f64_colmajor(Z, "Z")
        m = Z.shape[0]
        if Z.shape[1] != d:
            raise _lib.DimensionMismatch("Z and X must have the same number of columns")
        K = np.empty((m, n), dtype=np.float64, order="F")
        check(ctx._lib.fx_gram_f64(ctx._h, _ptr(Z), m, m, _ptr(X), n, n, d, kid, kp, variant, _ptr(K), m,
                                   C.byref(st)), "kernelmatrix")
    return (K, st.as_dict()) if return_stats else K


def centerkernelmatrix(K, *, ctx=None, return_stats=False):
    """centerkernelmatrix!(K) out of place -> (K~, colmean, grandmean)  [kpca.jl:52]"""
    ctx = ctx or default_context()
    K = np.array(K, dtype=np.float64, order="F", copy=True)
    if K.ndim != 2 or K.shape[0] != K.shape[1]:
        raise _lib.DimensionMismatch("Matrix K must be square.")
    n = K.shape[0]
    cm = np.empty(n)
    gm = np.empty(1)
    st = Stats()
    check(ctx._lib.fx_center_f64(ctx._h, _ptr(K), n, n, _ptr(cm), _ptr(gm), C.byref(st)), "center")
    return (K, cm, float(gm[0]), st.as_dict()) if return_stats else (K, cm, float(gm[0]))


def symm(K, V, *, ctx=None, variant=0):
    ctx = ctx or default_context()
    K = _f64_colmajor(K, "K"); V = _f64_colmajor(V, "V")
    n, b = V.shape
    Y = np.empty((n, b), dtype=np.float64, order="F")
    st = Stats()
    check(ctx._lib.fx_symm_f64(ctx._h, _ptr(K), n, n, _ptr(V), b, n, _ptr(Y), n, variant, C.byref(st)),
          "symm")
    return Y


def center_symm(K, V, *, ctx=None):
    """The centring pass as a fit runs it: returns (K~, K~ V, colmean, grandmean) from the fused
    centre-while-multiplying kernel (fx_center_symm_f64)."""
    ctx = ctx or default_context()
    K = np.array(K, dtype=np.float64, order="F", copy=True)
    V = _f64_colmajor(V, "V")
    if K.ndim != 2 or K.shape[0] != K.shape[1]:
        raise _lib.DimensionMismatch("Matrix K must be square.")
    n, b = V.shape
    if K.shape[0] != n:
        raise _lib.DimensionMismatch("K and V must have the same number of rows")
    Y = np.empty((n, b), dtype=np.float64, order="F")
    cm = np.empty(n)
    gm = np.empty(1)
    st = Stats()
    check(ctx._lib.fx_center_symm_f64(ctx._h, _ptr(K), n, n, _ptr(V), b, n, _ptr(Y), n, _ptr(cm), _ptr(gm),
                                      C.byref(st)), "center_symm")
    return K, Y, cm, float(gm[0])


def syevd_small(A, *, ctx=None):
    """Device Jacobi: returns (w descending, V, sweeps)."""
    ctx = ctx or default_context()
    A = np.array(A, dtype=np.float64, order="F", copy=True)
    if A.ndim != 2 or A.shape[0] != A.shape[1]:
        raise _lib.DimensionMismatch("Matrix A must be square.")   # common.jl:162
    m = A.shape[0]
    w = np.empty(m)
    sw = C.c_int()
    check(ctx._lib.fx_syevd_small_f64(ctx._h, _ptr(A), m, m, _ptr(w), C.byref(sw)), "syevd_small")
    return w, A, sw.value


def bench_kernel(which, n=0, d=0, b=0, m=0, kernel_id=0, reps=5, *, ctx=None):
    ctx = ctx or default_context()
    ms = np.zeros(reps)
    check(ctx._lib.fx_bench_kernel(ctx._h, which, n, d, b, m, kernel_id, reps, _ptr(ms)), "bench_kernel")
    return ms
