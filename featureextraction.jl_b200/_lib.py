"""ctypes binding of include/fxb200.h -- the executable stand-in for the Julia
``ccall`` wrapper (julia/FeatureExtractionB200.jl; Julia is not in this image).

There is deliberately NO fallback: if ``csrc/libfxb200.so`` is missing or does
not load, importing the product API raises immediately.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# FXB200_LIB_PATH: load an experimental build instead (kernel A/B measurements; tools/probes/)
LIB_PATH = os.environ.get("FXB200_LIB_PATH") or os.path.join(_HERE, "csrc", "libfxb200.so")


class SolverOpts(C.Structure):
    _fields_ = [("block", C.c_int32), ("max_blocks", C.c_int32), ("max_steps", C.c_int32),
                ("variant", C.c_int32), ("tol", C.c_double), ("seed", C.c_uint64),
                ("cg_max_iters", C.c_int32), ("klda_path", C.c_int32), ("cg_tol", C.c_double)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("total_ms", "h2d_ms", "d2h_ms", "gram_ms", "center_ms", "symm_ms", "ortho_ms",
                 "rr_ms", "comm_ms", "eig_ms", "transform_ms", "max_residual", "trace")] + \
               [(n, C.c_int64) for n in
                ("gram_launches", "center_launches", "symm_launches", "transform_launches",
                 "kernel_launches", "h2d_bytes", "d2h_bytes")] + \
               [(n, C.c_int32) for n in
                ("steps", "restarts", "converged", "block", "cg_iters", "dense_path")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


_P = C.c_void_p
_D = C.POINTER(C.c_double)
_I64 = C.c_int64

# name -> (restype, argtypes); every symbol include/fxb200.h declares
PROTOTYPES = {
    "fx_version": (C.c_int, []),
    "fx_last_error": (C.c_char_p, []),
    "fx_init": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "fx_comm_unique_id": (C.c_int, [_P]),
    "fx_init_distributed": (C.c_int, [C.c_int, C.c_int, C.c_int, _P, C.POINTER(_P)]),
    "fx_init_multi": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(_P)]),
    "fx_set_stream": (C.c_int, [_P, _P]),
    "fx_ctx_info": (C.c_int, [_P] + [C.POINTER(C.c_int)] * 4),
    "fx_ctx_peer_access": (C.c_int, [_P]),
    "fx_shutdown": (None, [_P]),
    "fx_kpca_fit_f64": (C.c_int, [_P, _P, _I64, _I64, _I64, C.c_int, _D, C.c_int, C.c_int,
                                  C.c_double, C.c_double, C.POINTER(SolverOpts), _P, _I64, _P,
                                  C.POINTER(_P), C.POINTER(Stats)]),
    "fx_kpca_fit_dev_f64": (C.c_int, [_P, _P, _I64, _I64, _I64, C.c_int, _D, C.c_int, C.c_int,
                                      C.c_double, C.c_double, C.POINTER(SolverOpts), _P, _I64, _P,
                                      C.POINTER(_P), C.POINTER(Stats)]),
    "fx_kpca_transform_f64": (C.c_int, [_P, _P, _P, _I64, _I64, _P, _I64, C.POINTER(Stats)]),
    "fx_kpca_transform_dev_f64": (C.c_int, [_P, _P, _P, _I64, _I64, _P, _I64, C.POINTER(Stats)]),
    "fx_kpca_transform_naive_f64": (C.c_int, [_P, _P, _P, _I64, _I64, _P, _I64, C.POINTER(Stats)]),
    "fx_model_create_f64": (C.c_int, [_P, _P, _I64, _I64, _I64, C.c_int, _D, _P, C.c_int, _I64,
                                      C.c_int, C.POINTER(_P), C.POINTER(Stats)]),
    "fx_model_info": (C.c_int, [_P, C.POINTER(_I64), C.POINTER(_I64), C.POINTER(C.c_int),
                                C.POINTER(C.c_int)]),
    "fx_model_free": (None, [_P]),
    "fx_klda_fit_f64": (C.c_int, [_P, _P, _I64, _I64, _I64, C.c_int, _D, _P, C.c_int, _P,
                                  C.c_double, C.c_double, C.c_double, C.c_double, C.c_int,
                                  C.POINTER(SolverOpts), _P, _I64, _P, C.POINTER(_P),
                                  C.POINTER(Stats)]),
    "fx_klda_transform_f64": (C.c_int, [_P, _P, _P, _I64, _I64, _P, _I64, C.POINTER(Stats)]),
    "fx_pca_f64": (C.c_int, [_P, _P, _I64, _I64, _I64, C.c_double, C.c_double, _P, _I64, _P, _P,
                             C.POINTER(Stats)]),
    "fx_lda_f64": (C.c_int, [_P, _P, _I64, _I64, _I64, _P, C.c_int, _P, C.c_double, C.c_double,
                             C.c_double, C.c_double, _P, _I64, _P, C.POINTER(Stats)]),
    "fx_project_f64": (C.c_int, [_P, _P, _I64, C.c_int, _I64, _P, _I64, _I64, _P, _I64,
                                 C.POINTER(Stats)]),
    "fx_gram_f64": (C.c_int, [_P, _P, _I64, _I64, _P, _I64, _I64, _I64, C.c_int, _D, C.c_int,
                              _P, _I64, C.POINTER(Stats)]),
    "fx_center_f64": (C.c_int, [_P, _P, _I64, _I64, _P, _P, C.POINTER(Stats)]),
    "fx_symm_f64": (C.c_int, [_P, _P, _I64, _I64, _P, C.c_int, _I64, _P, _I64, C.c_int,
                              C.POINTER(Stats)]),
    "fx_center_symm_f64": (C.c_int, [_P, _P, _I64, _I64, _P, C.c_int, _I64, _P, _I64, _P, _P,
                                     C.POINTER(Stats)]),
    "fx_syevd_small_f64": (C.c_int, [_P, _P, C.c_int, _I64, _P, C.POINTER(C.c_int)]),
    "fx_bench_kernel": (C.c_int, [_P, C.c_int, _I64, _I64, C.c_int, _I64, C.c_int, C.c_int, _P]),
}

_lib = None


class FxLibraryMissing(RuntimeError):
    pass


def load():
    """Loads libfxb200.so (once).  Raises FxLibraryMissing -- never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FxLibraryMissing(
            "CUDA library %s is missing: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `make -C featureextraction.jl_b200/csrc`.  There is no CPU fallback."
            % LIB_PATH)
    try:
        lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    except OSError as e:
        raise FxLibraryMissing("cannot load %s: %s (no CPU fallback)" % (LIB_PATH, e))
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)            # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class FxError(RuntimeError):
    """info > 0 from the library (runtime class), like LAPACKException (common.jl:90-92)."""
    def __init__(self, code, msg):
        super().__init__("fxb200 error %d: %s" % (code, msg))
        self.code = code


class DimensionMismatch(ValueError):
    """info < 0 for a shape argument, like Julia's DimensionMismatch (common.jl:16,24,33,42)."""


def check(rc, what=""):
    if rc == 0:
        return
    msg = (load().fx_last_error() or b"").decode("utf-8", "replace")
    if rc < 0:
        if "DimensionMismatch" in msg:
            raise DimensionMismatch("%s: %s" % (what, msg))
        raise ValueError("%s: invalid argument #%d: %s" % (what, -rc, msg))   # common.jl:87-89
    raise FxError(rc, "%s: %s" % (what, msg))
