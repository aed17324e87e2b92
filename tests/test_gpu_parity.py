"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded
inputs, against the committed golden vectors, and -- at larger sizes -- through
size-independent properties.  Tolerances (BASELINE.json north_star): eigenvalues 1e-10
relative, subspace angle below 1e-8; kernel matrices 1e-13 absolute / relative."""
import glob
import os

import numpy as np
import pytest

import oracle as O
import fxb200
from fxb200.synth import mix

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))

KERNELS = [
    ("gaussian", fxb200.GaussianKernel(0.05), O.GaussianKernel(0.05)),
    ("poly3", fxb200.PolynomialKernel(0.05, 1.0, 3.0), O.PolynomialKernel(0.05, 1.0, 3.0)),
    ("poly2.5", fxb200.PolynomialKernel(0.01, 2.0, 2.5), O.PolynomialKernel(0.01, 2.0, 2.5)),
    ("linear", fxb200.PolynomialKernel(1.0, 0.0, 1.0), O.PolynomialKernel(1.0, 0.0, 1.0)),
    ("sigmoid", fxb200.SigmoidKernel(0.02, 0.3), O.SigmoidKernel(0.02, 0.3)),
]


def sin_angle(A, B):
    Qa, _ = np.linalg.qr(A)
    Qb, _ = np.linalg.qr(B)
    return np.linalg.norm(Qb - Qa @ (Qa.T @ Qb), 2)


def relclose(a, b, tol):
    scale = max(np.abs(b).max(), 1e-300)
    err = np.abs(a - b).max() / scale
    assert err <= tol, "rel err %.3e > %.1e" % (err, tol)


# ------------------------------------------------------------------ Gram + kernel map epilogue
@pytest.mark.parametrize("name,kern,okern", KERNELS, ids=[k[0] for k in KERNELS])
@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("n,d", [(300, 10), (1000, 37), (128, 16)])
def test_gram_symmetric(gpu_ctx, name, kern, okern, variant, n, d):
    X, _ = mix(n, d, 4, 2.0, 31)
    K = fxb200.kernelmatrix(kern, X, ctx=gpu_ctx, variant=variant)
    relclose(K, O.kernelmatrix(okern, X), 1e-13)
    if name == "gaussian":
        assert np.all(np.diag(K) == 1.0)          # exact zero distance on the diagonal


@pytest.mark.parametrize("name,kern,okern", KERNELS[:3], ids=[k[0] for k in KERNELS[:3]])
def test_gram_cross(gpu_ctx, name, kern, okern):
    X, _ = mix(777, 21, 4, 2.0, 32)
    Z, _ = mix(130, 21, 4, 2.0, 33)
    relclose(fxb200.kernelmatrix(kern, X, Z, ctx=gpu_ctx), O.kernelmatrix(okern, Z, X), 1e-13)


# ------------------------------------------------------------------ centring
@pytest.mark.parametrize("n", [64, 301, 1000])
def test_center(gpu_ctx, n):
    X, _ = mix(n, 9, 3, 2.0, 34)
    K = O.kernelmatrix(O.GaussianKernel(0.1), X)
    Kc, cm, gm = fxb200.centerkernelmatrix(K, ctx=gpu_ctx)
    Ko, r, mu = O.centerkernelmatrix(K)
    relclose(Kc, Ko, 1e-14)
    relclose(cm, r, 1e-14)
    assert abs(gm - mu) <= 1e-14 * abs(mu)
    assert np.abs(Kc.sum(0)).max() < 1e-10       # double centring: zero column sums


# ------------------------------------------------------------------ K*V (SYMM)
@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("n,b", [(256, 8), (1000, 16), (1300, 24), (700, 40), (513, 64), (200, 5)])
def test_symm(gpu_ctx, variant, n, b):
    rng = np.random.default_rng(35)
    A = rng.standard_normal((n, n)); K = A + A.T
    V = rng.standard_normal((n, b))
    relclose(fxb200.symm(K, V, ctx=gpu_ctx, variant=variant), K @ V, 1e-13)


# ------------------------------------------------------------------ small dense eigensolver
@pytest.mark.parametrize("m", [1, 2, 5, 16, 33, 96, 160, 161, 300])
def test_syevd_small(gpu_ctx, m):
    rng = np.random.default_rng(36)
    A = rng.standard_normal((m, m)); A = A + A.T
    w, V, sweeps = fxb200.syevd_small(A, ctx=gpu_ctx)
    wo = np.linalg.eigvalsh(A)[::-1]
    scale = max(np.abs(wo).max(), 1e-300)
    assert np.abs(w - wo).max() <= 1e-13 * scale
    assert np.abs(V.T @ V - np.eye(m)).max() < 1e-12
    assert np.abs(A @ V - V * w[None, :]).max() <= 1e-12 * scale
    assert sweeps <= 20


# ------------------------------------------------------------------ KPCA fit
def check_fit(V, L, Vo, Lo, k):
    np.testing.assert_allclose(L, Lo[:k], rtol=1e-10)
    assert sin_angle(V, Vo[:, :k]) < 1e-8
    np.testing.assert_allclose(np.linalg.norm(V, axis=0), 1.0, atol=1e-12)


def test_reference_identity_kpca_vs_pca(gpu_ctx):
    """test_kpca.jl:44-56 through the CUDA path: KPCA with PolynomialKernel(1,0,1) == PCA
    (all n eigenpairs: the small-n dense path)."""
    n, p = 20, 5
    X = np.random.default_rng(7).random((n, p))
    Xc = X - X.mean(0)
    S = Xc.T @ Xc / (n - 1)
    for eps in (0.0, 1.0):
        V1, D1 = fxb200.kpca(X, fxb200.PolynomialKernel(1.0, 0.0, 1.0), 0.0, eps, ctx=gpu_ctx)
        V2, D2 = O.pca(X.copy(), 0.0, eps)
        a = (V1.T @ Xc @ Xc.T @ V1 / (n - 1))[:p, :p]
        b = V2.T @ S @ V2
        assert np.abs(a - b).max() < 1e-12
        np.testing.assert_allclose(np.abs(D1)[:p], np.abs(D2), rtol=1e-12)
        assert V1.shape == (n, n) and np.all(np.diff(D1) <= 1e-15)


@pytest.mark.parametrize("name,kern,okern", [KERNELS[0], KERNELS[1], KERNELS[4]],
                         ids=["gaussian", "poly3", "sigmoid"])
def test_kpca_fit_matches_oracle(gpu_ctx, name, kern, okern):
    X, _ = mix(1500, 20, 5, 3.0, 41)
    k = 4
    V, L, st = fxb200.kpca(X, kern, k=k, ctx=gpu_ctx, return_stats=True)
    Vo, Lo = O.kpca(X, okern)
    check_fit(V, L, Vo, Lo, k)
    assert st["converged"] == 1 and st["dense_path"] == 0 and st["kernel_launches"] > 10


def test_kpca_config_C1(gpu_ctx):
    """BASELINE.json configs[0]: RBF n=1000 d=10 k=2, fit + transform of 200 fresh points."""
    cfg = fxb200.synth.CONFIGS["C1"]
    X, _ = mix(cfg["n"], cfg["d"], cfg["c"], cfg["sep"], cfg["seed"])
    Z, _ = mix(200, cfg["d"], cfg["c"], cfg["sep"], 1002)
    kern, okern = fxb200.GaussianKernel(0.1), O.GaussianKernel(0.1)
    V, L, model = fxb200.kpca(X, kern, k=2, ctx=gpu_ctx, return_model=True)
    Vo, Lo = O.kpca(X, okern)
    check_fit(V, L, Vo, Lo, 2)
    T = model.transform(Z)
    relclose(T, O.transform_kpca(X, okern, V, Z), 1e-10)
    relclose(model.transform(Z, _naive=True), T, 1e-11)
    relclose(fxb200.transform_kpca(X, kern, V, Z, ctx=gpu_ctx), T, 1e-10)   # stateless signature
    relclose(model.transform(X), V * (L * (len(X) - 1))[None, :], 1e-9)     # in-sample identity
    model.free()


def test_kpca_naive_variant_agrees(gpu_ctx):
    X, _ = mix(900, 12, 4, 3.0, 42)
    kern = fxb200.GaussianKernel(1 / 12)
    V0, L0 = fxb200.kpca(X, kern, k=3, ctx=gpu_ctx)
    V1, L1 = fxb200.kpca(X, kern, k=3, ctx=gpu_ctx, opts=fxb200.make_opts(variant=1))
    np.testing.assert_allclose(L0, L1, rtol=1e-12)
    assert sin_angle(V0, V1) < 1e-9


def test_kpca_alpha_eps_center(gpu_ctx):
    X, _ = mix(700, 8, 4, 3.0, 43)
    kern, okern = fxb200.GaussianKernel(0.125), O.GaussianKernel(0.125)
    for alpha, eps in [(0.3, 0.0), (0.0, 0.5), (0.5, 1.0)]:
        V, L = fxb200.kpca(X, kern, alpha, eps, k=3, ctx=gpu_ctx)
        Vo, Lo = O.kpca(X, okern, alpha, eps)
        check_fit(V, L, Vo, Lo, 3)
    V, L = fxb200.kpca(X, kern, k=3, center=False, ctx=gpu_ctx)
    Vo, Lo = O.kpca(X, okern, center=False)
    check_fit(V, L, Vo, Lo, 3)
    # alpha = 1: K' = beta I, every eigenvalue equals tr(K~)/n/(n-1) (degenerate basis)
    _V, L = fxb200.kpca(X, kern, 1.0, 0.0, k=3, ctx=gpu_ctx)
    _Vo, Lo = O.kpca(X, okern, 1.0, 0.0)
    np.testing.assert_allclose(L, Lo[:3], rtol=1e-10)


def test_kpca_low_rank_kernel(gpu_ctx):
    """Linear kernel with d < block width: K~ has rank d, the Krylov block goes rank deficient."""
    X, _ = mix(2000, 5, 3, 3.0, 44)
    V, L = fxb200.kpca(X, fxb200.PolynomialKernel(1.0, 0.0, 1.0), k=5, ctx=gpu_ctx)
    Vo, Lo = O.kpca(X, O.PolynomialKernel(1.0, 0.0, 1.0))
    check_fit(V, L, Vo, Lo, 5)
    _V2, D2 = O.pca(X)
    np.testing.assert_allclose(L, D2, rtol=1e-10)      # == PCA eigenvalues (test_kpca.jl:55)


def test_kpca_slow_decay_restarts(gpu_ctx):
    rng = np.random.default_rng(45)
    X = rng.standard_normal((1400, 16))
    kern, okern = fxb200.GaussianKernel(1 / 16), O.GaussianKernel(1 / 16)
    V, L, st = fxb200.kpca(X, kern, k=6, ctx=gpu_ctx, return_stats=True,
                           opts=fxb200.make_opts(max_blocks=2))
    Vo, Lo = O.kpca(X, okern)
    np.testing.assert_allclose(L, Lo[:6], rtol=1e-10)
    assert st["restarts"] >= 1 and st["converged"] == 1


def test_kpca_medium(gpu_ctx):
    X, _ = mix(4096, 128, 9, 3.0, 46)
    kern, okern = fxb200.GaussianKernel(1 / 128), O.GaussianKernel(1 / 128)
    V, L, model = fxb200.kpca(X, kern, k=8, ctx=gpu_ctx, return_model=True)
    Vo, Lo = O.kpca(X, okern)
    check_fit(V, L, Vo, Lo, 8)
    Z, _ = mix(1000, 128, 9, 3.0, 47)
    relclose(model.transform(Z), O.transform_kpca(X, okern, V, Z), 1e-10)
    model.free()


# ------------------------------------------------------------------ golden vectors
GOLD = sorted(glob.glob(os.path.join(HERE, "golden", "kpca_*.npz")))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(f) for f in GOLD])
def test_golden(gpu_ctx, path):
    g = np.load(path)
    n, d, c, sep, seed, k, m, zseed = g["meta"]
    X, _ = mix(int(n), int(d), int(c), float(sep), int(seed))
    Z, _ = mix(int(m), int(d), int(c), float(sep), int(zseed))
    kern = fxb200.from_spec(str(g["kernel"]), tuple(g["params"]))
    K = fxb200.kernelmatrix(kern, X, ctx=gpu_ctx)
    relclose(K[g["rows"]], g["K_rows"], 1e-13)
    Kc, cm, gm = fxb200.centerkernelmatrix(K, ctx=gpu_ctx)
    relclose(Kc[g["rows"]], g["Kc_rows"], 1e-13)
    relclose(cm, g["colmean"], 1e-13)
    V, L, model = fxb200.kpca(X, kern, k=int(k), ctx=gpu_ctx, return_model=True)
    np.testing.assert_allclose(L, g["L"], rtol=1e-10)
    assert np.abs(np.abs(V) - g["V_abs"]).max() < 1e-8
    relclose(np.abs(model.transform(Z)), g["T_abs"], 1e-8)
    model.free()


# ------------------------------------------------------------------ size-independent properties
def test_properties_large(gpu_ctx):
    """n = 16,384 (BASELINE configs[1]): the dense oracle needs minutes here, so check the
    eigen-relation through the library itself: transform(X) = K~ V = V diag(lambda (n-1)),
    V'V = I, 1'V = 0, descending eigenvalues, trace bound."""
    cfg = fxb200.synth.CONFIGS["C2"]
    X, _ = mix(cfg["n"], cfg["d"], cfg["c"], cfg["sep"], cfg["seed"])
    n = len(X)
    kern = fxb200.GaussianKernel(cfg["params"][0])
    V, L, model, st = fxb200.kpca(X, kern, k=cfg["k"], ctx=gpu_ctx, return_model=True,
                                  return_stats=True)
    assert st["converged"] == 1
    assert np.abs(V.T @ V - np.eye(cfg["k"])).max() < 1e-12
    assert np.abs(V.sum(0)).max() < 1e-9
    assert np.all(np.diff(L) < 0) and L[-1] > 0
    lam = L * (n - 1)
    assert lam.sum() <= st["trace"] * (1 + 1e-12)
    R = model.transform(X) - V * lam[None, :]
    assert np.abs(R).max() <= 1e-9 * lam[0]
    model.free()


# ------------------------------------------------------------------ error behaviour
def test_errors(gpu_ctx):
    X = np.random.default_rng(1).random((50, 3))
    with pytest.raises(ValueError):
        fxb200.kpca(X, fxb200.GaussianKernel(1.0), k=51, ctx=gpu_ctx)
    with pytest.raises(ValueError):
        fxb200.kpca(X, fxb200.GaussianKernel(1.0), 1.5, 0.0, k=2, ctx=gpu_ctx)
    with pytest.raises(fxb200.DimensionMismatch):
        fxb200.centerkernelmatrix(np.zeros((3, 4)), ctx=gpu_ctx)
    with pytest.raises(fxb200.DimensionMismatch):
        fxb200.syevd_small(np.zeros((3, 4)), ctx=gpu_ctx)
    V, L, model = fxb200.kpca(X, fxb200.GaussianKernel(1.0), k=2, ctx=gpu_ctx, return_model=True)
    with pytest.raises(fxb200.DimensionMismatch):
        model.transform(np.zeros((4, 5)))
    with pytest.raises(fxb200.DimensionMismatch):
        fxb200.transform_kpca(X, fxb200.GaussianKernel(1.0), np.zeros((49, 2)), X, ctx=gpu_ctx)
    model.free()


# ------------------------------------------------------------------ KLDA (parity unpinned upstream)
def _klda_case(n, d, c, sep, seed, gamma):
    X, y0 = mix(n, d, c, sep, seed)
    y = y0 + 1
    freq = np.bincount(y0, minlength=c) / n
    return X, y, freq


@pytest.mark.parametrize("n,d,c,rel", [(600, 8, 3, 1e-3), (1500, 32, 5, 1e-6), (900, 16, 10, 1e-6),
                                       (2000, 64, 10, 1e-6)])
def test_klda_matches_oracle(gpu_ctx, n, d, c, rel):
    X, y, freq = _klda_case(n, d, c, 2.0, 71, 1.0 / d)
    kern, okern = fxb200.GaussianKernel(1.0 / d), O.GaussianKernel(1.0 / d)
    K = O.kernelmatrix(okern, X)
    H_b, H_w, _ = O.lda_matrices(K, y, c, freq)
    _Sb, Sw = O.scatter_matrices(H_b, 0, 0, H_w, 0, 0)
    eps_w = rel * np.trace(Sw) / n
    Ao, Lo = O.klda(X, okern, y, c, freq, eps_w=eps_w)
    A, L, model, st = fxb200.klda(X, kern, y, c, freq, eps_w=eps_w, ctx=gpu_ctx, return_model=True,
                                  return_stats=True, opts=fxb200.make_opts(klda_path=1))   # iterative path
    assert st["converged"] == 1 and st["cg_iters"] > 0
    np.testing.assert_allclose(L, Lo, rtol=1e-7)
    # vectors: compare the spanned subspace (clustered eigenvalues) and the B-normalisation
    assert sin_angle(A, Ao) < 1e-5
    Swr = Sw + eps_w * np.eye(n)
    np.testing.assert_allclose(np.diag(A.T @ Swr @ A), 1.0, rtol=1e-6)
    # relative eps_w (negative) resolves to the same problem
    A2, L2 = fxb200.klda(X, kern, y, c, freq, eps_w=-rel, ctx=gpu_ctx, opts=fxb200.make_opts(klda_path=1))
    np.testing.assert_allclose(L2, Lo, rtol=1e-6)
    # transform_klda = k(Z,X) A, no centring (klda.jl:71)
    Z, _ = mix(77, d, c, 2.0, 72)
    T = model.transform(Z)
    To = O.transform_klda(X, okern, A, Z)
    relclose(T, To, 1e-10)
    relclose(fxb200.transform_klda(X, kern, A, Z, ctx=gpu_ctx), To, 1e-10)
    model.free()


def test_klda_alpha_w_and_errors(gpu_ctx):
    n, d, c = 500, 6, 4
    X, y, freq = _klda_case(n, d, c, 2.0, 73, 1.0 / d)
    kern, okern = fxb200.GaussianKernel(1.0 / d), O.GaussianKernel(1.0 / d)
    Ao, Lo = O.klda(X, okern, y, c, freq, alpha_w=0.05)
    A, L = fxb200.klda(X, kern, y, c, freq, alpha_w=0.05, ctx=gpu_ctx, opts=fxb200.make_opts(klda_path=1))
    np.testing.assert_allclose(L, Lo, rtol=1e-7)
    ybad = y.copy(); ybad[3] = c + 1
    with pytest.raises(ValueError):                     # klda.jl:8 "Index out of range"
        fxb200.klda(X, kern, ybad, c, freq, eps_w=1e-3, ctx=gpu_ctx)
    with pytest.raises(fxb200.FxError):                 # Sigma_w singular without regularisation
        fxb200.klda(X, kern, y, c, freq, ctx=gpu_ctx)
    with pytest.raises(fxb200.DimensionMismatch):       # klda.jl:16
        fxb200.klda(X, kern, y[:-1], c, freq, eps_w=1e-3, ctx=gpu_ctx)
    # an ill-conditioned pencil that CG cannot finish in the allowed iterations reports
    # "not converged" (the analogue of LAPACKException, common.jl:90-92) instead of a wrong answer
    with pytest.raises(fxb200.FxError) as ei:
        fxb200.klda(X, kern, y, c, freq, eps_w=-1e-9, ctx=gpu_ctx, opts=fxb200.make_opts(cg_max_iters=20, klda_path=1))
    assert ei.value.code == 4
    # ... while the dense path (the default at this size) answers the same ill-conditioned pencil like sygvd
    H_b, H_w, _ = O.lda_matrices(O.kernelmatrix(okern, X), y, c, freq)
    tr_w = np.trace(O.scatter_matrices(H_b, 0, 0, H_w, 0, 0)[1])
    Ao, Lo = O.klda(X, okern, y, c, freq, eps_w=1e-9 * tr_w / n)
    A, L, st = fxb200.klda(X, kern, y, c, freq, eps_w=-1e-9, ctx=gpu_ctx, return_stats=True)
    assert st["dense_path"] == 1
    np.testing.assert_allclose(L, Lo, rtol=1e-4)


# ------------------------------------------------------------------ edge cases
@pytest.mark.parametrize("n,d,k", [(300, 1, 1), (257, 3, 1), (1025, 1000, 3), (640, 17, 64), (129, 5, 129)])
def test_kpca_shapes(gpu_ctx, n, d, k):
    """k = 1, d = 1, d not a multiple of 16, block width 64, n just above the dense threshold,
    and the all-components request of the reference (k = n, dense path)."""
    X, _ = mix(n, d, 4, 3.0, 81)
    gamma = 1.0 / d
    kern, okern = fxb200.GaussianKernel(gamma), O.GaussianKernel(gamma)
    V, L, st = fxb200.kpca(X, kern, k=k, ctx=gpu_ctx, return_stats=True)
    Vo, Lo = O.kpca(X, okern)
    scale = Lo[0]
    assert np.abs(L - Lo[:k]).max() <= 1e-10 * scale
    assert np.abs(V.T @ V - np.eye(k)).max() < 1e-10
    kk = min(k, 3)                      # leading, well separated part
    assert sin_angle(V[:, :kk], Vo[:, :kk]) < 1e-7
    assert st["dense_path"] == (1 if (n <= 256 or k > 64) else 0)


@pytest.mark.parametrize("n,d,k,gs", [(4096, 4, 256, 0.25), (2048, 4, 200, 0.25), (1500, 8, 70, 0.1)])
def test_kpca_more_than_64_components(gpu_ctx, n, d, k, gs):
    """k > 64 beyond the dense-Jacobi sizes: groups of 48 with locked (deflated) eigenvectors.  All k
    eigenvalues against dsyevd (the reference returns every pair, common.jl:164), orthonormality, the
    eigen-residual of every returned pair (1e-9 |lambda_1| is this mode's documented Ritz-residual
    tolerance), and the subspace of the well separated leading part."""
    X, _ = mix(n, d, 9, 3.0, 92)
    kern, okern = fxb200.GaussianKernel(gs / d), O.GaussianKernel(gs / d)
    V, L, model, st = fxb200.kpca(X, kern, k=k, ctx=gpu_ctx, return_model=True, return_stats=True)
    Vo, Lo = O.kpca(X, okern)
    assert st["converged"] == 1 and st["dense_path"] == 0
    assert np.abs(L - Lo[:k]).max() <= 1e-10 * Lo[0]
    assert np.all(np.diff(L) <= 1e-12 * Lo[0])
    assert np.abs(V.T @ V - np.eye(k)).max() < 1e-10
    Kc, _, _ = O.centerkernelmatrix(O.kernelmatrix(okern, X))
    R = Kc @ V - V * (L * (n - 1))[None, :]
    assert np.linalg.norm(R, axis=0).max() <= 2e-9 * Lo[0] * (n - 1)
    assert sin_angle(V[:, :8], Vo[:, :8]) < 1e-7
    Z, _ = mix(50, d, 9, 3.0, 93)
    relclose(model.transform(Z), O.transform_kpca(X, okern, V, Z), 1e-9)        # k > 64 through the model
    model.free()


@pytest.mark.parametrize("n,d,k", [(2048, 16, None), (1300, 5, 1300), (2500, 32, 600), (4096, 32, None)])
def test_kpca_full_spectrum_dense_path(gpu_ctx, n, d, k):
    """The reference's bare call kpca(X, kernel) returns ALL n pairs (common.jl:164).  For
    1024 < n <= 8192 that is served by the one-sided (Hestenes) Jacobi solver over all SMs: every
    eigenvalue against dsyevd, orthonormal V, K~ V = V diag(lambda)."""
    X, _ = mix(n, d, 6, 3.0, 94)
    kern, okern = fxb200.GaussianKernel(1.0 / d), O.GaussianKernel(1.0 / d)
    V, L, st = fxb200.kpca(X, kern, k=k, ctx=gpu_ctx, return_stats=True)
    kk = n if k is None else k
    Vo, Lo = O.kpca(X, okern)
    assert st["dense_path"] == 1 and st["converged"] == 1 and V.shape == (n, kk)
    assert np.abs(L - Lo[:kk]).max() <= 1e-10 * Lo[0]
    assert np.abs(V.T @ V - np.eye(kk)).max() < 1e-10
    Kc, _, _ = O.centerkernelmatrix(O.kernelmatrix(okern, X))
    R = Kc @ V - V * (L * (n - 1))[None, :]
    assert np.linalg.norm(R, axis=0).max() <= 1e-10 * Lo[0] * (n - 1)
    assert sin_angle(V[:, :5], Vo[:, :5]) < 1e-8


def test_kpca_degenerate_inputs(gpu_ctx):
    """Duplicate observations (rank-deficient K) and constant data (K~ = 0): eigenvalues must
    match the dense reference path; any orthonormal basis is acceptable for a zero spectrum."""
    rng = np.random.default_rng(82)
    base = rng.standard_normal((50, 6))
    X = np.repeat(base, 8, axis=0)                      # 400 rows, 50 distinct
    kern, okern = fxb200.GaussianKernel(0.2), O.GaussianKernel(0.2)
    V, L = fxb200.kpca(X, kern, k=5, ctx=gpu_ctx)
    Vo, Lo = O.kpca(X, okern)
    np.testing.assert_allclose(L, Lo[:5], rtol=1e-10)
    Xc = np.ones((300, 4))
    V, L = fxb200.kpca(Xc, kern, k=2, ctx=gpu_ctx)
    assert np.all(np.isfinite(V)) and np.all(np.isfinite(L))
    assert np.abs(L).max() < 1e-12


def test_sigmoid_indefinite_spectrum(gpu_ctx):
    """tanh kernels are not PSD: the leading (largest algebraic) pairs must still match."""
    X, _ = mix(900, 10, 4, 2.0, 83)
    kern, okern = fxb200.SigmoidKernel(0.5, -1.0), O.SigmoidKernel(0.5, -1.0)
    V, L = fxb200.kpca(X, kern, k=4, ctx=gpu_ctx)
    Vo, Lo = O.kpca(X, okern)
    assert Lo[-1] < 0                                   # indefinite indeed
    np.testing.assert_allclose(L, Lo[:4], rtol=1e-10)
    assert sin_angle(V, Vo[:, :4]) < 1e-8


@pytest.mark.parametrize("m", [1, 127, 129, 300])
def test_transform_ragged_query_counts(gpu_ctx, m):
    X, _ = mix(500, 7, 3, 3.0, 84)
    Z, _ = mix(m, 7, 3, 3.0, 85)
    for kern, okern in [(fxb200.GaussianKernel(0.1), O.GaussianKernel(0.1)),
                        (fxb200.PolynomialKernel(0.1, 1.0, 2.0), O.PolynomialKernel(0.1, 1.0, 2.0))]:
        V, L, model = fxb200.kpca(X, kern, k=3, ctx=gpu_ctx, return_model=True)
        relclose(model.transform(Z), O.transform_kpca(X, okern, V, Z), 1e-10)
        model.free()
    # uncentred fit: transform is the plain projection k(Z,X) V
    V, L, model = fxb200.kpca(X, fxb200.GaussianKernel(0.1), k=3, center=False, ctx=gpu_ctx, return_model=True)
    relclose(model.transform(Z), O.kernelmatrix(O.GaussianKernel(0.1), Z, X) @ V, 1e-10)
    model.free()


def test_contexts_are_independent(fx):
    """Two contexts on the same GPU give identical results (the library keeps no global state
    beyond its function-attribute caches)."""
    X, _ = mix(700, 9, 3, 3.0, 86)
    kern = fx.GaussianKernel(0.1)
    c1, c2 = fx.Context(0), fx.Context(0)
    V1, L1 = fx.kpca(X, kern, k=3, ctx=c1)
    V2, L2 = fx.kpca(X, kern, k=3, ctx=c2)
    assert np.array_equal(L1, L2) and np.array_equal(V1, V2)      # deterministic, bit for bit
    c1.close(); c2.close()


# ------------------------------------------------------------------ linear pca / lda (pinned upstream)
def test_pca_reference_tests_on_gpu(gpu_ctx):
    """test_kpca.jl:29-42 through the CUDA path: pca vs eig(Sigma/(n-1)) for every (alpha, eps),
    transform_pca, and test_kpca.jl:44-56 with BOTH sides on the GPU."""
    n, p = 20, 5
    X = np.random.default_rng(6).random((n, p))
    H = X - X.mean(0)
    S = H.T @ H / (n - 1)
    for alpha in (0.0, 0.5):
        for eps in (0.0, 1.0):
            V1, D1 = fxb200.pca(X, alpha, eps, ctx=gpu_ctx)
            D2, V2 = np.linalg.eigh(O.perturb(O.regularize(S, alpha, np.trace(S) / p), eps))
            np.testing.assert_allclose(np.abs(V1), np.abs(V2[:, ::-1]), atol=1e-10)
            np.testing.assert_allclose(np.abs(D1), np.abs(D2[::-1]), rtol=1e-12)
            np.testing.assert_allclose(np.abs(fxb200.transform_pca(V1, X, ctx=gpu_ctx)), np.abs(X @ V2[:, ::-1]),
                                       atol=1e-10)
    _V, D1 = fxb200.pca(X, 1.0, 1.0, ctx=gpu_ctx)          # alpha = 1: degenerate, eigenvalues only
    np.testing.assert_allclose(D1, O.pca(X, 1.0, 1.0)[1], rtol=1e-12)
    Xc = X - X.mean(0)
    for eps in (0.0, 1.0):
        V1, D1 = fxb200.kpca(X, fxb200.PolynomialKernel(1.0, 0.0, 1.0), 0.0, eps, ctx=gpu_ctx)
        V2, D2 = fxb200.pca(X, 0.0, eps, ctx=gpu_ctx)
        assert np.abs((V1.T @ Xc @ Xc.T @ V1 / (n - 1))[:p, :p] - V2.T @ S @ V2).max() < 1e-12
        np.testing.assert_allclose(np.abs(D1)[:p], np.abs(D2), rtol=1e-12)


@pytest.mark.parametrize("n,p,k", [(12, 3, 4), (500, 37, 5), (3000, 128, 10)])
def test_lda_matches_oracle(gpu_ctx, n, p, k):
    rng = np.random.default_rng(91)
    y = np.repeat(np.arange(1, k + 1), n // k)
    y = np.concatenate([y, np.full(n - len(y), 1)])
    X = rng.random((n, p)) + 3.0 * rng.random((k, p))[y - 1]
    freq = np.bincount(y - 1, minlength=k) / n
    for (ab, eb, aw, ew) in [(0.0, 0.0, 0.0, 1e-8), (0.2, 0.1, 0.3, 0.05)]:
        Vo, Lo = O.lda(X, y, k, freq, ab, eb, aw, ew)
        V, L = fxb200.lda(X, y, k, freq, ab, eb, aw, ew, ctx=gpu_ctx)
        scale = np.abs(Lo).max()
        assert np.abs(L - Lo).max() <= 1e-9 * scale
        nz = int(np.sum(Lo > 1e-8 * scale)) if (ab == 0 and eb == 0) else p
        nz = min(nz, k - 1) if (ab == 0 and eb == 0) else min(nz, 3)
        assert sin_angle(V[:, :nz], Vo[:, :nz]) < 1e-6
        H_b, H_w, _ = O.lda_matrices(X, y, k, freq)
        _Sb, Sw = O.scatter_matrices(H_b, ab, eb, H_w, aw, ew)
        np.testing.assert_allclose(V.T @ Sw @ V, np.eye(p), atol=1e-7)       # sygvd normalisation
        Z = rng.random((33, p))
        np.testing.assert_allclose(fxb200.transform_lda(V, Z, ctx=gpu_ctx), Z @ V, rtol=1e-12, atol=1e-12)
    # upstream pencil is singular without eps_w when p > k - 1 ... reported as "not positive definite"
    if p > n - k:
        with pytest.raises(fxb200.FxError):
            fxb200.lda(X, y, k, freq, ctx=gpu_ctx)


# ------------------------------------------------------------------ smallest sizes
@pytest.mark.parametrize("n,d", [(2, 1), (3, 2), (5, 3), (17, 4)])
def test_kpca_tiny(gpu_ctx, n, d):
    X = np.random.default_rng(100 + n).random((n, d))
    for kern, okern in [(fxb200.GaussianKernel(0.7), O.GaussianKernel(0.7)),
                        (fxb200.PolynomialKernel(1.0, 0.0, 1.0), O.PolynomialKernel(1.0, 0.0, 1.0))]:
        V, L = fxb200.kpca(X, kern, ctx=gpu_ctx)                   # k = n like the reference
        Vo, Lo = O.kpca(X, okern)
        assert V.shape == (n, n)
        assert np.abs(L - Lo).max() <= 1e-12 * max(np.abs(Lo).max(), 1e-300)
        assert np.abs(V.T @ V - np.eye(n)).max() < 1e-12


@pytest.mark.parametrize("n,d,c", [(300, 6, 3), (500, 12, 5), (512, 20, 10)])
def test_klda_dense_path_matches_sygvd(gpu_ctx, n, d, c):
    """n <= 512: the reference's LDA pipeline applied literally to K (klda.jl:34-61 with the intended
    pencil), including the between-class regularisation alpha_b, eps_b (klda.jl:36-37) that the
    reduced iterative path cannot take -- against the oracle's dsygvd, all n components."""
    X, y, freq = _klda_case(n, d, c, 2.0, 74, 1.0 / d)
    kern, okern = fxb200.GaussianKernel(1.0 / d), O.GaussianKernel(1.0 / d)
    K = O.kernelmatrix(okern, X)
    H_b, H_w, _ = O.lda_matrices(K, y, c, freq)
    for (ab, eb, aw, ew_rel) in [(0.0, 0.0, 0.0, 1e-4), (0.2, 0.01, 0.1, 1e-3), (0.0, 1e-3, 0.0, 1e-2)]:
        _Sb0, Sw0 = O.scatter_matrices(H_b, 0, 0, H_w, 0, 0)
        ew = ew_rel * np.trace(Sw0) / n
        Ao, Lo = O.klda(X, okern, y, c, freq, ab, eb, aw, ew, ncomp=n)
        ncomp = n if (ab or eb) else c - 1
        A, L, model, st = fxb200.klda(X, kern, y, c, freq, ab, eb, aw, ew, ncomp=ncomp, ctx=gpu_ctx,
                                      return_model=True, return_stats=True)
        assert st["dense_path"] == 1 and st["converged"] == 1
        scale = np.abs(Lo).max()
        assert np.abs(L - Lo[:ncomp]).max() <= 1e-8 * scale
        Sb, Sw = O.scatter_matrices(H_b, ab, eb, H_w, aw, ew)
        kk = min(ncomp, c - 1)
        R = Sb @ A[:, :kk] - (Sw @ A[:, :kk]) * L[None, :kk]
        assert (np.linalg.norm(R, axis=0) / np.linalg.norm(Sb @ A[:, :kk], axis=0)).max() < 1e-7
        np.testing.assert_allclose(A[:, :kk].T @ Sw @ A[:, :kk], np.eye(kk), atol=1e-7)
        Z, _ = mix(33, d, c, 2.0, 75)
        relclose(model.transform(Z), O.transform_klda(X, okern, A, Z), 1e-10)
        model.free()


def test_klda_orphan_test_sizes(gpu_ctx):
    """test/test_klda.jl's sizes (nc = 3 per class, p = 3, k = 4 classes), through KLDA."""
    rng = np.random.default_rng(8)
    nc, pp, k = 3, 3, 4
    y = np.repeat(np.arange(1, k + 1), nc)
    X = np.vstack([rng.random((nc, pp)) + 10 * rng.random((1, 3)) for _ in range(k)])
    freq = np.full(k, 1.0 / k)
    kern, okern = fxb200.GaussianKernel(0.05), O.GaussianKernel(0.05)
    Ao, Lo = O.klda(X, okern, y, k, freq, eps_w=1e-6)
    A, L = fxb200.klda(X, kern, y, k, freq, eps_w=1e-6, ctx=gpu_ctx)
    np.testing.assert_allclose(L, Lo, rtol=1e-6)
    assert sin_angle(A, Ao) < 1e-4


def test_float32_inputs_keep_their_eltype(gpu_ctx):
    """The reference is generic over (Float32, Float64) (test/runtests.jl:7, eltype checks at
    test_common.jl:66,73): Float32 in -> Float32 out, computed in float64 underneath."""
    X = np.random.default_rng(7).random((20, 5)).astype(np.float32)
    V, D = fxb200.kpca(X, fxb200.PolynomialKernel(1.0, 0.0, 1.0), ctx=gpu_ctx)
    assert V.dtype == np.float32 and D.dtype == np.float32
    V2, D2 = fxb200.pca(X, ctx=gpu_ctx)
    assert V2.dtype == np.float32 and D2.dtype == np.float32
    np.testing.assert_allclose(np.abs(D)[:5], np.abs(D2), rtol=1e-5)
    T = fxb200.transform_kpca(X, fxb200.GaussianKernel(0.5), V[:, :2].astype(np.float32), X, ctx=gpu_ctx)
    assert T.dtype == np.float32 and T.shape == (20, 2)


# ------------------------------------------------------------------ a plain C host
def test_plain_c_caller(tmp_path):
    """tests/host/abi_smoke.c (C99, no Python in the loop) fits, transforms and checks the in-sample
    identity through the C ABI."""
    import subprocess
    from test_library_abi import build_c_smoke
    exe = build_c_smoke(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "C_ABI_OK" in r.stdout, r.stdout + r.stderr
