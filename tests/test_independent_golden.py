"""Parity against fixtures that do NOT come from oracle/ nor from the CUDA library:

  tests/golden/sk_*.npz                 scikit-learn KernelPCA (rbf / poly-3 / sigmoid; n = 512, 4096)
  tests/golden/large_c2_dsyevd.npz      the reference's dense dsyevd path at BASELINE configs[1]
                                        (n = 16,384), minutes of CPU work done once
  tests/golden/large_headline_eigsh.npz ARPACK on the CPU-built centred matrix of the bench
                                        configuration (n = 65,536)

CPU tests pin the oracle to the scikit-learn fixtures; GPU tests compare the CUDA path (through the
C ABI) DIRECTLY with every fixture, so the chain  reference semantics -> independent CPU code ->
CUDA  has no self-referential link.  Tolerances: eigenvalues 1e-10 relative, subspace angle 1e-8
(BASELINE.json north_star), transform 1e-9 relative to its largest entry."""
import glob
import os

import numpy as np
import pytest

import oracle as O
import fxb200
from fxb200.synth import mix, CONFIGS

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
SK = sorted(glob.glob(os.path.join(GOLD, "sk_*.npz")))


def sin_angle(A, B):
    Qa, _ = np.linalg.qr(A)
    Qb, _ = np.linalg.qr(B)
    return np.linalg.norm(Qb - Qa @ (Qa.T @ Qb), 2)


def sk_case(path):
    g = np.load(path)
    n, d, c, sep, seed, k, m, zseed = g["meta"]
    X, _ = mix(int(n), int(d), int(c), float(sep), int(seed))
    Z, _ = mix(int(m), int(d), int(c), float(sep), int(zseed))
    np.testing.assert_allclose([X.sum(), (X * X).sum()], g["x_checksum"], rtol=1e-13)
    gamma, coef0, degree = g["params"]
    kname = str(g["kernel"])
    spec = {"rbf": ("gaussian", (gamma,)), "poly": ("polynomial", (gamma, coef0, degree)),
            "sigmoid": ("sigmoid", (gamma, coef0))}[kname]
    return g, X, Z, spec, int(k)


def okernel(spec):
    return {"gaussian": O.GaussianKernel, "polynomial": O.PolynomialKernel,
            "sigmoid": O.SigmoidKernel}[spec[0]](*spec[1])


def align_signs(T, Tref):
    s = np.sign(np.sum(T * Tref, axis=0))
    s[s == 0] = 1.0
    return T * s[None, :]


def check_against_sklearn(g, V, L, T, k):
    np.testing.assert_allclose(L, g["L"][:k], rtol=1e-10)
    assert sin_angle(V, g["V"]) < 1e-8
    Tr = g["T"]
    err = np.abs(align_signs(T, Tr) - Tr).max() / np.abs(Tr).max()
    assert err < 1e-9, "transform rel err %.3e" % err


# ------------------------------------------------------------------------------------ CPU
@pytest.mark.parametrize("path", SK, ids=[os.path.basename(f) for f in SK])
def test_oracle_matches_sklearn_fixture(path):
    """The oracle (restated reference + restated MLKernels definitions) reproduces scikit-learn's
    eigenvalues / (n-1), eigenvectors and K~_z V for the kernels upstream never tests."""
    g, X, Z, spec, k = sk_case(path)
    if len(X) > 1024:
        pytest.skip("n = 4096 oracle fit is covered by the GPU-side test; keep the CPU suite short")
    kern = okernel(spec)
    V, L, _ = O.kpca_topk(X, kern, k)
    check_against_sklearn(g, V, L, O.transform_kpca(X, kern, V, Z), k)


def test_large_fixtures_are_consistent():
    """Generator drift guard + internal consistency of the two large fixtures."""
    for name, cfgname in (("large_c2_dsyevd.npz", "C2"), ("large_headline_eigsh.npz", "HEADLINE")):
        g = np.load(os.path.join(GOLD, name))
        cfg = CONFIGS[cfgname]
        assert str(g["config"]) == cfgname
        L = g["L"]
        assert np.all(np.diff(L) < 0) and L[cfg["k"] - 1] > 5 * L[cfg["k"]]      # the spectral gap
        assert L.sum() * (cfg["n"] - 1) < float(g["trace"])
    g = np.load(os.path.join(GOLD, "large_headline_eigsh.npz"))
    assert g["residual"].max() < 1e-12                                           # ARPACK converged
    S = g["sketch"]
    assert np.abs(S - S.T).max() < 1e-9 and abs(np.trace(S) / S.shape[0] - CONFIGS["HEADLINE"]["k"]) < 8


# ------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("path", SK, ids=[os.path.basename(f) for f in SK])
def test_cuda_matches_sklearn_fixture(gpu_ctx, path):
    g, X, Z, spec, k = sk_case(path)
    kern = fxb200.from_spec(*spec)
    V, L, model, st = fxb200.kpca(X, kern, k=k, ctx=gpu_ctx, return_model=True, return_stats=True)
    assert st["converged"] == 1
    check_against_sklearn(g, V, L, model.transform(Z), k)
    model.free()


@pytest.mark.gpu
def test_cuda_c2_matches_dense_dsyevd_fixture(gpu_ctx):
    """BASELINE configs[1] (RBF n = 16,384 d = 128 k = 8) against the committed result of the
    reference's dense path (kpca.jl:49-57 restated: full dsyevd, 315 s of CPU)."""
    g = np.load(os.path.join(GOLD, "large_c2_dsyevd.npz"))
    cfg = CONFIGS["C2"]
    X, _ = mix(cfg["n"], cfg["d"], cfg["c"], cfg["sep"], cfg["seed"])
    np.testing.assert_allclose([X.sum(), (X * X).sum()], g["x_checksum"], rtol=1e-13)
    Z, _ = mix(256, cfg["d"], cfg["c"], cfg["sep"], int(g["z_seed"]))
    k = cfg["k"]
    kern = fxb200.GaussianKernel(*cfg["params"])
    V, L, model, st = fxb200.kpca(X, kern, k=k, ctx=gpu_ctx, return_model=True, return_stats=True)
    assert st["converged"] == 1 and st["dense_path"] == 0
    np.testing.assert_allclose(L, g["L"][:k], rtol=1e-10)
    assert sin_angle(V, g["V"]) < 1e-8
    assert abs(st["trace"] / float(g["trace"]) - 1) < 1e-12
    Tr = g["T"]
    T = model.transform(Z)
    # the fixture's T belongs to ITS eigenvector signs / rotation inside clusters: compare in the
    # rotation-invariant form T T' (= K~_z P K~_z', P the projector on the leading eigenspace)
    assert np.abs(T @ T.T - Tr @ Tr.T).max() <= 1e-9 * np.abs(Tr @ Tr.T).max()
    model.free()


@pytest.mark.gpu
def test_cuda_n8192_matches_live_dsyevd(gpu_ctx):
    """n = 8,192: the dense CPU path (oracle: LAPACK dsyevd 'V','U') is run live (tens of seconds on
    the box's host cores) next to the CUDA fit of the same input."""
    n, d, k = 8192, 128, 8
    X, _ = mix(n, d, k + 1, 3.0, 2101)
    kern, okern = fxb200.GaussianKernel(1.0 / d), O.GaussianKernel(1.0 / d)
    V, L, st = fxb200.kpca(X, kern, k=k, ctx=gpu_ctx, return_stats=True)
    Vo, Lo = O.kpca(X, okern)
    np.testing.assert_allclose(L, Lo[:k], rtol=1e-10)
    assert sin_angle(V, Vo[:, :k]) < 1e-8
    assert st["converged"] == 1


@pytest.mark.gpu
def test_cuda_headline_matches_cpu_eigsh_fixture(gpu_ctx):
    """The bench configuration itself (RBF n = 65,536 d = 256 k = 16; K = 34 GB) against ARPACK run on
    the CPU-built centred matrix: eigenvalues, trace, and the 48 x 48 sketch Om' V V' Om of the
    projector onto the leading eigenspace (bounds the subspace angle without an 8 MB fixture)."""
    g = np.load(os.path.join(GOLD, "large_headline_eigsh.npz"))
    cfg = CONFIGS["HEADLINE"]
    n, k = cfg["n"], cfg["k"]
    X, _ = mix(n, cfg["d"], cfg["c"], cfg["sep"], cfg["seed"])
    np.testing.assert_allclose([X.sum(), (X * X).sum()], g["x_checksum"], rtol=1e-13)
    V, L, st = fxb200.kpca(X, fxb200.GaussianKernel(*cfg["params"]), k=k, ctx=gpu_ctx, return_stats=True)
    assert st["converged"] == 1
    np.testing.assert_allclose(L, g["L"][:k], rtol=1e-10)
    assert abs(st["trace"] / float(g["trace"]) - 1) < 1e-12
    Om = np.random.default_rng(int(g["sketch_seed"])).standard_normal((n, int(g["sketch_cols"])))
    Y = V.T @ Om
    S = Y.T @ Y
    # For Gaussian Om the entries of Om'(P1 - P2)Om are ~ N(0, |P1 - P2|_F^2) (twice that on the
    # diagonal), so the largest of the 48^2 entries is about 3.5 |P1 - P2|_F, and sin(theta) =
    # |P1 - P2|_2 <= |P1 - P2|_F: a maximum below 1e-8 puts the subspace angle below ~3e-9.
    assert np.abs(S - g["sketch"]).max() < 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("n", [128, 301, 1000, 1237])
def test_fused_centring_writeback_elementwise(gpu_ctx, n):
    """The centring pass a fit actually runs (symm.cu CENTER variant: centred in registers inside the
    first K V pass, written back through the TMA store warp) against oracle.centerkernelmatrix,
    element by element, on ragged sizes -- kpca.jl:52."""
    X, _ = mix(n, 9, 3, 2.0, 34)
    K = O.kernelmatrix(O.GaussianKernel(0.1), X)
    V = np.random.default_rng(n).standard_normal((n, 16))
    Kc, Y, cm, gm = fxb200.center_symm(K, V, ctx=gpu_ctx)
    Ko, r, mu = O.centerkernelmatrix(K)
    scale = np.abs(Ko).max()
    assert np.abs(Kc - Ko).max() <= 1e-14 * scale
    assert np.abs(cm - r).max() <= 1e-14 * np.abs(r).max() and abs(gm - mu) <= 1e-14 * abs(mu)
    assert np.abs(Y - Ko @ V).max() <= 1e-12 * np.abs(Ko @ V).max()
    # and it is the same arithmetic, element for element, as the standalone kernel
    Kc2, _, _ = fxb200.centerkernelmatrix(K, ctx=gpu_ctx)
    assert np.array_equal(Kc, Kc2)


@pytest.mark.gpu
def test_model_with_more_than_64_components(gpu_ctx):
    """k = n > 64 with a model (the bias buffer held 64 entries): the dense path returns all
    n pairs like the reference (common.jl:164), and transform handles any k in chunks of 64.  With
    k = n the null eigenvector (proportional to 1) is among the components, so the transform needs
    the full double-centring formula, row-mean term included."""
    n, d = 200, 6
    X, _ = mix(n, d, 4, 3.0, 87)
    Z, _ = mix(37, d, 4, 3.0, 88)
    kern, okern = fxb200.GaussianKernel(0.2), O.GaussianKernel(0.2)
    V, L, model = fxb200.kpca(X, kern, ctx=gpu_ctx, return_model=True)
    assert V.shape == (n, n) and model.k == n
    T = model.transform(Z)
    To = O.transform_kpca(X, okern, V, Z)
    assert np.abs(T - To).max() <= 1e-10 * np.abs(To).max()
    assert np.abs(model.transform(Z, _naive=True) - To).max() <= 1e-10 * np.abs(To).max()
    model.free()


@pytest.mark.gpu
@pytest.mark.parametrize("k", [3, 8, 70])
def test_transform_arbitrary_components_full_centring(gpu_ctx, k):
    """transform_kpca(X, kernel, W, Z) with a caller-supplied W that is NOT orthogonal to 1:
    (k(Z,X) - 1 r' - r_z 1' + mu) W, the defined semantics of kpca.jl:64-68."""
    n, d, m = 333, 5, 141
    X, _ = mix(n, d, 3, 3.0, 89)
    Z, _ = mix(m, d, 3, 3.0, 90)
    W = np.random.default_rng(91).standard_normal((n, k)) + 0.5
    for kern, okern in [(fxb200.GaussianKernel(0.3), O.GaussianKernel(0.3)),
                        (fxb200.PolynomialKernel(0.2, 1.0, 2.0), O.PolynomialKernel(0.2, 1.0, 2.0))]:
        To = O.transform_kpca(X, okern, W, Z)
        T = fxb200.transform_kpca(X, kern, W, Z, ctx=gpu_ctx)
        assert np.abs(T - To).max() <= 1e-11 * np.abs(To).max()
        To2 = O.transform_kpca(X, okern, W, Z, center=False)
        T2 = fxb200.transform_klda(X, kern, W, Z, ctx=gpu_ctx)
        assert np.abs(T2 - To2).max() <= 1e-12 * np.abs(To2).max()


@pytest.mark.gpu
@pytest.mark.parametrize("n,d,c,rel", [(600, 8, 3, 1e-3), (1500, 32, 5, 1e-6), (2000, 64, 10, 1e-6)])
def test_klda_independent_residual(gpu_ctx, n, d, c, rel):
    """KLDA has no upstream test and no third-party implementation to compare with, so the check is
    the DEFINITION, evaluated in plain NumPy from explicitly formed scatter matrices (no oracle.klda,
    no sygvd): Sigma_b a = lambda Sigma_w' a and a' Sigma_w' a = 1 (klda.jl:34-41 with the intended
    pencil; sygvd itype-1 normalisation, common.jl:172), plus completeness: the returned lambdas are
    ALL the non-zero generalized eigenvalues (trace identity tr(Sigma_w'^-1 Sigma_b))."""
    X, y0 = mix(n, d, c, 2.0, 71)
    y = y0 + 1
    cnt = np.bincount(y0, minlength=c).astype(float)
    freq = cnt / n
    gamma = 1.0 / d
    sq = (X * X).sum(1)
    K = np.exp(-gamma * np.maximum(sq[:, None] + sq[None, :] - 2.0 * X @ X.T, 0.0))
    E = np.zeros((n, c)); E[np.arange(n), y0] = 1.0
    M = (E.T @ K) / cnt[:, None]                                   # class means of the rows of K
    Hw = (K - M[y0]) / (n - c)                                     # klda.jl:49-50
    mu = freq @ M
    Hb = np.sqrt(freq)[:, None] * (M - mu[None, :])                # klda.jl:51-53
    Sb, Sw = Hb.T @ Hb, Hw.T @ Hw
    eps_w = rel * np.trace(Sw) / n
    Swr = Sw + eps_w * np.eye(n)
    A, L, st = fxb200.klda(X, fxb200.GaussianKernel(gamma), y, c, freq, eps_w=eps_w, ctx=gpu_ctx,
                           return_stats=True, opts=fxb200.make_opts(cg_tol=1e-13))
    assert st["converged"] == 1 and L.shape == (c - 1,) and np.all(np.diff(L) <= 0) and L[-1] > 0
    R = Sb @ A - (Swr @ A) * L[None, :]
    res = np.linalg.norm(R, axis=0) / np.linalg.norm(Sb @ A, axis=0)
    assert res.max() <= 1e-9, "generalized eigen-residual %.3e" % res.max()
    G = A.T @ Swr @ A
    assert np.abs(G - np.eye(c - 1)).max() <= 1e-9                 # B-orthonormal
    tr = np.trace(np.linalg.solve(Swr, Sb))                        # sum of all generalized eigenvalues
    assert abs(L.sum() / tr - 1) <= 1e-8
