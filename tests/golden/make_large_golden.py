"""Mints the LARGE-size CPU fixtures that the GPU tests and bench.py compare against without
having to redo minutes of CPU work on the GPU box.  NumPy / SciPy only -- nothing in here touches
the CUDA library, so the fixtures are independent of the code under test.

    python tests/golden/make_large_golden.py c2         # ~5 min on 8 cores, 9 GB
    python tests/golden/make_large_golden.py headline   # ~15 min on 8 cores, 36 GB

c2        BASELINE.json configs[1] (RBF n=16,384 d=128 k=8): the reference's dense path restated
          (kpca.jl:49-57: kernelmatrix, centerkernelmatrix!, dsyevd 'V','U', reverse, /(n-1)) run in
          full; stores the 24 leading eigenvalues, the k leading eigenvectors, the trace and the
          transform of 256 fresh queries.  -> tests/golden/large_c2_dsyevd.npz
headline  the bench configuration (RBF n=65,536 d=256 k=16).  dsyevd is infeasible (137 GB, hours),
          so the centred kernel matrix is built once on the CPU (34 GB, in place) and its 24 leading
          eigenpairs come from ARPACK (scipy.sparse.linalg.eigsh, tol = 0 -> machine precision).
          Stores the eigenvalues /(n-1), the trace and a 48 x 48 sketch Om' V V' Om of the projector
          onto the leading k-dimensional eigenspace (Om = seeded Gaussian n x 48), which lets a test
          bound the subspace angle without shipping the 8 MB eigenvector block.
          -> tests/golden/large_headline_eigsh.npz
Inputs are regenerated from (generator, seed) by the tests; x_checksum guards the generator."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O                     # noqa: E402
from fxb200.synth import mix, CONFIGS  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SKETCH_COLS, SKETCH_SEED = 48, 777


def sketch_omega(n):
    return np.random.default_rng(SKETCH_SEED).standard_normal((n, SKETCH_COLS))


def make_c2():
    cfg = CONFIGS["C2"]
    n, d, k = cfg["n"], cfg["d"], cfg["k"]
    X, _ = mix(n, d, cfg["c"], cfg["sep"], cfg["seed"])
    Z, _ = mix(256, d, cfg["c"], cfg["sep"], cfg["seed"] + 1)
    kern = O.GaussianKernel(*cfg["params"])
    t0 = time.time()
    V, L, st = O.kpca(X, kern, return_state=True)         # full dsyevd, all n pairs
    print("c2: dense oracle fit %.1f s" % (time.time() - t0), flush=True)
    Vk = V[:, :k].copy()
    T = O.transform_kpca(X, kern, Vk, Z)
    Kc, _, _ = O.centerkernelmatrix(O.kernelmatrix(kern, X))
    np.savez_compressed(
        os.path.join(HERE, "large_c2_dsyevd.npz"),
        config=np.array("C2"), L=L[:24], V=Vk, trace=np.array(np.trace(Kc)), T=T,
        colmean_head=st["r"][:64], grandmean=np.array(st["mu"]),
        x_checksum=np.array([X.sum(), (X * X).sum()]), z_seed=np.array(cfg["seed"] + 1),
        how=np.array("oracle.kpca: NumPy Gram + map, centring, LAPACK dsyevd('V','U'), all n pairs"))
    print("c2: L[:9] =", L[:9], flush=True)


def make_headline():
    from scipy.sparse.linalg import eigsh
    cfg = CONFIGS["HEADLINE"]
    n, d, k = cfg["n"], cfg["d"], cfg["k"]
    X, _ = mix(n, d, cfg["c"], cfg["sep"], cfg["seed"])
    Xc = np.ascontiguousarray(X)                          # row-major copy for the blocked GEMM
    gamma = cfg["params"][0]
    nx = np.einsum("ij,ij->i", Xc, Xc)
    t0 = time.time()
    K = np.empty((n, n))
    B = 4096
    for i0 in range(0, n, B):                             # oracle.kernelmatrix, block by block
        G = Xc[i0:i0 + B] @ Xc.T
        G *= -2.0
        G += nx[i0:i0 + B, None]
        G += nx[None, :]
        np.maximum(G, 0.0, out=G)                         # cancellation guard (oracle/kernels.py)
        idx = np.arange(i0, min(i0 + B, n))
        G[idx - i0, idx] = 0.0                            # exact zero distance on the diagonal
        G *= -gamma
        np.exp(G, out=K[i0:i0 + B])
    print("headline: K built %.1f s" % (time.time() - t0), flush=True)
    r = np.zeros(n)
    for i0 in range(0, n, B):
        r += K[i0:i0 + B].sum(axis=0)
    r /= n
    mu = r.sum() / n
    for i0 in range(0, n, B):                             # centerkernelmatrix!, in place
        blk = K[i0:i0 + B]
        blk -= r[None, :]
        blk -= r[i0:i0 + B, None]
        blk += mu
    trace = float(np.trace(K))
    print("headline: centred %.1f s, trace %.6f" % (time.time() - t0, trace), flush=True)
    nev = 24
    w, U = eigsh(K, k=nev, which="LA", ncv=96, tol=0.0)
    order = np.argsort(-w)
    w, U = w[order], U[:, order]
    print("headline: eigsh done %.1f s" % (time.time() - t0), flush=True)
    res = np.linalg.norm(K @ U[:, :k] - U[:, :k] * w[:k], axis=0) / w[0]
    Om = sketch_omega(n)
    Y = U[:, :k].T @ Om
    np.savez_compressed(
        os.path.join(HERE, "large_headline_eigsh.npz"),
        config=np.array("HEADLINE"), L=w / (n - 1), trace=np.array(trace), grandmean=np.array(mu),
        colmean_head=r[:64], residual=res, sketch=Y.T @ Y,
        sketch_cols=np.array(SKETCH_COLS), sketch_seed=np.array(SKETCH_SEED),
        x_checksum=np.array([X.sum(), (X * X).sum()]), seconds=np.array(time.time() - t0),
        how=np.array("NumPy blocked Gram + map + centring (oracle definitions), ARPACK eigsh "
                     "k=24 ncv=96 tol=0 on the dense centred matrix"))
    print("headline: L =", w[:k] / (n - 1), "max residual", res.max(), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["c2", "headline"]
    if "c2" in which:
        make_c2()
    if "headline" in which:
        make_headline()
