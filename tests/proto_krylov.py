"""NumPy prototype of the device eigensolver (test infrastructure; mirrors
csrc/solver.cu step for step so the algorithm can be validated on the CPU and,
with ``ops``, across gloo ranks).

Thick-restart block Krylov (Krylov-Schur form) with explicit projections:
  * basis Q (n_loc x m) row-sharded, next block Qn orthonormal and orthogonal to Q
  * step:  W = K Qn ; H = [Q Qn]' W (allreduce) ; T[:, blk] = H ; W -= [Q Qn] H ;
           second CGS pass ; SVQB orthonormalisation twice (b x b Jacobi) giving the
           coupling C with  K [Q Qn] = [Q Qn] T + Qnew [0 .. 0 C]
  * Rayleigh-Ritz on T (Jacobi), residual_i = ||C s_i[last b]||
  * restart: keep r Ritz vectors, T = diag(theta)
"""
import numpy as np


class LocalOps:
    """Single-rank collectives."""
    def allreduce(self, a):
        return a

    def allgather_rows(self, a):
        return a


def hash_uniform(rows, cols, seed):
    """Deterministic pseudo-random start block, identical to the device hash
    (splitmix64 of (row, col, seed) -> uniform in [-1, 1))."""
    r = rows.astype(np.uint64)[:, None]
    c = cols.astype(np.uint64)[None, :]
    with np.errstate(over="ignore"):
        z = (r * np.uint64(0x9E3779B97F4A7C15)) ^ (c * np.uint64(0xBF58476D1CE4E5B9)) ^ np.uint64(seed)
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (2.0 ** -52) - 1.0


def chol_factors(Gs, thr):
    """Right-looking Cholesky of the scaled Gram matrix as chol_factors_kernel (smalldense.cu)
    does it: returns lower L with Gs = L L', or None as soon as a pivot is <= thr."""
    b = Gs.shape[0]
    L = np.tril(0.5 * (Gs + Gs.T))
    for j in range(b):
        piv = L[j, j]
        if not piv > thr:
            return None
        L[j:, j] /= np.sqrt(piv)
        for c in range(j + 1, b):
            L[c:, c] -= L[c:, j] * L[c, j]
    return L


def svqb(W, ops, tol=1e-14, chol_thr=1e-8):
    """One block orthonormalisation pass: G = W'W scaled by its diagonal; CholQR when every
    pivot of the Cholesky exceeds chol_thr (W_orth = W D^-1 L^-T), else SVQB: eigendecompose,
    W <- W D^-1/2 U L^-1/2.  Returns (W_orth, Cfac, ndef) with W_old = W_orth @ Cfac."""
    G = ops.allreduce(W.T @ W)
    d = np.sqrt(np.maximum(np.diag(G), 1e-300))
    Gs = G / d[:, None] / d[None, :]
    L = chol_factors(Gs, chol_thr) if chol_thr is not None else None
    if L is not None:
        Li = np.linalg.inv(L)
        M = Li.T / d[:, None]                             # W_orth = W M
        Cfac = L.T * d[None, :]                           # W = W_orth Cfac
        return W @ M, Cfac, 0
    lam, U = np.linalg.eigh(Gs)
    lmax = lam.max()
    ndef = int((lam < tol * lmax).sum())
    lam = np.maximum(lam, tol * lmax)
    M = (U / np.sqrt(lam)[None, :]) / d[:, None]          # W_orth = W M
    Cfac = (np.sqrt(lam)[:, None] * U.T) * d[None, :]     # W = W_orth Cfac
    return W @ M, Cfac, ndef


def topk(matvec, n_loc, row0, k, ops=None, b=None, max_blocks=5, tol=1e-12,
         max_steps=200, seed=1234, keep=None, verbose=False):
    """matvec(Vfull[n x b]) -> local rows of K V.  Returns (V_loc, theta, info)."""
    ops = ops or LocalOps()
    b = b or (k + 8 + 7) // 8 * 8
    mcap = max_blocks * b
    keep = keep or b
    rows = np.arange(row0, row0 + n_loc)
    Q = np.zeros((n_loc, mcap + b))
    T = np.zeros((mcap + b, mcap + b))
    W = hash_uniform(rows, np.arange(b), seed)
    W, _, _ = svqb(W, ops); W, _, _ = svqb(W, ops)
    m = 0
    Q[:, :b] = W
    steps = 0
    hist = []
    while True:
        Qn = Q[:, m:m + b]
        W = matvec(ops.allgather_rows(Qn))
        steps += 1
        Qa = Q[:, :m + b]
        H = ops.allreduce(Qa.T @ W)
        W = W - Qa @ H
        H2 = ops.allreduce(Qa.T @ W)
        W = W - Qa @ H2
        H = H + H2
        T[:m + b, m:m + b] = H
        T[m:m + b, :m] = H[:m].T
        blk = T[m:m + b, m:m + b]
        T[m:m + b, m:m + b] = 0.5 * (blk + blk.T)
        W, C1, nd1 = svqb(W, ops)
        W, C2, nd2 = svqb(W, ops)
        C = C2 @ C1
        m += b
        # Rayleigh-Ritz
        theta, S = np.linalg.eigh(T[:m, :m])
        theta, S = theta[::-1], S[:, ::-1]
        res = np.linalg.norm(C @ S[m - b:m, :k], axis=0)
        scale = max(abs(theta[0]), 1e-300)
        hist.append((steps, m, res.max() / scale))
        if verbose:
            print("step %d m=%d maxres/theta1=%.3e ndef=%d" % (steps, m, res.max() / scale, nd1))
        if res.max() <= tol * scale or steps >= max_steps:
            V = Q[:, :m] @ S[:, :k]
            return V, theta[:k], dict(steps=steps, res=res / scale, hist=hist,
                                      converged=bool(res.max() <= tol * scale))
        if m + b > mcap:
            r = keep
            Q[:, :r] = Q[:, :m] @ S[:, :r]
            Q[:, r:r + b] = W
            T[:] = 0
            T[np.arange(r), np.arange(r)] = theta[:r]
            m = r
        else:
            Q[:, m:m + b] = W


def jacobi_eigh(A, max_sweeps=30, tol=1e-30):
    """Parallel-order cyclic Jacobi exactly as the device kernel schedules it
    (round-robin tournament, all pairs of a round rotated 'simultaneously')."""
    A = np.array(A, dtype=np.float64, copy=True)
    m = A.shape[0]
    mp = m + (m & 1)
    V = np.eye(m)
    sweeps = 0
    for sweep in range(max_sweeps):
        Ao = A - np.diag(np.diag(A))
        off = (Ao * Ao).sum()                    # summed directly: no cancellation
        if off <= tol * (np.diag(A) ** 2).sum() or off == 0.0:
            break
        sweeps += 1
        for rnd in range(mp - 1):
            pairs = []
            for i in range(mp // 2):
                if i == 0:
                    p, q = mp - 1, rnd
                else:
                    p, q = (rnd + i) % (mp - 1), (rnd - i + (mp - 1)) % (mp - 1)
                if p > q:
                    p, q = q, p
                if q < m:
                    pairs.append((p, q))
            J = np.eye(m)
            for p, q in pairs:
                apq = A[p, q]
                if abs(apq) < 1e-300:
                    continue
                tau = (A[q, q] - A[p, p]) / (2 * apq)
                t = (1.0 if tau >= 0 else -1.0) / (abs(tau) + np.sqrt(1 + tau * tau))
                c = 1 / np.sqrt(1 + t * t); s = t * c
                J[p, p] = c; J[q, q] = c; J[p, q] = s; J[q, p] = -s
            A = J.T @ A @ J
            V = V @ J
    d = np.diag(A).copy()
    order = np.argsort(-d, kind="stable")
    return d[order], V[:, order], sweeps
