"""Step-for-step NumPy model of the iterative KLDA solve in csrc/klda.cu (test infrastructure).

Sigma_w' = cw K (I-P) K + s I = A - U U' with A = cw K^2 + s I, U = sqrt(cw) K E D^-1/2.  Because A and
K commute, Woodbury collapses to
    Sigma_w'^-1 (K E g) = (1/s) Z W^-1 D g,       Y = A^-1 E, Z = K Y, W = E'Y
    Sigma_w'^-1 r       = A^-1 r + (cw/s) Z W^-1 Z'r
A^-1 B comes from the Lanczos process of K (columns in lock-step) with the pentadiagonal Galerkin system
cw (T^2 + beta_m^2 e_m e_m') + s I; x is refined on the true residual evaluated through the entrywise
class-centred matrix Hc = (I-P) K (klda.jl:49, center_rows!).
"""
import numpy as np


def penta_solve(al, be, m, cw, shift, rhs0):
    """z of the Galerkin system of dimension m and the exact residual norm of Q z (needs al[m], be[m])."""
    d0 = np.zeros(m); d1 = np.zeros(m); d2 = np.zeros(m)
    for i in range(m):
        bl = be[i - 1] if i > 0 else 0.0
        br = be[i] if i < m - 1 else 0.0
        d0[i] = cw * (al[i] ** 2 + bl ** 2 + br ** 2) + shift
        if i + 1 < m:
            d1[i] = cw * be[i] * (al[i] + al[i + 1])
        if i + 2 < m:
            d2[i] = cw * be[i] * be[i + 1]
    d0[m - 1] += cw * be[m - 1] ** 2
    D = np.zeros(m); L1 = np.zeros(m); L2 = np.zeros(m)
    for i in range(m):
        v = d0[i]
        if i >= 1:
            v -= L1[i - 1] ** 2 * D[i - 1]
        if i >= 2:
            v -= L2[i - 2] ** 2 * D[i - 2]
        D[i] = v
        if i + 1 < m:
            t = d1[i]
            if i >= 1:
                t -= L1[i - 1] * L2[i - 1] * D[i - 1]
            L1[i] = t / v
        if i + 2 < m:
            L2[i] = d2[i] / v
    y = np.zeros(m); y[0] = rhs0
    for i in range(1, m):
        y[i] = -L1[i - 1] * y[i - 1] - (L2[i - 2] * y[i - 2] if i >= 2 else 0.0)
    y /= D
    for i in range(m - 2, -1, -1):
        y[i] -= L1[i] * y[i + 1] + (L2[i] * y[i + 2] if i + 2 < m else 0.0)
    tz = (be[m - 2] * y[m - 2] if m >= 2 else 0.0) + al[m - 1] * y[m - 1]
    est = cw * be[m - 1] * np.hypot(tz + al[m] * y[m - 1], be[m] * y[m - 1])
    return y, est


def solve_A(K, B, cw, shift, tol, maxit=4000):
    """Y = (cw K^2 + shift I)^-1 B, every column to relative residual tol; returns (Y, passes, estimates)."""
    n, c = B.shape
    nrm = np.linalg.norm(B, axis=0)
    V = [B / nrm]
    al = [[] for _ in range(c)]; be = [[] for _ in range(c)]
    vprev = np.zeros_like(B); bprev = np.zeros(c)
    mcol = [0] * c; zcol = [None] * c; ests = [None] * c
    for step in range(maxit):
        w = K @ V[-1] - vprev * bprev
        a = (V[-1] * w).sum(0); w -= V[-1] * a
        b = np.linalg.norm(w, axis=0)
        for col in range(c):
            al[col].append(a[col]); be[col].append(b[col])
        vprev = V[-1]; bprev = b
        V.append(w / b)
        m = step
        if m < 1:
            continue
        for col in range(c):
            if mcol[col]:
                continue
            z, est = penta_solve(al[col], be[col], m, cw, shift, nrm[col])
            if est <= tol * nrm[col]:
                mcol[col] = m; zcol[col] = z; ests[col] = est / nrm[col]
        if all(mcol):
            Y = np.zeros_like(B)
            for col in range(c):
                for i in range(mcol[col]):
                    Y[:, col] += zcol[col][i] * V[i][:, col]
            return Y, step + 1, ests
    raise RuntimeError("Lanczos solve did not converge")


def klda_solve(K, y0, c, freq, cw, shift, cg_tol=1e-10, max_ref=6):
    """x = Sigma_w'^-1 H_b' as klda.cu computes it; returns (x, passes, residual history)."""
    n = K.shape[0]
    E = np.zeros((n, c)); E[np.arange(n), y0] = 1.0
    cnt = E.sum(0)
    g = np.sqrt(freq)[None, :] * (np.diag(1.0 / cnt) - (freq / cnt)[:, None])
    h = K @ (E @ g)
    Hc = K - E @ ((E.T @ K) / cnt[:, None])          # entrywise class centring (center_rows!)

    def apply(x):
        t = Hc @ x
        t -= E @ ((E.T @ t) / cnt[:, None])
        return cw * (K @ t) + shift * x
    Y, passes, _ = solve_A(K, E, cw, shift, min(1e-13, cg_tol))
    Z = K @ Y; passes += 1
    W = E.T @ Y; W = 0.5 * (W + W.T)
    x = Z @ np.linalg.solve(W, np.diag(cnt) @ g) / shift
    hn = np.linalg.norm(h, axis=0)
    hist = []; prev = -1.0
    for it in range(max_ref + 1):
        r = h - apply(x); passes += 2
        rel = (np.linalg.norm(r, axis=0) / hn).max(); hist.append(rel)
        if rel <= cg_tol or (prev >= 0 and rel > 0.03 * prev) or it == max_ref:
            break
        prev = rel
        tol_in = min(1e-2, max(1e-14, 0.05 * cg_tol / rel))
        dlt, m, _ = solve_A(K, r, cw, shift, tol_in); passes += m
        x = x + dlt + Z @ np.linalg.solve(W, Z.T @ r) * (cw / shift)
    return x, h, passes, hist
