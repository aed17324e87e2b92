"""The host side of the iterative KLDA solve (csrc/klda_host.hpp: pentadiagonal Galerkin system with its
residual formula, small SPD solve) compiled with g++ from the very source klda.cu includes and compared
with NumPy -- no GPU needed."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from proto_klda_lanczos import penta_solve

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("g++ not available")
    out = str(tmp_path_factory.mktemp("kldahost") / "klda_host")
    subprocess.run([gxx, "-std=c++17", "-O2", "-o", out, os.path.join(ROOT, "tests", "host", "klda_host_main.cpp")],
                   check=True, capture_output=True, text=True)
    return out


def _run(exe, text):
    r = subprocess.run([exe], input=text, capture_output=True, text=True, check=True)
    return [np.array([float(v) for v in line.split()]) for line in r.stdout.strip().splitlines()]


@pytest.mark.parametrize("m,cw,shift", [(1, 1e-6, 1e-13), (2, 1.0, 0.5), (7, 3e-8, 2e-14), (60, 1e-7, 1e-15), (300, 1e-9, 1e-18)])
def test_galerkin_pentadiagonal_solve(exe, m, cw, shift):
    rng = np.random.default_rng(m)
    al = np.abs(rng.normal(2.0, 1.0, m + 1)) + 0.1
    be = np.abs(rng.normal(0.5, 0.3, m + 1)) + 1e-3
    rhs0 = 3.7
    txt = "penta %d %.17g %.17g %.17g\n%s\n%s\n" % (m, cw, shift, rhs0, " ".join("%.17g" % v for v in al),
                                                   " ".join("%.17g" % v for v in be))
    out = _run(exe, txt)[0]
    est, z = out[0], out[1:]
    # dense reference of the same Galerkin system
    T = np.diag(al[:m]) + np.diag(be[:m - 1], 1) + np.diag(be[:m - 1], -1)
    M = cw * (T @ T)
    M[m - 1, m - 1] += cw * be[m - 1] ** 2
    M += shift * np.eye(m)
    rhs = np.zeros(m); rhs[0] = rhs0
    zd = np.linalg.solve(M, rhs)
    assert np.abs(z - zd).max() <= 1e-9 * np.abs(zd).max()
    zp, estp = penta_solve(list(al), list(be), m, cw, shift, rhs0)      # the NumPy model of klda.cu
    assert np.abs(z - zp).max() <= 1e-10 * np.abs(zp).max() and abs(est - estp) <= 1e-9 * abs(estp) + 1e-300
    # the residual formula: extend T by one row/column and form the true residual of Q z in the (m+2)-dim basis
    Te = np.zeros((m + 2, m + 2))
    Te[:m + 1, :m + 1] = np.diag(al[:m + 1]) + np.diag(be[:m], 1) + np.diag(be[:m], -1)
    Te[m, m + 1] = Te[m + 1, m] = be[m]
    ze = np.zeros(m + 2); ze[:m] = z
    r = -(cw * (Te @ (Te @ ze)) + shift * ze)
    r[0] += rhs0
    r[m + 1] = -cw * be[m] * (Te @ ze)[m]                               # q_{m+2} component (row m+1 of Te^2 needs alpha_{m+2}: only its beta part counts)
    true = np.linalg.norm(np.concatenate([r[m:m + 1], r[m + 1:m + 2]]))
    assert np.abs(r[:m]).max() <= 1e-7 * rhs0                           # Galerkin condition
    assert abs(true - est) <= 1e-6 * est + 1e-25 * rhs0


@pytest.mark.parametrize("c,nb", [(1, 1), (3, 2), (10, 10), (64, 5)])
def test_small_spd_solve(exe, c, nb):
    rng = np.random.default_rng(c * 7 + nb)
    A = rng.normal(size=(c, c + 3))
    W = A @ A.T + 1e-3 * np.eye(c)
    B = rng.normal(size=(c, nb))
    txt = "spd %d %d\n%s\n%s\n" % (c, nb, " ".join("%.17g" % v for v in W.flatten(order="F")),
                                   " ".join("%.17g" % v for v in B.flatten(order="F")))
    out = _run(exe, txt)[0]
    assert out[0] == 1
    X = out[1:].reshape((c, nb), order="F")
    assert np.abs(W @ X - B).max() <= 1e-9 * np.abs(B).max() * np.linalg.cond(W)
    assert np.abs(X - np.linalg.solve(W, B)).max() <= 1e-10 * np.abs(X).max() * max(1.0, np.linalg.cond(W) * 1e-4)


def test_small_spd_solve_rejects_indefinite(exe):
    W = np.array([[1.0, 2.0], [2.0, 1.0]])
    txt = "spd 2 1\n%s\n1 1\n" % " ".join("%.17g" % v for v in W.flatten(order="F"))
    assert _run(exe, txt)[0][0] == 0
