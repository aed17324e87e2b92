"""GPU: row-sharded fit over NCCL with one rank per GPU (needs >= 2 GPUs; skipped otherwise)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world, port, env_extra=None):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs (covered on CPU by test_distributed_gloo.py)" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "dist_gpu_worker.py")]
    env = dict(os.environ)
    env.update(env_extra or {})
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("DIST_GPU_OK") >= 5
    return r.stdout


# The default schedule is rank-local (each rank computes its whole row block; no peer stores).
@pytest.mark.parametrize("world", [2, 3, 4])
def test_sharded_fit_rank_local_gram(world):
    out = _run(world, 29640 + world)
    assert "peer=1" not in out
    print(out)


# Opt-in cross-rank symmetric Gram (FXB200_PEER_SYM=1; see DESIGN.md section 7 for why it is not the
# default).  world 2: the two blocks are "opposite" (checkerboard split of the off-diagonal block);
# world 3: odd ring, every off-diagonal block has one owner; world 4: both kinds at once
@pytest.mark.parametrize("world", [2, 3, 4])
def test_sharded_fit_peer_symmetric_gram(world):
    out = _run(world, 29650 + world, {"FXB200_PEER_SYM": "1"})
    assert "peer=1" in out
    print(out)


def test_peer_switch_off_wins():
    """FXB200_NO_PEER=1 overrides the opt-in (the path taken when CUDA IPC / NVLink peer access is
    unavailable); same results."""
    out = _run(2, 29649, {"FXB200_NO_PEER": "1", "FXB200_PEER_SYM": "1"})
    assert "peer=1" not in out


@pytest.mark.parametrize("ndev", [2, 4])
def test_single_process_multi_gpu_context(ndev):
    """fx_init_multi: ONE process drives `ndev` GPUs -- one host thread per device
    inside the library, in-process NCCL communicators, direct peer access.  The caller makes one call
    per fit / transform with host buffers.  Results: against the oracle, and eigenvalues identical to
    the one-GPU context of the same process."""
    import numpy as np
    import torch
    import oracle as O
    import fxb200
    from fxb200.synth import mix
    if torch.cuda.device_count() < ndev:
        pytest.skip("needs %d GPUs" % ndev)
    ctx = fxb200.Context.multi(list(range(ndev)))
    one = fxb200.Context(0)
    assert ctx.world == ndev
    for (n, d, k) in [(4096, 64, 8), (1500, 20, 4), (1000, 10, 2)]:
        X, _ = mix(n, d, k + 1, 3.0, 61)
        kern, okern = fxb200.GaussianKernel(1.0 / d), O.GaussianKernel(1.0 / d)
        V, L, model, st = fxb200.kpca(X, kern, k=k, ctx=ctx, return_model=True, return_stats=True)
        V1, L1 = fxb200.kpca(X, kern, k=k, ctx=one)
        Vo, Lo = O.kpca(X, okern)
        np.testing.assert_allclose(L, Lo[:k], rtol=1e-10)
        np.testing.assert_allclose(L, L1, rtol=1e-13)
        Qa, _ = np.linalg.qr(V)
        assert np.linalg.norm(Vo[:, :k] - Qa @ (Qa.T @ Vo[:, :k]), 2) < 1e-8
        Z, _ = mix(1001, d, k + 1, 3.0, 62)
        T = model.transform(Z)                               # queries sharded over the devices
        To = O.transform_kpca(X, okern, V, Z)
        assert np.abs(T - To).max() <= 1e-10 * np.abs(To).max()
        model.free()
        print("MULTI_OK ndev=%d n=%d steps=%d peer=%d" % (ndev, n, st["steps"], ctx.peer_access))
    # KLDA through the same context (iterative path: the dense one is single-GPU)
    n, d, c = 1500, 32, 5
    X, y0 = mix(n, d, c, 2.0, 63)
    y = y0 + 1
    freq = np.bincount(y0, minlength=c) / n
    kern, okern = fxb200.GaussianKernel(1.0 / d), O.GaussianKernel(1.0 / d)
    H_b, H_w, _ = O.lda_matrices(O.kernelmatrix(okern, X), y, c, freq)
    eps_w = 1e-6 * np.trace(O.scatter_matrices(H_b, 0, 0, H_w, 0, 0)[1]) / n
    Ao, Lo = O.klda(X, okern, y, c, freq, eps_w=eps_w)
    A, L = fxb200.klda(X, kern, y, c, freq, eps_w=eps_w, ctx=ctx)
    np.testing.assert_allclose(L, Lo, rtol=1e-7)
    # building blocks on a multi-GPU context run on its first device
    K = fxb200.kernelmatrix(kern, X[:300], ctx=ctx)
    assert np.abs(K - O.kernelmatrix(okern, X[:300])).max() < 1e-13
    ctx.close(); one.close()
