"""One rank per GPU (NCCL): row-sharded KPCA fit through the C ABI, checked against the oracle
and against the sharding-independent properties.  Launched by tests/test_gpu_distributed.py or
by hand: torchrun --nproc-per-node N tests/dist_gpu_worker.py [n d k]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O                      # noqa: E402
import fxb200                           # noqa: E402
from fxb200.synth import mix            # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    ids = [fxb200.Context.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx = fxb200.Context.distributed(local, rank, world, ids[0])
    assert ctx.world == world and ctx.rank == rank

    args = [int(a) for a in sys.argv[1:]]
    cases = [tuple(args[:3])] if len(args) >= 3 else [(1500, 20, 4), (1000, 10, 2), (200, 6, 3), (4096, 64, 8)]
    for (n, d, k) in cases:
        X, _ = mix(n, d, k + 1, 3.0, 61)
        kern, okern = fxb200.GaussianKernel(1.0 / d), O.GaussianKernel(1.0 / d)
        V, L, model, st = fxb200.kpca(X, kern, k=k, ctx=ctx, return_model=True, return_stats=True)
        Vo, Lo = O.kpca(X, okern)
        np.testing.assert_allclose(L, Lo[:k], rtol=1e-10)
        Qa, _ = np.linalg.qr(V)
        ang = np.linalg.norm(Vo[:, :k] - Qa @ (Qa.T @ Vo[:, :k]), 2)
        assert ang < 1e-8, ang
        # every rank must hold the identical result
        t = torch.from_numpy(np.ascontiguousarray(L)).cuda()
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi)
        # transform: queries sharded across ranks, no collective
        Z, _ = mix(301, d, k + 1, 3.0, 62)
        q0, cnt = fxb200.sharding.query_partition(len(Z), world)[rank]
        T = model.transform(Z[q0:q0 + cnt]) if cnt else np.zeros((0, k))
        To = O.transform_kpca(X, okern, V, Z)[q0:q0 + cnt]
        if cnt:
            assert np.abs(T - To).max() <= 1e-10 * np.abs(To).max()
        model.free()
        if rank == 0:
            print("DIST_GPU_OK world=%d n=%d k=%d steps=%d dense=%d comm_ms=%.3f angle=%.1e peer=%d" %
                  (world, n, k, st["steps"], st["dense_path"], st["comm_ms"], ang, ctx.peer_access), flush=True)
    # KLDA on the same sharding (indicator K*V, class projector, CG all row-sharded)
    if len(args) < 3:
        n, d, c = 1500, 32, 5
        X, y0 = mix(n, d, c, 2.0, 63)
        y = y0 + 1
        freq = np.bincount(y0, minlength=c) / n
        kern, okern = fxb200.GaussianKernel(1.0 / d), O.GaussianKernel(1.0 / d)
        K = O.kernelmatrix(okern, X)
        H_b, H_w, _ = O.lda_matrices(K, y, c, freq)
        _Sb, Sw = O.scatter_matrices(H_b, 0, 0, H_w, 0, 0)
        eps_w = 1e-6 * np.trace(Sw) / n
        Ao, Lo = O.klda(X, okern, y, c, freq, eps_w=eps_w)
        A, L, st = fxb200.klda(X, kern, y, c, freq, eps_w=eps_w, ctx=ctx, return_stats=True)
        np.testing.assert_allclose(L, Lo, rtol=1e-7)
        Qa, _ = np.linalg.qr(A)
        assert np.linalg.norm(Ao - Qa @ (Qa.T @ Ao), 2) / np.linalg.norm(Ao, 2) < 1e-5
        if rank == 0:
            print("DIST_GPU_OK klda world=%d n=%d c=%d cg_iters=%d comm_ms=%.2f" %
                  (world, n, c, st["cg_iters"], st["comm_ms"]), flush=True)
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
