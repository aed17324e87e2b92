"""world_size-2 worker (gloo, CPU): runs the row-sharded eigensolver algorithm (the NumPy model
of csrc/solver.cu) with torch.distributed collectives in exactly the places the CUDA path uses
NCCL -- all-gather of the n x b block, all-reduce of the small projection matrices, all-gather
of the centring row sums -- and checks it against the dense oracle.  Also exercises the
unique-id style broadcast bench.py uses.  Launched by tests/test_distributed_gloo.py."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle as O                      # noqa: E402
import fxb200                           # noqa: E402
from fxb200.synth import mix            # noqa: E402
from fxb200.sharding import row_partition, query_partition   # noqa: E402
from proto_krylov import topk           # noqa: E402


class GlooOps:
    def __init__(self, ld, world):
        self.ld, self.world = ld, world

    def allreduce(self, a):
        t = torch.from_numpy(np.ascontiguousarray(a))
        dist.all_reduce(t)
        return t.numpy()

    def allgather_rows(self, a):
        # every rank sends its padded (ld x b) block; result is (world*ld) x b like the NCCL path
        pad = np.zeros((self.ld, a.shape[1]))
        pad[:a.shape[0]] = a
        outs = [torch.zeros_like(torch.from_numpy(pad)) for _ in range(self.world)]
        dist.all_gather(outs, torch.from_numpy(pad))
        return np.concatenate([o.numpy() for o in outs], axis=0)


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    # the 128-byte communicator id travels like this in bench.py (rank 0 -> all)
    ids = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    assert ids[0] == bytes(range(128))

    n, d, k = 700, 12, 4
    X, _ = mix(n, d, k + 1, 3.0, 51)
    kern = O.GaussianKernel(1.0 / d)
    ld, n_ld, parts = row_partition(n, world)
    row0, rows = parts[rank]
    # each rank builds only its row block of K (Gram shards by rows, no collective)
    Kloc = O.kernelmatrix(kern, X[row0:row0 + rows], X)          # rows x n
    # centring: local row sums, all-gather, then the local update
    rs = np.zeros(ld); rs[:rows] = Kloc.sum(axis=1)
    outs = [torch.zeros(ld, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(outs, torch.from_numpy(rs))
    rowsum = np.concatenate([o.numpy() for o in outs])[:n] if n_ld == n else \
        np.concatenate([o.numpy()[:parts[r][1]] for r, o in enumerate(outs)])
    colmean = rowsum / n
    mu = colmean.sum() / n
    Kc_loc = Kloc - colmean[None, :] - colmean[row0:row0 + rows, None] + mu
    ops = GlooOps(ld, world)

    def matvec(Vfull):
        # Vfull is the gathered (world*ld) x b block; real rows of rank r sit at [r*ld, r*ld+rows_r)
        V = np.concatenate([Vfull[r * ld:r * ld + parts[r][1]] for r in range(world)], axis=0)
        return Kc_loc @ V

    V_loc, theta, info = topk(matvec, rows, row0, k, ops=ops)
    assert info["converged"]
    Vo, Lo = O.kpca(X, kern)
    np.testing.assert_allclose(theta / (n - 1), Lo[:k], rtol=1e-10)
    full = ops.allgather_rows(V_loc)
    V = np.concatenate([full[r * ld:r * ld + parts[r][1]] for r in range(world)], axis=0)
    Qa, _ = np.linalg.qr(V)
    ang = np.linalg.norm(Vo[:, :k] - Qa @ (Qa.T @ Vo[:, :k]), 2)
    assert ang < 1e-8, ang
    # transform shards queries, no collective
    qp = query_partition(101, world)
    assert sum(c for _, c in qp) == 101
    dist.barrier()
    if rank == 0:
        print("DIST_OK steps=%d angle=%.2e" % (info["steps"], ang))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
