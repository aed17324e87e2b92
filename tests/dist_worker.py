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


def peer_symmetric_gram(kern, X, n, world, rank):
    """The cross-rank symmetric Gram of csrc/gram.cu (sym = 2) with the peer stores emulated by an
    all-reduce: every rank computes only the 128x64 tiles sharding.gram_tile_roles gives it and
    'stores' each mirrored tile transposed into the owning rank's rows; the column sums of a
    mirrored tile are the owner's row-sum partials.  Returns this rank's padded row block
    [ld][n_ld] and its row sums [ld]; asserts every element was written exactly once."""
    from fxb200.sharding import gram_tile_roles
    ld, n_ld, parts = row_partition(n, world)
    Xp = np.zeros((n_ld, X.shape[1])); Xp[:n] = X
    valid = np.arange(n_ld) < n
    role, target = gram_tile_roles(n, world, rank, peer=True)
    writes = np.zeros((n_ld, n_ld)); count = np.zeros((n_ld, n_ld)); rsum = np.zeros(n_ld)
    r0 = rank * ld
    for ty in range(role.shape[0]):
        for tx in range(role.shape[1]):
            if role[ty, tx] == 0:
                continue
            gi, gj = r0 + 128 * ty, 64 * tx
            tile = O.kernelmatrix(kern, Xp[gi:gi + 128], Xp[gj:gj + 64])
            tile = tile * valid[gi:gi + 128, None] * valid[None, gj:gj + 64]     # padding written as 0
            writes[gi:gi + 128, gj:gj + 64] += tile; count[gi:gi + 128, gj:gj + 64] += 1
            rsum[gi:gi + 128] += tile.sum(axis=1)
            if role[ty, tx] == 2:
                assert gj // ld == target[ty, tx]
                writes[gj:gj + 64, gi:gi + 128] += tile.T; count[gj:gj + 64, gi:gi + 128] += 1
                rsum[gj:gj + 64] += tile.sum(axis=0)
    for a in (writes, count, rsum):
        t = torch.from_numpy(a)
        dist.all_reduce(t)
    assert (count == 1).all(), "a tile was produced %s times" % np.unique(count)
    return writes[r0:r0 + ld], rsum[r0:r0 + ld]


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
    # the same block through the cross-rank symmetric ownership (mirrored tiles 'stored' into peers)
    Ksym, rs_sym = peer_symmetric_gram(kern, X, n, world, rank)
    assert np.array_equal(Ksym[rows:], np.zeros((ld - rows, n_ld))) and not Ksym[:, n:].any()
    np.testing.assert_allclose(Ksym[:rows, :n], Kloc, rtol=0, atol=1e-14)
    np.testing.assert_allclose(rs_sym[:rows], Kloc.sum(axis=1), rtol=1e-13)
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
