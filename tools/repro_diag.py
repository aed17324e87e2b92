"""Diagnostics of the rare multi-GPU run-to-run difference (torchrun): F fits of the same X, the row sums
and per-tile partial row sums of every fit (FXB200_DEBUG_DUMP) compared with the first fit's.
Usage: torchrun ... tools/repro_diag.py <n> <d> <k> <fits>"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dump = tempfile.mkdtemp(prefix="fxdump_")
os.environ["FXB200_DEBUG_DUMP"] = dump
import bench  # noqa: E402


def main():
    n, d, k, fits = (int(x) for x in sys.argv[1:5])
    h = bench.Harness()
    fxb200 = h.fx
    from fxb200.synth import mix
    X, _ = mix(n, d, k + 1, 3.0, 3002)
    kern = fxb200.GaussianKernel(1.0 / d)
    ld = fxb200.sharding.row_partition(n, h.world)[0]
    rs0 = rp0 = L0 = None
    events = 0
    for f in range(fits):
        V, L, st = fxb200.kpca(X, kern, k=k, ctx=h.ctx, return_stats=True)
        rs = np.fromfile(os.path.join(dump, "rowsum_r%d.bin" % h.rank))
        rp = np.fromfile(os.path.join(dump, "rowpart_r%d.bin" % h.rank)).reshape(-1, ld)
        if rs0 is None:
            rs0, rp0, L0 = rs, rp, L.copy()
            continue
        drs = np.nonzero(rs != rs0)[0]
        drp = np.argwhere(rp != rp0)
        lam_diff = np.abs(L / L0 - 1).max()
        if len(drs) or len(drp) or lam_diff > 0:
            events += 1
            if events <= 3:
                msg = ["rank %d fit %d: lambda rel diff %.2e, %d row sums differ, %d partials of this rank differ" %
                       (h.rank, f, lam_diff, len(drs), len(drp))]
                if len(drs):
                    owner = drs // ld
                    msg.append("  rows (global) %s ... owners %s; diffs %s" %
                               (drs[:12].tolist(), np.unique(owner).tolist(), (rs[drs] - rs0[drs])[:6].tolist()))
                if len(drp):
                    slots = np.unique(drp[:, 0]); rows = np.unique(drp[:, 1])
                    msg.append("  partials: slots %s (column blocks %s) rows %d..%d (%d rows)" %
                               (slots[:16].tolist(), np.unique(slots * 64 // ld).tolist(), rows.min(), rows.max(), len(rows)))
                    for (sl, r) in drp[:6]:
                        msg.append("    slot %d row %d: now %.17g first %.17g" % (sl, r, rp[sl, r], rp0[sl, r]))
                print("\n".join(msg), flush=True)
    t = h.torch.tensor([float(events)], dtype=h.torch.float64, device="cuda")
    if h.dist is not None:
        h.dist.all_reduce(t)
    if h.rank == 0:
        print("DIAG world=%d n=%d fits=%d: events summed over ranks %d" % (h.world, n, fits, int(t.item())), flush=True)
    h.close()


if __name__ == "__main__":
    main()
