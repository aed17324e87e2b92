"""Run-to-run determinism check of the multi-GPU fit (torchrun): F fits of the same X through the host
API, every rank's lambda / trace compared BITWISE with its first fit and across ranks.
Usage: torchrun ... tools/repro_determinism.py <n> <d> <k> <fits> [kernel: gaussian|poly]"""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    n, d, k, fits = (int(x) for x in sys.argv[1:5])
    kname = sys.argv[5] if len(sys.argv) > 5 else "gaussian"
    h = bench.Harness()
    fxb200 = h.fx
    from fxb200.synth import mix
    X, _ = mix(n, d, k + 1, 3.0, 3002)
    kern = fxb200.GaussianKernel(1.0 / d) if kname == "gaussian" else fxb200.PolynomialKernel(1.0 / d, 1.0, 3.0)
    first = None
    bad = 0
    for f in range(fits):
        try:
            V, L, st = fxb200.kpca(X, kern, k=k, ctx=h.ctx, return_stats=True)
        except fxb200.FxError as e:                 # e.g. not converged on a corrupted K: an event, keep going
            bad += 1
            print("rank %d fit %d FAILED: %s" % (h.rank, f, str(e)[:120]), flush=True)
            continue
        key = (hashlib.sha1(L.tobytes()).hexdigest()[:12], repr(st["trace"]))
        if first is None:
            first = (key, L.copy())
        elif key != first[0]:
            bad += 1
            print("rank %d fit %d DIFFERS: lambda max rel diff %.3e  trace %r vs %r" %
                  (h.rank, f, np.abs(L / first[1] - 1).max(), st["trace"], first[0][1]), flush=True)
    t = h.torch.tensor([float(bad), first[1][0], first[1][-1]], dtype=h.torch.float64, device="cuda")
    if h.dist is not None:
        g = [h.torch.zeros_like(t) for _ in range(h.world)]
        h.dist.all_gather(g, t)
    else:
        g = [t]
    if h.rank == 0:
        tot = sum(int(x[0].item()) for x in g)
        same = all(x[1].item() == g[0][1].item() and x[2].item() == g[0][2].item() for x in g)
        print("DETERMINISM world=%d n=%d d=%d k=%d fits=%d env=%s: %d differing fits over all ranks, first-fit lambda "
              "identical across ranks: %s, lambda1=%.17g steps=%d" %
              (h.world, n, d, k, fits, {e: os.environ[e] for e in os.environ if e.startswith("FXB200_")}, tot, same,
               first[1][0], st["steps"]), flush=True)
    h.close()


if __name__ == "__main__":
    main()
