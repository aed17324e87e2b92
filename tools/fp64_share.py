"""Do DMMA and DFMA share the FP64 datapath?  Times 4 DMMA warps, 4 DFMA warps (equal lane-cycles)
and both together on every SM."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fxb200
t = [float(np.min(fxb200.bench_kernel(8, d=m, reps=5))) for m in (0, 1, 2)]
print("dmma_only_ms %.4f dfma_only_ms %.4f both_ms %.4f -> both/(dmma+dfma) = %.3f, both/max = %.3f"
      % (t[0], t[1], t[2], t[2] / (t[0] + t[1]), t[2] / max(t[0], t[1])))
