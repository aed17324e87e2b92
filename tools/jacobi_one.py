import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fxb200
m = int(sys.argv[1]) if len(sys.argv) > 1 else 80
rng = np.random.default_rng(0)
A = rng.standard_normal((m, m)); A = A + A.T
w, V, sw = fxb200.syevd_small(A)
print("m", m, "sweeps", sw)
