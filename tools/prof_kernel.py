"""Runs one kernel class a few times on device-resident inputs -- the short command that
ncu wraps.  Usage: python tools/prof_kernel.py <which> <n> <d> <b> <m> [kernel_id] [reps]
which: 0 gram, 1 center, 2 symm, 3 transform, 4 dmma peak, 5 dfma peak, 6 copy, 7 dmma sweep"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fxb200  # noqa: E402

a = [int(x) for x in sys.argv[1:]]
which, n, d, b, m = (a + [0] * 5)[:5]
kid = a[5] if len(a) > 5 else 0
reps = a[6] if len(a) > 6 else 2
ms = fxb200.bench_kernel(which, n=n, d=d, b=b, m=m, kernel_id=kid, reps=reps)
print("which=%d n=%d d=%d b=%d m=%d kernel=%d ms=%s" % (which, n, d, b, m, kid, np.round(ms, 4).tolist()))
