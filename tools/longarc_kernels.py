"""Per-kernel durations of ONE full-phase BA iteration of a whole long arc (run under
`ncu --profile-from-start off --metrics gpu__time_duration.sum --csv`).  python tools/longarc_kernels.py T K gap_max"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from vinsat_b200 import _lib, synth

T, K, gap_max = (int(a) for a in sys.argv[1:4])
gen = _lib.Context(0)
pr = synth.make_problem(123, T, K, gap_max=gap_max, orbit_fn=lambda x0, n: gen.orbit_propagate(x0, n, 1, 1.0))
gen.close()
ctx = _lib.Context(0)
b = _lib.Batch(ctx, _lib.concat_problems([pr]))
lam = np.array([1e-4])
for it in range(13):
    if it == 12:
        ctx.synchronize(); torch.cuda.cudart().cudaProfilerStart()
    lam, ntr = b.ba_iterate(it, lam, initialize=it < 10)
ctx.synchronize(); torch.cuda.cudart().cudaProfilerStop()
print("ntrials", ntr)
