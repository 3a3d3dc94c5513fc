import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vinsat_b200 import _lib, longarc, synth
ctx = _lib.Context(0)
pr = synth.make_problem(77, 120, 6)
b = _lib.Batch(ctx, _lib.concat_problems([pr]))
lam, ntr = b.ba_iterate(10, 1e-4, initialize=False)
dbg = b.debug_fetch()
print("whole: ntr", ntr, "lam", lam, "dpose max", np.abs(dbg["dpose"]).max())
la = longarc.LongArc(pr, ctxs=[ctx], world=1, n_segments=int(sys.argv[1]) if len(sys.argv) > 1 else 1)
import math
L = longarc
it = 10; alpha = min(max(1 - (2 * (it / 5) - 1), 1), 2); Sigma = float(min(10000 * (it + 1) ** 2, 1000000))
la._stage(L.RESID); la._stage(L.SELECT_BEGIN, 2 * la.M)
for ps in range(6):
    la._stage(L.SELECT_HIST, ps); la._all_reduce(L.BUF_HIST, "sum"); la._stage(L.SELECT_PICK, ps)
la._stage(L.ASSEMBLE, d0=alpha); la._stage(L.DYNAMICS, 0); la._stage(L.SYSTEM, 0, 0, Sigma)
la._stage(L.SUMS_INIT, 0); print("sums init", la._global_sums())
la._stage(L.SET_LAM, d0=1e-4); la._stage(L.FORWARD); la._all_gather(L.BUF_PACK, L.BUF_GATHER)
la._stage(L.REDUCED, per_rank_i0=lambda p: p.rank * la.S); la._stage(L.BACKSUB)
d2 = la.parts[0].batch.debug_fetch()
print("la dpose max", np.abs(d2["dpose"]).max(), "diff", np.abs(d2["dpose"] - dbg["dpose"]).max())
print("D diff", np.abs(d2["D"] - dbg["D"]).max(), "U diff", np.abs(d2["U"] - dbg["U"]).max(), "rhs diff", np.abs(d2["rhs"] - dbg["rhs"]).max())
print("c_obs", d2["c_obs"], dbg["c_obs"], "w diff", np.abs(d2["weights"] - dbg["weights"]).max())
bad = np.abs(d2["dpose"] - dbg["dpose"]).max(axis=1)
print("per-frame diff (first 12, last 6):", bad[:12], bad[-6:])
