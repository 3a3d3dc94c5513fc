"""Per-kernel-family time of the long-arc BA: whole arc in one batch against the frame-window sharded driver with its windows
emulated on ONE GPU (no NCCL).  python tools/longarc_profile.py T K gap_max world  ->  one JSON record."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from vinsat_b200 import _lib, longarc, synth

T, K, gap_max, world = (int(a) for a in sys.argv[1:5])
gen = _lib.Context(0)
pr = synth.make_problem(123, T, K, gap_max=gap_max, orbit_fn=lambda x0, n: gen.orbit_propagate(x0, n, 1, 1.0))
gen.close()
out = {"T": T, "K": K, "gap_max": gap_max, "world_emulated": world}


def fam(ctx):
    return {k: (round(v[0], 3), v[1]) for k, v in ctx.timing().items() if v[1]}


ctx = _lib.Context(0)
b = _lib.Batch(ctx, _lib.concat_problems([pr]))
lam = np.array([1e-4])
for it in range(20):
    if it == 10:
        out["whole_init_phase"] = fam(ctx); ctx.reset_timing()
    if it == 9:
        ctx.enable_timing(True); ctx.reset_timing()
    ctx.synchronize(); t0 = time.perf_counter()
    lam, ntr = b.ba_iterate(it, lam, initialize=it < 10)
    ctx.synchronize()
    if it in (9, 19):
        out["whole_ms_it%d" % it] = 1e3 * (time.perf_counter() - t0)
out["whole_full_phase_10_iterations"] = fam(ctx)
out["whole_err_km"] = float(np.abs(b.get_states()[:, :3] - pr["states_gt"][:, :3]).max())
b.close(); ctx.close()

ctxs = [_lib.Context(0) for _ in range(world)]
la = longarc.LongArc(pr, ctxs=ctxs, world=world)
lam = 1e-4
for it in range(20):
    if it == 10:
        for c in ctxs:
            c.enable_timing(True); c.reset_timing()
    lam, ntr = la.ba_iterate(it, lam, initialize=it < 10)
out["sharded_full_phase_10_iterations_per_part"] = [fam(c) for c in ctxs]
out["sharded_err_km"] = float(np.abs(la.gather_states()[:, :3] - pr["states_gt"][:, :3]).max())
out["S_per_rank"] = la.S
la.close()
print(json.dumps(out))
