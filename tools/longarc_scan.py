"""Whole-arc BA on one GPU for growing arc lengths: where does the solve stop converging, and what does an iteration cost?
python tools/longarc_scan.py K T1 T2 ...   (env VINSAT_SEG_LEN etc. apply).  One JSON line per T."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from vinsat_b200 import _lib, synth

K = int(sys.argv[1])
gen = _lib.Context(0)
orbit_fn = (lambda x0, n: gen.orbit_propagate(x0, n, 1, 1.0)) if not os.environ.get("SCAN_HOST_ORBIT") else None
for T in [int(a) for a in sys.argv[2:]]:
    pr = synth.make_problem(123, T, K, gap_max=3, orbit_fn=orbit_fn)
    ctx = _lib.Context(0)
    b = _lib.Batch(ctx, _lib.concat_problems([pr]))
    lam = np.array([1e-4]); hist = []; tms = []
    for it in range(20):
        ctx.synchronize(); t0 = time.perf_counter()
        lam, ntr = b.ba_iterate(it, lam, initialize=it < 10)
        ctx.synchronize(); tms.append(1e3 * (time.perf_counter() - t0))
        st = b.get_states()
        hist.append((float(lam[0]), int(ntr[0]), float(np.abs(st[:, :3] - pr["states_gt"][:, :3]).max())))
    b.close(); ctx.close()
    print(json.dumps({"T": T, "K": K, "ms_init_median": float(np.median(tms[:10])), "ms_full_median": float(np.median(tms[10:])),
                      "final_max_pos_err_km": hist[-1][2], "history_lam_ntr_err": hist}), flush=True)
