"""How much do independent resident solves overlap on one GPU?  depth batches of P problems, each driven by its own
host thread / context / stream; reports solves/s for depth = 1, 2, 3, 4 at a fixed total of problems in flight."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vinsat_b200 import _lib, synth

Ptot = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
prs = synth.make_batch(Ptot, 1000, 10, seed0=0)
for depth in (1, 2, 4):
    Pb = Ptot // depth
    slots = []
    for d in range(depth):
        ctx = _lib.Context(0)
        arr = _lib.concat_problems(prs[d * Pb:(d + 1) * Pb])
        slots.append((ctx, _lib.Batch(ctx, arr), arr))
    def work(s, reps):
        ctx, b, arr = slots[s]
        for _ in range(reps):
            b.set_states(arr["states"])
            b.od_solve(20, 10, 1e-4)
        ctx.synchronize()
    for reps in (1, 4):
        ths = [threading.Thread(target=work, args=(s, reps)) for s in range(depth)]
        t0 = time.time()
        for t in ths: t.start()
        for t in ths: t.join()
        dt = time.time() - t0
    print("depth %d x %d problems: %.1f ms per %d problems, %.0f solves/s" % (depth, Pb, 1e3 * dt / 4, Ptot, 4 * Ptot / dt))
    for ctx, b, arr in slots:
        b.close(); ctx.close()
