"""Small fixed case for ncu captures: one batched OD solve schedule (few iterations)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vinsat_b200 import _lib, synth
P = int(sys.argv[1]) if len(sys.argv) > 1 else 148
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
ctx = _lib.Context(0)
prs = synth.make_batch(P, T, 10, seed0=0)
b = _lib.Batch(ctx, _lib.concat_problems(prs))
b.od_solve(12, 10, 1e-4)
b.eval_resjac()
ctx.synchronize()
print("done", ctx.launch_count())
