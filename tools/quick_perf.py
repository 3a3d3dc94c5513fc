"""Ad-hoc perf probe (not the bench): per-family device time of one batched OD solve + the headline kernel."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vinsat_b200 import _lib, synth

P = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
K = int(sys.argv[3]) if len(sys.argv) > 3 else 10
ctx = _lib.Context(0)
print("fp64 peak TFLOP/s", ctx.fp64_peak_tflops(), "copy GB/s", ctx.copy_bw_gbs(1 << 30))
t0 = time.time(); prs = synth.make_batch(P, T, K, seed0=0); print("synth s", time.time() - t0)
arrays = _lib.concat_problems(prs)
t0 = time.time(); b = _lib.Batch(ctx, arrays); print("create+upload s", time.time() - t0)
for rep in range(2):
    b.upload(arrays)
    ctx.synchronize(); t0 = time.time(); b.od_solve(20, 10, 1e-4); ctx.synchronize(); dt = time.time() - t0
    print("od_solve wall s", dt, "solves/s", P / dt)
st = b.get_states()
err = [np.abs(st[arrays['frame_off'][p]:arrays['frame_off'][p+1], :3] - prs[p]['states_gt'][:, :3]).max() for p in range(P)]
print("max pos err vs truth km: median %.3f max %.3f" % (np.median(err), np.max(err)))
b.upload(arrays); ctx.enable_timing(True); ctx.reset_timing()
b.od_solve(20, 10, 1e-4)
tm = ctx.timing(); ctx.enable_timing(False)
tot = sum(v[0] for v in tm.values())
for k, v in sorted(tm.items(), key=lambda kv: -kv[1][0]):
    if v[1]: print("%-20s %9.3f ms %6d launches %5.1f%%" % (k, v[0], v[1], 100 * v[0] / tot))
# headline kernel
M = b.M
ctx.enable_timing(True); ctx.reset_timing()
for _ in range(10): b.eval_resjac()
tm = ctx.timing(); ctx.enable_timing(False)
ms = tm["project_resjac"][0] / tm["project_resjac"][1]
print("resjac: M=%d  %.3f ms  %.3e obs/s  %.1f GB/s (156 B/obs)" % (M, ms, M / ms * 1e3, M * 156 / ms / 1e6))
