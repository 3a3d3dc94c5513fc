"""Measurements of the BASELINE.json configs that are not the bench line (numbers quoted in DESIGN.md section 7):
  c4 : per-GPU share of configs[3] (65,536 problems over 8 GPUs = 8192 problems x 1000 frames on one GPU)
  c5 : configs[4], SatCam visibility sweep (all MGRS landmark centroids x nadir poses), poses/s and pairs/s
  c3 : configs[2], single long arc T=100,000 x 50 obs/frame on one GPU (whole-arc, partitioned sweep)
usage: python tools/measure_configs.py c4|c5|c3 [size]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vinsat_b200 import _lib, synth

what = sys.argv[1]
ctx = _lib.Context(0)
out = {"config": what}
if what == "c4":
    P = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    t0 = time.time()
    base = synth.make_batch(1024, 1000, 10, seed0=0)
    prs = [base[i % 1024] for i in range(P)]                 # replicated inputs: throughput only
    arrays = _lib.concat_problems(prs)
    out["synth_s"] = time.time() - t0
    b = _lib.Batch(ctx, arrays)
    b.od_solve(20, 10, 1e-4); ctx.synchronize()
    ts = []
    for _ in range(2):
        b.upload(arrays); ctx.synchronize()
        t0 = time.time(); b.od_solve(20, 10, 1e-4); ctx.synchronize(); ts.append(time.time() - t0)
    ctx.enable_timing(True); ctx.reset_timing(); b.upload(arrays); b.od_solve(20, 10, 1e-4); tm = ctx.timing(); ctx.enable_timing(False)
    out.update(P=P, T=1000, K=10, od_solve_s=min(ts), solves_per_s=P / min(ts),
               kernel_ms={k: round(v[0], 2) for k, v in tm.items() if v[1]})
elif what == "c5":
    n_pose = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    from vinsat_b200.sim import SatCam as SC
    from vinsat_b200 import trajgen_pipe as tp, hostmath as hm
    L, _, _ = SC.all_landmark_centroids_ecef()
    L = np.ascontiguousarray(L, dtype=np.float64).reshape(-1, 3)
    n_arc = (n_pose + 10800) // 10801
    rng = np.random.default_rng(0)
    x0 = np.stack([hm.oe2eci_values(6978 + rng.uniform(-50, 50), rng.uniform(0, 0.01), np.pi / 2 + rng.uniform(-.1, .1),
                                    *rng.uniform(0, 2 * np.pi, 3)) for _ in range(n_arc)])
    traj = tp.propagate_orbits(x0, 10800, 1.0).reshape(-1, 6)[:n_pose]
    # nadir camera frames (orbit_gen.py:322 get_nadir_attitude_vectors); the inertial positions are used as
    # Earth-fixed ones (no Earth rotation) -- immaterial for a throughput measurement
    pos = traj[:, :3] * 1e3
    up = pos / np.linalg.norm(pos, axis=1, keepdims=True)
    east = np.cross(np.array([0.0, 0.0, 1.0]), up); east /= np.linalg.norm(east, axis=1, keepdims=True)
    north = np.cross(up, east)
    poses = np.concatenate([pos, -up, north, east], axis=1)
    out["note"] = "poses from %d seeded 3 h polar arcs at 1 Hz, nadir pointing" % n_arc
    SC.inframe_sweep(poses[:8192], L, chunk=8192)
    t0 = time.time(); counts = SC.inframe_sweep(poses, L, chunk=8192); dt = time.time() - t0
    out.update(n_poses=len(poses), n_landmarks=len(L), sweep_s=dt, poses_per_s=len(poses) / dt, pairs_per_s=len(poses) * len(L) / dt,
               mean_inframe=float(counts.mean()), max_inframe=int(counts.max()))
    SC.visibility_sweep(poses[:8192])
    t0 = time.time(); vis = SC.visibility_sweep(poses); dt = time.time() - t0
    out.update(visibility_s=dt, visibility_poses_per_s=len(poses) / dt, n_visible=int(vis.sum()))
elif what == "c3":
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
    t0 = time.time(); pr = synth.make_problem(5, T, 50, gap_max=3); out["synth_s"] = time.time() - t0
    b = _lib.Batch(ctx, _lib.concat_problems([pr]))
    lam = 1e-4
    for it in range(10):
        lam, _ = b.ba_iterate(it, lam, initialize=True)
    ts = []
    for it in range(10, 20):
        ctx.synchronize(); t0 = time.time(); lam, ntr = b.ba_iterate(it, lam, initialize=False); ctx.synchronize(); ts.append(time.time() - t0)
    st = b.get_states()
    ctx.enable_timing(True); ctx.reset_timing(); lam, ntr = b.ba_iterate(20, lam, initialize=False); tm = ctx.timing(); ctx.enable_timing(False)
    out["kernel_ms"] = {k: round(v[0], 3) for k, v in tm.items() if v[1]}
    out["launches"] = {k: v[1] for k, v in tm.items() if v[1]}
    out.update(T=T, M=int(b.M), ms_per_full_iteration=1e3 * float(np.median(ts)),
               max_pos_err_km=float(np.abs(st[:, :3] - pr["states_gt"][:, :3]).max()))
print(json.dumps(out))
