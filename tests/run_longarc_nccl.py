"""Real NCCL check + timing of the frame-window sharded long arc.  Launch with
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tests/run_longarc_nccl.py [T] [K] [iters]
Rank 0 also solves the whole arc on its own GPU and compares (1 m / 1 mm/s and identical LM schedules)."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vinsat_b200 import _lib, longarc, synth  # noqa: E402


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    ctx = _lib.Context(local)
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx._bound_to_torch = True
    pr = synth.make_problem(123, T, K, gap_max=3 if T > 5000 else 20)
    la = longarc.LongArc(pr, ctxs=[ctx], use_dist=world > 1, world=world)
    n_init = min(10, iters // 2)
    graphs = la.enable_graphs() and os.environ.get("VINSAT_LA_HOST_LM") is None

    def solve():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        lam = 1e-4
        sched = []
        for it in range(iters):
            if graphs:
                lam, ntr = la.ba_iterate_device_lm(it, 1e-4 if it == 0 else None, initialize=it < n_init)
            else:
                lam, ntr = la.ba_iterate(it, lam, initialize=it < n_init)
            sched.append((lam, ntr))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        return time.perf_counter() - t0, sched

    dt_first, sched = solve()
    la.reset_states()
    dt, sched_b = solve()                      # second solve: every CUDA graph replays
    assert sched == sched_b
    st = la.gather_states()
    out = dict(world=world, T=T, M=T * K, iters=iters, seconds=dt, ms_per_iteration=1e3 * dt / iters,
               first_solve_ms_per_iteration=1e3 * dt_first / iters, device_lm_and_graphs=bool(graphs and la.use_graphs),
               graph_error=getattr(la, "graph_error", None), graph_replays=getattr(la, "n_graph_replays", 0),
               segments_per_rank=la.S, collectives=la.n_collectives,
               max_pos_err_vs_truth_km=float(np.abs(st[:, :3] - pr["states_gt"][:, :3]).max()))
    if rank == 0:
        ctx2 = _lib.Context(local)
        b = _lib.Batch(ctx2, _lib.concat_problems([pr]))
        ctx2.synchronize(); t1 = time.perf_counter()
        lam2 = np.array([1e-4]); sched2 = []
        for it in range(iters):
            lam2, ntr2 = b.ba_iterate(it, lam2, initialize=it < n_init)
            sched2.append((float(lam2[0]), int(ntr2[0])))
        ctx2.synchronize()
        out["whole_arc_1gpu_seconds"] = time.perf_counter() - t1
        ref = b.get_states()
        out["dpos_m"] = float(np.abs(st[:, :3] - ref[:, :3]).max() * 1e3)
        out["dvel_mm_s"] = float(np.abs(st[:, 7:] - ref[:, 7:]).max() * 1e6)
        out["same_lm_schedule"] = sched == sched2
        out["ok"] = bool(out["dpos_m"] < 1.0 and out["dvel_mm_s"] < 1.0 and out["same_lm_schedule"])
        print(json.dumps(out))
    la.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
