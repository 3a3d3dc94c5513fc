"""N>1 host logic on CPU: world_size-2 `gloo` process groups (no GPU).  Covers the Monte-Carlo shard ranges,
the frame-window plan of the long arc (every frame / observation owned exactly once, ghosts consistent) and the
exchange choreography of the exact distributed median (same radix schedule as the kernels)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vinsat_b200 import longarc, synth
    import ba_oracle as o
    out = {}
    # --- Monte-Carlo shard ranges (eval/batch_runner.py::run_od_monte_carlo) ---
    n = 11
    lo, hi = (n * rank) // world, (n * (rank + 1)) // world
    ids = torch.zeros(n, dtype=torch.int64); ids[lo:hi] = 1
    dist.all_reduce(ids)
    out["mc_partition"] = bool((ids == 1).all())
    # --- long-arc windows ---
    pr = synth.make_problem(5, 37, 4, empty_frame_frac=0.2)
    T, M = pr["states0"].shape[0], pr["xyz"].shape[0]
    wlo, whi = longarc.plan_windows(T, world)[rank]
    a, olo, ohi = longarc.window_arrays(pr, wlo, whi)
    own_f = torch.zeros(T, dtype=torch.int64); own_f[wlo:whi] = 1
    own_o = torch.zeros(M, dtype=torch.int64)
    k0 = int(np.searchsorted(pr["ii"], wlo)); own_o[k0:k0 + a["obs_off"][1]] = 1
    dist.all_reduce(own_f); dist.all_reduce(own_o)
    out["frames_once"] = bool((own_f == 1).all()); out["obs_once"] = bool((own_o == 1).all())
    g0 = wlo - olo
    out["ghosts_ok"] = bool(np.array_equal(a["states"], pr["states0"][g0:g0 + a["frame_off"][1]]) and
                            np.array_equal(a["ii"] + g0, pr["ii"][k0:k0 + a["obs_off"][1]]) and
                            (olo == (1 if rank > 0 else 0)) and (a["frame_off"][1] - ohi == (1 if rank < world - 1 else 0)))
    # --- exact distributed median == torch.median semantics on the union ---
    uv = o.landmark_project(pr["states0"], pr["xyz"], pr["intr"], pr["ii"], jacobian=False)
    r = (pr["uv"] - uv)
    local = r[k0:k0 + a["obs_off"][1]].reshape(-1)

    def ar(x):
        t = torch.from_numpy(np.ascontiguousarray(x)); dist.all_reduce(t); return t.numpy()
    med = longarc.distributed_lower_median_host(local, ar)
    out["median_exact"] = (med == o.lower_median(np.abs(r)))
    # --- accept sums are plain all-reduces of per-window partial sums ---
    part = torch.tensor([np.abs(local).sum()], dtype=torch.float64); dist.all_reduce(part)
    out["sum_close"] = bool(abs(part.item() - np.abs(r).sum()) <= 1e-9 * np.abs(r).sum())
    # --- dynamic chunk pool (vinsat_b200/pool.py): every chunk handed out exactly once across ranks x workers ---
    from vinsat_b200 import pool
    import time
    n_chunks = 23
    counter = pool.make_counter("test_pool_%d" % world, world)
    mine = pool.drain(counter, n_chunks, lambda w, c: time.sleep(0.002 * (1 + rank)), n_workers=2)
    got = torch.zeros(n_chunks, dtype=torch.int64)
    for _, c in mine:
        got[c] += 1
    dist.all_reduce(got)
    out["pool_each_chunk_once"] = bool((got == 1).all())
    cnt = torch.tensor([len(mine)], dtype=torch.int64); cmin = cnt.clone(); dist.all_reduce(cmin, op=dist.ReduceOp.MIN)
    out["pool_every_rank_worked"] = bool(cmin.item() >= 1)
    # --- SatCam sweep: contiguous pose slices cover [0, n) once (sim/SatCam.py::rank_slice) ---
    from vinsat_b200.sim.SatCam import rank_slice
    lo2, hi2 = rank_slice(1_000_003, rank, world)
    cover = torch.tensor([hi2 - lo2], dtype=torch.int64); dist.all_reduce(cover)
    edges = torch.zeros(world + 1, dtype=torch.int64); edges[rank] = lo2; edges[rank + 1] += 0
    out["satcam_slices_cover"] = bool(cover.item() == 1_000_003 and lo2 == (1_000_003 * rank) // world)
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_multirank_host_logic_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, out in res:
        assert all(out.values()), (rank, out)
