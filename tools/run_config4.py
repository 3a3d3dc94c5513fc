"""configs[3] of BASELINE.json: 65,536 OD problems with a pixel-noise sweep, sharded over the GPUs of one box by the
dynamic chunk pool.  Launch: python tools/run_config4.py  |  torchrun --nproc-per-node N tools/run_config4.py
Prints one JSON record on rank 0 (copied to profiles/)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

_real_stdout = os.dup(1)          # NCCL prints its version banner on fd 1: keep stdout for the ONE JSON record
os.dup2(2, 1)

from vinsat_b200 import config
from vinsat_b200.eval import batch_runner

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
n_problems = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
workers = int(os.environ.get("VINSAT_MC_WORKERS", "3"))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
config.device = local
# setup (base arcs on the host, resident batches), a warm-up pool, then the timed pool
odp = batch_runner.ODPool(rank=rank, world_size=world, workers=workers, device=local)
odp.run(2 * workers * world * 1024, "warm")
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
res = odp.run(n_problems, "timed")
torch.cuda.synchronize()
t_own = time.perf_counter() - t0
odp.close()
if world > 1:
    dist.barrier()
wall = time.perf_counter() - t0
gathered = [None] * world
if world > 1:
    dist.all_gather_object(gathered, (rank, t_own, res))
else:
    gathered = [(rank, t_own, res)]
if rank == 0:
    allres = [r for g in gathered for r in g[2]]
    chunks = sorted(r["chunk"] for r in allres)
    assert chunks == list(range(len(chunks))), "every chunk exactly once"
    os.write(_real_stdout, (json.dumps({
        "config": "configs[3]: %d OD problems (1000 frames x 10 obs), noise sweep sigma_px in %s, %d GPU(s), %d solves in flight per GPU, "
                  "dynamic chunk pool (1024 problems per chunk)" % (n_problems, list(batch_runner.NOISE_SWEEP_PX), world, workers),
        "wall_s": wall, "solves_per_s": n_problems / wall,
        "timing": "wall clock from a barrier to the barrier after the last chunk (setup of the resident base arcs excluded); per chunk: "
                  "8-byte seed in, device draws of pixel noise + initial guess, 20 BA iterations, 16 KB of per-problem errors out", "chunks_per_rank": {g[0]: len(g[2]) for g in gathered},
        "noise_sweep": {str(k): v for k, v in batch_runner.summarize_noise_sweep(allres).items()}}) + "\n").encode())
if world > 1:
    dist.destroy_process_group()
