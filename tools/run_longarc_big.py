"""configs[2] beyond the bench size: one arc of T frames x K obs/frame, frame-window sharded over the ranks (bench.longarc_leg),
with the whole arc solved on rank 0's GPU beside it.  The truth orbit comes from the device propagator (`vinsat_orbit_propagate`),
the host RK4 loop of `synth` would take minutes at these lengths.
Keep T x mean gap below ~2.5e6 s: the reference's J2 term is not conservative, an orbit integrated with it turns
eccentric after ~3e6 s and escapes after ~4.5e6 s, so longer synthetic arcs are not OD problems any more.
Launch: torchrun --nproc-per-node N tools/run_longarc_big.py T K [gap_max]   ->  one JSON record on rank 0."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

_real_stdout = os.dup(1)          # NCCL prints its banner on fd 1: keep stdout for the ONE JSON record
os.dup2(2, 1)

import bench
from vinsat_b200 import _lib

T = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
gap_max = int(sys.argv[3]) if len(sys.argv) > 3 else 1
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
_lib.bind_host_thread_to_gpu(local)


def sync_all():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


ctx = _lib.Context(local)
stream = torch.cuda.Stream(device=dev)
ctx.set_stream(stream.cuda_stream)
torch.cuda.set_stream(stream)
gen = _lib.Context(local)         # own context for input generation: its scratch is released before the solve
orbit_fn = lambda x0, n: gen.orbit_propagate(x0, n, 1, 1.0)
out = bench.longarc_leg(ctx, rank, world, max_over_ranks, sync_all, T=T, K=K, make_kw=dict(orbit_fn=orbit_fn, gap_max=gap_max))
gen.close()
out["n_gpus"] = world
out["gap_max_s"] = gap_max
if rank == 0:
    os.write(_real_stdout, (json.dumps(out) + "\n").encode())
sync_all()
ctx.close()
if world > 1:
    dist.destroy_process_group()
