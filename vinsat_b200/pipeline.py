"""Host-side pipelining of batched OD solves: `depth` batches in flight on one GPU (default 3: measured 17.6 k
solves/s one at a time, 23.8 k with two in flight, 25.2 k with three, 26.0 k with inputs resident).

One `vinsat_batch_od_solve` keeps the GPU busy for tens of milliseconds while the PCIe copy engines idle, and a
`vinsat_batch_upload` does the opposite.  `PipelinedSolver` owns `depth` (context, stream, device batch) slots,
each driven by its own host thread (ctypes releases the GIL inside the C ABI): while one slot solves, the next
one uploads its inputs and the previous one downloads its states.  Every solve still does its own host->device
copy of ALL inputs and its own device->host copy of the solved states; nothing is cached between solves.

This is an ADDED entry point (the reference is batch-size-1 and synchronous).
"""
import threading

from . import _lib


class PipelinedSolver:
    def __init__(self, device, template_arrays, depth=3):
        """template_arrays: dict as `_lib.concat_problems` returns; fixes (P, T, M) and the frame offsets of every
        batch that will be pushed through this solver."""
        self.slots = []
        for _ in range(depth):
            ctx = _lib.Context(device)
            self.slots.append((ctx, _lib.Batch(ctx, template_arrays)))

    def solve_many(self, inputs, outputs, num_iters=20, n_init=10, lamda_init=1e-4, mode=_lib.MODE_STEP1S):
        """inputs[i]: host arrays (pinned for full copy bandwidth) of job i; outputs[i]: host (T, 10) array that
        receives the solved states of job i.  Jobs are dealt round-robin to the slots; returns when all are done."""
        errs = []

        def worker(s):
            ctx, batch = self.slots[s]
            try:
                for i in range(s, len(inputs), len(self.slots)):
                    batch.upload(inputs[i])
                    batch.od_solve(num_iters, n_init, lamda_init, mode)
                    batch.get_states(outputs[i])
            except Exception as e:      # surfaced to the caller below
                errs.append(e)

        ths = [threading.Thread(target=worker, args=(s,)) for s in range(len(self.slots))]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        if errs:
            raise errs[0]

    def solve_pool(self, counter, n_jobs, inputs_of, output_of, num_iters=20, n_init=10, lamda_init=1e-4,
                   mode=_lib.MODE_STEP1S, on_done=None):
        """Dynamic version of `solve_many` for a work pool shared by several GPUs / processes (vinsat_b200/pool.py):
        every slot pulls job indices c from `counter` until `n_jobs` are handed out; job c reads `inputs_of(c)` (host
        arrays) and writes the solved states to `output_of(slot)` (one host buffer per slot, consumed by
        `on_done(c, slot)` before the slot's next job).  Returns the number of jobs this solver executed."""
        from . import pool

        def work(s, c):
            ctx, batch = self.slots[s]
            batch.upload(inputs_of(c))
            batch.od_solve(num_iters, n_init, lamda_init, mode)
            batch.get_states(output_of(s))
            if on_done:
                on_done(c, s)

        return len(pool.drain(counter, n_jobs, work, len(self.slots)))

    def close(self):
        for ctx, batch in self.slots:
            batch.close()
            ctx.close()
        self.slots = []
