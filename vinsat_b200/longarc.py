"""Frame-window sharded BA of ONE long arc (BASELINE.json configs[2]).

The arc's frames are cut into contiguous windows, one per rank (one process per GPU).  Every rank runs the
ordinary per-frame / per-observation kernels on its window (+ one ghost frame per side) and the host exchanges
small device buffers between the stages of ``vinsat_la_stage`` (include/vinsat_b200.h):

    robust scale   6 x all-reduce(SUM) of a 2048-bin digit histogram   -> exact global lower median
                   all-reduce(MAX) of the largest weight
    LM trial       all-gather of the per-segment reduced-system records -> every rank solves the small
                   reduced block-tridiagonal chain redundantly, then back-substitutes its own interiors
                   all-gather of the two edge states per rank           -> ghost frames after the retraction
    accept test    all-reduce(SUM) of two partial sums

With ``torch.distributed`` (backend nccl) the buffers never leave the GPUs (NVLink / NVSwitch); all messages are
well under 1 MB, i.e. latency bound, as SURVEY 8(e) predicts.  The same driver can hold several windows in ONE
process (``parts``), which is how the single-GPU tests emulate a multi-rank run without NCCL.

Results equal ``vinsat_batch_ba_iterate`` on the whole arc up to summation order (tested to 1e-9 relative on the
step; converged states within 1 m / 1 mm/s).
"""
import ctypes as C
import math

import numpy as np

from . import _lib

(RESID, SELECT_BEGIN, SELECT_HIST, SELECT_PICK, ASSEMBLE, DYNAMICS, SYSTEM, SUMS_INIT, SET_LAM, SOLVE_INIT, FORWARD,
 REDUCED, BACKSUB, RETRACT, PACK_EDGES, APPLY_GHOSTS, TRIAL, SUMS_TRIAL, COMMIT, BEGIN_ITER, ACCEPT, COMMIT_COPY) = range(22)
BUF_HIST, BUF_WMAX, BUF_SUMS, BUF_PACK, BUF_GATHER, BUF_EDGE, BUF_EDGES_ALL, BUF_FLAGS, BUF_LAM_NEXT, BUF_NTRIALS = range(10)
_TYPESTR = {BUF_HIST: "<i4", BUF_WMAX: "<i8", BUF_FLAGS: "<i4", BUF_NTRIALS: "<i4"}


def plan_windows(T, world):
    """Contiguous frame windows [lo, hi) per rank."""
    return [(T * r // world, T * (r + 1) // world) for r in range(world)]


def plan_segments(T, world, sm_count=148):
    """Common number of segments per rank.  Short arcs: segment length ~ sqrt(0.67 T) balances the parallel interior
    sweeps (len x ~3 us) against the sequential reduced chain (world x segments x ~2 us).  Long arcs: as many segments as a GPU
    runs chains at once (8 one-warp CTAs x 3 chains per SM), each at least 24 frames long -- the gathered reduced chain is then
    partitioned a second time inside the library (`Level2`, csrc/batch.h) instead of being walked by one warp."""
    import os
    seg_len = max(8.0, math.sqrt(0.67 * T))
    S = max(1, int(round((T / world) / seg_len)))
    l2_min = int(os.environ.get("VINSAT_L2_MIN", "256"))
    S2 = min(8 * sm_count * 3, (T // world) // 24)
    if l2_min > 0 and S2 * world >= l2_min and S2 > S:
        S = S2
    return S


def window_arrays(pr, lo, hi):
    """Local arrays of a rank: owned frames [lo,hi) plus one ghost frame per side, observations of owned frames."""
    T = pr["states0"].shape[0]
    g0, g1 = max(lo - 1, 0), min(hi + 1, T)
    ii = np.asarray(pr["ii"], dtype=np.int64)
    k0, k1 = np.searchsorted(ii, lo, side="left"), np.searchsorted(ii, hi, side="left")
    a = dict(frame_off=np.array([0, g1 - g0], dtype=np.int64), obs_off=np.array([0, k1 - k0], dtype=np.int64),
             states=np.ascontiguousarray(pr["states0"][g0:g1], dtype=np.float64),
             intrinsics=np.ascontiguousarray(pr["intr"][g0:g1], dtype=np.float64),
             cum_rot=np.ascontiguousarray(pr["cum_rot"][g0:g1], dtype=np.float64),
             time_idx=np.ascontiguousarray(pr["time_idx"][g0:g1], dtype=np.int64),
             landmarks_xyz=np.ascontiguousarray(pr["xyz"][k0:k1], dtype=np.float64),
             landmarks_uv=np.ascontiguousarray(pr["uv"][k0:k1], dtype=np.float64),
             confidences=np.ascontiguousarray(pr["conf"][k0:k1], dtype=np.float64),
             ii=np.ascontiguousarray(ii[k0:k1] - g0, dtype=np.int64))
    return a, lo - g0, hi - g0


class _DevArray:
    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class Part:
    """One frame window on one GPU."""

    def __init__(self, ctx, pr, lo, hi, rank, n_segments):
        import torch
        self.ctx, self.rank, self.lo, self.hi = ctx, rank, lo, hi
        arrays, self.own_lo, self.own_hi = window_arrays(pr, lo, hi)
        self.arrays = arrays
        self.lib = ctx.lib
        self.batch = _lib.Batch(ctx, arrays, window=(self.own_lo, self.own_hi, n_segments))
        h = self.batch.h
        self.n_seg = int(ctx.lib.vinsat_la_num_segments(h))
        self.torch = torch
        self.device = torch.device("cuda", ctx.device)
        self._bufs = {}

    def alloc_reduced(self, S_total, world):
        self.ctx.check(self.lib.vinsat_la_alloc_reduced(self.batch.h, int(S_total), int(world)))
        self._bufs = {}

    def stage(self, stage, i0=0, i1=0, d0=0.0):
        self.ctx.check(self.lib.vinsat_la_stage(self.batch.h, int(stage), int(i0), int(i1), float(d0)))

    def buf(self, which):
        """torch view (zero copy) of a device buffer of the library."""
        if which not in self._bufs:
            p, n = C.c_void_p(), C.c_int64()
            self.ctx.check(self.lib.vinsat_la_ptr(self.batch.h, int(which), C.byref(p), C.byref(n)))
            self._bufs[which] = self.torch.as_tensor(_DevArray(p.value, n.value, _TYPESTR.get(which, "<f8")), device=self.device)
        return self._bufs[which]

    def owned_states(self):
        st = self.batch.get_states()
        return st[self.own_lo:self.own_hi]

    def close(self):
        self.batch.close()


class LongArc:
    """Driver.  `parts` are the windows held by THIS process (normally one); `dist_group` spans the processes."""

    def __init__(self, pr, ctxs=None, world=None, rank=None, use_dist=False, n_segments=None, mode=_lib.MODE_STEP1S):
        import torch
        self.torch = torch
        self.pr = pr
        self.T = pr["states0"].shape[0]
        self.M = pr["xyz"].shape[0]
        self.mode = mode
        self.use_dist = use_dist
        if use_dist:
            import torch.distributed as dist
            self.dist = dist
            self.world = dist.get_world_size()
            ranks = [dist.get_rank()]
        else:
            self.world = world if world is not None else 1
            ranks = list(range(self.world)) if rank is None else [rank]
        wins = plan_windows(self.T, self.world)
        S = n_segments if n_segments is not None else plan_segments(self.T, self.world)
        S = max(1, min(S, min(hi - lo for lo, hi in wins)))
        ctxs = ctxs or [_lib.default_context(0)]
        self.parts = [Part(ctxs[i % len(ctxs)], pr, wins[r][0], wins[r][1], r, S) for i, r in enumerate(ranks)]
        for p in self.parts:
            assert p.n_seg == S
        self.S, self.S_total = S, S * self.world
        for p in self.parts:
            p.alloc_reduced(self.S_total, self.world)
        self.n_collectives = 0
        self.use_graphs = False
        self._exchange_ghosts(current=True)        # ghost rows came from the host arrays; keeps the code path uniform

    # ---- collectives over (local parts x processes) ------------------------------------------------------
    def _sync(self):
        for p in self.parts:
            p.ctx.synchronize()

    def _all_reduce(self, which, op):
        torch = self.torch
        self.n_collectives += 1
        if self.use_dist:
            (p,) = self.parts
            self._stream_handoff(p)
            self.dist.all_reduce(p.buf(which), op=self.dist.ReduceOp.SUM if op == "sum" else self.dist.ReduceOp.MAX)
            return
        if len(self.parts) == 1:
            return
        self._sync()
        stack = torch.stack([p.buf(which).to(self.parts[0].device) for p in self.parts])
        red = stack.sum(0) if op == "sum" else stack.max(0).values
        for p in self.parts:
            p.buf(which).copy_(red.to(p.device))
        torch.cuda.synchronize()

    def _all_gather(self, src, dst):
        torch = self.torch
        self.n_collectives += 1
        if self.use_dist:
            (p,) = self.parts
            self._stream_handoff(p)
            self.dist.all_gather_into_tensor(p.buf(dst), p.buf(src))
            return
        if len(self.parts) == 1 and self.world == 1 and getattr(self.parts[0].ctx, "_bound_to_torch", False):
            self.parts[0].buf(dst).copy_(self.parts[0].buf(src))      # same stream as the library: ordered, capturable
            return
        self._sync()
        cat = torch.cat([p.buf(src).to(self.parts[0].device) for p in sorted(self.parts, key=lambda q: q.rank)])
        for p in self.parts:
            p.buf(dst).copy_(cat.to(p.device))
        torch.cuda.synchronize()

    def _stream_handoff(self, p):
        """The library launches on the context stream; NCCL runs on torch's.  Both are the same stream when the
        caller bound them (ctx.set_stream(torch stream)); otherwise order them by a host sync."""
        if not getattr(p.ctx, "_bound_to_torch", False):
            p.ctx.synchronize()

    def _after_collective(self):
        if self.use_dist and not getattr(self.parts[0].ctx, "_bound_to_torch", False):
            self.torch.cuda.synchronize()

    def _stage(self, stage, i0=0, i1=0, d0=0.0, per_rank_i0=None):
        for p in self.parts:
            p.stage(stage, per_rank_i0(p) if per_rank_i0 else i0, i1, d0)

    def _exchange_ghosts(self, current):
        self._stage(PACK_EDGES, 1 if current else 0)
        self._all_gather(BUF_EDGE, BUF_EDGES_ALL)
        self._after_collective()
        for p in self.parts:
            p.stage(APPLY_GHOSTS, p.rank, 1 if current else 0)

    def _read(self, which):
        """Host copy of a (replicated) device buffer, ordered after the library's and torch's streams."""
        self.parts[0].ctx.synchronize()
        self.torch.cuda.synchronize(self.parts[0].device)
        return self.parts[0].buf(which).cpu().numpy()

    def _global_sums(self):
        self._all_reduce(BUF_SUMS, "sum")
        self._after_collective()
        s = self._read(BUF_SUMS)
        return float(s[0]), float(s[1])

    # ---- one BA iteration (BA_filtering.py:4-98) on the sharded arc ----------------------------------------
    def ba_iterate(self, it, lamda_init, initialize=False):
        alpha = min(max(1 - (2 * (it / 5) - 1), 1), 2)
        Sigma = float(min(10000 * (it + 1) ** 2, 1000000))
        sq = math.sqrt(Sigma)
        init = 1 if initialize else 0
        self._stage(RESID)
        self._stage(SELECT_BEGIN, 2 * self.M)
        for ps in range(6):
            self._stage(SELECT_HIST, ps)
            self._all_reduce(BUF_HIST, "sum")
            self._after_collective()
            self._stage(SELECT_PICK, ps)
        self._stage(ASSEMBLE, d0=alpha)
        self._all_reduce(BUF_WMAX, "max")
        self._after_collective()
        if not initialize:
            self._stage(DYNAMICS, self.mode)
            self._stage(SYSTEM, 0, 0, Sigma)
        self._stage(SUMS_INIT, init)                      # local sums -> BUF_SUMS[0:2]; reduced together with the trial's
        n = 2.0 * self.M + (6.0 if initialize else 7.0) * max(self.T - 1, 0)
        init_res = None
        lam = float(lamda_init)
        ntrials = 0
        while True:
            self._stage(SET_LAM, d0=lam)
            if initialize:
                self._stage(SOLVE_INIT)
            else:
                self._stage(FORWARD)
                self._all_gather(BUF_PACK, BUF_GATHER)
                self._after_collective()
                self._stage(REDUCED, per_rank_i0=lambda p: p.rank * self.S)
                self._stage(BACKSUB)
            self._stage(RETRACT)
            self._exchange_ghosts(current=False)
            self._stage(TRIAL, self.mode, init)
            # trial sums -> BUF_SUMS[2:4] (observation part already divided by the global max weight on the device);
            # from the second trial on the already-reduced init sums are zeroed first
            self._stage(SUMS_TRIAL, init, 1 if ntrials > 0 else 0)
            self._all_reduce(BUF_SUMS, "sum")
            self._after_collective()
            s4 = self._read(BUF_SUMS)                     # the ONE host read of the trial
            if init_res is None:
                init_res = (float(s4[0]) + sq * float(s4[1])) / n
            residual = (float(s4[2]) + sq * float(s4[3])) / n
            ntrials += 1
            lam = lam * 10
            if residual < init_res or lam > 1e4:
                break
        self._stage(COMMIT)
        return max(min(1e-1, lam * 0.01), 1e-4), ntrials

    # ---- the same iteration with the LM bookkeeping on the device: one CUDA graph per (alpha, Sigma, phase) -------
    def _graphable(self):
        return len(self.parts) == 1          # one window per process (the emulated multi-part mode stages through the host)

    def _issue_trial(self, init, first, n, sq):
        """Everything of one LM trial, asynchronously (no host decision inside): solve, retract, ghosts, trial residuals,
        ONE all-reduce of the four sums, device-side accept test."""
        if init:
            self._stage(SOLVE_INIT)
        else:
            self._stage(FORWARD)
            self._all_gather(BUF_PACK, BUF_GATHER)
            self._stage(REDUCED, per_rank_i0=lambda p: p.rank * self.S)
            self._stage(BACKSUB)
        self._stage(RETRACT)
        self._exchange_ghosts(current=False)
        self._stage(TRIAL, self.mode, init)
        self._stage(SUMS_TRIAL, init, 0 if first else 1)
        self._all_reduce(BUF_SUMS, "sum")
        self._stage(ACCEPT, init, n, sq)

    def _issue_head(self, it, init, alpha, Sigma, n, sq):
        """Linearisation + first trial (everything up to the first host decision)."""
        self._stage(BEGIN_ITER)
        self._stage(RESID)
        self._stage(SELECT_BEGIN, 2 * self.M)
        for ps in range(6):
            self._stage(SELECT_HIST, ps)
            self._all_reduce(BUF_HIST, "sum")
            self._stage(SELECT_PICK, ps)
        self._stage(ASSEMBLE, d0=alpha)
        self._all_reduce(BUF_WMAX, "max")
        if not init:
            self._stage(DYNAMICS, self.mode)
            self._stage(SYSTEM, 0, 0, Sigma)
        self._stage(SUMS_INIT, init)
        self._issue_trial(init, True, n, sq)

    def _run(self, key, issue):
        """Runs `issue()` as a CUDA graph: the first time a key is seen it runs eagerly (one-time host actions such as
        function attributes and scratch growth happen there), the second time it is captured -- library launches
        and NCCL collectives alike -- and from then on replayed."""
        torch = self.torch
        if not self.use_graphs:
            issue()
            return
        g = self._graphs.get(key)
        if g is not None:
            g.replay()
            self.n_graph_replays += 1
            return
        if key not in self._graph_seen:
            self._graph_seen.add(key)
            issue()
            return
        (p,) = self.parts
        try:
            g = torch.cuda.CUDAGraph()
            cap = self._capture_stream
            cap.wait_stream(torch.cuda.current_stream())
            p.ctx.set_stream(cap.cuda_stream)       # (synchronises: must happen outside the capture)
            try:
                with torch.cuda.graph(g, stream=cap):
                    issue()
            finally:
                p.ctx.set_stream(self._home_stream.cuda_stream)
            self._graphs[key] = g
            g.replay()
            self.n_graph_replays += 1
        except Exception as e:
            # a capture that fails half way leaves this rank's stream (and, with NCCL, the communicator) in an undefined
            # state while the other ranks carry on: do not try to continue
            self.use_graphs = False
            self.graph_error = repr(e)[:300]
            raise

    def reset_states(self):
        """Back to the initial guess (same buffers, so captured graphs stay valid)."""
        for p in self.parts:
            p.batch.set_states(p.arrays["states"])
        self._exchange_ghosts(current=True)

    def enable_graphs(self):
        """Call once after construction with the context bound to torch's current stream (ctx.set_stream)."""
        torch = self.torch
        if not self._graphable() or not getattr(self.parts[0].ctx, "_bound_to_torch", False):
            return False
        self._home_stream = torch.cuda.current_stream()
        self._capture_stream = torch.cuda.Stream(device=self.parts[0].device)
        self._graphs, self._graph_seen = {}, set()
        self.n_graph_replays = 0
        self.graph_error = None
        self.use_graphs = True
        self._status = torch.zeros(4, dtype=torch.int32).pin_memory()
        self._lam_host = torch.zeros(1, dtype=torch.float64).pin_memory()
        return True

    def ba_iterate_device_lm(self, it, lamda_init=None, initialize=False):
        """`ba_iterate` with the damping / accept logic on the device (stages BEGIN_ITER, ACCEPT, COMMIT_COPY) and the
        head and the extra trials of an iteration issued as CUDA graphs (`enable_graphs`).  `lamda_init` is only needed
        for the first call (afterwards the device carries lamda from call to call).  Returns (lamda_next, ntrials)."""
        torch = self.torch
        (p,) = self.parts
        alpha = min(max(1 - (2 * (it / 5) - 1), 1), 2)
        Sigma = float(min(10000 * (it + 1) ** 2, 1000000))
        sq = math.sqrt(Sigma)
        init = 1 if initialize else 0
        n = int(2 * self.M + (6 if initialize else 7) * max(self.T - 1, 0))
        if lamda_init is not None:
            self._stage(BEGIN_ITER, d0=float(lamda_init))        # seeds lam_next; the head's own BEGIN_ITER copies it
        self._run(("head", alpha, Sigma if not initialize else 0.0, init), lambda: self._issue_head(it, init, alpha, Sigma, n, sq))
        ntrials = 0
        while True:
            self._status.copy_(p.buf(BUF_FLAGS), non_blocking=True)
            torch.cuda.current_stream().synchronize()            # the ONE host wait of the trial
            ntrials += 1
            if int(self._status[0]) == 0:
                break
            self._run(("trial", Sigma if not initialize else 0.0, init), lambda: self._issue_trial(init, False, n, sq))
        self._stage(COMMIT_COPY)
        self._lam_host.copy_(p.buf(BUF_LAM_NEXT), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(self._lam_host[0]), ntrials

    def od_solve(self, num_iters=20, n_init=10, lamda_init=1e-4):
        lam = lamda_init
        for it in range(num_iters):
            lam, _ = self.ba_iterate(it, lam, initialize=(it < n_init))
        return lam

    def od_solve_device_lm(self, num_iters=20, n_init=10, lamda_init=1e-4):
        lam, total = lamda_init, 0
        for it in range(num_iters):
            lam, ntr = self.ba_iterate_device_lm(it, lamda_init if it == 0 else None, initialize=(it < n_init))
            total += ntr
        return lam, total

    def owned_states(self):
        """{rank: states of its owned frames}."""
        return {p.rank: p.owned_states() for p in self.parts}

    def gather_states(self):
        """Full (T,10) states on every process."""
        own = self.owned_states()
        if not self.use_dist:
            return np.concatenate([own[r] for r in sorted(own)])
        torch, dist = self.torch, self.dist
        (p,) = self.parts
        sizes = [hi - lo for lo, hi in plan_windows(self.T, self.world)]
        mx = max(sizes)
        loc = torch.zeros((mx, 10), dtype=torch.float64, device=p.device)
        loc[:sizes[p.rank]] = torch.from_numpy(own[p.rank]).to(p.device)
        out = torch.empty((self.world * mx, 10), dtype=torch.float64, device=p.device)
        dist.all_gather_into_tensor(out, loc)
        out = out.cpu().numpy().reshape(self.world, mx, 10)
        return np.concatenate([out[r, :sizes[r]] for r in range(self.world)])

    def close(self):
        # captured graphs hold NCCL kernels: release them (and drain the device) before the batches and, later, the
        # process group go away
        if getattr(self, "_graphs", None):
            self._graphs.clear()
            import gc
            gc.collect()
            self.torch.cuda.synchronize()
        for p in self.parts:
            p.close()


# ---- host reference of the exchange choreography (used by the CPU / gloo tests) ---------------------------
SELECT_SHIFTS = (53, 42, 31, 20, 9, 0)
SELECT_BITS = (11, 11, 11, 11, 11, 9)


def distributed_lower_median_host(values_local, all_reduce_sum):
    """Exact lower median (torch.median semantics) of the union of every rank's `values_local` (>= 0) with the
    SAME radix-select schedule as the kernels (k_select_hist / k_select_pick): 6 passes over the bit pattern,
    one all-reduce of a 2048-bin histogram per pass.  `all_reduce_sum(int64 ndarray) -> ndarray` is the collective."""
    keys = np.abs(np.asarray(values_local, dtype=np.float64)).view(np.uint64)
    n_total = int(all_reduce_sum(np.array([keys.size], dtype=np.int64))[0])
    if n_total == 0:
        return float("nan")
    rank = (n_total - 1) // 2
    prefix = np.uint64(0)
    for shift, bits in zip(SELECT_SHIFTS, SELECT_BITS):
        hs = shift + bits
        match = keys if hs >= 64 else keys[(keys >> np.uint64(hs)) == (prefix >> np.uint64(hs))]
        digit = ((match >> np.uint64(shift)) & np.uint64((1 << bits) - 1)).astype(np.int64)
        hist = all_reduce_sum(np.bincount(digit, minlength=2048).astype(np.int64))
        cum = np.cumsum(hist)
        d = int(np.searchsorted(cum, rank, side="right"))
        rank -= int(cum[d - 1]) if d > 0 else 0
        prefix = prefix | (np.uint64(d) << np.uint64(shift))
    return float(np.array([prefix], dtype=np.uint64).view(np.float64)[0])
