"""Dynamic work pool for independent OD problems (Monte-Carlo configs 2 and 4 of BASELINE.json).

The problems of a Monte-Carlo run are independent (the reference's own "Monte Carlo" is a sequential loop,
od_pipe.py:1063-1086), so no data-path collective exists -- but a STATIC split makes the whole job wait for its
slowest GPU (round 1: 4 of 8 GPUs ran the same kernels 15-20 % slower => 0.82 scaling efficiency).  Here the job
is a pool of equally sized CHUNKS (one chunk = one batched OD solve); every worker (one or more per GPU, each with
its own context / stream / device batch) pulls the next chunk index from a shared counter until the pool is
empty, so the job time is set by the mean GPU, not the slowest one.

The counter is the only shared state: `torch.distributed`'s key-value store (`store.add`, atomic on the rank-0
TCPStore server) when there are several processes, a lock-protected integer otherwise.  No tensor ever crosses
ranks.
"""
import threading


class LocalCounter:
    """Shared counter for the workers (threads) of one process."""

    def __init__(self):
        self._n = 0
        self._lock = threading.Lock()

    def next(self):
        with self._lock:
            n = self._n
            self._n += 1
        return n


class StoreCounter:
    """Shared counter for all ranks of a `torch.distributed` job: `store.add(key, 1)` is atomic and returns the new
    value.  `store` defaults to the process group's own store; `key` must be fresh for every pool."""

    def __init__(self, key, store=None):
        if store is None:
            import torch.distributed as dist
            store = dist.distributed_c10d._get_default_store()
        self.store = store
        self.key = key

    def next(self):
        return int(self.store.add(self.key, 1)) - 1


def make_counter(key, world_size=1, store=None):
    return StoreCounter(key, store) if world_size > 1 else LocalCounter()


def drain(counter, n_chunks, work, n_workers=1):
    """Runs `work(worker_index, chunk_index)` for chunk indices pulled from `counter` until `n_chunks` are handed
    out, on `n_workers` host threads.  Returns the list of (worker, chunk) pairs THIS process executed, in
    completion order.  The first exception of any worker is re-raised after all workers stopped."""
    done, errs = [], []
    lock = threading.Lock()

    def loop(w):
        try:
            while not errs:
                c = counter.next()
                if c >= n_chunks:
                    return
                work(w, c)
                with lock:
                    done.append((w, c))
        except Exception as e:          # surfaced below
            errs.append(e)

    if n_workers == 1:
        loop(0)
    else:
        ths = [threading.Thread(target=loop, args=(w,)) for w in range(n_workers)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
    if errs:
        raise errs[0]
    return done
