"""PipelinedSolver (two batched OD solves in flight per GPU) must return exactly what a plain Batch returns, and
the two-sided block-tridiagonal sweep must agree with the one-sided sweep and with the oracle's dense solve for
every problem length (1, 2, 3 frames ... odd / even middles)."""
import numpy as np
import pytest

from vinsat_b200 import _lib, synth
from vinsat_b200.pipeline import PipelinedSolver

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = _lib.Context(0)
    yield c
    c.close()


def test_pipelined_solves_equal_plain_batch(ctx):
    jobs = [_lib.concat_problems(synth.make_batch(6, 60, 8, seed0=100 * j)) for j in range(5)]
    ref = []
    b = _lib.Batch(ctx, jobs[0])
    for a in jobs:
        b.upload(a)
        b.od_solve(20, 10, 1e-4)
        ref.append(b.get_states().copy())
    b.close()
    pipe = PipelinedSolver(0, jobs[0], depth=2)
    outs = [np.empty_like(ref[0]) for _ in jobs]
    pipe.solve_many(jobs, outs)
    pipe.close()
    for r, x in zip(ref, outs):
        assert np.array_equal(r, x)


@pytest.mark.parametrize("two_sided", [True, False])
@pytest.mark.parametrize("lengths", [(1, 2, 3), (4, 5, 7, 8), (31, 32, 33), (2, 200, 3, 101)])
def test_two_sided_sweep_ragged_lengths(ctx, monkeypatch, lengths, two_sided):
    """one full (non-initialize) BA iteration on problems of ragged lengths: the LM step solves the same
    block-tridiagonal system as a dense LAPACK solve.  two_sided=True forces the Monte-Carlo path (two chains per
    problem meeting at the middle frame, columns assembled inside the sweep) that batches of >= 259 problems take;
    False leaves the small-batch policy (partitioned sweep, materialised system)."""
    if two_sided:
        monkeypatch.setenv("VINSAT_SEG_LEN", "1000000")        # one segment per problem => not partitioned
    prs = [synth.make_problem(900 + i, T, 6) for i, T in enumerate(lengths)]
    b = _lib.Batch(ctx, _lib.concat_problems(prs))
    lam, ntr = b.ba_iterate(12, 1e-4, initialize=False)
    dbg = b.debug_fetch()
    st = b.get_states()
    b.close()
    fo = np.concatenate([[0], np.cumsum(lengths)])
    for p, pr in enumerate(prs):
        T = lengths[p]
        D, U, rhs = dbg["D"][fo[p]:fo[p + 1]], dbg["U"][fo[p]:fo[p + 1]], dbg["rhs"][fo[p]:fo[p + 1]]
        A = np.zeros((9 * T, 9 * T))
        lam32 = float(np.float32(1e-4)) if ntr[p] == 1 else None
        for i in range(T):
            A[9 * i:9 * i + 9, 9 * i:9 * i + 9] = D[i]
            if i + 1 < T:
                A[9 * i:9 * i + 9, 9 * i + 9:9 * i + 18] = U[i]
                A[9 * i + 9:9 * i + 18, 9 * i:9 * i + 9] = U[i].T
        if lam32 is None:
            continue                      # rejected first trial: the fetched step belongs to a larger damping
        Al = A + lam32 * np.eye(9 * T)
        got = dbg["dpose"][fo[p]:fo[p + 1]].reshape(-1)
        # normwise backward error (the systems of short, weakly observed arcs are ill-conditioned, so compare
        # residuals, not solutions): as small as LAPACK's own
        bwd = np.abs(Al @ got - rhs.reshape(-1)).max() / (np.abs(Al).sum(1).max() * np.abs(got).max() + np.abs(rhs).max())
        assert bwd < 1e-13, (p, T, bwd)
        x = np.linalg.solve(Al, rhs.reshape(-1))
        assert np.allclose(got, x, rtol=1e-4, atol=1e-6 * np.abs(x).max()), (p, T, np.abs(got - x).max())
    assert np.all(np.isfinite(st))
