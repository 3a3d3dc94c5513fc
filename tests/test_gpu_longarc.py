"""Frame-window sharded long arc (configs[2]): the staged, exchange-driven solve must equal the ordinary
batched solve of the whole arc.  Multi-rank runs are emulated with several windows in one process on one GPU
(same stages, same buffers; reductions done by the test driver instead of NCCL); the real NCCL path is
exercised by tests/run_longarc_nccl.py under torchrun."""
import numpy as np
import pytest

import ba_oracle as o
from vinsat_b200 import _lib, longarc, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = _lib.Context(0)
    yield c
    c.close()


def _whole(ctx, pr, iters):
    b = _lib.Batch(ctx, _lib.concat_problems([pr]))
    lam = np.array([1e-4]); hist = []
    for it, init in iters:
        lam, ntr = b.ba_iterate(it, lam, initialize=init)
        hist.append((b.get_states().copy(), float(lam[0]), int(ntr[0])))
    b.close()
    return hist


@pytest.mark.parametrize("world,nseg", [(1, 1), (1, 4), (2, 1), (2, 3), (3, 2), (4, 5)])
def test_sharded_iterations_equal_whole_arc(ctx, world, nseg):
    pr = synth.make_problem(77, 120, 6)
    iters = [(0, True), (1, True), (9, True), (10, False), (11, False), (14, False), (19, False)]
    ref = _whole(ctx, pr, iters)
    la = longarc.LongArc(pr, ctxs=[ctx], world=world, n_segments=nseg)
    lam = 1e-4
    for k, (it, init) in enumerate(iters):
        lam, ntr = la.ba_iterate(it, lam, initialize=init)
        st = la.gather_states()
        s_ref, lam_ref, ntr_ref = ref[k]
        assert ntr == ntr_ref and lam == lam_ref, (world, nseg, it)
        assert np.abs(st[:, :3] - s_ref[:, :3]).max() < 1e-6, (world, nseg, it)      # 1 mm
        assert np.abs(st[:, 7:] - s_ref[:, 7:]).max() < 1e-9, (world, nseg, it)
        assert np.abs(st[:, 3:7] - s_ref[:, 3:7]).max() < 1e-10, (world, nseg, it)
    assert la.n_collectives > 0
    la.close()


def test_sharded_od_solve_converges_like_oracle(ctx):
    pr = synth.make_problem(78, 90, 5)
    la = longarc.LongArc(pr, ctxs=[ctx], world=3, n_segments=2)
    la.od_solve(20, 10, 1e-4)
    st = la.gather_states()
    ref, _, _ = o.od_solve(pr["states0"].copy(), pr["cum_rot"], pr["uv"], pr["xyz"], pr["ii"], pr["time_idx"], pr["intr"], pr["conf"])
    assert np.abs(st[:, :3] - ref[:, :3]).max() < 1e-3 and np.abs(st[:, 7:] - ref[:, 7:]).max() < 1e-6
    la.close()


def test_window_plan():
    assert longarc.plan_windows(10, 3) == [(0, 3), (3, 6), (6, 10)]
    pr = synth.make_problem(1, 12, 2)
    a, lo, hi = longarc.window_arrays(pr, 4, 8)
    assert a["frame_off"][1] == 6 and (lo, hi) == (1, 5)                    # one ghost each side
    assert a["obs_off"][1] == 8 and a["ii"].min() == 1 and a["ii"].max() == 4
    a, lo, hi = longarc.window_arrays(pr, 0, 4)
    assert a["frame_off"][1] == 5 and (lo, hi) == (0, 4)


def test_device_side_lm_and_cuda_graphs_equal_host_driven_iterations():
    """The long-arc iteration with the LM bookkeeping on the device and the head / extra trials replayed as CUDA graphs
    (world = 1 here; the NCCL collectives are captured the same way on N > 1, tests/run_longarc_nccl.py): same LM
    schedule and states as the host-driven `ba_iterate`, on a noisy problem whose LM loop rejects trials; a second solve
    after `reset_states` (all replays) is bit-identical to the first."""
    import torch
    from vinsat_b200 import _lib
    ctx = _lib.Context(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx._bound_to_torch = True
    try:
        pr = synth.make_problem(79, 150, 6, sigma_px=25.0, gap_max=40)
        ref = longarc.LongArc(pr, ctxs=[ctx], world=1, n_segments=3)
        lam, sched_ref = 1e-4, []
        for it in range(20):
            lam, ntr = ref.ba_iterate(it, lam, initialize=it < 10)
            sched_ref.append((lam, ntr))
        st_ref = ref.gather_states()
        ref.close()
        assert max(n for _, n in sched_ref) > 1, "no LM rejection: the extra-trial graph would be untested"
        la = longarc.LongArc(pr, ctxs=[ctx], world=1, n_segments=3)
        assert la.enable_graphs()
        for rep in range(2):
            sched = []
            for it in range(20):
                lam, ntr = la.ba_iterate_device_lm(it, 1e-4 if it == 0 else None, initialize=it < 10)
                sched.append((lam, ntr))
            st = la.gather_states()
            assert sched == sched_ref, rep
            assert np.abs(st - st_ref).max() < 1e-9, rep
            if rep == 0:
                st0 = st
                la.reset_states()
        assert np.array_equal(st, st0)
        assert la.graph_error is None and la.use_graphs and la.n_graph_replays > 20
        la.close()
    finally:
        torch.cuda.set_stream(torch.cuda.default_stream())
        ctx.close()


@pytest.mark.parametrize("world,nseg", [(1, 40), (2, 25), (3, 17)])
def test_sharded_arc_with_second_partition_level(ctx, monkeypatch, world, nseg):
    """With many segments per rank the gathered reduced chain is partitioned a second time (Level2, csrc/batch.h): forced here
    on a small arc (VINSAT_L2_MIN=16 => ~sqrt(S_total) level-2 segments); the whole-arc reference runs with level 2 off."""
    pr = synth.make_problem(81, 400, 5)
    iters = [(0, True), (9, True), (10, False), (11, False), (15, False), (19, False)]
    monkeypatch.setenv("VINSAT_L2_MIN", "0")
    ref = _whole(ctx, pr, iters)
    monkeypatch.setenv("VINSAT_L2_MIN", "16")
    la = longarc.LongArc(pr, ctxs=[ctx], world=world, n_segments=nseg)
    lam = 1e-4
    for k, (it, init) in enumerate(iters):
        lam, ntr = la.ba_iterate(it, lam, initialize=init)
        st = la.gather_states()
        s_ref, lam_ref, ntr_ref = ref[k]
        assert ntr == ntr_ref and lam == lam_ref, (world, nseg, it)
        assert np.abs(st[:, :3] - s_ref[:, :3]).max() < 1e-6, (world, nseg, it)
        assert np.abs(st[:, 7:] - s_ref[:, 7:]).max() < 1e-9, (world, nseg, it)
        assert np.abs(st[:, 3:7] - s_ref[:, 3:7]).max() < 1e-10, (world, nseg, it)
    la.close()


def test_long_arc_sums_in_two_stages(ctx):
    """Arcs of more than 4096 frames per problem take the two-stage (chunked) accept sums: same LM schedule and states as the
    sharded driver (whose sums are chunked per window) and convergence to the truth."""
    pr = synth.make_problem(82, 9000, 3, gap_max=2)
    iters = [(it, it < 10) for it in range(20)]
    ref = _whole(ctx, pr, iters)
    assert np.abs(ref[-1][0][:, :3] - pr["states_gt"][:, :3]).max() < 2.0
    la = longarc.LongArc(pr, ctxs=[ctx], world=2)
    lam = 1e-4
    for k, (it, init) in enumerate(iters):
        lam, ntr = la.ba_iterate(it, lam, initialize=init)
        assert (lam, ntr) == (ref[k][1], ref[k][2]), it
    st = la.gather_states()
    assert np.abs(st[:, :3] - ref[-1][0][:, :3]).max() < 1e-5
    la.close()
