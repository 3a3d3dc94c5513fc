"""Device input preparation: compute_omega_from_quat + precompute_cum_rotations and
the batched attitude simulator against their host restatements (vinsat_b200/hostmath.py follows
BA/BA_utils.py:949-1000, 278-288, 1361-1367; trajgen_pipe.attitude_step follows trajgen_pipe.py:155-207)."""
import numpy as np
import pytest

from vinsat_b200 import _lib, hostmath as hm, trajgen_pipe as tp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = _lib.Context(0)
    yield c
    c.close()


def _quat_walk(n, seed):
    rng = np.random.default_rng(seed)
    q = np.zeros((n, 4)); q[0] = [0.1, -0.2, 0.3, 0.9]; q[0] /= np.linalg.norm(q[0])
    for k in range(1, n):
        q[k] = hm.quaternion_multiply(q[k - 1], hm.quaternion_exp(rng.normal(0, 0.01, 3)))
    return q


def test_cum_rotations_match_host(ctx):
    n, dt = 400, 1.0
    q = _quat_walk(n, 3)
    rng = np.random.default_rng(4)
    time_idx = np.concatenate([[0], np.cumsum(rng.integers(1, 21, 37))]).astype(np.int64)
    time_idx = time_idx[time_idx < n]
    cr, om = ctx.cum_rotations(q, time_idx, dt, want_omega=True)
    om_ref = hm.compute_omega_from_quat(q, dt)
    assert np.allclose(om, om_ref, rtol=1e-12, atol=1e-15)
    T = len(time_idx)
    N = int(np.max(np.diff(time_idx)))
    omegas = np.zeros((T, N, 3))
    for i in range(1, T):                                              # od_pipe.py:948-952
        omegas[i - 1, :time_idx[i] - time_idx[i - 1]] = om_ref[time_idx[i - 1]:time_idx[i]]
    rot = hm.quaternion_exp(dt * omegas)
    cum = [rot[:, 0]]
    for j in range(1, N):                                              # BA_utils.py:283-286
        cum.append(hm.quaternion_multiply(cum[-1], rot[:, j]))
    cum = np.stack(cum, axis=1)
    assert np.allclose(cr, cum[:, -1], rtol=0, atol=1e-14)
    assert np.array_equal(cr[-1], [0.0, 0.0, 0.0, 1.0])
    # general form: all prefixes
    full = ctx.precompute_cum_rotations(omegas, dt)
    assert full.shape == (T, N, 4) and np.allclose(full, cum, rtol=0, atol=1e-14)


def test_precompute_cum_rotations_vs_reference_golden(ctx):
    """the reference's own outputs (tests/golden/make_golden.py ran BA_utils.precompute_cum_rotations /
    compute_omega_from_quat from /root/reference)"""
    import torch
    from conftest import load_golden
    from vinsat_b200.BA import BA_utils as U
    g = load_golden("helpers")
    got = U.precompute_cum_rotations(torch.tensor(g["omegas"]), 1.0).numpy()
    assert got.shape == g["cum_rot"].shape and np.abs(got - g["cum_rot"]).max() < 1e-14
    cr, om = U.cum_rotations_from_quat(torch.tensor(g["qtrack"]), np.arange(0, len(g["qtrack"]), 3), 1.0)
    assert np.abs(om.numpy() - g["omega_from_quat"]).max() < 1e-12


def test_cum_rotations_reject_bad_time_idx(ctx):
    q = _quat_walk(50, 1)
    with pytest.raises(Exception):
        ctx.cum_rotations(q, np.array([0, 10, 10, 20]), 1.0)
    with pytest.raises(Exception):
        ctx.cum_rotations(q, np.array([0, 10, 80]), 1.0)


def test_attitude_propagate_matches_attitude_step(ctx):
    rng = np.random.default_rng(9)
    x0 = np.zeros((5, 7))
    x0[:, :4] = rng.normal(size=(5, 4)); x0[:, :4] /= np.linalg.norm(x0[:, :4], axis=1, keepdims=True)
    x0[:, 4:] = rng.normal(0, 0.05, (5, 3))
    out = ctx.attitude_propagate(x0, 300, stride=3, h=1.0)
    assert out.shape == (5, 101, 7)
    for i in range(5):
        x = x0[i].copy()
        for k in range(300):
            if k % 3 == 0:
                assert np.allclose(out[i, k // 3], x, rtol=0, atol=1e-12), (i, k)
            x = _host_attitude_step(x.copy(), 1.0)
        assert np.allclose(out[i, 100], x, rtol=0, atol=1e-12)


def _host_attitude_step(x, h):
    """attitude_step (trajgen_pipe.py:198-207) on the host, from the mirror's derivative function."""
    f = tp.attitude_dynamics
    f1 = f(x); f2 = f(x + 0.5 * h * f1); f3 = f(x + 0.5 * h * f2); f4 = f(x + h * f3)
    xn = x + (h / 6.0) * (f1 + 2 * f2 + 2 * f3 + f4)
    xn[:4] /= np.linalg.norm(xn[:4])
    return xn


def test_single_step_helpers_run_on_device_and_match_reference_goldens(ctx):
    """trajgen_pipe.orbit_step / attitude_step are one-step device calls: against the reference's own outputs."""
    from conftest import load_golden
    g = load_golden("helpers")
    got = np.stack([tp.orbit_step(x, 1.0) for x in g["x"]])
    assert np.abs(got - g["step_np"]).max() <= 1e-13 * np.abs(g["step_np"]).max()
    xa = g["att_traj"][0].copy()
    for k in range(5):
        xa = tp.attitude_step(xa.copy(), 1.0)
        assert np.abs(xa - g["att_traj"][k + 1]).max() < 1e-14


def test_generate_new_traj_shapes(ctx):
    np.random.seed(0)
    traj, tsamp = tp.generate_new_traj('polar')
    assert traj.shape == (13, 10801) and len(tsamp) == 10801
    assert np.allclose(np.linalg.norm(traj[6:10], axis=0), 1.0, atol=1e-12)
    with pytest.raises(ValueError):
        np.random.seed(0)
        tp.generate_new_traj('polar', strict=True)
