"""Generate golden vectors by running the UNMODIFIED VINSat reference (CPU `predict` path).

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py [ba|helpers|indexing|streaming|all]
Outputs ``tests/golden/*.npz`` (small, committed).  The reference is pure Python, so it is imported
through ``oracle/ref_loader.py`` (three stub modules); nothing is copied.
Inputs come from ``vinsat_b200.synth`` (seeded) or are built here; every array the tests need is
stored so the tests never touch /root/reference.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import ref_loader  # noqa: E402
from vinsat_b200 import synth  # noqa: E402

ns = ref_loader.load()
U, F = ns.BA_utils, ns.BA_filtering
PV = np.array([0, 1, 2, 6, 7, 8])


def t(x):
    return torch.tensor(np.asarray(x))


def imu_from_cum_rot(cum_rot, time_idx):
    T = len(time_idx)
    N = int(np.diff(time_idx).max()) if T > 1 else 1
    imu = torch.zeros(1, T, N, 10, dtype=torch.float64)
    imu[0, :, -1, 6:10] = t(cum_rot)
    return imu


def ref_predict_blocks(states, cum_rot, time_idx):
    """Run reference `predict(jacobian=True)` and cut the dense outputs into their nonzero blocks."""
    T = states.shape[0]
    imu = imu_from_cum_rot(cum_rot, time_idx)
    r, _, _, _, _, Jf, Hq, qg = U.predict(t(states)[None].clone(), imu, time_idx, 100, 100, jacobian=True)
    Jf = Jf[0].detach().numpy().reshape(T - 1, 6, T, 9)
    Hq = Hq[0].detach().numpy().reshape(T, 9, T, 9)
    qg = qg[0].detach().numpy()
    # structure claims: everything outside these blocks is exactly zero
    Jmask = np.ones_like(Jf, dtype=bool)
    Hmask = np.ones_like(Hq, dtype=bool)
    for i in range(T - 1):
        Jmask[i, :, i] = False
        Jmask[i, :, i + 1] = False
    for i in range(T):
        for j in (i - 1, i, i + 1):
            if 0 <= j < T:
                Hmask[i, 3:6, j, 3:6] = False
    assert np.abs(Jf[Jmask]).max() == 0 and np.abs(Hq[Hmask]).max() == 0
    assert np.abs(Jf[:, :, :, 3:6]).max() == 0 and np.abs(qg[:, PV]).max() == 0
    return dict(r_pred=r[0].detach().numpy(),
                Jf_self=np.stack([Jf[i][:, i][:, PV] for i in range(T - 1)]),      # D Phi_i
                Jf_next=np.stack([Jf[i][:, i + 1][:, PV] for i in range(T - 1)]),  # -D
                qgrad=qg[:, 3:6],
                Hq_diag=np.stack([Hq[i, 3:6, i, 3:6] for i in range(T)]),
                Hq_off=np.stack([Hq[i, 3:6, i + 1, 3:6] for i in range(T - 1)]),
                Hq_low=np.stack([Hq[i + 1, 3:6, i, 3:6] for i in range(T - 1)]))


def run_ref_ba(pr, num_iters=20, n_init=10, lamda_init=1e-4):
    T = pr["states0"].shape[0]
    imu = imu_from_cum_rot(pr["cum_rot"], pr["time_idx"])
    states = t(pr["states0"])[None]
    vel = t(pr["velocities"])[None]
    lam = lamda_init
    hist, lams, hess = [], [], []
    for it in range(num_iters):
        states, _, lam, H = F.BA(it, states.detach(), vel, imu, t(pr["uv"])[None], t(pr["xyz"])[None],
                                 pr["ii"], pr["time_idx"], t(pr["intr"])[None], t(pr["conf"]), 1e-3, 1e-3,
                                 lam, t(pr["states_gt"][:, :7]), initialize=(it < n_init))
        hist.append(states[0].detach().numpy().copy())
        lams.append(lam)
        hess.append(H[0].detach().numpy().copy())
    return np.stack(hist), np.array(lams), np.stack(hess)


def golden_ba():
    cases = {
        # name: (seed, T, K, kwargs)
        "ba_T30": (3, 30, 8, dict(faithful_cum_rot=True)),
        "ba_T24_ragged": (11, 24, 5, dict(faithful_cum_rot=True, empty_frame_frac=0.25, conf_lo=0.8,
                                           sigma_px=2.0)),
        "ba_T40_noisy": (7, 40, 3, dict(faithful_cum_rot=True, sigma_px=40.0, gap_max=60)),
    }
    for name, (seed, T, K, kw) in cases.items():
        pr = synth.make_problem(seed, T, K, **kw)
        st = t(pr["states0"])[None]
        uv, Jg = U.landmark_project(st, t(pr["xyz"])[None], t(pr["intr"])[None], pr["ii"], jacobian=True)
        blocks = ref_predict_blocks(pr["states0"], pr["cum_rot"], pr["time_idx"])
        # skip-mode forward propagation (the propagator `predict_gpu` uses, BA_utils.py:52-71)
        sp, sv = U.propagate_orbit_dynamics_skip(st[:, :, :3], st[:, :, 7:], pr["time_idx"], 1)
        hist, lams, hess = run_ref_ba(pr)
        out = {("in_" + k): v for k, v in pr.items()}
        out.update(uv=uv[0].detach().numpy(), Jg=Jg.detach().numpy(), skip_pos=sp[0].numpy(), skip_vel=sv[0].numpy(),
                   states_hist=hist, lamda_hist=lams, hessian_hist=hess,
                   **{("pred_" + k): v for k, v in blocks.items()})
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        err = np.abs(hist[-1][:, :3] - pr["states_gt"][:, :3]).max()
        print(name, "saved; final |pos-gt|max = %.3f km; lamdas" % err, lams)


def golden_ba_large():
    """T=200, K=10 (the no-pivot block LU is most exposed at T >= 200 in the full phase, cond 1e11-5e12).
    Histories only (the dense Jf / Hq blocks are pinned by the small cases).  ~2 min of reference time."""
    pr = synth.make_problem(17, 200, 10)
    hist, lams, hess = run_ref_ba(pr)
    out = {("in_" + k): v for k, v in pr.items()}
    out.update(states_hist=hist, lamda_hist=lams, hessian_hist=hess)
    np.savez_compressed(os.path.join(HERE, "ba_T200.npz"), **out)
    print("ba_T200 saved; final |pos-gt|max = %.3f km; lamdas" % np.abs(hist[-1][:, :3] - pr["states_gt"][:, :3]).max(), lams)


def golden_ba_skip():
    """20 BA iterations through `predict_gpu` (BA_filtering.py:16-17), i.e. with propagate_orbit_dynamics_skip
    (steps of up to 100 s).  The reference takes this branch when torch.cuda.is_available(); there is no GPU in the
    build container, so availability is forced and `Tensor.cuda()` is made the identity -- the arithmetic is the same
    ATen code on the CPU device.  Gaps up to 250 s so that several 100 s hops and zero-length last hops occur."""
    pr = synth.make_problem(23, 36, 6, gap_max=250)
    avail, cuda = torch.cuda.is_available, torch.Tensor.cuda
    torch.cuda.is_available = lambda: True
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        hist, lams, hess = run_ref_ba(pr)
    finally:
        torch.cuda.is_available, torch.Tensor.cuda = avail, cuda
    out = {("in_" + k): v for k, v in pr.items()}
    out.update(states_hist=hist, lamda_hist=lams, hessian_hist=hess)
    np.savez_compressed(os.path.join(HERE, "ba_T36_skip100.npz"), **out)
    print("ba_T36_skip100 saved; gaps max", int(np.diff(pr["time_idx"]).max()), "final |pos-gt|max = %.3f km; lamdas"
          % np.abs(hist[-1][:, :3] - pr["states_gt"][:, :3]).max(), lams)


def golden_long_gap():
    """Gaps > 100 s so that skip mode takes several 100 s hops and a zero-length last hop."""
    rng = np.random.default_rng(5)
    T = 8
    gaps = np.array([100, 250, 1, 99, 300, 101, 17])
    time_idx = np.concatenate([[0], np.cumsum(gaps)]).astype(np.int64)
    pr = synth.make_problem(5, T, 2)
    states = pr["states_gt"] + rng.normal(0, 1e-3, size=(T, 10))
    st = t(states)[None]
    sp, sv = U.propagate_orbit_dynamics_skip(st[:, :, :3], st[:, :, 7:], time_idx, 1)
    p1, v1 = U.propagate_orbit_dynamics(st[:, :, :3], st[:, :, 7:], time_idx, 1)
    np.savez_compressed(os.path.join(HERE, "long_gap.npz"), states=states, time_idx=time_idx,
                        skip_pos=sp[0].numpy(), skip_vel=sv[0].numpy(), step_pos=p1[0].numpy(), step_vel=v1[0].numpy())
    print("long_gap saved; |skip-step1s| max [m] =", np.abs(sp - p1).max().item() * 1e3)


def golden_helpers():
    rng = np.random.default_rng(0)
    tp = ns.trajgen_pipe
    out = {}
    q1 = rng.normal(size=(16, 4)); q2 = rng.normal(size=(16, 4))
    q1 /= np.linalg.norm(q1, axis=-1, keepdims=True)
    d = rng.normal(size=(16, 3)) * 0.3
    d[0] = 0.0
    d[1] = 1e-18
    out.update(q1=q1, q2=q2, d=d,
               qmul=U.quaternion_multiply(t(q1), t(q2)).numpy(), qexp=U.quaternion_exp(t(d)).numpy(),
               qlog=U.quaternion_log(t(q1)).numpy(), qconj=U.quaternion_conjugate(t(q1)).numpy(),
               Gq=U.attitude_jacobian(t(q1)).numpy())
    om = rng.normal(size=(1, 5, 7, 3)) * 0.01
    om[0, 2, 4:] = 0.0
    out.update(omegas=om, cum_rot=U.precompute_cum_rotations(t(om), 1.0).numpy())
    qs = U.quaternion_exp(t(np.cumsum(rng.normal(size=(12, 3)) * 0.02, axis=0)))
    out.update(qtrack=qs.numpy(), omega_from_quat=U.compute_omega_from_quat(qs, 1.0).numpy())
    pos = rng.normal(size=(9, 3)) * 4000 + np.array([3000.0, -2000.0, 1000.0])
    out.update(pos=pos, nadir_quat=U.convert_pos_to_quaternion(pos), vel_fd=U.compute_velocity_from_pos(pos, 1.0))
    times = np.arange(9) * 137.0
    ex, ey, ez = U.ecef_to_eci(pos[:, 0], pos[:, 1], pos[:, 2], times=times)
    lat = rng.uniform(-80, 80, size=9); lon = rng.uniform(-180, 180, size=9)
    out.update(times=times, ecef2eci=np.stack([ex, ey, ez], -1), eci2ecef=U.eci_to_ecef(pos, times),
               lat=lat, lon=lon, latlon_cart=U.convert_latlong_to_cartesian(lat, lon, times))
    # orbit dynamics / RK4 (torch version used by BA and NumPy version used by the simulator)
    x = np.concatenate([pos / np.linalg.norm(pos, axis=-1, keepdims=True) * 6978.0, rng.normal(size=(9, 3)) * 4], -1)
    f, _ = U.orbit_dynamics(t(x))
    out.update(x=x, f_torch=f.numpy(), rk4_1s=U.RK4(t(x), 0, 1.0).numpy(), rk4_100s=U.RK4(t(x), 0, 100.0).numpy(),
               f_np=np.stack([tp.orbit_dynamics(xx) for xx in x]), step_np=np.stack([tp.orbit_step(xx, 1.0) for xx in x]))
    oe = tp.OrbitalElements(6978.0 + 13.0, 0.004, np.pi / 2 + 0.03, 1.1, 2.2, 4.4)
    oe2 = tp.OrbitalElements(6800.0, 0.0, 0.9, 0.3, 0.0, 1.0)
    out.update(oe=np.array([oe.a, oe.e, oe.i, oe.Omega, oe.omega, oe.nu]), oe_eci=tp.oe2eci(oe),
               oe2=np.array([oe2.a, oe2.e, oe2.i, oe2.Omega, oe2.omega, oe2.nu]), oe2_eci=tp.oe2eci(oe2))
    xa = np.concatenate([np.ones(4) * 0.5, 2 * (np.pi / 180) * np.array([0.5, -0.3, 0.8])])
    traj = [xa]
    for _ in range(5):
        traj.append(tp.attitude_step(traj[-1].copy(), 1.0))
    out.update(att_traj=np.stack(traj))
    # propagate_dynamics_init (BA_utils.py:114-129)
    pr = synth.make_problem(2, 6, 2)
    st0 = pr["states_gt"][0]
    omega = rng.normal(size=(1, 9, 3)) * 0.002
    s_t, v_t, s_full, v_full = U.propagate_dynamics_init(t(st0)[None], t(pr["velocities"][0])[None], t(omega), 4, 5, 1)
    out.update(pdi_state=st0, pdi_vel=pr["velocities"][0], pdi_omega=omega[0], pdi_states_t=s_t[0].numpy(),
               pdi_vel_t=v_t[0].numpy(), pdi_states_full=s_full[0].numpy(), pdi_vel_full=v_full[0].numpy())
    # scatter (torch_scatter semantics as used at BA_utils.py:1379-1382) -- pins the stub too
    b = rng.normal(size=(1, 10, 3)); idx = np.array([0, 0, 2, 2, 2, 3, 5, 5, 5, 5])
    out.update(sc_b=b, sc_idx=idx, sc_sum=U.safe_scatter_add_vec(t(b), t(idx), 7).numpy(),
               sc_mean=U.safe_scatter_add_vec(t(b), t(idx), 7, mean=True).numpy())
    np.savez_compressed(os.path.join(HERE, "helpers.npz"), **out)
    print("helpers saved")


def make_mgrs_table():
    """sim/getMGRS.py of the reference: the ordered zone table (tests/golden/mgrs_table.json)."""
    import importlib.util
    import json
    spec = importlib.util.spec_from_file_location("ref_mgrs", "/root/reference/sim/getMGRS.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    table = m.getMGRS()
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "mgrs_table.json"), "w") as f:
        json.dump([[k] + [int(x) for x in v] for k, v in table.items()], f, separators=(",", ":"))


def make_batch_runner_commands():
    """eval/batch_runner.py is a module-level loop of subprocess.call()s: run it with the call recorded, not executed."""
    import json
    import runpy
    import subprocess
    calls = []
    real = subprocess.call
    subprocess.call = lambda cmd, **kw: calls.append([cmd, kw]) or 0
    try:
        runpy.run_path("/root/reference/eval/batch_runner.py", run_name="__main__")
    finally:
        subprocess.call = real
    with open(os.path.join(HERE, "batch_runner_commands.json"), "w") as f:
        json.dump(calls, f, indent=0)
    print("batch_runner_commands saved:", len(calls), "calls")


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("ba", "all"):
        golden_ba()
        golden_long_gap()
    if what in ("ba_large", "all"):
        golden_ba_large()
    if what in ("ba_skip", "all"):
        golden_ba_skip()
    if what in ("helpers", "all"):
        golden_helpers()
        make_mgrs_table()
        make_batch_runner_commands()
    if what in ("indexing", "streaming", "all"):
        import make_golden_streaming  # noqa: F401  (kept separate: slower)
        make_golden_streaming.main(what)
