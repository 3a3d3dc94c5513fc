"""Seeded synthetic OD problems: simulated LEO nadir arcs, landmark
observations and the perturbed initial guess of ``od_pipe.py:962-969``.

Used by the tests, ``smoke()`` and ``bench.py`` -- there are no datasets in this environment.
Every problem p is seeded with ``seed0 + p``.  All arrays are float64 / int64 NumPy.
"""
import numpy as np

from . import hostmath as hm

INTRINSICS_ROW0 = np.array([3547.85, 3547.85, 2304.0, 1296.0])   # landmarks/intrinsics.csv:1


def _accel(r):
    n = np.linalg.norm(r, axis=-1, keepdims=True)
    rm = np.array([[6.0, -1.5, -1.5], [6.0, -1.5, -1.5], [3.0, -4.5, -4.5]])
    return -(hm.MU / n ** 3) * r + (hm.J2C / n ** 7) * ((r * r) @ rm.T) * r


def _f(x):
    return np.concatenate([x[..., 3:6], _accel(x[..., 0:3])], axis=-1)


def orbit_step_batch(x, h=1.0):
    """trajgen_pipe.py:145-152 vectorised over leading dims (host input generation only)."""
    f1 = _f(x)
    f2 = _f(x + 0.5 * h * f1)
    f3 = _f(x + 0.5 * h * f2)
    f4 = _f(x + h * f3)
    return x + (h / 6.0) * (f1 + 2 * f2 + 2 * f3 + f4)


def random_polar_elements(rng):
    """trajgen_pipe.py:218."""
    return (600.0 + 6378.0 + (100 * rng.random() - 50), 0.0 + 0.01 * rng.random(),
            (np.pi / 2) + (0.2 * rng.random() - 0.1), 2 * np.pi * rng.random(),
            2 * np.pi * rng.random(), 2 * np.pi * rng.random())


def _project_truth(states, xyz, intr, ii):
    """Pinhole projection on the host, only to synthesise pixel measurements."""
    p, q = states[ii, 0:3], states[ii, 3:7]
    qn = q / np.linalg.norm(q, axis=-1, keepdims=True)
    v = np.concatenate([xyz - p, np.zeros((len(ii), 1))], axis=-1)
    pc = hm.quaternion_multiply(hm.quaternion_conjugate(qn), hm.quaternion_multiply(v, qn))
    d = 1.0 / np.maximum(pc[:, 2], 0.1)
    fx, fy, cx, cy = intr[ii].T
    return np.stack([fx * d * pc[:, 0] + cx, fy * d * pc[:, 1] + cy], axis=-1)


def make_batch(P, T, K, seed0=0, gap_max=20, sigma_px=1.0, lm_sigma_km=60.0, pos_sigma=100.0,
               rot_sigma=0.2, vel_frac=0.1, conf_lo=1.0, faithful_cum_rot=False, empty_frame_frac=0.0, orbit_fn=None):
    """P independent problems with T frames and K observations per frame each.

    `orbit_fn(x0 (P,6), n_steps) -> (P, n_steps+1, 6)`, when given, replaces the host RK4 loop that integrates the truth
    orbits at 1 Hz (e.g. `Context.orbit_propagate` for arcs of millions of frames, where the Python loop takes minutes).

    Returns a list of P dicts with keys: states0 (T,10) perturbed guess, states_gt (T,10),
    velocities (T,3) [gt, forward difference as process_ground_truths], cum_rot (T,4),
    uv (M,2), xyz (M,3) ECI km, ii (M,) int64 sorted, time_idx (T,) int64, intr (T,4), conf (M,).
    """
    rngs = [np.random.default_rng(seed0 + p) for p in range(P)]
    gaps = np.stack([r.integers(1, gap_max + 1, size=T - 1) for r in rngs]) if T > 1 else np.zeros((P, 0), int)
    time_idx = np.concatenate([np.zeros((P, 1), dtype=np.int64), np.cumsum(gaps, axis=1)], axis=1).astype(np.int64)
    x = np.stack([hm.oe2eci_values(*random_polar_elements(r)) for r in rngs])          # (P,6)
    D = int(time_idx[:, -1].max()) + 2
    pos_t = np.zeros((P, T, 3)); pos_t1 = np.zeros((P, T, 3)); vel_true = np.zeros((P, T, 3))
    ptr = np.zeros(P, dtype=np.int64)
    ar = np.arange(P)
    full = [] if faithful_cum_rot else None
    if orbit_fn is not None and full is None:
        traj = np.asarray(orbit_fn(x, D))                        # (P, D+1, 6): states at t = 0 .. D
        for p in range(P):
            pos_t[p], vel_true[p], pos_t1[p] = traj[p, time_idx[p], 0:3], traj[p, time_idx[p], 3:6], traj[p, time_idx[p] + 1, 0:3]
        D = 0
    for t in range(D):
        if full is not None:
            full.append(x[:, 0:3].copy())
        # frames whose time is t
        hit = (ptr < T) & (time_idx[ar, np.minimum(ptr, T - 1)] == t)
        if hit.any():
            pos_t[hit, ptr[hit]] = x[hit, 0:3]
            vel_true[hit, ptr[hit]] = x[hit, 3:6]
        xn = orbit_step_batch(x, 1.0)
        if hit.any():
            pos_t1[hit, ptr[hit]] = xn[hit, 0:3]
            ptr[hit] += 1
        x = xn
    out = []
    for p in range(P):
        rng = rngs[p]
        pos = pos_t[p]
        quat = hm.convert_pos_to_quaternion(pos)
        vel = (pos_t1[p] - pos) / 1.0                       # compute_velocity_from_pos at the frame times
        states_gt = np.concatenate([pos, quat, vel], axis=-1)
        if faithful_cum_rot:
            # od_pipe.py:945-961: 1 Hz omegas from the full nadir quaternion track, zero-padded per gap
            pos_full = np.stack([f[p] for f in full])
            q_full = hm.convert_pos_to_quaternion(pos_full)
            om = hm.compute_omega_from_quat(q_full, 1.0)
            N = int(gaps[p].max()) if T > 1 else 1
            omegas = np.zeros((T, N, 3))
            for i in range(1, T):
                g = time_idx[p, i] - time_idx[p, i - 1]
                omegas[i - 1, :g] = om[time_idx[p, i - 1]:time_idx[p, i]]
            rot = hm.quaternion_exp(omegas)
            cr = rot[:, 0]
            for k in range(1, N):
                cr = hm.quaternion_multiply(cr, rot[:, k])
            cum_rot = cr
        else:
            cum_rot = np.zeros((T, 4)); cum_rot[:, 3] = 1.0
            if T > 1:
                dq = hm.quaternion_multiply(hm.quaternion_conjugate(quat[:-1]), quat[1:])
                cum_rot[:-1] = dq / np.linalg.norm(dq, axis=-1, keepdims=True)
        # landmarks: K per frame around the sub-satellite point, on the sphere through it
        nobs = np.full(T, K, dtype=np.int64)
        if empty_frame_frac > 0:
            nobs[rng.random(T) < empty_frame_frac] = 0
        ii = np.repeat(np.arange(T, dtype=np.int64), nobs)
        M = len(ii)
        R = hm.nadir_frames(pos)                               # columns x_c y_c z_c
        off = rng.normal(0.0, lm_sigma_km, size=(M, 2))
        up = pos / np.linalg.norm(pos, axis=-1, keepdims=True)
        ground = up[ii] * 6371.0 + R[ii, :, 0] * off[:, 0:1] + R[ii, :, 1] * off[:, 1:2]
        xyz = ground / np.linalg.norm(ground, axis=-1, keepdims=True) * 6371.0
        intr = np.tile(INTRINSICS_ROW0, (T, 1))
        uv = _project_truth(states_gt, xyz, intr, ii) + rng.normal(0.0, sigma_px, size=(M, 2))
        conf = np.ones(M) if conf_lo >= 1.0 else rng.uniform(conf_lo, 1.0, size=M)
        # perturbed initial guess, od_pipe.py:962-969
        position = pos + rng.normal(0.0, pos_sigma, size=(T, 3))
        orientation = hm.quaternion_exp(hm.quaternion_log(quat) + rng.normal(0.0, rot_sigma, size=(T, 3)))
        vels = vel + rng.normal(0.0, 1.0, size=(T, 3)) * np.abs(vel).mean() * vel_frac
        states0 = np.concatenate([position, orientation, vels], axis=-1)
        out.append(dict(states0=states0, states_gt=states_gt, velocities=vel, cum_rot=cum_rot, uv=uv,
                        xyz=xyz, ii=ii, time_idx=time_idx[p].copy(), intr=intr, conf=conf,
                        vel_true=vel_true[p]))
    return out


def make_problem(seed, T, K, **kw):
    return make_batch(1, T, K, seed0=seed, **kw)[0]


def make_sequence(seed, n_orbit=1300, windows=((5, 150, 5), (400, 520, 5), (800, 900, 5), (1150, 1230, 5)),
                  dets_per_frame=6, sigma_px=1.0, lonlat_sigma_deg=1.0, conf_lo=0.7):
    """A synthetic detection sequence in the reference's on-disk formats:
      detections (n,6) rows [frame, lon_deg, lat_deg, u_px, v_px, conf] sorted by frame;
      orbit (n_orbit,12) rows [ECEF position in metres, dir(3), up(3), right(3)], one row per second.
    Pixels are the nadir-pointing pinhole projection of the landmark (ECI at t = frame) plus noise."""
    rng = np.random.default_rng(seed)
    x = hm.oe2eci_values(*random_polar_elements(rng))
    eci = np.zeros((n_orbit, 6))
    for t in range(n_orbit):
        eci[t] = x
        x = orbit_step_batch(x, 1.0)
    times = np.arange(n_orbit, dtype=np.float64)
    ecef = hm.eci_to_ecef(eci[:, :3], times)
    R = hm.nadir_frames(ecef)                                   # placeholders for dir / up / right
    orbit = np.concatenate([ecef * 1000.0, R[:, :, 2], -R[:, :, 1], -R[:, :, 0]], axis=1)
    quat = hm.convert_pos_to_quaternion(eci[:, :3])
    rows = []
    for (t0, t1, step) in windows:
        for t in range(t0, min(t1, n_orbit - 1), step):
            p = ecef[t]
            lon0 = np.degrees(np.arctan2(p[1], p[0]))
            lat0 = np.degrees(np.arctan2(p[2], np.hypot(p[0], p[1])))
            lon = lon0 + rng.normal(0, lonlat_sigma_deg, dets_per_frame) / max(np.cos(np.radians(lat0)), 0.2)
            lat = np.clip(lat0 + rng.normal(0, lonlat_sigma_deg, dets_per_frame), -89.0, 89.0)
            xyz = hm.convert_latlong_to_cartesian(lat, lon, np.full(dets_per_frame, float(t)))
            st = np.concatenate([eci[t, :3], quat[t], eci[t, 3:]])[None]
            uv = _project_truth(st, xyz, INTRINSICS_ROW0[None], np.zeros(dets_per_frame, dtype=np.int64))
            uv = uv + rng.normal(0, sigma_px, uv.shape)
            conf = rng.uniform(conf_lo, 1.0, dets_per_frame)
            for k in range(dets_per_frame):
                rows.append([float(t), lon[k], lat[k], uv[k, 0], uv[k, 1], conf[k]])
    return np.array(rows), orbit
