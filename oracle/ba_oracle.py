"""CPU oracle for the VINSat BA / OD hot path -- NumPy fp64 restatement.

TEST INFRASTRUCTURE ONLY.  This file is the *checker*: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  The product path (``vinsat_b200/``) never does; it
fails loudly when the CUDA library is missing.

It restates, in closed form, what the reference computes
with PyTorch autograd and dense LAPACK.  Citations are ``path:line`` under
``/root/reference/estimation``.  Parity is PINNED: ``tests/golden/make_golden.py``
runs the unmodified reference (imported via ``oracle/ref_loader.py``) on seeded
synthetic problems and stores its outputs under ``tests/golden/*.npz``;
``tests/test_oracle_vs_golden.py`` checks every function here against them.

Conventions: km, km/s, s, pixels; quaternions xyzw; state row =
[p(3) q(4) v(3)]; tangent row = [dp(3) dtheta(3) dv(3)].
"""
import numpy as np

MU = 398600.4418          # BA/BA_utils.py:883
J2C = 1.75553e10          # BA/BA_utils.py:883
R_MAT = np.array([[6.0, -1.5, -1.5],
                  [6.0, -1.5, -1.5],
                  [3.0, -4.5, -4.5]])   # BA/BA_utils.py:888-892 (non-textbook)
QUAT_COEFF = 100.0        # BA/BA_filtering.py:11
VEL_COEFF = 100.0         # BA/BA_filtering.py:12


# ----------------------------------------------------------------------------
# quaternion algebra (BA/BA_utils.py:949-1000, 19-28)
# ----------------------------------------------------------------------------
def quat_mul(q1, q2):
    """BA_utils.py:992-1000 (xyzw Hamilton product)."""
    x1, y1, z1, w1 = np.moveaxis(q1, -1, 0)
    x2, y2, z2, w2 = np.moveaxis(q2, -1, 0)
    w = w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2
    x = w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2
    y = w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2
    z = w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2
    return np.stack([x, y, z, w], axis=-1)


def quat_conj(q):
    """BA_utils.py:987-990."""
    return np.concatenate([-q[..., :3], q[..., 3:]], axis=-1)


def quat_exp(dtheta):
    """BA_utils.py:970-985: vec = d*sin(t/2)/(t+1e-16), w = cos(t/2); identity if t<1e-16."""
    t = np.linalg.norm(dtheta, axis=-1, keepdims=True)
    q = np.concatenate([dtheta * np.sin(t / 2) / (t + 1e-16), np.cos(t / 2)], axis=-1)
    ident = np.concatenate([np.zeros_like(dtheta), np.ones_like(t)], axis=-1)
    mask = (t < 1e-16).astype(np.float64)
    return ident * mask + q * (1 - mask)


def quat_log(q):
    """BA_utils.py:949-967."""
    q = np.clip(q / np.linalg.norm(q, axis=-1, keepdims=True), -1, 1)
    theta = 2 * np.arccos(q[..., 3])
    with np.errstate(divide="ignore", invalid="ignore"):
        n = q[..., :3] / np.sin(theta / 2)[..., None]
    return n * theta[..., None]


def attitude_jacobian(q):
    """BA_utils.py:19-28: 4x3 Gq with q (x) (d,0) = Gq(q) d."""
    q1, q2, q3, q0 = np.moveaxis(q, -1, 0)
    return np.stack([
        np.stack([q0, -q3, q2], axis=-1),
        np.stack([q3, q0, -q1], axis=-1),
        np.stack([-q2, q1, q0], axis=-1),
        np.stack([-q1, -q2, -q3], axis=-1)], axis=-2)


def right_mul_matrix(r):
    """4x4 M(r) with q (x) r = M(r) q (xyzw), from the product at BA_utils.py:992-1000."""
    x, y, z, w = np.moveaxis(r, -1, 0)
    return np.stack([
        np.stack([w, z, -y, x], axis=-1),
        np.stack([-z, w, x, y], axis=-1),
        np.stack([y, -x, w, z], axis=-1),
        np.stack([-x, -y, -z, w], axis=-1)], axis=-2)


def hat(v):
    x, y, z = np.moveaxis(v, -1, 0)
    o = np.zeros_like(x)
    return np.stack([np.stack([o, -z, y], -1), np.stack([z, o, -x], -1), np.stack([-y, x, o], -1)], -2)


def precompute_cum_rotations(omegas, dt):
    """BA_utils.py:278-288.  omegas (..., N, 3) -> ordered cumulative products (..., N, 4)."""
    rot = quat_exp(dt * omegas)
    out = [rot[..., 0, :]]
    for i in range(1, rot.shape[-2]):
        out.append(quat_mul(out[-1], rot[..., i, :]))
    return np.stack(out, axis=-2)


def compute_omega_from_quat(quat, dt):
    """BA_utils.py:1361-1367."""
    dq = quat_mul(quat_conj(quat[:-1]), quat[1:])
    dq = dq / np.linalg.norm(dq, axis=-1, keepdims=True)
    phi = quat_log(dq)
    return np.concatenate([phi / dt, np.zeros((1, 3))], axis=0)


# ----------------------------------------------------------------------------
# a1: landmark projection + Jacobian (BA_utils.py:7-50, 1052-1069)
# ----------------------------------------------------------------------------
def landmark_project(states, landmarks_xyz, intrinsics, ii, jacobian=True):
    """states (T,10), landmarks_xyz (M,3), intrinsics (T,4), ii (M,) -> uv (M,2)[, Jg (M,2,9)].

    p_c = conj(qn) (x) (X-p, 0) (x) qn (BA_utils.py:1052-1069); d = 1/max(Z,0.1) (:13);
    Jg = [-Pi R^T | 2 Pi hat(p_c) | 0]  (the factor 2 is the reference's).
    """
    ii = np.asarray(ii, dtype=np.int64)
    p = states[ii, 0:3]
    q = states[ii, 3:7]
    qn = q / np.linalg.norm(q, axis=-1, keepdims=True)
    v = np.concatenate([landmarks_xyz - p, np.zeros((len(ii), 1))], axis=-1)
    pc = quat_mul(quat_conj(qn), quat_mul(v, qn))[:, :3]
    fx, fy, cx, cy = np.moveaxis(intrinsics[ii], -1, 0)
    Xc, Yc, Zc = pc[:, 0], pc[:, 1], pc[:, 2]
    d = 1.0 / np.maximum(Zc, 0.1)
    uv = np.stack([fx * (d * Xc) + cx, fy * (d * Yc) + cy], axis=-1)
    if not jacobian:
        return uv
    live = (Zc >= 0.1).astype(np.float64)    # clamp has zero gradient below the bound
    M = len(ii)
    Pi = np.zeros((M, 2, 3))
    Pi[:, 0, 0] = fx * d
    Pi[:, 0, 2] = -fx * Xc * d * d * live
    Pi[:, 1, 1] = fy * d
    Pi[:, 1, 2] = -fy * Yc * d * d * live
    # R(qn)^T columns: rotate the basis vectors the same way the points are rotated
    eye = np.eye(3)
    Rt = np.stack([quat_mul(quat_conj(qn), quat_mul(
        np.broadcast_to(np.concatenate([eye[k], [0.0]]), (M, 4)), qn))[:, :3] for k in range(3)], axis=-1)
    Jg = np.zeros((M, 2, 9))
    Jg[:, :, 0:3] = -np.einsum("mij,mjk->mik", Pi, Rt)
    Jg[:, :, 3:6] = 2.0 * np.einsum("mij,mjk->mik", Pi, hat(pc))
    return uv, Jg


# ----------------------------------------------------------------------------
# a3: orbit dynamics, RK4 and its exact discrete state-transition matrix
# ----------------------------------------------------------------------------
def orbit_accel(r):
    """BA_utils.py:883-899: -mu r/|r|^3 + (J2/|r|^7) (R_MAT r^2) (.) r."""
    n = np.linalg.norm(r, axis=-1, keepdims=True)
    s = (r * r) @ R_MAT.T
    return -(MU / n ** 3) * r + (J2C / n ** 7) * s * r


def orbit_accel_grad(r):
    """G = d(accel)/dr (3x3), derivative of the expression above."""
    n = np.linalg.norm(r, axis=-1)[..., None, None]
    s = (r * r) @ R_MAT.T
    rr = r[..., :, None] * r[..., None, :]
    eye = np.eye(3)
    G = -MU * (eye / n ** 3 - 3.0 * rr / n ** 5)
    sr = (s * r)[..., :, None] * r[..., None, :]            # (s_i r_i) r_j
    G = G + J2C * (-7.0 * sr / n ** 9
                   + (2.0 * R_MAT * rr + s[..., :, None] * eye) / n ** 7)
    return G


def _f(x):
    return np.concatenate([x[..., 3:6], orbit_accel(x[..., 0:3])], axis=-1)


def rk4_step(x, h):
    """BA_utils.py:901-912 (classic RK4; h scalar or broadcastable)."""
    f1 = _f(x)
    f2 = _f(x + 0.5 * h * f1)
    f3 = _f(x + 0.5 * h * f2)
    f4 = _f(x + h * f3)
    return x + (h / 6.0) * (f1 + 2 * f2 + 2 * f3 + f4)


def _A_times(x, dX):
    """[[0,I],[G(r),0]] @ dX for dX (...,6,6)."""
    G = orbit_accel_grad(x[..., 0:3])
    return np.concatenate([dX[..., 3:6, :], G @ dX[..., 0:3, :]], axis=-2)


def rk4_step_stm(x, Phi, h):
    """RK4 on the augmented (state, variational) system = exact Jacobian of rk4_step."""
    hh = h[..., None] if np.ndim(h) else h
    f1 = _f(x); d1 = _A_times(x, Phi)
    x2 = x + 0.5 * h * f1; f2 = _f(x2); d2 = _A_times(x2, Phi + 0.5 * hh * d1)
    x3 = x + 0.5 * h * f2; f3 = _f(x3); d3 = _A_times(x3, Phi + 0.5 * hh * d2)
    x4 = x + h * f3; f4 = _f(x4); d4 = _A_times(x4, Phi + hh * d3)
    return (x + (h / 6.0) * (f1 + 2 * f2 + 2 * f3 + f4),
            Phi + (hh / 6.0) * (d1 + 2 * d2 + 2 * d3 + d4))


def step_schedule(gap, mode):
    """Step sizes for one frame gap.  mode 'step1s': [1]*gap (BA_utils.py:73-87);
    'skip100': [100]*(gap//100) + [gap%100] (BA_utils.py:52-71; the last step may be 0)."""
    if mode == "step1s":
        return [1.0] * int(gap)
    return [100.0] * int(gap // 100) + [float(gap % 100)]


def propagate_pairs(states, times, mode="step1s", stm=True):
    """For every frame i propagate (p_i, v_i) over gap_i = t_{i+1}-t_i (dummy gap 1 for the last
    frame, BA_utils.py:75).  Returns x_pred (T,6) and Phi (T,6,6) in [p,v] ordering."""
    T = states.shape[0]
    gaps = np.concatenate([np.diff(np.asarray(times, dtype=np.int64)), [1]])
    x = np.concatenate([states[:, 0:3], states[:, 7:10]], axis=-1).copy()
    Phi = np.broadcast_to(np.eye(6), (T, 6, 6)).copy() if stm else None
    if mode == "step1s":
        nsteps = gaps.copy()
        hs = None
    else:
        nsteps = gaps // 100 + 1
    for k in range(int(nsteps.max())):
        act = nsteps > k
        if not act.any():
            break
        if mode == "step1s":
            h = np.ones((int(act.sum()), 1))
        else:
            hop = gaps[act] // 100
            h = np.where(hop > k, 100.0, (gaps[act] % 100).astype(np.float64))[:, None]
        if stm:
            x[act], Phi[act] = rk4_step_stm(x[act], Phi[act], h)
        else:
            x[act] = rk4_step(x[act], h)
    return x, Phi


# ----------------------------------------------------------------------------
# a2/a4: dynamics residual, its Jacobian blocks, quaternion gradient and Hessian blocks
# ----------------------------------------------------------------------------
def predict(states, cum_rot, times, quat_coeff=QUAT_COEFF, vel_coeff=VEL_COEFF, jacobian=True,
            initialize=False, mode="step1s"):
    """Restates BA_utils.py:457-527 (`predict`).  states (T,10); cum_rot (T,4) = imu_meas[0,:,-1,6:10]
    (only that slice is consumed, BA_utils.py:295); times (T,) int.

    Returns dict: r_pred (T-1,7) [(T-1,6) zeros when initialize, :463-466]; with jacobian also
      Phi (T-1,6,6) [p,v order; Jf pair block = [D Phi_i | -D], D=diag(1,1,1,c_v,c_v,c_v)],
      qgrad (T,3) rot slots, Hq_diag (T,3,3), Hq_off (T-1,3,3) = Hq[i,i+1] (Hq[i+1,i] is its transpose).
    """
    T = states.shape[0]
    if initialize:
        out = {"r_pred": np.zeros((T - 1, 6))}
        if jacobian:
            out.update(Phi=np.zeros((T - 1, 6, 6)), qgrad=np.zeros((T, 3)),
                       Hq_diag=np.zeros((T, 3, 3)), Hq_off=np.zeros((T - 1, 3, 3)), initialize=True)
        return out
    x_pred, Phi = propagate_pairs(states, times, mode=mode, stm=jacobian)
    q = states[:, 3:7]
    q_pred = quat_mul(q, cum_rot)                                   # BA_utils.py:297
    dot = (q_pred[:-1] * q[1:]).sum(-1)
    r = np.concatenate([x_pred[:-1, 0:3] - states[1:, 0:3],
                        (x_pred[:-1, 3:6] - states[1:, 7:10]) * vel_coeff,
                        (quat_coeff * (1 - np.abs(dot)))[:, None]], axis=-1)   # BA_utils.py:478
    out = {"r_pred": r, "x_pred": x_pred, "q_pred": q_pred}
    if not jacobian:
        return out
    s = np.sign(dot)                                                # d|x|/dx; sign(0)=0 like torch
    M = right_mul_matrix(cum_rot)                                   # q (x) R = M q
    a = np.zeros((T, 4))
    a[:-1] += -quat_coeff * s[:, None] * np.einsum("tji,tj->ti", M[:-1], q[1:])   # M_i^T q_{i+1}
    a[1:] += -quat_coeff * s[:, None] * q_pred[:-1]                                # M_{i-1} q_{i-1}
    Gq = attitude_jacobian(q)
    g = np.einsum("tji,tj->ti", Gq, a)                              # Gq^T a
    beta = -(q * a).sum(-1)
    Hd = beta[:, None, None] * np.eye(3) + hat(g)
    Ho = -quat_coeff * s[:, None, None] * np.einsum("tji,tkj,tkl->til", Gq[:-1], M[:-1], Gq[1:])
    out.update(Phi=Phi[:-1], qgrad=g, Hq_diag=Hd, Hq_off=Ho, initialize=False)
    return out


# ----------------------------------------------------------------------------
# a5: robust weights (BA_filtering.py:21-25)
# ----------------------------------------------------------------------------
def lower_median(x):
    """torch.median semantics: lower of the two middle values for even counts."""
    x = np.sort(np.asarray(x).reshape(-1))
    return x[(len(x) - 1) // 2]


def robust_alpha(it):
    return min(max(1 - (2 * (it / 5) - 1), 1), 2)                   # BA_filtering.py:22


def robust_weights(r_obs, confidences, it):
    """r_obs (M,2) -> (w (M,), c_obs).  Reproduces the IEEE corner at alpha==2: x/0 -> inf (nan for
    0/0) and pow(inf|nan, 0) == 1 (SURVEY section 7 hard parts)."""
    alpha = robust_alpha(it)
    c = lower_median(np.abs(r_obs))
    with np.errstate(divide="ignore", invalid="ignore"):
        base = ((r_obs / c) ** 2) / abs(alpha - 2) + 1
        w = np.power(base, alpha / 2 - 1)
    if alpha / 2 - 1 == 0:
        w = np.ones_like(w)                                          # pow(., 0) == 1 even for nan/inf
    w = (w / c ** 2).mean(axis=-1)
    w = w / w.max() * confidences
    return w, c


# ----------------------------------------------------------------------------
# a6: normal equations in block-tridiagonal form (BA_filtering.py:28-48)
# ----------------------------------------------------------------------------
PV = np.array([0, 1, 2, 6, 7, 8])
ROT = np.array([3, 4, 5])


def assemble(Jg, w, r_obs, ii, T, dyn, Sigma, vel_coeff=VEL_COEFF):
    """Returns (Dg (T,9,9) WITHOUT the damping term, U (T-1,9,9) = block (i,i+1), b (T,9)).
    Block (i+1,i) is U[i]^T."""
    ii = np.asarray(ii, dtype=np.int64)
    Dg = np.zeros((T, 9, 9))
    b = np.zeros((T, 9))
    np.add.at(Dg, ii, np.einsum("mai,m,maj->mij", Jg, w, Jg))                 # :32-37
    np.add.at(b, ii, np.einsum("mai,m,ma->mi", Jg, w, r_obs))                 # :44
    U = np.zeros((max(T - 1, 0), 9, 9))
    if not dyn.get("initialize", False) and T > 1:
        Dv = np.array([1, 1, 1, vel_coeff, vel_coeff, vel_coeff], dtype=np.float64)
        Phi = dyn["Phi"]
        r6 = dyn["r_pred"][:, :6]
        DPhi = Dv[None, :, None] * Phi                                       # J block = D Phi
        ix = np.ix_(PV, PV)
        Dg[:-1, ix[0], ix[1]] += Sigma * np.einsum("tki,tkj->tij", DPhi, DPhi)
        Dg[1:, ix[0], ix[1]] += Sigma * np.diag(Dv * Dv)
        U[:, ix[0], ix[1]] += -Sigma * np.einsum("tki,k->tik", DPhi, Dv)
        b[:-1, PV] += -Sigma * np.einsum("tki,tk->ti", DPhi, r6)              # :47
        b[1:, PV] += Sigma * Dv * r6
        rx = np.ix_(ROT, ROT)
        Dg[:, rx[0], rx[1]] += Sigma * dyn["Hq_diag"]                          # :54  Sigma*Hq
        U[:, rx[0], rx[1]] += Sigma * dyn["Hq_off"]
        b[:, ROT] += -Sigma * dyn["qgrad"]                                     # :48
    return Dg, U, b


def damping(lamda):
    """torch.eye(n)*lamda is float32 (BA_filtering.py:54)."""
    return float(np.float32(lamda))


def dense_from_blocks(Dg, U, lam32):
    T = Dg.shape[0]
    A = np.zeros((9 * T, 9 * T))
    for i in range(T):
        A[9 * i:9 * i + 9, 9 * i:9 * i + 9] = Dg[i] + lam32 * np.eye(9)
        if i + 1 < T:
            A[9 * i:9 * i + 9, 9 * i + 9:9 * i + 18] = U[i]
            A[9 * i + 9:9 * i + 18, 9 * i:9 * i + 9] = U[i].T
    return A


def solve_blocktridiag(Dg, U, b, lam32, dense=False):
    """delta = A^{-1} b.  dense=True mirrors the reference literally (dense LU, BA_filtering.py:55);
    otherwise a banded LU with partial pivoting (bandwidth 17), same answer to ~1e-9 km."""
    T = Dg.shape[0]
    if dense:
        return np.linalg.solve(dense_from_blocks(Dg, U, lam32), b.reshape(-1)).reshape(T, 9)
    from scipy.linalg import solve_banded
    n = 9 * T
    kl = ku = 17
    ab = np.zeros((kl + ku + 1, n))
    eye = lam32 * np.eye(9)
    rows = np.arange(9)[:, None]
    cols = np.arange(9)[None, :]
    for i in range(T):
        blk = Dg[i] + eye
        r = 9 * i + rows
        c = 9 * i + cols
        ab[ku + r - c, c] = blk
        if i + 1 < T:
            c2 = c + 9
            ab[ku + r - c2, c2] = U[i]
            r2 = r + 9
            ab[ku + r2 - c, c] = U[i].T
    return solve_banded((kl, ku), ab, b.reshape(-1)).reshape(T, 9)


# ----------------------------------------------------------------------------
# a7: retraction + one full BA iteration (BA_filtering.py:4-98)
# ----------------------------------------------------------------------------
def retract(states, dpose):
    """BA_filtering.py:56-60."""
    pos = states[:, 0:3] + dpose[:, 0:3]
    vel = states[:, 7:10] + dpose[:, 6:9]
    rot = quat_mul(states[:, 3:7], quat_exp(dpose[:, 3:6]))
    rot = rot / np.linalg.norm(rot, axis=-1, keepdims=True)
    return np.concatenate([pos, rot, vel], axis=-1)


def ba_iteration(it, states, cum_rot, landmarks_uv, landmarks_xyz, ii, time_idx, intrinsics,
                 confidences, lamda_init, initialize=False, mode="step1s", dense=False, trace=None):
    """One damped Gauss-Newton / LM iteration.  Returns (states_new, lamda_next, last_hessian, info)."""
    T = states.shape[0]
    uv, Jg = landmark_project(states, landmarks_xyz, intrinsics, ii, jacobian=True)      # :15
    dyn = predict(states, cum_rot, time_idx, jacobian=True, initialize=initialize, mode=mode)  # :19
    r_obs = landmarks_uv - uv                                                           # :21
    w, c_obs = robust_weights(r_obs, confidences, it)                                   # :22-25
    Sigma = min(10000 * (it + 1) ** 2, 1000000)                                         # :26
    Dg, U, b = assemble(Jg, w, r_obs, ii, T, dyn, Sigma)
    r_pred = dyn["r_pred"]
    sq = np.sqrt(Sigma)
    init_residual = np.abs(np.concatenate([r_obs.reshape(-1), r_pred.reshape(-1) * sq])).mean()   # :51
    lamda = lamda_init
    ntrials = 0
    while True:
        lam32 = damping(lamda)
        dpose = solve_blocktridiag(Dg, U, b, lam32, dense=dense)                        # :54-55
        states_new = retract(states, dpose)                                             # :56-60
        uv1 = landmark_project(states_new, landmarks_xyz, intrinsics, ii, jacobian=False)
        r_pred1 = predict(states_new, cum_rot, time_idx, jacobian=False, initialize=initialize,
                          mode=mode)["r_pred"]
        r_obs1 = (landmarks_uv - uv1) * w[:, None]                                      # :66
        residual = np.abs(np.concatenate([r_obs1.reshape(-1), r_pred1.reshape(-1) * sq])).mean()
        ntrials += 1
        lamda = lamda * 10                                                              # :72
        if residual < init_residual:
            break
        if lamda > 1e4:
            break
    lamda_next = max(min(1e-1, lamda * 0.01), 1e-4)                                     # :79
    last_hessian = Dg[-1] + lam32 * np.eye(9)                                           # :97
    info = dict(uv=uv, Jg=Jg, r_obs=r_obs, w=w, c_obs=c_obs, Sigma=Sigma, Dg=Dg, U=U, b=b,
                dpose=dpose, init_residual=init_residual, residual=residual, ntrials=ntrials, dyn=dyn,
                lam32=lam32)
    if trace is not None:
        trace.append(info)
    return states_new, lamda_next, last_hessian, info


def od_solve(states, cum_rot, landmarks_uv, landmarks_xyz, ii, time_idx, intrinsics, confidences,
             num_iters=20, n_init=10, lamda_init=1e-4, mode="step1s", dense=False, trace=None):
    """One OD solve on one window = streaming_version's fixed schedule (od_pipe.py:918,1036-1040):
    num_iters BA iterations, the first n_init with initialize=True."""
    lam = lamda_init
    for it in range(num_iters):
        states, lam, H, _ = ba_iteration(it, states, cum_rot, landmarks_uv, landmarks_xyz, ii, time_idx,
                                         intrinsics, confidences, lam, initialize=(it < n_init),
                                         mode=mode, dense=dense, trace=trace)
    return states, lam, H


# ----------------------------------------------------------------------------
# a8: propagate_dynamics_init (BA_utils.py:89-129)
# ----------------------------------------------------------------------------
def propagate_dynamics_init(state, velocity, omega, tdiff, duration, dt=1.0):
    """state (10,), velocity (3,), omega (tdiff+duration, 3).  Returns (states_t (duration+1,10),
    velocities_t (duration+1,3), states_full, velocities_full) as BA_utils.py:114-129 (batch dim dropped).
    NOTE the reference propagates with the *separate* velocities argument, not states[7:10]."""
    def orbit_chain(p, v, n):
        x = np.concatenate([p, v])
        out = [x]
        for _ in range(n):
            x = rk4_step(x, dt)
            out.append(x)
        return np.stack(out)

    def rot_chain(q, w):
        out = [q]
        for k in range(len(w)):
            q = quat_mul(q, quat_exp(dt * w[k]))
            out.append(q)
        return np.stack(out)

    xb = orbit_chain(state[0:3], velocity, tdiff)
    qb = rot_chain(state[3:7], omega[:tdiff])
    xt = orbit_chain(xb[-1, 0:3], xb[-1, 3:6], duration)
    qt = rot_chain(qb[-1], omega[tdiff:tdiff + duration])
    states_beg = np.concatenate([xb[:, 0:3], qb, xb[:, 3:6]], axis=-1)[1:-1]
    states_t = np.concatenate([xt[:, 0:3], qt, xt[:, 3:6]], axis=-1)
    states_full = np.concatenate([states_beg, states_t], axis=0)
    vel_full = np.concatenate([xb[:, 3:6], xt[:, 3:6]], axis=0)[1:-1]
    return states_t, xt[:, 3:6], states_full, vel_full
