"""CPU oracle for SURVEY section 8 (f)4: `prior_gpu` (BA/BA_utils.py:604-676), the covariance propagation
`propagate_dynamics_cov_init` (BA/BA_utils.py:130-248) and `BA_reg` (BA/BA_filtering.py:100-210) -- NumPy closed forms.

TEST INFRASTRUCTURE ONLY (same rules as ba_oracle.py).  PARITY PINNED by tests/golden/ba_reg.npz, generated from the
unmodified reference by tests/golden/make_golden_reg.py (tests/test_oracle_vs_golden.py::test_reg_*).

Facts about the reference that the closed forms reproduce:
  * `prior_gpu`'s state residual is H_i e_i with e_i = [p_prop - p ; (v_prop - v) vel_coeff]; its Jacobian block is
    -H_i diag(1,1,1,vc,vc,vc) on the position / velocity columns, rotation columns zero (:626, :651-654).
  * the rotation residual is quat_coeff (1 - |q_prop^T Gq(q_prop) H_rot Gq(q)^T q|) with both Gq DETACHED (:623-627).
    Gq(q)^T q == 0 identically, so the scalar inside |.| is rounding noise (~1e-17): the residual is quat_coeff, its
    gradient / Hessian are ~1e-15 |H_rot|.  They are computed anyway, with the sign of that noise.
  * BA_reg calls prior_gpu(states, prior, quat_coeff_prior, vel_coeff_prior, ...) -- the two coefficients land in each
    other's parameter (both are 1 at the linearisation); in the LM trial it passes (quat_coeff_prior, vel_coeff) =
    (1, 100), so the trial's rotation prior residual is 100 per frame against 1 per frame in `init_residual`, and the
    trial's `predict` gets quat_coeff 1 instead of 100 (:166-171).  All reproduced.
  * the covariance propagation feeds an xyzw quaternion into `qtoQ`, which reads it as scalar-FIRST (:203-206).
"""
import numpy as np

import ba_oracle as o

PV = o.PV
ROT = o.ROT


def prior(states, prop_states, vel_coeff, quat_coeff, Hs, Hr, jacobian=True):
    """-> r (N,7) [, Jp (N,6,9) diagonal blocks, Hqp (N,9,9) diagonal blocks, qgrad (N,9)]."""
    N = states.shape[0]
    p, q, v = states[:, :3], states[:, 3:7], states[:, 7:]
    pp, qp, vp = prop_states[:, :3], prop_states[:, 3:7], prop_states[:, 7:]
    e = np.concatenate([pp - p, (vp - v) * vel_coeff], axis=1)
    r_state = np.einsum("nij,nj->ni", Hs, e)                                            # :626
    Gq = o.attitude_jacobian(q)                                                         # (N,4,3)
    Gqp = o.attitude_jacobian(qp)
    a = np.einsum("nr,nri->ni", qp, Gqp)                                                # q_prop^T Gq(q_prop)  (~0)
    w = np.einsum("nri,nr->ni", Gq, q)                                                  # Gq(q)^T q            (~0)
    s = np.einsum("ni,nij,nj->n", a, Hr, w)
    r_rot = quat_coeff * (1 - np.abs(s))                                                # :627
    r = np.concatenate([r_state, r_rot[:, None]], axis=1)
    if not jacobian:
        return r
    Dv = np.array([1, 1, 1, vel_coeff, vel_coeff, vel_coeff], dtype=np.float64)
    Jp = np.zeros((N, 6, 9))
    Jp[:, :, PV] = -Hs * Dv[None, None, :]
    # gradient of the rotation residual with both Gq detached: d/dq = -quat_coeff sign(s) Gq (H_rot^T a); projected by
    # the (attached) Gq(q): qgrad_rot = Gq^T g4; Hessian = d(Gq(q)^T g4)/dq Gq(q) with g4 held fixed
    u = np.einsum("nij,ni->nj", Hr, a)                                                  # H_rot^T a ... (a^T H)_j
    g4 = -quat_coeff * np.sign(s)[:, None] * np.einsum("nrj,nj->nr", Gq, u)
    qgrad = np.zeros((N, 9))
    qgrad[:, ROT] = np.einsum("nri,nr->ni", Gq, g4)
    g0, g1, g2, g3 = g4.T
    dG = np.stack([np.stack([-g3, -g2, g1, g0], -1), np.stack([g2, -g3, -g0, g1], -1), np.stack([-g1, g0, -g3, g2], -1)], 1)
    Hqp = np.zeros((N, 9, 9))
    Hqp[np.ix_(np.arange(N), ROT, ROT)] = np.einsum("nic,ncj->nij", dG, Gq)
    return r, Jp, Hqp, qgrad


def _qtoQ_scalar_first(q):
    """BA_utils.py:193-206 applied to whatever 4-vector it is given (s = q[0], v = q[1:])."""
    s, v = q[0], q[1:]
    L = np.zeros((4, 4))
    L[0, 0] = s; L[0, 1:] = -v; L[1:, 0] = v
    L[1:, 1:] = s * np.eye(3) + np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
    Tm = np.diag([1.0, -1.0, -1.0, -1.0])
    H = np.vstack([np.zeros((1, 3)), np.eye(3)])
    return H.T @ (Tm @ L) @ (Tm @ L) @ H


def propagate_dynamics_cov_init(state, velocity, hessian, omega, tdiff, duration, dt=1.0):
    """BA_utils.py:222-248 (batch dim dropped) -> states_t (duration+1,10), velocities_t (duration+1,3),
    hessian_state_t (duration+1,6,6), hessian_rot_t (duration+1,3,3)."""
    hs = hessian[np.ix_(PV, PV)]
    cov_s = np.linalg.inv(hs)
    cov_r = np.linalg.inv(hessian[3:6, 3:6])
    x = np.concatenate([state[:3], velocity])
    q = state[3:7].copy()

    def step(x, q, cov_s, cov_r, w):
        Phi = np.eye(6)
        xn, Phi = o.rk4_step_stm(x, Phi, dt)                                            # compute_orbit_jacobian (:130-133)
        cov_s = Phi @ cov_s @ Phi.T
        Y = _qtoQ_scalar_first(o.quat_exp((-dt * w)[None])[0])                          # compute_rot_jacobian (:208-211)
        qn = o.quat_mul(q[None], o.quat_exp((dt * w)[None]))[0]
        cov_r = Y @ cov_r @ Y.T
        return xn, qn, cov_s, cov_r

    for k in range(tdiff):
        x, q, cov_s, cov_r = step(x, q, cov_s, cov_r, omega[k])
    xs, qs, cs, cr = [x], [q], [cov_s], [cov_r]
    for k in range(duration):
        x, q, cov_s, cov_r = step(x, q, cov_s, cov_r, omega[tdiff + k])
        xs.append(x); qs.append(q); cs.append(cov_s); cr.append(cov_r)
    xs, qs = np.stack(xs), np.stack(qs)
    states_t = np.concatenate([xs[:, :3], qs, xs[:, 3:]], axis=1)
    return states_t, xs[:, 3:], np.linalg.inv(np.stack(cs)), np.linalg.inv(np.stack(cr))


def ba_reg_iteration(it, states, states_prior, Hs, Hr, cum_rot, landmarks_uv, landmarks_xyz, ii, time_idx, intrinsics,
                     confidences, lamda_init, mode="step1s", dense=False):
    """BA_reg (BA_filtering.py:100-210), initialize=False.  Returns (states_new, lamda_next, last_hessian, info)."""
    T = states.shape[0]
    quat_coeff_prior, vel_coeff_prior, vel_coeff = 1.0, 1.0, 100.0
    uv, Jg = o.landmark_project(states, landmarks_xyz, intrinsics, ii, jacobian=True)
    dyn = o.predict(states, cum_rot, time_idx, jacobian=True, mode=mode)
    r_prior, Jp, Hqp, qgradp = prior(states, states_prior, quat_coeff_prior, vel_coeff_prior, Hs, Hr)     # :122
    r_obs = landmarks_uv - uv
    w, c_obs = o.robust_weights(r_obs, confidences, it)
    Sigma = min(10000 * (it + 1) ** 2, 1000000)
    Dg, U, b = o.assemble(Jg, w, r_obs, ii, T, dyn, Sigma)
    Dg = Dg + np.einsum("nki,nkj->nij", Jp, Jp) + Hqp                                   # :165  JpTJp + Hqp
    b = b - np.einsum("nki,nk->ni", Jp, r_prior[:, :6]) - qgradp                         # :159-160
    sq = np.sqrt(Sigma)
    init_residual = np.abs(np.concatenate([r_obs.reshape(-1), dyn["r_pred"].reshape(-1) * sq, r_prior.reshape(-1)])).mean()
    lamda = lamda_init
    ntrials = 0
    while True:
        lam32 = o.damping(lamda)
        dpose = o.solve_blocktridiag(Dg, U, b, lam32, dense=dense)
        states_new = o.retract(states, dpose)
        uv1 = o.landmark_project(states_new, landmarks_xyz, intrinsics, ii, jacobian=False)
        r_pred1 = o.predict(states_new, cum_rot, time_idx, quat_coeff=quat_coeff_prior, vel_coeff=vel_coeff, jacobian=False,
                            mode=mode)["r_pred"]                                       # :169: quat_coeff_prior !
        r_prior1 = prior(states_new, states_prior, quat_coeff_prior, vel_coeff, Hs, Hr, jacobian=False)    # :170: (1, 100)
        r_obs1 = (landmarks_uv - uv1) * w[:, None]
        residual = np.abs(np.concatenate([r_obs1.reshape(-1), r_pred1.reshape(-1) * sq, r_prior1.reshape(-1)])).mean()
        ntrials += 1
        lamda = lamda * 10
        if residual < init_residual or lamda > 1e4:
            break
    lamda_next = max(min(1e-1, lamda * 0.01), 1e-4)
    last_hessian = Dg[-1] + lam32 * np.eye(9)
    info = dict(Dg=Dg, U=U, b=b, dpose=dpose, ntrials=ntrials, init_residual=init_residual, residual=residual, r_prior=r_prior,
                c_obs=c_obs, w=w)
    return states_new, lamda_next, last_hessian, info
