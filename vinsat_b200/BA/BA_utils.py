"""Mirror of the reference module ``estimation/BA/BA_utils.py`` for the OD hot path: same function names,
argument meaning and return shapes (torch float64 CPU tensors in / out), computed by the CUDA library
through the C ABI (``include/vinsat_b200.h``).  Citations: path:line under <reference>/estimation.

Heavy operators -> CUDA: landmark_project, predict / predict_gpu, propagate_dynamics_init.
Small host helpers (quaternion algebra, frames / geodesy) stay NumPy/torch on the host exactly where the
reference also runs them on the host.
"""
import numpy as np
import torch

from .. import _lib, config
from .. import hostmath as hm

# ---- constants leaked into BA_utils by `from BA.utils import *` (BA/utils.py:89) and its own ----------
theta_G0_deg = hm.THETA_G0_DEG
omega_earth_deg_per_sec = hm.OMEGA_EARTH_DEG
a = hm.A_KM
b = hm.B_KM
e = hm.ECC


def _np(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().double().numpy()
    return np.asarray(x, dtype=np.float64)


def _ctx():
    return _lib.default_context(config.device)


# ---- a1 ---------------------------------------------------------------------------------------------
def landmark_project(poses, landmarks_xyz, intrinsics, ii, jacobian=True):
    """BA_utils.py:30-50.  poses (1,T,>=7), landmarks_xyz (1,M,3), intrinsics (1,T,4), ii (M,) ->
    landmark_est (1,M,2) [, Jg (M,2,9)]."""
    st = _np(poses)
    bsz = st.shape[0]
    assert bsz == 1, "the reference is batch-size-1 only (SURVEY 0.11)"
    st = st[0]
    if st.shape[1] < 10:
        st = np.concatenate([st, np.zeros((st.shape[0], 10 - st.shape[1]))], axis=1)
    xyz = _np(landmarks_xyz).reshape(-1, 3)
    intr = _np(intrinsics).reshape(-1, 4)
    ii = np.asarray(ii, dtype=np.int64).reshape(-1)
    out = _ctx().landmark_project(st, xyz, intr, ii, jacobian=jacobian)
    if jacobian:
        return torch.from_numpy(out[0]).reshape(bsz, -1, 2), torch.from_numpy(out[1])
    return torch.from_numpy(out).reshape(bsz, -1, 2)


# ---- a2-a4 ------------------------------------------------------------------------------------------
def _cum_rot_of(imu_meas):
    return _np(imu_meas)[0, :, -1, 6:10]            # BA_utils.py:295: only the last slice is consumed


def _predict(states, imu_meas, times, quat_coeff, vel_coeff, dt, jacobian, initialize, mode):
    st = _np(states)
    bsz, N = st.shape[0], st.shape[1]
    assert bsz == 1
    assert dt == 1, "the reference only ever calls predict with dt=1"
    num_res = (N - 1) * 6
    if initialize:                                   # BA_utils.py:463-466
        if jacobian:
            return (torch.zeros((bsz, N - 1, 6)), 0, 0, 0, 0, torch.zeros((bsz, num_res, N * 9)),
                    torch.zeros((bsz, N * 9, N * 9)), torch.zeros((bsz, N, 9)))
        return torch.zeros((bsz, N - 1, 6)), 0, 0
    cr = _cum_rot_of(imu_meas)
    d = _ctx().predict(st[0], cr, np.asarray(times, dtype=np.int64), quat_coeff, vel_coeff, jacobian=jacobian,
                       mode=mode, want_x_pred=True)
    res_pred = torch.from_numpy(d["r_pred"])[None]
    q_pred = hm.quaternion_multiply(st[0][:, 3:7], cr)
    x = d["x_pred"]
    pose_pred = torch.from_numpy(np.concatenate([x[:, :3], q_pred, x[:, 3:]], axis=1))[None]
    vel_pred = torch.from_numpy(x[:, 3:].copy())[None]
    if not jacobian:
        return res_pred, pose_pred, vel_pred
    # dense outputs with the reference's shapes, filled from the nonzero blocks
    PV = np.array([0, 1, 2, 6, 7, 8])
    Dv = np.array([1, 1, 1, vel_coeff, vel_coeff, vel_coeff], dtype=np.float64)
    Jf = np.zeros((N - 1, 6, N, 9))
    Hq = np.zeros((N, 9, N, 9))
    idx = np.arange(N - 1)
    Jf[idx[:, None, None], np.arange(6)[None, :, None], idx[:, None, None], PV[None, None, :]] = Dv[None, :, None] * d["Phi"]
    Jf[idx[:, None], np.arange(6)[None, :], idx[:, None] + 1, PV[None, :]] = -Dv[None, :]
    r3 = np.arange(3)
    fi = np.arange(N)
    Hq[fi[:, None, None], 3 + r3[None, :, None], fi[:, None, None], 3 + r3[None, None, :]] = d["Hq_diag"]
    Hq[idx[:, None, None], 3 + r3[None, :, None], idx[:, None, None] + 1, 3 + r3[None, None, :]] = d["Hq_off"]
    Hq[idx[:, None, None] + 1, 3 + r3[None, :, None], idx[:, None, None], 3 + r3[None, None, :]] = np.swapaxes(d["Hq_off"], 1, 2)
    qgrad = np.zeros((N, 9))
    qgrad[:, 3:6] = d["qgrad"]
    return (res_pred, pose_pred, vel_pred, 0, 0, torch.from_numpy(Jf.reshape(1, num_res, N * 9)),
            torch.from_numpy(Hq.reshape(1, N * 9, N * 9)), torch.from_numpy(qgrad)[None])


def predict(states, imu_meas, times, quat_coeff, vel_coeff, dt=1, jacobian=True, initialize=False):
    """BA_utils.py:457-527 (1 s RK4 steps)."""
    return _predict(states, imu_meas, times, quat_coeff, vel_coeff, dt, jacobian, initialize, _lib.MODE_STEP1S)


def predict_gpu(states, imu_meas, times, quat_coeff, vel_coeff, dt=1, jacobian=True, initialize=False):
    """BA_utils.py:529-602 (steps of up to 100 s, `propagate_orbit_dynamics_skip`)."""
    return _predict(states, imu_meas, times, quat_coeff, vel_coeff, dt, jacobian, initialize, _lib.MODE_SKIP100)


# ---- a8 ---------------------------------------------------------------------------------------------
def propagate_dynamics_init(states, velocities, omega, tdiff, duration, dt):
    """BA_utils.py:114-129.  states (1,10), velocities (1,3), omega (1,tdiff+duration,3) ->
    (states_t (1,duration+1,10), velocities_t (1,duration+1,3), states_full, velocities_full)."""
    st = _np(states).reshape(10)
    v0 = _np(velocities).reshape(3)
    om = _np(omega).reshape(-1, 3)[:tdiff + duration]
    assert om.shape[0] == tdiff + duration, "omega must cover tdiff + duration steps"
    chain = _ctx().propagate_chain(st, v0, om, float(dt))               # rows 0 .. tdiff+duration
    states_t = chain[tdiff:]
    states_beg = chain[1:tdiff]                                           # [:, 1:-1] of the first chain
    states_full = np.concatenate([states_beg, states_t], axis=0)
    vel_all = np.concatenate([chain[:tdiff + 1, 7:10], chain[tdiff:, 7:10]], axis=0)[1:-1]
    return (torch.from_numpy(states_t.copy())[None], torch.from_numpy(states_t[:, 7:10].copy())[None],
            torch.from_numpy(states_full)[None], torch.from_numpy(vel_all)[None])


# ---- scatter (BA_utils.py:1376-1382; torch_scatter semantics) ----------------------------------------
def safe_scatter_add_vec(b, ii, n, mean=False):
    ii = torch.as_tensor(ii, dtype=torch.long)
    size = list(b.shape)
    size[1] = n
    out = torch.zeros(size, dtype=b.dtype).index_add_(1, ii, b)
    if mean:
        cnt = torch.zeros(n, dtype=b.dtype).index_add_(0, ii, torch.ones(len(ii), dtype=b.dtype)).clamp_(min=1)
        out = out / cnt.view([1, n] + [1] * (b.dim() - 2))
    return out


def safe_scatter_add_mat(A, ii, jj, n, m):
    return safe_scatter_add_vec(A, torch.as_tensor(ii) * m + torch.as_tensor(jj), n * m)


# ---- quaternion algebra on torch tensors (BA_utils.py:949-1000, 19-28) -------------------------------
def quaternion_multiply(q1, q2):
    return torch.from_numpy(hm.quaternion_multiply(_np(q1), _np(q2)))


def quaternion_exp(d_theta):
    return torch.from_numpy(hm.quaternion_exp(_np(d_theta)))


def quaternion_log(q):
    return torch.from_numpy(hm.quaternion_log(_np(q)))


def quaternion_conjugate(q):
    return torch.from_numpy(hm.quaternion_conjugate(_np(q)))


def attitude_jacobian(q):
    q = _np(q)
    q1, q2, q3, q0 = np.moveaxis(q, -1, 0)
    return torch.from_numpy(np.stack([np.stack([q0, -q3, q2], -1), np.stack([q3, q0, -q1], -1),
                                      np.stack([-q2, q1, q0], -1), np.stack([-q1, -q2, -q3], -1)], -2))


def precompute_cum_rotations(omegas, dt):
    """BA_utils.py:278-288: (1,T,N,3) -> (1,T,N,4), on the device (`vinsat_precompute_cum_rotations`)."""
    om = _np(omegas)
    return torch.from_numpy(_ctx().precompute_cum_rotations(om[0], dt)[None])


def cum_rotations_from_quat(gt_quat_full, time_idx, dt):
    """The slice of precompute_cum_rotations that `predict` reads (BA_utils.py:295), straight from the full-rate
    quaternions: compute_omega_from_quat + per-gap product on the device (od_pipe.py:944-953 without the
    (1,T,N,3) zero-padded omegas tensor).  Returns (cum_rot (T,4), gt_omega (n_full,3))."""
    cr, om = _ctx().cum_rotations(_np(gt_quat_full), np.asarray(time_idx), dt, want_omega=True)
    return cr, torch.from_numpy(om)


def compute_omega_from_quat(quat, dt):
    return torch.from_numpy(hm.compute_omega_from_quat(_np(quat), dt))


# ---- frames / geodesy (host NumPy, identical call signatures) ------------------------------------------
compute_velocity_from_pos = hm.compute_velocity_from_pos
ecef_to_eci = hm.ecef_to_eci
eci_to_ecef = hm.eci_to_ecef
get_Rz = hm.get_Rz
geodetic_to_ecef = hm.geodetic_to_ecef
convert_latlong_to_cartesian = hm.convert_latlong_to_cartesian
convert_pos_to_quaternion = hm.convert_pos_to_quaternion
deg_to_rad = np.deg2rad


# ---- (f)4: prior-regularised variant ---------------------------------------------------------------------------------
def prior_gpu(states, prop_states, vel_coeff, quat_coeff, hessian_state_t, hessian_rot_t, jacobian=True, initialize=False):
    """BA_utils.py:604-676.  states / prop_states (1,N,10), hessian_state_t (1,N,6,6), hessian_rot_t (1,N,3,3).
    Returns res (1,N,7) [, Jp (1,6N,9N), Hqp (1,9N,9N), qgrad (1,N,9)] like the reference: the dense matrices are built
    from the device's diagonal blocks (the reference's are exactly zero elsewhere).  initialize=True returns the
    reference's zero tensors (:611-614)."""
    st = _np(states)
    assert st.shape[0] == 1, "the reference is batch-size-1 only (SURVEY 0.11)"
    N = st.shape[1]
    if initialize:
        if jacobian:
            return (torch.zeros((1, N, 6)), torch.zeros((1, N * 6, N * 9)), torch.zeros((1, N * 9, N * 9)),
                    torch.zeros((1, N, 9)))
        return torch.zeros((1, N, 6))
    out = _ctx().prior(st[0], _np(prop_states)[0], vel_coeff, quat_coeff, _np(hessian_state_t)[0], _np(hessian_rot_t)[0],
                       jacobian=jacobian)
    if not jacobian:
        return torch.from_numpy(out)[None]
    r, Jp, Hqp, qg = out
    Jd = np.zeros((N, 6, N, 9))
    Hd = np.zeros((N, 9, N, 9))
    idx = np.arange(N)
    Jd[idx, :, idx, :] = Jp
    Hd[idx, :, idx, :] = Hqp
    return (torch.from_numpy(r)[None], torch.from_numpy(Jd.reshape(1, N * 6, N * 9)),
            torch.from_numpy(Hd.reshape(1, N * 9, N * 9)), torch.from_numpy(qg)[None])


def propagate_dynamics_cov_init(states, velocities, hessian, omega, tdiff, duration, dt):
    """BA_utils.py:222-248: states (1,10), velocities (1,3), hessian (1,9,9), omega (1,tdiff+duration,3) ->
    (states_t (1,duration+1,10), velocities_t (1,duration+1,3), hessian_state_t (1,duration+1,6,6),
    hessian_rot_t (1,duration+1,3,3)); one device call (orbit + attitude chains with their covariances)."""
    st, hs, hr = _ctx().propagate_chain_cov(_np(states)[0], _np(velocities)[0], _np(hessian).reshape(-1, 9, 9)[0],
                                            _np(omega)[0], int(tdiff), int(duration), float(dt))
    t = torch.from_numpy
    return t(st)[None], t(np.ascontiguousarray(st[:, 7:]))[None], t(hs)[None], t(hr)[None]
