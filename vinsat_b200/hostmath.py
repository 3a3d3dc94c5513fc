"""Host-side (NumPy fp64) geometry used for input preparation, exactly where the reference
also runs NumPy on the host: frames/geodesy, quaternion algebra on small arrays, orbital
elements.  Heavy per-observation / per-frame arithmetic is NOT here -- it lives in the CUDA
library (``csrc/``).  Citations are ``path:line`` under ``/root/reference/estimation``.
"""
import numpy as np

MU = 398600.4418
J2C = 1.75553e10
THETA_G0_DEG = 280.16                       # BA/BA_utils.py:1172
OMEGA_EARTH_DEG = 360 / 86164.100352        # BA/BA_utils.py:1173
A_KM = 6378.137                             # BA/BA_utils.py:1178
B_KM = 6356.752                             # BA/BA_utils.py:1179
ECC = np.sqrt(1 - (B_KM ** 2 / A_KM ** 2))  # BA/BA_utils.py:1180


# -- quaternions (xyzw), BA/BA_utils.py:949-1000 ------------------------------------------
def quaternion_multiply(q1, q2):
    x1, y1, z1, w1 = np.moveaxis(np.asarray(q1, dtype=np.float64), -1, 0)
    x2, y2, z2, w2 = np.moveaxis(np.asarray(q2, dtype=np.float64), -1, 0)
    return np.stack([w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                     w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
                     w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2,
                     w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2], axis=-1)


def quaternion_conjugate(q):
    q = np.asarray(q, dtype=np.float64)
    return np.concatenate([-q[..., :3], q[..., 3:]], axis=-1)


def quaternion_exp(d_theta):
    d_theta = np.asarray(d_theta, dtype=np.float64)
    theta = np.linalg.norm(d_theta, axis=-1, keepdims=True)
    q = np.concatenate([d_theta * np.sin(theta / 2) / (theta + 1e-16), np.cos(theta / 2)], axis=-1)
    ident = np.concatenate([np.zeros_like(d_theta), np.ones_like(theta)], axis=-1)
    mask = (theta < 1e-16).astype(np.float64)
    return ident * mask + q * (1 - mask)


def quaternion_log(q):
    q = np.asarray(q, dtype=np.float64)
    q = np.clip(q / np.linalg.norm(q, axis=-1, keepdims=True), -1, 1)
    theta = 2 * np.arccos(q[..., 3])
    with np.errstate(divide="ignore", invalid="ignore"):
        n = q[..., :3] / np.sin(theta / 2)[..., None]
    return n * theta[..., None]


def compute_omega_from_quat(quat, dt):
    """BA/BA_utils.py:1361-1367."""
    dq = quaternion_multiply(quaternion_conjugate(quat[:-1]), quat[1:])
    dq = dq / np.linalg.norm(dq, axis=-1, keepdims=True)
    return np.concatenate([quaternion_log(dq) / dt, np.zeros((1, 3))], axis=0)


def precompute_cum_rotations(omegas, dt):
    """BA/BA_utils.py:278-288 on the host: (..., N, 3) -> (..., N, 4) cumulative rotations along axis -2."""
    rot = quaternion_exp(dt * np.asarray(omegas, dtype=np.float64))
    out = [rot[..., 0, :]]
    for i in range(1, rot.shape[-2]):
        out.append(quaternion_multiply(out[-1], rot[..., i, :]))
    return np.stack(out, axis=-2)


def compute_velocity_from_pos(pos, dt):
    """BA/BA_utils.py:1370-1373 (forward difference, zero last row)."""
    return np.concatenate([(pos[1:] - pos[:-1]) / dt, np.zeros((1, 3))], axis=0)


# -- frames / geodesy, BA/BA_utils.py:1172-1251 ---------------------------------------------
def ecef_to_eci(x_ecef, y_ecef, z_ecef, gmst=None, times=None):
    if times is not None:
        gmst = THETA_G0_DEG + OMEGA_EARTH_DEG * times
    theta = np.deg2rad(gmst)
    return (x_ecef * np.cos(theta) - y_ecef * np.sin(theta),
            x_ecef * np.sin(theta) + y_ecef * np.cos(theta), z_ecef)


def get_Rz(times):
    th = np.deg2rad(THETA_G0_DEG + OMEGA_EARTH_DEG * np.asarray(times, dtype=np.float64))
    zero, one = np.zeros_like(th), np.ones_like(th)
    return np.stack([np.stack([np.cos(th), np.sin(th), zero], -1),
                     np.stack([-np.sin(th), np.cos(th), zero], -1),
                     np.stack([zero, zero, one], -1)], -2)


def eci_to_ecef(r_eci, times):
    return (get_Rz(times) * r_eci[:, None, :]).sum(axis=-1)


def geodetic_to_ecef(latitude, longitude, altitude):
    phi, lam = np.deg2rad(latitude), np.deg2rad(longitude)
    N = A_KM / np.sqrt(1 - (ECC ** 2 * np.sin(phi) ** 2))
    return ((N + altitude) * np.cos(phi) * np.cos(lam),
            (N + altitude) * np.cos(phi) * np.sin(lam),
            ((B_KM ** 2 / A_KM ** 2) * N + altitude) * np.sin(phi))


def convert_latlong_to_cartesian(lat, long, times, altitude=None):
    if altitude is None:
        altitude = np.zeros(lat.shape[0])
    x, y, z = geodetic_to_ecef(lat, long, altitude)
    gmst_deg = THETA_G0_DEG + OMEGA_EARTH_DEG * times
    return np.stack(ecef_to_eci(x, y, z, gmst_deg), axis=-1)


def nadir_frames(pos):
    """Camera axes of the nadir attitude, BA/BA_utils.py:1276-1289: z_c=-p/|p|, r_c=zhat x z_c (unit),
    x_c=-r_c, y_c=r_c x z_c.  Returns R (...,3,3) with columns [x_c y_c z_c]."""
    zc = -pos / np.linalg.norm(pos, axis=-1, keepdims=True)
    rc = np.cross(np.broadcast_to(np.array([0.0, 0.0, 1.0]), zc.shape), zc)
    rc = rc / np.linalg.norm(rc, axis=-1, keepdims=True)
    yc = np.cross(rc, zc)
    return np.stack([-rc, yc, zc], axis=-1)


def convert_pos_to_quaternion(pos_eci):
    """BA/BA_utils.py:1276-1292 (scipy Rotation.from_matrix(...).as_quat(), xyzw)."""
    from scipy.spatial import transform
    return transform.Rotation.from_matrix(nadir_frames(np.asarray(pos_eci, dtype=np.float64))).as_quat()


# -- orbital elements, trajgen_pipe.py:13-64 ------------------------------------------------
def rotz(g):
    return np.array([[np.cos(g), -np.sin(g), 0], [np.sin(g), np.cos(g), 0], [0, 0, 1]])


def rotx(a):
    return np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])


def anom2E(nu, e):
    E = np.arccos((e + np.cos(nu)) / (1 + e * np.cos(nu)))
    if nu > np.pi:
        E = 2 * np.pi - E
    return E


def oe2eci_values(a, e, i, Omega, omega, nu, mu=MU):
    """trajgen_pipe.py:13-44 on scalars."""
    n = np.sqrt(mu / a ** 3)
    E = anom2E(nu, e)
    r_peri = np.array([a * (np.cos(E) - e), a * np.sqrt(1 - e ** 2) * np.sin(E), 0])
    v_peri = (a * n) / (1 - e * np.cos(E)) * np.array([-np.sin(E), np.sqrt(1 - e ** 2) * np.cos(E), 0])
    if i == 0 and e != 0:
        R1, R2, R3 = np.eye(3), np.eye(3), rotz(omega)
    elif e == 0 and i != 0:
        R1, R2, R3 = rotz(Omega), rotx(i), np.eye(3)
    elif i == 0 and e == 0:
        R1, R2, R3 = np.eye(3), np.eye(3), np.eye(3)
    else:
        R1, R2, R3 = rotz(Omega), rotx(i), rotz(omega)
    R = np.dot(R1, np.dot(R2, R3))
    return np.concatenate([np.dot(R, r_peri), np.dot(R, v_peri)])


# ---- right-hand sides used by trajgen_pipe's single-state helpers (trajgen_pipe.py:130-143,185-196) -------------------
_J2_ROWS = np.array([[6.0, -1.5, -1.5], [6.0, -1.5, -1.5], [3.0, -4.5, -4.5]])


def orbit_rhs(x, mu=MU, j2=J2C):
    """[v, a(r)] for the 6-state: a = -mu r/|r|^3 + j2/|r|^7 (C r.^2) .* r with the reference's coefficient rows."""
    r, v = x[:3], x[3:6]
    n = np.linalg.norm(r)
    return np.concatenate([v, -(mu / n ** 3) * r + (j2 / n ** 7) * np.dot(_J2_ROWS, r ** 2) * r])


def hat(v):
    """hat(v) w = v x w."""
    out = np.zeros((3, 3))
    out[0, 1], out[0, 2], out[1, 2] = -v[2], v[1], -v[0]
    return out - out.T


def quat_left_matrix_wxyz(q):
    """4x4 M with q (x) p = M p for scalar-FIRST quaternions: [[s, -v^T], [v, s I + hat(v)]]."""
    s, v = q[0], q[1:]
    M = np.empty((4, 4))
    M[0, 0], M[0, 1:], M[1:, 0] = s, -v, v
    M[1:, 1:] = s * np.eye(3) + hat(v)
    return M


def attitude_rhs(x, J):
    """[q_dot, omega_dot] of a torque-free rigid body with inertia J (scalar-first quaternion, already unit norm)."""
    q, w = x[:4], x[4:]
    q_dot = 0.5 * quat_left_matrix_wxyz(q)[:, 1:] @ w
    w_dot = -np.linalg.solve(J, np.cross(w, J @ w))
    return np.hstack((q_dot, w_dot))
