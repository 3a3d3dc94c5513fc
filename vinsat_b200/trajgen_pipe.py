"""Mirror of ``estimation/trajgen_pipe.py`` (identical in ``sim/orbit_gen.py:13-207``): orbital elements,
two-body+J2 orbit dynamics / RK4 step, rigid-body attitude step, trajectory generation.  Single-state
helpers are NumPy (they are 6- and 7-vectors); whole trajectories run on the device
(``vinsat_orbit_propagate``, a11 of SURVEY.md section 8).  Citations: estimation/trajgen_pipe.py:line.
"""
import numpy as np

from . import _lib, config
from . import hostmath as hm
from .hostmath import rotx, rotz, anom2E  # noqa: F401


class OrbitalElements:                                          # :4-11
    def __init__(self, a, e, i, Omega, omega, nu):
        self.a, self.e, self.i, self.Omega, self.omega, self.nu = a, e, i, Omega, omega, nu


def oe2eci(oe, mu=398600.4418):                                 # :13-44
    return hm.oe2eci_values(oe.a, oe.e, oe.i, oe.Omega, oe.omega, oe.nu, mu)


def orbit_dynamics(x_orbit, mu=398600.4418, J2=1.75553e10):     # :130-143
    r = x_orbit[:3]
    v = x_orbit[3:6]
    r_mat = np.array([[6, -1.5, -1.5], [6, -1.5, -1.5], [3, -4.5, -4.5]])
    v_dot = -(mu / np.linalg.norm(r) ** 3) * r + (J2 / np.linalg.norm(r) ** 7) * np.dot(r_mat, r ** 2) * r
    return np.concatenate([v, v_dot])


def orbit_step(xk, h):                                          # :145-152
    f1 = orbit_dynamics(xk)
    f2 = orbit_dynamics(xk + 0.5 * h * f1)
    f3 = orbit_dynamics(xk + 0.5 * h * f2)
    f4 = orbit_dynamics(xk + h * f3)
    return xk + (h / 6.0) * (f1 + 2 * f2 + 2 * f3 + f4)


def propagate_orbits(x0, n_steps, h=1.0, stride=1):
    """Batched `for k: x = orbit_step(x, h)` on the device: x0 (n,6) -> (n, n_steps/stride+1, 6)."""
    return _lib.default_context(config.device).orbit_propagate(x0, n_steps, stride, h)


# ---- attitude (3U CubeSat), :155-207 ------------------------------------------------------------------
m = 4.0
J = np.diag([(m / 12) * (.1 ** 2 + .34 ** 2), (m / 12) * (.1 ** 2 + .34 ** 2), (m / 12) * (.1 ** 2 + .1 ** 2)])


def hat(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


def L(q):
    s, v = q[0], q[1:]
    return np.concatenate([np.concatenate([[s], -v.T])[None], np.concatenate([v[:, None], s * np.eye(3) + hat(v)], axis=1)], axis=0)


H = np.concatenate([np.zeros((1, 3)), np.eye(3)], axis=0)


def G(q):
    return L(q) @ H


def attitude_dynamics(x_attitude, J=np.diag((1 / 3) * np.array([(.1 ** 2 + .34 ** 2), (.1 ** 2 + .34 ** 2), (.1 ** 2 + .1 ** 2)]))):
    q = x_attitude[:4]
    q /= np.linalg.norm(q)                                      # in place, as the reference (:187)
    omega = x_attitude[4:]
    q_dot = 0.5 * G(q) @ omega
    omega_dot = -np.linalg.solve(J, (hat(omega) @ J @ omega))
    return np.hstack((q_dot, omega_dot))


def attitude_step(xk, h):                                       # :198-207
    f1 = attitude_dynamics(xk)
    f2 = attitude_dynamics(xk + 0.5 * h * f1)
    f3 = attitude_dynamics(xk + 0.5 * h * f2)
    f4 = attitude_dynamics(xk + h * f3)
    xn = xk + (h / 6.0) * (f1 + 2 * f2 + 2 * f3 + f4)
    xn[:4] /= np.linalg.norm(xn[:4])
    return xn


def propagate_attitudes(x0, n_steps, h=1.0, stride=1):
    """Batched `for k: x = attitude_step(x, h)` on the device: x0 (n,7) -> (n, n_steps/stride+1, 7)."""
    return _lib.default_context(config.device).attitude_propagate(x0, n_steps, stride, h)


def generate_new_traj(orbit_type='polar', strict=False):
    """:209-251.  The reference draws random elements (:218/:222), simulates 3 h at 1 Hz and then FAILS at its
    `return` (it concatenates a 6xN and a 7xN array on axis 1).  With strict=True that ValueError
    is reproduced; by default the two trajectories are stacked on axis 0 (13 x N), which is what the return
    statement evidently meant.  Returns (traj (13,N), tsamp)."""
    if orbit_type == 'polar':
        oe = OrbitalElements(600.0 + 6378.0 + (100 * np.random.rand() - 50), 0.0 + 0.01 * np.random.rand(),
                             (np.pi / 2) + (0.2 * np.random.rand() - 0.1), 2 * np.pi * np.random.rand(),
                             2 * np.pi * np.random.rand(), 2 * np.pi * np.random.rand())
    else:
        oe = OrbitalElements(420.0 + 6378.0 + (100 * np.random.rand() - 50), 0.00034 + 0.01 * np.random.rand(),
                             (51.5 * np.pi / 180) + (0.2 * np.random.rand() - 0.1), 2 * np.pi * np.random.rand(),
                             2 * np.pi * np.random.rand(), 2 * np.pi * np.random.rand())
    x0_orbit = oe2eci(oe)
    tf = 3 * 60 * 60
    tsamp = np.arange(0, tf + 1, 1)
    xtraj_orbit = propagate_orbits(x0_orbit[None], len(tsamp) - 1, 1.0)[0].T          # (6, N) on the device
    q0 = np.ones(4) * 0.5
    q0 /= np.linalg.norm(q0)
    omega0 = 2 * (np.pi / 180) * np.ones(3) * 0.5
    xtraj_attitude = propagate_attitudes(np.hstack((q0, omega0))[None], len(tsamp) - 1, 1.0)[0].T   # (7, N), device
    if strict:
        return np.concatenate([xtraj_orbit, xtraj_attitude], axis=1), tsamp          # raises ValueError (:251)
    return np.concatenate([xtraj_orbit, xtraj_attitude], axis=0), tsamp
