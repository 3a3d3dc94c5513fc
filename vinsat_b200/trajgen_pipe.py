"""Mirror of ``estimation/trajgen_pipe.py`` (identical in ``sim/orbit_gen.py:13-207``): orbital elements,
two-body+J2 orbit dynamics / RK4 step, rigid-body attitude step, trajectory generation.  The RK4 steps and whole
trajectories run on the device (``vinsat_orbit_propagate`` / ``vinsat_attitude_propagate``, a11 of SURVEY.md section 8);
the derivative functions are thin views of ``hostmath``.  Citations: estimation/trajgen_pipe.py:line.
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
    """d/dt [r, v] with the reference's two-body + J2 acceleration (non-textbook coefficient rows 6,-1.5,-1.5 /
    6,-1.5,-1.5 / 3,-4.5,-4.5): a = -mu r/|r|^3 + J2/|r|^7 (C r^2) .* r.  Host helper for single states;
    trajectories go through `propagate_orbits` (device)."""
    x = np.asarray(x_orbit, dtype=np.float64)
    return hm.orbit_rhs(x[:6], mu, J2)


def orbit_step(xk, h):                                          # :145-152
    """One classic RK4 step of the 6-state, on the device (`vinsat_orbit_propagate` with one trajectory, one step)."""
    return propagate_orbits(np.asarray(xk, dtype=np.float64)[None, :6], 1, float(h))[0, 1]


def propagate_orbits(x0, n_steps, h=1.0, stride=1):
    """Batched `for k: x = orbit_step(x, h)` on the device: x0 (n,6) -> (n, n_steps/stride+1, 6)."""
    return _lib.default_context(config.device).orbit_propagate(x0, n_steps, stride, h)


# ---- attitude (3U CubeSat), :155-207 ------------------------------------------------------------------
m = 4.0
J = np.diag([(m / 12) * (.1 ** 2 + .34 ** 2), (m / 12) * (.1 ** 2 + .34 ** 2), (m / 12) * (.1 ** 2 + .1 ** 2)])
_J_DEFAULT = (1 / 3) * np.array([(.1 ** 2 + .34 ** 2), (.1 ** 2 + .34 ** 2), (.1 ** 2 + .1 ** 2)])    # :185 default argument
H = np.vstack([np.zeros((1, 3)), np.eye(3)])                    # :172


def hat(v):                                                     # :160-163: cross-product matrix, hat(v) w = v x w
    return hm.hat(np.asarray(v, dtype=np.float64))


def L(q):                                                       # :165-168: left-multiplication matrix, scalar-first q
    return hm.quat_left_matrix_wxyz(np.asarray(q, dtype=np.float64))


def G(q):                                                       # :177-178
    return L(q)[:, 1:]


def attitude_dynamics(x_attitude, J=np.diag(_J_DEFAULT)):       # :185-196
    """[q_dot, omega_dot] of a torque-free rigid body: q_dot = 1/2 q (x) (0, omega), J omega_dot = -omega x J omega.
    Like the reference it normalises the quaternion part of its ARGUMENT in place (:187)."""
    x_attitude[:4] /= np.linalg.norm(x_attitude[:4])
    return hm.attitude_rhs(x_attitude, np.asarray(J, dtype=np.float64))


def attitude_step(xk, h):                                       # :198-207
    """One RK4 step + quaternion re-normalisation on the device (`vinsat_attitude_propagate`, one trajectory, one
    step).  The reference's first stage normalises xk[:4] in place; that side effect is kept."""
    xk[:4] /= np.linalg.norm(xk[:4])
    return propagate_attitudes(np.asarray(xk, dtype=np.float64)[None], 1, float(h))[0, 1]


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
