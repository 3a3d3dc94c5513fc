"""ctypes binding of ``libvinsat_b200.so`` (C ABI: ``include/vinsat_b200.h``).

There is NO CPU fallback: if the shared library is missing, or no CUDA device is usable, every
operator raises ``VinsatError``.  PyTorch is not needed here; tensors are handed over as raw
pointers by the mirror modules (``BA/``, ``od_pipe.py`` ...).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VINSAT_LIB") or os.path.join(_HERE, "libvinsat_b200.so")     # VINSAT_LIB: A/B builds of the kernels

MEM_HOST, MEM_DEVICE = 0, 1
MODE_STEP1S, MODE_SKIP100 = 0, 1

c_dp = C.POINTER(C.c_double)
c_i64p = C.POINTER(C.c_int64)
c_i32p = C.POINTER(C.c_int32)
c_u8p = C.POINTER(C.c_uint8)


class VinsatError(RuntimeError):
    pass


class ProblemDesc(C.Structure):
    _fields_ = [("n_problems", C.c_int64), ("frame_off", c_i64p), ("obs_off", c_i64p), ("states", c_dp),
                ("intrinsics", c_dp), ("cum_rot", c_dp), ("time_idx", c_i64p), ("landmarks_xyz", c_dp),
                ("landmarks_uv", c_dp), ("confidences", c_dp), ("ii", c_i64p)]


class StreamDesc(C.Structure):
    _fields_ = [("n_frames", C.c_int64), ("n_obs", C.c_int64), ("states", c_dp), ("velocities", c_dp),
                ("intrinsics", c_dp), ("cum_rot", c_dp), ("time_idx", c_i64p), ("landmarks_xyz", c_dp),
                ("landmarks_uv", c_dp), ("confidences", c_dp), ("ii", c_i64p), ("n_omega", C.c_int64), ("omega", c_dp),
                ("n_windows", C.c_int64), ("t_final", c_i64p), ("i_final", c_i64p)]


# name -> (restype, argtypes); every function declared in include/vinsat_b200.h
SIGNATURES = {
    "vinsat_abi_version": (C.c_int, []),
    "vinsat_device_count": (C.c_int, []),
    "vinsat_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "vinsat_ctx_destroy": (C.c_int, [C.c_void_p]),
    "vinsat_ctx_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vinsat_ctx_reset_stream": (C.c_int, [C.c_void_p]),
    "vinsat_ctx_synchronize": (C.c_int, [C.c_void_p]),
    "vinsat_last_error": (C.c_char_p, [C.c_void_p]),
    "vinsat_landmark_project": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vinsat_predict": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                 C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p]),
    "vinsat_propagate_chain": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p]),
    "vinsat_orbit_propagate": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_double,
                                         C.c_void_p, C.c_void_p]),
    "vinsat_cum_rotations": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_double, C.c_int64, C.c_void_p,
                                       C.c_void_p, C.c_void_p]),
    "vinsat_precompute_cum_rotations": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_double,
                                                  C.c_void_p]),
    "vinsat_attitude_propagate": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_double,
                                            C.c_void_p, C.c_void_p, C.c_void_p]),
    "vinsat_batch_create": (C.c_int, [C.c_void_p, C.POINTER(ProblemDesc), C.POINTER(C.c_void_p)]),
    "vinsat_batch_create_window": (C.c_int, [C.c_void_p, C.POINTER(ProblemDesc), C.c_int64, C.c_int64, C.c_int64,
                                             C.POINTER(C.c_void_p)]),
    "vinsat_la_num_segments": (C.c_int, [C.c_void_p]),
    "vinsat_la_alloc_reduced": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64]),
    "vinsat_la_stage": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_double]),
    "vinsat_la_ptr": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "vinsat_batch_upload": (C.c_int, [C.c_void_p, C.POINTER(ProblemDesc)]),
    "vinsat_batch_set_states": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "vinsat_batch_get_states": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "vinsat_batch_destroy": (C.c_int, [C.c_void_p]),
    "vinsat_batch_ba_iterate": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "vinsat_batch_od_solve": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int]),
    "vinsat_batch_last_hessian": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vinsat_stream_solve": (C.c_int, [C.c_void_p, C.POINTER(StreamDesc), C.c_int, C.c_int, C.c_double, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vinsat_prior": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vinsat_propagate_chain_cov": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_double, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vinsat_batch_set_prior": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vinsat_batch_ba_reg_iterate": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "vinsat_batch_mc_set_truth": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vinsat_batch_mc_perturb": (C.c_int, [C.c_void_p, C.c_uint64, C.c_double, C.c_double, C.c_double, C.c_double]),
    "vinsat_batch_mc_errors": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "vinsat_batch_debug_fetch": (C.c_int, [C.c_void_p] + [C.c_void_p] * 7),
    "vinsat_batch_eval_resjac": (C.c_int, [C.c_void_p]),
    "vinsat_batch_fetch_resjac": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "vinsat_ctx_enable_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "vinsat_ctx_reset_timing": (C.c_int, [C.c_void_p]),
    "vinsat_ctx_get_timing": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), c_dp, c_i64p]),
    "vinsat_ctx_launch_count": (C.c_int64, [C.c_void_p]),
    "vinsat_satcam_project": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                        C.c_double, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vinsat_satcam_corners": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_double, C.c_int32,
                                        C.c_int32, C.c_void_p, C.c_void_p]),
    "vinsat_satcam_corner_rays": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_double, C.c_int32,
                                            C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vinsat_satcam_cam_matrix": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_double, C.c_int32,
                                           C.c_int32, C.c_void_p]),
    "vinsat_satcam_table_create": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                             C.c_void_p, C.POINTER(C.c_void_p)]),
    "vinsat_satcam_table_destroy": (C.c_int, [C.c_void_p]),
    "vinsat_satcam_visibility": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_double,
                                           C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vinsat_index_detections": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                          C.c_void_p, C.c_void_p]),
    "vinsat_remove_elems_index": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vinsat_measure_fp64_peak": (C.c_int, [C.c_void_p, c_dp]),
    "vinsat_measure_copy_bw": (C.c_int, [C.c_void_p, C.c_int64, c_dp]),
}

_lib = None


def load():
    """Loads the shared library (once).  Raises VinsatError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VinsatError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(make -C vinsat_b200/csrc).  There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)     # AttributeError => ABI mismatch, fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.vinsat_abi_version() != 1:
        raise VinsatError("ABI version mismatch")
    _lib = lib
    return lib


def _ptr(a):
    """Raw pointer of a NumPy array / torch tensor / int / None."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


class Context:
    """Owns a ``vinsat_ctx`` (device, stream, scratch)."""

    def __init__(self, device=0):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.vinsat_ctx_create(int(device), C.byref(h))
        if rc != 0:
            raise VinsatError("vinsat_ctx_create(%d) failed (%d): %s" % (
                device, rc, (self.lib.vinsat_last_error(None) or b"").decode()))
        self.h = h
        self.device = device

    def check(self, rc):
        if rc != 0:
            raise VinsatError("vinsat error %d: %s" % (rc, (self.lib.vinsat_last_error(self.h) or b"").decode()))

    def close(self):
        if getattr(self, "h", None):
            for t in self.__dict__.pop("_satcam_tables", {}).values():      # device tables created through this context
                t.close()
            self.lib.vinsat_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        self.check(self.lib.vinsat_ctx_synchronize(self.h))

    def set_stream(self, cuda_stream):
        """cuda_stream: integer handle (torch.cuda.Stream.cuda_stream); 0 = the legacy default stream."""
        self.check(self.lib.vinsat_ctx_set_stream(self.h, C.c_void_p(int(cuda_stream))))

    def reset_stream(self):
        self.check(self.lib.vinsat_ctx_reset_stream(self.h))

    def launch_count(self):
        return int(self.lib.vinsat_ctx_launch_count(self.h))

    def enable_timing(self, on=True):
        self.check(self.lib.vinsat_ctx_enable_timing(self.h, 1 if on else 0))

    def reset_timing(self):
        self.check(self.lib.vinsat_ctx_reset_timing(self.h))

    def timing(self):
        names = (C.c_char_p * 32)()
        ms = (C.c_double * 32)()
        n_l = (C.c_int64 * 32)()
        n = self.lib.vinsat_ctx_get_timing(self.h, 32, names, ms, n_l)
        return {names[i].decode(): (ms[i], n_l[i]) for i in range(n)}

    def fp64_peak_tflops(self):
        v = C.c_double()
        self.check(self.lib.vinsat_measure_fp64_peak(self.h, C.byref(v)))
        return v.value

    def copy_bw_gbs(self, nbytes=1 << 30):
        v = C.c_double()
        self.check(self.lib.vinsat_measure_copy_bw(self.h, int(nbytes), C.byref(v)))
        return v.value

    # ---- stand-alone operators (host NumPy in / out) --------------------------------------------
    def landmark_project(self, states, landmarks_xyz, intrinsics, ii, jacobian=True):
        states, xyz, intr, ii = f64(states), f64(landmarks_xyz), f64(intrinsics), i64(ii)
        T, M = states.shape[0], xyz.shape[0]
        uv = np.empty((M, 2))
        Jg = np.empty((M, 2, 9)) if jacobian else None
        self.check(self.lib.vinsat_landmark_project(self.h, MEM_HOST, T, M, _ptr(states), _ptr(intr), _ptr(xyz),
                                                    _ptr(ii), _ptr(uv), _ptr(Jg)))
        return (uv, Jg) if jacobian else uv

    def predict(self, states, cum_rot, time_idx, quat_coeff=100.0, vel_coeff=100.0, jacobian=True,
                mode=MODE_STEP1S, want_x_pred=False):
        states, cum_rot, time_idx = f64(states), f64(cum_rot), i64(time_idx)
        T = states.shape[0]
        out = {"r_pred": np.empty((T - 1, 7))}
        if want_x_pred:
            out["x_pred"] = np.empty((T, 6))
        if jacobian:
            out.update(Phi=np.empty((T - 1, 6, 6)), qgrad=np.empty((T, 3)), Hq_diag=np.empty((T, 3, 3)),
                       Hq_off=np.empty((T - 1, 3, 3)))
        self.check(self.lib.vinsat_predict(self.h, MEM_HOST, T, _ptr(states), _ptr(cum_rot), _ptr(time_idx),
                                           float(quat_coeff), float(vel_coeff), int(mode), _ptr(out["r_pred"]),
                                           _ptr(out.get("x_pred")), _ptr(out.get("Phi")), _ptr(out.get("qgrad")),
                                           _ptr(out.get("Hq_diag")), _ptr(out.get("Hq_off"))))
        return out

    def propagate_chain(self, state0, vel0, omega, dt=1.0):
        state0, vel0, omega = f64(state0), f64(vel0), f64(omega).reshape(-1, 3)
        n = omega.shape[0]
        out = np.empty((n + 1, 10))
        self.check(self.lib.vinsat_propagate_chain(self.h, MEM_HOST, n, float(dt), _ptr(state0), _ptr(vel0),
                                                   _ptr(omega), _ptr(out)))
        return out

    def orbit_propagate(self, x0, n_steps, stride=1, h=1.0):
        x0 = f64(x0).reshape(-1, 6)
        out = np.empty((x0.shape[0], n_steps // stride + 1, 6))
        self.check(self.lib.vinsat_orbit_propagate(self.h, MEM_HOST, x0.shape[0], int(n_steps), int(stride),
                                                   float(h), _ptr(x0), _ptr(out)))
        return out

    def cum_rotations(self, quat_full, time_idx, dt=1.0, want_omega=False):
        """compute_omega_from_quat + precompute_cum_rotations(...)[0, :, -1] (od_pipe.py:944-953) on the device."""
        q, ti = f64(quat_full).reshape(-1, 4), i64(time_idx)
        n, T = q.shape[0], ti.shape[0]
        om = np.empty((n, 3)) if want_omega else None
        cr = np.empty((T, 4))
        self.check(self.lib.vinsat_cum_rotations(self.h, MEM_HOST, n, _ptr(q), float(dt), T, _ptr(ti), _ptr(om),
                                                 _ptr(cr)))
        return (cr, om) if want_omega else cr

    def precompute_cum_rotations(self, omegas, dt=1.0):
        """BA_utils.py:278-288: omegas (T, N, 3) -> cumulative rotations (T, N, 4)."""
        om = f64(omegas)
        T, N = om.shape[0], om.shape[1]
        out = np.empty((T, N, 4))
        self.check(self.lib.vinsat_precompute_cum_rotations(self.h, MEM_HOST, T, N, _ptr(om), float(dt), _ptr(out)))
        return out

    def attitude_propagate(self, x0, n_steps, stride=1, h=1.0, inertia_diag=None):
        """Batched trajgen_pipe.attitude_step: x0 (n,7) [q scalar-first, omega] -> (n, n_steps/stride+1, 7)."""
        x0 = f64(x0).reshape(-1, 7)
        J = f64(inertia_diag if inertia_diag is not None else
                (1 / 3) * np.array([(.1 ** 2 + .34 ** 2), (.1 ** 2 + .34 ** 2), (.1 ** 2 + .1 ** 2)]))
        out = np.empty((x0.shape[0], n_steps // stride + 1, 7))
        self.check(self.lib.vinsat_attitude_propagate(self.h, MEM_HOST, x0.shape[0], int(n_steps), int(stride),
                                                      float(h), _ptr(J), _ptr(x0), _ptr(out)))
        return out

    def index_detections(self, frames, n_orbit):
        """read_detections' indexing (od_pipe.py:214-247) on the device -> (time_idx with knots, ii)."""
        fr = f64(frames).reshape(-1)
        n = fr.shape[0]
        cap = n + int(n_orbit) // 1000 + 2
        tix, ii = np.empty(cap, dtype=np.int64), np.empty(n, dtype=np.int64)
        nf = C.c_int64()
        self.check(self.lib.vinsat_index_detections(self.h, MEM_HOST, n, _ptr(fr), int(n_orbit), cap, _ptr(tix), _ptr(ii),
                                                    C.byref(nf)))
        return tix[:nf.value].copy(), ii

    def remove_elems_index(self, mask, ii, time_idx):
        """remove_elems' re-indexing (od_pipe.py:253-288) on the device -> (ii_new, time_idx_new, frame keep mask)."""
        mk = np.ascontiguousarray(mask, dtype=np.uint8).reshape(-1)
        ii, tix = i64(ii), i64(time_idx)
        n, T = mk.shape[0], tix.shape[0]
        ii_out, t_out = np.empty(max(n, 1), dtype=np.int64), np.empty(T, dtype=np.int64)
        keep = np.empty(T, dtype=np.uint8)
        cnt = np.zeros(2, dtype=np.int64)
        self.check(self.lib.vinsat_remove_elems_index(self.h, MEM_HOST, n, T, _ptr(mk), _ptr(ii), _ptr(tix), _ptr(ii_out),
                                                      _ptr(t_out), _ptr(keep), _ptr(cnt)))
        return ii_out[:cnt[0]].copy(), t_out[:cnt[1]].copy(), keep.astype(bool)

    def prior(self, states, prop_states, vel_coeff, quat_coeff, hessian_state, hessian_rot, jacobian=True):
        """prior_gpu (BA_utils.py:604-676) in block-diagonal form: r (N,7) [, Jp (N,6,9), Hqp (N,9,9), qgrad (N,9)]."""
        st, pr = f64(states).reshape(-1, 10), f64(prop_states).reshape(-1, 10)
        N = st.shape[0]
        hs, hr = f64(hessian_state).reshape(N, 6, 6), f64(hessian_rot).reshape(N, 3, 3)
        r = np.empty((N, 7))
        Jp = np.empty((N, 6, 9)) if jacobian else None
        Hqp = np.empty((N, 9, 9)) if jacobian else None
        qg = np.empty((N, 9)) if jacobian else None
        self.check(self.lib.vinsat_prior(self.h, MEM_HOST, N, _ptr(st), _ptr(pr), float(vel_coeff), float(quat_coeff),
                                         _ptr(hs), _ptr(hr), _ptr(r), _ptr(Jp), _ptr(Hqp), _ptr(qg)))
        return (r, Jp, Hqp, qg) if jacobian else r

    def propagate_chain_cov(self, state0, vel0, hessian, omega, tdiff, duration, dt=1.0):
        """propagate_dynamics_cov_init (BA_utils.py:222-248): -> states_t (duration+1,10), Hs_t (duration+1,6,6),
        Hr_t (duration+1,3,3)."""
        s0, v0, hh = f64(state0).reshape(10), f64(vel0).reshape(3), f64(hessian).reshape(9, 9)
        om = f64(omega).reshape(-1, 3)
        st = np.empty((duration + 1, 10)); hs = np.empty((duration + 1, 6, 6)); hr = np.empty((duration + 1, 3, 3))
        self.check(self.lib.vinsat_propagate_chain_cov(self.h, MEM_HOST, int(tdiff), int(duration), float(dt), _ptr(s0),
                                                       _ptr(v0), _ptr(hh), _ptr(om), _ptr(st), _ptr(hs), _ptr(hr)))
        return st, hs, hr

    def stream_solve(self, states, velocities, intrinsics, cum_rot, time_idx, landmarks_xyz, landmarks_uv, confidences,
                     ii, omega, t_final, i_final, num_iters=20, n_init_first=10, lamda_init=1e-4, mode=MODE_STEP1S):
        """streaming_version's window loop (od_pipe.py:987-1060) in one device call; see include/vinsat_b200.h.
        Returns dict(states (t_final[-1],10), seed_states (T_all,10), window_last_state (W,10), last_hessian (9,9))."""
        keep = dict(states=f64(states), velocities=f64(velocities), intrinsics=f64(intrinsics), cum_rot=f64(cum_rot),
                    time_idx=i64(time_idx), xyz=f64(landmarks_xyz), uv=f64(landmarks_uv), conf=f64(confidences), ii=i64(ii),
                    omega=f64(omega).reshape(-1, 3), t_final=i64(t_final), i_final=i64(i_final))
        d = StreamDesc()
        d.n_frames, d.n_obs = keep["states"].shape[0], keep["xyz"].reshape(-1, 3).shape[0]
        cast = lambda x, t: C.cast(_ptr(x), t)
        d.states = cast(keep["states"], c_dp); d.velocities = cast(keep["velocities"], c_dp)
        d.intrinsics = cast(keep["intrinsics"], c_dp); d.cum_rot = cast(keep["cum_rot"], c_dp)
        d.time_idx = cast(keep["time_idx"], c_i64p); d.landmarks_xyz = cast(keep["xyz"], c_dp)
        d.landmarks_uv = cast(keep["uv"], c_dp); d.confidences = cast(keep["conf"], c_dp); d.ii = cast(keep["ii"], c_i64p)
        d.n_omega = keep["omega"].shape[0]; d.omega = cast(keep["omega"], c_dp)
        d.n_windows = len(keep["t_final"]); d.t_final = cast(keep["t_final"], c_i64p); d.i_final = cast(keep["i_final"], c_i64p)
        W, T_all, T_last = d.n_windows, d.n_frames, int(keep["t_final"][-1])
        out = dict(states=np.empty((T_last, 10)), seed_states=np.empty((T_all, 10)), window_last_state=np.empty((W, 10)),
                   last_hessian=np.zeros((9, 9)))
        self.check(self.lib.vinsat_stream_solve(self.h, C.byref(d), int(num_iters), int(n_init_first), float(lamda_init),
                                                int(mode), _ptr(out["states"]), _ptr(out["seed_states"]),
                                                _ptr(out["window_last_state"]), _ptr(out["last_hessian"])))
        return out

    def satcam_project(self, poses, landmarks_ecef, hfov, w_px, h_px, want_uv=True, want_mask=True,
                       want_count=True):
        poses, lm = f64(poses).reshape(-1, 12), f64(landmarks_ecef).reshape(-1, 3)
        P, L = poses.shape[0], lm.shape[0]
        uv = np.empty((P, L, 2)) if want_uv else None
        mask = np.empty((P, L), dtype=np.uint8) if want_mask else None
        cnt = np.empty(P, dtype=np.int32) if want_count else None
        self.check(self.lib.vinsat_satcam_project(self.h, MEM_HOST, P, L, _ptr(poses), _ptr(lm), float(hfov),
                                                  int(w_px), int(h_px), _ptr(uv), _ptr(mask), _ptr(cnt)))
        return uv, mask, cnt

    def satcam_corners(self, poses, hfov, w_px, h_px, want_rays=False):
        poses = f64(poses).reshape(-1, 12)
        P = poses.shape[0]
        corners = np.empty((P, 4, 3))
        hit = np.empty((P, 4), dtype=np.uint8)
        vec = np.empty((P, 4, 3)) if want_rays else None
        self.check(self.lib.vinsat_satcam_corner_rays(self.h, MEM_HOST, P, _ptr(poses), float(hfov), int(w_px),
                                                      int(h_px), _ptr(corners), _ptr(hit), _ptr(vec)))
        return (corners, hit, vec) if want_rays else (corners, hit)

    def satcam_cam_matrix(self, poses, hfov, w_px, h_px):
        poses = f64(poses).reshape(-1, 12)
        Cm = np.empty((poses.shape[0], 3, 4))
        self.check(self.lib.vinsat_satcam_cam_matrix(self.h, MEM_HOST, poses.shape[0], _ptr(poses), float(hfov),
                                                     int(w_px), int(h_px), _ptr(Cm)))
        return Cm

    def satcam_table(self, region_codes, region_off, centroid_lonlat, active_codes):
        """Device-resident landmark table for `satcam_visibility` (see include/vinsat_b200.h)."""
        return SatcamTable(self, region_codes, region_off, centroid_lonlat, active_codes)

    def satcam_visibility(self, table, poses, hfov, w_px, h_px, want_count=False, want_corners=False):
        """check_for_all_landmarks for every pose -> visible (P,) bool [, count (P,), corner lon/lat (P,4,2),
        corner region codes (P,4)]."""
        poses = f64(poses).reshape(-1, 12)
        P = poses.shape[0]
        vis = np.zeros(P, dtype=np.uint8)
        cnt = np.zeros(P, dtype=np.int32) if want_count else None
        ll = np.empty((P, 4, 2)) if want_corners else None
        reg = np.empty((P, 4), dtype=np.int32) if want_corners else None
        self.check(self.lib.vinsat_satcam_visibility(self.h, table.h, MEM_HOST, P, _ptr(poses), float(hfov),
                                                     int(w_px), int(h_px), _ptr(vis), _ptr(cnt), _ptr(ll), _ptr(reg)))
        out = [vis.astype(bool)]
        if want_count:
            out.append(cnt)
        if want_corners:
            out += [ll, reg]
        return out[0] if len(out) == 1 else tuple(out)


class SatcamTable:
    def __init__(self, ctx, region_codes, region_off, centroid_lonlat, active_codes):
        self.ctx = ctx
        codes = np.ascontiguousarray(region_codes, dtype=np.int32)
        off = i64(region_off)
        ll = f64(centroid_lonlat).reshape(-1, 2)
        act = np.ascontiguousarray(active_codes, dtype=np.int32)
        h = C.c_void_p()
        ctx.check(ctx.lib.vinsat_satcam_table_create(ctx.h, len(codes), _ptr(codes), _ptr(off), _ptr(ll), len(act),
                                                     _ptr(act), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.lib.vinsat_satcam_table_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def bind_host_thread_to_gpu(device):
    """Pin the calling process to the CPU cores NVML reports as local to `device` (NUMA node of the GPU's PCIe
    root).  One process per GPU drives ~200 launches and 20 host round trips per OD solve; on a two-socket host a
    rank scheduled on the far socket ran 15-20 % slower (bench, 8 x B200).  Returns the core list or None."""
    try:
        import os
        import pynvml
        pynvml.nvmlInit()
        # NVML enumerates all GPUs; map the CUDA ordinal through CUDA_VISIBLE_DEVICES when it is a plain list
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = device
        if vis:
            parts = [v.strip() for v in vis.split(",") if v.strip()]
            if device < len(parts) and parts[device].isdigit():
                idx = int(parts[device])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cores = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1 and 64 * w + b < n_cpu]
        if cores:
            os.sched_setaffinity(0, cores)
            return cores
    except Exception:
        pass
    return None


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


def concat_problems(problems):
    """list of problem dicts (keys as vinsat_b200.synth) -> concatenated host arrays + offsets."""
    frame_off = np.zeros(len(problems) + 1, dtype=np.int64)
    obs_off = np.zeros(len(problems) + 1, dtype=np.int64)
    for p, pr in enumerate(problems):
        frame_off[p + 1] = frame_off[p] + pr["states0"].shape[0]
        obs_off[p + 1] = obs_off[p] + pr["xyz"].shape[0]
    cat = lambda k, dt: np.ascontiguousarray(np.concatenate([np.asarray(pr[k], dtype=dt) for pr in problems]))
    return dict(frame_off=frame_off, obs_off=obs_off, states=cat("states0", np.float64),
                intrinsics=cat("intr", np.float64), cum_rot=cat("cum_rot", np.float64),
                time_idx=cat("time_idx", np.int64), landmarks_xyz=cat("xyz", np.float64),
                landmarks_uv=cat("uv", np.float64), confidences=cat("conf", np.float64), ii=cat("ii", np.int64))


class Batch:
    """Device-resident batch of independent OD problems (``vinsat_batch``)."""

    def __init__(self, ctx, arrays, window=None):
        """arrays: dict as returned by concat_problems (NumPy or pinned torch tensors; kept alive here).
        window=(own_lo, own_hi, n_segments): frame-window of a sharded long arc (vinsat_b200/longarc.py)."""
        self.ctx = ctx
        self.lib = ctx.lib
        self.P = len(arrays["frame_off"]) - 1
        self.frame_off = np.asarray(arrays["frame_off"]).copy()
        self.obs_off = np.asarray(arrays["obs_off"]).copy()
        self.T = int(self.frame_off[-1])
        self.M = int(self.obs_off[-1])
        self._desc = self._make_desc(arrays)
        h = C.c_void_p()
        if window is None:
            ctx.check(self.lib.vinsat_batch_create(ctx.h, C.byref(self._desc), C.byref(h)))
        else:
            ctx.check(self.lib.vinsat_batch_create_window(ctx.h, C.byref(self._desc), int(window[0]), int(window[1]),
                                                          int(window[2]), C.byref(h)))
        self.h = h

    def _make_desc(self, a):
        self._keep = a
        d = ProblemDesc()
        d.n_problems = self.P
        cast = lambda x, t: C.cast(_ptr(x), t)
        d.frame_off = cast(a["frame_off"], c_i64p); d.obs_off = cast(a["obs_off"], c_i64p)
        d.states = cast(a["states"], c_dp); d.intrinsics = cast(a["intrinsics"], c_dp)
        d.cum_rot = cast(a["cum_rot"], c_dp); d.time_idx = cast(a["time_idx"], c_i64p)
        d.landmarks_xyz = cast(a["landmarks_xyz"], c_dp); d.landmarks_uv = cast(a["landmarks_uv"], c_dp)
        d.confidences = cast(a["confidences"], c_dp); d.ii = cast(a["ii"], c_i64p)
        return d

    def upload(self, arrays):
        self._desc = self._make_desc(arrays)
        self.ctx.check(self.lib.vinsat_batch_upload(self.h, C.byref(self._desc)))

    def close(self):
        if getattr(self, "h", None):
            self.lib.vinsat_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_states(self, states):
        states = f64(states)
        assert states.shape == (self.T, 10)
        self.ctx.check(self.lib.vinsat_batch_set_states(self.h, MEM_HOST, _ptr(states)))

    def get_states(self, out=None):
        out = np.empty((self.T, 10)) if out is None else out
        self.ctx.check(self.lib.vinsat_batch_get_states(self.h, MEM_HOST, _ptr(out)))
        return out

    def ba_iterate(self, it, lamda, initialize=False, mode=MODE_STEP1S):
        lam = np.ascontiguousarray(np.broadcast_to(np.asarray(lamda, dtype=np.float64), (self.P,))).copy()
        ntr = np.zeros(self.P, dtype=np.int32)
        self.ctx.check(self.lib.vinsat_batch_ba_iterate(self.h, int(it), 1 if initialize else 0, int(mode),
                                                        _ptr(lam), _ptr(ntr)))
        return lam, ntr

    def od_solve(self, num_iters=20, n_init=10, lamda_init=1e-4, mode=MODE_STEP1S):
        self.ctx.check(self.lib.vinsat_batch_od_solve(self.h, int(num_iters), int(n_init), float(lamda_init),
                                                      int(mode)))

    def set_prior(self, states_prior, hessian_state, hessian_rot):
        sp = f64(states_prior).reshape(self.T, 10)
        hs, hr = f64(hessian_state).reshape(self.T, 6, 6), f64(hessian_rot).reshape(self.T, 3, 3)
        self.ctx.check(self.lib.vinsat_batch_set_prior(self.h, MEM_HOST, _ptr(sp), _ptr(hs), _ptr(hr)))

    def ba_reg_iterate(self, it, lamda, mode=MODE_STEP1S):
        lam = np.ascontiguousarray(np.broadcast_to(np.asarray(lamda, dtype=np.float64), (self.P,))).copy()
        ntr = np.zeros(self.P, dtype=np.int32)
        self.ctx.check(self.lib.vinsat_batch_ba_reg_iterate(self.h, int(it), int(mode), _ptr(lam), _ptr(ntr)))
        return lam, ntr

    def mc_set_truth(self, states_true, uv_true, vel_true=None):
        st, uv = f64(states_true), f64(uv_true)
        vt = f64(vel_true) if vel_true is not None else None
        self.ctx.check(self.lib.vinsat_batch_mc_set_truth(self.h, _ptr(st), _ptr(uv), _ptr(vt)))

    def mc_perturb(self, seed, sigma_px=1.0, pos_sigma=100.0, rot_sigma=0.2, vel_sigma=0.44):
        self.ctx.check(self.lib.vinsat_batch_mc_perturb(self.h, int(seed), float(sigma_px), float(pos_sigma),
                                                        float(rot_sigma), float(vel_sigma)))

    def mc_errors(self):
        pe, ve = np.empty(self.P), np.empty(self.P)
        self.ctx.check(self.lib.vinsat_batch_mc_errors(self.h, _ptr(pe), _ptr(ve)))
        return pe, ve

    def last_hessian(self):
        out = np.empty((self.P, 9, 9))
        self.ctx.check(self.lib.vinsat_batch_last_hessian(self.h, _ptr(out)))
        return out

    def debug_fetch(self):
        T, M, P = self.T, self.M, self.P
        out = dict(r_obs=np.empty((M, 2)), weights=np.empty(M), c_obs=np.empty(P), D=np.empty((T, 9, 9)),
                   U=np.empty((T, 9, 9)), rhs=np.empty((T, 9)), dpose=np.empty((T, 9)))
        self.ctx.check(self.lib.vinsat_batch_debug_fetch(self.h, *[_ptr(out[k]) for k in
                                                                   ("r_obs", "weights", "c_obs", "D", "U", "rhs", "dpose")]))
        return out

    def eval_resjac(self):
        self.ctx.check(self.lib.vinsat_batch_eval_resjac(self.h))

    def fetch_resjac(self):
        r = np.empty((self.M, 2))
        J = np.empty((self.M, 2, 6))
        self.ctx.check(self.lib.vinsat_batch_fetch_resjac(self.h, _ptr(r), _ptr(J)))
        return r, J
