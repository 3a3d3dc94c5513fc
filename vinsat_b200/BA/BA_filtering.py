"""Mirror of ``estimation/BA/BA_filtering.py``: ``BA(...)`` with the reference's 15-argument signature
(BA_filtering.py:4) and return tuple (:98), executed as ONE batched-BA call on the device."""
import numpy as np
import torch

from .. import _lib, config
from .BA_utils import _np, _ctx, _cum_rot_of

_cache = {}


def _fingerprint(a):
    """Cheap identity of a static input: where it lives, its size and its end values.  BA() is called 20x per window
    with the same observation / frame arrays (od_pipe.py:1036-1040); a changed array changes at least one of these."""
    a = np.asarray(a)
    flat = a.reshape(-1)
    return (a.__array_interface__["data"][0], a.shape, a.dtype.str,
            (float(flat[0]), float(flat[-1]), float(flat[len(flat) // 2])) if len(flat) else ())


def _batch_for(arrays, sources):
    """One cached device batch per (T, M).  Returns (batch, resident): resident=True when every static input is the
    one already on the device AND `states` is exactly what the previous BA() call returned -- then nothing is
    uploaded and the iteration continues from the device-resident states (and residuals)."""
    key = (config.device, len(arrays["time_idx"]), len(arrays["ii"]))
    fp = tuple(_fingerprint(x) for x in sources)
    ent = _cache.get(key)
    if ent is None:
        if len(_cache) > 8:
            for old in _cache.values():
                old["batch"].close()
            _cache.clear()
        ent = _cache[key] = dict(batch=_lib.Batch(_ctx(), arrays), fp=fp, last=None)
        return ent, False
    if ent["fp"] == fp and ent["last"] is not None and np.array_equal(ent["last"], arrays["states"]):
        return ent, True
    ent["batch"].upload(arrays)
    ent["fp"] = fp
    return ent, False


def BA(iter, states, velocities, imu_meas, landmarks, landmarks_xyz, ii, time_idx, intrinsics, confidences,
       Sigma, V, lamda_init, poses_gt_eci, initialize=False):
    """One damped Gauss-Newton / LM iteration (BA_filtering.py:4-98).

    states (1,T,10); velocities (1,T,3) passed through; imu_meas (1,T,N,10) of which only
    [..., -1, 6:10] is read; landmarks (1,M,2) pixels; landmarks_xyz (1,M,3); ii (M,) ints; time_idx (T,);
    intrinsics (1,T,4); confidences (M,).  `Sigma` and `V` are overwritten inside the reference (:26-27)
    and are ignored here too.  Returns (states_new (1,T,10), velocities, lamda_init, last_hessian (1,9,9)).
    """
    st = _np(states)
    assert st.shape[0] == 1, "the reference is batch-size-1 only (SURVEY 0.11)"
    T = st.shape[1]
    ii_np = np.asarray(ii, dtype=np.int64).reshape(-1)
    M = len(ii_np)
    uv = _np(landmarks).reshape(M, 2)
    xyz = _np(landmarks_xyz).reshape(M, 3)
    conf = _np(confidences).reshape(M)
    if M > 1 and np.any(np.diff(ii_np) < 0):            # the kernels want frame-sorted observations
        order = np.argsort(ii_np, kind="stable")
        ii_np, uv, xyz, conf = ii_np[order], uv[order], xyz[order], conf[order]
    arrays = dict(frame_off=np.array([0, T], dtype=np.int64), obs_off=np.array([0, M], dtype=np.int64),
                  states=np.ascontiguousarray(st[0]), intrinsics=np.ascontiguousarray(_np(intrinsics).reshape(T, 4)),
                  cum_rot=np.ascontiguousarray(_cum_rot_of(imu_meas)),
                  time_idx=np.ascontiguousarray(time_idx, dtype=np.int64), landmarks_xyz=np.ascontiguousarray(xyz),
                  landmarks_uv=np.ascontiguousarray(uv), confidences=np.ascontiguousarray(conf),
                  ii=np.ascontiguousarray(ii_np))
    ent, resident = _batch_for(arrays, (landmarks, landmarks_xyz, confidences, ii, time_idx, intrinsics, imu_meas))
    b = ent["batch"]
    lam, ntr = b.ba_iterate(int(iter), float(lamda_init), initialize=bool(initialize), mode=config.mode())
    if lam[0] * 100 > 1e4 and ntr[0] >= 8:
        pass   # the reference prints "lamda too large" (:76); kept silent here
    new = b.get_states()
    ent["last"] = new.copy()        # private copy: the caller may modify the tensor it gets back
    states_new = torch.from_numpy(new)[None]
    last_hessian = torch.from_numpy(b.last_hessian())
    if iter > 18:                                        # :86-87
        gt = _np(poses_gt_eci)
        d = np.abs(states_new[0, :, :3].numpy() - gt[:, :3]).mean(axis=0)
        print("final pos: ", torch.from_numpy(d), float(np.linalg.norm(d)))
    return states_new, velocities, float(lam[0]), last_hessian



def BA_reg(iter, states, velocities, states_prior, velocity_prior, hessian_state_t, hessian_rot_t, imu_meas, landmarks,
           landmarks_xyz, ii, time_idx, intrinsics, confidences, Sigma, V, lamda_init, poses_gt_eci, initialize=False,
           use_reg=True):
    """The prior-regularised iteration (BA_filtering.py:100-210): BA() plus the prior terms of `prior_gpu` around
    `states_prior` with the information matrices `hessian_state_t` (1,T,6,6) / `hessian_rot_t` (1,T,3,3).  Same return
    tuple as BA().  Only initialize=False exists on the device -- it is the only way the reference calls it
    (od_pipe.py:893); `velocity_prior` and `use_reg` are unused by the reference as well."""
    if initialize:
        raise NotImplementedError("BA_reg(initialize=True) is never used by the reference (od_pipe.py:893)")
    st = _np(states)
    assert st.shape[0] == 1, "the reference is batch-size-1 only (SURVEY 0.11)"
    T = st.shape[1]
    ii_np = np.asarray(ii, dtype=np.int64).reshape(-1)
    M = len(ii_np)
    uv = _np(landmarks).reshape(M, 2)
    xyz = _np(landmarks_xyz).reshape(M, 3)
    conf = _np(confidences).reshape(M)
    if M > 1 and np.any(np.diff(ii_np) < 0):
        order = np.argsort(ii_np, kind="stable")
        ii_np, uv, xyz, conf = ii_np[order], uv[order], xyz[order], conf[order]
    arrays = dict(frame_off=np.array([0, T], dtype=np.int64), obs_off=np.array([0, M], dtype=np.int64),
                  states=np.ascontiguousarray(st[0]), intrinsics=np.ascontiguousarray(_np(intrinsics).reshape(T, 4)),
                  cum_rot=np.ascontiguousarray(_cum_rot_of(imu_meas)),
                  time_idx=np.ascontiguousarray(time_idx, dtype=np.int64), landmarks_xyz=np.ascontiguousarray(xyz),
                  landmarks_uv=np.ascontiguousarray(uv), confidences=np.ascontiguousarray(conf),
                  ii=np.ascontiguousarray(ii_np))
    ent, resident = _batch_for(arrays, (landmarks, landmarks_xyz, confidences, ii, time_idx, intrinsics, imu_meas))
    b = ent["batch"]
    b.set_prior(_np(states_prior)[0], _np(hessian_state_t)[0], _np(hessian_rot_t)[0])
    lam, ntr = b.ba_reg_iterate(int(iter), float(lamda_init), mode=config.mode())
    new = b.get_states()
    ent["last"] = new.copy()
    states_new = torch.from_numpy(new)[None]
    last_hessian = torch.from_numpy(b.last_hessian())
    gt = _np(poses_gt_eci)
    d = np.abs(states_new[0, :, :3].numpy() - gt[:, :3]).mean(axis=0)
    print("final pos: ", torch.from_numpy(d), float(np.linalg.norm(d)))          # :194
    return states_new, velocities, float(lam[0]), last_hessian
