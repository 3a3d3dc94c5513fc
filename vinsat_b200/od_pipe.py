"""Mirror of ``estimation/od_pipe.py`` (the live OD driver): ``read_detections``, ``process_ground_truths``,
``remove_elems``, ``identify_next_batch_new``, ``streaming_version`` keep the reference's names, arguments
and return values.  The integer indexing (unique frames, knots, `ii`, re-indexing after the visibility mask) runs as
device kernels, the whole window loop of ``streaming_version`` is one device call (``vinsat_stream_solve``).  Citations: path:line under <reference>/estimation.

The debug drivers of the reference (``od_pipe`` [stale], ``full_batch_optimization``, ``*_debugging``)
stop in ``ipdb.set_trace()`` and are out of scope.
"""
import os

import numpy as np
import torch

from . import _lib, config
from . import hostmath as hm
from .BA.BA_filtering import BA  # noqa: F401  (same import surface as the reference)
from .BA.BA_utils import (landmark_project, propagate_dynamics_init, quaternion_exp, quaternion_log,
                          precompute_cum_rotations, compute_omega_from_quat, cum_rotations_from_quat, _ctx)

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def _load_intrinsics():
    """od_pipe.py:248 reads landmarks/intrinsics.csv relative to the cwd; fall back to the packaged copy."""
    path = "landmarks/intrinsics.csv"
    if not os.path.exists(path):
        path = os.path.join(_DATA, "intrinsics.csv")
    return np.genfromtxt(path, delimiter=",")[0]


def seeding(seed):
    """od_pipe.py:307-309."""
    torch.manual_seed(seed)
    np.random.seed(seed)


def read_detections(sample_dets=False, detections=None, orbit_np=None, orbit_file_name=None, detections_file_name=None):
    """od_pipe.py:185-251.  Returns (orbit [ECI km, converted IN PLACE like the reference], landmarks_dict,
    intrinsics row 0, time_idx with knot frames at multiples of 1000 s, ii)."""
    landmarks = detections if detections is not None else np.load(detections_file_name, allow_pickle=True)
    landmarks_dict = {"frame": landmarks[:, 0], "uv": landmarks[:, 3:5], "lonlat": landmarks[:, 1:3],
                      "confidence": landmarks[:, 5]}
    orbit = orbit_np if orbit_np is not None else np.load(orbit_file_name, allow_pickle=True)
    # :214-228,242-247 -- unique frames, knot insertion and the obs -> frame slot map as integer kernels on the device
    time_idx_new, ii = _ctx().index_detections(landmarks[:, 0], orbit.shape[0])
    orbit[:, 0], orbit[:, 1], orbit[:, 2] = hm.ecef_to_eci(orbit[:, 0] / 1000, orbit[:, 1] / 1000, orbit[:, 2] / 1000,
                                                             times=np.arange(orbit.shape[0]))     # :240
    return orbit, landmarks_dict, _load_intrinsics(), time_idx_new, ii


def process_ground_truths(orbit, landmarks_dict, intrinsics, dt, time_idx):
    """od_pipe.py:94-123 (nadir branch)."""
    gt_pos_eci = orbit[time_idx, :3]
    gt_pos_eci_full = orbit[:, :3]
    gt_vel_eci = hm.compute_velocity_from_pos(orbit[:, :3], dt)
    gt_quat_eci = hm.convert_pos_to_quaternion(gt_pos_eci)
    gt_quat_eci_full = hm.convert_pos_to_quaternion(gt_pos_eci_full)
    poses_gt_eci = np.concatenate([gt_pos_eci, gt_quat_eci], axis=1)
    landmarks_xyz = hm.convert_latlong_to_cartesian(landmarks_dict["lonlat"][:, 1], landmarks_dict["lonlat"][:, 0],
                                                    landmarks_dict["frame"])
    gt_acceleration = hm.compute_velocity_from_pos(gt_vel_eci, dt)
    t = torch.tensor
    return (t(gt_pos_eci), t(gt_vel_eci), t(poses_gt_eci), t(gt_quat_eci), t(gt_quat_eci_full), t(landmarks_xyz),
            t(landmarks_dict["uv"]).double(), t(intrinsics).unsqueeze(0).repeat(len(gt_pos_eci), 1),
            t(gt_acceleration), t(gt_pos_eci_full))


def remove_elems(mask, gt_pos_eci, gt_vel_eci, poses_gt_eci, gt_quat_eci, gt_quat_eci_full, landmarks_xyz,
                 landmarks_uv, intrinsics, gt_acceleration, ii, time_idx):
    """od_pipe.py:253-288: frames kept = frames with a surviving observation or knots; surviving observations are
    re-indexed by an exclusive prefix sum over the dropped frames -- device kernels
    (`vinsat_remove_elems_index`: mark, scan, compact)."""
    mask_np = mask.numpy() if isinstance(mask, torch.Tensor) else np.asarray(mask)
    ii_new, time_idx_new, keep = _ctx().remove_elems_index(mask_np, ii, time_idx)
    return (gt_pos_eci[keep], gt_vel_eci, poses_gt_eci[keep], gt_quat_eci[keep], gt_quat_eci_full, landmarks_xyz,
            landmarks_uv, intrinsics, gt_acceleration, ii_new, time_idx_new, mask)


def identify_next_batch_new(ii, time_idx, i, t):
    """od_pipe.py:898-905."""
    contiguous_patch_count = 0
    for j in range(i + 1, len(ii)):
        gap = time_idx[ii[j]] - time_idx[ii[j - 1]]
        if gap < 100:
            contiguous_patch_count += 1
        if gap > 200 and contiguous_patch_count > 4:
            return ii[j - 1] + 1, j, False
    return ii[-1] + 1, len(ii), True


def compute_residuals(states, gt_states):
    return (states - gt_states)[..., :3].reshape(-1, 3).norm(dim=-1)


def visibility_mask(landmark_uv_proj, landmarks_uv, confidence):
    """od_pipe.py:930."""
    p = landmark_uv_proj
    return ((p[:, :, 0] > 0) * (p[:, :, 1] > 0) * (p[:, :, 0] < 4700) * (p[:, :, 1] < 2600)
            * ((p - landmarks_uv[None]).norm(dim=-1) < 1000) * (torch.as_tensor(confidence) > 0.8))[0]


def window_schedule(ii, time_idx):
    """The (t_final, i_final) pairs the `while not seq_end` loop of od_pipe.py:987-992 visits, by repeated
    identify_next_batch_new (:898-905) -- integers only, bit-exact by construction."""
    t = i = 0
    seq_end = False
    t_final, i_final = [], []
    while not seq_end:
        t, i, seq_end = identify_next_batch_new(ii, time_idx, i, t)
        t_final.append(int(t))
        i_final.append(int(i))
    return np.array(t_final, dtype=np.int64), np.array(i_final, dtype=np.int64)


def streaming_version(detections=None, orbit_np=None, orbit_file_name=None, detections_file_name=None):
    """od_pipe.py:911-1062.  Returns (errors, first_detection, times).

    Host part = the reference's own preprocessing (ingest, ground truth, visibility mask, remove_elems, initial guess
    with the reference's RNG call order).  The window loop (:987-1060) is ONE device call, `vinsat_stream_solve`: the
    growing-window schedule is precomputed here (integers), the device seeds every window's new frames with
    propagate_dynamics_init, runs the 20 BA iterations per window and keeps the solved states resident.  The error /
    time lists are then assembled from the states the call returns, in the order the reference appends them."""
    seeding(0)
    h = 1
    dt = 1 / h
    num_iters = 20
    orbit, landmarks_dict, intrinsics, time_idx, ii = read_detections(False, detections=detections, orbit_np=orbit_np,
                                                                      orbit_file_name=orbit_file_name,
                                                                      detections_file_name=detections_file_name)
    (gt_pos_eci, gt_vel_eci, poses_gt_eci, gt_quat_eci, gt_quat_eci_full, landmarks_xyz, landmarks_uv, intrinsics,
     gt_acceleration, gt_pos_eci_full) = process_ground_truths(orbit, landmarks_dict, intrinsics, dt, time_idx)
    states_gt_eci = torch.cat([poses_gt_eci, gt_vel_eci[time_idx]], dim=-1)
    landmark_uv_proj = landmark_project(states_gt_eci.unsqueeze(0), landmarks_xyz.unsqueeze(0), intrinsics.unsqueeze(0),
                                        ii, jacobian=False)
    mask = visibility_mask(landmark_uv_proj, landmarks_uv, landmarks_dict["confidence"])
    (gt_pos_eci, gt_vel_eci, poses_gt_eci, gt_quat_eci, gt_quat_eci_full, landmarks_xyz, landmarks_uv, intrinsics,
     gt_acceleration, ii, time_idx, mask) = remove_elems(mask, gt_pos_eci, gt_vel_eci, poses_gt_eci, gt_quat_eci,
                                                         gt_quat_eci_full, landmarks_xyz, landmarks_uv, intrinsics,
                                                         gt_acceleration, ii, time_idx)
    landmarks_xyz, landmarks_uv, landmark_uv_proj = landmarks_xyz[mask], landmarks_uv[mask], landmark_uv_proj[:, mask]
    confidences = torch.tensor(landmarks_dict["confidence"])[mask].double()
    print("mean landmark difference : ", ((landmark_uv_proj[0, :] - landmarks_uv)).abs().mean(dim=0))
    noise_level = 1.0
    landmarks_uv += (landmark_uv_proj[0, :] - landmarks_uv) * (1 - noise_level)

    # initial guess (od_pipe.py:941-973); the torch RNG call order is the reference's
    T = len(gt_pos_eci)
    velocities = gt_vel_eci[time_idx].double()
    # od_pipe.py:944-953 (compute_omega_from_quat, zero-padded omegas (1,T,N,3), precompute_cum_rotations) as one
    # device call that returns the only slice `predict` reads, cum_rotations[0, :, -1]
    cum_rot, gt_omega = cum_rotations_from_quat(gt_quat_eci_full, time_idx, dt)
    position_offset = torch.randn((T, 3)) * 100
    orientation_offset = torch.randn([T, 3]) * 0.2
    velocity_offset = torch.randn([T, 3]) * velocities.abs().mean() * 0.1
    position = poses_gt_eci.double()[:, :3] + position_offset
    orientation = quaternion_exp(quaternion_log(poses_gt_eci.double()[:, 3:]) + orientation_offset)
    states = torch.cat([position, orientation, velocities + velocity_offset], dim=-1)

    t_final, i_final = window_schedule(ii, time_idx)
    out = _ctx().stream_solve(states.numpy(), velocities.numpy(), intrinsics.double().numpy(), cum_rot, time_idx,
                              landmarks_xyz.double().numpy(), landmarks_uv.double().numpy(), confidences.numpy(), ii,
                              gt_omega.double().numpy(), t_final, i_final, num_iters=num_iters, n_init_first=10,
                              lamda_init=1e-4, mode=config.mode())
    seed = torch.from_numpy(out["seed_states"])
    win_last = torch.from_numpy(out["window_last_state"])
    gt = poses_gt_eci.double()
    errors, times = [], []
    first_detection = time_idx[t_final[0] - 1]
    for w, tf in enumerate(t_final):
        if w > 0:                                              # error of the propagated seeds, last one dropped (:1023-1026)
            ti = t_final[w - 1]
            errors.append(compute_residuals(seed[ti:tf, :3], gt[ti:tf, :3])[:-1])
            times.append(time_idx[ti:tf][:-1])
        errors.append((win_last[w:w + 1, :3] - gt[tf - 1:tf, :3]).norm(dim=-1))        # :1042-1044
        times.append(time_idx[tf - 1:tf])
    if t_final[-1] < len(time_idx):                            # :1046-1060: propagate to the end of the sequence
        ti = t_final[-1]
        errors.append(compute_residuals(seed[ti:, :3], gt[ti:, :3]))
        times.append(time_idx[ti:])
    errors = torch.cat(errors)
    return errors, first_detection, times
