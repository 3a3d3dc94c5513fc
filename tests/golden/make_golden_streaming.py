"""Golden vectors for the driver-side indexing (a9) and the streaming driver (f1), from the unmodified
reference: read_detections, the visibility mask, remove_elems, identify_next_batch_new and
streaming_version on a seeded synthetic sequence (vinsat_b200.synth.make_sequence)."""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_loader  # noqa: E402
from vinsat_b200 import synth  # noqa: E402


def main(what="all"):
    ns = ref_loader.load()
    od = ns.od_pipe
    U = ns.BA_utils
    cwd = os.getcwd()
    os.chdir(ns.est_dir)          # read_detections opens landmarks/intrinsics.csv relative to cwd
    try:
        for name, seed, kw in (("seq_a", 4, {}),
                               ("seq_b", 9, dict(n_orbit=2300, windows=((995, 1100, 7), (1400, 1460, 3), (1990, 2060, 5)),
                                                 dets_per_frame=4))):
            dets, orbit = synth.make_sequence(seed, **kw)
            out = dict(dets=dets, orbit=orbit)
            # --- indexing pieces, step by step -------------------------------------------------------
            orb, ld, intr, time_idx, ii = od.read_detections(False, detections=dets.copy(), orbit_np=orbit.copy())
            out.update(rd_orbit=orb, rd_time_idx=time_idx, rd_ii=ii, rd_intr=intr)
            g = od.process_ground_truths(orb, ld, intr, 1.0, time_idx)
            (gt_pos, gt_vel, poses_gt, gt_quat, gt_quat_full, lm_xyz, lm_uv, intr_t, gt_acc, gt_pos_full) = g
            out.update(pg_poses_gt=poses_gt.numpy(), pg_gt_vel=gt_vel.numpy(), pg_lm_xyz=lm_xyz.numpy(),
                       pg_quat_full=gt_quat_full.numpy())
            states_gt = torch.cat([poses_gt, gt_vel[time_idx]], dim=-1)
            proj = U.landmark_project(states_gt.unsqueeze(0), lm_xyz.unsqueeze(0), intr_t.unsqueeze(0), ii, jacobian=False)
            mask = ((proj[:, :, 0] > 0) * (proj[:, :, 1] > 0) * (proj[:, :, 0] < 4700) * (proj[:, :, 1] < 2600)
                    * ((proj - lm_uv[None]).norm(dim=-1) < 1000) * (torch.tensor(ld["confidence"]) > 0.8))[0]
            out.update(vis_proj=proj[0].detach().numpy(), vis_mask=mask.numpy().copy())
            r = od.remove_elems(mask.clone(), gt_pos, gt_vel, poses_gt, gt_quat, gt_quat_full, lm_xyz, lm_uv, intr_t,
                                gt_acc, ii, time_idx)
            ii_new, time_idx_new, mask_new = r[9], r[10], r[11]
            out.update(re_ii=np.asarray(ii_new), re_time_idx=np.asarray(time_idx_new), re_mask=mask_new.numpy(),
                       re_poses_gt=r[2].numpy())
            # streaming split points
            splits = []
            i = 0
            t = 0
            end = False
            while not end:
                t, i, end = od.identify_next_batch_new(ii_new, time_idx_new, i, t)
                splits.append((int(t), int(i), bool(end)))
            out.update(splits=np.array(splits, dtype=np.int64))
            if what in ("streaming", "all"):
                t0 = time.time()
                errors, first_det, times = od.streaming_version(detections=dets.copy(), orbit_np=orbit.copy())
                print(name, "streaming_version took %.1f s" % (time.time() - t0))
                out.update(sv_errors=errors.detach().numpy(), sv_first_detection=np.int64(first_det),
                           sv_times=np.concatenate([np.asarray(x).reshape(-1) for x in times]).astype(np.int64),
                           sv_times_len=np.array([len(np.asarray(x).reshape(-1)) for x in times], dtype=np.int64))
            np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
            print(name, "saved: T=%d (after removal %d), M=%d (kept %d), splits=%s" % (
                len(time_idx), len(time_idx_new), len(ii), int(mask_new.sum()), splits))
    finally:
        os.chdir(cwd)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "all")
