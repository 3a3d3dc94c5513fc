"""``eval/batch_runner.py`` of the reference is an 11-line script that loops over 16 MGRS regions and runs
``eval_landmarks.py`` (YOLO detector evaluation) twice per region in subprocesses; it has no callable
API and no estimation code.  `main()` issues exactly the reference's two command strings per region
(eval/batch_runner.py:7-11).  The Monte-Carlo OD runner BASELINE.json's configs name is ADDED here as
`run_od_monte_carlo` / `run_od_pool` without changing what `python batch_runner.py` does.
"""
import subprocess

import numpy as np

regions = ['10S', '10T', '11R', '12R', '16T', '17R', '17T', '18S',                      # eval/batch_runner.py:3-5
           '32S', '32T', '33S', '33T', '52S', '53S', '54S', '54T']
REGIONS = regions


def region_commands(region):
    """The two shell commands of eval/batch_runner.py:8 and :10 for one region (detector error evaluation, then the
    best-class / best-confidence search)."""
    return ("python eval_landmarks.py --model_path ../sim/models/" + region + ".pt --im_path datasets/" + region
            + "/images --lab_path datasets/" + region + "/labels --output_path " + region
            + "_err.npy --calculate_err --save_err",
            "python eval_landmarks.py --err_path " + region + "_err.npy --best_classes --save_best_conf"
            " --best_classes_path ../sim/best_classes/" + region + "_best_classes.npy --best_conf_path ../sim/best_confs/"
            + region + "_best_conf.npy --px_threshold 10")


def main(call=subprocess.call):
    """eval/batch_runner.py:7-11."""
    for region in regions:
        for command in region_commands(region):
            call(command, shell=True)


def run_od_monte_carlo(n_problems, frames=1000, obs_per_frame=10, seed0=0, sigma_px=1.0, device=None, rank=0,
                       world_size=1, num_iters=20, n_init=10, mode=None):
    """Independent Monte-Carlo OD problems (problem p seeded with seed0+p), sharded by contiguous blocks across
    `world_size` ranks with no communication.  Returns dict(problem_ids, pos_err_km, vel_err_kms, states)."""
    from .. import _lib, config, synth
    lo = (n_problems * rank) // world_size
    hi = (n_problems * (rank + 1)) // world_size
    if world_size > 1:
        _lib.bind_host_thread_to_gpu(config.device if device is None else device)
    ctx = _lib.default_context(config.device if device is None else device)
    prs = synth.make_batch(hi - lo, frames, obs_per_frame, seed0=seed0 + lo, sigma_px=sigma_px)
    arrays = _lib.concat_problems(prs)
    b = _lib.Batch(ctx, arrays)
    b.od_solve(num_iters, n_init, 1e-4, config.mode() if mode is None else mode)
    st = b.get_states()
    b.close()
    fo = arrays["frame_off"]
    pos = np.array([np.abs(st[fo[p]:fo[p + 1], :3] - prs[p]["states_gt"][:, :3]).max() for p in range(hi - lo)])
    vel = np.array([np.abs(st[fo[p]:fo[p + 1], 7:] - prs[p]["vel_true"]).max() for p in range(hi - lo)])
    return dict(problem_ids=np.arange(lo, hi), pos_err_km=pos, vel_err_kms=vel, states=st, frame_off=fo)


if __name__ == "__main__":
    main()
