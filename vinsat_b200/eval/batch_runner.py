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


NOISE_SWEEP_PX = (0.25, 0.5, 1.0, 2.0, 4.0)         # sigma_px of the noise sweep (configs[3])


class ODPool:
    """configs[3]: a Monte-Carlo noise sweep over many OD problems, sharded across ranks and workers by a dynamic
    chunk pool (vinsat_b200/pool.py; no data-path collective).

    Chunk c = `chunk` problems: the c-th noise realisation of `chunk` simulated base arcs (seeded base_seed + p), with
    pixel noise sigma_px = sigmas[c % len(sigmas)] and a fresh perturbed initial guess (od_pipe.py:962-969), both drawn
    on the device from the chunk's seed.  Every worker (own context / stream / device batch) keeps the base arcs
    resident; per chunk only the seed goes in and the per-problem errors come out."""

    def __init__(self, chunk=1024, frames=1000, obs_per_frame=10, sigmas=NOISE_SWEEP_PX, device=None, rank=0,
                 world_size=1, workers=2, base_seed=0, mode=None):
        from .. import _lib, config, synth
        self.dev = config.device if device is None else device
        self.rank, self.world, self.chunk, self.sigmas, self.base_seed = rank, world_size, chunk, tuple(sigmas), base_seed
        self.mode = config.mode() if mode is None else mode
        if world_size > 1:
            _lib.bind_host_thread_to_gpu(self.dev)
        prs = synth.make_batch(chunk, frames, obs_per_frame, seed0=base_seed, sigma_px=0.0)        # noise-free pixels
        arrays = _lib.concat_problems(prs)
        st_true = np.ascontiguousarray(np.concatenate([pr["states_gt"] for pr in prs]))
        vel_true = np.ascontiguousarray(np.concatenate([pr["vel_true"] for pr in prs]))
        self.vel_sigma = float(np.abs(st_true[:, 7:]).mean() * 0.1)
        self.slots = []
        for _ in range(workers):
            ctx = _lib.Context(self.dev)
            b = _lib.Batch(ctx, arrays)
            b.mc_set_truth(st_true, arrays["landmarks_uv"], vel_true)
            self.slots.append((ctx, b))

    def run(self, n_problems, pool_key, num_iters=20, n_init=10, store=None):
        """Drains the pool of ceil(n_problems / chunk) chunks (shared with the other ranks through `pool_key`).
        Returns this rank's list of dict(chunk, sigma_px, pos_err_km, vel_err_kms)."""
        from .. import pool
        n_chunks = (n_problems + self.chunk - 1) // self.chunk
        counter = pool.make_counter(pool_key, self.world, store)
        results = []

        def work(w, c):
            ctx, b = self.slots[w]
            sig = float(self.sigmas[c % len(self.sigmas)])
            b.mc_perturb(1000003 * (self.base_seed + 1) + c, sigma_px=sig, vel_sigma=self.vel_sigma)
            b.od_solve(num_iters, n_init, 1e-4, self.mode)
            pe, ve = b.mc_errors()
            results.append(dict(chunk=c, sigma_px=sig, pos_err_km=pe, vel_err_kms=ve))

        pool.drain(counter, n_chunks, work, len(self.slots))
        return results

    def close(self):
        for ctx, b in self.slots:
            b.close()
            ctx.close()
        self.slots = []


def run_od_pool(n_problems=65536, pool_key="vinsat_mc_pool", **kw):
    """One-shot form of `ODPool`: build, drain `n_problems`, close."""
    run_kw = {k: kw.pop(k) for k in ("num_iters", "n_init", "store") if k in kw}
    p = ODPool(**kw)
    try:
        return p.run(n_problems, pool_key, **run_kw)
    finally:
        p.close()


def summarize_noise_sweep(results):
    """Per sigma_px: problems solved, median / 95th percentile / max of the per-problem max position error [m] and
    velocity error [mm/s]."""
    out = {}
    for sig in sorted({r["sigma_px"] for r in results}):
        pe = np.concatenate([r["pos_err_km"] for r in results if r["sigma_px"] == sig]) * 1e3
        ve = np.concatenate([r["vel_err_kms"] for r in results if r["sigma_px"] == sig]) * 1e6
        out[sig] = dict(n=int(len(pe)), pos_m_median=float(np.median(pe)), pos_m_p95=float(np.percentile(pe, 95)),
                        pos_m_max=float(pe.max()), vel_mms_median=float(np.median(ve)), vel_mms_p95=float(np.percentile(ve, 95)))
    return out
