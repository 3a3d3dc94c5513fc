"""GPU tests of the reference-interface mirror (same names / signatures as the reference's modules) against
goldens from the unmodified reference."""
import numpy as np
import pytest
import torch

from conftest import load_golden, problem_from_golden
from vinsat_b200 import od_pipe, trajgen_pipe
from vinsat_b200.BA import BA_utils as U
from vinsat_b200.BA.BA_filtering import BA

pytestmark = pytest.mark.gpu
PV = np.array([0, 1, 2, 6, 7, 8])


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def _imu(pr):
    T = len(pr["time_idx"]); N = int(np.diff(pr["time_idx"]).max())
    imu = torch.zeros(1, T, N, 10, dtype=torch.float64)
    imu[0, :, -1, 6:10] = torch.tensor(pr["cum_rot"])
    return imu


@pytest.mark.parametrize("name", ["ba_T30", "ba_T40_noisy"])
def test_BA_signature_and_iterates(name):
    g = load_golden(name); pr = problem_from_golden(g)
    t = torch.tensor
    states, vel, imu, lam = t(pr["states0"])[None], t(pr["velocities"])[None], _imu(pr), 1e-4
    for it in range(20):
        states, vel_out, lam, H = BA(it, states, vel, imu, t(pr["uv"])[None], t(pr["xyz"])[None], pr["ii"], pr["time_idx"],
                                     t(pr["intr"])[None], t(pr["conf"]), 1e-3, 1e-3, lam, t(pr["states_gt"][:, :7]),
                                     initialize=(it < 10))
        assert states.shape == (1, len(pr["time_idx"]), 10) and states.dtype == torch.float64
        assert vel_out is vel and isinstance(lam, float) and H.shape == (1, 9, 9)
        ref = g["states_hist"][it]; s = states[0].numpy()
        assert np.abs(s[:, :3] - ref[:, :3]).max() < 1e-3 and np.abs(s[:, 7:] - ref[:, 7:]).max() < 1e-6
        assert lam == g["lamda_hist"][it]
        assert rel(H[0].numpy(), g["hessian_hist"][it]) < 1e-6


def test_BA_accepts_unsorted_ii():
    g = load_golden("ba_T30"); pr = problem_from_golden(g)
    t = torch.tensor
    perm = np.random.default_rng(0).permutation(len(pr["ii"]))
    s, _, lam, _ = BA(0, t(pr["states0"])[None], t(pr["velocities"])[None], _imu(pr), t(pr["uv"][perm])[None],
                      t(pr["xyz"][perm])[None], pr["ii"][perm], pr["time_idx"], t(pr["intr"])[None], t(pr["conf"][perm]),
                      1e-3, 1e-3, 1e-4, t(pr["states_gt"][:, :7]), initialize=True)
    assert np.abs(s[0].numpy()[:, :3] - g["states_hist"][0][:, :3]).max() < 1e-3


def test_landmark_project_and_predict_shapes_match_reference():
    g = load_golden("ba_T30"); pr = problem_from_golden(g)
    t = torch.tensor
    T = len(pr["time_idx"])
    uv, Jg = U.landmark_project(t(pr["states0"])[None], t(pr["xyz"])[None], t(pr["intr"])[None], pr["ii"], jacobian=True)
    assert uv.shape == (1, len(pr["ii"]), 2) and Jg.shape == (len(pr["ii"]), 2, 9)
    assert np.array_equal(uv[0].numpy(), g["uv"]) and rel(Jg.numpy(), g["Jg"]) < 1e-9
    out = U.predict(t(pr["states0"])[None], _imu(pr), pr["time_idx"], 100, 100, jacobian=True)
    r_pred, pose_pred, vel_pred, z0, z1, Jf, Hq, qgrad = out
    assert r_pred.shape == (1, T - 1, 7) and Jf.shape == (1, 6 * (T - 1), 9 * T) and Hq.shape == (1, 9 * T, 9 * T)
    assert qgrad.shape == (1, T, 9) and pose_pred.shape == (1, T, 10) and vel_pred.shape == (1, T, 3) and z0 == 0
    Jf4 = Jf[0].numpy().reshape(T - 1, 6, T, 9); Hq4 = Hq[0].numpy().reshape(T, 9, T, 9)
    for i in range(T - 1):
        assert rel(Jf4[i][:, i][:, PV], g["pred_Jf_self"][i]) < 1e-9
        assert np.array_equal(Jf4[i][:, i + 1][:, PV], g["pred_Jf_next"][i])
        assert rel(Hq4[i, 3:6, i + 1, 3:6], g["pred_Hq_off"][i]) < 1e-9 and rel(Hq4[i + 1, 3:6, i, 3:6], g["pred_Hq_low"][i]) < 1e-9
    assert rel(np.stack([Hq4[i, 3:6, i, 3:6] for i in range(T)]), g["pred_Hq_diag"]) < 1e-9
    nz = np.abs(Jf4).sum() - sum(np.abs(Jf4[i][:, [i, i + 1]]).sum() for i in range(T - 1))
    assert nz == 0
    r3 = U.predict(t(pr["states0"])[None], _imu(pr), pr["time_idx"], 100, 100, jacobian=False)
    assert len(r3) == 3 and rel(r3[0][0].numpy(), g["pred_r_pred"]) < 1e-9
    init = U.predict(t(pr["states0"])[None], _imu(pr), pr["time_idx"], 100, 100, jacobian=True, initialize=True)
    assert init[0].shape == (1, T - 1, 6) and float(init[5].abs().sum()) == 0
    gp = U.predict_gpu(t(pr["states0"])[None], _imu(pr), pr["time_idx"], 100, 100, jacobian=False)
    assert rel(gp[1][0].numpy()[:, :3], g["skip_pos"]) < 1e-12


def test_propagate_dynamics_init_matches_reference():
    g = load_golden("helpers")
    t = torch.tensor
    s_t, v_t, s_full, v_full = U.propagate_dynamics_init(t(g["pdi_state"])[None], t(g["pdi_vel"])[None], t(g["pdi_omega"])[None], 4, 5, 1)
    assert rel(s_t[0].numpy(), g["pdi_states_t"]) < 1e-12 and rel(v_t[0].numpy(), g["pdi_vel_t"]) < 1e-12
    assert rel(s_full[0].numpy(), g["pdi_states_full"]) < 1e-12 and rel(v_full[0].numpy(), g["pdi_vel_full"]) < 1e-12


@pytest.mark.parametrize("name", ["seq_a", "seq_b"])
def test_ingest_kernels_bit_exact_vs_reference(name):
    """(f)2: read_detections / remove_elems of the mirror (device kernels for the integer indexing) against the
    reference's own outputs."""
    g = load_golden(name)
    orbit, ld, intr, time_idx, ii = od_pipe.read_detections(False, detections=g["dets"].copy(), orbit_np=g["orbit"].copy())
    assert np.array_equal(time_idx, g["rd_time_idx"]) and time_idx.dtype == g["rd_time_idx"].dtype
    assert np.array_equal(ii, g["rd_ii"]) and np.array_equal(orbit, g["rd_orbit"]) and np.array_equal(intr, g["rd_intr"])
    T = len(time_idx)
    dummy = torch.zeros((T, 3), dtype=torch.float64)
    poses = torch.arange(T * 7, dtype=torch.float64).reshape(T, 7)
    r = od_pipe.remove_elems(torch.tensor(g["vis_mask"]), dummy, dummy, poses, dummy, dummy, None, None, None, None, ii, time_idx)
    assert np.array_equal(r[9], g["re_ii"]) and np.array_equal(r[10], g["re_time_idx"])
    assert np.array_equal(r[11].numpy(), g["re_mask"]) and r[2].shape[0] == len(g["re_time_idx"])


def test_ingest_kernels_equal_oracle_on_random_and_edge_inputs():
    import ingest_oracle as io
    from vinsat_b200 import _lib
    ctx = _lib.default_context(0)
    rng = np.random.default_rng(0)
    cases = [np.array([500.0, 2000.0, 2000.0, 2500.0]),          # a detection ON a knot that was just inserted: duplicate time
             np.array([1000.0, 2000.0, 3000.0, 3001.0]),         # consecutive detections on knots
             np.array([7.0]), np.array([0.0, 0.0, 999.0, 1000.0, 1001.0, 5000.0])]
    for n in (50, 3000, 70000):
        fr = np.sort(rng.choice(np.arange(0, 12000), size=min(n, 4000) if n < 70000 else 9000, replace=False))
        cases.append(np.repeat(fr, rng.integers(1, 12, size=len(fr))).astype(np.float64))
    for frames in cases:
        for n_orbit in (int(frames[-1]) + 1, 10801, 500):
            t_o, ii_o = io.index_detections(frames, n_orbit)
            t_d, ii_d = ctx.index_detections(frames, n_orbit)
            assert np.array_equal(t_d, t_o) and np.array_equal(ii_d, ii_o), (len(frames), n_orbit)
            for dens in (0.0, 0.3, 1.0):
                mask = rng.random(len(frames)) < dens
                a = io.remove_elems_index(mask, ii_o, t_o)
                b = ctx.remove_elems_index(mask, ii_o, t_o)
                assert all(np.array_equal(x, y) for x, y in zip(a, b)), (len(frames), n_orbit, dens)
    with pytest.raises(_lib.VinsatError, match="sorted"):
        ctx.index_detections(np.array([5.0, 3.0]), 100)
    # remove_elems edge cases of the reference's semantics: one survivor / nothing survives (knots stay)
    time_idx = np.array([3, 8, 1000, 1004, 1010, 2000, 2005]); ii = np.array([0, 0, 1, 3, 3, 4, 6, 6])
    a = ctx.remove_elems_index(np.array([0, 0, 0, 1, 0, 0, 0, 0], dtype=bool), ii, time_idx)
    assert list(a[0]) == [1] and list(a[1]) == [1000, 1004, 2000]
    a = ctx.remove_elems_index(np.zeros(8, dtype=bool), ii, time_idx)
    assert len(a[0]) == 0 and list(a[1]) == [1000, 2000]


@pytest.mark.parametrize("name", ["seq_a", "seq_b"])
def test_visibility_mask_bit_exact(name):
    g = load_golden(name)
    orbit, ld, intr, time_idx, ii = od_pipe.read_detections(False, detections=g["dets"].copy(), orbit_np=g["orbit"].copy())
    r = od_pipe.process_ground_truths(orbit, ld, intr, 1.0, time_idx)
    states_gt = torch.cat([r[2], r[1][time_idx]], dim=-1)
    proj = U.landmark_project(states_gt.unsqueeze(0), r[5].unsqueeze(0), r[7].unsqueeze(0), ii, jacobian=False)
    assert np.array_equal(proj[0].numpy(), g["vis_proj"])
    mask = od_pipe.visibility_mask(proj, r[6], ld["confidence"])
    assert np.array_equal(mask.numpy(), g["vis_mask"])


@pytest.mark.parametrize("name", ["seq_a", "seq_b"])
def test_streaming_version_matches_reference(name):
    g = load_golden(name)
    errors, first_detection, times = od_pipe.streaming_version(detections=g["dets"].copy(), orbit_np=g["orbit"].copy())
    tcat = np.concatenate([np.asarray(x).reshape(-1) for x in times]).astype(np.int64)
    assert np.array_equal(tcat, g["sv_times"])
    assert np.array_equal(np.array([len(np.asarray(x).reshape(-1)) for x in times]), g["sv_times_len"])
    assert int(first_detection) == int(g["sv_first_detection"])
    e = errors.numpy()
    assert e.shape == g["sv_errors"].shape
    assert np.abs(e - g["sv_errors"]).max() < 1e-3, np.abs(e - g["sv_errors"]).max()       # 1 m


def test_generate_new_traj():
    np.random.seed(3)
    traj, tsamp = trajgen_pipe.generate_new_traj('polar')
    assert traj.shape == (13, 10801) and len(tsamp) == 10801
    x = traj[:6, 0].copy()
    for _ in range(50):
        x = trajgen_pipe.orbit_step(x, 1.0)
    assert rel(traj[:6, 50], x) < 1e-12
    with pytest.raises(ValueError):
        trajgen_pipe.generate_new_traj('iss', strict=True)      # the reference's own failure


def test_satcam_class_and_sweep():
    import satcam_oracle as so
    from vinsat_b200.sim import SatCam as SC
    lm_ecef, rows, names = SC.all_landmark_centroids_ecef()
    assert lm_ecef.shape == (16825, 3) and len(names) == 34
    # a nadir pose above a landmark of region 17R
    r17 = SC.load_landmarks(["17R"])["17R"]
    tgt = SC.lonlat_to_ecef(r17[100, 0], r17[100, 1])
    up = tgt / np.linalg.norm(tgt)
    pos = up * (np.linalg.norm(tgt) + 600e3)
    d = -up
    east = np.cross(np.array([0, 0, 1.0]), up); east /= np.linalg.norm(east)
    north = np.cross(up, east)
    pose = np.concatenate([pos, d, north, east])
    cam = SC.SatCam(SC.SatellitePose(pose), 66.0, 4608, 2592)
    px = cam.ecef_pos_to_px(tgt)
    assert np.abs(px - np.array([2304.0, 1296.0])).max() < 1e-6            # the boresight hits the image centre
    uvo, _ = so.project(pose[None], tgt[None], 66.0, 4608, 2592)
    assert np.array_equal(px, uvo[0, 0])
    cl = cam.get_corner_lonlats()
    co, ho = so.corners(pose[None], 66.0, 4608, 2592)
    for k, key in enumerate(("tl", "tr", "br", "bl")):
        lon, lat = so.ecef_to_lonlat(co[0, k])
        assert abs(cl[key][0] - float(lon)) < 1e-13 and abs(cl[key][1] - float(lat)) < 1e-13   # device atan2 vs libm
    regions = cam.find_current_regions()
    assert "17R" in regions
    n = cam.check_for_landmarks_in_region("17R")
    tl, br = cl["tl"], cl["br"]
    assert n == min(3, int(so.landmarks_in_footprint(tl, br, r17[:, 0], r17[:, 1]).sum()))
    assert cam.check_for_all_landmarks() == (n >= 3)
    # batched projection sweep == oracle, bit-exact in-frame sets
    rng = np.random.default_rng(0)
    poses = np.tile(pose, (40, 1))
    shift = rng.normal(0, 3e5, size=(40, 3))
    poses[:, :3] += shift - (shift @ up)[:, None] * up
    counts, mask = SC.inframe_sweep(poses, lm_ecef, chunk=16, want_mask=True)
    _, mo = so.project(poses, lm_ecef, 66.0, 4608, 2592)
    assert np.array_equal(mask.astype(bool), mo) and np.array_equal(counts, mo.sum(1)) and counts.max() > 10


def test_monte_carlo_runner_shards_without_overlap():
    from vinsat_b200.eval import batch_runner
    full = batch_runner.run_od_monte_carlo(6, frames=30, obs_per_frame=6, seed0=40)
    a = batch_runner.run_od_monte_carlo(6, frames=30, obs_per_frame=6, seed0=40, rank=0, world_size=2)
    b = batch_runner.run_od_monte_carlo(6, frames=30, obs_per_frame=6, seed0=40, rank=1, world_size=2)
    assert list(a["problem_ids"]) + list(b["problem_ids"]) == list(full["problem_ids"])
    assert np.array_equal(np.concatenate([a["states"], b["states"]]), full["states"])   # sharding changes nothing
    assert full["pos_err_km"].max() < 30.0
