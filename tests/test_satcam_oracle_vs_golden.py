"""Pins the SatCam oracle (oracle/satcam_oracle.c) to outputs of the UNMODIFIED reference class
sim/SatCam.py (tests/golden/satcam.npz, produced by tests/golden/make_golden_satcam.py).  CPU only."""
import hashlib

import numpy as np

import satcam_oracle as so
from conftest import load_golden


def landmark_table():
    """Product data file (sim/landmark_csvs/*.csv as one table) in the oracle's layout."""
    from vinsat_b200.sim import SatCam as SC
    lm = SC.load_landmarks()
    names = sorted(lm)
    rows = np.concatenate([lm[n] for n in names])
    off = np.cumsum([0] + [len(lm[n]) for n in names])
    codes = np.array([so.region_code(n) for n in names], dtype=np.int32)
    return names, codes, off, rows


def test_landmark_table_is_the_reference_csvs():
    g = load_golden("satcam")
    _, _, _, rows = landmark_table()
    assert hashlib.sha256(np.ascontiguousarray(rows[:, :2]).tobytes()).hexdigest() == str(g["lm_lonlat_sha256"])


def test_camera_matrix_bit_exact():
    g = load_golden("satcam")
    C = so.cam_matrix(g["poses"], float(g["hfov"]), int(g["w_px"]), int(g["h_px"]))
    assert np.array_equal(C, g["C_cw"])
    assert np.array_equal(so.K_inv(float(g["hfov"]), int(g["w_px"]), int(g["h_px"])), g["K_inv"])
    assert so.intrinsics(float(g["hfov"]), int(g["w_px"]), int(g["h_px"]))[0] == float(g["f"])


def test_corner_rays_and_ellipsoid_hits_bit_exact():
    g = load_golden("satcam")
    vec, pts, hit = so.corner_rays(g["poses"], float(g["hfov"]), int(g["w_px"]), int(g["h_px"]))
    assert np.array_equal(vec, g["corner_vec"])
    assert np.array_equal(hit, g["corner_hit"].astype(bool))
    assert np.array_equal(pts, g["corner_pts"])
    lon, lat = so.ecef_to_lonlat(pts)
    h = hit
    assert np.array_equal(lon[h], g["corner_lonlat"][..., 0][h])
    assert np.array_equal(lat[h], g["corner_lonlat"][..., 1][h])


def test_pixel_projection_bit_exact():
    g = load_golden("satcam")
    uv, _ = so.project(g["poses"][g["pose_idx"]], g["lm_ecef"], float(g["hfov"]), int(g["w_px"]), int(g["h_px"]))
    assert np.array_equal(uv, g["uv"])


def test_get_region_lattice():
    g = load_golden("satcam")
    assert np.array_equal(so.get_region(g["grid_lon"], g["grid_lat"]), g["grid_region"])


def test_current_regions_and_visibility_exact():
    g = load_golden("satcam")
    from vinsat_b200.sim import SatCam as SC
    _, codes, off, rows = landmark_table()
    active = np.array([so.region_code(n) for n in SC.DEFAULT_REGIONS], dtype=np.int32)
    vis, ll, hit, curs = so.check_for_all_landmarks(g["poses"], float(g["hfov"]), int(g["w_px"]), int(g["h_px"]),
                                                    codes, off, rows[:, :2], active)
    assert np.array_equal(vis, g["visible"].astype(bool))
    assert int(vis.sum()) > 300
    co = g["cur_off"]
    for i in range(len(vis)):
        assert np.array_equal(curs[i], g["cur_codes"][co[i]:co[i + 1]]), i
    creg = so.get_region(np.nan_to_num(ll[..., 0], nan=1e9), np.nan_to_num(ll[..., 1], nan=1e9))
    assert np.array_equal(creg[hit], g["corner_region"][hit])
