"""CPU oracle for the SatCam geometry (sim/SatCam.py) -- ctypes front end of ``satcam_oracle.c``.

TEST INFRASTRUCTURE ONLY (same rules as ba_oracle.py): imported by ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline leg; never by ``vinsat_b200/``.

The arithmetic lives in plain C (``oracle/satcam_oracle.c``, built by ``oracle/Makefile`` into
``oracle/_build/libsatcam_oracle.so``; compiled on first import when the prebuilt file is missing) because the
reference's results depend on where NumPy/OpenBLAS fuses multiply-adds, and C has an exact ``fma()``.

PARITY PINNED: ``tests/golden/satcam.npz`` holds outputs of the unmodified reference class on 3039 poses
(``tests/golden/make_golden_satcam.py``); ``tests/test_oracle_vs_golden.py`` compares this oracle with them bit for
bit (camera matrices, corner rays, ellipsoid hits, pixel coordinates) and exactly (regions, visibility).
Two boundaries stay unpinned by the reference because its dependencies are absent: astropy's
geocentric->geodetic (SatCam.py:181) and pyproj's geodetic->geocentric (:194-199) are replaced by the WGS84
closed forms below, in the golden generator too.
"""
import ctypes
import os
import subprocess

import numpy as np

A_M = 6378137.0
C_M = 6356752.314245
_E2 = 1.0 - (C_M * C_M) / (A_M * A_M)
_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsatcam_oracle.so")
_lib = None

_dp = ctypes.POINTER(ctypes.c_double)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_ip = ctypes.POINTER(ctypes.c_int)
_i64p = ctypes.POINTER(ctypes.c_int64)


def _ptr(a, t):
    return None if a is None else a.ctypes.data_as(t)


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "satcam_oracle.c")
        if not os.path.exists(_SO) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_SO)):
            subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
        L = ctypes.CDLL(_SO)
        L.so_intrinsics.argtypes = [ctypes.c_double, ctypes.c_int, ctypes.c_int, _dp, _dp]
        L.so_cam_matrix.argtypes = [ctypes.c_int64, _dp, ctypes.c_double, ctypes.c_int, ctypes.c_int, _dp]
        L.so_project.argtypes = [ctypes.c_int64, _dp, ctypes.c_int64, _dp, ctypes.c_double, ctypes.c_int,
                                 ctypes.c_int, _dp, _u8p]
        L.so_corners.argtypes = [ctypes.c_int64, _dp, ctypes.c_double, ctypes.c_int, ctypes.c_int, _dp, _dp, _u8p]
        L.so_lonlat.argtypes = [ctypes.c_int64, _dp, _dp, _dp]
        L.so_get_region.argtypes = [ctypes.c_double, ctypes.c_double]
        L.so_get_region.restype = ctypes.c_int
        L.so_current_regions.argtypes = [_dp, _dp, _u8p, _ip, ctypes.c_int]
        L.so_current_regions.restype = ctypes.c_int
        L.so_check_for_all_landmarks.argtypes = [_dp, _dp, _u8p, ctypes.c_int, _ip, _i64p, _dp, ctypes.c_int, _ip,
                                                 _ip, ctypes.c_int, _ip]
        L.so_check_for_all_landmarks.restype = ctypes.c_int
        _lib = L
    return _lib


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a if shape is None else a.reshape(shape)


def region_code(name):
    """'10S' -> 10*32 + (ord('S')-64); None -> -1 (the encoding of tests/golden/satcam.npz)."""
    return -1 if name is None else int(name[:2]) * 32 + (ord(name[2]) - 64)


def region_name(code):
    return None if code < 0 else "%02d%s" % (code // 32, chr(64 + code % 32))


def intrinsics(hfov_deg, w_px, h_px):
    """SatCam.py:44-49,57-59 -> (f, cx, cy)."""
    f = ctypes.c_double()
    K = np.zeros(9)
    lib().so_intrinsics(float(hfov_deg), int(w_px), int(h_px), ctypes.byref(f), _ptr(K, _dp))
    return f.value, w_px / 2.0, h_px / 2.0


def K_inv(hfov_deg, w_px, h_px):
    f = ctypes.c_double()
    K = np.zeros(9)
    lib().so_intrinsics(float(hfov_deg), int(w_px), int(h_px), ctypes.byref(f), _ptr(K, _dp))
    return K.reshape(3, 3)


def cam_matrix(poses, hfov_deg, w_px, h_px):
    """C_cw = K [R_cw | -R_cw p] (SatCam.py:87-92).  poses (P,12) -> (P,3,4)."""
    poses = _f64(poses, (-1, 12))
    C = np.zeros((poses.shape[0], 3, 4))
    lib().so_cam_matrix(poses.shape[0], _ptr(poses, _dp), float(hfov_deg), int(w_px), int(h_px), _ptr(C, _dp))
    return C


def project(poses, landmarks_ecef, hfov_deg, w_px, h_px):
    """uv = (C X)_{0:2}/(C X)_2 (SatCam.py:149-154) for all pose x landmark pairs.
    Returns uv (P,L,2), inframe (P,L) bool: in front, 0<=u<w, 0<=v<h."""
    poses = _f64(poses, (-1, 12))
    lm = _f64(landmarks_ecef, (-1, 3))
    uv = np.zeros((poses.shape[0], lm.shape[0], 2))
    inf = np.zeros((poses.shape[0], lm.shape[0]), dtype=np.uint8)
    lib().so_project(poses.shape[0], _ptr(poses, _dp), lm.shape[0], _ptr(lm, _dp), float(hfov_deg), int(w_px),
                     int(h_px), _ptr(uv, _dp), _ptr(inf, _u8p))
    return uv, inf.astype(bool)


def corner_rays(poses, hfov_deg, w_px, h_px):
    """get_corner_vectors + cast_ray_to_earth (SatCam.py:94-147) for tl, tr, br, bl.
    Returns unit vectors (P,4,3), points (P,4,3) ECEF metres (0 where missed), hit (P,4) bool."""
    poses = _f64(poses, (-1, 12))
    P = poses.shape[0]
    vec = np.zeros((P, 4, 3))
    pts = np.zeros((P, 4, 3))
    hit = np.zeros((P, 4), dtype=np.uint8)
    lib().so_corners(P, _ptr(poses, _dp), float(hfov_deg), int(w_px), int(h_px), _ptr(vec, _dp), _ptr(pts, _dp),
                     _ptr(hit, _u8p))
    return vec, pts, hit.astype(bool)


def corners(poses, hfov_deg, w_px, h_px):
    _, pts, hit = corner_rays(poses, hfov_deg, w_px, h_px)
    return pts, hit


def ecef_to_lonlat(p):
    """WGS84 geocentric -> geodetic lon/lat in degrees for points ON the ellipsoid (closed form:
    tan(lat) = z / ((1-e^2) sqrt(x^2+y^2)); stands in for astropy EarthLocation.from_geocentric, SatCam.py:181)."""
    p = _f64(p)
    flat = p.reshape(-1, 3)
    lon = np.zeros(flat.shape[0])
    lat = np.zeros(flat.shape[0])
    lib().so_lonlat(flat.shape[0], _ptr(flat, _dp), _ptr(lon, _dp), _ptr(lat, _dp))
    return lon.reshape(p.shape[:-1]), lat.reshape(p.shape[:-1])


def lonlat_to_ecef(lon_deg, lat_deg, alt=0.0):
    """WGS84 geodetic -> ECEF metres (BA_utils.py:1221-1236 form; stands in for pyproj, SatCam.py:194-199)."""
    phi, lam = np.radians(lat_deg), np.radians(lon_deg)
    N = A_M / np.sqrt(1 - _E2 * np.sin(phi) ** 2)
    return np.stack([(N + alt) * np.cos(phi) * np.cos(lam), (N + alt) * np.cos(phi) * np.sin(lam),
                     ((1 - _E2) * N + alt) * np.sin(phi)], axis=-1)


def get_region(lon, lat):
    """get_region (SatCam.py:187-191) -> region code (-1 = None); scalars or arrays."""
    L = lib()
    lon, lat = np.broadcast_arrays(np.asarray(lon, dtype=np.float64), np.asarray(lat, dtype=np.float64))
    out = np.array([L.so_get_region(float(a), float(b)) for a, b in zip(lon.ravel(), lat.ravel())], dtype=np.int32)
    return out.reshape(lon.shape)


def landmarks_in_footprint(corner_lonlat_tl, corner_lonlat_br, lm_lon, lm_lat):
    """check_for_landmarks_in_region box test (SatCam.py:247): strict inequalities."""
    tl_lon, tl_lat = corner_lonlat_tl
    br_lon, br_lat = corner_lonlat_br
    return (lm_lon > tl_lon) & (lm_lon < br_lon) & (lm_lat > br_lat) & (lm_lat < tl_lat)


def check_for_all_landmarks(poses, hfov_deg, w_px, h_px, lm_codes, lm_off, lm_lonlat, active_codes):
    """find_current_regions + check_for_all_landmarks (SatCam.py:203-262) per pose.
    Landmark table: region code lm_codes[r] owns rows lm_off[r]:lm_off[r+1] of lm_lonlat (centroid lon, lat; CSV
    order).  Returns visible (P,) bool, corner lon/lat (P,4,2; NaN where the ray missed), hit (P,4) bool and the
    list of per-pose current-region code lists."""
    L = lib()
    _, pts, hit = corner_rays(poses, hfov_deg, w_px, h_px)
    lon, lat = ecef_to_lonlat(pts)
    P = pts.shape[0]
    lm_codes = np.ascontiguousarray(lm_codes, dtype=np.int32)
    lm_off = np.ascontiguousarray(lm_off, dtype=np.int64)
    lm_lonlat = _f64(lm_lonlat, (-1, 2))
    active = np.ascontiguousarray(active_codes, dtype=np.int32)
    vis = np.zeros(P, dtype=bool)
    cur = np.zeros(512, dtype=np.int32)
    ncur = ctypes.c_int()
    curs = []
    hit8 = np.ascontiguousarray(hit, dtype=np.uint8)
    for i in range(P):
        lo = np.ascontiguousarray(lon[i])
        la = np.ascontiguousarray(lat[i])
        vis[i] = L.so_check_for_all_landmarks(_ptr(lo, _dp), _ptr(la, _dp), _ptr(hit8[i], _u8p), len(lm_codes),
                                              _ptr(lm_codes, _ip), _ptr(lm_off, _i64p), _ptr(lm_lonlat, _dp),
                                              len(active), _ptr(active, _ip), _ptr(cur, _ip), 512, ctypes.byref(ncur))
        curs.append(cur[:ncur.value].copy())
    ll = np.stack([lon, lat], axis=-1)
    ll[~hit] = np.nan
    return vis, ll, hit, curs
