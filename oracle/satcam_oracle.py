"""CPU oracle for the SatCam geometry (sim/SatCam.py) -- NumPy fp64 restatement.

TEST INFRASTRUCTURE ONLY (same rules as ba_oracle.py).

PARITY UNPINNED at two third-party boundaries: ``sim/SatCam.py`` cannot be imported in the build
container (astropy, pyproj, rasterio are not installed), the reference has no
tests or golden vectors for it, and astropy's geocentric->geodetic / pyproj's geodetic->geocentric are
third-party code with unpinned versions.  This file therefore restates the reference's own arithmetic
line by line (citations: path:line under /root/reference/sim) and *defines* the evaluation order the CUDA
kernels are compared against bit for bit: every expression is evaluated left to right with separately
rounded operations (NumPy never contracts to FMA), dot products are accumulated sequentially.
The WGS84 geodetic conversions use the closed forms of BA_utils.py:1221-1236 (scaled to metres).
"""
import math

import numpy as np

A_M = 6378137.0
C_M = 6356752.314245


def intrinsics(hfov_deg, w_px, h_px):
    """SatCam.py:44-49,57-59 -> (f, cx, cy)."""
    half_angle = (hfov_deg * (math.pi / 180.0)) / 2.0
    f = (w_px / 2.0) / math.tan(half_angle)
    return f, w_px / 2.0, h_px / 2.0


def _rows(poses):
    """Rows of R_cw = right, -up, dir (SatCam.py:50-56,81-84)."""
    return np.stack([poses[:, 9:12], -poses[:, 6:9], poses[:, 3:6]], axis=1)       # (P,3,3)


def cam_matrix(poses, hfov_deg, w_px, h_px):
    """C_cw = K [R_cw | -R_cw p] (SatCam.py:87-92).  poses (P,12) -> (P,3,4)."""
    f, cx, cy = intrinsics(hfov_deg, w_px, h_px)
    R = _rows(poses)
    p = poses[:, 0:3]
    t = (R[:, :, 0] * p[:, None, 0] + R[:, :, 1] * p[:, None, 1]) + R[:, :, 2] * p[:, None, 2]    # (P,3)
    E = np.concatenate([R, -t[:, :, None]], axis=2)                                                  # (P,3,4)
    return np.stack([f * E[:, 0] + cx * E[:, 2], f * E[:, 1] + cy * E[:, 2], E[:, 2]], axis=1)


def project(poses, landmarks_ecef, hfov_deg, w_px, h_px):
    """uv = (C X)_{0:2}/(C X)_2 (SatCam.py:149-154) for all pose x landmark pairs.
    Returns uv (P,L,2), inframe (P,L) bool: in front, 0<=u<w, 0<=v<h."""
    Cm = cam_matrix(poses, hfov_deg, w_px, h_px)[:, None]      # (P,1,3,4)
    X = landmarks_ecef[None]                                   # (1,L,3)
    uvw = ((Cm[..., 0] * X[..., None, 0] + Cm[..., 1] * X[..., None, 1]) + Cm[..., 2] * X[..., None, 2]) + Cm[..., 3]
    with np.errstate(divide="ignore", invalid="ignore"):
        u = uvw[..., 0] / uvw[..., 2]
        v = uvw[..., 1] / uvw[..., 2]
    inframe = (uvw[..., 2] > 0) & (u >= 0) & (u < w_px) & (v >= 0) & (v < h_px)
    return np.stack([u, v], axis=-1), inframe


def corners(poses, hfov_deg, w_px, h_px):
    """get_corner_vectors + cast_ray_to_earth (SatCam.py:94-147) for tl, tr, br, bl.
    Returns corners (P,4,3) ECEF metres (0 where missed) and hit (P,4) bool."""
    f, cx, cy = intrinsics(hfov_deg, w_px, h_px)
    P = poses.shape[0]
    px = np.array([0.0, w_px, w_px, 0.0])
    py = np.array([0.0, 0.0, h_px, h_px])
    kx = ((px - cx) / f)[None]           # K^-1 [px,py,1] in closed form
    ky = ((py - cy) / f)[None]
    rw = np.stack([poses[:, 9:12], -poses[:, 6:9], poses[:, 3:6]], axis=2)       # columns right,-up,dir
    v3 = (rw[:, None, :, 0] * kx[..., None] + rw[:, None, :, 1] * ky[..., None]) + rw[:, None, :, 2] * 1.0
    nrm = np.sqrt((v3[..., 0] * v3[..., 0] + v3[..., 1] * v3[..., 1]) + v3[..., 2] * v3[..., 2])
    u, v, w = v3[..., 0] / nrm, v3[..., 1] / nrm, v3[..., 2] / nrm
    x, y, z = poses[:, None, 0], poses[:, None, 1], poses[:, None, 2]
    a = b = A_M
    c = C_M
    a2, b2, c2 = a * a, b * b, c * c
    a2b2, a2c2, b2c2 = a2 * b2, a2 * c2, b2 * c2
    value = ((-a2b2) * w * z - a2c2 * v * y) - b2c2 * u * x                                   # SatCam.py:133
    w2, v2, u2, x2, y2, z2 = w * w, v * v, u * u, x * x, y * y, z * z
    rad = a2b2 * w2                                                                           # SatCam.py:134
    rad = rad + a2c2 * v2
    rad = rad - a2 * v2 * z2
    rad = rad + 2.0 * a2 * v * w * y * z
    rad = rad - a2 * w2 * y2
    rad = rad + b2c2 * u2
    rad = rad - b2 * u2 * z2
    rad = rad + 2.0 * b2 * u * w * x * z
    rad = rad - b2 * w2 * x2
    rad = rad - c2 * u2 * y2
    rad = rad + 2.0 * c2 * u * v * x * y
    rad = rad - c2 * v2 * x2
    mag = (a2b2 * w2 + a2c2 * v2) + b2c2 * u2                                                 # SatCam.py:135
    ok = ~(rad < 0)
    with np.errstate(invalid="ignore"):
        d = (value - a * b * c * np.sqrt(rad)) / mag                                          # SatCam.py:139
    ok = ok & ~(d < 0)
    pts = np.stack([x + d * u, y + d * v, z + d * w], axis=-1)
    pts = np.where(ok[..., None], pts, 0.0)
    return pts, ok


def ecef_to_lonlat(p):
    """WGS84 geocentric -> geodetic lon/lat in degrees for points ON the ellipsoid (closed form:
    tan(lat) = z / ((1-e^2) sqrt(x^2+y^2)); stands in for astropy EarthLocation.from_geocentric,
    SatCam.py:181 -- parity unpinned there)."""
    e2 = 1.0 - (C_M * C_M) / (A_M * A_M)
    lon = np.degrees(np.arctan2(p[..., 1], p[..., 0]))
    lat = np.degrees(np.arctan2(p[..., 2], (1.0 - e2) * np.sqrt(p[..., 0] ** 2 + p[..., 1] ** 2)))
    return lon, lat


def lonlat_to_ecef(lon_deg, lat_deg, alt=0.0):
    """WGS84 geodetic -> ECEF metres (BA_utils.py:1221-1236 form; stands in for pyproj, SatCam.py:194-199)."""
    phi, lam = np.radians(lat_deg), np.radians(lon_deg)
    e2 = 1.0 - (C_M * C_M) / (A_M * A_M)
    N = A_M / np.sqrt(1 - e2 * np.sin(phi) ** 2)
    return np.stack([(N + alt) * np.cos(phi) * np.cos(lam), (N + alt) * np.cos(phi) * np.sin(lam),
                     ((1 - e2) * N + alt) * np.sin(phi)], axis=-1)


def landmarks_in_footprint(corner_lonlat_tl, corner_lonlat_br, lm_lon, lm_lat):
    """check_for_landmarks_in_region box test (SatCam.py:247): strict inequalities."""
    tl_lon, tl_lat = corner_lonlat_tl
    br_lon, br_lat = corner_lonlat_br
    return (lm_lon > tl_lon) & (lm_lon < br_lon) & (lm_lat > br_lat) & (lm_lat < tl_lat)
