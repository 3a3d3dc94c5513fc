"""Mirror of ``sim/SatCam.py`` (live part, :15-262): pinhole camera in ECEF, world->pixel matrix, corner
ray-cast to the WGS84 ellipsoid, MGRS region lookup, landmark-in-footprint test -- plus the batched entry
points the trajectory-generation sweep needs (config 5 of BASELINE.json).  Geometry runs on the device
(``vinsat_satcam_project`` / ``vinsat_satcam_corner_rays`` / ``vinsat_satcam_visibility``), in the reference's own
rounding order (pinned bit for bit by ``tests/golden/satcam.npz``, generated from the reference class).

Third-party boundaries of the reference that are NOT available here (astropy geocentric->geodetic,
SatCam.py:181; pyproj geodetic->geocentric, :194-199) are replaced by WGS84 closed forms; parity there is
unpinned.  The `best_classes/*.npy` landmark filter (:234,244) is a missing blob and is
treated as "all classes".  Image access (rasterio / cv2, :264-660) is out of scope.
"""
import os

import numpy as np

from .. import _lib, config
from .getMGRS import getMGRS

_DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "landmarks_mgrs.npz")
A_M = 6378137.0
C_M = 6356752.314245
_E2 = 1.0 - (C_M * C_M) / (A_M * A_M)
DEFAULT_REGIONS = ['10S', '10T', '11R', '12R', '16T', '17R', '17T', '18S',                 # SatCam.py:64-65
                   '32S', '32T', '33S', '33T', '52S', '53S', '54S', '54T']


def load_landmarks(regions=None):
    """-> dict region -> (n,6) rows [centroid lon, lat, tl lon, lat, br lon, lat] (sim/landmark_csvs/*.csv)."""
    z = np.load(_DATA)
    names = [str(r) for r in z["regions"]]
    off = z["offsets"]
    out = {n: z["rows"][off[i]:off[i + 1]] for i, n in enumerate(names)}
    if regions is not None:
        out = {k: v for k, v in out.items() if k in regions}
    return out


def lonlat_to_ecef(lon_deg, lat_deg, alt=0.0):
    """WGS84 geodetic -> ECEF metres (stands in for pyproj, SatCam.py:194-199)."""
    phi, lam = np.radians(lat_deg), np.radians(lon_deg)
    N = A_M / np.sqrt(1 - _E2 * np.sin(phi) ** 2)
    return np.stack([(N + alt) * np.cos(phi) * np.cos(lam), (N + alt) * np.cos(phi) * np.sin(lam),
                     ((1 - _E2) * N + alt) * np.sin(phi)], axis=-1)


def ecef_to_lonlat(p):
    """Geodetic lon/lat (deg) of points ON the ellipsoid (stands in for astropy, SatCam.py:181)."""
    lon = np.degrees(np.arctan2(p[..., 1], p[..., 0]))
    lat = np.degrees(np.arctan2(p[..., 2], (1.0 - _E2) * np.sqrt(p[..., 0] ** 2 + p[..., 1] ** 2)))
    return lon, lat


class Vector3D:
    def __init__(self, xyz):
        self.x, self.y, self.z = xyz[0], xyz[1], xyz[2]

    def get(self):
        return np.array((self.x, self.y, self.z))


class SatellitePose:                                            # SatCam.py:23-37
    def __init__(self, pose_array):
        self.array = np.asarray(pose_array, dtype=np.float64)[:12].copy()
        self.position = Vector3D(pose_array[0:3])
        self.dir_vec = Vector3D(pose_array[3:6])
        self.up_vec = Vector3D(pose_array[6:9])
        self.right_vec = Vector3D(pose_array[9:12])

    def get_position(self): return self.position.get()
    def get_dir_vec(self): return self.dir_vec.get()
    def get_up_vec(self): return self.up_vec.get()
    def get_right_vec(self): return self.right_vec.get()


def region_code(name):
    """'10S' -> 10*32 + 19 (letter code 'A'=1): the integer form the C ABI uses for MGRS cells; None -> -1."""
    return -1 if name is None else int(name[:2]) * 32 + (ord(name[2]) - 64)


def region_name(code):
    return None if code < 0 else "%02d%s" % (code // 32, chr(64 + code % 32))


def landmark_table(ctx, regions=None):
    """Device-resident landmark table (all shipped CSVs) with `regions` as the active set (SatCam.py:63-67);
    cached ON the context per active set (it lives and dies with the context's device memory)."""
    active = tuple(DEFAULT_REGIONS if regions is None else regions)
    tables = ctx.__dict__.setdefault("_satcam_tables", {})
    if active not in tables:
        lm = load_landmarks()
        names = sorted(lm)
        off = np.cumsum([0] + [len(lm[n]) for n in names])
        rows = np.concatenate([lm[n][:, :2] for n in names])
        tables[active] = ctx.satcam_table([region_code(n) for n in names], off, rows, [region_code(n) for n in active])
    return tables[active]


class SatCam:
    def __init__(self, sat_pose, hfov, w_px, h_px, regions=None):          # SatCam.py:39-75
        self.hfov, self.w_px, self.h_px = hfov, w_px, h_px
        half_angle = np.deg2rad(hfov) / 2
        self.f = (w_px / 2) / np.tan(half_angle)
        self.vfov = np.rad2deg(2 * np.arctan((h_px / 2) / self.f))
        self.K = np.array([[self.f, 0, w_px / 2], [0, self.f, h_px / 2], [0, 0, 1]])
        self.regions = regions if regions is not None else list(DEFAULT_REGIONS)
        self.grid = getMGRS()
        self._ctx = _lib.default_context(config.device)
        self._table = landmark_table(self._ctx, self.regions)
        self.update_pose(sat_pose)

    def update_pose(self, sat_pose):                                        # SatCam.py:77-85
        self.sat_pose = sat_pose
        self._pose12 = np.concatenate([sat_pose.get_position(), sat_pose.get_dir_vec(), sat_pose.get_up_vec(),
                                       sat_pose.get_right_vec()]).astype(np.float64)
        self.sat_pos = sat_pose.get_position()
        self.dir_vec = sat_pose.get_dir_vec()
        self.up_vec = -sat_pose.get_up_vec()
        self.right_vec = sat_pose.get_right_vec()
        self.R_wc = np.stack((self.right_vec, self.up_vec, self.dir_vec), axis=1)
        self.R_cw = self.R_wc.T
        self.C_cw = self.world_to_pixel_mat()

    def world_to_pixel_mat(self):                                           # SatCam.py:87-92
        return self._ctx.satcam_cam_matrix(self._pose12[None], self.hfov, self.w_px, self.h_px)[0]

    def get_corner_vectors(self):                                           # SatCam.py:98-104
        _, _, vec = self._ctx.satcam_corners(self._pose12[None], self.hfov, self.w_px, self.h_px, want_rays=True)
        return {k: vec[0, i] for i, k in enumerate(('tl', 'tr', 'br', 'bl'))}

    def ecef_pos_to_px(self, pos):                                          # SatCam.py:149-154
        pos = np.asarray(pos, dtype=np.float64)
        pts = pos.reshape(3, -1).T if pos.ndim == 2 and pos.shape[0] == 3 else pos.reshape(-1, 3)
        uv, _, _ = self._ctx.satcam_project(self._pose12[None], pts, self.hfov, self.w_px, self.h_px,
                                            want_mask=False, want_count=False)
        uv = uv[0]
        return uv[0] if pos.ndim == 1 else uv.T

    def lonlat_to_pixel_coords(self, lon, lat):                             # SatCam.py:193-201
        return self.ecef_pos_to_px(lonlat_to_ecef(np.asarray(lon, dtype=np.float64), np.asarray(lat, dtype=np.float64)).T)

    def _visibility(self):
        vis, cnt, ll, reg = self._ctx.satcam_visibility(self._table, self._pose12[None], self.hfov, self.w_px,
                                                        self.h_px, want_count=True, want_corners=True)
        return bool(vis[0]), int(cnt[0]), ll[0], reg[0]

    def get_corner_lonlats(self):                                           # SatCam.py:175-185
        _, _, ll, _ = self._visibility()
        return {key: (None if np.isnan(ll[k, 0]) else (float(ll[k, 0]), float(ll[k, 1])))
                for k, key in enumerate(('tl', 'tr', 'br', 'bl'))}

    def get_region(self, lon, lat):                                         # SatCam.py:187-191
        for key, bounds in self.grid.items():
            if bounds[0] <= lon <= bounds[2] and bounds[1] <= lat <= bounds[3]:
                return key
        return None

    def find_current_regions(self):                                         # SatCam.py:203-230
        _, _, ll, reg = self._visibility()
        self.corner_lonlats = {key: (None if np.isnan(ll[k, 0]) else (float(ll[k, 0]), float(ll[k, 1])))
                               for k, key in enumerate(('tl', 'tr', 'br', 'bl'))}
        regions = [region_name(int(reg[k])) for k in range(4) if not np.isnan(ll[k, 0])]
        if not regions:
            raise UnboundLocalError("no image corner intersects the Earth (the reference fails here too, SatCam.py:217)")
        nums = [int(r[:2]) for r in regions if r is not None]
        chars = [r[2] for r in regions if r is not None]
        if regions[-1] is not None:       # the reference tests the loop variable of its last iteration (:217)
            num_range = [58, 59, 60, 1, 2, 3] if (min(nums) < 4 and max(nums) > 57) else range(min(nums), max(nums) + 1)
            self.current_regions = [str(n).zfill(2) + chr(c) for n in num_range
                                    for c in range(ord(min(chars)), ord(max(chars)) + 1)]
        else:
            self.current_regions = regions
        return self.current_regions

    def check_for_landmarks_in_region(self, region):                        # SatCam.py:232-251
        """Count (capped at 3 by the reference's early return) of the region's centroids inside the tl/br box."""
        cl = self.corner_lonlats
        lm = load_landmarks().get(region)
        if cl['tl'] is None or cl['br'] is None or lm is None:
            return 0
        tl_lon, tl_lat = cl['tl']
        br_lon, br_lat = cl['br']
        inside = (lm[:, 0] > tl_lon) & (lm[:, 0] < br_lon) & (lm[:, 1] > br_lat) & (lm[:, 1] < tl_lat)
        return int(min(inside.sum(), 3))

    def check_for_all_landmarks(self):                                      # SatCam.py:254-262
        vis, _, ll, _ = self._visibility()
        self.corner_lonlats = {key: (None if np.isnan(ll[k, 0]) else (float(ll[k, 0]), float(ll[k, 1])))
                               for k, key in enumerate(('tl', 'tr', 'br', 'bl'))}
        return vis


# ---- batched sweeps (config 5): many poses at once ------------------------------------------------------------
def all_landmark_centroids_ecef(regions=None):
    lm = load_landmarks(regions)
    names = sorted(lm)
    rows = np.concatenate([lm[n] for n in names])
    return lonlat_to_ecef(rows[:, 0], rows[:, 1]), rows, names


def rank_slice(n, rank, world_size):
    """Contiguous block of poses owned by `rank` (poses are independent: no collective)."""
    return (n * rank) // world_size, (n * (rank + 1)) // world_size


def visibility_sweep(poses, hfov=66.0, w_px=4608, h_px=2592, regions=None, rank=0, world_size=1, want_count=False,
                     device=None):
    """The reference's per-pose predicate `SatCam.check_for_all_landmarks()` (sim/nadir_sim.py:175,198) for all
    poses in ONE device call: poses (P,12) -> visible (P_rank,) bool [, landmark count].  With world_size > 1 the
    rank evaluates its contiguous slice `rank_slice(P, rank, world_size)`; callers concatenate (or all-gather)."""
    ctx = _lib.default_context(config.device if device is None else device)
    poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 12)
    lo, hi = rank_slice(poses.shape[0], rank, world_size)
    return ctx.satcam_visibility(landmark_table(ctx, regions), poses[lo:hi], hfov, w_px, h_px, want_count=want_count)


def inframe_sweep(poses, landmarks_ecef, hfov=66.0, w_px=4608, h_px=2592, chunk=4096, want_mask=False):
    """poses (P,12) x landmarks (L,3) -> per-pose count of landmarks that project inside the image (and optionally
    the (P,L) uint8 mask), computed on the device in pose chunks.  (Not a reference function: nadir_sim only
    projects the detections of visible frames.)"""
    ctx = _lib.default_context(config.device)
    poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 12)
    counts = np.zeros(poses.shape[0], dtype=np.int32)
    masks = []
    for p0 in range(0, poses.shape[0], chunk):
        _, m, c = ctx.satcam_project(poses[p0:p0 + chunk], landmarks_ecef, hfov, w_px, h_px, want_uv=False,
                                     want_mask=want_mask, want_count=True)
        counts[p0:p0 + chunk] = c
        if want_mask:
            masks.append(m)
    return (counts, np.concatenate(masks)) if want_mask else counts
