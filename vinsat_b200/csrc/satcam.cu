// a10: SatCam batched projection / visibility (sim/SatCam.py:39-92,125-154,175-262).
//
// Rounding order.  The reference does its small matrix products through NumPy -> OpenBLAS, which (on the host the
// goldens were generated on, tests/golden/make_golden_satcam.py) evaluates them as FMA chains in a fixed order:
//   gemm / Fortran-ordered gemv / dot : c = a0*b0; c = fma(a1,b1,c); c = fma(a2,b2,c)
//   C-ordered 3-column gemv           : c = a1*x1; c = fma(a0,x0,c); c = fma(a2,x2,c)
//   C-ordered 4-column gemv           : (a0*x0 + a2*x2) + (a1*x1 + a3*x3), products rounded separately
// and cast_ray_to_earth is scalar Python arithmetic, left to right.  The kernels below reproduce exactly that with
// explicit __fma_rn / __dmul_rn / __dadd_rn, so camera matrices, corner rays and pixel coordinates are bit-identical
// to the reference class (tests/test_gpu_satcam.py against tests/golden/satcam.npz).  The one exception is documented
// there: NumPy evaluates `w**2` with libm pow(), which glibc does not always round like w*w.
#include "common.cuh"
#include "launch.h"

using namespace vs;

namespace {

struct CamIntr { double f, cx, cy, k00, k02, k12; };   // k00 = K_inv[0,0] = K_inv[1,1], k02 = K_inv[0,2], k12 = K_inv[1,2]

__device__ __forceinline__ double xfma(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ double chain3(double a0, double b0, double a1, double b1, double a2, double b2) {
  return xfma(a2, b2, xfma(a1, b1, xmul(a0, b0)));
}

// C_cw = K_hom @ [R_cw | -R_cw p ; 0 0 0 1] (SatCam.py:87-92), rows of R_cw = right, -up, dir (:50-56,81-84).
__device__ __forceinline__ void cam_matrix(const double* __restrict__ s, const CamIntr& ci, double* c) {
  const double px = s[0], py = s[1], pz = s[2];
  const double R[3][3] = {{s[9], s[10], s[11]}, {-s[6], -s[7], -s[8]}, {s[3], s[4], s[5]}};
  double t[3];
#pragma unroll
  for (int r = 0; r < 3; r++) t[r] = chain3(R[r][0], px, R[r][1], py, R[r][2], pz);
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const double e0 = j < 3 ? R[0][j] : -t[0];
    const double e1 = j < 3 ? R[1][j] : -t[1];
    const double e2 = j < 3 ? R[2][j] : -t[2];
    c[0 * 4 + j] = xfma(ci.cx, e2, xmul(ci.f, e0));   // K_hom row [f 0 cx 0]: the zero terms leave the chain unchanged
    c[1 * 4 + j] = xfma(ci.cy, e2, xmul(ci.f, e1));
    c[2 * 4 + j] = e2;
  }
}

__global__ void __launch_bounds__(128) k_cam_matrix(int64_t P, const double* __restrict__ poses, CamIntr ci,
                                                    double* __restrict__ C) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= P) return;
  double c[12];
  cam_matrix(poses + i * 12, ci, c);
#pragma unroll
  for (int k = 0; k < 12; k++) C[i * 12 + k] = c[k];
}

constexpr int kPoseTile = 64;

// One thread per landmark, a tile of poses per CTA in shared memory; outputs are written with the
// landmark index fastest (coalesced).  uv = (C X)_{0:2} / (C X)_2 (SatCam.py:149-154).
__global__ void __launch_bounds__(256) k_satcam_project(int64_t P, int64_t L, const double* __restrict__ C,
                                                        const double* __restrict__ lm, double w_px, double h_px,
                                                        double* __restrict__ uv_out, uint8_t* __restrict__ inframe,
                                                        int32_t* __restrict__ count) {
  __shared__ double sC[kPoseTile * 12];
  const int64_t p0 = (int64_t)blockIdx.y * kPoseTile;
  const int np = (int)min((int64_t)kPoseTile, P - p0);
  for (int i = threadIdx.x; i < np * 12; i += blockDim.x) sC[i] = C[p0 * 12 + i];
  __syncthreads();
  const int64_t l = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const bool valid = l < L;
  double x = 0, y = 0, z = 0;
  if (valid) { x = lm[l * 3]; y = lm[l * 3 + 1]; z = lm[l * 3 + 2]; }
  for (int i = 0; i < np; i++) {
    const double* c = sC + i * 12;
    // C-ordered 4-column gemv with [x y z 1]
    const double a = xadd(xadd(xmul(c[0], x), xmul(c[2], z)), xadd(xmul(c[1], y), c[3]));
    const double b = xadd(xadd(xmul(c[4], x), xmul(c[6], z)), xadd(xmul(c[5], y), c[7]));
    const double w = xadd(xadd(xmul(c[8], x), xmul(c[10], z)), xadd(xmul(c[9], y), c[11]));
    const double u = a / w, v = b / w;
    const bool in = valid && w > 0.0 && u >= 0.0 && u < w_px && v >= 0.0 && v < h_px;
    const int64_t o = (p0 + i) * L + l;
    if (valid) {
      if (uv_out) { uv_out[o * 2] = u; uv_out[o * 2 + 1] = v; }
      if (inframe) inframe[o] = in ? 1 : 0;
    }
    if (count) {
      const unsigned m = __ballot_sync(0xffffffffu, in);
      if ((threadIdx.x & 31) == 0 && m) atomicAdd(&count[p0 + i], __popc(m));
    }
  }
}

// get_pixel_vector for corner k (tl, tr, br, bl; SatCam.py:94-104) and cast_ray_to_earth (:125-147).
// Returns hit; vec = unit ray, pt = intersection with the WGS84 ellipsoid (ECEF m).
__device__ __forceinline__ bool corner_cast(const double* __restrict__ s, const CamIntr& ci, int k, double w_px,
                                            double h_px, double* vec, double* pt) {
  const double x = s[0], y = s[1], z = s[2];
  const double px = (k == 1 || k == 2) ? w_px : 0.0;
  const double py = (k >= 2) ? h_px : 0.0;
  double v3[3];
#pragma unroll
  for (int r = 0; r < 3; r++) {
    // row r of R_wc = [right_r, -up_r, dir_r]; M = R_wc @ K_inv (gemm chain; K_inv = [[k00 0 k02],[0 k00 k12],[0 0 1]])
    const double rr = s[9 + r], ru = -s[6 + r], rd = s[3 + r];
    const double m0 = xmul(rr, ci.k00);                  // + fma(ru, 0, .) + fma(rd, 0, .)
    const double m1 = xfma(ru, ci.k00, xmul(rr, 0.0));
    const double m2 = xfma(rd, 1.0, xfma(ru, ci.k12, xmul(rr, ci.k02)));
    // vec = M @ [px, py, 1]: C-ordered 3-column gemv
    v3[r] = xfma(m2, 1.0, xfma(m0, px, xmul(m1, py)));
  }
  const double nrm = sqrt(chain3(v3[0], v3[0], v3[1], v3[1], v3[2], v3[2]));
  const double u = v3[0] / nrm, v = v3[1] / nrm, w = v3[2] / nrm;
  vec[0] = u; vec[1] = v; vec[2] = w;
  const double a = 6378137.0, c = 6356752.314245;
  const double a2 = xmul(a, a), c2 = xmul(c, c);
  const double a2b2 = xmul(a2, a2), a2c2 = xmul(a2, c2), b2c2 = a2c2;
  // value = -a^2 b^2 w z - a^2 c^2 v y - b^2 c^2 u x, evaluated left to right (SatCam.py:133)
  const double value = xsub(xsub(xmul(xmul(-a2b2, w), z), xmul(xmul(a2c2, v), y)), xmul(xmul(b2c2, u), x));
  const double w2 = xmul(w, w), v2 = xmul(v, v), u2 = xmul(u, u), x2 = xmul(x, x), y2 = xmul(y, y), z2 = xmul(z, z);
  // radical, term by term in the order of SatCam.py:134
  double rad = xmul(a2b2, w2);
  rad = xadd(rad, xmul(a2c2, v2));
  rad = xsub(rad, xmul(xmul(a2, v2), z2));
  rad = xadd(rad, xmul(xmul(xmul(xmul(xmul(2.0, a2), v), w), y), z));
  rad = xsub(rad, xmul(xmul(a2, w2), y2));
  rad = xadd(rad, xmul(b2c2, u2));
  rad = xsub(rad, xmul(xmul(a2, u2), z2));
  rad = xadd(rad, xmul(xmul(xmul(xmul(xmul(2.0, a2), u), w), x), z));
  rad = xsub(rad, xmul(xmul(a2, w2), x2));
  rad = xsub(rad, xmul(xmul(c2, u2), y2));
  rad = xadd(rad, xmul(xmul(xmul(xmul(xmul(2.0, c2), u), v), x), y));
  rad = xsub(rad, xmul(xmul(c2, v2), x2));
  const double mag = xadd(xadd(xmul(a2b2, w2), xmul(a2c2, v2)), xmul(b2c2, u2));
  bool ok = !(rad < 0.0);
  double d = 0.0;
  if (ok) {
    d = xsub(value, xmul(xmul(a2, c), sqrt(rad))) / mag;
    ok = !(d < 0.0);
  }
  pt[0] = ok ? xadd(x, xmul(d, u)) : 0.0;
  pt[1] = ok ? xadd(y, xmul(d, v)) : 0.0;
  pt[2] = ok ? xadd(z, xmul(d, w)) : 0.0;
  return ok;
}

__global__ void __launch_bounds__(128) k_satcam_corners(int64_t P, const double* __restrict__ poses, CamIntr ci,
                                                        double w_px, double h_px, double* __restrict__ corners,
                                                        uint8_t* __restrict__ hit, double* __restrict__ vecs) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t i = t >> 2;
  const int k = (int)(t & 3);
  if (i >= P) return;
  double vec[3], pt[3];
  const bool ok = corner_cast(poses + i * 12, ci, k, w_px, h_px, vec, pt);
  hit[i * 4 + k] = ok ? 1 : 0;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    corners[(i * 4 + k) * 3 + c] = pt[c];
    if (vecs) vecs[(i * 4 + k) * 3 + c] = vec[c];
  }
}

// ---- MGRS grid (sim/getMGRS.py:5-30) -------------------------------------------------------------------------
// Cell (row i = latitude band C..X, column j = zone 1..60) in the dict's insertion order (row-major).  The dict
// re-assigns the X row (72..84), 31V/32V and 31X/33X/35X/37X in place and deletes 32X/34X/36X.
__device__ __forceinline__ bool cell_bounds(int i, int j, double& lo0, double& la0, double& lo1, double& la1) {
  lo0 = -180.0 + 6.0 * j; lo1 = lo0 + 6.0;
  la0 = -80.0 + 8.0 * i;  la1 = la0 + 8.0;
  if (i == 19) {                       // X
    la1 = 84.0;
    if (j == 31 || j == 33 || j == 35) return false;          // 32X, 34X, 36X deleted
    if (j == 30) { lo0 = 0.0; lo1 = 9.0; }                    // 31X
    if (j == 32) { lo0 = 9.0; lo1 = 21.0; }                   // 33X
    if (j == 34) { lo0 = 21.0; lo1 = 33.0; }                  // 35X
    if (j == 36) { lo0 = 33.0; lo1 = 42.0; }                  // 37X
  } else if (i == 17) {                // V
    if (j == 30) { lo0 = 0.0; lo1 = 3.0; }                    // 31V
    if (j == 31) { lo0 = 3.0; lo1 = 12.0; }                   // 32V
  }
  return true;
}

// Band letter code (A=1 ... Z=26) of row i: labels "CDEFGHJKLMNPQRSTUVWX" skip I and O.
__device__ __forceinline__ int band_code(int i) { return i + 3 + (i >= 6) + (i >= 11); }

// get_region (SatCam.py:187-191): first cell, in dict order, whose INCLUSIVE bounds hold the point; -1 = None.
// Region code = zone*32 + letter code.  Only the cells around the point's own row / column can contain it.
__device__ __forceinline__ int get_region(double lon, double lat) {
  if (!(lon >= -180.0 && lon <= 180.0 && lat >= -80.0 && lat <= 84.0)) return -1;
  const int i0 = (int)floor((lat + 80.0) / 8.0), j0 = (int)floor((lon + 180.0) / 6.0);
  const int ia = max(0, i0 - 1), ib = min(19, i0), ja = max(0, j0 - 2), jb = min(59, j0 + 2);
  for (int i = ia; i <= ib; i++)
    for (int j = ja; j <= jb; j++) {
      double lo0, la0, lo1, la1;
      if (cell_bounds(i, j, lo0, la0, lo1, la1) && lo0 <= lon && lon <= lo1 && la0 <= lat && lat <= la1)
        return (j + 1) * 32 + band_code(i);
    }
  return -1;
}

// Geodetic lon/lat (degrees) of a point ON the ellipsoid: closed form standing in for astropy's
// EarthLocation.from_geocentric (SatCam.py:181; third party, absent from the reference checkout and the image).
__device__ __forceinline__ void lonlat_on_ellipsoid(const double* p, double& lon, double& lat) {
  const double a = 6378137.0, c = 6356752.314245;
  const double e2 = xsub(1.0, xmul(c, c) / xmul(a, a));
  const double k = 180.0 / 3.14159265358979323846;
  lon = xmul(atan2(p[1], p[0]), k);
  lat = xmul(atan2(p[2], xmul(xsub(1.0, e2), sqrt(xadd(xmul(p[0], p[0]), xmul(p[1], p[1]))))), k);
}

struct LmTable {
  const double* lon;          // [n] centroid longitude, region by region in CSV row order
  const double* lat;          // [n]
  const int32_t* off;         // [n_regions + 1]
  const int16_t* slot;        // [2048] region code -> table index, or -1 when the region is not in self.regions
};

// check_for_all_landmarks (SatCam.py:254-262) for every pose: corners -> lon/lat -> get_region ->
// find_current_regions (:203-230) -> box test of the region's centroids against the tl / br corners (:232-251).
// One lane per pose for the geometry; the box tests of a pose are then spread over the warp.
// visible[i] = 1 iff the reference would return True.  count[i] (optional) = landmarks counted without the
// reference's early exits (sum over the region sequence).  With no corner on the Earth the reference raises
// (unbound loop variable, :217): reported as not visible.
template <bool COUNT>
__global__ void __launch_bounds__(128) k_satcam_visibility(int64_t P, const double* __restrict__ poses, CamIntr ci,
                                                           double w_px, double h_px, LmTable tb,
                                                           uint8_t* __restrict__ visible, int32_t* __restrict__ count,
                                                           double* __restrict__ lonlat_out,
                                                           int32_t* __restrict__ region_out) {
  __shared__ int16_t s_slot[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) s_slot[i] = tb.slot[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const bool valid = i < P;
  double lon[4], lat[4];
  int reg[4];
  unsigned hit = 0;
  if (valid) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      double vec[3], pt[3];
      reg[k] = -1;
      lon[k] = lat[k] = __longlong_as_double(0x7ff8000000000000ll);
      if (corner_cast(poses + i * 12, ci, k, w_px, h_px, vec, pt)) {
        hit |= 1u << k;
        lonlat_on_ellipsoid(pt, lon[k], lat[k]);
        reg[k] = get_region(lon[k], lat[k]);
      }
      if (lonlat_out) { lonlat_out[(i * 4 + k) * 2] = lon[k]; lonlat_out[(i * 4 + k) * 2 + 1] = lat[k]; }
      if (region_out) region_out[i * 4 + k] = reg[k];
    }
  }
  // find_current_regions: bounds over the corners that hit and whose region is not None
  int nmin = 1000, nmax = -1, cmin = 1000, cmax = -1, last = -1;
#pragma unroll
  for (int k = 0; k < 4; k++)
    if (hit & (1u << k)) {
      last = reg[k];
      if (reg[k] >= 0) {
        nmin = min(nmin, reg[k] >> 5); nmax = max(nmax, reg[k] >> 5);
        cmin = min(cmin, reg[k] & 31); cmax = max(cmax, reg[k] & 31);
      }
    }
  // mode 0: nothing to test; 1: rectangle num_range x char_range (:218-227); 2: the corner regions themselves (:229)
  int mode = 0;
  const bool box_ok = (hit & 1u) && (hit & 4u);         // tl and br present (:237-238)
  if (valid && hit && box_ok) mode = (last >= 0) ? 1 : 2;
  const bool wrap = (nmin < 4 && nmax > 57);
  // does this pose touch any active region?  (cheap scan; most poses do not)
  bool any = false;
  if (mode == 1) {
    const int nn = wrap ? 6 : (nmax - nmin + 1);
    for (int a = 0; a < nn && !any; a++) {
      const int num = wrap ? (a < 3 ? 58 + a : a - 2) : nmin + a;
      for (int c = cmin; c <= cmax; c++) any |= s_slot[num * 32 + c] >= 0;
    }
  } else if (mode == 2) {
#pragma unroll
    for (int k = 0; k < 4; k++) any |= (hit & (1u << k)) && reg[k] >= 0 && s_slot[reg[k]] >= 0;
  }
  int total = 0;
  unsigned todo = __ballot_sync(0xffffffffu, any);
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    const double tl_lon = __shfl_sync(0xffffffffu, lon[0], src), tl_lat = __shfl_sync(0xffffffffu, lat[0], src);
    const double br_lon = __shfl_sync(0xffffffffu, lon[2], src), br_lat = __shfl_sync(0xffffffffu, lat[2], src);
    const int m = __shfl_sync(0xffffffffu, mode, src);
    const int s_nmin = __shfl_sync(0xffffffffu, nmin, src), s_nmax = __shfl_sync(0xffffffffu, nmax, src);
    const int s_cmin = __shfl_sync(0xffffffffu, cmin, src), s_cmax = __shfl_sync(0xffffffffu, cmax, src);
    const unsigned s_hit = __shfl_sync(0xffffffffu, hit, src);
    int s_reg[4];
#pragma unroll
    for (int k = 0; k < 4; k++) s_reg[k] = __shfl_sync(0xffffffffu, reg[k], src);
    const bool s_wrap = (s_nmin < 4 && s_nmax > 57);
    const int nn = (m == 1) ? (s_wrap ? 6 : (s_nmax - s_nmin + 1)) : 4;
    const int nc = (m == 1) ? (s_cmax - s_cmin + 1) : 1;
    int sum = 0;
    for (int a = 0; a < nn; a++) {
      for (int cc = 0; cc < nc; cc++) {
        int code;
        if (m == 1) {
          const int num = s_wrap ? (a < 3 ? 58 + a : a - 2) : s_nmin + a;
          code = num * 32 + s_cmin + cc;
        } else {
          code = (s_hit & (1u << a)) ? s_reg[a] : -1;
        }
        const int slot = code >= 0 ? s_slot[code] : -1;
        if (slot < 0) continue;
        const int b0 = tb.off[slot], b1 = tb.off[slot + 1];
        int cnt = 0;
        for (int j = b0 + lane; j < b1 + lane; j += 32) {
          bool in = false;
          if (j < b1) {
            const double clon = tb.lon[j], clat = tb.lat[j];
            in = clon > tl_lon && clon < br_lon && clat > br_lat && clat < tl_lat;     // strict (:247)
          }
          cnt += __popc(__ballot_sync(0xffffffffu, in));
          if (!COUNT && cnt >= 3) break;
        }
        sum += cnt;
        if (!COUNT && sum >= 3) { a = nn; break; }
      }
    }
    if (lane == src) total = sum;
  }
  if (valid) {
    visible[i] = total >= 3 ? 1 : 0;
    if (COUNT) count[i] = total;
  }
}

CamIntr make_intr(double hfov_deg, int32_t w_px, int32_t h_px) {
  // SatCam.py:44-49: f = (w/2)/tan(deg2rad(hfov)/2); K_inv = inv(K) (:62) of an upper-triangular K
  CamIntr ci;
  const double half_angle = (hfov_deg * (M_PI / 180.0)) / 2.0;
  ci.f = ((double)w_px / 2.0) / tan(half_angle);
  ci.cx = (double)w_px / 2.0;
  ci.cy = (double)h_px / 2.0;
  ci.k00 = 1.0 / ci.f;
  ci.k02 = -ci.cx / ci.f;
  ci.k12 = -ci.cy / ci.f;
  return ci;
}

}  // namespace

struct vinsat_satcam_table {
  vinsat_ctx* ctx = nullptr;
  double* lon = nullptr;
  double* lat = nullptr;
  int32_t* off = nullptr;
  int16_t* slot = nullptr;
  int64_t n = 0;
  int n_regions = 0;
};

extern "C" {

int vinsat_satcam_project(vinsat_ctx* ctx, int mem, int64_t n_poses, int64_t n_landmarks, const double* poses,
                          const double* landmarks_ecef, double hfov_deg, int32_t w_px, int32_t h_px, double* uv_out,
                          uint8_t* inframe_out, int32_t* count_out) {
  VS_CHECK_ARG(ctx, ctx != nullptr);
  VS_CHECK_ARG(ctx, n_poses >= 0 && n_landmarks >= 0 && poses && landmarks_ecef && w_px > 0 && h_px > 0);
  VS_CUDA(ctx, cudaSetDevice(ctx->device));
  if (n_poses == 0) return VINSAT_OK;
  const int64_t P = n_poses, L = n_landmarks;
  const CamIntr ci = make_intr(hfov_deg, w_px, h_px);
  const bool host = mem != VINSAT_MEM_DEVICE;
  DevBuf<double> d_poses, d_lm, d_C, d_uv;
  DevBuf<uint8_t> d_in;
  DevBuf<int32_t> d_cnt;
  const double* p_poses = poses;
  const double* p_lm = landmarks_ecef;
  double* p_uv = uv_out;
  uint8_t* p_in = inframe_out;
  int32_t* p_cnt = count_out;
  if (host) {
    VS_CUDA(ctx, d_poses.alloc(P * 12));
    VS_CUDA(ctx, d_lm.alloc(L * 3));
    VS_CUDA(ctx, cudaMemcpyAsync(d_poses.p, poses, P * 12 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    VS_CUDA(ctx, cudaMemcpyAsync(d_lm.p, landmarks_ecef, L * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    p_poses = d_poses.p; p_lm = d_lm.p;
    if (uv_out) { VS_CUDA(ctx, d_uv.alloc(P * L * 2)); p_uv = d_uv.p; }
    if (inframe_out) { VS_CUDA(ctx, d_in.alloc(P * L)); p_in = d_in.p; }
    if (count_out) { VS_CUDA(ctx, d_cnt.alloc(P)); p_cnt = d_cnt.p; }
  }
  VS_CUDA(ctx, d_C.alloc(P * 12));
  if (p_cnt) VS_CUDA(ctx, cudaMemsetAsync(p_cnt, 0, P * sizeof(int32_t), ctx->stream));
  VS_LAUNCH(ctx, F_SATCAM, k_cam_matrix, ceil_div(P, 128), 128, 0, P, p_poses, ci, d_C.p);
  if (L > 0) {
    const int64_t gy_max = 65535;
    for (int64_t pb = 0; pb < P; pb += gy_max * kPoseTile) {
      const int64_t pn = std::min<int64_t>(P - pb, gy_max * kPoseTile);
      dim3 grid((unsigned)ceil_div(L, 256), (unsigned)ceil_div(pn, kPoseTile));
      VS_LAUNCH(ctx, F_SATCAM, k_satcam_project, grid, 256, 0, pn, L, d_C.p + pb * 12, p_lm, (double)w_px,
                (double)h_px, p_uv ? p_uv + pb * L * 2 : nullptr, p_in ? p_in + pb * L : nullptr,
                p_cnt ? p_cnt + pb : nullptr);
    }
  }
  if (host) {
    if (uv_out) VS_CUDA(ctx, cudaMemcpyAsync(uv_out, p_uv, P * L * 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (inframe_out) VS_CUDA(ctx, cudaMemcpyAsync(inframe_out, p_in, P * L, cudaMemcpyDeviceToHost, ctx->stream));
    if (count_out) VS_CUDA(ctx, cudaMemcpyAsync(count_out, p_cnt, P * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  }
  VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return VINSAT_OK;
}

int vinsat_satcam_cam_matrix(vinsat_ctx* ctx, int mem, int64_t n_poses, const double* poses, double hfov_deg,
                             int32_t w_px, int32_t h_px, double* C_out) {
  VS_CHECK_ARG(ctx, ctx != nullptr);
  VS_CHECK_ARG(ctx, n_poses >= 0 && poses && C_out && w_px > 0 && h_px > 0);
  VS_CUDA(ctx, cudaSetDevice(ctx->device));
  if (n_poses == 0) return VINSAT_OK;
  const int64_t P = n_poses;
  const CamIntr ci = make_intr(hfov_deg, w_px, h_px);
  const bool host = mem != VINSAT_MEM_DEVICE;
  DevBuf<double> d_poses, d_C;
  const double* pp = poses;
  double* pc = C_out;
  if (host) {
    VS_CUDA(ctx, d_poses.alloc(P * 12));
    VS_CUDA(ctx, d_C.alloc(P * 12));
    VS_CUDA(ctx, cudaMemcpyAsync(d_poses.p, poses, P * 12 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    pp = d_poses.p; pc = d_C.p;
  }
  VS_LAUNCH(ctx, F_SATCAM, k_cam_matrix, ceil_div(P, 128), 128, 0, P, pp, ci, pc);
  if (host) VS_CUDA(ctx, cudaMemcpyAsync(C_out, pc, P * 12 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return VINSAT_OK;
}

int vinsat_satcam_corners(vinsat_ctx* ctx, int mem, int64_t n_poses, const double* poses, double hfov_deg,
                          int32_t w_px, int32_t h_px, double* corners_out, uint8_t* hit_out) {
  return vinsat_satcam_corner_rays(ctx, mem, n_poses, poses, hfov_deg, w_px, h_px, corners_out, hit_out, nullptr);
}

int vinsat_satcam_corner_rays(vinsat_ctx* ctx, int mem, int64_t n_poses, const double* poses, double hfov_deg,
                              int32_t w_px, int32_t h_px, double* corners_out, uint8_t* hit_out, double* vec_out) {
  VS_CHECK_ARG(ctx, ctx != nullptr);
  VS_CHECK_ARG(ctx, n_poses >= 0 && poses && corners_out && hit_out && w_px > 0 && h_px > 0);
  VS_CUDA(ctx, cudaSetDevice(ctx->device));
  if (n_poses == 0) return VINSAT_OK;
  const int64_t P = n_poses;
  const CamIntr ci = make_intr(hfov_deg, w_px, h_px);
  const bool host = mem != VINSAT_MEM_DEVICE;
  DevBuf<double> d_poses, d_c, d_v;
  DevBuf<uint8_t> d_h;
  const double* pp = poses;
  double* pc = corners_out;
  double* pv = vec_out;
  uint8_t* ph = hit_out;
  if (host) {
    VS_CUDA(ctx, d_poses.alloc(P * 12));
    VS_CUDA(ctx, d_c.alloc(P * 12));
    VS_CUDA(ctx, d_h.alloc(P * 4));
    if (vec_out) { VS_CUDA(ctx, d_v.alloc(P * 12)); pv = d_v.p; }
    VS_CUDA(ctx, cudaMemcpyAsync(d_poses.p, poses, P * 12 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    pp = d_poses.p; pc = d_c.p; ph = d_h.p;
  }
  VS_LAUNCH(ctx, F_SATCAM, k_satcam_corners, ceil_div(P * 4, 128), 128, 0, P, pp, ci, (double)w_px, (double)h_px, pc,
            ph, pv);
  if (host) {
    VS_CUDA(ctx, cudaMemcpyAsync(corners_out, pc, P * 12 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    VS_CUDA(ctx, cudaMemcpyAsync(hit_out, ph, P * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (vec_out) VS_CUDA(ctx, cudaMemcpyAsync(vec_out, pv, P * 12 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  }
  VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return VINSAT_OK;
}

int vinsat_satcam_table_create(vinsat_ctx* ctx, int32_t n_regions, const int32_t* region_codes,
                               const int64_t* region_off, const double* centroid_lonlat, int32_t n_active,
                               const int32_t* active_codes, vinsat_satcam_table** out) {
  VS_CHECK_ARG(ctx, ctx != nullptr);
  VS_CHECK_ARG(ctx, out && n_regions >= 0 && n_regions < 32767 && region_codes && region_off && n_active >= 0);
  VS_CHECK_ARG(ctx, n_regions == 0 || centroid_lonlat);
  VS_CHECK_ARG(ctx, n_active == 0 || active_codes);
  VS_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t n = region_off[n_regions];
  VS_CHECK_ARG(ctx, region_off[0] == 0 && n >= 0 && n < (int64_t)1 << 31);
  std::vector<int16_t> slot(2048, -1);
  std::vector<int32_t> off(n_regions + 1);
  for (int r = 0; r <= n_regions; r++) {
    if (r && region_off[r] < region_off[r - 1]) return set_error(ctx, VINSAT_EINVAL, "region_off must be non-decreasing");
    off[r] = (int32_t)region_off[r];
  }
  for (int r = 0; r < n_regions; r++) {
    const int code = region_codes[r];
    if (code < 0 || code >= 2048) return set_error(ctx, VINSAT_EINVAL, "region code %d out of range", code);
    bool act = false;
    for (int a = 0; a < n_active; a++) act |= active_codes[a] == code;
    if (act) slot[code] = (int16_t)r;       // a region listed twice keeps the later table (as a dict would)
  }
  std::vector<double> lon(n ? n : 1), lat(n ? n : 1);
  for (int64_t j = 0; j < n; j++) { lon[j] = centroid_lonlat[2 * j]; lat[j] = centroid_lonlat[2 * j + 1]; }
  vinsat_satcam_table* t = new vinsat_satcam_table();
  t->ctx = ctx; t->n = n; t->n_regions = n_regions;
  cudaError_t e = cudaMalloc((void**)&t->lon, (n ? n : 1) * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc((void**)&t->lat, (n ? n : 1) * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc((void**)&t->off, (n_regions + 1) * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc((void**)&t->slot, 2048 * sizeof(int16_t));
  if (e == cudaSuccess) e = cudaMemcpy(t->lon, lon.data(), n * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(t->lat, lat.data(), n * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(t->off, off.data(), (n_regions + 1) * sizeof(int32_t), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(t->slot, slot.data(), 2048 * sizeof(int16_t), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    vinsat_satcam_table_destroy(t);
    return set_error(ctx, VINSAT_ECUDA, "satcam table upload failed: %s", cudaGetErrorString(e));
  }
  *out = t;
  return VINSAT_OK;
}

int vinsat_satcam_table_destroy(vinsat_satcam_table* t) {
  if (!t) return VINSAT_OK;
  cudaSetDevice(t->ctx->device);
  cudaFree(t->lon); cudaFree(t->lat); cudaFree(t->off); cudaFree(t->slot);
  delete t;
  return VINSAT_OK;
}

int vinsat_satcam_visibility(vinsat_ctx* ctx, const vinsat_satcam_table* table, int mem, int64_t n_poses,
                             const double* poses, double hfov_deg, int32_t w_px, int32_t h_px, uint8_t* visible_out,
                             int32_t* count_out, double* corner_lonlat_out, int32_t* corner_region_out) {
  VS_CHECK_ARG(ctx, ctx != nullptr);
  VS_CHECK_ARG(ctx, table && table->ctx == ctx && n_poses >= 0 && poses && visible_out && w_px > 0 && h_px > 0);
  VS_CUDA(ctx, cudaSetDevice(ctx->device));
  if (n_poses == 0) return VINSAT_OK;
  const CamIntr ci = make_intr(hfov_deg, w_px, h_px);
  const LmTable tb = {table->lon, table->lat, table->off, table->slot};
  const bool host = mem != VINSAT_MEM_DEVICE;
  if (!host) {
    const int64_t P = n_poses;
    if (count_out)
      VS_LAUNCH(ctx, F_SATCAM, k_satcam_visibility<true>, ceil_div(P, 128), 128, 0, P, poses, ci, (double)w_px,
                (double)h_px, tb, visible_out, count_out, corner_lonlat_out, corner_region_out);
    else
      VS_LAUNCH(ctx, F_SATCAM, k_satcam_visibility<false>, ceil_div(P, 128), 128, 0, P, poses, ci, (double)w_px,
                (double)h_px, tb, visible_out, count_out, corner_lonlat_out, corner_region_out);
    return VINSAT_OK;                          // asynchronous on the context stream, like a kernel launch
  }
  // host buffers: staged chunk by chunk (bounded device memory); results are gathered on the device and copied once
  const int64_t chunk = 1 << 20;                // 1,048,576 poses = 100 MB per chunk
  const int64_t cn = std::min<int64_t>(chunk, n_poses);
  // all staging comes from the context's persistent scratch (one allocation that only grows): a sweep called per orbit
  // must not pay a 100 MB cudaMalloc / cudaFree pair every time
  const size_t b_poses = (size_t)cn * 12 * sizeof(double);
  const size_t b_ll = corner_lonlat_out ? (size_t)n_poses * 8 * sizeof(double) : 0;
  const size_t b_cnt = count_out ? (size_t)n_poses * sizeof(int32_t) : 0;
  const size_t b_reg = corner_region_out ? (size_t)n_poses * 4 * sizeof(int32_t) : 0;
  const size_t b_vis = ((size_t)n_poses + 15) & ~(size_t)15;
  char* sc = (char*)ctx_scratch(ctx, b_poses + b_ll + b_cnt + b_reg + b_vis + 64);
  if (!sc) return set_error(ctx, VINSAT_ENOMEM, "scratch allocation failed");
  struct { double* p; } d_poses{(double*)sc}, d_ll{(double*)(sc + b_poses)};
  struct { int32_t* p; } d_cnt{(int32_t*)(sc + b_poses + b_ll)}, d_reg{(int32_t*)(sc + b_poses + b_ll + b_cnt)};
  struct { uint8_t* p; } d_vis{(uint8_t*)(sc + b_poses + b_ll + b_cnt + b_reg)};
  int rc = VINSAT_OK;
  for (int64_t p0 = 0; p0 < n_poses && rc == VINSAT_OK; p0 += cn) {
    const int64_t pn = std::min<int64_t>(cn, n_poses - p0);
    double* dp = d_poses.p;
    cudaError_t e = cudaMemcpyAsync(dp, poses + p0 * 12, pn * 12 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) { rc = set_error(ctx, VINSAT_ECUDA, "H2D failed: %s", cudaGetErrorString(e)); break; }
    timing_begin(ctx, F_SATCAM);
    if (count_out)
      k_satcam_visibility<true><<<(unsigned)ceil_div(pn, 128), 128, 0, ctx->stream>>>(
          pn, dp, ci, (double)w_px, (double)h_px, tb, d_vis.p + p0, d_cnt.p + p0,
          corner_lonlat_out ? d_ll.p + p0 * 8 : nullptr, corner_region_out ? d_reg.p + p0 * 4 : nullptr);
    else
      k_satcam_visibility<false><<<(unsigned)ceil_div(pn, 128), 128, 0, ctx->stream>>>(
          pn, dp, ci, (double)w_px, (double)h_px, tb, d_vis.p + p0, nullptr,
          corner_lonlat_out ? d_ll.p + p0 * 8 : nullptr, corner_region_out ? d_reg.p + p0 * 4 : nullptr);
    timing_end(ctx);
    e = cudaGetLastError();
    if (e != cudaSuccess) rc = set_error(ctx, VINSAT_ECUDA, "launch k_satcam_visibility failed: %s", cudaGetErrorString(e));
  }
  if (rc == VINSAT_OK) {
    cudaError_t e = cudaMemcpyAsync(visible_out, d_vis.p, n_poses, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && count_out)
      e = cudaMemcpyAsync(count_out, d_cnt.p, n_poses * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && corner_lonlat_out)
      e = cudaMemcpyAsync(corner_lonlat_out, d_ll.p, n_poses * 8 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && corner_region_out)
      e = cudaMemcpyAsync(corner_region_out, d_reg.p, n_poses * 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = set_error(ctx, VINSAT_ECUDA, "visibility D2H failed: %s", cudaGetErrorString(e));
  } else {
    cudaStreamSynchronize(ctx->stream);
  }
  return rc;
}

}  // extern "C"
