// a10: SatCam batched projection / visibility (sim/SatCam.py:39-92,125-154).
//
// Every arithmetic step is written with explicitly rounded operations in a fixed order (no FMA contraction)
// so that the in-frame decisions are bit-identical to the CPU oracle (oracle/satcam_oracle.py), which
// evaluates the same expressions in the same order in NumPy.
#include "common.cuh"
#include "launch.h"

using namespace vs;

namespace {

#define VS_TRY(expr)            \
  do {                          \
    int _rc = (expr);           \
    if (_rc != VINSAT_OK) return _rc; \
  } while (0)

struct CamIntr { double f, cx, cy; };

// C_cw = K [R_cw | -R_cw p] (SatCam.py:87-92), rows of R_cw = right, -up, dir (SatCam.py:52-56,81-84).
__global__ void __launch_bounds__(128) k_cam_matrix(int64_t P, const double* __restrict__ poses, CamIntr ci,
                                                    double* __restrict__ C) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= P) return;
  const double* s = poses + i * 12;
  const double px = s[0], py = s[1], pz = s[2];
  const double R[3][3] = {{s[9], s[10], s[11]}, {-s[6], -s[7], -s[8]}, {s[3], s[4], s[5]}};
  double t[3];
#pragma unroll
  for (int r = 0; r < 3; r++) t[r] = xadd(xadd(xmul(R[r][0], px), xmul(R[r][1], py)), xmul(R[r][2], pz));
  double* c = C + i * 12;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const double e0 = j < 3 ? R[0][j] : -t[0];
    const double e1 = j < 3 ? R[1][j] : -t[1];
    const double e2 = j < 3 ? R[2][j] : -t[2];
    c[0 * 4 + j] = xadd(xmul(ci.f, e0), xmul(ci.cx, e2));
    c[1 * 4 + j] = xadd(xmul(ci.f, e1), xmul(ci.cy, e2));
    c[2 * 4 + j] = e2;
  }
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
    const double a = xadd(xadd(xadd(xmul(c[0], x), xmul(c[1], y)), xmul(c[2], z)), c[3]);
    const double b = xadd(xadd(xadd(xmul(c[4], x), xmul(c[5], y)), xmul(c[6], z)), c[7]);
    const double w = xadd(xadd(xadd(xmul(c[8], x), xmul(c[9], y)), xmul(c[10], z)), c[11]);
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

// Ray / WGS84-ellipsoid intersection for the four image corners (SatCam.py:94-147).
__global__ void __launch_bounds__(128) k_satcam_corners(int64_t P, const double* __restrict__ poses, CamIntr ci,
                                                        double w_px, double h_px, double* __restrict__ corners,
                                                        uint8_t* __restrict__ hit) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t i = t >> 2;
  const int k = (int)(t & 3);
  if (i >= P) return;
  const double* s = poses + i * 12;
  const double x = s[0], y = s[1], z = s[2];
  // pixel of corner k: tl, tr, br, bl (SatCam.py:98-104)
  const double px = (k == 1 || k == 2) ? w_px : 0.0;
  const double py = (k >= 2) ? h_px : 0.0;
  // K^-1 [px,py,1] in closed form: ((px - cx)/f, (py - cy)/f, 1)
  const double kx = xsub(px, ci.cx) / ci.f, ky = xsub(py, ci.cy) / ci.f, kz = 1.0;
  // R_wc columns: right, -up, dir
  const double rw[3][3] = {{s[9], -s[6], s[3]}, {s[10], -s[7], s[4]}, {s[11], -s[8], s[5]}};
  double v3[3];
#pragma unroll
  for (int r = 0; r < 3; r++) v3[r] = xadd(xadd(xmul(rw[r][0], kx), xmul(rw[r][1], ky)), xmul(rw[r][2], kz));
  const double nrm = sqrt(xadd(xadd(xmul(v3[0], v3[0]), xmul(v3[1], v3[1])), xmul(v3[2], v3[2])));
  const double u = v3[0] / nrm, v = v3[1] / nrm, w = v3[2] / nrm;
  const double a = 6378137.0, b = 6378137.0, c = 6356752.314245;
  const double a2 = xmul(a, a), b2 = xmul(b, b), c2 = xmul(c, c);
  const double a2b2 = xmul(a2, b2), a2c2 = xmul(a2, c2), b2c2 = xmul(b2, c2);
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
  rad = xsub(rad, xmul(xmul(b2, u2), z2));
  rad = xadd(rad, xmul(xmul(xmul(xmul(xmul(2.0, b2), u), w), x), z));
  rad = xsub(rad, xmul(xmul(b2, w2), x2));
  rad = xsub(rad, xmul(xmul(c2, u2), y2));
  rad = xadd(rad, xmul(xmul(xmul(xmul(xmul(2.0, c2), u), v), x), y));
  rad = xsub(rad, xmul(xmul(c2, v2), x2));
  const double mag = xadd(xadd(xmul(a2b2, w2), xmul(a2c2, v2)), xmul(b2c2, u2));
  double* o = corners + (i * 4 + k) * 3;
  bool ok = !(rad < 0.0);
  double d = 0.0;
  if (ok) {
    d = xsub(value, xmul(xmul(xmul(a, b), c), sqrt(rad))) / mag;
    ok = !(d < 0.0);
  }
  hit[i * 4 + k] = ok ? 1 : 0;
  o[0] = ok ? xadd(x, xmul(d, u)) : 0.0;
  o[1] = ok ? xadd(y, xmul(d, v)) : 0.0;
  o[2] = ok ? xadd(z, xmul(d, w)) : 0.0;
}

CamIntr make_intr(double hfov_deg, int32_t w_px, int32_t h_px) {
  // SatCam.py:44-49: f = (w/2)/tan(deg2rad(hfov)/2)
  CamIntr ci;
  const double half_angle = (hfov_deg * (M_PI / 180.0)) / 2.0;
  ci.f = ((double)w_px / 2.0) / tan(half_angle);
  ci.cx = (double)w_px / 2.0;
  ci.cy = (double)h_px / 2.0;
  return ci;
}

}  // namespace

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

int vinsat_satcam_corners(vinsat_ctx* ctx, int mem, int64_t n_poses, const double* poses, double hfov_deg,
                          int32_t w_px, int32_t h_px, double* corners_out, uint8_t* hit_out) {
  VS_CHECK_ARG(ctx, ctx != nullptr);
  VS_CHECK_ARG(ctx, n_poses >= 0 && poses && corners_out && hit_out && w_px > 0 && h_px > 0);
  VS_CUDA(ctx, cudaSetDevice(ctx->device));
  if (n_poses == 0) return VINSAT_OK;
  const int64_t P = n_poses;
  const CamIntr ci = make_intr(hfov_deg, w_px, h_px);
  const bool host = mem != VINSAT_MEM_DEVICE;
  DevBuf<double> d_poses, d_c;
  DevBuf<uint8_t> d_h;
  const double* pp = poses;
  double* pc = corners_out;
  uint8_t* ph = hit_out;
  if (host) {
    VS_CUDA(ctx, d_poses.alloc(P * 12));
    VS_CUDA(ctx, d_c.alloc(P * 12));
    VS_CUDA(ctx, d_h.alloc(P * 4));
    VS_CUDA(ctx, cudaMemcpyAsync(d_poses.p, poses, P * 12 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    pp = d_poses.p; pc = d_c.p; ph = d_h.p;
  }
  VS_LAUNCH(ctx, F_SATCAM, k_satcam_corners, ceil_div(P * 4, 128), 128, 0, P, pp, ci, (double)w_px, (double)h_px, pc, ph);
  if (host) {
    VS_CUDA(ctx, cudaMemcpyAsync(corners_out, pc, P * 12 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    VS_CUDA(ctx, cudaMemcpyAsync(hit_out, ph, P * 4, cudaMemcpyDeviceToHost, ctx->stream));
  }
  VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return VINSAT_OK;
}

}  // extern "C"
