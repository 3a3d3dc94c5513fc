// Observation-side kernels (a1, a5, a6 of SURVEY.md section 8): pinhole projection of landmarks, closed-form
// Jacobian, robust weights, and the segmented (per-frame) reduction of J^T W J / J^T W r.
//
// Layout: observations are SoA (X[3][M], uv[2][M], conf[M], oframe[M]) sorted by frame, so a warp reads
// 32 consecutive doubles per array (fully coalesced 256 B requests); per-frame data (state row 80 B,
// intrinsics 32 B) is AoS and is broadcast-read by the lanes that share a frame (L1 hits).
#include <cuda_pipeline.h>

#include <algorithm>

#include "common.cuh"
#include "launch.h"

namespace vs {

// ---------------------------------------------------------------------------------------------------------
// layout conversion through shared memory: both sides coalesced
// ---------------------------------------------------------------------------------------------------------
template <int NCOL>
__global__ void __launch_bounds__(256) k_aos_to_soa(const double* __restrict__ aos, double* __restrict__ soa, int64_t n) {
  __shared__ double tile[256 * NCOL];
  const int64_t row0 = (int64_t)blockIdx.x * 256;
  const int64_t rows = min((int64_t)256, n - row0);
  const double* src = aos + row0 * NCOL;
  for (int i = threadIdx.x; i < rows * NCOL; i += 256) tile[i] = src[i];
  __syncthreads();
  if (threadIdx.x < rows) {
#pragma unroll
    for (int c = 0; c < NCOL; c++) soa[(int64_t)c * n + row0 + threadIdx.x] = tile[threadIdx.x * NCOL + c];
  }
}

template <int NCOL>
__global__ void __launch_bounds__(256) k_soa_to_aos(const double* __restrict__ soa, double* __restrict__ aos, int64_t n) {
  __shared__ double tile[256 * NCOL];
  const int64_t row0 = (int64_t)blockIdx.x * 256;
  const int64_t rows = min((int64_t)256, n - row0);
  if (threadIdx.x < rows) {
#pragma unroll
    for (int c = 0; c < NCOL; c++) tile[threadIdx.x * NCOL + c] = soa[(int64_t)c * n + row0 + threadIdx.x];
  }
  __syncthreads();
  double* dst = aos + row0 * NCOL;
  for (int i = threadIdx.x; i < rows * NCOL; i += 256) dst[i] = tile[i];
}

int launch_aos_to_soa(vinsat_ctx* ctx, const double* aos, double* soa, int64_t n, int ncol) {
  if (n == 0) return VINSAT_OK;
  const int64_t grid = ceil_div(n, 256);
  switch (ncol) {
    case 1: VS_CUDA(ctx, cudaMemcpyAsync(soa, aos, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream)); break;
    case 2: VS_LAUNCH(ctx, F_LAYOUT, k_aos_to_soa<2>, grid, 256, 0, aos, soa, n); break;
    case 3: VS_LAUNCH(ctx, F_LAYOUT, k_aos_to_soa<3>, grid, 256, 0, aos, soa, n); break;
    case 6: VS_LAUNCH(ctx, F_LAYOUT, k_aos_to_soa<6>, grid, 256, 0, aos, soa, n); break;
    case 12: VS_LAUNCH(ctx, F_LAYOUT, k_aos_to_soa<12>, grid, 256, 0, aos, soa, n); break;
    default: return set_error(ctx, VINSAT_EINVAL, "aos_to_soa: unsupported ncol %d", ncol);
  }
  return VINSAT_OK;
}

int launch_soa_to_aos(vinsat_ctx* ctx, const double* soa, double* aos, int64_t n, int ncol) {
  if (n == 0) return VINSAT_OK;
  const int64_t grid = ceil_div(n, 256);
  switch (ncol) {
    case 1: VS_CUDA(ctx, cudaMemcpyAsync(aos, soa, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream)); break;
    case 2: VS_LAUNCH(ctx, F_LAYOUT, k_soa_to_aos<2>, grid, 256, 0, soa, aos, n); break;
    case 3: VS_LAUNCH(ctx, F_LAYOUT, k_soa_to_aos<3>, grid, 256, 0, soa, aos, n); break;
    case 6: VS_LAUNCH(ctx, F_LAYOUT, k_soa_to_aos<6>, grid, 256, 0, soa, aos, n); break;
    case 12: VS_LAUNCH(ctx, F_LAYOUT, k_soa_to_aos<12>, grid, 256, 0, soa, aos, n); break;
    default: return set_error(ctx, VINSAT_EINVAL, "soa_to_aos: unsupported ncol %d", ncol);
  }
  return VINSAT_OK;
}

// ---------------------------------------------------------------------------------------------------------
// observation indexing: global frame of each observation, CSR by frame, sortedness check
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int find_problem(const int64_t* __restrict__ off, int P, int64_t k) {
  int lo = 0, hi = P;   // off[lo] <= k < off[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (off[mid] <= k) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256) k_obs_index(int64_t M, int64_t T, int P, const int64_t* __restrict__ obs_off,
                                                   const int64_t* __restrict__ frame_off,
                                                   const int64_t* __restrict__ ii_local, int32_t* __restrict__ oframe,
                                                   int32_t* __restrict__ obs_start, int32_t* __restrict__ flags) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= M) return;
  const int p = find_problem(obs_off, P, k);
  const int64_t nf = frame_off[p + 1] - frame_off[p];
  const int64_t il = ii_local[k];
  if (il < 0 || il >= nf) { atomicOr(&flags[1], 1); return; }
  const int64_t f = frame_off[p] + il;
  oframe[k] = (int32_t)f;
  int64_t prev = -1;
  if (k > 0) {
    const int pp = find_problem(obs_off, P, k - 1);
    const int64_t ilp = ii_local[k - 1];
    // an out-of-range predecessor is flagged by its own thread; never derive a fill range from it
    if (ilp < 0 || ilp >= frame_off[pp + 1] - frame_off[pp]) return;
    prev = frame_off[pp] + ilp;
  }
  if (f < prev) { atomicOr(&flags[1], 2); return; }    // not sorted by frame
  for (int64_t j = prev + 1; j <= f; j++) obs_start[j] = (int32_t)k;
  if (k == M - 1)
    for (int64_t j = f + 1; j <= T; j++) obs_start[j] = (int32_t)M;
}

int launch_obs_index(vinsat_batch* b, const int64_t* d_ii_local) {
  vinsat_ctx* ctx = b->ctx;
  if (b->M == 0) {
    VS_CUDA(ctx, cudaMemsetAsync(b->obs_start, 0, (b->T + 1) * sizeof(int32_t), ctx->stream));
    return VINSAT_OK;
  }
  VS_LAUNCH(ctx, F_LAYOUT, k_obs_index, ceil_div(b->M, 256), 256, 0, b->M, b->T, (int)b->P, b->d_obs_off,
            b->d_frame_off, d_ii_local, b->oframe, b->obs_start, b->flags);
  return VINSAT_OK;
}

// ---------------------------------------------------------------------------------------------------------
// a1 stand-alone, reference (AoS) layout, arbitrary ii
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_project_aos(int64_t T, int64_t M, const double* __restrict__ states,
                                                     const double* __restrict__ intr, const double* __restrict__ xyz,
                                                     const int64_t* __restrict__ ii, double* __restrict__ uv_out,
                                                     double* __restrict__ Jg_out, int32_t* __restrict__ err) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= M) return;
  const int64_t f = ii[k];
  if (f < 0 || f >= T) { atomicOr(err, 1); return; }
  const double* s = states + f * 10;
  const double* c = intr + f * 4;
  Quat q = {s[3], s[4], s[5], s[6]};
  const double fx = c[0], fy = c[1];
  ProjOut o = project_exact(s[0], s[1], s[2], q, xyz[k * 3 + 0], xyz[k * 3 + 1], xyz[k * 3 + 2], fx, fy, c[2], c[3]);
  uv_out[k * 2 + 0] = o.u;
  uv_out[k * 2 + 1] = o.v;
  if (Jg_out) {
    double ju[6], jv[6];
    project_jacobian(o, fx, fy, ju, jv);
    double* J = Jg_out + k * 18;
#pragma unroll
    for (int a = 0; a < 6; a++) { J[a] = ju[a]; J[9 + a] = jv[a]; }
#pragma unroll
    for (int a = 6; a < 9; a++) { J[a] = 0.0; J[9 + a] = 0.0; }
  }
}

int launch_project_aos(vinsat_ctx* ctx, int64_t T, int64_t M, const double* states, const double* intr,
                       const double* xyz, const int64_t* ii, double* uv_out, double* Jg_out, int32_t* err_flag) {
  if (M == 0) return VINSAT_OK;
  VS_LAUNCH(ctx, F_PROJECT, k_project_aos, ceil_div(M, 256), 256, 0, T, M, states, intr, xyz, ii, uv_out, Jg_out,
            err_flag);
  return VINSAT_OK;
}

// ---------------------------------------------------------------------------------------------------------
// headline kernel: residual + Jacobian, SoA in / SoA out (unfused a1)
// Algorithmic bytes per observation: read X 24 + uv 16 + frame id 4, write r 16 + J (2x6 nonzeros) 96
// = 156 B, plus the frame's state row / intrinsics (88 B) amortised over its observations.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_resjac(int64_t M, const double* __restrict__ X, const double* __restrict__ uv,
                                                const int32_t* __restrict__ oframe, const double* __restrict__ st,
                                                const double* __restrict__ intr, double* __restrict__ r,
                                                double* __restrict__ J) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= M) return;
  const int f = oframe[k];
  const double x = X[k], y = X[M + k], z = X[2 * M + k];
  const double mu = uv[k], mv = uv[M + k];
  const double* s = st + (int64_t)f * 10;
  const double4 c = *reinterpret_cast<const double4*>(intr + (int64_t)f * 4);
  Quat q = {s[3], s[4], s[5], s[6]};
  ProjOut o = project_exact(s[0], s[1], s[2], q, x, y, z, c.x, c.y, c.z, c.w);
  double ju[6], jv[6];
  project_jacobian(o, c.x, c.y, ju, jv);
  r[k] = xsub(mu, o.u);
  r[M + k] = xsub(mv, o.v);
#pragma unroll
  for (int a = 0; a < 6; a++) {
    J[(int64_t)a * M + k] = ju[a];
    J[(int64_t)(6 + a) * M + k] = jv[a];
  }
}

int launch_resjac(vinsat_batch* b) {
  vinsat_ctx* ctx = b->ctx;
  if (b->M == 0) return VINSAT_OK;
  VS_LAUNCH(ctx, F_PROJECT, k_resjac, ceil_div(b->M, 256), 256, 0, b->M, b->X, b->uv, b->oframe, b->st, b->intr,
            b->r, b->J);
  return VINSAT_OK;
}

// r = uv - project(st): input of the robust scale c = median |r| (BA_filtering.py:21,23)
__global__ void __launch_bounds__(256) k_obs_residual(int64_t M, const double* __restrict__ X,
                                                      const double* __restrict__ uv,
                                                      const int32_t* __restrict__ oframe,
                                                      const double* __restrict__ st, const double* __restrict__ intr,
                                                      double* __restrict__ r) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= M) return;
  const int f = oframe[k];
  const double* s = st + (int64_t)f * 10;
  const double4 c = *reinterpret_cast<const double4*>(intr + (int64_t)f * 4);
  Quat q = {s[3], s[4], s[5], s[6]};
  ProjOut o = project_exact(s[0], s[1], s[2], q, X[k], X[M + k], X[2 * M + k], c.x, c.y, c.z, c.w);
  r[k] = xsub(uv[k], o.u);
  r[M + k] = xsub(uv[M + k], o.v);
}

int launch_obs_residual(vinsat_batch* b) {
  vinsat_ctx* ctx = b->ctx;
  if (b->M == 0) return VINSAT_OK;
  VS_LAUNCH(ctx, F_OBS_RESID, k_obs_residual, ceil_div(b->M, 256), 256, 0, b->M, b->X, b->uv, b->oframe, b->st,
            b->intr, b->r);
  return VINSAT_OK;
}

// ---------------------------------------------------------------------------------------------------------
// fused projection + Jacobian + robust weight + segmented reduction per frame (a1 + a5 + a6)
// A group of 8 lanes owns one frame and strides over its observations; the 28 partial sums are combined
// with xor-shuffles (fixed tree => deterministic) and written as one 224 B record.
// ---------------------------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ double group_sum(double v) {
  if (G >= 8) v += __shfl_xor_sync(0xffffffffu, v, 4);
  if (G >= 4) v += __shfl_xor_sync(0xffffffffu, v, 2);
  if (G >= 2) v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}
template <int G>
__device__ __forceinline__ double group_max(double v) {
  if (G >= 8) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, 4));
  if (G >= 4) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, 2));
  if (G >= 2) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return v;
}

struct WeightParams {
  double am2;       // |alpha - 2| (the reference divides by it)
  double ex;        // alpha/2 - 1
  int alpha_is_two;
  int ex_is_mhalf;  // alpha == 1 (every iteration >= 3): pow(x, -0.5) == 1/sqrt(x)
};

// BA_filtering.py:24, one component: ((r/c)^2/|alpha-2| + 1)^(alpha/2-1) / c^2
__device__ __forceinline__ double robust_component(double r, double c, const WeightParams& wp) {
  const double c2 = c * c;
  if (wp.alpha_is_two) return 1.0 / c2;     // pow(inf|nan, 0) == 1 (IEEE / torch), SURVEY section 7
  const double q = r / c;
  const double base = q * q / wp.am2 + 1.0;
  if (wp.ex_is_mhalf) return (1.0 / sqrt(base)) / c2;
  return pow(base, wp.ex) / c2;
}

// Same weight with the scale-dependent divisions hoisted out of the per-observation code (1/c, 1/|alpha-2|,
// 1/c^2 computed once per frame; rsqrt for alpha == 1): a few ulp from robust_component, well inside 1e-12.
struct FrameWeight { double inv_c, inv_am2, inv_c2; };
__device__ __forceinline__ FrameWeight frame_weight(double c, const WeightParams& wp) {
  FrameWeight fw;
  fw.inv_c = 1.0 / c;
  fw.inv_am2 = wp.alpha_is_two ? 0.0 : 1.0 / wp.am2;
  fw.inv_c2 = 1.0 / (c * c);
  return fw;
}
__device__ __forceinline__ double robust_component_fast(double r, const FrameWeight& fw, const WeightParams& wp) {
  if (wp.alpha_is_two) return fw.inv_c2;
  const double q = r * fw.inv_c;
  const double base = q * q * fw.inv_am2 + 1.0;
  if (wp.ex_is_mhalf) return rsqrt(base) * fw.inv_c2;
  return pow(base, wp.ex) * fw.inv_c2;
}

// ---------------------------------------------------------------------------------------------------------
// Camera-frame accumulation.  With J = [-Pi R^T | 2 Pi H], H = hat(p_c), N = w Pi^T Pi (3x3, 5 distinct non-zeros)
// and m = w Pi^T r, the per-frame sums are
//     J^T W J = [ R (sum N) R^T      -2 R (sum N H) ]        J^T W r = [ -R (sum m)     ]
//               [      .             4 sum H^T N H  ]                  [ 2 sum H^T m   ]
// so each observation only adds 26 camera-frame numbers (no per-observation rotation of the Jacobian) and the
// rotation into the world frame happens once per frame.
// ---------------------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------------------
// Same camera-frame accumulation, observations STAGED THROUGH SHARED MEMORY: a CTA owns 128 consecutive frames
// (one thread each); their observations form one contiguous range of the SoA arrays, which is copied in chunks
// with fully coalesced loads into a shared tile; every thread then walks its own frame's observations in the
// tile (LDS latency instead of L2/HBM latency per observation) and the un-normalised weights are written back
// through the same tile, coalesced.
// ---------------------------------------------------------------------------------------------------------
template <int kStageThreads, int kStageChunk, int kUnroll, int kMinBlocks>
__global__ void __launch_bounds__(kStageThreads, kMinBlocks) k_obs_assemble_staged(int64_t T, int64_t M,
                                                                         const int32_t* __restrict__ obs_start,
                                                                         const int32_t* __restrict__ fprob,
                                                                         const double* __restrict__ X,
                                                                         const double* __restrict__ uv,
                                                                         const double* __restrict__ conf,
                                                                         const double* __restrict__ st,
                                                                         const double* __restrict__ intr,
                                                                         const double* __restrict__ c_obs, WeightParams wp,
                                                                         double* __restrict__ wu_out,
                                                                         double* __restrict__ grec,
                                                                         unsigned long long* __restrict__ wmax) {
  extern __shared__ __align__(16) double tile[];      // [6][kStageChunk]: X0 X1 X2 u v conf(->w)
  const int tid = threadIdx.x;
  const int64_t f0 = (int64_t)blockIdx.x * kStageThreads;
  const int64_t f = f0 + tid;
  const bool valid = f < T;
  const int kb = obs_start[f0];
  const int ke = obs_start[min(f0 + kStageThreads, T)];
  // asynchronous staging (cp.async, no register round trip): every copy of the chunk is in flight at once and
  // overlaps the per-frame prologue below (ncu of the synchronous version: 2/3 of the stall samples sat on the
  // staging loop's load->store pairs and on the prologue's dependent loads)
  auto issue_chunk = [&](int base) {
    const int n = min(kStageChunk, ke - base);
    for (int i = tid; i < n; i += kStageThreads) {
      const int k = base + i;
      __pipeline_memcpy_async(&tile[i], &X[k], 8);
      __pipeline_memcpy_async(&tile[kStageChunk + i], &X[M + k], 8);
      __pipeline_memcpy_async(&tile[2 * kStageChunk + i], &X[2 * M + k], 8);
      __pipeline_memcpy_async(&tile[3 * kStageChunk + i], &uv[k], 8);
      __pipeline_memcpy_async(&tile[4 * kStageChunk + i], &uv[M + k], 8);
      __pipeline_memcpy_async(&tile[5 * kStageChunk + i], &conf[k], 8);
    }
    __pipeline_commit();
  };
  issue_chunk(kb);
  int k0 = 0, k1 = 0;
  double sN[5] = {0, 0, 0, 0, 0}, sNH[9], sHNH[6] = {0, 0, 0, 0, 0, 0}, sm[3] = {0, 0, 0}, sHm[3] = {0, 0, 0}, sabs = 0.0;
#pragma unroll
  for (int i = 0; i < 9; i++) sNH[i] = 0.0;
  double wloc = 0.0, c = 1.0, px = 0, py = 0, pz = 0;
  FrameWeight fw = {1.0, 1.0, 1.0};
  double4 ci = make_double4(0, 0, 0, 0);
  int p = 0;
  double Rt[9];
#pragma unroll
  for (int i = 0; i < 9; i++) Rt[i] = 0.0;
  if (valid) {
    k0 = obs_start[f]; k1 = obs_start[f + 1];
    if (k1 > k0) {
      p = fprob[f];
      c = c_obs[p];
      fw = frame_weight(c, wp);
      const double* s = st + f * 10;
      ci = *reinterpret_cast<const double4*>(intr + f * 4);
      px = s[0]; py = s[1]; pz = s[2];
      const double qn = 1.0 / sqrt(s[3] * s[3] + s[4] * s[4] + s[5] * s[5] + s[6] * s[6]);
      const double x = s[3] * qn, y = s[4] * qn, z = s[5] * qn, w = s[6] * qn;
      Rt[0] = 1 - 2 * (y * y + z * z); Rt[1] = 2 * (x * y + z * w);     Rt[2] = 2 * (x * z - y * w);
      Rt[3] = 2 * (x * y - z * w);     Rt[4] = 1 - 2 * (x * x + z * z); Rt[5] = 2 * (y * z + x * w);
      Rt[6] = 2 * (x * z + y * w);     Rt[7] = 2 * (y * z - x * w);     Rt[8] = 1 - 2 * (x * x + y * y);
    }
  }
  for (int base = kb; base < ke; base += kStageChunk) {
    const int n = min(kStageChunk, ke - base);
    __pipeline_wait_prior(0);
    __syncthreads();
    const int lo = max(k0, base), hi = min(k1, base + n);
#pragma unroll kUnroll
    for (int k = lo; k < hi; k++) {
      const int i = k - base;
      const double dx = tile[i] - px, dy = tile[kStageChunk + i] - py, dz = tile[2 * kStageChunk + i] - pz;
      const double Xc = Rt[0] * dx + Rt[1] * dy + Rt[2] * dz;
      const double Yc = Rt[3] * dx + Rt[4] * dy + Rt[5] * dz;
      const double Zc = Rt[6] * dx + Rt[7] * dy + Rt[8] * dz;
      const double live = (Zc >= 0.1) ? 1.0 : 0.0;
      const double d = 1.0 / fmax(Zc, 0.1);
      const double a = ci.x * d, bb = ci.y * d;
      const double ru = tile[3 * kStageChunk + i] - (a * Xc + ci.z), rv = tile[4 * kStageChunk + i] - (bb * Yc + ci.w);
      const double cc = -a * Xc * d * live, ee = -bb * Yc * d * live;
      const double wraw = 0.5 * (robust_component_fast(ru, fw, wp) + robust_component_fast(rv, fw, wp));
      const double w = wraw * tile[5 * kStageChunk + i];
      tile[5 * kStageChunk + i] = w;
      wloc = fmax(wloc, wraw);
      sabs += fabs(ru) + fabs(rv);
      const double wa = w * a, wb = w * bb;
      const double n00 = wa * a, n02 = wa * cc, n11 = wb * bb, n12 = wb * ee, n22 = w * (cc * cc + ee * ee);
      sN[0] += n00; sN[1] += n02; sN[2] += n11; sN[3] += n12; sN[4] += n22;
      const double h00 = -n02 * Yc, h01 = n02 * Xc - n00 * Zc, h02 = n00 * Yc;
      const double h10 = n11 * Zc - n12 * Yc, h11 = n12 * Xc, h12 = -n11 * Xc;
      const double h20 = n12 * Zc - n22 * Yc, h21 = n22 * Xc - n02 * Zc, h22 = n02 * Yc - n12 * Xc;
      sNH[0] += h00; sNH[1] += h01; sNH[2] += h02; sNH[3] += h10; sNH[4] += h11; sNH[5] += h12;
      sNH[6] += h20; sNH[7] += h21; sNH[8] += h22;
      sHNH[0] += Zc * h10 - Yc * h20; sHNH[1] += Zc * h11 - Yc * h21; sHNH[2] += Zc * h12 - Yc * h22;
      sHNH[3] += Xc * h21 - Zc * h01; sHNH[4] += Xc * h22 - Zc * h02; sHNH[5] += Yc * h02 - Xc * h12;
      const double m0 = wa * ru, m1 = wb * rv, m2 = w * (cc * ru + ee * rv);
      sm[0] += m0; sm[1] += m1; sm[2] += m2;
      sHm[0] += Zc * m1 - Yc * m2; sHm[1] += Xc * m2 - Zc * m0; sHm[2] += Yc * m0 - Xc * m1;
    }
    __syncthreads();
    for (int i = tid; i < n; i += kStageThreads) wu_out[base + i] = tile[5 * kStageChunk + i];
    __syncthreads();
    if (base + kStageChunk < ke) issue_chunk(base + kStageChunk);
  }
  if (valid) {
    const double N[9] = {sN[0], 0.0, sN[1], 0.0, sN[2], sN[3], sN[1], sN[3], sN[4]};
    double NR[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) NR[i * 3 + j] = N[i * 3] * Rt[j] + N[i * 3 + 1] * Rt[3 + j] + N[i * 3 + 2] * Rt[6 + j];
    double out[VS_GREC];
    int idx = 0;
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
      for (int bcol = a; bcol < 3; bcol++)
        out[idx++] = Rt[a] * NR[bcol] + Rt[3 + a] * NR[3 + bcol] + Rt[6 + a] * NR[6 + bcol];
#pragma unroll
      for (int bcol = 0; bcol < 3; bcol++)
        out[idx++] = -2.0 * (Rt[a] * sNH[bcol] + Rt[3 + a] * sNH[3 + bcol] + Rt[6 + a] * sNH[6 + bcol]);
    }
    out[idx++] = 4.0 * sHNH[0]; out[idx++] = 4.0 * sHNH[1]; out[idx++] = 4.0 * sHNH[2];
    out[idx++] = 4.0 * sHNH[3]; out[idx++] = 4.0 * sHNH[4];
    out[idx++] = 4.0 * sHNH[5];
#pragma unroll
    for (int a = 0; a < 3; a++) out[21 + a] = -(Rt[a] * sm[0] + Rt[3 + a] * sm[1] + Rt[6 + a] * sm[2]);
#pragma unroll
    for (int a = 0; a < 3; a++) out[24 + a] = 2.0 * sHm[a];
    out[27] = sabs;
    double* g = grec + f * VS_GREC;
#pragma unroll
    for (int i = 0; i < VS_GREC; i += 2) *reinterpret_cast<double2*>(g + i) = make_double2(out[i], out[i + 1]);
    if (wloc > 0.0) atomicMax(&wmax[p], (unsigned long long)__double_as_longlong(wloc));
  }
}

// ---------------------------------------------------------------------------------------------------------
// Two-pass variant (default).  The walk above is one long dependent chain per observation (division, two rsqrt,
// ~40 dependent FP64 operations) repeated sequentially for a frame's observations: time = observations x chain
// latency / resident threads, and the 27 accumulators cap the resident threads at 12 warps per SM.  Here ONE WARP
// owns 32 consecutive frames and splits the work by what is parallel in it:
//   pass 1  one lane per OBSERVATION of the tile (independent iterations, no loop-carried state): camera-frame
//           point, reciprocal depth, residual, robust weight -> written back over the staged inputs in shared memory
//   pass 2  one lane per FRAME: only the products and the 27 accumulations remain (dependency depth ~6, 27
//           independent chains), reading the per-observation basics from shared memory
// The per-frame constants pass 1 needs (rotation, position, intrinsics, 1/c, 1/c^2) go through a small shared table.
// ---------------------------------------------------------------------------------------------------------
constexpr int k2pChunk = 320;                 // observations per chunk (32 frames x 10; more obs => more chunks)
constexpr int k2pSlots = 7;                   // Xc Yc Zc | ru rv | w | d    (staged as X0 X1 X2 | u v | conf | -)
constexpr int k2pFrameRec = 19;               // Rt 9 | p 3 | fx fy cx cy | 1/c | 1/c^2 | (pad: odd pitch)

__global__ void __launch_bounds__(32, 9) k_obs_assemble_2pass(int64_t T, int64_t M, const int32_t* __restrict__ obs_start,
                                                              const int32_t* __restrict__ oframe,
                                                              const int32_t* __restrict__ fprob,
                                                              const double* __restrict__ X, const double* __restrict__ uv,
                                                              const double* __restrict__ conf,
                                                              const double* __restrict__ st,
                                                              const double* __restrict__ intr,
                                                              const double* __restrict__ c_obs, WeightParams wp,
                                                              double* __restrict__ wu_out, double* __restrict__ grec,
                                                              unsigned long long* __restrict__ wmax,
                                                              const int32_t* __restrict__ gate) {
  if (gate && *gate) return;      // speculative launch behind an LM loop that is not finished (batch.cu)
  extern __shared__ __align__(16) double sm2[];
  double* tile = sm2;                                       // [k2pSlots][k2pChunk]
  double* fdat = sm2 + k2pSlots * k2pChunk;                 // [32][k2pFrameRec]
  int32_t* ofr = reinterpret_cast<int32_t*>(fdat + 32 * k2pFrameRec);     // [k2pChunk] frame of each staged observation
  const int lane = threadIdx.x;
  const int64_t f0 = (int64_t)blockIdx.x * 32;
  const int64_t f = f0 + lane;
  const bool valid = f < T;
  const int kb = obs_start[f0];
  const int ke = obs_start[min(f0 + 32, T)];
  auto issue_chunk = [&](int base) {
    const int n = min(k2pChunk, ke - base);
    for (int i = lane; i < n; i += 32) {
      const int k = base + i;
      __pipeline_memcpy_async(&tile[i], &X[k], 8);
      __pipeline_memcpy_async(&tile[k2pChunk + i], &X[M + k], 8);
      __pipeline_memcpy_async(&tile[2 * k2pChunk + i], &X[2 * M + k], 8);
      __pipeline_memcpy_async(&tile[3 * k2pChunk + i], &uv[k], 8);
      __pipeline_memcpy_async(&tile[4 * k2pChunk + i], &uv[M + k], 8);
      __pipeline_memcpy_async(&tile[5 * k2pChunk + i], &conf[k], 8);
      __pipeline_memcpy_async(&ofr[i], &oframe[k], 4);
    }
    __pipeline_commit();
  };
  issue_chunk(kb);
  int k0 = 0, k1 = 0, p = 0;
  double Rt[9];
#pragma unroll
  for (int i = 0; i < 9; i++) Rt[i] = 0.0;
  double4 ci = make_double4(0, 0, 0, 0);
  {
    double* fd = fdat + lane * k2pFrameRec;
    double px = 0, py = 0, pz = 0;
    FrameWeight fw = {1.0, 1.0, 1.0};
    if (valid) {
      k0 = obs_start[f]; k1 = obs_start[f + 1];
      p = fprob[f];
      if (k1 > k0) {
        fw = frame_weight(c_obs[p], wp);
        const double* s = st + f * 10;
        ci = *reinterpret_cast<const double4*>(intr + f * 4);
        px = s[0]; py = s[1]; pz = s[2];
        const double qn = 1.0 / sqrt(s[3] * s[3] + s[4] * s[4] + s[5] * s[5] + s[6] * s[6]);
        const double x = s[3] * qn, y = s[4] * qn, z = s[5] * qn, w = s[6] * qn;
        Rt[0] = 1 - 2 * (y * y + z * z); Rt[1] = 2 * (x * y + z * w);     Rt[2] = 2 * (x * z - y * w);
        Rt[3] = 2 * (x * y - z * w);     Rt[4] = 1 - 2 * (x * x + z * z); Rt[5] = 2 * (y * z + x * w);
        Rt[6] = 2 * (x * z + y * w);     Rt[7] = 2 * (y * z - x * w);     Rt[8] = 1 - 2 * (x * x + y * y);
      }
    }
#pragma unroll
    for (int i = 0; i < 9; i++) fd[i] = Rt[i];
    fd[9] = px; fd[10] = py; fd[11] = pz;
    fd[12] = ci.x; fd[13] = ci.y; fd[14] = ci.z; fd[15] = ci.w;
    fd[16] = fw.inv_c; fd[17] = fw.inv_c2;
  }
  // the tile's frames usually belong to one problem: then the largest raw weight is reduced in registers
  const int p_first = __shfl_sync(0xffffffffu, p, 0);
  const bool one_problem = __all_sync(0xffffffffu, !valid || p == p_first);
  const double inv_am2 = wp.alpha_is_two ? 0.0 : 1.0 / wp.am2;
  double sN[5] = {0, 0, 0, 0, 0}, sNH[9], sHNH[6] = {0, 0, 0, 0, 0, 0}, sm[3] = {0, 0, 0}, sHm[3] = {0, 0, 0}, sabs = 0.0;
#pragma unroll
  for (int i = 0; i < 9; i++) sNH[i] = 0.0;
  double wloc = 0.0;
  for (int base = kb; base < ke; base += k2pChunk) {
    const int n = min(k2pChunk, ke - base);
    __pipeline_wait_prior(0);
    __syncwarp();
    // ---- pass 1: one lane per observation
#pragma unroll 2
    for (int i = lane; i < n; i += 32) {
      const int lf = ofr[i] - (int)f0;
      const double* fd = fdat + lf * k2pFrameRec;
      const double dx = tile[i] - fd[9], dy = tile[k2pChunk + i] - fd[10], dz = tile[2 * k2pChunk + i] - fd[11];
      const double Xc = fd[0] * dx + fd[1] * dy + fd[2] * dz;
      const double Yc = fd[3] * dx + fd[4] * dy + fd[5] * dz;
      const double Zc = fd[6] * dx + fd[7] * dy + fd[8] * dz;
      const double d = 1.0 / fmax(Zc, 0.1);
      const double a = fd[12] * d, bb = fd[13] * d;
      const double ru = tile[3 * k2pChunk + i] - (a * Xc + fd[14]), rv = tile[4 * k2pChunk + i] - (bb * Yc + fd[15]);
      FrameWeight fw; fw.inv_c = fd[16]; fw.inv_am2 = inv_am2; fw.inv_c2 = fd[17];
      const double wraw = 0.5 * (robust_component_fast(ru, fw, wp) + robust_component_fast(rv, fw, wp));
      const double w = wraw * tile[5 * k2pChunk + i];
      tile[i] = Xc; tile[k2pChunk + i] = Yc; tile[2 * k2pChunk + i] = Zc;
      tile[3 * k2pChunk + i] = ru; tile[4 * k2pChunk + i] = rv;
      tile[5 * k2pChunk + i] = w;
      tile[6 * k2pChunk + i] = d;
      if (one_problem) wloc = fmax(wloc, wraw);
      else if (wraw > 0.0) atomicMax(&wmax[fprob[f0 + lf]], (unsigned long long)__double_as_longlong(wraw));
    }
    __syncwarp();
    // ---- pass 2: one lane per frame
    const int lo = max(k0, base), hi = min(k1, base + n);
    for (int k = lo; k < hi; k++) {
      const int i = k - base;
      const double Xc = tile[i], Yc = tile[k2pChunk + i], Zc = tile[2 * k2pChunk + i];
      const double ru = tile[3 * k2pChunk + i], rv = tile[4 * k2pChunk + i];
      const double w = tile[5 * k2pChunk + i], d = tile[6 * k2pChunk + i];
      const double live = (Zc >= 0.1) ? 1.0 : 0.0;
      const double a = ci.x * d, bb = ci.y * d;
      const double cc = -a * Xc * d * live, ee = -bb * Yc * d * live;
      sabs += fabs(ru) + fabs(rv);
      const double wa = w * a, wb = w * bb;
      const double n00 = wa * a, n02 = wa * cc, n11 = wb * bb, n12 = wb * ee, n22 = w * (cc * cc + ee * ee);
      sN[0] += n00; sN[1] += n02; sN[2] += n11; sN[3] += n12; sN[4] += n22;
      const double h00 = -n02 * Yc, h01 = n02 * Xc - n00 * Zc, h02 = n00 * Yc;
      const double h10 = n11 * Zc - n12 * Yc, h11 = n12 * Xc, h12 = -n11 * Xc;
      const double h20 = n12 * Zc - n22 * Yc, h21 = n22 * Xc - n02 * Zc, h22 = n02 * Yc - n12 * Xc;
      sNH[0] += h00; sNH[1] += h01; sNH[2] += h02; sNH[3] += h10; sNH[4] += h11; sNH[5] += h12;
      sNH[6] += h20; sNH[7] += h21; sNH[8] += h22;
      sHNH[0] += Zc * h10 - Yc * h20; sHNH[1] += Zc * h11 - Yc * h21; sHNH[2] += Zc * h12 - Yc * h22;
      sHNH[3] += Xc * h21 - Zc * h01; sHNH[4] += Xc * h22 - Zc * h02; sHNH[5] += Yc * h02 - Xc * h12;
      const double m0 = wa * ru, m1 = wb * rv, m2 = w * (cc * ru + ee * rv);
      sm[0] += m0; sm[1] += m1; sm[2] += m2;
      sHm[0] += Zc * m1 - Yc * m2; sHm[1] += Xc * m2 - Zc * m0; sHm[2] += Yc * m0 - Xc * m1;
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) wu_out[base + i] = tile[5 * k2pChunk + i];
    __syncwarp();
    if (base + k2pChunk < ke) issue_chunk(base + k2pChunk);
  }
  if (one_problem) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wloc = fmax(wloc, __shfl_xor_sync(0xffffffffu, wloc, o));
    if (lane == 0 && wloc > 0.0) atomicMax(&wmax[p_first], (unsigned long long)__double_as_longlong(wloc));
  }
  if (valid) {
    const double N[9] = {sN[0], 0.0, sN[1], 0.0, sN[2], sN[3], sN[1], sN[3], sN[4]};
    double NR[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) NR[i * 3 + j] = N[i * 3] * Rt[j] + N[i * 3 + 1] * Rt[3 + j] + N[i * 3 + 2] * Rt[6 + j];
    double out[VS_GREC];
    int idx = 0;
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
      for (int bcol = a; bcol < 3; bcol++)
        out[idx++] = Rt[a] * NR[bcol] + Rt[3 + a] * NR[3 + bcol] + Rt[6 + a] * NR[6 + bcol];
#pragma unroll
      for (int bcol = 0; bcol < 3; bcol++)
        out[idx++] = -2.0 * (Rt[a] * sNH[bcol] + Rt[3 + a] * sNH[3 + bcol] + Rt[6 + a] * sNH[6 + bcol]);
    }
    out[idx++] = 4.0 * sHNH[0]; out[idx++] = 4.0 * sHNH[1]; out[idx++] = 4.0 * sHNH[2];
    out[idx++] = 4.0 * sHNH[3]; out[idx++] = 4.0 * sHNH[4];
    out[idx++] = 4.0 * sHNH[5];
#pragma unroll
    for (int a = 0; a < 3; a++) out[21 + a] = -(Rt[a] * sm[0] + Rt[3 + a] * sm[1] + Rt[6 + a] * sm[2]);
#pragma unroll
    for (int a = 0; a < 3; a++) out[24 + a] = 2.0 * sHm[a];
    out[27] = sabs;
    double* g = grec + f * VS_GREC;
#pragma unroll
    for (int i = 0; i < VS_GREC; i += 2) *reinterpret_cast<double2*>(g + i) = make_double2(out[i], out[i + 1]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Persistent, double-buffered variant of the two-pass kernel (default).  ncu on the two-pass kernel (r01/r02): 13.7 %
// of the warp slots, FP64 pipe 27 % busy, a CTA lives ~25 k cycles of which ~4 k are FP64 issue: the rest is the
// exposed latency of a CTA that first learns its observation range, then waits for its tile, then for the frame
// constants.  Here one warp per scheduler stays resident (4 one-warp CTAs per SM, 51 KB of shared memory each) and
// walks tiles of 32 frames: while tile i is processed the observation tile AND the frame rows (states, intrinsics,
// CSR offsets, problem ids) of tile i+1 arrive by cp.async, and the observation range of tile i+2 is already being
// read.  Math and accumulation order are those of the two-pass kernel (bit-identical records).
// ---------------------------------------------------------------------------------------------------------
constexpr int kPaSlotBytes = k2pChunk * 8;
constexpr int kPaInBytes = 7 * kPaSlotBytes + k2pChunk * 4;                // 7 double slots + frame ids
constexpr int kPaFinBytes = 32 * 10 * 8 + 32 * 4 * 8 + 40 * 4 + 32 * 4;    // states | intrinsics | obs_start[33] (padded) | fprob
constexpr int kPaSmemBytes = 2 * kPaInBytes + 2 * kPaFinBytes + 32 * k2pFrameRec * 8;

__global__ void __launch_bounds__(32, 4) k_obs_assemble_pers(int64_t T, int64_t M, const int32_t* __restrict__ obs_start,
                                                             const int32_t* __restrict__ oframe,
                                                             const int32_t* __restrict__ fprob,
                                                             const double* __restrict__ X, const double* __restrict__ uv,
                                                             const double* __restrict__ conf,
                                                             const double* __restrict__ st,
                                                             const double* __restrict__ intr,
                                                             const double* __restrict__ c_obs, WeightParams wp,
                                                             double* __restrict__ wu_out, double* __restrict__ grec,
                                                             unsigned long long* __restrict__ wmax,
                                                             const int32_t* __restrict__ gate) {
  if (gate && *gate) return;
  extern __shared__ __align__(16) unsigned char smp[];
  const int lane = threadIdx.x;
  const int64_t n_tiles = (T + 31) / 32;
  auto in_tile = [&](int b) { return reinterpret_cast<double*>(smp + b * kPaInBytes); };
  auto in_ofr = [&](int b) { return reinterpret_cast<int32_t*>(smp + b * kPaInBytes + 7 * kPaSlotBytes); };
  auto fin_st = [&](int b) { return reinterpret_cast<double*>(smp + 2 * kPaInBytes + b * kPaFinBytes); };
  auto fin_intr = [&](int b) { return fin_st(b) + 320; };
  auto fin_os = [&](int b) { return reinterpret_cast<int32_t*>(fin_intr(b) + 128); };
  auto fin_fp = [&](int b) { return fin_os(b) + 40; };
  double* fdat = reinterpret_cast<double*>(smp + 2 * kPaInBytes + 2 * kPaFinBytes);

  auto issue_obs = [&](double* tile, int32_t* ofr, int base, int n) {
    for (int i = lane; i < n; i += 32) {
      const int k = base + i;
      __pipeline_memcpy_async(&tile[i], &X[k], 8);
      __pipeline_memcpy_async(&tile[k2pChunk + i], &X[M + k], 8);
      __pipeline_memcpy_async(&tile[2 * k2pChunk + i], &X[2 * M + k], 8);
      __pipeline_memcpy_async(&tile[3 * k2pChunk + i], &uv[k], 8);
      __pipeline_memcpy_async(&tile[4 * k2pChunk + i], &uv[M + k], 8);
      __pipeline_memcpy_async(&tile[5 * k2pChunk + i], &conf[k], 8);
      __pipeline_memcpy_async(&ofr[i], &oframe[k], 4);
    }
  };
  // frame rows of a tile: 32 x 80 B of states and 32 x 32 B of intrinsics are contiguous -> 16-byte copies
  auto issue_frames = [&](int b, int64_t f0) {
    const int nf = (int)min((int64_t)32, T - f0);
    const double* s0 = st + f0 * 10;
    const double* i0 = intr + f0 * 4;
    for (int i = lane; i < nf * 5; i += 32) __pipeline_memcpy_async(fin_st(b) + 2 * i, s0 + 2 * i, 16);
    for (int i = lane; i < nf * 2; i += 32) __pipeline_memcpy_async(fin_intr(b) + 2 * i, i0 + 2 * i, 16);
    for (int i = lane; i < nf + 1; i += 32) __pipeline_memcpy_async(fin_os(b) + i, obs_start + f0 + i, 4);
    if (lane < nf) __pipeline_memcpy_async(fin_fp(b) + lane, fprob + f0 + lane, 4);
  };
  auto tile_range = [&](int64_t tile, int& kb, int& ke) {
    const int64_t f0 = tile * 32;
    kb = obs_start[f0];
    ke = obs_start[min(f0 + 32, T)];
  };

  int64_t tile = blockIdx.x;
  if (tile >= n_tiles) return;
  int kb, ke, kb_n = 0, ke_n = 0;
  tile_range(tile, kb, ke);
  issue_frames(0, tile * 32);
  issue_obs(in_tile(0), in_ofr(0), kb, min(k2pChunk, ke - kb));
  __pipeline_commit();
  int64_t tile_n = tile + gridDim.x;
  if (tile_n < n_tiles) tile_range(tile_n, kb_n, ke_n);
  const double inv_am2 = wp.alpha_is_two ? 0.0 : 1.0 / wp.am2;
  int buf = 0;
  for (; tile < n_tiles; tile = tile_n, tile_n += gridDim.x, buf ^= 1) {
    const int64_t f0 = tile * 32;
    const int64_t f = f0 + lane;
    const bool valid = f < T;
    // prefetch the next tile (its range was read one iteration ago) and start reading the range of the one after it
    const bool has_next = tile_n < n_tiles;
    int kb_nn = 0, ke_nn = 0;
    if (has_next) {
      issue_frames(buf ^ 1, tile_n * 32);
      issue_obs(in_tile(buf ^ 1), in_ofr(buf ^ 1), kb_n, min(k2pChunk, ke_n - kb_n));
    }
    __pipeline_commit();
    if (tile_n + gridDim.x < n_tiles) tile_range(tile_n + gridDim.x, kb_nn, ke_nn);
    __pipeline_wait_prior(1);
    __syncwarp();
    double* tilep = in_tile(buf);
    int32_t* ofr = in_ofr(buf);
    int k0 = 0, k1 = 0, p = 0;
    double Rt[9];
#pragma unroll
    for (int i = 0; i < 9; i++) Rt[i] = 0.0;
    double4 ci = make_double4(0, 0, 0, 0);
    {
      double* fd = fdat + lane * k2pFrameRec;
      double px = 0, py = 0, pz = 0;
      FrameWeight fw = {1.0, 1.0, 1.0};
      if (valid) {
        k0 = fin_os(buf)[lane]; k1 = fin_os(buf)[lane + 1];
        p = fin_fp(buf)[lane];
        if (k1 > k0) {
          const double c = c_obs[p];                      // L2-resident, issued before the rotation arithmetic
          const double* s = fin_st(buf) + lane * 10;
          ci = *reinterpret_cast<const double4*>(fin_intr(buf) + lane * 4);
          px = s[0]; py = s[1]; pz = s[2];
          const double qn = 1.0 / sqrt(s[3] * s[3] + s[4] * s[4] + s[5] * s[5] + s[6] * s[6]);
          const double x = s[3] * qn, y = s[4] * qn, z = s[5] * qn, w = s[6] * qn;
          Rt[0] = 1 - 2 * (y * y + z * z); Rt[1] = 2 * (x * y + z * w);     Rt[2] = 2 * (x * z - y * w);
          Rt[3] = 2 * (x * y - z * w);     Rt[4] = 1 - 2 * (x * x + z * z); Rt[5] = 2 * (y * z + x * w);
          Rt[6] = 2 * (x * z + y * w);     Rt[7] = 2 * (y * z - x * w);     Rt[8] = 1 - 2 * (x * x + y * y);
          fw = frame_weight(c, wp);
        }
      }
#pragma unroll
      for (int i = 0; i < 9; i++) fd[i] = Rt[i];
      fd[9] = px; fd[10] = py; fd[11] = pz;
      fd[12] = ci.x; fd[13] = ci.y; fd[14] = ci.z; fd[15] = ci.w;
      fd[16] = fw.inv_c; fd[17] = fw.inv_c2;
    }
    const int p_first = __shfl_sync(0xffffffffu, p, 0);
    const bool one_problem = __all_sync(0xffffffffu, !valid || p == p_first);
    double sN[5] = {0, 0, 0, 0, 0}, sNH[9], sHNH[6] = {0, 0, 0, 0, 0, 0}, sm[3] = {0, 0, 0}, sHm[3] = {0, 0, 0}, sabs = 0.0;
#pragma unroll
    for (int i = 0; i < 9; i++) sNH[i] = 0.0;
    double wloc = 0.0;
    __syncwarp();
    for (int base = kb; base < ke; base += k2pChunk) {
      const int n = min(k2pChunk, ke - base);
      if (base > kb) {
        // a tile with more than one chunk of observations (dense frames): the extra chunks are not prefetched
        issue_obs(tilep, ofr, base, n);
        __pipeline_commit();
        __pipeline_wait_prior(0);
        __syncwarp();
      }
      // ---- pass 1: one lane per observation (independent iterations: five chains in flight per lane)
#pragma unroll 5
      for (int i = lane; i < n; i += 32) {
        const int lf = ofr[i] - (int)f0;
        const double* fd = fdat + lf * k2pFrameRec;
        const double dx = tilep[i] - fd[9], dy = tilep[k2pChunk + i] - fd[10], dz = tilep[2 * k2pChunk + i] - fd[11];
        const double Xc = fd[0] * dx + fd[1] * dy + fd[2] * dz;
        const double Yc = fd[3] * dx + fd[4] * dy + fd[5] * dz;
        const double Zc = fd[6] * dx + fd[7] * dy + fd[8] * dz;
        const double d = 1.0 / fmax(Zc, 0.1);
        const double a = fd[12] * d, bb = fd[13] * d;
        const double ru = tilep[3 * k2pChunk + i] - (a * Xc + fd[14]), rv = tilep[4 * k2pChunk + i] - (bb * Yc + fd[15]);
        FrameWeight fw; fw.inv_c = fd[16]; fw.inv_am2 = inv_am2; fw.inv_c2 = fd[17];
        const double wraw = 0.5 * (robust_component_fast(ru, fw, wp) + robust_component_fast(rv, fw, wp));
        const double w = wraw * tilep[5 * k2pChunk + i];
        tilep[i] = Xc; tilep[k2pChunk + i] = Yc; tilep[2 * k2pChunk + i] = Zc;
        tilep[3 * k2pChunk + i] = ru; tilep[4 * k2pChunk + i] = rv;
        tilep[5 * k2pChunk + i] = w;
        tilep[6 * k2pChunk + i] = d;
        if (one_problem) wloc = fmax(wloc, wraw);
        else if (wraw > 0.0) atomicMax(&wmax[fin_fp(buf)[lf]], (unsigned long long)__double_as_longlong(wraw));
      }
      __syncwarp();
      // ---- pass 2: one lane per frame
      const int lo = max(k0, base), hi = min(k1, base + n);
      for (int k = lo; k < hi; k++) {
        const int i = k - base;
        const double Xc = tilep[i], Yc = tilep[k2pChunk + i], Zc = tilep[2 * k2pChunk + i];
        const double ru = tilep[3 * k2pChunk + i], rv = tilep[4 * k2pChunk + i];
        const double w = tilep[5 * k2pChunk + i], d = tilep[6 * k2pChunk + i];
        const double live = (Zc >= 0.1) ? 1.0 : 0.0;
        const double a = ci.x * d, bb = ci.y * d;
        const double cc = -a * Xc * d * live, ee = -bb * Yc * d * live;
        sabs += fabs(ru) + fabs(rv);
        const double wa = w * a, wb = w * bb;
        const double n00 = wa * a, n02 = wa * cc, n11 = wb * bb, n12 = wb * ee, n22 = w * (cc * cc + ee * ee);
        sN[0] += n00; sN[1] += n02; sN[2] += n11; sN[3] += n12; sN[4] += n22;
        const double h00 = -n02 * Yc, h01 = n02 * Xc - n00 * Zc, h02 = n00 * Yc;
        const double h10 = n11 * Zc - n12 * Yc, h11 = n12 * Xc, h12 = -n11 * Xc;
        const double h20 = n12 * Zc - n22 * Yc, h21 = n22 * Xc - n02 * Zc, h22 = n02 * Yc - n12 * Xc;
        sNH[0] += h00; sNH[1] += h01; sNH[2] += h02; sNH[3] += h10; sNH[4] += h11; sNH[5] += h12;
        sNH[6] += h20; sNH[7] += h21; sNH[8] += h22;
        sHNH[0] += Zc * h10 - Yc * h20; sHNH[1] += Zc * h11 - Yc * h21; sHNH[2] += Zc * h12 - Yc * h22;
        sHNH[3] += Xc * h21 - Zc * h01; sHNH[4] += Xc * h22 - Zc * h02; sHNH[5] += Yc * h02 - Xc * h12;
        const double m0 = wa * ru, m1 = wb * rv, m2 = w * (cc * ru + ee * rv);
        sm[0] += m0; sm[1] += m1; sm[2] += m2;
        sHm[0] += Zc * m1 - Yc * m2; sHm[1] += Xc * m2 - Zc * m0; sHm[2] += Yc * m0 - Xc * m1;
      }
      __syncwarp();
      for (int i = lane; i < n; i += 32) wu_out[base + i] = tilep[5 * k2pChunk + i];
      __syncwarp();
    }
    if (one_problem) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) wloc = fmax(wloc, __shfl_xor_sync(0xffffffffu, wloc, o));
      if (lane == 0 && wloc > 0.0) atomicMax(&wmax[p_first], (unsigned long long)__double_as_longlong(wloc));
    }
    if (valid) {
      const double N[9] = {sN[0], 0.0, sN[1], 0.0, sN[2], sN[3], sN[1], sN[3], sN[4]};
      double NR[9];
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) NR[i * 3 + j] = N[i * 3] * Rt[j] + N[i * 3 + 1] * Rt[3 + j] + N[i * 3 + 2] * Rt[6 + j];
      double out[VS_GREC];
      int idx = 0;
#pragma unroll
      for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int bcol = a; bcol < 3; bcol++)
          out[idx++] = Rt[a] * NR[bcol] + Rt[3 + a] * NR[3 + bcol] + Rt[6 + a] * NR[6 + bcol];
#pragma unroll
        for (int bcol = 0; bcol < 3; bcol++)
          out[idx++] = -2.0 * (Rt[a] * sNH[bcol] + Rt[3 + a] * sNH[3 + bcol] + Rt[6 + a] * sNH[6 + bcol]);
      }
      out[idx++] = 4.0 * sHNH[0]; out[idx++] = 4.0 * sHNH[1]; out[idx++] = 4.0 * sHNH[2];
      out[idx++] = 4.0 * sHNH[3]; out[idx++] = 4.0 * sHNH[4];
      out[idx++] = 4.0 * sHNH[5];
#pragma unroll
      for (int a = 0; a < 3; a++) out[21 + a] = -(Rt[a] * sm[0] + Rt[3 + a] * sm[1] + Rt[6 + a] * sm[2]);
#pragma unroll
      for (int a = 0; a < 3; a++) out[24 + a] = 2.0 * sHm[a];
      out[27] = sabs;
      double* g = grec + f * VS_GREC;
#pragma unroll
      for (int i = 0; i < VS_GREC; i += 2) *reinterpret_cast<double2*>(g + i) = make_double2(out[i], out[i + 1]);
    }
    __syncwarp();       // fdat and the consumed buffers are rewritten by the next iteration
    kb = kb_n; ke = ke_n;
    kb_n = kb_nn; ke_n = ke_nn;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Half-tile variant: ONE WARP OWNS 16 FRAMES, two lanes per frame.  Measurements that led here (B200, 1024 x 1000 x 10):
// the 32-frame two-pass kernel runs 9 warps per SM (24 KB of shared memory and 158 registers per warp) at 0.31 ms; the
// persistent double-buffered version of it -- every load prefetched one tile ahead, 4 warps per SM -- takes 0.49 ms.
// So the kernel is not waiting for memory: a lone warp issues one instruction per ~5 cycles (FP64 dependent-issue
// latency 8.25 cycles, issue interval 2.3, measured by tools/ubench/fp64_lat.cu) and only more resident warps fill the
// pipe.  Halving the tile halves the shared memory per warp (12 KB), the two lanes of a frame take its even / odd
// observations (5 instead of 10 sequential accumulation steps) and are combined by one xor-shuffle per sum.
// ---------------------------------------------------------------------------------------------------------
constexpr int kHtFrames = 16;
constexpr int kHtChunk = 160;                 // observations per chunk (16 frames x 10; more obs => more chunks)
constexpr int kHtSmemBytes = (k2pSlots * kHtChunk + kHtFrames * k2pFrameRec) * 8 + kHtChunk * 4;

template <int kMinBlocks>
__global__ void __launch_bounds__(32, kMinBlocks) k_obs_assemble_half(int64_t T, int64_t M,
                                                                      const int32_t* __restrict__ obs_start,
                                                                      const int32_t* __restrict__ oframe,
                                                                      const int32_t* __restrict__ fprob,
                                                                      const double* __restrict__ X,
                                                                      const double* __restrict__ uv,
                                                                      const double* __restrict__ conf,
                                                                      const double* __restrict__ st,
                                                                      const double* __restrict__ intr,
                                                                      const double* __restrict__ c_obs, WeightParams wp,
                                                                      double* __restrict__ wu_out, double* __restrict__ grec,
                                                                      unsigned long long* __restrict__ wmax,
                                                                      const int32_t* __restrict__ gate) {
  if (gate && *gate) return;
  extern __shared__ __align__(16) double smh[];
  double* tile = smh;                                                   // [k2pSlots][kHtChunk]
  double* fdat = smh + k2pSlots * kHtChunk;                             // [16][k2pFrameRec]
  int32_t* ofr = reinterpret_cast<int32_t*>(fdat + kHtFrames * k2pFrameRec);
  const int lane = threadIdx.x;
  const int half = lane & 1;                  // which observations of the frame this lane accumulates
  const int lfr = lane >> 1;                  // local frame of the lane pair
  const int64_t f0 = (int64_t)blockIdx.x * kHtFrames;
  const int64_t f = f0 + lfr;
  const bool valid = f < T;
  const int kb = obs_start[f0];
  const int ke = obs_start[min(f0 + kHtFrames, T)];
  auto issue_chunk = [&](int base) {
    const int n = min(kHtChunk, ke - base);
    for (int i = lane; i < n; i += 32) {
      const int k = base + i;
      __pipeline_memcpy_async(&tile[i], &X[k], 8);
      __pipeline_memcpy_async(&tile[kHtChunk + i], &X[M + k], 8);
      __pipeline_memcpy_async(&tile[2 * kHtChunk + i], &X[2 * M + k], 8);
      __pipeline_memcpy_async(&tile[3 * kHtChunk + i], &uv[k], 8);
      __pipeline_memcpy_async(&tile[4 * kHtChunk + i], &uv[M + k], 8);
      __pipeline_memcpy_async(&tile[5 * kHtChunk + i], &conf[k], 8);
      __pipeline_memcpy_async(&ofr[i], &oframe[k], 4);
    }
    __pipeline_commit();
  };
  issue_chunk(kb);
  int k0 = 0, k1 = 0, p = 0;
  double cix = 0.0, ciy = 0.0;
  {
    // frame constants: computed by the even lane of the pair, shared through fdat
    double* fd = fdat + lfr * k2pFrameRec;
    if (valid) {
      k0 = obs_start[f]; k1 = obs_start[f + 1];
      p = fprob[f];
    }
    if (half == 0) {
      double Rt[9];
#pragma unroll
      for (int i = 0; i < 9; i++) Rt[i] = 0.0;
      double4 ci = make_double4(0, 0, 0, 0);
      double px = 0, py = 0, pz = 0;
      FrameWeight fw = {1.0, 1.0, 1.0};
      if (valid && k1 > k0) {
        fw = frame_weight(c_obs[p], wp);
        const double* s = st + f * 10;
        ci = *reinterpret_cast<const double4*>(intr + f * 4);
        px = s[0]; py = s[1]; pz = s[2];
        const double qn = 1.0 / sqrt(s[3] * s[3] + s[4] * s[4] + s[5] * s[5] + s[6] * s[6]);
        const double x = s[3] * qn, y = s[4] * qn, z = s[5] * qn, w = s[6] * qn;
        Rt[0] = 1 - 2 * (y * y + z * z); Rt[1] = 2 * (x * y + z * w);     Rt[2] = 2 * (x * z - y * w);
        Rt[3] = 2 * (x * y - z * w);     Rt[4] = 1 - 2 * (x * x + z * z); Rt[5] = 2 * (y * z + x * w);
        Rt[6] = 2 * (x * z + y * w);     Rt[7] = 2 * (y * z - x * w);     Rt[8] = 1 - 2 * (x * x + y * y);
      }
#pragma unroll
      for (int i = 0; i < 9; i++) fd[i] = Rt[i];
      fd[9] = px; fd[10] = py; fd[11] = pz;
      fd[12] = ci.x; fd[13] = ci.y; fd[14] = ci.z; fd[15] = ci.w;
      fd[16] = fw.inv_c; fd[17] = fw.inv_c2;
    }
    __syncwarp();
    cix = fd[12]; ciy = fd[13];
  }
  const int p_first = __shfl_sync(0xffffffffu, p, 0);
  const bool one_problem = __all_sync(0xffffffffu, !valid || p == p_first);
  const double inv_am2 = wp.alpha_is_two ? 0.0 : 1.0 / wp.am2;
  double sN[5] = {0, 0, 0, 0, 0}, sNH[9], sHNH[6] = {0, 0, 0, 0, 0, 0}, sm[3] = {0, 0, 0}, sHm[3] = {0, 0, 0}, sabs = 0.0;
#pragma unroll
  for (int i = 0; i < 9; i++) sNH[i] = 0.0;
  double wloc = 0.0;
  for (int base = kb; base < ke; base += kHtChunk) {
    const int n = min(kHtChunk, ke - base);
    __pipeline_wait_prior(0);
    __syncwarp();
    // ---- pass 1: one lane per observation
#pragma unroll 5
    for (int i = lane; i < n; i += 32) {
      const int lf = ofr[i] - (int)f0;
      const double* fd = fdat + lf * k2pFrameRec;
      const double dx = tile[i] - fd[9], dy = tile[kHtChunk + i] - fd[10], dz = tile[2 * kHtChunk + i] - fd[11];
      const double Xc = fd[0] * dx + fd[1] * dy + fd[2] * dz;
      const double Yc = fd[3] * dx + fd[4] * dy + fd[5] * dz;
      const double Zc = fd[6] * dx + fd[7] * dy + fd[8] * dz;
      const double d = 1.0 / fmax(Zc, 0.1);
      const double a = fd[12] * d, bb = fd[13] * d;
      const double ru = tile[3 * kHtChunk + i] - (a * Xc + fd[14]), rv = tile[4 * kHtChunk + i] - (bb * Yc + fd[15]);
      FrameWeight fw; fw.inv_c = fd[16]; fw.inv_am2 = inv_am2; fw.inv_c2 = fd[17];
      const double wraw = 0.5 * (robust_component_fast(ru, fw, wp) + robust_component_fast(rv, fw, wp));
      const double w = wraw * tile[5 * kHtChunk + i];
      tile[i] = Xc; tile[kHtChunk + i] = Yc; tile[2 * kHtChunk + i] = Zc;
      tile[3 * kHtChunk + i] = ru; tile[4 * kHtChunk + i] = rv;
      tile[5 * kHtChunk + i] = w;
      tile[6 * kHtChunk + i] = d;
      if (one_problem) wloc = fmax(wloc, wraw);
      else if (wraw > 0.0) atomicMax(&wmax[fprob[f0 + lf]], (unsigned long long)__double_as_longlong(wraw));
    }
    __syncwarp();
    // ---- pass 2: two lanes per frame, even / odd observations of the frame (fixed split => deterministic sums)
    const int lo = max(k0, base), hi = min(k1, base + n);
    for (int k = lo + ((half + 2 - ((lo - k0) & 1)) & 1); k < hi; k += 2) {     // k - k0 has the parity of `half`
      const int i = k - base;
      const double Xc = tile[i], Yc = tile[kHtChunk + i], Zc = tile[2 * kHtChunk + i];
      const double ru = tile[3 * kHtChunk + i], rv = tile[4 * kHtChunk + i];
      const double w = tile[5 * kHtChunk + i], d = tile[6 * kHtChunk + i];
      const double live = (Zc >= 0.1) ? 1.0 : 0.0;
      const double a = cix * d, bb = ciy * d;
      const double cc = -a * Xc * d * live, ee = -bb * Yc * d * live;
      sabs += fabs(ru) + fabs(rv);
      const double wa = w * a, wb = w * bb;
      const double n00 = wa * a, n02 = wa * cc, n11 = wb * bb, n12 = wb * ee, n22 = w * (cc * cc + ee * ee);
      sN[0] += n00; sN[1] += n02; sN[2] += n11; sN[3] += n12; sN[4] += n22;
      const double h00 = -n02 * Yc, h01 = n02 * Xc - n00 * Zc, h02 = n00 * Yc;
      const double h10 = n11 * Zc - n12 * Yc, h11 = n12 * Xc, h12 = -n11 * Xc;
      const double h20 = n12 * Zc - n22 * Yc, h21 = n22 * Xc - n02 * Zc, h22 = n02 * Yc - n12 * Xc;
      sNH[0] += h00; sNH[1] += h01; sNH[2] += h02; sNH[3] += h10; sNH[4] += h11; sNH[5] += h12;
      sNH[6] += h20; sNH[7] += h21; sNH[8] += h22;
      sHNH[0] += Zc * h10 - Yc * h20; sHNH[1] += Zc * h11 - Yc * h21; sHNH[2] += Zc * h12 - Yc * h22;
      sHNH[3] += Xc * h21 - Zc * h01; sHNH[4] += Xc * h22 - Zc * h02; sHNH[5] += Yc * h02 - Xc * h12;
      const double m0 = wa * ru, m1 = wb * rv, m2 = w * (cc * ru + ee * rv);
      sm[0] += m0; sm[1] += m1; sm[2] += m2;
      sHm[0] += Zc * m1 - Yc * m2; sHm[1] += Xc * m2 - Zc * m0; sHm[2] += Yc * m0 - Xc * m1;
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) wu_out[base + i] = tile[5 * kHtChunk + i];
    __syncwarp();
    if (base + kHtChunk < ke) issue_chunk(base + kHtChunk);
  }
  if (one_problem) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wloc = fmax(wloc, __shfl_xor_sync(0xffffffffu, wloc, o));
    if (lane == 0 && wloc > 0.0) atomicMax(&wmax[p_first], (unsigned long long)__double_as_longlong(wloc));
  }
  // combine the two halves of every frame (even + odd, always in this order)
  auto pair_sum = [&](double v) {
    const double o = __shfl_xor_sync(0xffffffffu, v, 1);
    return half == 0 ? v + o : o + v;
  };
#pragma unroll
  for (int i = 0; i < 5; i++) sN[i] = pair_sum(sN[i]);
#pragma unroll
  for (int i = 0; i < 9; i++) sNH[i] = pair_sum(sNH[i]);
#pragma unroll
  for (int i = 0; i < 6; i++) sHNH[i] = pair_sum(sHNH[i]);
#pragma unroll
  for (int i = 0; i < 3; i++) { sm[i] = pair_sum(sm[i]); sHm[i] = pair_sum(sHm[i]); }
  sabs = pair_sum(sabs);
  if (valid) {
    // the rotation back to the world frame is split between the two lanes: even lane rows 0..13, odd lane rows 14..27
    const double* Rt = fdat + lfr * k2pFrameRec;
    const double N[9] = {sN[0], 0.0, sN[1], 0.0, sN[2], sN[3], sN[1], sN[3], sN[4]};
    double NR[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) NR[i * 3 + j] = N[i * 3] * Rt[j] + N[i * 3 + 1] * Rt[3 + j] + N[i * 3 + 2] * Rt[6 + j];
    double out[VS_GREC];
    int idx = 0;
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
      for (int bcol = a; bcol < 3; bcol++)
        out[idx++] = Rt[a] * NR[bcol] + Rt[3 + a] * NR[3 + bcol] + Rt[6 + a] * NR[6 + bcol];
#pragma unroll
      for (int bcol = 0; bcol < 3; bcol++)
        out[idx++] = -2.0 * (Rt[a] * sNH[bcol] + Rt[3 + a] * sNH[3 + bcol] + Rt[6 + a] * sNH[6 + bcol]);
    }
    out[idx++] = 4.0 * sHNH[0]; out[idx++] = 4.0 * sHNH[1]; out[idx++] = 4.0 * sHNH[2];
    out[idx++] = 4.0 * sHNH[3]; out[idx++] = 4.0 * sHNH[4];
    out[idx++] = 4.0 * sHNH[5];
#pragma unroll
    for (int a = 0; a < 3; a++) out[21 + a] = -(Rt[a] * sm[0] + Rt[3 + a] * sm[1] + Rt[6 + a] * sm[2]);
#pragma unroll
    for (int a = 0; a < 3; a++) out[24 + a] = 2.0 * sHm[a];
    out[27] = sabs;
    double* g = grec + f * VS_GREC;
#pragma unroll
    for (int i = 0; i < VS_GREC; i += 2)
      if ((i < 14) == (half == 0)) *reinterpret_cast<double2*>(g + i) = make_double2(out[i], out[i + 1]);
  }
}

int launch_obs_assemble(vinsat_batch* b, double alpha) {
  vinsat_ctx* ctx = b->ctx;
  WeightParams wp;
  wp.am2 = fabs(alpha - 2.0);
  wp.ex = alpha / 2.0 - 1.0;
  wp.alpha_is_two = (wp.ex == 0.0) ? 1 : 0;
  wp.ex_is_mhalf = (wp.ex == -0.5) ? 1 : 0;
  if (b->gate_arg) {
    int rc = launch_gated_zero(b, b->wmax, b->P, nullptr, 0);
    if (rc != VINSAT_OK) return rc;
  } else {
    VS_CUDA(ctx, cudaMemsetAsync(b->wmax, 0, b->P * sizeof(unsigned long long), ctx->stream));
  }
  if (b->T == 0) return VINSAT_OK;
  static int variant = getenv("VINSAT_ASM_VARIANT") ? atoi(getenv("VINSAT_ASM_VARIANT")) : 0;
  if (variant == 32) {            // thread-per-frame walk (kept for comparison)
    const int smem = 6 * 352 * (int)sizeof(double);
    VS_SMEM_OPTIN(ctx, SM_ASM_STAGED, (k_obs_assemble_staged<32, 352, 1, 12>), smem);
    VS_LAUNCH(ctx, F_OBS_ASSEMBLE, (k_obs_assemble_staged<32, 352, 1, 12>), ceil_div(b->T, 32), 32, smem, b->T, b->M,
              b->obs_start, b->fprob, b->X, b->uv, b->conf, b->st, b->intr, b->c_obs, wp, b->wu, b->grec, b->wmax);
    return VINSAT_OK;
  }
  if (variant == 0 || variant == 16 || variant == 12) {      // half-tile kernel (default)
    const int grid = (int)ceil_div(b->T, kHtFrames);
    if (variant == 12) {
      VS_SMEM_OPTIN(ctx, SM_SPARE1, k_obs_assemble_half<12>, kHtSmemBytes);
      VS_LAUNCH(ctx, F_OBS_ASSEMBLE, k_obs_assemble_half<12>, grid, 32, kHtSmemBytes, b->T, b->M, b->obs_start, b->oframe,
                b->fprob, b->X, b->uv, b->conf, b->st, b->intr, b->c_obs, wp, b->wu, b->grec, b->wmax, b->gate_arg);
    } else {
      VS_SMEM_OPTIN(ctx, SM_SPARE2, k_obs_assemble_half<16>, kHtSmemBytes);
      VS_LAUNCH(ctx, F_OBS_ASSEMBLE, k_obs_assemble_half<16>, grid, 32, kHtSmemBytes, b->T, b->M, b->obs_start, b->oframe,
                b->fprob, b->X, b->uv, b->conf, b->st, b->intr, b->c_obs, wp, b->wu, b->grec, b->wmax, b->gate_arg);
    }
    return VINSAT_OK;
  }
  if (variant == 4) {             // persistent double-buffered kernel (kept for the record: slower, see above)
    const int64_t n_tiles = ceil_div(b->T, 32);
    const int grid = (int)std::min<int64_t>(n_tiles, (int64_t)ctx->sm_count * 4);
    VS_SMEM_OPTIN(ctx, SM_SPARE0, k_obs_assemble_pers, kPaSmemBytes);
    VS_LAUNCH(ctx, F_OBS_ASSEMBLE, k_obs_assemble_pers, grid, 32, kPaSmemBytes, b->T, b->M, b->obs_start, b->oframe,
              b->fprob, b->X, b->uv, b->conf, b->st, b->intr, b->c_obs, wp, b->wu, b->grec, b->wmax, b->gate_arg);
    return VINSAT_OK;
  }
  const int smem = (k2pSlots * k2pChunk + 32 * k2pFrameRec) * (int)sizeof(double) + k2pChunk * (int)sizeof(int32_t);
  VS_SMEM_OPTIN(ctx, SM_ASM_2PASS, k_obs_assemble_2pass, smem);
  VS_LAUNCH(ctx, F_OBS_ASSEMBLE, k_obs_assemble_2pass, ceil_div(b->T, 32), 32, smem, b->T, b->M, b->obs_start, b->oframe,
            b->fprob, b->X, b->uv, b->conf, b->st, b->intr, b->c_obs, wp, b->wu, b->grec, b->wmax, b->gate_arg);
  return VINSAT_OK;
}

// trial residual, observation part: e_obs[f] = sum_k wu_k (|ru| + |rv|) at st_new (BA_filtering.py:61,66)
constexpr int kTrialUnroll = 5;
template <int kGroup>
__global__ void __launch_bounds__(256) k_obs_trial(int64_t T, int64_t M, const int32_t* __restrict__ obs_start,
                                                   const int32_t* __restrict__ fprob,
                                                   const int32_t* __restrict__ active, const double* __restrict__ X,
                                                   const double* __restrict__ uv, const double* __restrict__ wu,
                                                   const double* __restrict__ st, const double* __restrict__ intr,
                                                   double* __restrict__ e_obs, double* __restrict__ r_next,
                                                   const int32_t* __restrict__ gate) {
  if (gate && *gate) return;      // speculative launch behind an LM loop that is not finished (batch.cu)
  const int gl = threadIdx.x & (kGroup - 1);
  const int64_t f = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / kGroup;
  const bool valid = f < T;
  double e = 0.0;
  bool live = false;
  if (valid) {
    live = active[fprob[f]] != 0;
    const int k0 = obs_start[f], k1 = obs_start[f + 1];
    if (live && k1 > k0) {
      const double* s = st + f * 10;
      const double4 ci = *reinterpret_cast<const double4*>(intr + f * 4);
      const Quat q = {s[3], s[4], s[5], s[6]};
      // kTrialUnroll observations per pass with ALL their loads issued before the first use: the kernel is bound by
      // memory latency (ncu r01: long_scoreboard 20 of 24 stall cycles per issue at 36 % occupancy), so the loads in
      // flight per thread are what sets its bandwidth.  Out-of-range slots re-read the last observation and are masked.
      for (int kb = k0 + gl; kb < k1; kb += kGroup * kTrialUnroll) {
        double x[kTrialUnroll], y[kTrialUnroll], z[kTrialUnroll], mu[kTrialUnroll], mv[kTrialUnroll], w[kTrialUnroll];
#pragma unroll
        for (int j = 0; j < kTrialUnroll; j++) {
          const int k = min(kb + j * kGroup, k1 - 1);
          x[j] = X[k]; y[j] = X[M + k]; z[j] = X[2 * M + k];
          mu[j] = uv[k]; mv[j] = uv[M + k]; w[j] = wu[k];
        }
#pragma unroll
        for (int j = 0; j < kTrialUnroll; j++) {
          const int k = kb + j * kGroup;
          if (k < k1) {
            ProjOut o = project_exact(s[0], s[1], s[2], q, x[j], y[j], z[j], ci.x, ci.y, ci.z, ci.w);
            const double ru = xsub(mu[j], o.u), rv = xsub(mv[j], o.v);
            r_next[k] = ru;
            r_next[M + k] = rv;
            e += fabs(ru * w[j]) + fabs(rv * w[j]);
          }
        }
      }
    }
  }
  e = group_sum<kGroup>(e);
  if (valid && live && gl == 0) e_obs[f] = e;
}

int launch_obs_trial(vinsat_batch* b) {
  vinsat_ctx* ctx = b->ctx;
  if (b->T == 0) return VINSAT_OK;
  if (b->M <= 16 * b->T) {       // sparse frames: 2 lanes per frame (measured best of 1/2/4/8 at K = 10)
    VS_LAUNCH(ctx, F_TRIAL, k_obs_trial<2>, ceil_div(b->T * 2, 256), 256, 0, b->T, b->M, b->obs_start, b->fprob,
              b->active, b->X, b->uv, b->wu, b->st_new, b->intr, b->e_obs, b->r_next, b->gate_arg);
  } else {
    VS_LAUNCH(ctx, F_TRIAL, k_obs_trial<8>, ceil_div(b->T * 8, 256), 256, 0, b->T, b->M, b->obs_start, b->fprob,
              b->active, b->X, b->uv, b->wu, b->st_new, b->intr, b->e_obs, b->r_next, b->gate_arg);
  }
  return VINSAT_OK;
}

}  // namespace vs
