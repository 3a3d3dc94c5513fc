// Monte-Carlo input generation and scoring on the device (configs[3] of BASELINE.json: many OD problems with noise
// sweeps).  The reference's "Monte Carlo" is a sequential loop over stored sequences (od_pipe.py:1063-1086) whose only
// random ingredients are the perturbed initial guess (od_pipe.py:962-969: position + N(0, 100 km), rotation
// exp(log q + N(0, 0.2 rad)), velocity + N(0, 0.1 mean|v|)) and the detector's pixel noise.  Re-drawing those on the
// host costs ~0.5 s per 1024-problem chunk against 40 ms of solve, so they are drawn here from a counter-based
// generator (Philox4x32-10): chunk seed + element index -> the same numbers on any GPU, in any order.
#include "common.cuh"
#include "launch.h"

using namespace vs;

namespace {

#define VS_TRY(expr)            \
  do {                          \
    int _rc = (expr);           \
    if (_rc != VINSAT_OK) return _rc; \
  } while (0)

struct U4 { uint32_t x, y, z, w; };

__device__ __forceinline__ U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = {hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0};
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}

// two independent N(0,1) from one Philox block (Box-Muller on 2 x 53-bit uniforms in (0,1])
__device__ __forceinline__ void normal2(uint64_t seed, uint32_t stream, uint64_t index, double& n0, double& n1) {
  const U4 r = philox4x32_10({(uint32_t)index, (uint32_t)(index >> 32), stream, 0x5eedu}, (uint32_t)seed, (uint32_t)(seed >> 32));
  const uint64_t a = ((uint64_t)r.x << 32) | r.y, b = ((uint64_t)r.z << 32) | r.w;
  const double u0 = ((double)(a >> 11) + 1.0) * (1.0 / 9007199254740992.0);
  const double u1 = ((double)(b >> 11) + 1.0) * (1.0 / 9007199254740992.0);
  const double rad = sqrt(-2.0 * log(u0));
  double s, c;
  sincospi(2.0 * u1, &s, &c);
  n0 = rad * c;
  n1 = rad * s;
}

// BA_utils.py:949-967
__device__ __forceinline__ void quat_log(const Quat& q, double& lx, double& ly, double& lz) {
  const double n = sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  auto clip = [](double v) { return v < -1.0 ? -1.0 : (v > 1.0 ? 1.0 : v); };
  const double x = clip(q.x / n), y = clip(q.y / n), z = clip(q.z / n), w = clip(q.w / n);
  const double theta = 2.0 * acos(w);
  const double sh = sin(0.5 * theta);
  if (sh == 0.0) { lx = ly = lz = 0.0; return; }
  lx = (x / sh) * theta; ly = (y / sh) * theta; lz = (z / sh) * theta;
}

// initial guess of od_pipe.py:962-969 around the true states
__global__ void __launch_bounds__(128) k_mc_perturb_states(int64_t T, const double* __restrict__ st_true, uint64_t seed,
                                                           double pos_sigma, double rot_sigma, double vel_sigma,
                                                           double* __restrict__ st) {
  const int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (f >= T) return;
  const double* s = st_true + f * 10;
  double n[10];
#pragma unroll
  for (int k = 0; k < 5; k++) normal2(seed, 1u + k, (uint64_t)f, n[2 * k], n[2 * k + 1]);
  double* o = st + f * 10;
  o[0] = s[0] + pos_sigma * n[0]; o[1] = s[1] + pos_sigma * n[1]; o[2] = s[2] + pos_sigma * n[2];
  const Quat q = {s[3], s[4], s[5], s[6]};
  double lx, ly, lz;
  quat_log(q, lx, ly, lz);
  // quaternion_exp (BA_utils.py:970-985) takes the ROTATION VECTOR (half-angle inside): exp(log q + noise)
  const Quat e = qexp(lx + rot_sigma * n[3], ly + rot_sigma * n[4], lz + rot_sigma * n[5]);
  o[3] = e.x; o[4] = e.y; o[5] = e.z; o[6] = e.w;
  o[7] = s[7] + vel_sigma * n[6]; o[8] = s[8] + vel_sigma * n[7]; o[9] = s[9] + vel_sigma * n[8];
}

__global__ void __launch_bounds__(256) k_mc_perturb_uv(int64_t M, const double* __restrict__ uv_true, uint64_t seed,
                                                       double sigma_px, double* __restrict__ uv) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= M) return;
  double n0, n1;
  normal2(seed, 0u, (uint64_t)k, n0, n1);
  uv[k] = uv_true[k] + sigma_px * n0;
  uv[M + k] = uv_true[M + k] + sigma_px * n1;
}

// per problem: max |p - p_true| and max |v - v_true| over its frames (one warp per problem)
__global__ void __launch_bounds__(128) k_mc_errors(int64_t P, const int64_t* __restrict__ frame_off,
                                                   const double* __restrict__ st, const double* __restrict__ st_true,
                                                   const double* __restrict__ vel_true, double* __restrict__ pos_err,
                                                   double* __restrict__ vel_err) {
  const int64_t p = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (p >= P) return;
  double ep = 0.0, ev = 0.0;
  for (int64_t f = frame_off[p] + lane; f < frame_off[p + 1]; f += 32) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      ep = fmax(ep, fabs(st[f * 10 + k] - st_true[f * 10 + k]));
      const double vt = vel_true ? vel_true[f * 3 + k] : st_true[f * 10 + 7 + k];
      ev = fmax(ev, fabs(st[f * 10 + 7 + k] - vt));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ep = fmax(ep, __shfl_xor_sync(0xffffffffu, ep, o));
    ev = fmax(ev, __shfl_xor_sync(0xffffffffu, ev, o));
  }
  if (lane == 0) { pos_err[p] = ep; vel_err[p] = ev; }
}

}  // namespace

extern "C" {

int vinsat_batch_mc_set_truth(vinsat_batch* b, const double* states_true, const double* uv_true, const double* vel_true) {
  if (!b || !states_true) return set_error(b ? b->ctx : nullptr, VINSAT_EINVAL, "vinsat_batch_mc_set_truth: NULL argument");
  vinsat_ctx* ctx = b->ctx;
  VS_CHECK_ARG(ctx, b->M == 0 || uv_true);
  VS_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t T = b->T, M = b->M;
  if (!b->mc_st_true) VS_CUDA(ctx, cudaMalloc((void**)&b->mc_st_true, T * 10 * sizeof(double)));
  if (!b->mc_uv_true) VS_CUDA(ctx, cudaMalloc((void**)&b->mc_uv_true, std::max<int64_t>(M, 1) * 2 * sizeof(double)));
  if (vel_true && !b->mc_vel_true) VS_CUDA(ctx, cudaMalloc((void**)&b->mc_vel_true, T * 3 * sizeof(double)));
  if (!b->mc_err) VS_CUDA(ctx, cudaMalloc((void**)&b->mc_err, b->P * 2 * sizeof(double)));
  VS_CUDA(ctx, cudaMemcpyAsync(b->mc_st_true, states_true, T * 10 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (vel_true)
    VS_CUDA(ctx, cudaMemcpyAsync(b->mc_vel_true, vel_true, T * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (M > 0) {
    double* tmp = (double*)ctx_scratch(ctx, (size_t)M * 2 * sizeof(double));
    if (!tmp) return set_error(ctx, VINSAT_ENOMEM, "scratch allocation failed");
    VS_CUDA(ctx, cudaMemcpyAsync(tmp, uv_true, M * 2 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    VS_TRY(launch_aos_to_soa(ctx, tmp, b->mc_uv_true, M, 2));
  }
  VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return VINSAT_OK;
}

int vinsat_batch_mc_perturb(vinsat_batch* b, uint64_t seed, double sigma_px, double pos_sigma, double rot_sigma,
                            double vel_sigma) {
  if (!b) return set_error(nullptr, VINSAT_EINVAL, "vinsat_batch_mc_perturb: NULL batch");
  vinsat_ctx* ctx = b->ctx;
  if (!b->mc_st_true) return set_error(ctx, VINSAT_EINVAL, "vinsat_batch_mc_set_truth has not been called");
  VS_CHECK_ARG(ctx, sigma_px >= 0 && pos_sigma >= 0 && rot_sigma >= 0 && vel_sigma >= 0);
  VS_CUDA(ctx, cudaSetDevice(ctx->device));
  VS_LAUNCH(ctx, F_SIM, k_mc_perturb_states, ceil_div(b->T, 128), 128, 0, b->T, b->mc_st_true, seed, pos_sigma, rot_sigma,
            vel_sigma, b->st);
  if (b->M > 0)
    VS_LAUNCH(ctx, F_SIM, k_mc_perturb_uv, ceil_div(b->M, 256), 256, 0, b->M, b->mc_uv_true, seed, sigma_px, b->uv);
  b->r_valid = false;
  b->have_iter = false;
  return VINSAT_OK;
}

int vinsat_batch_mc_errors(vinsat_batch* b, double* pos_err_out, double* vel_err_out) {
  if (!b || !pos_err_out || !vel_err_out)
    return set_error(b ? b->ctx : nullptr, VINSAT_EINVAL, "vinsat_batch_mc_errors: NULL argument");
  vinsat_ctx* ctx = b->ctx;
  if (!b->mc_st_true) return set_error(ctx, VINSAT_EINVAL, "vinsat_batch_mc_set_truth has not been called");
  VS_CUDA(ctx, cudaSetDevice(ctx->device));
  VS_LAUNCH(ctx, F_ACCEPT, k_mc_errors, ceil_div(b->P * 32, 128), 128, 0, b->P, b->d_frame_off, b->st, b->mc_st_true,
            b->mc_vel_true, b->mc_err, b->mc_err + b->P);
  VS_CUDA(ctx, cudaMemcpyAsync(pos_err_out, b->mc_err, b->P * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  VS_CUDA(ctx, cudaMemcpyAsync(vel_err_out, b->mc_err + b->P, b->P * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return VINSAT_OK;
}

}  // extern "C"
