// Stand-alone operators of the C ABI on reference-layout buffers (host or device pointers):
// landmark_project, predict, propagate chain, orbit propagation.
#include "common.cuh"
#include "launch.h"

using namespace vs;

namespace {

#define VS_TRY(expr)            \
  do {                          \
    int _rc = (expr);           \
    if (_rc != VINSAT_OK) return _rc; \
  } while (0)

// Input staging: returns a device pointer for `src` (copying if it lives on the host).
template <typename T>
struct InBuf {
  DevBuf<T> own;
  const T* p = nullptr;
  int init(vinsat_ctx* ctx, int mem, const T* src, int64_t n) {
    if (mem == VINSAT_MEM_DEVICE || n == 0) { p = src; return VINSAT_OK; }
    if (own.alloc(n) != cudaSuccess) { cudaGetLastError(); return set_error(ctx, VINSAT_ENOMEM, "cudaMalloc failed"); }
    VS_CUDA(ctx, cudaMemcpyAsync(own.p, src, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    p = own.p;
    return VINSAT_OK;
  }
};

template <typename T>
struct OutBuf {
  DevBuf<T> own;
  T* p = nullptr;
  T* host = nullptr;
  int64_t n = 0;
  int init(vinsat_ctx* ctx, int mem, T* dst, int64_t n_) {
    n = n_;
    if (!dst) { p = nullptr; return VINSAT_OK; }
    if (mem == VINSAT_MEM_DEVICE) { p = dst; return VINSAT_OK; }
    if (own.alloc(n) != cudaSuccess) { cudaGetLastError(); return set_error(ctx, VINSAT_ENOMEM, "cudaMalloc failed"); }
    p = own.p;
    host = dst;
    return VINSAT_OK;
  }
  int finish(vinsat_ctx* ctx) {
    if (host && n) VS_CUDA(ctx, cudaMemcpyAsync(host, p, n * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    return VINSAT_OK;
  }
};

__global__ void k_gaps(int64_t T, const int64_t* __restrict__ time_idx, int32_t* __restrict__ gap,
                       int32_t* __restrict__ err) {
  const int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (f >= T) return;
  int32_t g = 0;
  if (f + 1 < T) {
    const int64_t d = time_idx[f + 1] - time_idx[f];
    if (d <= 0 || d > 100000000) atomicOr(err, 1);
    g = (int32_t)d;
  }
  gap[f] = g;
}

__global__ void k_predict_extract(int64_t T, const double* __restrict__ drec, double* __restrict__ r_pred,
                                  double* __restrict__ Phi, double* __restrict__ qgrad,
                                  double* __restrict__ Hd, double* __restrict__ Ho) {
  const int64_t f = blockIdx.x;
  const double* d = drec + f * VS_DREC;
  for (int e = threadIdx.x; e < VS_DREC; e += blockDim.x) {
    const double v = d[e];
    if (e < 36) { if (Phi && f + 1 < T) Phi[f * 36 + e] = v; }
    else if (e < 43) { if (r_pred && f + 1 < T) r_pred[f * 7 + (e - 36)] = v; }
    else if (e < 46) { if (qgrad) qgrad[f * 3 + (e - 43)] = v; }
    else if (e < 55) { if (Hd) Hd[f * 9 + (e - 46)] = v; }
    else { if (Ho && f + 1 < T) Ho[f * 9 + (e - 55)] = v; }
  }
}

}  // namespace

extern "C" {

int vinsat_landmark_project(vinsat_ctx* ctx, int mem, int64_t n_frames, int64_t n_obs, const double* states,
                            const double* intrinsics, const double* landmarks_xyz, const int64_t* ii,
                            double* uv_out, double* Jg_out) {
  VS_CHECK_ARG(ctx, ctx != nullptr);
  VS_CHECK_ARG(ctx, n_frames >= 1 && n_obs >= 0);
  VS_CHECK_ARG(ctx, states && intrinsics && (n_obs == 0 || (landmarks_xyz && ii && uv_out)));
  VS_CUDA(ctx, cudaSetDevice(ctx->device));
  InBuf<double> st, in, xyz;
  InBuf<int64_t> idx;
  OutBuf<double> uv, J;
  DevBuf<int32_t> err;
  VS_TRY(st.init(ctx, mem, states, n_frames * 10));
  VS_TRY(in.init(ctx, mem, intrinsics, n_frames * 4));
  VS_TRY(xyz.init(ctx, mem, landmarks_xyz, n_obs * 3));
  VS_TRY(idx.init(ctx, mem, ii, n_obs));
  VS_TRY(uv.init(ctx, mem, uv_out, n_obs * 2));
  VS_TRY(J.init(ctx, mem, Jg_out, n_obs * 18));
  VS_CUDA(ctx, err.alloc(1));
  VS_CUDA(ctx, cudaMemsetAsync(err.p, 0, sizeof(int32_t), ctx->stream));
  VS_TRY(launch_project_aos(ctx, n_frames, n_obs, st.p, in.p, xyz.p, idx.p, uv.p, J.p, err.p));
  VS_TRY(uv.finish(ctx));
  VS_TRY(J.finish(ctx));
  int32_t h_err = 0;
  VS_CUDA(ctx, cudaMemcpyAsync(&h_err, err.p, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (h_err) return set_error(ctx, VINSAT_EINVAL, "landmark_project: ii out of range [0,%lld)", (long long)n_frames);
  return VINSAT_OK;
}

int vinsat_predict(vinsat_ctx* ctx, int mem, int64_t n_frames, const double* states, const double* cum_rot,
                   const int64_t* time_idx, double quat_coeff, double vel_coeff, int mode, double* r_pred_out,
                   double* x_pred_out, double* Phi_out, double* qgrad_out, double* Hq_diag_out, double* Hq_off_out) {
  VS_CHECK_ARG(ctx, ctx != nullptr);
  VS_CHECK_ARG(ctx, n_frames >= 1 && states && cum_rot && time_idx);
  VS_CHECK_ARG(ctx, mode == VINSAT_MODE_STEP1S || mode == VINSAT_MODE_SKIP100);
  const bool jac = Phi_out != nullptr;
  VS_CHECK_ARG(ctx, jac == (qgrad_out != nullptr) && jac == (Hq_diag_out != nullptr) && jac == (Hq_off_out != nullptr));
  VS_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t T = n_frames;
  InBuf<double> st, cr;
  InBuf<int64_t> ti;
  OutBuf<double> rp, xp, ph, qg, hd, ho;
  DevBuf<int32_t> gap, err;
  DevBuf<double> drec;
  VS_TRY(st.init(ctx, mem, states, T * 10));
  VS_TRY(cr.init(ctx, mem, cum_rot, T * 4));
  VS_TRY(ti.init(ctx, mem, time_idx, T));
  VS_TRY(rp.init(ctx, mem, r_pred_out, (T - 1) * 7));
  VS_TRY(xp.init(ctx, mem, x_pred_out, T * 6));
  VS_TRY(ph.init(ctx, mem, Phi_out, (T - 1) * 36));
  VS_TRY(qg.init(ctx, mem, qgrad_out, T * 3));
  VS_TRY(hd.init(ctx, mem, Hq_diag_out, T * 9));
  VS_TRY(ho.init(ctx, mem, Hq_off_out, (T - 1) * 9));
  VS_CUDA(ctx, gap.alloc(T));
  VS_CUDA(ctx, err.alloc(1));
  VS_CUDA(ctx, cudaMemsetAsync(err.p, 0, sizeof(int32_t), ctx->stream));
  VS_LAUNCH(ctx, F_LAYOUT, k_gaps, ceil_div(T, 256), 256, 0, T, ti.p, gap.p, err.p);
  if (jac || xp.p) {
    VS_CUDA(ctx, drec.alloc(T * VS_DREC));
    VS_TRY(launch_dynamics_stm(ctx, T, nullptr, st.p, gap.p, vel_coeff, mode, drec.p, xp.p, nullptr));
    VS_TRY(launch_quat_terms(ctx, T, st.p, cr.p, gap.p, quat_coeff, drec.p));
    VS_LAUNCH(ctx, F_LAYOUT, k_predict_extract, (unsigned)T, 64, 0, T, drec.p, rp.p, ph.p, qg.p, hd.p, ho.p);
  } else if (rp.p && T > 1) {
    VS_TRY(launch_dyn_trial(ctx, T - 1, nullptr, st.p, cr.p, gap.p, nullptr, nullptr, quat_coeff, vel_coeff, mode,
                            nullptr, rp.p));
  }
  VS_TRY(rp.finish(ctx)); VS_TRY(xp.finish(ctx)); VS_TRY(ph.finish(ctx));
  VS_TRY(qg.finish(ctx)); VS_TRY(hd.finish(ctx)); VS_TRY(ho.finish(ctx));
  int32_t h_err = 0;
  VS_CUDA(ctx, cudaMemcpyAsync(&h_err, err.p, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (h_err) return set_error(ctx, VINSAT_EINVAL, "predict: time_idx must be strictly increasing");
  return VINSAT_OK;
}

int vinsat_propagate_chain(vinsat_ctx* ctx, int mem, int64_t n_steps, double dt, const double* state0,
                           const double* vel0, const double* omega, double* states_out) {
  VS_CHECK_ARG(ctx, ctx != nullptr);
  VS_CHECK_ARG(ctx, n_steps >= 0 && state0 && vel0 && states_out && (n_steps == 0 || omega));
  VS_CUDA(ctx, cudaSetDevice(ctx->device));
  InBuf<double> s0, v0, om;
  OutBuf<double> out;
  VS_TRY(s0.init(ctx, mem, state0, 10));
  VS_TRY(v0.init(ctx, mem, vel0, 3));
  VS_TRY(om.init(ctx, mem, omega, n_steps * 3));
  VS_TRY(out.init(ctx, mem, states_out, (n_steps + 1) * 10));
  VS_TRY(launch_chain(ctx, n_steps, dt, s0.p, v0.p, om.p, out.p));
  VS_TRY(out.finish(ctx));
  VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return VINSAT_OK;
}

int vinsat_orbit_propagate(vinsat_ctx* ctx, int mem, int64_t n_traj, int64_t n_steps, int64_t stride, double h,
                           const double* x0, double* out) {
  VS_CHECK_ARG(ctx, ctx != nullptr);
  VS_CHECK_ARG(ctx, n_traj >= 0 && n_steps >= 0 && stride >= 1 && x0 && out);
  VS_CUDA(ctx, cudaSetDevice(ctx->device));
  InBuf<double> in;
  OutBuf<double> o;
  VS_TRY(in.init(ctx, mem, x0, n_traj * 6));
  VS_TRY(o.init(ctx, mem, out, n_traj * (n_steps / stride + 1) * 6));
  VS_TRY(launch_orbit_propagate(ctx, n_traj, n_steps, stride, h, in.p, o.p));
  VS_TRY(o.finish(ctx));
  VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return VINSAT_OK;
}

int vinsat_cum_rotations(vinsat_ctx* ctx, int mem, int64_t n_full, const double* quat_full, double dt,
                         int64_t n_frames, const int64_t* time_idx, double* omega_out, double* cum_rot_out) {
  VS_CHECK_ARG(ctx, ctx != nullptr);
  VS_CHECK_ARG(ctx, n_full >= 0 && n_frames >= 0 && dt != 0.0 && (n_full == 0 || quat_full));
  VS_CHECK_ARG(ctx, n_frames == 0 || (time_idx && cum_rot_out));
  VS_CUDA(ctx, cudaSetDevice(ctx->device));
  InBuf<double> q;
  InBuf<int64_t> ti;
  OutBuf<double> om, cr;
  DevBuf<double> om_own;
  DevBuf<int32_t> err;
  VS_TRY(q.init(ctx, mem, quat_full, n_full * 4));
  VS_TRY(ti.init(ctx, mem, time_idx, n_frames));
  VS_TRY(om.init(ctx, mem, omega_out, n_full * 3));
  VS_TRY(cr.init(ctx, mem, cum_rot_out, n_frames * 4));
  double* omega = om.p;
  if (!omega) { VS_CUDA(ctx, om_own.alloc(n_full * 3)); omega = om_own.p; }
  VS_CUDA(ctx, err.alloc(1));
  VS_CUDA(ctx, cudaMemsetAsync(err.p, 0, sizeof(int32_t), ctx->stream));
  VS_TRY(launch_omega_from_quat(ctx, n_full, q.p, dt, omega));
  VS_TRY(launch_cum_rot_frames(ctx, n_frames, n_full, ti.p, omega, dt, cr.p, err.p));
  VS_TRY(om.finish(ctx)); VS_TRY(cr.finish(ctx));
  int32_t h_err = 0;
  VS_CUDA(ctx, cudaMemcpyAsync(&h_err, err.p, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (h_err) return set_error(ctx, VINSAT_EINVAL, "cum_rotations: time_idx must be strictly increasing and < n_full");
  return VINSAT_OK;
}

int vinsat_precompute_cum_rotations(vinsat_ctx* ctx, int mem, int64_t n_frames, int64_t n_slots, const double* omegas,
                                    double dt, double* cum_out) {
  VS_CHECK_ARG(ctx, ctx != nullptr);
  VS_CHECK_ARG(ctx, n_frames >= 0 && n_slots >= 0 && (n_frames * n_slots == 0 || (omegas && cum_out)));
  VS_CUDA(ctx, cudaSetDevice(ctx->device));
  InBuf<double> in;
  OutBuf<double> o;
  VS_TRY(in.init(ctx, mem, omegas, n_frames * n_slots * 3));
  VS_TRY(o.init(ctx, mem, cum_out, n_frames * n_slots * 4));
  VS_TRY(launch_cum_rot_prefix(ctx, n_frames, n_slots, in.p, dt, o.p));
  VS_TRY(o.finish(ctx));
  VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return VINSAT_OK;
}

int vinsat_attitude_propagate(vinsat_ctx* ctx, int mem, int64_t n_traj, int64_t n_steps, int64_t stride, double h,
                              const double* inertia_diag, const double* x0, double* out) {
  VS_CHECK_ARG(ctx, ctx != nullptr);
  VS_CHECK_ARG(ctx, n_traj >= 0 && n_steps >= 0 && stride >= 1 && inertia_diag && x0 && out);
  VS_CHECK_ARG(ctx, inertia_diag[0] > 0 && inertia_diag[1] > 0 && inertia_diag[2] > 0);
  VS_CUDA(ctx, cudaSetDevice(ctx->device));
  InBuf<double> in;
  OutBuf<double> o;
  VS_TRY(in.init(ctx, mem, x0, n_traj * 7));
  VS_TRY(o.init(ctx, mem, out, n_traj * (n_steps / stride + 1) * 7));
  VS_TRY(launch_attitude_propagate(ctx, n_traj, n_steps, stride, h, inertia_diag, in.p, o.p));
  VS_TRY(o.finish(ctx));
  VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return VINSAT_OK;
}

}  // extern "C"
