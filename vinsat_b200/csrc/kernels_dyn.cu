// Dynamics-side kernels (a2, a3, a4, a8, a11 of SURVEY.md section 8): RK4 orbit propagation with its exact
// discrete state-transition matrix, quaternion-smoothness gradient / Hessian blocks, trial residuals,
// and the sequential chains used to seed streaming windows / simulate orbits.
//
// FP64-pipe bound: ~1.2 kflop per RK4+STM step, a few hundred bytes per frame pair.  Pairs are processed
// longest-gap-first (`order`) so that the lanes of a warp run the same number of steps.
#include "common.cuh"
#include "launch.h"

namespace vs {

// ---------------------------------------------------------------------------------------------------------
// RK4 + STM.  Two threads per frame pair, three STM columns each (keeps the live set under ~170 registers;
// both threads integrate the 6-state redundantly, which is bit-identical and needs no exchange).
// drec[f]: Phi row-major [0,36) | r6 [36,42)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_dynamics_stm(int64_t n_pairs, const int32_t* __restrict__ order,
                                                      const double* __restrict__ st,
                                                      const int32_t* __restrict__ gap, double vel_coeff, int mode,
                                                      double* __restrict__ drec, double* __restrict__ x_pred,
                                                      double* __restrict__ mrec, const int32_t* __restrict__ gate) {
  if (gate && *gate) return;      // speculative launch behind an LM loop that is not finished (batch.cu)
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int half = (int)(t & 1);
  const bool live = (t >> 1) < n_pairs;
  const int64_t j = live ? (t >> 1) : n_pairs - 1;      // dead tail threads shadow the last pair (no early exit:
  const int64_t f = order ? order[j] : j;               // the lane pair exchanges its columns with shuffles below)
  const double* s = st + f * 10;
  double x[6] = {s[0], s[1], s[2], s[7], s[8], s[9]};
  double phi[3][6];
#pragma unroll
  for (int c = 0; c < 3; c++)
#pragma unroll
    for (int k = 0; k < 6; k++) phi[c][k] = (k == half * 3 + c) ? 1.0 : 0.0;
  int g = gap[f];
  const bool has_next = g > 0;
  if (!has_next) g = 1;           // BA_utils.py:75: the last frame gets a dummy gap of 1, result discarded
  const int nh = num_hops(g, mode);
  for (int k = 0; k < nh; k++) rk4_step_stm<3>(x, phi, hop_size(g, mode, k));
  double* d = drec + f * VS_DREC;
  if (live) {
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
      for (int k = 0; k < 6; k++) d[k * 6 + half * 3 + c] = phi[c][k];
  }
  double r6[6] = {0, 0, 0, 0, 0, 0};
  if (has_next) {
    const double* sn = s + 10;
    r6[0] = x[0] - sn[0]; r6[1] = x[1] - sn[1]; r6[2] = x[2] - sn[2];
    r6[3] = (x[3] - sn[7]) * vel_coeff; r6[4] = (x[4] - sn[8]) * vel_coeff; r6[5] = (x[5] - sn[9]) * vel_coeff;
  }
  if (live && half == 0) {
    if (x_pred) {
#pragma unroll
      for (int k = 0; k < 6; k++) x_pred[f * 6 + k] = x[k];
    }
#pragma unroll
    for (int k = 0; k < 6; k++) d[36 + k] = r6[k];
  }
  if (mrec) {
    // Phi^T D^2 Phi (upper triangle) and Phi^T D r for the normal equations: the partner's three columns come
    // over with shuffles; this thread produces the entries whose FIRST column it owns.
    double other[3][6];
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
      for (int k = 0; k < 6; k++) other[c][k] = __shfl_xor_sync(0xffffffffu, phi[c][k], 1);
    const double dv1[6] = {1.0, 1.0, 1.0, vel_coeff, vel_coeff, vel_coeff};
    double* m = mrec + f * VS_MREC;
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const int a = half * 3 + c;
      double wcol[6];                                   // D^2 Phi[:, a]
#pragma unroll
      for (int k = 0; k < 6; k++) wcol[k] = dv1[k] * dv1[k] * phi[c][k];
      // entries (a, b) with b >= a: own columns b = a .. 3*half+2, and (half 0 only) the partner's columns 3..5;
      // every entry is stored at both (a, b) and (b, a) so the 6x6 block is exactly symmetric
#pragma unroll
      for (int c2 = 0; c2 < 3; c2++) {
        if (c2 >= c) {
          double sacc = 0.0;
#pragma unroll
          for (int k = 0; k < 6; k++) sacc = fma(wcol[k], phi[c2][k], sacc);
          const int bcol = half * 3 + c2;
          if (live) { m[a * 6 + bcol] = sacc; m[bcol * 6 + a] = sacc; }
        }
      }
      if (half == 0) {
#pragma unroll
        for (int c2 = 0; c2 < 3; c2++) {
          double sacc = 0.0;
#pragma unroll
          for (int k = 0; k < 6; k++) sacc = fma(wcol[k], other[c2][k], sacc);
          const int bcol = 3 + c2;
          if (live) { m[a * 6 + bcol] = sacc; m[bcol * 6 + a] = sacc; }
        }
      }
      double vacc = 0.0;
#pragma unroll
      for (int k = 0; k < 6; k++) vacc = fma(phi[c][k] * dv1[k], r6[k], vacc);
      if (live) m[36 + a] = vacc;
    }
  }
}

int launch_dynamics_stm(vinsat_ctx* ctx, int64_t n_pairs, const int32_t* order, const double* st,
                        const int32_t* gap, double vel_coeff, int mode, double* drec, double* x_pred, double* mrec,
                        const int32_t* gate) {
  if (n_pairs == 0) return VINSAT_OK;
  static const int th = getenv("VINSAT_DYN_THREADS") ? atoi(getenv("VINSAT_DYN_THREADS")) : 32;   // 198 registers: 32-thread CTAs pack 10 warps per SM (128-thread CTAs: 8)
  VS_LAUNCH(ctx, F_DYNAMICS, k_dynamics_stm, ceil_div(n_pairs * 2, th), th, 0, n_pairs, order, st, gap, vel_coeff,
            mode, drec, x_pred, mrec, gate);
  return VINSAT_OK;
}

// ---------------------------------------------------------------------------------------------------------
// quaternion smoothness (BA_utils.py:478,484-495,519-523 via closed forms)
// drec[f]: rho [42] | qgrad [43,46) | Hq_diag [46,55) | Hq_off [55,64)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ Quat qconj(const Quat& q) { return Quat{-q.x, -q.y, -q.z, q.w}; }
__device__ __forceinline__ double qdot(const Quat& a, const Quat& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ double sgn(double d) { return (double)((d > 0.0) - (d < 0.0)); }
__device__ __forceinline__ Quat load_q(const double* s) { return Quat{s[3], s[4], s[5], s[6]}; }
__device__ __forceinline__ Quat load_q4(const double* s) { return Quat{s[0], s[1], s[2], s[3]}; }

// One thread per frame; the state rows / cumulative rotations of the CTA's 128 frames (+ one halo row each side)
// are staged with coalesced loads and the 22 outputs per frame leave through a shared tile as contiguous runs
// (the direct version issued 22 stores + 14 loads per thread that each touched 32 sectors: ncu lg_throttle).
constexpr int kQtFrames = 128;
constexpr int kQtOut = 23;      // 22 outputs, odd pitch

__global__ void __launch_bounds__(kQtFrames) k_quat_terms(int64_t T, const double* __restrict__ st,
                                                          const double* __restrict__ crot,
                                                          const int32_t* __restrict__ gap, double c,
                                                          double* __restrict__ drec, const int32_t* __restrict__ gate) {
  if (gate && *gate) return;      // speculative launch behind an LM loop that is not finished (batch.cu)
  __shared__ double s_q[(kQtFrames + 2) * 4];      // quaternion of frames f0-1 .. f0+128
  __shared__ double s_r[(kQtFrames + 1) * 4];      // cum rotation of frames f0-1 .. f0+127
  __shared__ int32_t s_gap[kQtFrames + 1];         // gap of frames f0-1 .. f0+127
  __shared__ double s_out[kQtFrames * kQtOut];
  const int tid = threadIdx.x;
  const int64_t f0 = (int64_t)blockIdx.x * kQtFrames;
  const int nf = (int)min((int64_t)kQtFrames, T - f0);
  for (int i = tid; i < (kQtFrames + 2) * 4; i += kQtFrames) {
    const int64_t f = f0 - 1 + i / 4;
    s_q[i] = (f >= 0 && f < T) ? st[f * 10 + 3 + (i & 3)] : 0.0;
  }
  for (int i = tid; i < (kQtFrames + 1) * 4; i += kQtFrames) {
    const int64_t f = f0 - 1 + i / 4;
    s_r[i] = (f >= 0 && f < T) ? crot[f * 4 + (i & 3)] : 0.0;
  }
  for (int i = tid; i < kQtFrames + 1; i += kQtFrames) {
    const int64_t f = f0 - 1 + i;
    s_gap[i] = (f >= 0 && f < T) ? gap[f] : 0;
  }
  __syncthreads();
  if (tid < nf) {
    const bool has_next = s_gap[tid + 1] > 0;
    const bool has_prev = s_gap[tid] > 0;
    const Quat q = load_q4(s_q + (tid + 1) * 4);
    const Quat qc = qconj(q);
    double ax = 0, ay = 0, az = 0, aw = 0, rho = 0;
    double* d = s_out + tid * kQtOut - 42;           // d[42..63] as in the record
    if (has_next) {
      const Quat qn = load_q4(s_q + (tid + 2) * 4);
      const Quat R = load_q4(s_r + (tid + 1) * 4);
      const Quat Rc = qconj(R);
      const double dt = qdot(qmul(q, R), qn);
      const double s = sgn(dt);
      rho = c * (1.0 - fabs(dt));
      const Quat m = qmul(qn, Rc);                    // M(R)^T q_{i+1}
      ax += -c * s * m.x; ay += -c * s * m.y; az += -c * s * m.z; aw += -c * s * m.w;
      // Hq[i,i+1] column j = -c s vec(conj(q) (x) ((q_next (x) e_j) (x) conj(R)))
      const Quat e[3] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}};
#pragma unroll
      for (int jx = 0; jx < 3; jx++) {
        const Quat v = qmul(qc, qmul(qmul(qn, e[jx]), Rc));
        d[55 + 0 * 3 + jx] = -c * s * v.x;
        d[55 + 1 * 3 + jx] = -c * s * v.y;
        d[55 + 2 * 3 + jx] = -c * s * v.z;
      }
    } else {
#pragma unroll
      for (int k = 55; k < 64; k++) d[k] = 0.0;
    }
    if (has_prev) {
      const Quat qp = load_q4(s_q + tid * 4);
      const Quat Rp = load_q4(s_r + tid * 4);
      const Quat pred = qmul(qp, Rp);                 // M_{i-1} q_{i-1}
      const double s = sgn(qdot(pred, q));
      ax += -c * s * pred.x; ay += -c * s * pred.y; az += -c * s * pred.z; aw += -c * s * pred.w;
    }
    const Quat a = {ax, ay, az, aw};
    const Quat gq = qmul(qc, a);                      // Gq(q)^T a = vec(conj(q) (x) a)
    const double beta = -qdot(q, a);
    d[42] = rho;
    d[43] = gq.x; d[44] = gq.y; d[45] = gq.z;
    // beta I + hat(g)
    d[46] = beta;  d[47] = -gq.z; d[48] = gq.y;
    d[49] = gq.z;  d[50] = beta;  d[51] = -gq.x;
    d[52] = -gq.y; d[53] = gq.x;  d[54] = beta;
  }
  __syncthreads();
  for (int i = tid; i < nf * 22; i += kQtFrames) {
    const int lf = i / 22, e = i - lf * 22;
    drec[(f0 + lf) * VS_DREC + 42 + e] = s_out[lf * kQtOut + e];
  }
}

int launch_quat_terms(vinsat_ctx* ctx, int64_t T, const double* st, const double* crot, const int32_t* gap,
                      double quat_coeff, double* drec, const int32_t* gate) {
  if (T == 0) return VINSAT_OK;
  VS_LAUNCH(ctx, F_QUAT, k_quat_terms, ceil_div(T, 128), 128, 0, T, st, crot, gap, quat_coeff, drec, gate);
  return VINSAT_OK;
}

// ---------------------------------------------------------------------------------------------------------
// trial residual, dynamics part (BA_filtering.py:65,67): one thread per pair, state-only RK4
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_dyn_trial(int64_t n_pairs, const int32_t* __restrict__ order,
                                                   const double* __restrict__ st, const double* __restrict__ crot,
                                                   const int32_t* __restrict__ gap,
                                                   const int32_t* __restrict__ active,
                                                   const int32_t* __restrict__ fprob, double qc, double vc, int mode,
                                                   double* __restrict__ e_dyn, double* __restrict__ r7_out,
                                                   const int32_t* __restrict__ gate) {
  if (gate && *gate) return;      // speculative launch behind an LM loop that is not finished (batch.cu)
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= n_pairs) return;
  const int64_t f = order ? order[j] : j;
  if (active && !active[fprob[f]]) return;
  const int g = gap[f];
  if (g <= 0) {
    if (e_dyn) e_dyn[f] = 0.0;
    return;
  }
  const double* s = st + f * 10;
  const double* sn = s + 10;
  double x[6] = {s[0], s[1], s[2], s[7], s[8], s[9]};
  const int nh = num_hops(g, mode);
  for (int k = 0; k < nh; k++) rk4_step<true>(x, hop_size(g, mode, k));
  const double r0 = x[0] - sn[0], r1 = x[1] - sn[1], r2 = x[2] - sn[2];
  const double r3 = (x[3] - sn[7]) * vc, r4 = (x[4] - sn[8]) * vc, r5 = (x[5] - sn[9]) * vc;
  const double dt = qdot(qmul(load_q(s), load_q4(crot + f * 4)), load_q(sn));
  const double r6 = qc * (1.0 - fabs(dt));
  if (e_dyn) e_dyn[f] = fabs(r0) + fabs(r1) + fabs(r2) + fabs(r3) + fabs(r4) + fabs(r5) + fabs(r6);
  if (r7_out) {
    double* o = r7_out + f * 7;
    o[0] = r0; o[1] = r1; o[2] = r2; o[3] = r3; o[4] = r4; o[5] = r5; o[6] = r6;
  }
}

int launch_dyn_trial(vinsat_ctx* ctx, int64_t n_pairs, const int32_t* order, const double* st,
                     const double* crot, const int32_t* gap, const int32_t* active, const int32_t* fprob,
                     double quat_coeff, double vel_coeff, int mode, double* e_dyn, double* r7_out,
                     const int32_t* gate) {
  if (n_pairs == 0) return VINSAT_OK;
  VS_LAUNCH(ctx, F_TRIAL, k_dyn_trial, ceil_div(n_pairs, 128), 128, 0, n_pairs, order, st, crot, gap, active, fprob,
            quat_coeff, vel_coeff, mode, e_dyn, r7_out, gate);
  return VINSAT_OK;
}

// ---------------------------------------------------------------------------------------------------------
// a8: one state chained forward (BA_utils.py:89-112).  Thread 0 integrates the orbit, thread 1 the attitude.
// ---------------------------------------------------------------------------------------------------------
__global__ void k_chain(int64_t n_steps, double dt, const double* __restrict__ state0,
                        const double* __restrict__ vel0, const double* __restrict__ omega,
                        double* __restrict__ out) {
  if (threadIdx.x == 0) {
    double x[6] = {state0[0], state0[1], state0[2], vel0[0], vel0[1], vel0[2]};
    for (int64_t k = 0;; k++) {
      double* o = out + k * 10;
      o[0] = x[0]; o[1] = x[1]; o[2] = x[2]; o[7] = x[3]; o[8] = x[4]; o[9] = x[5];
      if (k == n_steps) break;
      rk4_step(x, dt);
    }
  } else if (threadIdx.x == 1) {
    Quat q = {state0[3], state0[4], state0[5], state0[6]};
    for (int64_t k = 0;; k++) {
      double* o = out + k * 10;
      o[3] = q.x; o[4] = q.y; o[5] = q.z; o[6] = q.w;
      if (k == n_steps) break;
      q = qmul(q, qexp(dt * omega[k * 3], dt * omega[k * 3 + 1], dt * omega[k * 3 + 2]));
    }
  }
}

int launch_chain(vinsat_ctx* ctx, int64_t n_steps, double dt, const double* state0, const double* vel0,
                 const double* omega, double* states_out) {
  VS_LAUNCH(ctx, F_SIM, k_chain, 1, 32, 0, n_steps, dt, state0, vel0, omega, states_out);
  return VINSAT_OK;
}

// a11: independent trajectories, one thread each; every `stride`-th state is stored.
__global__ void __launch_bounds__(128) k_orbit_propagate(int64_t n_traj, int64_t n_steps, int64_t stride, double h,
                                                         const double* __restrict__ x0, double* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n_traj) return;
  const int64_t n_out = n_steps / stride + 1;
  double x[6];
#pragma unroll
  for (int k = 0; k < 6; k++) x[k] = x0[i * 6 + k];
  double* o = out + i * n_out * 6;
  for (int64_t s = 0;; s++) {
    if (s % stride == 0) {
#pragma unroll
      for (int k = 0; k < 6; k++) o[(s / stride) * 6 + k] = x[k];
    }
    if (s == n_steps) break;
    rk4_step(x, h);
  }
}

int launch_orbit_propagate(vinsat_ctx* ctx, int64_t n_traj, int64_t n_steps, int64_t stride, double h,
                           const double* x0, double* out) {
  if (n_traj == 0) return VINSAT_OK;
  VS_LAUNCH(ctx, F_SIM, k_orbit_propagate, ceil_div(n_traj, 128), 128, 0, n_traj, n_steps, stride, h, x0, out);
  return VINSAT_OK;
}


// ---------------------------------------------------------------------------------------------------------
// Input preparation on the device
// ---------------------------------------------------------------------------------------------------------
// BA_utils.py:962-968: q <- clip(q/|q|, -1, 1); theta = 2 acos(w); log = xyz / sin(theta/2) * theta   (0/0 -> NaN, as torch)
__device__ __forceinline__ void qlog(const Quat& q, double& lx, double& ly, double& lz) {
  const double n = qnorm_exact(q);
  auto clip = [](double v) { return v < -1.0 ? -1.0 : (v > 1.0 ? 1.0 : v); };
  const double x = clip(q.x / n), y = clip(q.y / n), z = clip(q.z / n), w = clip(q.w / n);
  const double theta = 2.0 * acos(w);
  const double sh = sin(0.5 * theta);
  lx = (x / sh) * theta; ly = (y / sh) * theta; lz = (z / sh) * theta;
}

// compute_omega_from_quat (BA_utils.py:1361-1367): omega_s = log(normalize(conj(q_s) (x) q_{s+1})) / dt, last row 0
__global__ void __launch_bounds__(256) k_omega_from_quat(int64_t n, const double* __restrict__ quat, double dt,
                                                         double* __restrict__ omega) {
  const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (s >= n) return;
  double ox = 0.0, oy = 0.0, oz = 0.0;
  if (s + 1 < n) {
    const Quat a = load_q4(quat + s * 4), b = load_q4(quat + (s + 1) * 4);
    Quat dq = qmul_exact(qconj(a), b);
    const double nn = qnorm_exact(dq);
    dq.x /= nn; dq.y /= nn; dq.z /= nn; dq.w /= nn;
    qlog(dq, ox, oy, oz);
    ox /= dt; oy /= dt; oz /= dt;
  }
  omega[s * 3] = ox; omega[s * 3 + 1] = oy; omega[s * 3 + 2] = oz;
}

// precompute_cum_rotations (BA_utils.py:278-288) restricted to the slice `predict` reads (cum_rotations[:, :, -1],
// :295): frame i gets the ordered product of exp(dt*omega_s), s = time_idx[i] .. time_idx[i+1]-1; the reference's
// zero padding multiplies by the identity quaternion (exact), the last frame is the identity.
__global__ void __launch_bounds__(128) k_cum_rot_frames(int64_t T, int64_t n_full, const int64_t* __restrict__ time_idx,
                                                        const double* __restrict__ omega, double dt,
                                                        double* __restrict__ cum_rot, int32_t* __restrict__ err) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= T) return;
  Quat c = {0.0, 0.0, 0.0, 1.0};
  if (i + 1 < T) {
    const int64_t s0 = time_idx[i], s1 = time_idx[i + 1];
    if (s0 < 0 || s1 <= s0 || s1 > n_full) { atomicOr(err, 1); }
    else {
      for (int64_t s = s0; s < s1; s++) {
        const Quat r = qexp(dt * omega[s * 3], dt * omega[s * 3 + 1], dt * omega[s * 3 + 2]);
        c = (s == s0) ? r : qmul_exact(c, r);
      }
    }
  }
  cum_rot[i * 4] = c.x; cum_rot[i * 4 + 1] = c.y; cum_rot[i * 4 + 2] = c.z; cum_rot[i * 4 + 3] = c.w;
}

// the general form (all prefixes): omegas [T, N, 3] -> cum [T, N, 4]; one thread per frame
__global__ void __launch_bounds__(128) k_cum_rot_prefix(int64_t T, int64_t N, const double* __restrict__ omegas,
                                                        double dt, double* __restrict__ cum) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= T) return;
  Quat c = {0.0, 0.0, 0.0, 1.0};
  for (int64_t j = 0; j < N; j++) {
    const double* o = omegas + (i * N + j) * 3;
    const Quat r = qexp(dt * o[0], dt * o[1], dt * o[2]);
    c = (j == 0) ? r : qmul_exact(c, r);
    double* d = cum + (i * N + j) * 4;
    d[0] = c.x; d[1] = c.y; d[2] = c.z; d[3] = c.w;
  }
}

int launch_omega_from_quat(vinsat_ctx* ctx, int64_t n, const double* quat, double dt, double* omega) {
  if (n == 0) return VINSAT_OK;
  VS_LAUNCH(ctx, F_SIM, k_omega_from_quat, ceil_div(n, 256), 256, 0, n, quat, dt, omega);
  return VINSAT_OK;
}
int launch_cum_rot_frames(vinsat_ctx* ctx, int64_t T, int64_t n_full, const int64_t* time_idx, const double* omega,
                          double dt, double* cum_rot, int32_t* err) {
  if (T == 0) return VINSAT_OK;
  VS_LAUNCH(ctx, F_SIM, k_cum_rot_frames, ceil_div(T, 128), 128, 0, T, n_full, time_idx, omega, dt, cum_rot, err);
  return VINSAT_OK;
}
int launch_cum_rot_prefix(vinsat_ctx* ctx, int64_t T, int64_t N, const double* omegas, double dt, double* cum) {
  if (T == 0 || N == 0) return VINSAT_OK;
  VS_LAUNCH(ctx, F_SIM, k_cum_rot_prefix, ceil_div(T, 128), 128, 0, T, N, omegas, dt, cum);
  return VINSAT_OK;
}

// Rigid-body attitude simulation (trajgen_pipe.py:155-207, scalar-FIRST quaternion [s, v], state [q, omega]):
//   q_dot = 1/2 L(q) H omega,  omega_dot = -J^-1 (omega x J omega),  J = diag(jx, jy, jz) (trajgen_pipe.py:185),
// classic RK4, then q normalised (:205).  attitude_dynamics normalises ITS ARGUMENT in place (:187): the stage
// arguments are temporaries, but the first call normalises the step's own state before the later stages use it.
struct Att { double s, x, y, z, wx, wy, wz; };
__device__ __forceinline__ Att att_axpy(const Att& a, double h, const Att& f) {
  return Att{a.s + h * f.s, a.x + h * f.x, a.y + h * f.y, a.z + h * f.z, a.wx + h * f.wx, a.wy + h * f.wy, a.wz + h * f.wz};
}
__device__ __forceinline__ Att att_dyn(Att& a, double jx, double jy, double jz) {
  const double n = sqrt(a.s * a.s + a.x * a.x + a.y * a.y + a.z * a.z);
  a.s /= n; a.x /= n; a.y /= n; a.z /= n;
  Att f;
  f.s = 0.5 * (-a.x * a.wx - a.y * a.wy - a.z * a.wz);
  f.x = 0.5 * (a.s * a.wx - a.z * a.wy + a.y * a.wz);
  f.y = 0.5 * (a.z * a.wx + a.s * a.wy - a.x * a.wz);
  f.z = 0.5 * (-a.y * a.wx + a.x * a.wy + a.s * a.wz);
  const double hx = jx * a.wx, hy = jy * a.wy, hz = jz * a.wz;      // J omega
  f.wx = -(a.wy * hz - a.wz * hy) / jx;
  f.wy = -(a.wz * hx - a.wx * hz) / jy;
  f.wz = -(a.wx * hy - a.wy * hx) / jz;
  return f;
}

__global__ void __launch_bounds__(128) k_attitude_propagate(int64_t n_traj, int64_t n_steps, int64_t stride, double h,
                                                            double jx, double jy, double jz,
                                                            const double* __restrict__ x0, double* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n_traj) return;
  const int64_t n_out = n_steps / stride + 1;
  const double* a0 = x0 + i * 7;
  Att a = {a0[0], a0[1], a0[2], a0[3], a0[4], a0[5], a0[6]};
  double* o = out + i * n_out * 7;
  for (int64_t s = 0;; s++) {
    if (s % stride == 0) {
      double* d = o + (s / stride) * 7;
      d[0] = a.s; d[1] = a.x; d[2] = a.y; d[3] = a.z; d[4] = a.wx; d[5] = a.wy; d[6] = a.wz;
    }
    if (s == n_steps) break;
    const Att f1 = att_dyn(a, jx, jy, jz);                 // normalises a (see above)
    Att t = att_axpy(a, 0.5 * h, f1);
    const Att f2 = att_dyn(t, jx, jy, jz);
    t = att_axpy(a, 0.5 * h, f2);
    const Att f3 = att_dyn(t, jx, jy, jz);
    t = att_axpy(a, h, f3);
    const Att f4 = att_dyn(t, jx, jy, jz);
    const double h6 = h / 6.0;
    a.s += h6 * (f1.s + 2 * f2.s + 2 * f3.s + f4.s);
    a.x += h6 * (f1.x + 2 * f2.x + 2 * f3.x + f4.x);
    a.y += h6 * (f1.y + 2 * f2.y + 2 * f3.y + f4.y);
    a.z += h6 * (f1.z + 2 * f2.z + 2 * f3.z + f4.z);
    a.wx += h6 * (f1.wx + 2 * f2.wx + 2 * f3.wx + f4.wx);
    a.wy += h6 * (f1.wy + 2 * f2.wy + 2 * f3.wy + f4.wy);
    a.wz += h6 * (f1.wz + 2 * f2.wz + 2 * f3.wz + f4.wz);
    const double n = sqrt(a.s * a.s + a.x * a.x + a.y * a.y + a.z * a.z);
    a.s /= n; a.x /= n; a.y /= n; a.z /= n;
  }
}

int launch_attitude_propagate(vinsat_ctx* ctx, int64_t n_traj, int64_t n_steps, int64_t stride, double h,
                              const double* inertia, const double* x0, double* out) {
  if (n_traj == 0) return VINSAT_OK;
  VS_LAUNCH(ctx, F_SIM, k_attitude_propagate, ceil_div(n_traj, 128), 128, 0, n_traj, n_steps, stride, h, inertia[0],
            inertia[1], inertia[2], x0, out);
  return VINSAT_OK;
}

}  // namespace vs
