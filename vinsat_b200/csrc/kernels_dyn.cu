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
                                                      double* __restrict__ mrec) {
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
                        const int32_t* gap, double vel_coeff, int mode, double* drec, double* x_pred, double* mrec) {
  if (n_pairs == 0) return VINSAT_OK;
  VS_LAUNCH(ctx, F_DYNAMICS, k_dynamics_stm, ceil_div(n_pairs * 2, 128), 128, 0, n_pairs, order, st, gap, vel_coeff,
            mode, drec, x_pred, mrec);
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

__global__ void __launch_bounds__(128) k_quat_terms(int64_t T, const double* __restrict__ st,
                                                    const double* __restrict__ crot,
                                                    const int32_t* __restrict__ gap, double c,
                                                    double* __restrict__ drec) {
  const int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (f >= T) return;
  const bool has_next = gap[f] > 0;
  const bool has_prev = f > 0 && gap[f - 1] > 0;
  const Quat q = load_q(st + f * 10);
  const Quat qc = qconj(q);
  double ax = 0, ay = 0, az = 0, aw = 0, rho = 0;
  double* d = drec + f * VS_DREC;
  if (has_next) {
    const Quat qn = load_q(st + (f + 1) * 10);
    const Quat R = load_q4(crot + f * 4);
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
    const Quat qp = load_q(st + (f - 1) * 10);
    const Quat Rp = load_q4(crot + (f - 1) * 4);
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

int launch_quat_terms(vinsat_ctx* ctx, int64_t T, const double* st, const double* crot, const int32_t* gap,
                      double quat_coeff, double* drec) {
  if (T == 0) return VINSAT_OK;
  VS_LAUNCH(ctx, F_QUAT, k_quat_terms, ceil_div(T, 128), 128, 0, T, st, crot, gap, quat_coeff, drec);
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
                                                   double* __restrict__ e_dyn, double* __restrict__ r7_out) {
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
  for (int k = 0; k < nh; k++) rk4_step(x, hop_size(g, mode, k));
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
                     double quat_coeff, double vel_coeff, int mode, double* e_dyn, double* r7_out) {
  if (n_pairs == 0) return VINSAT_OK;
  VS_LAUNCH(ctx, F_TRIAL, k_dyn_trial, ceil_div(n_pairs, 128), 128, 0, n_pairs, order, st, crot, gap, active, fprob,
            quat_coeff, vel_coeff, mode, e_dyn, r7_out);
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

}  // namespace vs
