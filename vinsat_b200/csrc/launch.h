// Kernel launchers (device pointers; all run on ctx->stream).  Internal.
#pragma once
#include "batch.h"

namespace vs {

// ---- kernels_obs.cu ---------------------------------------------------------------------------------
// Stand-alone a1 on reference-layout (AoS) device buffers.
int launch_project_aos(vinsat_ctx* ctx, int64_t T, int64_t M, const double* states, const double* intr,
                       const double* xyz, const int64_t* ii, double* uv_out, double* Jg_out, int32_t* err_flag);
// Headline: residual + Jacobian for resident SoA observations -> r[2][M], J[12][M].
int launch_resjac(vinsat_batch* b);
// r = uv - project(st) for every observation (input of the robust scale).
int launch_obs_residual(vinsat_batch* b);
// Fused projection + Jacobian + robust weight + per-frame JtWJ / JtWr (+ sum|r|, max w).
int launch_obs_assemble(vinsat_batch* b, double alpha);
// Trial: per-frame sum of wu*|uv - project(st_new)|.
int launch_obs_trial(vinsat_batch* b);
// Observation indexing on the device: oframe, obs_start (CSR), sortedness check.
int launch_obs_index(vinsat_batch* b, const int64_t* d_ii_local);
// AoS [n][ncol] <-> SoA [ncol][n] through shared memory.
int launch_aos_to_soa(vinsat_ctx* ctx, const double* aos, double* soa, int64_t n, int ncol);
int launch_soa_to_aos(vinsat_ctx* ctx, const double* soa, double* aos, int64_t n, int ncol);

// ---- kernels_dyn.cu ---------------------------------------------------------------------------------
// RK4 + STM for the pairs listed in `order` (2 threads per pair): writes Phi, r6 into drec; x_pred optional.
int launch_dynamics_stm(vinsat_ctx* ctx, int64_t n_pairs, const int32_t* order, const double* st,
                        const int32_t* gap, double vel_coeff, int mode, double* drec, double* x_pred, double* mrec,
                        const int32_t* gate = nullptr);
// Quaternion smoothness terms per frame: rho, qgrad, Hq_diag, Hq_off into drec.
int launch_quat_terms(vinsat_ctx* ctx, int64_t T, const double* st, const double* crot, const int32_t* gap,
                      double quat_coeff, double* drec, const int32_t* gate = nullptr);
// Trial: residual-only propagation; e_dyn[f] = sum |r_pred(f)| over the 7 components (0 where no pair).
int launch_dyn_trial(vinsat_ctx* ctx, int64_t n_pairs, const int32_t* order, const double* st,
                     const double* crot, const int32_t* gap, const int32_t* active, const int32_t* fprob,
                     double quat_coeff, double vel_coeff, int mode, double* e_dyn, double* r7_out,
                     const int32_t* gate = nullptr);
int launch_chain(vinsat_ctx* ctx, int64_t n_steps, double dt, const double* state0, const double* vel0,
                 const double* omega, double* states_out);
int launch_orbit_propagate(vinsat_ctx* ctx, int64_t n_traj, int64_t n_steps, int64_t stride, double h,
                           const double* x0, double* out);

// ---- kernels_solve.cu -------------------------------------------------------------------------------
int launch_select_median(vinsat_batch* b);                       // c_obs[p] = lower median of |r|
int launch_select_begin(vinsat_batch* b, int64_t global_values);  // -1: per-problem counts; else global count (long arc)
int launch_select_hist(vinsat_batch* b, int pass);               // histogram of digit `pass` into sel_hist
int launch_select_pick(vinsat_batch* b, int pass);               // narrow the prefix from sel_hist (and zero it)
int launch_system_build(vinsat_batch* b, int initialize, double Sigma, double vel_coeff);
int launch_init_residual(vinsat_batch* b, int initialize, double Sigma, double lamda_host_default,
                         const double* d_lam_in, const double* e_prior = nullptr);
int launch_solve_retract(vinsat_batch* b, int initialize);
int launch_solve_init_only(vinsat_batch* b);
int launch_retract_only(vinsat_batch* b);
int launch_accept(vinsat_batch* b, int initialize, double Sigma, const double* e_prior = nullptr);
// stage 1 of the long-arc sums: part[chunk][3]; mode 0 linearisation / 1 trial; hi > lo: owned frames of a window batch
int launch_sum_partials(vinsat_batch* b, int mode, int initialize, const double* e_prior, int64_t lo, int64_t hi);
// ---- prior.cu (BA_reg) ----
int launch_prior_linearize(vinsat_batch* b, double vc, double qc);   // srec += prior blocks; e_pr_init[f] = sum |r_prior|
int launch_prior_trial(vinsat_batch* b, double vc, double qc);       // e_pr[f] = sum |r_prior(st_new)|
int launch_gather_last_hessian(vinsat_batch* b, double* out_dev);    // out[P][81] = JTwJ[-9:, -9:] incl. damping
int launch_stream_gather(vinsat_ctx* ctx, int64_t n, const double* chain, const int64_t* time_idx, int64_t f0, int64_t t0,
                         double* st, double* vel, double* seed);

// ---- kernels_chain.cu -------------------------------------------------------------------------------
int launch_chain_solve(vinsat_batch* b);   // delta = A^-1 rhs for every active problem (partitioned block LU)
// pieces of the partitioned solve, for the frame-window sharded long arc (longarc.cu)
int launch_seg_forward(vinsat_batch* b);                                    // segments: forward + back-recurrence -> redrec
int launch_seg_backsub(vinsat_batch* b);                                    // interiors from separator solutions in delta
int launch_reduced_packed(vinsat_batch* b, int64_t S_total, const double* pack, double* rsys, double* rlow,
                          double* rwrec, double* xsep, const int32_t* one_chain /* {0, S_total, 0} on device */);

// ---- input preparation / simulation (kernels_dyn.cu) --------------------------------------------------
int launch_omega_from_quat(vinsat_ctx* ctx, int64_t n, const double* quat, double dt, double* omega);
int launch_cum_rot_frames(vinsat_ctx* ctx, int64_t T, int64_t n_full, const int64_t* time_idx, const double* omega,
                          double dt, double* cum_rot, int32_t* err);
int launch_cum_rot_prefix(vinsat_ctx* ctx, int64_t T, int64_t N, const double* omegas, double dt, double* cum);
int launch_attitude_propagate(vinsat_ctx* ctx, int64_t n_traj, int64_t n_steps, int64_t stride, double h,
                              const double* inertia, const double* x0, double* out);

// ---- speculation gate (kernels_solve.cu; see vinsat_batch_od_solve) ---------------------------------------
int launch_gated_zero(vinsat_batch* b, unsigned long long* p64, int64_t n64, int32_t* p32, int n32);
int launch_gate_publish(vinsat_batch* b);

}  // namespace vs
