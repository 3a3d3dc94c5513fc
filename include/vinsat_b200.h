/*
 * vinsat_b200 -- C ABI of the B200-native VINSat estimation hot path.
 *
 * The reference (CMUAbstract/VINSat) is pure Python and has no FFI / plugin interface; its
 * boundary for this path is the set of Python call signatures that estimation/od_pipe.py uses
 * (SURVEY.md section 8(b)).  Every entry point below names the reference function(s) it replaces,
 * as path:line under <reference>/estimation.  The Python mirror in vinsat_b200/ binds these with
 * ctypes (see INTEGRATION.md for the stub a maintainer would add to the reference).
 *
 * Conventions
 *   - fp64 everywhere; km, km/s, s, pixels; quaternions xyzw.
 *   - state row   = [p(3) q(4) v(3)]              (reference `states`, (1,T,10))
 *   - tangent row = [dp(3) dtheta(3) dv(3)]        (BA_filtering.py:56-60)
 *   - all functions return 0 on success, a negative VINSAT_E* code on failure; the message is
 *     available from vinsat_last_error().  No exceptions cross the boundary.
 *   - `mem` says where the caller's buffers live: VINSAT_MEM_HOST (pageable or pinned host memory;
 *     the library stages them through device memory) or VINSAT_MEM_DEVICE (CUDA device pointers on
 *     the context's device; no copies).
 *   - plain pointers and sizes only; no torch / numpy types.
 */
#ifndef VINSAT_B200_H
#define VINSAT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VINSAT_ABI_VERSION 1

enum {
  VINSAT_OK = 0,
  VINSAT_EINVAL = -1,  /* bad argument */
  VINSAT_ECUDA = -2,   /* CUDA runtime error (message has the CUDA error string) */
  VINSAT_ENOMEM = -3,
  VINSAT_ENODEV = -4   /* no usable CUDA device */
};

enum { VINSAT_MEM_HOST = 0, VINSAT_MEM_DEVICE = 1 };

/* Orbit propagator used inside the dynamics residual:
 *   STEP1S  = BA_utils.py:73-87  `propagate_orbit_dynamics`       (CPU `predict`, the parity reference)
 *   SKIP100 = BA_utils.py:52-71  `propagate_orbit_dynamics_skip`  (`predict_gpu`) */
enum { VINSAT_MODE_STEP1S = 0, VINSAT_MODE_SKIP100 = 1 };

typedef struct vinsat_ctx vinsat_ctx;
typedef struct vinsat_batch vinsat_batch;

/* ---- context ---------------------------------------------------------------------------- */
int vinsat_abi_version(void);
int vinsat_device_count(void);
int vinsat_ctx_create(int device, vinsat_ctx** out);
int vinsat_ctx_destroy(vinsat_ctx* ctx);
/* Run on a caller-owned stream (e.g. torch.cuda.current_stream().cuda_stream).  NULL is the CUDA legacy
 * default stream, exactly as for a kernel launch.  vinsat_ctx_reset_stream returns to the library's own
 * (non-blocking) stream. */
int vinsat_ctx_set_stream(vinsat_ctx* ctx, void* cuda_stream);
int vinsat_ctx_reset_stream(vinsat_ctx* ctx);
int vinsat_ctx_synchronize(vinsat_ctx* ctx);
const char* vinsat_last_error(const vinsat_ctx* ctx); /* ctx may be NULL: last global error */

/* ---- a1: landmark_project  (BA/BA_utils.py:30-50, proj :7-17, apply_inverse_pose_transformation
 *          :1052-1069, attitude_jacobian :19-28) -------------------------------------------------
 * states [T,10], intrinsics [T,4] (fx fy cx cy), landmarks_xyz [M,3], ii [M] (frame of each obs,
 * any order).  uv_out [M,2].  Jg_out [M,2,9] = [-Pi R^T | 2 Pi hat(p_c) | 0] or NULL (jacobian=False).
 * uv_out is bit-identical to the reference (no FMA contraction in the forward projection). */
int vinsat_landmark_project(vinsat_ctx* ctx, int mem, int64_t n_frames, int64_t n_obs,
                            const double* states, const double* intrinsics, const double* landmarks_xyz,
                            const int64_t* ii, double* uv_out, double* Jg_out);

/* ---- a2-a4: predict / predict_gpu  (BA/BA_utils.py:457-527 / :529-602) -------------------------
 * states [T,10]; cum_rot [T,4] = imu_meas[0,:,-1,6:10] (the only slice `predict` consumes, :295);
 * time_idx [T].  r_pred_out [T-1,7].  Block outputs (all nullable together via Phi_out==NULL):
 *   Phi_out [T-1,6,6]   state-transition matrix of pair i in [p,v] order; the reference's dense Jf has
 *                       pair block [D Phi_i | -D], D = diag(1,1,1,vel_coeff x3), rot columns zero
 *   qgrad_out [T,3], Hq_diag_out [T,3,3], Hq_off_out [T-1,3,3] (= Hq[i,i+1]; Hq[i+1,i] is its transpose)
 * x_pred_out [T,6] nullable: propagated [p,v] of every frame (pose_pred / vel_pred of the reference). */
int vinsat_predict(vinsat_ctx* ctx, int mem, int64_t n_frames, const double* states, const double* cum_rot,
                   const int64_t* time_idx, double quat_coeff, double vel_coeff, int mode,
                   double* r_pred_out, double* x_pred_out, double* Phi_out, double* qgrad_out,
                   double* Hq_diag_out, double* Hq_off_out);

/* ---- a8: propagate_dynamics_init (BA/BA_utils.py:89-129) --------------------------------------
 * One state chained for n_steps RK4 steps of dt (orbit) and n_steps quaternion increments exp(dt*omega_k).
 * state0 [10] (velocity taken from vel0 [3], as the reference does), omega [n_steps,3].
 * states_out [(n_steps+1),10] including the start row. */
int vinsat_propagate_chain(vinsat_ctx* ctx, int mem, int64_t n_steps, double dt, const double* state0,
                           const double* vel0, const double* omega, double* states_out);

/* ---- a11: batched orbit simulation (trajgen_pipe.py:145-152 / sim/orbit_gen.py:145-152) ----------
 * n_traj initial [p,v] states propagated n_steps RK4 steps of h; every `stride`-th state is stored:
 * out [n_traj, n_steps/stride + 1, 6]. */
int vinsat_orbit_propagate(vinsat_ctx* ctx, int mem, int64_t n_traj, int64_t n_steps, int64_t stride, double h,
                           const double* x0, double* out);

/* ---- a4 / (f)3: input preparation (BA/BA_utils.py:278-288, 1361-1367; od_pipe.py:944-953) -------
 * vinsat_cum_rotations: compute_omega_from_quat on the full-rate quaternions quat_full [n_full,4] (xyzw), then
 * for every frame the ordered product of exp(dt*omega_s) over its gap s = time_idx[i] .. time_idx[i+1]-1 --
 * exactly precompute_cum_rotations(omegas, dt)[0, :, -1], the only slice `predict` reads (BA_utils.py:295); the
 * last frame gets the identity.  omega_out [n_full,3] nullable, cum_rot_out [n_frames,4].
 * vinsat_precompute_cum_rotations: the general form, omegas [n_frames, n_slots, 3] -> cum_out [n_frames, n_slots, 4]. */
int vinsat_cum_rotations(vinsat_ctx* ctx, int mem, int64_t n_full, const double* quat_full, double dt,
                         int64_t n_frames, const int64_t* time_idx, double* omega_out, double* cum_rot_out);
int vinsat_precompute_cum_rotations(vinsat_ctx* ctx, int mem, int64_t n_frames, int64_t n_slots, const double* omegas,
                                    double dt, double* cum_out);

/* ---- a11 / (f)3: batched rigid-body attitude simulation (trajgen_pipe.py:155-207, attitude_step) ----
 * State [q (scalar FIRST, as the reference's L(q)), omega] = 7 doubles; inertia_diag [3] = diag(J) (HOST pointer).
 * out [n_traj, n_steps/stride + 1, 7]. */
int vinsat_attitude_propagate(vinsat_ctx* ctx, int mem, int64_t n_traj, int64_t n_steps, int64_t stride, double h,
                              const double* inertia_diag, const double* x0, double* out);

/* ---- batched BA / OD (BA/BA_filtering.py:4-98 driven by od_pipe.py:1036-1040) -----------------------
 * A batch holds P independent problems, concatenated; problem p owns frames
 * [frame_off[p], frame_off[p+1]) and observations [obs_off[p], obs_off[p+1]).  `ii` holds frame indices
 * LOCAL to the problem and must be non-decreasing inside each problem (read_detections builds it that
 * way, od_pipe.py:214-228); the Python mirror sorts otherwise. */
typedef struct {
  int64_t n_problems;
  const int64_t* frame_off;     /* [P+1] */
  const int64_t* obs_off;       /* [P+1] */
  const double* states;         /* [Ttot,10] initial guess */
  const double* intrinsics;     /* [Ttot,4]  */
  const double* cum_rot;        /* [Ttot,4]  */
  const int64_t* time_idx;      /* [Ttot]    */
  const double* landmarks_xyz;  /* [Mtot,3]  */
  const double* landmarks_uv;   /* [Mtot,2]  */
  const double* confidences;    /* [Mtot]    */
  const int64_t* ii;            /* [Mtot]    */
} vinsat_problem_desc;

/* Allocates device storage for the sizes in `desc` and uploads it (host pointers). */
int vinsat_batch_create(vinsat_ctx* ctx, const vinsat_problem_desc* desc, vinsat_batch** out);
/* Re-upload new data of the SAME sizes (async on the context stream; host buffers should be pinned). */
int vinsat_batch_upload(vinsat_batch* b, const vinsat_problem_desc* desc);
int vinsat_batch_set_states(vinsat_batch* b, int mem, const double* states /* [Ttot,10] */);
int vinsat_batch_get_states(vinsat_batch* b, int mem, double* states_out /* [Ttot,10] */);
int vinsat_batch_destroy(vinsat_batch* b);

/* One BA() call per problem (BA_filtering.py:4-98): residuals + Jacobians, robust weights, block
 * tridiagonal normal equations, LM trials until accepted or lamda > 1e4, retraction.
 * lamda_io [P] host: in = lamda_init, out = lamda_init for the next call (:79).  ntrials_out [P] nullable. */
int vinsat_batch_ba_iterate(vinsat_batch* b, int iter, int initialize, int mode, double* lamda_io,
                            int32_t* ntrials_out);
/* streaming_version's schedule on one window (od_pipe.py:918,1036-1040): num_iters BA iterations, the first
 * n_init with initialize=1.  Equivalent to calling vinsat_batch_ba_iterate in a loop. */
int vinsat_batch_od_solve(vinsat_batch* b, int num_iters, int n_init, double lamda_init, int mode);
/* JTwJ[:, -9:, -9:] of the last trial of the last iteration (BA_filtering.py:97).  out [P,9,9]. */
int vinsat_batch_last_hessian(vinsat_batch* b, double* out);

/* ---- (f)1: streaming_version's outer loop (od_pipe.py:987-1060) in one call ------------------------------------
 * The whole sequence after read_detections / remove_elems (T_all frames incl. knots, M_all observations, `ii` GLOBAL
 * frame indices, non-decreasing) plus the window schedule: window w covers frames [0, t_final[w]) and observations
 * [0, i_final[w]) -- the (t_final, i_final) pairs identify_next_batch_new returns (:898-905, integers, computed by
 * the caller).  Window 0 starts from `states`; the frames a later window adds are seeded on the device by
 * propagate_dynamics_init (:1011, BA_utils.py:114-129) from the last solved state and the carried `velocities` row
 * (omega = full-rate compute_omega_from_quat output, [n_omega,3]); every window then runs num_iters BA iterations
 * (the first n_init_first of window 0 with initialize=1, :1036-1040), starting from lamda_init (:1032).  If the last
 * window ends before T_all the remaining frames are propagated only (:1046-1060).
 * Outputs (host): states_out [t_final[W-1],10] solved states of the last window; seed_states_out [T_all,10] nullable
 * = the state every frame entered its window with (states for window 0, propagated seeds afterwards, tail included);
 * window_last_state_out [W,10] nullable = last frame's state after each window's solve (:1042); last_hessian_out
 * [9,9] nullable (BA_filtering.py:97 of the last call). */
typedef struct {
  int64_t n_frames, n_obs;
  const double* states;        /* [T_all,10] */
  const double* velocities;    /* [T_all,3]  the `velocities` array carried beside the states (od_pipe.py:942) */
  const double* intrinsics;    /* [T_all,4]  */
  const double* cum_rot;       /* [T_all,4]  */
  const int64_t* time_idx;     /* [T_all]    */
  const double* landmarks_xyz; /* [M_all,3]  */
  const double* landmarks_uv;  /* [M_all,2]  */
  const double* confidences;   /* [M_all]    */
  const int64_t* ii;           /* [M_all]    */
  int64_t n_omega;
  const double* omega;         /* [n_omega,3] */
  int64_t n_windows;
  const int64_t* t_final;      /* [W] */
  const int64_t* i_final;      /* [W] */
} vinsat_stream_desc;
int vinsat_stream_solve(vinsat_ctx* ctx, const vinsat_stream_desc* desc, int num_iters, int n_init_first,
                        double lamda_init, int mode, double* states_out, double* seed_states_out,
                        double* window_last_state_out, double* last_hessian_out);

/* ---- (f)4: the prior-regularised variant (BA_reg, BA/BA_filtering.py:100-210) ------------------------------------------
 * vinsat_prior = prior_gpu (BA/BA_utils.py:604-676): states / prop_states [N,10], hessian_state [N,6,6] (position /
 * velocity information, row-major), hessian_rot [N,3,3].  r_out [N,7] = [H e ; quat_coeff (1 - |.|)].  Block outputs
 * (all three or none): Jp_out [N,6,9] = the diagonal blocks of the reference's dense Jacobian (everything off the block
 * diagonal is exactly zero there), Hqp_out [N,9,9], qgrad_out [N,9].  NOTE the parameter order (vel_coeff before
 * quat_coeff) is prior_gpu's own; BA_reg passes its quat / vel coefficients the other way round (:122) -- the mirror
 * keeps that call as it is.
 * vinsat_propagate_chain_cov = propagate_dynamics_cov_init (BA/BA_utils.py:222-248): state0 [10], vel0 [3], hessian
 * [9,9] (last_hessian of the previous window), omega [(tdiff+duration),3] -> states_out [(duration+1),10],
 * hessian_state_out [(duration+1),6,6], hessian_rot_out [(duration+1),3,3].
 * vinsat_batch_set_prior + vinsat_batch_ba_reg_iterate: one BA_reg() call per problem of the batch (initialize=False, the
 * only way the reference calls it, od_pipe.py:893), same lamda_io / ntrials_out contract as vinsat_batch_ba_iterate. */
int vinsat_prior(vinsat_ctx* ctx, int mem, int64_t n_frames, const double* states, const double* prop_states,
                 double vel_coeff, double quat_coeff, const double* hessian_state, const double* hessian_rot,
                 double* r_out, double* Jp_out, double* Hqp_out, double* qgrad_out);
int vinsat_propagate_chain_cov(vinsat_ctx* ctx, int mem, int64_t tdiff, int64_t duration, double dt, const double* state0,
                               const double* vel0, const double* hessian, const double* omega, double* states_out,
                               double* hessian_state_out, double* hessian_rot_out);
int vinsat_batch_set_prior(vinsat_batch* b, int mem, const double* states_prior, const double* hessian_state,
                           const double* hessian_rot);
int vinsat_batch_ba_reg_iterate(vinsat_batch* b, int iter, int mode, double* lamda_io, int32_t* ntrials_out);

/* ---- Monte-Carlo noise sweeps on a resident batch (configs[3]; the reference's Monte Carlo is the sequential loop of
 *      od_pipe.py:1063-1086) ------------------------------------------------------------------------------------------
 * set_truth: the true states [Ttot,10] and noise-free pixels [Mtot,2] the draws are centred on (+ optional true
 * velocities [Ttot,3] for the velocity error).  perturb: a fresh initial guess as od_pipe.py:962-969 (position +
 * N(0,pos_sigma), rotation exp(log q + N(0,rot_sigma)), velocity + N(0,vel_sigma)) and fresh pixel noise N(0,sigma_px),
 * drawn ON THE DEVICE from a counter-based generator keyed by `seed` (same numbers on any GPU).  errors: per problem
 * max |p - p_true| and max |v - v_true| over its frames, host outputs [P]. */
int vinsat_batch_mc_set_truth(vinsat_batch* b, const double* states_true, const double* uv_true, const double* vel_true);
int vinsat_batch_mc_perturb(vinsat_batch* b, uint64_t seed, double sigma_px, double pos_sigma, double rot_sigma,
                            double vel_sigma);
int vinsat_batch_mc_errors(vinsat_batch* b, double* pos_err_out, double* vel_err_out);

/* Per-iteration diagnostics of the LAST vinsat_batch_ba_iterate call, for parity tests (host outputs,
 * any may be NULL): r_obs [Mtot,2], weights [Mtot] (after /max and *conf), c_obs [P], D [Ttot,9,9] (without
 * damping), U [Ttot,9,9] (block (i,i+1); last of each problem unused), rhs [Ttot,9], dpose [Ttot,9]. */
int vinsat_batch_debug_fetch(vinsat_batch* b, double* r_obs, double* weights, double* c_obs, double* D,
                             double* U, double* rhs, double* dpose);

/* Headline kernel, timed alone: residual + Jacobian for every resident observation (a1, unfused).
 * Outputs stay on the device in SoA form (r[2][M], J[12][M]); `fetch` copies them out as
 * r_out [Mtot,2], J_out [Mtot,2,6] (nonzero columns only), both nullable. */
int vinsat_batch_eval_resjac(vinsat_batch* b);
int vinsat_batch_fetch_resjac(vinsat_batch* b, double* r_out, double* J_out);

/* Device-time of the library's kernels since the last reset, by kernel family, measured with CUDA
 * events on the context stream (only when profiling is enabled; adds a sync per launch).  */
int vinsat_ctx_enable_timing(vinsat_ctx* ctx, int on);
int vinsat_ctx_reset_timing(vinsat_ctx* ctx);
/* names_out: up to `cap` pointers to static strings; ms_out/launches_out parallel arrays. Returns count. */
int vinsat_ctx_get_timing(vinsat_ctx* ctx, int cap, const char** names_out, double* ms_out, int64_t* launches_out);
/* Total kernel launches issued by this context since creation (bench.py's `gpu_launches`). */
int64_t vinsat_ctx_launch_count(const vinsat_ctx* ctx);

/* ---- frame-window sharded long arc (config 3) ----------------------------------------------
 * One problem split by contiguous frame windows across ranks (one process per GPU).  The rank's batch holds
 * its owned frames [own_lo, own_hi) plus at most one GHOST frame per side (the neighbour's edge frame: state,
 * intrinsics, cum_rot, time) and the observations of the owned frames only.  The host drives the iteration
 * stage by stage and exchanges the exposed device buffers between stages with NCCL (vinsat_b200/longarc.py
 * documents the choreography).  Math and results are those of vinsat_batch_ba_iterate on the whole arc. */
int vinsat_batch_create_window(vinsat_ctx* ctx, const vinsat_problem_desc* desc, int64_t own_lo, int64_t own_hi,
                               int64_t n_segments, vinsat_batch** out);
int vinsat_la_num_segments(const vinsat_batch* b);
int vinsat_la_alloc_reduced(vinsat_batch* b, int64_t S_total, int64_t n_ranks);
enum {
  VINSAT_LA_RESID = 0, VINSAT_LA_SELECT_BEGIN, VINSAT_LA_SELECT_HIST, VINSAT_LA_SELECT_PICK, VINSAT_LA_ASSEMBLE,
  VINSAT_LA_DYNAMICS, VINSAT_LA_SYSTEM, VINSAT_LA_SUMS_INIT, VINSAT_LA_SET_LAM, VINSAT_LA_SOLVE_INIT,
  VINSAT_LA_FORWARD, VINSAT_LA_REDUCED, VINSAT_LA_BACKSUB, VINSAT_LA_RETRACT, VINSAT_LA_PACK_EDGES,
  VINSAT_LA_APPLY_GHOSTS, VINSAT_LA_TRIAL, VINSAT_LA_SUMS_TRIAL, VINSAT_LA_COMMIT,
  /* device-side LM bookkeeping (lets one iteration be captured as one CUDA graph, kernels + NCCL): */
  VINSAT_LA_BEGIN_ITER, VINSAT_LA_ACCEPT, VINSAT_LA_COMMIT_COPY
};
enum {
  VINSAT_LA_BUF_HIST = 0,   /* uint32[2048]   all-reduce SUM between SELECT_HIST and SELECT_PICK */
  VINSAT_LA_BUF_WMAX,       /* int64[1]       all-reduce MAX after ASSEMBLE (bit pattern of a double >= 0) */
  VINSAT_LA_BUF_SUMS,       /* float64[4]     all-reduce SUM after SUMS_INIT / SUMS_TRIAL */
  VINSAT_LA_BUF_PACK,       /* float64[n_seg*514]  all-gather into GATHER after FORWARD */
  VINSAT_LA_BUF_GATHER,     /* float64[S_total*514] */
  VINSAT_LA_BUF_EDGE,       /* float64[20]    all-gather into EDGES_ALL after PACK_EDGES */
  VINSAT_LA_BUF_EDGES_ALL,  /* float64[n_ranks*20] */
  VINSAT_LA_BUF_FLAGS,      /* int32[4]: [0] = 1 while the LM loop of the iteration wants another trial */
  VINSAT_LA_BUF_LAM_NEXT,   /* float64[1]: lamda_init of the next BA call */
  VINSAT_LA_BUF_NTRIALS     /* int32[1] */
};
int vinsat_la_stage(vinsat_batch* b, int stage, int64_t i0, int64_t i1, double d0);
int vinsat_la_ptr(vinsat_batch* b, int which, void** ptr, int64_t* count);

/* ---- a10: SatCam batched projection / visibility (sim/SatCam.py:87-92,125-154,175-262) -------------
 * poses [P,12] = [ECEF pos (m), dir, up, right] (SatCam.py:23-28; note up is negated, :52,81);
 * landmarks_ecef [L,3] (m).  uv_out [P,L,2] nullable; inframe_out [P,L] uint8 nullable
 * (0<=u<w_px, 0<=v<h_px, in front of the camera); count_out [P] int32 nullable (# in frame). */
int vinsat_satcam_project(vinsat_ctx* ctx, int mem, int64_t n_poses, int64_t n_landmarks, const double* poses,
                          const double* landmarks_ecef, double hfov_deg, int32_t w_px, int32_t h_px,
                          double* uv_out, uint8_t* inframe_out, int32_t* count_out);
/* Footprint corners by ray casting to the WGS84 ellipsoid (SatCam.py:94-147): corners_out [P,4,3] ECEF m
 * (tl,tr,br,bl), hit_out [P,4] uint8. */
int vinsat_satcam_corners(vinsat_ctx* ctx, int mem, int64_t n_poses, const double* poses, double hfov_deg,
                          int32_t w_px, int32_t h_px, double* corners_out, uint8_t* hit_out);
/* Same, plus the unit corner rays of get_corner_vectors (SatCam.py:94-104): vec_out [P,4,3] nullable. */
int vinsat_satcam_corner_rays(vinsat_ctx* ctx, int mem, int64_t n_poses, const double* poses, double hfov_deg,
                              int32_t w_px, int32_t h_px, double* corners_out, uint8_t* hit_out, double* vec_out);
/* world_to_pixel_mat (SatCam.py:87-92) for every pose: C_out [P,3,4]. */
int vinsat_satcam_cam_matrix(vinsat_ctx* ctx, int mem, int64_t n_poses, const double* poses, double hfov_deg,
                             int32_t w_px, int32_t h_px, double* C_out);

/* ---- a10: the reference's visibility predicate, batched (sim/SatCam.py:175-262 driven per pose by
 *          sim/nadir_sim.py:175,198) ------------------------------------------------------------------------
 * Landmark table, resident on the device: region r (code = zone*32 + letter, 'A'=1; '10S' = 10*32+19) owns rows
 * region_off[r] .. region_off[r+1] of centroid_lonlat [n,2] (the 'Centroid Longitude', 'Centroid Latitude' columns of
 * sim/landmark_csvs/<region>_top_salient.csv in row order; host pointers).  active_codes = `self.regions`
 * (SatCam.py:63-67): only those regions are ever tested (:258). */
typedef struct vinsat_satcam_table vinsat_satcam_table;
int vinsat_satcam_table_create(vinsat_ctx* ctx, int32_t n_regions, const int32_t* region_codes,
                               const int64_t* region_off, const double* centroid_lonlat, int32_t n_active,
                               const int32_t* active_codes, vinsat_satcam_table** out);
int vinsat_satcam_table_destroy(vinsat_satcam_table* table);
/* check_for_all_landmarks (SatCam.py:254-262) for every pose: corner rays -> WGS84 ellipsoid (cast_ray_to_earth,
 * :125-147) -> geodetic lon/lat (get_corner_lonlats, :175-185; astropy's conversion is replaced by the closed form for
 * a point on the ellipsoid) -> get_region (:187-191) -> find_current_regions (:203-230) ->
 * check_for_landmarks_in_region (:232-251; the best_classes blob is not shipped: all classes).
 * visible_out [P] uint8 = what the reference returns.  Nullable extras: count_out [P] = landmarks counted without
 * the early exits; corner_lonlat_out [P,4,2] (NaN where the ray misses); corner_region_out [P,4] region codes (-1 = None).
 * With VINSAT_MEM_DEVICE the call is asynchronous on the context stream. */
int vinsat_satcam_visibility(vinsat_ctx* ctx, const vinsat_satcam_table* table, int mem, int64_t n_poses,
                             const double* poses, double hfov_deg, int32_t w_px, int32_t h_px, uint8_t* visible_out,
                             int32_t* count_out, double* corner_lonlat_out, int32_t* corner_region_out);

/* ---- (f)2: ingest indexing as integer kernels (od_pipe.py:214-247, 253-288) ------------------------------------------
 * vinsat_index_detections: `frames` [n_det] = detections[:, 0] (float64 frame ids, sorted non-decreasing as the
 * reference assumes).  time_idx_out (capacity cap_frames >= #unique frames + n_orbit/1000 + 2) = unique frames with a
 * knot frame at every multiple of 1000 s between / after the detections (:216-226,242-245), ii_out [n_det] = frame slot
 * of every detection (:227-228), *n_frames_out (host) = length of time_idx.
 * vinsat_remove_elems_index: mask [n_det] (1 = keep, the visibility mask of :930), ii / time_idx as above.
 * frame_keep_out [n_frames] = frames with a surviving observation or knots (:259-263); ii_out = re-indexed surviving
 * observations, ii - #dropped frames below (:272-282); time_idx_out = time_idx[keep]; counts_out (host) = {#obs kept,
 * #frames kept}.  Bit-exact against the reference (integers only). */
int vinsat_index_detections(vinsat_ctx* ctx, int mem, int64_t n_det, const double* frames, int64_t n_orbit,
                            int64_t cap_frames, int64_t* time_idx_out, int64_t* ii_out, int64_t* n_frames_out);
int vinsat_remove_elems_index(vinsat_ctx* ctx, int mem, int64_t n_det, int64_t n_frames, const uint8_t* mask,
                              const int64_t* ii, const int64_t* time_idx, int64_t* ii_out, int64_t* time_idx_out,
                              uint8_t* frame_keep_out, int64_t* counts_out);

/* ---- measurement helpers ------------------------------------------------------------------ */
/* FP64 FMA peak of the device (DFMA microbenchmark), TFLOP/s; and a device copy bandwidth, GB/s. */
int vinsat_measure_fp64_peak(vinsat_ctx* ctx, double* tflops_out);
int vinsat_measure_copy_bw(vinsat_ctx* ctx, int64_t bytes, double* gbs_out);

#ifdef __cplusplus
}
#endif
#endif /* VINSAT_B200_H */
