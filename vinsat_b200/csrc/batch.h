// Device-resident batch of OD problems (internal).
#pragma once
#include <cmath>
#include <cstdlib>
#include <map>
#include <vector>
#include <set>
#include "internal.h"

// Per-frame record sizes (doubles).
#define VS_GREC 28    // obs normal block: 21 sym (6x6 upper) + 6 rhs + 1 sum|r|
#define VS_DREC 64    // dynamics: Phi 36 | r6 6 | rho 1 | qgrad 3 | Hq_diag 9 | Hq_off 9
#define VS_MREC 42    // Phi^T D^2 Phi (6x6, both triangles, [a*6+b]) | Phi^T D r (6)  (written by k_dynamics_stm)
#define VS_SREC 172   // system: D 81 | U 81 | b 9 | pad 1
#define VS_WREC 172   // solver: W = S^-1 U (col-major 81) | y 9 | Z spike (col-major 81) | pad
constexpr int kSumChunk = 2048;     // frames per CTA of the two-stage per-problem sums (long arcs, kernels_solve.cu)
constexpr int kSumThreads = 256;
#define VS_RREC 342   // reduced-system contributions of a segment: left {Dl 81, Ll 81, bl 9} | right {Dr 81, Ur 81, br 9}

// Second partition level of the block-tridiagonal solve (kernels_chain.cu): the REDUCED chain of a long problem (one element
// per level-1 separator) is itself cut into segments, so that the sequential depth of a solve is
// len(level-1 segment) + len(level-2 segment) + #level-2 segments instead of len + #level-1 segments.
// Indices a / b / left are level-1 separator numbers.
struct Level2 {
  int64_t n = 0;            // level-2 segments (0 = off)
  int64_t n_chains = 0;     // reduced-2 chains (one per problem)
  int32_t *a = nullptr, *b = nullptr, *left = nullptr, *prob = nullptr, *has_next = nullptr;   // [n]
  int32_t *red_a = nullptr, *red_b = nullptr, *red_prob = nullptr;                            // [n_chains]
  double *redrec = nullptr, *rsys = nullptr, *rlow = nullptr, *rwrec = nullptr;               // [n] records
};

struct Level2Host {
  std::vector<int32_t> a, b, left, prob, has_next, red_a, red_b, red_prob;
};

// Plan: problems whose reduced chain has >= min_sep elements get ~sqrt(S) level-2 segments, the others one.
// Empty (no level 2) when no problem is that long.
inline Level2Host plan_level2(const std::vector<int32_t>& red_a, const std::vector<int32_t>& red_b,
                              const std::vector<int32_t>& prob, int64_t min_sep) {
  Level2Host h;
  bool any = false;
  for (size_t p = 0; p < red_a.size(); p++) any = any || (red_b[p] - red_a[p] >= min_sep);
  if (!any) return h;
  for (size_t p = 0; p < red_a.size(); p++) {
    const int64_t S = red_b[p] - red_a[p];
    const int64_t S2 = S >= min_sep ? std::min<int64_t>(S, std::max<int64_t>(2, llround(sqrt((double)S)))) : (S > 0 ? 1 : 0);
    h.red_a.push_back((int32_t)h.a.size());
    for (int64_t k = 0; k < S2; k++) {
      const int64_t lo = red_a[p] + (S * k) / S2, hi = red_a[p] + (S * (k + 1)) / S2;   // elements [lo, hi), separator hi-1
      h.a.push_back((int32_t)lo);
      h.b.push_back((int32_t)(hi - 1));
      h.left.push_back(k > 0 ? (int32_t)(lo - 1) : -1);
      h.prob.push_back(prob[p]);
      h.has_next.push_back(k + 1 < S2 ? 1 : 0);
    }
    h.red_b.push_back((int32_t)h.a.size());
    h.red_prob.push_back(prob[p]);
  }
  return h;
}

inline int64_t level2_min_separators() {
  const int64_t v = getenv("VINSAT_L2_MIN") ? atoll(getenv("VINSAT_L2_MIN")) : 256;    // <= 0: level 2 off (read at batch creation)
  return v > 0 ? v : (int64_t)1 << 62;
}

struct vinsat_batch {
  vinsat_ctx* ctx = nullptr;
  int64_t P = 0, T = 0, M = 0;      // problems, total frames, total observations
  int64_t n_pairs = 0;
  int64_t max_obs_per_problem = 0;
  // host copies of the offsets
  std::vector<int64_t> frame_off, obs_off;
  // ---- device, per frame ----
  double* st = nullptr;        // [T][10] current states (AoS rows)
  double* st_new = nullptr;    // [T][10] trial states
  double* intr = nullptr;      // [T][4]
  double* crot = nullptr;      // [T][4]
  int32_t* gap = nullptr;      // [T] seconds to the next frame of the same problem, 0 = no pair
  int32_t* fprob = nullptr;    // [T] problem of the frame
  int32_t* dyn_order = nullptr;  // [n_pairs] frame index of each pair, longest gap first
  int32_t* obs_start = nullptr;  // [T+1] CSR of observations by frame
  double* grec = nullptr;      // [T][VS_GREC]
  double* drec = nullptr;      // [T][VS_DREC]
  double* mrec = nullptr;      // [T][VS_MREC]
  double* srec = nullptr;      // [T][VS_SREC]
  double* wrec = nullptr;      // [T][VS_WREC]
  double* delta = nullptr;     // [T][9]
  double* zeros = nullptr;     // [64] zero page read by lanes of the fused sweep that have no term of a kind
  // ---- partitioned block-tridiagonal solve (kernels_chain.cu) ----
  bool partitioned = false;
  int64_t n_seg = 0;
  int32_t *seg_a = nullptr, *seg_b = nullptr, *seg_left = nullptr, *seg_prob = nullptr, *seg_has_next = nullptr;  // [n_seg]
  int32_t *pl_a = nullptr, *pl_b = nullptr, *pl_prob = nullptr;   // [P] whole-problem chains
  int32_t *red_a = nullptr, *red_b = nullptr;                     // [P] reduced chains (segment index ranges)
  // two-sided sweep: chains 2p (top) / 2p+1 (bottom): first element, end sentinel (= middle), direction, problem, middle
  int32_t *bb_a = nullptr, *bb_e = nullptr, *bb_dir = nullptr, *bb_prob = nullptr, *bb_mid = nullptr;   // [2P]
  double* midrec = nullptr;                                       // [2P][96]
  double* redrec = nullptr;    // [n_seg][VS_RREC]
  double* rsys = nullptr;      // [n_seg][VS_SREC] reduced system rows
  double* rlow = nullptr;      // [n_seg][81] explicit lower blocks of the reduced system
  double* rwrec = nullptr;     // [n_seg][VS_WREC]
  Level2 l2;                   // second partition level over the reduced chains (long problems)
  double* xsep = nullptr;      // [n_seg][9] level-1 separator solutions (level-2 path)
  double* e_obs = nullptr;     // [T] trial partial sums (obs part)
  double* e_dyn = nullptr;     // [T] trial partial sums (dynamics part)
  // ---- device, per observation (SoA) ----
  double* X = nullptr;         // [3][M]
  double* uv = nullptr;        // [2][M]
  double* conf = nullptr;      // [M]
  int32_t* oframe = nullptr;   // [M] global frame of the observation
  double* r = nullptr;         // [2][M] residuals
  double* r_next = nullptr;    // [2][M] residuals at the last trial's states = next iteration's residuals
  bool r_valid = false;
  // CUDA graphs of one BA iteration up to and including its first LM trial, keyed by everything the captured
  // launches depend on (iteration index, phase, propagator, which of the ping-pong buffers is current)
  struct IterGraph { cudaGraphExec_t exec = nullptr; int64_t n_launches = 0; };
  std::map<uint64_t, IterGraph> graphs;
  std::set<uint64_t> graph_bad;       // keys whose capture / instantiation failed: plain launches from then on
  std::set<uint64_t> graph_warm;      // keys that ran eagerly once (function attributes set, scratch grown)
  const double* st_base = nullptr;    // identity of the ping-pong buffers at creation
  const double* r_base = nullptr;
  double* wu = nullptr;        // [M] conf * w_raw (un-normalised robust weight)
  double* J = nullptr;         // [12][M] headline kernel output (allocated on first use)
  // ---- device, per problem ----
  int64_t* d_frame_off = nullptr;  // [P+1]
  int64_t* d_obs_off = nullptr;    // [P+1]
  double* c_obs = nullptr;     // [P] robust scale (lower median of |r|)
  unsigned long long* wmax = nullptr;   // [P] bits of max w_raw
  double* lam = nullptr;       // [P] current damping
  double* lam_next = nullptr;  // [P] lamda_init for the next BA call
  double* lam32_last = nullptr;  // [P] damping actually added in the last trial
  double* init_res = nullptr;  // [P]
  int32_t* active = nullptr;   // [P]
  int32_t* ntrials = nullptr;  // [P]
  unsigned long long* sel_prefix = nullptr;  // [P]
  unsigned long long* sel_rank = nullptr;    // [P]
  unsigned int* sel_hist = nullptr;          // [P][2048]
  int32_t* flags = nullptr;    // [4]: 0 = n_active, 1 = index error
  int32_t* h_flags = nullptr;  // pinned mirror ([0]: trial counter; [2], [3]: ring for the speculative pipeline)
  int32_t* gate = nullptr;     // device word: != 0 while an LM loop behind speculatively launched work is unfinished
  const int32_t* gate_arg = nullptr;   // what the launches of the moment pass to their kernels (null = ungated)
  // ---- frame-window sharded long arc (longarc.cu): this batch holds ONE problem = owned frames + ghosts ----
  bool window = false;
  int64_t own_lo = 0, own_hi = 0;     // owned local frames [own_lo, own_hi); ghosts (if any) at own_lo-1 / own_hi
  int64_t forced_segments = 0;
  int64_t S_total = 0, n_ranks = 0;
  double* la_pack = nullptr;          // [n_seg][VS_RREC + VS_SREC]
  double* la_gath = nullptr;          // [S_total][VS_RREC + VS_SREC]
  double* la_rsys = nullptr;          // [S_total][VS_SREC]
  double* la_rlow = nullptr;          // [S_total][81]
  double* la_rwrec = nullptr;         // [S_total][VS_WREC]
  double* la_xsep = nullptr;          // [S_total][9]
  double* la_sums = nullptr;          // [4]
  // two-stage per-problem sums of long arcs (k_sum_partials)
  double* sum_part = nullptr;         // [sum_chunks_cap][3]
  int32_t* sum_chunk_off = nullptr;   // [P+1] first chunk of every problem
  int64_t sum_chunks = 0, sum_chunks_cap = 0;
  double* la_edge = nullptr;          // [2][10] new states of the first / last owned frame
  double* la_edges_all = nullptr;     // [n_ranks][2][10]
  int32_t* la_chain = nullptr;        // {0, S_total, 0}
  Level2 la_l2;                       // second partition level over the gathered reduced chain
  // ---- Monte-Carlo noise sweeps (mc.cu): true states / pixels the perturbations are drawn around, error scratch ----
  double *mc_st_true = nullptr, *mc_uv_true = nullptr, *mc_vel_true = nullptr, *mc_err = nullptr;
  // ---- BA_reg (prior.cu): prior states / information matrices and the per-frame sums of |r_prior| ----
  double *pr_st = nullptr, *pr_Hs = nullptr, *pr_Hr = nullptr, *e_pr_init = nullptr, *e_pr = nullptr;
  bool reg_iter = false;       // the BA() call in flight is a BA_reg() call (prior terms on)
  bool in_arena = false;       // device buffers come from ctx->arena (never cudaFree'd individually)
  bool have_iter = false;
  bool srec_valid = false;
  double last_sigma = 0.0;
  double cur_sigma = 0.0;      // Sigma / velocity weight of the iteration in flight (fused forward sweep)
  double cur_vc = 100.0;
  bool fused_system = false;   // the sweep builds its columns from grec/drec/mrec (no srec round trip)
  int last_initialize = 0;
};
