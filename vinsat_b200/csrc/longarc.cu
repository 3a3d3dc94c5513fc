// Frame-window sharded long-arc BA (config 3 of BASELINE.json).
//
// One problem is split by contiguous frame windows across ranks.  A rank's batch holds its owned frames plus one
// ghost frame per side (state, gap, cum_rot of the neighbour's edge frame), so every per-frame / per-pair kernel
// runs unchanged on local data.  What needs the other ranks is exchanged by the HOST between the stages below
// (NCCL over NVLink through torch.distributed, vinsat_b200/longarc.py), always as tiny device buffers:
//   - robust scale: the 2048-bin digit histogram of each radix-select pass is all-reduced (exact global median),
//     max weight all-reduced (max);
//   - normal equations: every rank eliminates the interiors of its segments, the per-segment records of the
//     REDUCED block-tridiagonal system are all-gathered, every rank solves the small reduced chain redundantly;
//   - accept test: two partial sums all-reduced; ghost states re-exchanged after every retraction.
#include "common.cuh"
#include "launch.h"

using namespace vs;

namespace {

#define VS_TRY(expr)            \
  do {                          \
    int _rc = (expr);           \
    if (_rc != VINSAT_OK) return _rc; \
  } while (0)

constexpr int kPack = VS_RREC + VS_SREC;

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sums over the OWNED frames: stage 1 = k_sum_partials (kernels_solve.cu: one CTA per chunk of kSumChunk frames -> part[c][3]),
// stage 2 = this kernel, one CTA adding the chunk sums in a fixed order (deterministic).  A single warp walking the window took
// 11 ms per call at 1.2 M frames per rank.
//   which=0: sums[0] = sum grec[f][27] (|r_obs| unweighted), sums[1] = sum_{owned pairs} |r_pred| (7 components)
//   which=1: sums[0] = sum e_obs[f],                         sums[1] = sum_{owned pairs} e_dyn[f]
// The linearisation's sums go to sums[0..1], the trial's to sums[2..3] (the weighted observation sum already divided
// by the GLOBAL largest weight, which the all-reduce(MAX) after ASSEMBLE left in wmax), so that ONE all-reduce of the
// four doubles and ONE host read at the end of a trial carry everything the accept test needs.
__global__ void __launch_bounds__(256) k_la_sums_final(int n_chunks, int which, const double* __restrict__ part,
                                                       const unsigned long long* __restrict__ wmax, double* __restrict__ sums) {
  __shared__ double s_part[2][8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double so = 0.0, sd = 0.0;
  for (int c = tid; c < n_chunks; c += 256) { so += part[(int64_t)c * 3]; sd += part[(int64_t)c * 3 + 1]; }
  so = warp_sum_d(so);
  sd = warp_sum_d(sd);
  if (lane == 0) { s_part[0][warp] = so; s_part[1][warp] = sd; }
  __syncthreads();
  if (tid == 0) {
    so = 0.0; sd = 0.0;
#pragma unroll
    for (int w = 0; w < 8; w++) { so += s_part[0][w]; sd += s_part[1][w]; }
    if (which == 0) { sums[0] = so; sums[1] = sd; sums[2] = 0.0; sums[3] = 0.0; }
    else {
      const unsigned long long wb = wmax[0];
      const double w = __longlong_as_double((long long)wb);
      sums[2] = (wb && w > 0.0) ? so / w : 0.0;
      sums[3] = sd;
    }
  }
}

__global__ void k_la_set_lam(double lam_v, double* __restrict__ lam, int32_t* __restrict__ active) {
  lam[0] = lam_v;
  active[0] = 1;
}

// Device-side LM bookkeeping, so that a whole iteration (kernels + NCCL exchanges) can be ONE CUDA graph: the damping
// never passes through the host.  begin: lam <- lam_next (what the previous BA call handed on, BA_filtering.py:79).
__global__ void k_la_begin_iter(const double* __restrict__ lam_next, double* __restrict__ lam, int32_t* __restrict__ active,
                                int32_t* __restrict__ ntrials, int32_t* __restrict__ flags) {
  lam[0] = lam_next[0];
  active[0] = 1;
  ntrials[0] = 0;
  flags[0] = 1;
}

// accept test on the all-reduced sums (identical on every rank): BA_filtering.py:51,66-79
__global__ void k_la_accept(const double* __restrict__ sums, double n, double sqrt_sigma, double* __restrict__ init_res,
                            double* __restrict__ lam, double* __restrict__ lam_next, int32_t* __restrict__ ntrials,
                            int32_t* __restrict__ flags) {
  if (ntrials[0] == 0) init_res[0] = (sums[0] + sqrt_sigma * sums[1]) / n;
  const double residual = (sums[2] + sqrt_sigma * sums[3]) / n;
  const double l = lam[0] * 10.0;
  lam[0] = l;
  ntrials[0] += 1;
  const bool done = (residual < init_res[0]) || (l > 1e4);
  flags[0] = done ? 0 : 1;
  if (done) lam_next[0] = fmax(fmin(1e-1, l * 0.01), 1e-4);
}

__global__ void __launch_bounds__(256) k_la_pack(int n_seg, const int32_t* __restrict__ seg_b,
                                                 const double* __restrict__ redrec, const double* __restrict__ srec,
                                                 double* __restrict__ pack) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int s = (int)(t / kPack), e = (int)(t % kPack);
  if (s >= n_seg) return;
  pack[t] = e < VS_RREC ? redrec[(int64_t)s * VS_RREC + e] : srec[(int64_t)seg_b[s] * VS_SREC + (e - VS_RREC)];
}

__global__ void __launch_bounds__(128) k_la_scatter(int n_seg, int64_t seg0, int64_t own_lo,
                                                    const int32_t* __restrict__ seg_b,
                                                    const double* __restrict__ xsep, double* __restrict__ delta) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int s = t / 9, r = t % 9;
  if (s < n_seg) delta[(int64_t)seg_b[s] * 9 + r] = xsep[(seg0 + s) * 9 + r];
  else if (s == n_seg && own_lo > 0 && seg0 > 0) delta[(own_lo - 1) * 9 + r] = xsep[(seg0 - 1) * 9 + r];
}

__global__ void k_la_edges(int64_t lo, int64_t hi, const double* __restrict__ st, double* __restrict__ edge) {
  const int t = threadIdx.x;
  if (t < 10) edge[t] = st[lo * 10 + t];
  else if (t < 20) edge[t] = st[(hi - 1) * 10 + (t - 10)];
}

__global__ void k_la_apply_ghosts(int64_t lo, int64_t hi, int64_t T, int rank, int n_ranks,
                                  const double* __restrict__ edges_all, double* __restrict__ st) {
  const int t = threadIdx.x;
  if (t < 10) {
    if (lo > 0 && rank > 0) st[(lo - 1) * 10 + t] = edges_all[(int64_t)(rank - 1) * 20 + 10 + t];   // neighbour's LAST frame
  } else if (t < 20) {
    if (hi < T && rank + 1 < n_ranks) st[hi * 10 + (t - 10)] = edges_all[(int64_t)(rank + 1) * 20 + (t - 10)];  // FIRST
  }
}

template <typename T>
int la_alloc(vinsat_ctx* ctx, T** p, int64_t n) {
  if (*p) { cudaFree(*p); *p = nullptr; }
  if (cudaMalloc((void**)p, (size_t)std::max<int64_t>(n, 1) * sizeof(T)) != cudaSuccess) {
    cudaGetLastError();
    return set_error(ctx, VINSAT_ENOMEM, "cudaMalloc failed (long-arc buffers)");
  }
  return VINSAT_OK;
}

}  // namespace

extern "C" {

int vinsat_la_num_segments(const vinsat_batch* b) { return b ? (int)b->n_seg : 0; }

int vinsat_la_alloc_reduced(vinsat_batch* b, int64_t S_total, int64_t n_ranks) {
  if (!b || !b->window) return set_error(b ? b->ctx : nullptr, VINSAT_EINVAL, "not a window batch");
  vinsat_ctx* ctx = b->ctx;
  VS_CHECK_ARG(ctx, S_total >= b->n_seg && n_ranks >= 1);
  VS_CUDA(ctx, cudaSetDevice(ctx->device));
  VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  b->S_total = S_total;
  b->n_ranks = n_ranks;
  VS_TRY(la_alloc(ctx, &b->la_gath, S_total * kPack));
  VS_TRY(la_alloc(ctx, &b->la_rsys, S_total * VS_SREC));
  VS_TRY(la_alloc(ctx, &b->la_rlow, S_total * 81));
  VS_TRY(la_alloc(ctx, &b->la_rwrec, S_total * VS_WREC));
  VS_TRY(la_alloc(ctx, &b->la_xsep, S_total * 9));
  VS_TRY(la_alloc(ctx, &b->la_edges_all, n_ranks * 20));
  VS_TRY(la_alloc(ctx, &b->la_chain, 4));
  const int32_t h[4] = {0, (int32_t)S_total, 0, 0};
  VS_CUDA(ctx, cudaMemcpy(b->la_chain, h, sizeof(h), cudaMemcpyHostToDevice));
  // second partition level over the gathered reduced chain (every rank solves it redundantly): with thousands of segments per
  // rank one warp walking S_total separators would dominate the trial
  {
    const Level2Host hp = plan_level2({0}, {(int32_t)S_total}, {0}, level2_min_separators());
    Level2& L = b->la_l2;
    L.n = (int64_t)hp.a.size();
    L.n_chains = (int64_t)hp.red_a.size();
    if (L.n > 0) {
      VS_TRY(la_alloc(ctx, &L.a, L.n)); VS_TRY(la_alloc(ctx, &L.b, L.n)); VS_TRY(la_alloc(ctx, &L.left, L.n));
      VS_TRY(la_alloc(ctx, &L.prob, L.n)); VS_TRY(la_alloc(ctx, &L.has_next, L.n));
      VS_TRY(la_alloc(ctx, &L.red_a, L.n_chains)); VS_TRY(la_alloc(ctx, &L.red_b, L.n_chains)); VS_TRY(la_alloc(ctx, &L.red_prob, L.n_chains));
      VS_TRY(la_alloc(ctx, &L.redrec, L.n * VS_RREC)); VS_TRY(la_alloc(ctx, &L.rsys, L.n * VS_SREC));
      VS_TRY(la_alloc(ctx, &L.rlow, L.n * 81)); VS_TRY(la_alloc(ctx, &L.rwrec, L.n * VS_WREC));
      auto up = [&](int32_t* dst, const std::vector<int32_t>& v) {
        return cudaMemcpy(dst, v.data(), v.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
      };
      VS_CUDA(ctx, up(L.a, hp.a)); VS_CUDA(ctx, up(L.b, hp.b)); VS_CUDA(ctx, up(L.left, hp.left)); VS_CUDA(ctx, up(L.prob, hp.prob));
      VS_CUDA(ctx, up(L.has_next, hp.has_next)); VS_CUDA(ctx, up(L.red_a, hp.red_a)); VS_CUDA(ctx, up(L.red_b, hp.red_b));
      VS_CUDA(ctx, up(L.red_prob, hp.red_prob));
    }
  }
  return VINSAT_OK;
}

int vinsat_la_ptr(vinsat_batch* b, int which, void** ptr, int64_t* count) {
  if (!b || !b->window || !ptr || !count) return set_error(b ? b->ctx : nullptr, VINSAT_EINVAL, "vinsat_la_ptr: bad argument");
  switch (which) {
    case VINSAT_LA_BUF_HIST: *ptr = b->sel_hist; *count = 2048; break;               // uint32
    case VINSAT_LA_BUF_WMAX: *ptr = b->wmax; *count = 1; break;                      // int64 bits of a double >= 0
    case VINSAT_LA_BUF_SUMS: *ptr = b->la_sums; *count = 4; break;                   // float64
    case VINSAT_LA_BUF_PACK: *ptr = b->la_pack; *count = b->n_seg * kPack; break;    // float64
    case VINSAT_LA_BUF_GATHER: *ptr = b->la_gath; *count = b->S_total * kPack; break;
    case VINSAT_LA_BUF_EDGE: *ptr = b->la_edge; *count = 20; break;
    case VINSAT_LA_BUF_EDGES_ALL: *ptr = b->la_edges_all; *count = b->n_ranks * 20; break;
    case VINSAT_LA_BUF_FLAGS: *ptr = b->flags; *count = 4; break;                     // int32: [0] = LM loop still active
    case VINSAT_LA_BUF_LAM_NEXT: *ptr = b->lam_next; *count = 1; break;               // float64
    case VINSAT_LA_BUF_NTRIALS: *ptr = b->ntrials; *count = 1; break;                 // int32
    default: return set_error(b->ctx, VINSAT_EINVAL, "vinsat_la_ptr: unknown buffer %d", which);
  }
  if (!*ptr) return set_error(b->ctx, VINSAT_EINVAL, "vinsat_la_ptr: buffer %d not allocated yet", which);
  return VINSAT_OK;
}

int vinsat_la_stage(vinsat_batch* b, int stage, int64_t i0, int64_t i1, double d0) {
  if (!b || !b->window) return set_error(b ? b->ctx : nullptr, VINSAT_EINVAL, "vinsat_la_stage: not a window batch");
  vinsat_ctx* ctx = b->ctx;
  VS_CUDA(ctx, cudaSetDevice(ctx->device));
  const double qc = 100.0, vc = 100.0;
  switch (stage) {
    case VINSAT_LA_RESID: return launch_obs_residual(b);
    case VINSAT_LA_SELECT_BEGIN: return launch_select_begin(b, i0);               // i0 = 2 * M_total
    case VINSAT_LA_SELECT_HIST: return launch_select_hist(b, (int)i0);
    case VINSAT_LA_SELECT_PICK: return launch_select_pick(b, (int)i0);
    case VINSAT_LA_ASSEMBLE: return launch_obs_assemble(b, d0);                   // d0 = alpha
    case VINSAT_LA_DYNAMICS:
      VS_TRY(launch_dynamics_stm(ctx, b->n_pairs, b->dyn_order, b->st, b->gap, vc, (int)i0, b->drec, nullptr, b->mrec));
      return launch_quat_terms(ctx, b->T, b->st, b->crot, b->gap, qc, b->drec);
    case VINSAT_LA_SYSTEM:
      b->srec_valid = true; b->last_sigma = d0; b->last_initialize = (int)i0; b->have_iter = true;
      return launch_system_build(b, (int)i0, d0, vc);                             // i0 = initialize, d0 = Sigma
    case VINSAT_LA_SUMS_INIT:
      VS_TRY(launch_sum_partials(b, 0, (int)i0, nullptr, b->own_lo, b->own_hi));
      VS_LAUNCH(ctx, F_ACCEPT, k_la_sums_final, 1, 256, 0, (int)ceil_div(b->own_hi - b->own_lo, kSumChunk), 0, b->sum_part, b->wmax,
                b->la_sums);
      return VINSAT_OK;
    case VINSAT_LA_SET_LAM:
      VS_LAUNCH(ctx, F_ACCEPT, k_la_set_lam, 1, 1, 0, d0, b->lam, b->active);
      return VINSAT_OK;
    case VINSAT_LA_SOLVE_INIT: return launch_solve_init_only(b);
    case VINSAT_LA_FORWARD:
      VS_TRY(launch_seg_forward(b));
      VS_LAUNCH(ctx, F_LAYOUT, k_la_pack, ceil_div(b->n_seg * kPack, 256), 256, 0, (int)b->n_seg, b->seg_b, b->redrec,
                b->srec, b->la_pack);
      return VINSAT_OK;
    case VINSAT_LA_REDUCED: {                                                      // i0 = first global segment of this rank
      if (!b->la_gath) return set_error(ctx, VINSAT_EINVAL, "vinsat_la_alloc_reduced has not been called");
      VS_TRY(launch_reduced_packed(b, b->S_total, b->la_gath, b->la_rsys, b->la_rlow, b->la_rwrec, b->la_xsep, b->la_chain));
      VS_LAUNCH(ctx, F_LAYOUT, k_la_scatter, ceil_div((b->n_seg + 1) * 9, 128), 128, 0, (int)b->n_seg, i0, b->own_lo,
                b->seg_b, b->la_xsep, b->delta);
      return VINSAT_OK;
    }
    case VINSAT_LA_BACKSUB: return launch_seg_backsub(b);
    case VINSAT_LA_RETRACT: return launch_retract_only(b);
    case VINSAT_LA_PACK_EDGES:
      VS_LAUNCH(ctx, F_LAYOUT, k_la_edges, 1, 32, 0, b->own_lo, b->own_hi, i0 ? b->st : b->st_new, b->la_edge);   // i0: 1 = current states
      return VINSAT_OK;
    case VINSAT_LA_APPLY_GHOSTS:                                                   // i0 = rank, i1 = 1: current states
      if (!b->la_edges_all) return set_error(ctx, VINSAT_EINVAL, "vinsat_la_alloc_reduced has not been called");
      VS_LAUNCH(ctx, F_LAYOUT, k_la_apply_ghosts, 1, 32, 0, b->own_lo, b->own_hi, b->T, (int)i0, (int)b->n_ranks,
                b->la_edges_all, i1 ? b->st : b->st_new);
      return VINSAT_OK;
    case VINSAT_LA_TRIAL:                                                          // i0 = mode, i1 = initialize
      VS_TRY(launch_obs_trial(b));
      if (!i1)
        VS_TRY(launch_dyn_trial(ctx, b->n_pairs, b->dyn_order, b->st_new, b->crot, b->gap, b->active, b->fprob, qc, vc,
                                (int)i0, b->e_dyn, nullptr));
      return VINSAT_OK;
    case VINSAT_LA_SUMS_TRIAL:
      if (i1) VS_CUDA(ctx, cudaMemsetAsync(b->la_sums, 0, 2 * sizeof(double), ctx->stream));   // i1: drop the (already reduced) init sums
      VS_TRY(launch_sum_partials(b, 1, (int)i0, nullptr, b->own_lo, b->own_hi));
      VS_LAUNCH(ctx, F_ACCEPT, k_la_sums_final, 1, 256, 0, (int)ceil_div(b->own_hi - b->own_lo, kSumChunk), 1, b->sum_part, b->wmax,
                b->la_sums);
      return VINSAT_OK;
    case VINSAT_LA_COMMIT:
      std::swap(b->st, b->st_new);
      return VINSAT_OK;
    case VINSAT_LA_BEGIN_ITER:                                                     // d0 >= 0: also (re)sets lam_next from the host
      if (d0 > 0.0) VS_LAUNCH(ctx, F_ACCEPT, k_la_set_lam, 1, 1, 0, d0, b->lam_next, b->active);
      VS_LAUNCH(ctx, F_ACCEPT, k_la_begin_iter, 1, 1, 0, b->lam_next, b->lam, b->active, b->ntrials, b->flags);
      return VINSAT_OK;
    case VINSAT_LA_ACCEPT:                                                         // i1 = number of residual components, d0 = sqrt(Sigma)
      VS_LAUNCH(ctx, F_ACCEPT, k_la_accept, 1, 1, 0, b->la_sums, (double)i1, d0, b->init_res, b->lam, b->lam_next, b->ntrials,
                b->flags);
      return VINSAT_OK;
    case VINSAT_LA_COMMIT_COPY:                                                    // graph mode: fixed buffer roles, copy instead of swap
      VS_CUDA(ctx, cudaMemcpyAsync(b->st, b->st_new, b->T * 10 * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
      return VINSAT_OK;
    default: return set_error(ctx, VINSAT_EINVAL, "vinsat_la_stage: unknown stage %d", stage);
  }
}

}  // extern "C"
