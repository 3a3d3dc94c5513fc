// Device-resident batch of OD problems: allocation, upload, BA iteration schedule (host side of a5-a7).
#include <math.h>
#include <string.h>

#include <algorithm>

#include "launch.h"

using namespace vs;

namespace {

template <typename T>
int dev_alloc(vinsat_ctx* ctx, T** p, int64_t n) {
  if (n < 1) n = 1;
  if (ctx->arena_on) {
    const size_t bytes = ((size_t)n * sizeof(T) + 255) & ~(size_t)255;
    if (ctx->arena_off + bytes > ctx->arena_bytes)
      return set_error(ctx, VINSAT_ENOMEM, "batch arena of %zu bytes exhausted", ctx->arena_bytes);
    *p = (T*)(ctx->arena + ctx->arena_off);
    ctx->arena_off += bytes;
    return VINSAT_OK;
  }
  cudaError_t e = cudaMalloc((void**)p, (size_t)n * sizeof(T));
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_error(ctx, VINSAT_ENOMEM, "cudaMalloc of %lld bytes failed: %s", (long long)(n * sizeof(T)),
                     cudaGetErrorString(e));
  }
  return VINSAT_OK;
}

#define VS_TRY(expr)            \
  do {                          \
    int _rc = (expr);           \
    if (_rc != VINSAT_OK) return _rc; \
  } while (0)

void free_all(vinsat_batch* b) {
  void* ptrs[] = {b->st, b->st_new, b->intr, b->crot, b->gap, b->fprob, b->dyn_order, b->obs_start, b->grec, b->drec, b->mrec,
                  b->srec, b->wrec, b->delta, b->zeros, b->e_obs, b->e_dyn, b->X, b->uv, b->conf, b->oframe, b->r, b->r_next, b->wu,
                  b->d_frame_off, b->d_obs_off, b->c_obs, b->wmax, b->lam, b->lam_next, b->lam32_last, b->init_res,
                  b->active, b->ntrials, b->sel_prefix, b->sel_rank, b->sel_hist, b->flags, b->seg_a, b->seg_b,
                  b->seg_left, b->seg_prob, b->seg_has_next, b->pl_a, b->pl_b, b->pl_prob, b->red_a, b->red_b,
                  b->bb_a, b->bb_e, b->bb_dir, b->bb_prob, b->bb_mid, b->midrec,
                  b->redrec, b->rsys, b->rlow, b->rwrec, b->la_pack, b->la_gath, b->la_rsys, b->la_rlow, b->la_rwrec,
                  b->la_xsep, b->la_sums, b->la_edge, b->la_edges_all, b->la_chain, b->sum_part, b->sum_chunk_off,
                  b->l2.a, b->l2.b, b->l2.left, b->l2.prob, b->l2.has_next, b->l2.red_a, b->l2.red_b, b->l2.red_prob,
                  b->l2.redrec, b->l2.rsys, b->l2.rlow, b->l2.rwrec, b->xsep};
  if (!b->in_arena)
    for (void* p : ptrs)
      if (p) cudaFree(p);
  if (b->J) cudaFree(b->J);
  for (void* p : {(void*)b->la_l2.a, (void*)b->la_l2.b, (void*)b->la_l2.left, (void*)b->la_l2.prob, (void*)b->la_l2.has_next,
                  (void*)b->la_l2.red_a, (void*)b->la_l2.red_b, (void*)b->la_l2.red_prob, (void*)b->la_l2.redrec,
                  (void*)b->la_l2.rsys, (void*)b->la_l2.rlow, (void*)b->la_l2.rwrec})
    if (p) cudaFree(p);            // vinsat_la_alloc_reduced: plain cudaMalloc
  for (void* p : {(void*)b->mc_st_true, (void*)b->mc_uv_true, (void*)b->mc_vel_true, (void*)b->mc_err, (void*)b->pr_st,
                  (void*)b->pr_Hs, (void*)b->pr_Hr, (void*)b->e_pr_init, (void*)b->e_pr})
    if (p) cudaFree(p);            // allocated lazily with cudaMalloc, never from the arena
  for (auto& kv : b->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  b->graphs.clear();
  if (b->h_flags) cudaFreeHost(b->h_flags);
}

int validate_desc(vinsat_ctx* ctx, const vinsat_problem_desc* d) {
  VS_CHECK_ARG(ctx, d != nullptr);
  VS_CHECK_ARG(ctx, d->n_problems >= 1);
  VS_CHECK_ARG(ctx, d->frame_off && d->obs_off);
  VS_CHECK_ARG(ctx, d->frame_off[0] == 0 && d->obs_off[0] == 0);
  for (int64_t p = 0; p < d->n_problems; p++) {
    VS_CHECK_ARG(ctx, d->frame_off[p + 1] >= d->frame_off[p]);
    VS_CHECK_ARG(ctx, d->obs_off[p + 1] >= d->obs_off[p]);
  }
  const int64_t T = d->frame_off[d->n_problems], M = d->obs_off[d->n_problems];
  VS_CHECK_ARG(ctx, T >= 1 && T < (1ll << 31) - 64 && M < (1ll << 31) - 64);
  VS_CHECK_ARG(ctx, d->states && d->intrinsics && d->cum_rot && d->time_idx);
  VS_CHECK_ARG(ctx, M == 0 || (d->landmarks_xyz && d->landmarks_uv && d->confidences && d->ii));
  return VINSAT_OK;
}

// Frame-side indexing on the host (O(T)): gaps, problem ids, and the longest-gap-first pair order.
int build_frame_index(vinsat_batch* b, const vinsat_problem_desc* d, std::vector<int32_t>& gap,
                      std::vector<int32_t>& fprob, std::vector<int32_t>& order) {
  vinsat_ctx* ctx = b->ctx;
  const int64_t T = b->T;
  gap.assign(T, 0);
  fprob.assign(T, 0);
  int32_t gmax = 0;
  for (int64_t p = 0; p < b->P; p++) {
    for (int64_t f = d->frame_off[p]; f < d->frame_off[p + 1]; f++) {
      fprob[f] = (int32_t)p;
      if (f + 1 < d->frame_off[p + 1]) {
        const int64_t g = d->time_idx[f + 1] - d->time_idx[f];
        if (g <= 0 || g > 100000000)
          return set_error(ctx, VINSAT_EINVAL, "time_idx must be strictly increasing inside a problem (frame %lld)",
                           (long long)f);
        gap[f] = (int32_t)g;
        gmax = std::max(gmax, gap[f]);
      }
    }
  }
  // counting sort by gap, descending (stable)
  const int32_t nb = std::min<int32_t>(gmax, 1 << 20) + 1;
  std::vector<int64_t> cnt(nb + 1, 0);
  auto bucket = [&](int32_t g) { return std::min<int32_t>(g, nb - 1); };
  int64_t np = 0;
  for (int64_t f = 0; f < T; f++)
    if (gap[f] > 0) { cnt[bucket(gap[f])]++; np++; }
  std::vector<int64_t> start(nb + 1, 0);
  int64_t run = 0;
  for (int32_t g = nb - 1; g >= 0; g--) { start[g] = run; run += cnt[g]; }
  order.assign(np, 0);
  for (int64_t f = 0; f < T; f++)
    if (gap[f] > 0) order[start[bucket(gap[f])]++] = (int32_t)f;
  b->n_pairs = np;
  return VINSAT_OK;
}

// Cut every problem into segments for the partitioned solve (kernels_chain.cu).  Depends only on frame_off.
struct Segmentation {
  std::vector<int32_t> a, b, left, prob, has_next, pl_a, pl_b, pl_prob, red_a, red_b;
  std::vector<int32_t> bb_a, bb_e, bb_dir, bb_prob, bb_mid;
  Level2Host l2;
  bool partitioned = false;
};

Segmentation make_segments(const vinsat_ctx* ctx, int64_t P, const int64_t* frame_off, const vinsat_batch* wb = nullptr) {
  Segmentation s;
  if (wb && wb->window) {
    // frame-window sharded arc: segments over the OWNED frames only; the left ghost (if any) is the left separator
    const int64_t lo = wb->own_lo, hi = wb->own_hi, Tp = hi - lo;
    const int64_t S = std::max<int64_t>(1, std::min<int64_t>(wb->forced_segments, Tp));
    s.pl_a.push_back(0); s.pl_b.push_back((int32_t)frame_off[1]); s.pl_prob.push_back(0);
    s.red_a.push_back(0);
    for (int64_t k = 0; k < S; k++) {
      const int64_t a = lo + (Tp * k) / S, b = lo + (Tp * (k + 1)) / S;
      s.a.push_back((int32_t)a);
      s.b.push_back((int32_t)(b - 1));
      s.left.push_back(a > 0 ? (int32_t)(a - 1) : -1);
      s.prob.push_back(0);
      s.has_next.push_back(k + 1 < S ? 1 : 0);
    }
    s.red_b.push_back((int32_t)S);
    s.partitioned = true;
    return s;
  }
  const int64_t T = frame_off[P];
  int64_t forced = 0;
  if (const char* e = getenv("VINSAT_SEG_LEN")) forced = atoll(e);
  const double target_chains = 8.0 * ctx->sm_count * 4;      // ~8 warps of 4 chains per SM
  for (int64_t p = 0; p < P; p++) {
    const int64_t f0 = frame_off[p], f1 = frame_off[p + 1], Tp = f1 - f0;
    s.pl_a.push_back((int32_t)f0); s.pl_b.push_back((int32_t)f1); s.pl_prob.push_back((int32_t)p);
    {
      const int32_t m = Tp > 0 ? (int32_t)(f0 + Tp / 2) : -1;
      s.bb_a.push_back((int32_t)f0);     s.bb_e.push_back(Tp > 0 ? m : (int32_t)f0);     s.bb_dir.push_back(1);
      s.bb_a.push_back((int32_t)f1 - 1); s.bb_e.push_back(Tp > 0 ? m : (int32_t)f1 - 1); s.bb_dir.push_back(-1);
      s.bb_prob.push_back((int32_t)p); s.bb_prob.push_back((int32_t)p);
      s.bb_mid.push_back(m); s.bb_mid.push_back(m);
    }
    s.red_a.push_back((int32_t)s.a.size());
    if (Tp > 0) {
      int64_t S;
      if (forced > 0) S = (Tp + forced - 1) / forced;
      else {
        S = (int64_t)llround((double)Tp * target_chains / (double)std::max<int64_t>(T, 1));
        S = std::min<int64_t>(S, (int64_t)sqrt((double)Tp));
        // long problems: as many segments as the GPU runs chains at once (8 one-warp CTAs x 3 chains per SM), each at least
        // 24 frames long; the reduced chain over them is then partitioned a second time (Level2) instead of being walked by
        // one warp.  Sequential depth of a 2.4 M-frame arc: 676 + 60 + 60 eliminations instead of 1549 + 1549.
        {
          const double one_wave = 8.0 * ctx->sm_count * 3;
          const int64_t S2 = std::min<int64_t>((int64_t)llround((double)Tp * one_wave / (double)std::max<int64_t>(T, 1)), Tp / 24);
          if (S2 >= level2_min_separators() && S2 > S) S = S2;
        }
        if (Tp < 64) S = 1;
        // measured on B200 (T=1000, two-sided fused sweep vs partitioned sweep + materialised system): 14.7 k solves/s
        // either way at P = 256, 21.2 k vs 16.3 k at P = 512, 9.3 k vs 11.8 k at P = 128 => partition below ~1.75
        // problems per SM
        if (4 * P >= 7 * (int64_t)ctx->sm_count) S = 1;
      }
      S = std::max<int64_t>(1, std::min<int64_t>(S, Tp));
      if (S > 1) s.partitioned = true;
      for (int64_t k = 0; k < S; k++) {
        const int64_t lo = f0 + (Tp * k) / S, hi = f0 + (Tp * (k + 1)) / S;   // frames [lo, hi), separator hi-1
        s.a.push_back((int32_t)lo);
        s.b.push_back((int32_t)(hi - 1));
        s.left.push_back(k > 0 ? (int32_t)(lo - 1) : -1);
        s.prob.push_back((int32_t)p);
        s.has_next.push_back(k + 1 < S ? 1 : 0);
      }
    }
    s.red_b.push_back((int32_t)s.a.size());
  }
  if (s.partitioned) s.l2 = plan_level2(s.red_a, s.red_b, s.pl_prob, level2_min_separators());
  return s;
}

int do_upload(vinsat_batch* b, const vinsat_problem_desc* d) {
  vinsat_ctx* ctx = b->ctx;
  VS_CUDA(ctx, cudaSetDevice(ctx->device));
  if (d->n_problems != b->P || d->frame_off[d->n_problems] != b->T || d->obs_off[d->n_problems] != b->M)
    return set_error(ctx, VINSAT_EINVAL, "vinsat_batch_upload: sizes differ from the batch (P,T,M)");
  cudaStream_t s = ctx->stream;
  const int64_t P = b->P, T = b->T, M = b->M;
  // captured iteration graphs bake in launch shapes derived from the offsets: drop them if the layout changes
  if (!b->graphs.empty() && (!std::equal(d->frame_off, d->frame_off + P + 1, b->frame_off.begin()) ||
                             !std::equal(d->obs_off, d->obs_off + P + 1, b->obs_off.begin()))) {
    VS_CUDA(ctx, cudaStreamSynchronize(s));
    for (auto& kv : b->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    b->graphs.clear();
    b->graph_warm.clear();
    b->graph_bad.clear();
  }
  b->frame_off.assign(d->frame_off, d->frame_off + P + 1);
  b->obs_off.assign(d->obs_off, d->obs_off + P + 1);
  b->max_obs_per_problem = 0;
  for (int64_t p = 0; p < P; p++) b->max_obs_per_problem = std::max(b->max_obs_per_problem, d->obs_off[p + 1] - d->obs_off[p]);
  std::vector<int32_t> gap, fprob, order;
  VS_TRY(build_frame_index(b, d, gap, fprob, order));
  VS_CUDA(ctx, cudaMemcpyAsync(b->d_frame_off, d->frame_off, (P + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s));
  VS_CUDA(ctx, cudaMemcpyAsync(b->d_obs_off, d->obs_off, (P + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s));
  std::vector<int32_t> chunk_off(P + 1, 0);
  for (int64_t p = 0; p < P; p++)
    chunk_off[p + 1] = chunk_off[p] + (int32_t)ceil_div(d->frame_off[p + 1] - d->frame_off[p], kSumChunk);
  b->sum_chunks = chunk_off[P];
  VS_CUDA(ctx, cudaMemcpyAsync(b->sum_chunk_off, chunk_off.data(), (P + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, s));
  VS_CUDA(ctx, cudaMemcpyAsync(b->gap, gap.data(), T * sizeof(int32_t), cudaMemcpyHostToDevice, s));
  VS_CUDA(ctx, cudaMemcpyAsync(b->fprob, fprob.data(), T * sizeof(int32_t), cudaMemcpyHostToDevice, s));
  if (b->n_pairs)
    VS_CUDA(ctx, cudaMemcpyAsync(b->dyn_order, order.data(), b->n_pairs * sizeof(int32_t), cudaMemcpyHostToDevice, s));
  VS_CUDA(ctx, cudaMemcpyAsync(b->st, d->states, T * 10 * sizeof(double), cudaMemcpyHostToDevice, s));
  VS_CUDA(ctx, cudaMemcpyAsync(b->intr, d->intrinsics, T * 4 * sizeof(double), cudaMemcpyHostToDevice, s));
  VS_CUDA(ctx, cudaMemcpyAsync(b->crot, d->cum_rot, T * 4 * sizeof(double), cudaMemcpyHostToDevice, s));
  {
    Segmentation sg = make_segments(ctx, P, d->frame_off, b);
    if ((int64_t)sg.a.size() != b->n_seg || (int64_t)sg.l2.a.size() != b->l2.n || (int64_t)sg.l2.red_a.size() != b->l2.n_chains)
      return set_error(ctx, VINSAT_EINVAL, "vinsat_batch_upload: segmentation differs (frame_off must not change)");
    b->partitioned = sg.partitioned;
    auto up = [&](int32_t* dst, const std::vector<int32_t>& v) {
      return v.empty() ? cudaSuccess : cudaMemcpyAsync(dst, v.data(), v.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s);
    };
    VS_CUDA(ctx, up(b->seg_a, sg.a)); VS_CUDA(ctx, up(b->seg_b, sg.b)); VS_CUDA(ctx, up(b->seg_left, sg.left));
    VS_CUDA(ctx, up(b->seg_prob, sg.prob)); VS_CUDA(ctx, up(b->seg_has_next, sg.has_next));
    VS_CUDA(ctx, up(b->pl_a, sg.pl_a)); VS_CUDA(ctx, up(b->pl_b, sg.pl_b)); VS_CUDA(ctx, up(b->pl_prob, sg.pl_prob));
    VS_CUDA(ctx, up(b->red_a, sg.red_a)); VS_CUDA(ctx, up(b->red_b, sg.red_b));
    VS_CUDA(ctx, up(b->bb_a, sg.bb_a)); VS_CUDA(ctx, up(b->bb_e, sg.bb_e)); VS_CUDA(ctx, up(b->bb_dir, sg.bb_dir));
    VS_CUDA(ctx, up(b->bb_prob, sg.bb_prob)); VS_CUDA(ctx, up(b->bb_mid, sg.bb_mid));
    if (b->l2.n > 0) {
      VS_CUDA(ctx, up(b->l2.a, sg.l2.a)); VS_CUDA(ctx, up(b->l2.b, sg.l2.b)); VS_CUDA(ctx, up(b->l2.left, sg.l2.left));
      VS_CUDA(ctx, up(b->l2.prob, sg.l2.prob)); VS_CUDA(ctx, up(b->l2.has_next, sg.l2.has_next));
      VS_CUDA(ctx, up(b->l2.red_a, sg.l2.red_a)); VS_CUDA(ctx, up(b->l2.red_b, sg.l2.red_b));
      VS_CUDA(ctx, up(b->l2.red_prob, sg.l2.red_prob));
    }
    VS_CUDA(ctx, cudaStreamSynchronize(s));     // the host vectors die at the end of this scope
  }
  VS_CUDA(ctx, cudaMemsetAsync(b->flags, 0, 4 * sizeof(int32_t), s));
  if (M > 0) {
    // stage AoS observation arrays + ii in scratch, then transpose to SoA / index on the device
    const size_t need = (size_t)M * (3 + 2) * sizeof(double) + (size_t)M * sizeof(int64_t);
    char* sc = (char*)ctx_scratch(ctx, need);
    if (!sc) return set_error(ctx, VINSAT_ENOMEM, "scratch allocation of %zu bytes failed", need);
    double* sx = (double*)sc;
    double* suv = sx + 3 * M;
    int64_t* sii = (int64_t*)(suv + 2 * M);
    VS_CUDA(ctx, cudaMemcpyAsync(sx, d->landmarks_xyz, M * 3 * sizeof(double), cudaMemcpyHostToDevice, s));
    VS_CUDA(ctx, cudaMemcpyAsync(suv, d->landmarks_uv, M * 2 * sizeof(double), cudaMemcpyHostToDevice, s));
    VS_CUDA(ctx, cudaMemcpyAsync(sii, d->ii, M * sizeof(int64_t), cudaMemcpyHostToDevice, s));
    VS_CUDA(ctx, cudaMemcpyAsync(b->conf, d->confidences, M * sizeof(double), cudaMemcpyHostToDevice, s));
    VS_TRY(launch_aos_to_soa(ctx, sx, b->X, M, 3));
    VS_TRY(launch_aos_to_soa(ctx, suv, b->uv, M, 2));
    VS_TRY(launch_obs_index(b, sii));
  } else {
    VS_TRY(launch_obs_index(b, nullptr));
  }
  VS_CUDA(ctx, cudaMemcpyAsync(b->h_flags, b->flags, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  VS_CUDA(ctx, cudaStreamSynchronize(s));
  if (b->h_flags[1] & 1) return set_error(ctx, VINSAT_EINVAL, "ii out of range for its problem's frame count");
  if (b->h_flags[1] & 2) return set_error(ctx, VINSAT_EINVAL, "ii must be non-decreasing inside each problem");
  b->have_iter = false;
  b->r_valid = false;
  return VINSAT_OK;
}

}  // namespace

extern "C" {

static int create_impl(vinsat_ctx* ctx, const vinsat_problem_desc* d, int64_t own_lo, int64_t own_hi, int64_t n_segments,
                       vinsat_batch** out) {
  if (!ctx || !out) return set_error(ctx, VINSAT_EINVAL, "vinsat_batch_create: NULL argument");
  *out = nullptr;
  VS_TRY(validate_desc(ctx, d));
  VS_CUDA(ctx, cudaSetDevice(ctx->device));
  vinsat_batch* b = new vinsat_batch();
  b->ctx = ctx;
  b->in_arena = ctx->arena_on;
  if (n_segments > 0) {
    if (d->n_problems != 1 || own_lo < 0 || own_hi <= own_lo || own_hi > d->frame_off[1] || own_lo > 1 ||
        own_hi < d->frame_off[1] - 1) {
      delete b;
      return set_error(ctx, VINSAT_EINVAL, "window: one problem, owned range plus at most one ghost frame per side");
    }
    b->window = true;
    b->own_lo = own_lo;
    b->own_hi = own_hi;
    b->forced_segments = n_segments;
  }
  b->P = d->n_problems;
  b->T = d->frame_off[b->P];
  b->M = d->obs_off[b->P];
  const int64_t P = b->P, T = b->T, M = b->M;
  {
    const Segmentation sg0 = make_segments(ctx, P, d->frame_off, b);
    b->n_seg = (int64_t)sg0.a.size();
    b->l2.n = (int64_t)sg0.l2.a.size();
    b->l2.n_chains = (int64_t)sg0.l2.red_a.size();
  }
  const int64_t NS = b->n_seg, N2 = b->l2.n, C2 = b->l2.n_chains;
  int rc = VINSAT_OK;
#define A(ptr, n) if (rc == VINSAT_OK) rc = dev_alloc(ctx, &b->ptr, (n))
  A(seg_a, NS); A(seg_b, NS); A(seg_left, NS); A(seg_prob, NS); A(seg_has_next, NS);
  A(pl_a, P); A(pl_b, P); A(pl_prob, P); A(red_a, P); A(red_b, P);
  A(bb_a, 2 * P); A(bb_e, 2 * P); A(bb_dir, 2 * P); A(bb_prob, 2 * P); A(bb_mid, 2 * P); A(midrec, 2 * P * 96);
  A(redrec, NS * VS_RREC); A(rsys, NS * VS_SREC); A(rlow, NS * 81); A(rwrec, NS * VS_WREC);
  if (N2 > 0) {
    A(l2.a, N2); A(l2.b, N2); A(l2.left, N2); A(l2.prob, N2); A(l2.has_next, N2);
    A(l2.red_a, C2); A(l2.red_b, C2); A(l2.red_prob, C2);
    A(l2.redrec, N2 * VS_RREC); A(l2.rsys, N2 * VS_SREC); A(l2.rlow, N2 * 81); A(l2.rwrec, N2 * VS_WREC);
    A(xsep, NS * 9);
  }
  if (b->window) { A(la_pack, NS * (VS_RREC + VS_SREC)); A(la_sums, 4); A(la_edge, 20); }
  A(st, T * 10); A(st_new, T * 10); A(intr, T * 4); A(crot, T * 4); A(gap, T); A(fprob, T); A(dyn_order, T);
  A(obs_start, T + 1); A(grec, T * VS_GREC); A(drec, T * VS_DREC); A(mrec, T * VS_MREC); A(srec, T * VS_SREC); A(wrec, T * VS_WREC);
  A(delta, T * 9); A(e_obs, T); A(e_dyn, T); A(zeros, 64);
  A(X, M * 3); A(uv, M * 2); A(conf, M); A(oframe, M); A(r, M * 2); A(r_next, M * 2); A(wu, M);
  A(d_frame_off, P + 1); A(d_obs_off, P + 1); A(c_obs, P); A(wmax, P); A(lam, P); A(lam_next, P); A(lam32_last, P);
  A(init_res, P); A(active, P); A(ntrials, P); A(sel_prefix, P); A(sel_rank, P); A(sel_hist, P * 2048); A(flags, 4);
  b->sum_chunks_cap = T / kSumChunk + P + 1;
  A(sum_part, b->sum_chunks_cap * 3); A(sum_chunk_off, P + 1);
#undef A
  if (rc == VINSAT_OK && cudaMallocHost((void**)&b->h_flags, 4 * sizeof(int32_t)) != cudaSuccess) {
    cudaGetLastError();
    rc = set_error(ctx, VINSAT_ENOMEM, "cudaMallocHost failed");
  }
  if (rc == VINSAT_OK) {
    cudaMemsetAsync(b->sel_hist, 0, (size_t)P * 2048 * sizeof(unsigned int), ctx->stream);
    cudaMemsetAsync(b->e_obs, 0, (size_t)T * sizeof(double), ctx->stream);
    cudaMemsetAsync(b->zeros, 0, 64 * sizeof(double), ctx->stream);
    cudaMemsetAsync(b->e_dyn, 0, (size_t)T * sizeof(double), ctx->stream);
    cudaMemsetAsync(b->drec, 0, (size_t)T * VS_DREC * sizeof(double), ctx->stream);
    rc = do_upload(b, d);
  }
  if (rc != VINSAT_OK) {
    free_all(b);
    delete b;
    return rc;
  }
  *out = b;
  return VINSAT_OK;
}

int vinsat_batch_create(vinsat_ctx* ctx, const vinsat_problem_desc* d, vinsat_batch** out) {
  return create_impl(ctx, d, 0, 0, 0, out);
}

int vinsat_batch_create_window(vinsat_ctx* ctx, const vinsat_problem_desc* d, int64_t own_lo, int64_t own_hi,
                               int64_t n_segments, vinsat_batch** out) {
  if (n_segments < 1) return set_error(ctx, VINSAT_EINVAL, "vinsat_batch_create_window: n_segments must be >= 1");
  return create_impl(ctx, d, own_lo, own_hi, n_segments, out);
}

int vinsat_batch_upload(vinsat_batch* b, const vinsat_problem_desc* d) {
  if (!b) return set_error(nullptr, VINSAT_EINVAL, "vinsat_batch_upload: NULL batch");
  VS_TRY(validate_desc(b->ctx, d));
  return do_upload(b, d);
}

int vinsat_batch_destroy(vinsat_batch* b) {
  if (!b) return VINSAT_OK;
  cudaSetDevice(b->ctx->device);
  cudaStreamSynchronize(b->ctx->stream);
  free_all(b);
  delete b;
  return VINSAT_OK;
}

int vinsat_batch_set_states(vinsat_batch* b, int mem, const double* states) {
  if (!b || !states) return set_error(b ? b->ctx : nullptr, VINSAT_EINVAL, "vinsat_batch_set_states: NULL argument");
  vinsat_ctx* ctx = b->ctx;
  VS_CUDA(ctx, cudaSetDevice(ctx->device));
  VS_CUDA(ctx, cudaMemcpyAsync(b->st, states, b->T * 10 * sizeof(double),
                               mem == VINSAT_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                               ctx->stream));
  if (mem != VINSAT_MEM_DEVICE) VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  b->r_valid = false;
  return VINSAT_OK;
}

int vinsat_batch_get_states(vinsat_batch* b, int mem, double* states_out) {
  if (!b || !states_out) return set_error(b ? b->ctx : nullptr, VINSAT_EINVAL, "vinsat_batch_get_states: NULL argument");
  vinsat_ctx* ctx = b->ctx;
  VS_CUDA(ctx, cudaSetDevice(ctx->device));
  VS_CUDA(ctx, cudaMemcpyAsync(states_out, b->st, b->T * 10 * sizeof(double),
                               mem == VINSAT_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                               ctx->stream));
  VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return VINSAT_OK;
}

// Wait for the "still active" counter of the trial in flight.  The host polls the pinned word the trial's
// device-to-host copy overwrites (h_flags[0] was set to -1 before the launch) instead of entering
// cudaStreamSynchronize: with 8 ranks on one host the driver's blocking wait cost ~0.4 ms per trial.
static int wait_flag(vinsat_batch* b) {
  vinsat_ctx* ctx = b->ctx;
  static const bool no_spin = getenv("VINSAT_NO_SPIN_WAIT") != nullptr;
  if (!no_spin) {
    volatile int32_t* f = b->h_flags;
    for (int64_t spins = 0;; spins++) {
      if (f[0] != -1) return VINSAT_OK;
      // every ~64k polls ask the driver: a finished stream (the copy has landed or will never land) or a faulted
      // one ends the spin at once instead of burning a core until a timeout
      if ((spins & 0xffff) == 0xffff) {
        const cudaError_t q = cudaStreamQuery(ctx->stream);
        if (q == cudaSuccess) return VINSAT_OK;
        if (q != cudaErrorNotReady)
          return set_error(ctx, VINSAT_ECUDA, "stream failed while waiting for an LM trial: %s", cudaGetErrorString(q));
      }
#if defined(__x86_64__) || defined(__i386__)
      __builtin_ia32_pause();
#endif
    }
  }
  VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return VINSAT_OK;
}

// One LM trial: solve, retract, trial residuals, accept test, copy of the "still active" counter to the host.
static int issue_trial(vinsat_batch* b, int initialize, int mode, double Sigma, double quat_coeff, double vel_coeff,
                       int32_t* h_dst = nullptr) {
  vinsat_ctx* ctx = b->ctx;
  const bool reg = b->reg_iter;
  VS_TRY(launch_solve_retract(b, initialize));
  VS_TRY(launch_obs_trial(b));
  // BA_reg evaluates the trial's dynamics residual with quat_coeff_prior = 1 and its prior with (vel, quat) coefficients
  // (1, 100) (BA_filtering.py:169-170): reproduced as is
  if (!initialize)
    VS_TRY(launch_dyn_trial(ctx, b->n_pairs, b->dyn_order, b->st_new, b->crot, b->gap, b->active, b->fprob,
                            reg ? 1.0 : quat_coeff, vel_coeff, mode, b->e_dyn, nullptr, b->gate_arg));
  if (reg) VS_TRY(launch_prior_trial(b, 1.0, vel_coeff));
  if (b->gate_arg) VS_TRY(launch_gated_zero(b, nullptr, 0, b->flags, 1));
  else VS_CUDA(ctx, cudaMemsetAsync(b->flags, 0, sizeof(int32_t), ctx->stream));
  VS_TRY(launch_accept(b, initialize, Sigma, reg ? b->e_pr : nullptr));
  if (b->gate_arg) VS_TRY(launch_gate_publish(b));
  VS_CUDA(ctx, cudaMemcpyAsync(h_dst ? h_dst : b->h_flags, b->flags, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  return VINSAT_OK;
}

// Linearisation + first trial of one BA() call (everything that does not depend on a host decision).
static int issue_iteration_head(vinsat_batch* b, int iter, int initialize, int mode, const double* lam_dev_in,
                                bool have_residuals, int32_t* h_dst = nullptr) {
  vinsat_ctx* ctx = b->ctx;
  const double quat_coeff = 100.0, vel_coeff = 100.0;                                  // BA_filtering.py:11-12
  const double alpha = std::min(std::max(1.0 - (2.0 * ((double)iter / 5.0) - 1.0), 1.0), 2.0);   // :22
  const double it1 = (double)iter + 1.0;
  const double Sigma = std::min(10000.0 * it1 * it1, 1000000.0);                       // :26
  if (!have_residuals) VS_TRY(launch_obs_residual(b));
  VS_TRY(launch_select_median(b));
  VS_TRY(launch_obs_assemble(b, alpha));
  if (!initialize) {
    VS_TRY(launch_dynamics_stm(ctx, b->n_pairs, b->dyn_order, b->st, b->gap, vel_coeff, mode, b->drec, nullptr, b->mrec,
                               b->gate_arg));
    VS_TRY(launch_quat_terms(ctx, b->T, b->st, b->crot, b->gap, quat_coeff, b->drec, b->gate_arg));
  }
  // initialize phase: block-diagonal system, solved straight from the observation records (k_solve_init);
  // the full records are only materialised on demand (last_hessian / debug_fetch).
  // The plain (unpartitioned) sweep builds its columns on the fly from the per-frame records (fused system
  // build); the partitioned sweep and the diagnostics read the materialised records.
  if (!initialize && !b->fused_system) VS_TRY(launch_system_build(b, 0, Sigma, vel_coeff));
  if (b->reg_iter) VS_TRY(launch_prior_linearize(b, 1.0, 1.0));      // prior_gpu(.., quat_coeff_prior = 1, vel_coeff_prior = 1, ..) (:122)
  VS_TRY(launch_init_residual(b, initialize, Sigma, 0.0, lam_dev_in, b->reg_iter ? b->e_pr_init : nullptr));
  return issue_trial(b, initialize, mode, Sigma, quat_coeff, vel_coeff, h_dst);
}

// One BA() call for every problem of the batch; lam_dev_in holds lamda_init per problem on the device.
static int ba_iterate_device(vinsat_batch* b, int iter, int initialize, int mode, const double* lam_dev_in) {
  vinsat_ctx* ctx = b->ctx;
  const double quat_coeff = 100.0, vel_coeff = 100.0;
  const double it1 = (double)iter + 1.0;
  const double Sigma = std::min(10000.0 * it1 * it1, 1000000.0);
  // the last LM trial of the previous call already evaluated uv - project(states) at the returned states
  // (bit-identical arithmetic), so its residuals are reused instead of projecting every observation again
  const bool have_residuals = b->r_valid;
  if (have_residuals) std::swap(b->r, b->r_next);
  static const bool no_fuse = getenv("VINSAT_NO_FUSED_SYSTEM") != nullptr || getenv("VINSAT_ONE_SIDED_SWEEP") != nullptr;
  b->fused_system = !initialize && !b->partitioned && !no_fuse && !b->reg_iter;      // BA_reg adds its blocks to the records
  b->srec_valid = !initialize && !b->fused_system;
  b->last_sigma = Sigma;
  b->cur_sigma = Sigma;
  b->cur_vc = vel_coeff;

  // Small batches are launch bound: the head of the iteration (10-12 launches) is replayed as ONE CUDA graph.  A key
  // is captured the second time it is seen, so that every one-time host action (function attributes, scratch
  // growth) has happened.
  static const bool no_graph = getenv("VINSAT_NO_GRAPH") != nullptr;
  bool done = false;
  b->h_flags[0] = -1;                 // sentinel: overwritten by the trial's device-to-host copy of the counter
  // Measured on B200: P = 64 x T = 1000 gains 24 % from the replay (launch bound), P = 1024 LOSES 5 % (20 distinct
  // graphs of long kernels: the per-graph launch cost exceeds the 10 stream launches it replaces) => small batches only.
  static const int64_t graph_max_frames = getenv("VINSAT_GRAPH_MAX_FRAMES") ? atoll(getenv("VINSAT_GRAPH_MAX_FRAMES")) : 200000;
  if (!no_graph && b->T <= graph_max_frames && !ctx->timing && !b->window && lam_dev_in == b->lam_next && !b->reg_iter) {
    if (!b->st_base) { b->st_base = b->st; b->r_base = b->r; }
    const uint64_t key = (uint64_t)(iter & 0xffff) | ((uint64_t)(initialize ? 1 : 0) << 16) | ((uint64_t)(mode & 0xf) << 17) |
                         ((uint64_t)(b->st == b->st_base ? 1 : 0) << 21) | ((uint64_t)(b->r == b->r_base ? 1 : 0) << 22) |
                         ((uint64_t)(have_residuals ? 1 : 0) << 23) | ((uint64_t)(b->fused_system ? 1 : 0) << 24);
    auto it = b->graphs.find(key);
    // the legacy / per-thread default streams cannot be captured: plain launches there
    const bool capturable = ctx->stream != nullptr && ctx->stream != cudaStreamLegacy && ctx->stream != cudaStreamPerThread;
    if (it == b->graphs.end() && b->graph_warm.count(key) && capturable && !b->graph_bad.count(key)) {
      cudaGraph_t g = nullptr;
      const int64_t l0 = ctx->launches;
      if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        const int rc = issue_iteration_head(b, iter, initialize, mode, lam_dev_in, have_residuals);
        const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &g);
        vinsat_batch::IterGraph ig;
        ig.n_launches = ctx->launches - l0;       // launches recorded into the graph, counted again at every replay
        ctx->launches = l0;
        cudaError_t ie = cudaErrorUnknown;
        if (rc == VINSAT_OK && ce == cudaSuccess && g) ie = cudaGraphInstantiate(&ig.exec, g, 0);
        if (g) cudaGraphDestroy(g);
        if (ie == cudaSuccess) {
          it = b->graphs.emplace(key, ig).first;
        } else {
          cudaGetLastError();                 // capture / instantiate failed: clear it and never try this key again
          b->graph_bad.insert(key);
        }
      } else {
        cudaGetLastError();
        b->graph_bad.insert(key);
      }
    }
    if (it != b->graphs.end()) {
      VS_CUDA(ctx, cudaGraphLaunch(it->second.exec, ctx->stream));
      ctx->launches += it->second.n_launches;
      done = true;
    } else {
      b->graph_warm.insert(key);
    }
  }
  if (!done) VS_TRY(issue_iteration_head(b, iter, initialize, mode, lam_dev_in, have_residuals));
  for (int trial = 0; trial < 16; trial++) {
    VS_TRY(wait_flag(b));
    if (b->h_flags[0] == 0) break;
    if (trial == 15) break;
    b->h_flags[0] = -1;
    VS_TRY(issue_trial(b, initialize, mode, Sigma, quat_coeff, vel_coeff));
  }
  std::swap(b->st, b->st_new);     // every problem's last trial is returned, accepted or not (:60,98)
  b->r_valid = true;
  b->have_iter = true;
  b->last_initialize = initialize;
  return VINSAT_OK;
}

int vinsat_batch_ba_iterate(vinsat_batch* b, int iter, int initialize, int mode, double* lamda_io,
                            int32_t* ntrials_out) {
  if (!b || !lamda_io) return set_error(b ? b->ctx : nullptr, VINSAT_EINVAL, "vinsat_batch_ba_iterate: NULL argument");
  vinsat_ctx* ctx = b->ctx;
  VS_CHECK_ARG(ctx, mode == VINSAT_MODE_STEP1S || mode == VINSAT_MODE_SKIP100);
  // The LM loop multiplies lamda by 10 until lamda > 1e4 (BA_filtering.py:74-77); the device loop is bounded at 16
  // trials, which covers every lamda >= 1e-10 (the reference itself never hands back less than 1e-4, :79).
  for (int64_t p = 0; p < b->P; p++)
    if (!(lamda_io[p] >= 1e-10 && lamda_io[p] <= 1e300))
      return set_error(ctx, VINSAT_EINVAL, "lamda_io[%lld] = %g outside [1e-10, inf)", (long long)p, lamda_io[p]);
  VS_CUDA(ctx, cudaSetDevice(ctx->device));
  VS_CUDA(ctx, cudaMemcpyAsync(b->lam_next, lamda_io, b->P * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  VS_TRY(ba_iterate_device(b, iter, initialize, mode, b->lam_next));
  VS_CUDA(ctx, cudaMemcpyAsync(lamda_io, b->lam_next, b->P * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (ntrials_out)
    VS_CUDA(ctx, cudaMemcpyAsync(ntrials_out, b->ntrials, b->P * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return VINSAT_OK;
}

int vinsat_batch_set_prior(vinsat_batch* b, int mem, const double* states_prior, const double* hessian_state,
                           const double* hessian_rot) {
  if (!b || !states_prior || !hessian_state || !hessian_rot)
    return set_error(b ? b->ctx : nullptr, VINSAT_EINVAL, "vinsat_batch_set_prior: NULL argument");
  vinsat_ctx* ctx = b->ctx;
  VS_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t T = b->T;
  if (!b->pr_st) {
    VS_CUDA(ctx, cudaMalloc((void**)&b->pr_st, T * 10 * sizeof(double)));
    VS_CUDA(ctx, cudaMalloc((void**)&b->pr_Hs, T * 36 * sizeof(double)));
    VS_CUDA(ctx, cudaMalloc((void**)&b->pr_Hr, T * 9 * sizeof(double)));
    VS_CUDA(ctx, cudaMalloc((void**)&b->e_pr_init, T * sizeof(double)));
    VS_CUDA(ctx, cudaMalloc((void**)&b->e_pr, T * sizeof(double)));
  }
  const cudaMemcpyKind kind = mem == VINSAT_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  VS_CUDA(ctx, cudaMemcpyAsync(b->pr_st, states_prior, T * 10 * sizeof(double), kind, ctx->stream));
  VS_CUDA(ctx, cudaMemcpyAsync(b->pr_Hs, hessian_state, T * 36 * sizeof(double), kind, ctx->stream));
  VS_CUDA(ctx, cudaMemcpyAsync(b->pr_Hr, hessian_rot, T * 9 * sizeof(double), kind, ctx->stream));
  VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return VINSAT_OK;
}

int vinsat_batch_ba_reg_iterate(vinsat_batch* b, int iter, int mode, double* lamda_io, int32_t* ntrials_out) {
  if (!b || !lamda_io) return set_error(b ? b->ctx : nullptr, VINSAT_EINVAL, "vinsat_batch_ba_reg_iterate: NULL argument");
  vinsat_ctx* ctx = b->ctx;
  if (!b->pr_st) return set_error(ctx, VINSAT_EINVAL, "vinsat_batch_set_prior has not been called");
  if (b->window) return set_error(ctx, VINSAT_EINVAL, "BA_reg is not available on frame-window shards");
  b->reg_iter = true;
  const int rc = vinsat_batch_ba_iterate(b, iter, 0, mode, lamda_io, ntrials_out);
  b->reg_iter = false;
  // the trial that produced r_next used the same projection, so the residual reuse of the next call stays valid
  return rc;
}

int vinsat_batch_od_solve(vinsat_batch* b, int num_iters, int n_init, double lamda_init, int mode) {
  if (!b) return set_error(nullptr, VINSAT_EINVAL, "vinsat_batch_od_solve: NULL batch");
  vinsat_ctx* ctx = b->ctx;
  VS_CHECK_ARG(ctx, mode == VINSAT_MODE_STEP1S || mode == VINSAT_MODE_SKIP100);
  VS_CHECK_ARG(ctx, num_iters >= 0);
  VS_CHECK_ARG(ctx, lamda_init >= 1e-10 && lamda_init <= 1e300);     // see vinsat_batch_ba_iterate
  VS_CUDA(ctx, cudaSetDevice(ctx->device));
  double* h = (double*)ctx_pinned(ctx, b->P * sizeof(double));
  if (!h) return set_error(ctx, VINSAT_ENOMEM, "pinned scratch failed");
  for (int64_t p = 0; p < b->P; p++) h[p] = lamda_init;
  VS_CUDA(ctx, cudaMemcpyAsync(b->lam_next, h, b->P * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  // Speculative pipeline (Monte-Carlo path): the head of iteration it+1 is enqueued BEFORE the host learns whether
  // iteration it's LM loop ended with its first trial.  Every kernel of a head carries the device gate word: the
  // first trial publishes the number of still-active problems into it, and while it is non-zero the kernels behind
  // it return at once.  In the common case the GPU never waits for the host; otherwise the host finishes the loop
  // with ungated trials, clears the gate and enqueues the head again.
  // OPT-IN (VINSAT_SPECULATE=1).  Measured on B200: it removes the 20 host round trips per solve, which costs 0.3 ms
  // per solve on a host whose round trips are already short (3 extra tiny launches per iteration) and did not change
  // the 8-GPU result (the ranks that run slow there stay slow), so the default is the plain loop.
  const bool no_spec = getenv("VINSAT_SPECULATE") == nullptr;      // read per call (tests switch it)
  const bool smem_select = 2 * b->max_obs_per_problem > 0 && 2 * b->max_obs_per_problem <= 24000 &&
                           getenv("VINSAT_SELECT_GLOBAL") == nullptr;
  if (no_spec || ctx->timing || b->window || b->partitioned || !smem_select || num_iters < 2) {
    for (int it = 0; it < num_iters; it++) VS_TRY(ba_iterate_device(b, it, it < n_init ? 1 : 0, mode, b->lam_next));
    return VINSAT_OK;
  }
  const double quat_coeff = 100.0, vel_coeff = 100.0;
  static const bool no_fuse = getenv("VINSAT_NO_FUSED_SYSTEM") != nullptr || getenv("VINSAT_ONE_SIDED_SWEEP") != nullptr;
  auto sigma_of = [](int it) { const double it1 = (double)it + 1.0; return std::min(10000.0 * it1 * it1, 1000000.0); };
  bool swapped_r = false;
  auto begin_iter = [&](int it) {                    // host-side state of ba_iterate_device's prologue
    const int init = it < n_init ? 1 : 0;
    swapped_r = b->r_valid;
    if (swapped_r) std::swap(b->r, b->r_next);
    b->fused_system = !init && !no_fuse;
    b->srec_valid = !init && !b->fused_system;
    b->last_sigma = b->cur_sigma = sigma_of(it);
    b->cur_vc = vel_coeff;
    return swapped_r;
  };
  auto end_iter = [&](int it) {                      // ... and of its epilogue
    std::swap(b->st, b->st_new);
    b->r_valid = true;
    b->have_iter = true;
    b->last_initialize = it < n_init ? 1 : 0;
  };
  b->gate = b->flags + 3;
  VS_CUDA(ctx, cudaMemsetAsync(b->gate, 0, sizeof(int32_t), ctx->stream));
  b->gate_arg = b->gate;
  volatile int32_t* ring = b->h_flags + 2;
  auto fail = [&](int rc) { b->gate_arg = nullptr; return rc; };
  {
    const bool had = begin_iter(0);
    ring[0] = -1;
    const int rc = issue_iteration_head(b, 0, 0 < n_init ? 1 : 0, mode, b->lam_next, had, b->h_flags + 2);
    if (rc != VINSAT_OK) return fail(rc);
  }
  for (int it = 0; it < num_iters; it++) {
    const int init = it < n_init ? 1 : 0;
    const bool spec = it + 1 < num_iters;
    const bool r_valid_before = b->r_valid;
    bool had_next = false;
    if (spec) {
      end_iter(it);
      had_next = begin_iter(it + 1);
      ring[(it + 1) & 1] = -1;
      const int rc = issue_iteration_head(b, it + 1, it + 1 < n_init ? 1 : 0, mode, b->lam_next, had_next,
                                          b->h_flags + 2 + ((it + 1) & 1));
      if (rc != VINSAT_OK) return fail(rc);
    }
    {                                                 // wait for iteration it's first-trial counter
      int64_t spins = 0;
      while (ring[it & 1] == -1 && ++spins < 400000000ll) {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
      }
      if (ring[it & 1] == -1) {
        const cudaError_t ce = cudaStreamSynchronize(ctx->stream);
        if (ce != cudaSuccess) { b->gate_arg = nullptr; VS_CUDA(ctx, ce); }
      }
    }
    if (ring[it & 1] == 0) {
      if (!spec) end_iter(it);
      continue;
    }
    // some problem's LM loop goes on: the head enqueued behind it is skipped on the device
    {
      const cudaError_t ce = cudaStreamSynchronize(ctx->stream);
      if (ce != cudaSuccess) { b->gate_arg = nullptr; VS_CUDA(ctx, ce); }
    }
    if (spec) {                                       // take back the host-side effects of the skipped head
      if (had_next) std::swap(b->r, b->r_next);
      std::swap(b->st, b->st_new);
      b->r_valid = r_valid_before;
      b->fused_system = !init && !no_fuse;
      b->srec_valid = !init && !b->fused_system;
      b->last_sigma = b->cur_sigma = sigma_of(it);
    }
    b->gate_arg = nullptr;                            // the remaining trials of iteration it run ungated
    for (int trial = 1; trial < 16; trial++) {
      b->h_flags[0] = -1;
      int rc = issue_trial(b, init, mode, sigma_of(it), quat_coeff, vel_coeff);
      if (rc == VINSAT_OK) rc = wait_flag(b);
      if (rc != VINSAT_OK) return rc;
      if (b->h_flags[0] == 0) break;
    }
    VS_CUDA(ctx, cudaMemsetAsync(b->gate, 0, sizeof(int32_t), ctx->stream));
    b->gate_arg = b->gate;
    end_iter(it);
    if (spec) {
      had_next = begin_iter(it + 1);
      ring[(it + 1) & 1] = -1;
      const int rc = issue_iteration_head(b, it + 1, it + 1 < n_init ? 1 : 0, mode, b->lam_next, had_next,
                                          b->h_flags + 2 + ((it + 1) & 1));
      if (rc != VINSAT_OK) return fail(rc);
    }
    // invariant restored: at the top of pass it+1 the head of iteration it+1 is in flight
  }
  b->gate_arg = nullptr;
  return VINSAT_OK;
}

static int ensure_srec(vinsat_batch* b) {
  if (b->srec_valid) return VINSAT_OK;
  VS_TRY(launch_system_build(b, b->last_initialize, b->last_sigma, 100.0));
  b->srec_valid = true;
  return VINSAT_OK;
}

int vinsat_batch_last_hessian(vinsat_batch* b, double* out) {
  if (!b || !out) return set_error(b ? b->ctx : nullptr, VINSAT_EINVAL, "vinsat_batch_last_hessian: NULL argument");
  vinsat_ctx* ctx = b->ctx;
  if (!b->have_iter) return set_error(ctx, VINSAT_EINVAL, "no BA iteration has run on this batch");
  VS_CUDA(ctx, cudaSetDevice(ctx->device));
  VS_TRY(ensure_srec(b));
  double* d_out = (double*)ctx_scratch(ctx, (size_t)b->P * 81 * sizeof(double));
  if (!d_out) return set_error(ctx, VINSAT_ENOMEM, "scratch allocation failed");
  VS_TRY(launch_gather_last_hessian(b, d_out));          // JTwJ includes eye*lamda (:54,97)
  VS_CUDA(ctx, cudaMemcpyAsync(out, d_out, (size_t)b->P * 81 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return VINSAT_OK;
}

int vinsat_batch_debug_fetch(vinsat_batch* b, double* r_obs, double* weights, double* c_obs, double* D, double* U,
                             double* rhs, double* dpose) {
  if (!b) return set_error(nullptr, VINSAT_EINVAL, "vinsat_batch_debug_fetch: NULL batch");
  vinsat_ctx* ctx = b->ctx;
  if (!b->have_iter) return set_error(ctx, VINSAT_EINVAL, "no BA iteration has run on this batch");
  VS_CUDA(ctx, cudaSetDevice(ctx->device));
  VS_TRY(ensure_srec(b));
  VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const int64_t P = b->P, T = b->T, M = b->M;
  if (r_obs && M) {
    std::vector<double> soa(2 * M);
    VS_CUDA(ctx, cudaMemcpy(soa.data(), b->r, 2 * M * sizeof(double), cudaMemcpyDeviceToHost));
    for (int64_t k = 0; k < M; k++) { r_obs[2 * k] = soa[k]; r_obs[2 * k + 1] = soa[M + k]; }
  }
  if (c_obs) VS_CUDA(ctx, cudaMemcpy(c_obs, b->c_obs, P * sizeof(double), cudaMemcpyDeviceToHost));
  if (weights && M) {
    std::vector<unsigned long long> wm(P);
    VS_CUDA(ctx, cudaMemcpy(weights, b->wu, M * sizeof(double), cudaMemcpyDeviceToHost));
    VS_CUDA(ctx, cudaMemcpy(wm.data(), b->wmax, P * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    for (int64_t p = 0; p < P; p++) {
      double w;
      memcpy(&w, &wm[p], 8);
      for (int64_t k = b->obs_off[p]; k < b->obs_off[p + 1]; k++) weights[k] = weights[k] / w;
    }
  }
  if (D || U || rhs) {
    std::vector<double> rec((size_t)T * VS_SREC);
    VS_CUDA(ctx, cudaMemcpy(rec.data(), b->srec, rec.size() * sizeof(double), cudaMemcpyDeviceToHost));
    for (int64_t f = 0; f < T; f++) {
      if (D) memcpy(D + f * 81, &rec[f * VS_SREC], 81 * sizeof(double));
      if (U) memcpy(U + f * 81, &rec[f * VS_SREC + 81], 81 * sizeof(double));
      if (rhs) memcpy(rhs + f * 9, &rec[f * VS_SREC + 162], 9 * sizeof(double));
    }
  }
  if (dpose) VS_CUDA(ctx, cudaMemcpy(dpose, b->delta, T * 9 * sizeof(double), cudaMemcpyDeviceToHost));
  return VINSAT_OK;
}

int vinsat_batch_eval_resjac(vinsat_batch* b) {
  if (!b) return set_error(nullptr, VINSAT_EINVAL, "vinsat_batch_eval_resjac: NULL batch");
  vinsat_ctx* ctx = b->ctx;
  VS_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!b->J) {
    const bool arena = ctx->arena_on;      // J outlives any arena: always a plain allocation
    ctx->arena_on = false;
    const int rc = dev_alloc(ctx, &b->J, b->M * 12);
    ctx->arena_on = arena;
    if (rc != VINSAT_OK) return rc;
  }
  return launch_resjac(b);
}

int vinsat_batch_fetch_resjac(vinsat_batch* b, double* r_out, double* J_out) {
  if (!b) return set_error(nullptr, VINSAT_EINVAL, "vinsat_batch_fetch_resjac: NULL batch");
  vinsat_ctx* ctx = b->ctx;
  VS_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t M = b->M;
  if (M == 0) return VINSAT_OK;
  if (!b->J) return set_error(ctx, VINSAT_EINVAL, "vinsat_batch_eval_resjac has not run");
  double* tmp = (double*)ctx_scratch(ctx, (size_t)M * 12 * sizeof(double));
  if (!tmp) return set_error(ctx, VINSAT_ENOMEM, "scratch failed");
  if (r_out) {
    VS_TRY(launch_soa_to_aos(ctx, b->r, tmp, M, 2));
    VS_CUDA(ctx, cudaMemcpyAsync(r_out, tmp, M * 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  if (J_out) {
    VS_TRY(launch_soa_to_aos(ctx, b->J, tmp, M, 12));
    VS_CUDA(ctx, cudaMemcpyAsync(J_out, tmp, M * 12 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return VINSAT_OK;
}


// ---------------------------------------------------------------------------------------------------------------
// (f)1: streaming_version's outer loop on the device (od_pipe.py:987-1060).
// The window schedule (integers from identify_next_batch_new, :898-905) comes from the host; everything between
// "window w starts" and "window w is solved" runs here without handing states back: the new frames are seeded by
// propagate_dynamics_init (:1011, BA_utils.py:114-129) from the last solved state, the window's num_iters BA calls run
// back to back (:1035-1040; the host only polls the pinned LM-trial counter), and the solved states stay in device
// memory for the next window.  Windows always start at frame 0 and grow (:994-1000,1015-1023), so every window's batch
// is carved out of ONE arena sized for the last window.
// ---------------------------------------------------------------------------------------------------------------
int vinsat_stream_solve(vinsat_ctx* ctx, const vinsat_stream_desc* d, int num_iters, int n_init_first, double lamda_init,
                        int mode, double* states_out, double* seed_states_out, double* window_last_state_out,
                        double* last_hessian_out) {
  VS_CHECK_ARG(ctx, ctx != nullptr && d != nullptr);
  VS_CHECK_ARG(ctx, mode == VINSAT_MODE_STEP1S || mode == VINSAT_MODE_SKIP100);
  VS_CHECK_ARG(ctx, d->n_frames >= 1 && d->n_obs >= 0 && d->n_windows >= 1 && d->t_final && d->i_final);
  VS_CHECK_ARG(ctx, d->states && d->velocities && d->intrinsics && d->cum_rot && d->time_idx && states_out);
  VS_CHECK_ARG(ctx, d->n_obs == 0 || (d->landmarks_xyz && d->landmarks_uv && d->confidences && d->ii));
  VS_CHECK_ARG(ctx, num_iters >= 0 && lamda_init >= 1e-10);
  VS_CHECK_ARG(ctx, !ctx->arena_on);
  const int64_t T_all = d->n_frames, W = d->n_windows;
  for (int64_t w = 0; w < W; w++) {
    VS_CHECK_ARG(ctx, d->t_final[w] >= 1 && d->t_final[w] <= T_all && d->i_final[w] >= 0 && d->i_final[w] <= d->n_obs);
    VS_CHECK_ARG(ctx, w == 0 || (d->t_final[w] > d->t_final[w - 1] && d->i_final[w] >= d->i_final[w - 1]));
  }
  const int64_t T_last = d->t_final[W - 1], M_last = d->i_final[W - 1];
  const bool tail = T_last < T_all;                       // od_pipe.py:1046: frames after the last window are propagated only
  const int64_t span = d->time_idx[T_all - 1] - d->time_idx[0];
  VS_CHECK_ARG(ctx, span >= 0 && span < (1ll << 31));
  if (W > 1 || tail) VS_CHECK_ARG(ctx, d->omega && d->n_omega >= d->time_idx[(tail ? T_all : T_last) - 1]);
  VS_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  // persistent device state across windows
  DevBuf<double> st_carry, vel_carry, seed, chain, omega, win_last;
  DevBuf<int64_t> tidx;
  VS_CUDA(ctx, st_carry.alloc(T_all * 10));
  VS_CUDA(ctx, vel_carry.alloc(T_all * 3));
  VS_CUDA(ctx, seed.alloc(T_all * 10));
  VS_CUDA(ctx, chain.alloc((span + 2) * 10));
  VS_CUDA(ctx, omega.alloc(std::max<int64_t>(d->n_omega, 1) * 3));
  VS_CUDA(ctx, win_last.alloc(W * 10));
  VS_CUDA(ctx, tidx.alloc(T_all));
  VS_CUDA(ctx, cudaMemcpyAsync(vel_carry.p, d->velocities, T_all * 3 * sizeof(double), cudaMemcpyHostToDevice, s));
  VS_CUDA(ctx, cudaMemcpyAsync(seed.p, d->states, T_all * 10 * sizeof(double), cudaMemcpyHostToDevice, s));
  VS_CUDA(ctx, cudaMemcpyAsync(tidx.p, d->time_idx, T_all * sizeof(int64_t), cudaMemcpyHostToDevice, s));
  if (d->n_omega > 0 && d->omega)
    VS_CUDA(ctx, cudaMemcpyAsync(omega.p, d->omega, d->n_omega * 3 * sizeof(double), cudaMemcpyHostToDevice, s));
  // one arena for every window's batch, sized for the last (largest) window
  {
    const int64_t nseg = (int64_t)sqrt((double)T_last) + 2;
    const size_t need = (size_t)T_last * 4600 + (size_t)M_last * 104 + (size_t)nseg * 6400 + (1u << 20);
    if (ctx->arena_bytes < need) {
      if (ctx->arena) { cudaStreamSynchronize(s); cudaFree(ctx->arena); ctx->arena = nullptr; ctx->arena_bytes = 0; }
      VS_CUDA(ctx, cudaMalloc((void**)&ctx->arena, need));
      ctx->arena_bytes = need;
    }
  }
  const int64_t frame_off[2] = {0, 0}, obs_off[2] = {0, 0};
  int rc = VINSAT_OK;
  int64_t T_prev = 0;
  for (int64_t w = 0; w < W && rc == VINSAT_OK; w++) {
    const int64_t Tw = d->t_final[w], Mw = d->i_final[w];
    int64_t fo[2] = {frame_off[0], Tw}, oo[2] = {obs_off[0], Mw};
    vinsat_problem_desc pd;
    pd.n_problems = 1;
    pd.frame_off = fo; pd.obs_off = oo;
    pd.states = d->states; pd.intrinsics = d->intrinsics; pd.cum_rot = d->cum_rot; pd.time_idx = d->time_idx;
    pd.landmarks_xyz = d->landmarks_xyz; pd.landmarks_uv = d->landmarks_uv; pd.confidences = d->confidences; pd.ii = d->ii;
    vinsat_batch* b = nullptr;
    ctx->arena_off = 0;
    ctx->arena_on = true;
    rc = create_impl(ctx, &pd, 0, 0, 0, &b);
    ctx->arena_on = false;
    if (rc != VINSAT_OK) break;
    auto step = [&]() -> int {
      if (w > 0) {
        // seed the new frames: chain from the last solved state with the carried `velocities` row (the reference passes
        // velocities_t[:, -1], not the state's own velocity, :1011), tdiff + duration steps of 1 s
        const int64_t t0 = d->time_idx[T_prev - 1];
        const int64_t n_steps = d->time_idx[Tw - 1] - t0;
        VS_TRY(launch_chain(ctx, n_steps, 1.0, st_carry.p + (T_prev - 1) * 10, vel_carry.p + (T_prev - 1) * 3,
                            omega.p + t0 * 3, chain.p));
        VS_CUDA(ctx, cudaMemcpyAsync(b->st, st_carry.p, T_prev * 10 * sizeof(double), cudaMemcpyDeviceToDevice, s));
        VS_TRY(launch_stream_gather(ctx, Tw - T_prev, chain.p, tidx.p, T_prev, t0, b->st, vel_carry.p, seed.p));
      }
      VS_TRY(vinsat_batch_od_solve(b, num_iters, w == 0 ? n_init_first : 0, lamda_init, mode));
      VS_CUDA(ctx, cudaMemcpyAsync(st_carry.p, b->st, Tw * 10 * sizeof(double), cudaMemcpyDeviceToDevice, s));
      VS_CUDA(ctx, cudaMemcpyAsync(win_last.p + w * 10, b->st + (Tw - 1) * 10, 10 * sizeof(double), cudaMemcpyDeviceToDevice, s));
      if (w == W - 1 && last_hessian_out && num_iters > 0) VS_TRY(vinsat_batch_last_hessian(b, last_hessian_out));
      return VINSAT_OK;
    };
    rc = step();
    cudaStreamSynchronize(s);          // the arena is recycled by the next window
    free_all(b);
    delete b;
    T_prev = Tw;
  }
  if (rc != VINSAT_OK) return rc;
  if (tail) {
    const int64_t t0 = d->time_idx[T_last - 1];
    const int64_t n_steps = d->time_idx[T_all - 1] - t0;
    VS_TRY(launch_chain(ctx, n_steps, 1.0, st_carry.p + (T_last - 1) * 10, vel_carry.p + (T_last - 1) * 3, omega.p + t0 * 3,
                        chain.p));
    VS_TRY(launch_stream_gather(ctx, T_all - T_last, chain.p, tidx.p, T_last, t0, nullptr, vel_carry.p, seed.p));
  }
  VS_CUDA(ctx, cudaMemcpyAsync(states_out, st_carry.p, T_last * 10 * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (seed_states_out)
    VS_CUDA(ctx, cudaMemcpyAsync(seed_states_out, seed.p, T_all * 10 * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (window_last_state_out)
    VS_CUDA(ctx, cudaMemcpyAsync(window_last_state_out, win_last.p, W * 10 * sizeof(double), cudaMemcpyDeviceToHost, s));
  VS_CUDA(ctx, cudaStreamSynchronize(s));
  return VINSAT_OK;
}

}  // extern "C"
