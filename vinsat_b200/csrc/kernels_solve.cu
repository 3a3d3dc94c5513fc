// Per-problem kernels (a5 scale, a6 system build, a7 LM step of SURVEY.md section 8): exact lower-median by radix
// select, block-tridiagonal normal equations, batched block LU solve with fused retraction, accept test.
#include <cuda_pipeline.h>

#include "common.cuh"
#include "launch.h"

namespace vs {

// ---------------------------------------------------------------------------------------------------------
// exact lower median of |r| per problem (torch.median semantics, BA_filtering.py:23): MSD radix select on
// the bit pattern of the non-negative doubles, 11-bit digits, 6 passes.
// ---------------------------------------------------------------------------------------------------------
constexpr int kSelBins = 2048;

__global__ void k_select_init(int P, const int64_t* __restrict__ obs_off, unsigned long long* __restrict__ prefix,
                              unsigned long long* __restrict__ rank) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const long long n = 2 * (obs_off[p + 1] - obs_off[p]);
  prefix[p] = 0ull;
  rank[p] = n > 0 ? (unsigned long long)((n - 1) / 2) : 0ull;
}

__global__ void __launch_bounds__(256) k_select_hist(int64_t M, const int64_t* __restrict__ obs_off,
                                                     const double* __restrict__ r,
                                                     const unsigned long long* __restrict__ prefix, int shift,
                                                     int nbits, unsigned int* __restrict__ hist) {
  __shared__ unsigned int sh[kSelBins];
  const int p = blockIdx.y;
  const int64_t k0 = obs_off[p], k1 = obs_off[p + 1];
  const int64_t n = 2 * (k1 - k0);
  const int64_t per = (n + gridDim.x - 1) / gridDim.x;
  const int64_t i0 = blockIdx.x * per, i1 = min(n, i0 + per);
  if (i0 >= i1) return;
  for (int i = threadIdx.x; i < kSelBins; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const unsigned long long pre = prefix[p];
  const int hs = shift + nbits;
  const unsigned int dmask = (1u << nbits) - 1u;
  const int64_t Mp = k1 - k0;
  for (int64_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    // element i of the problem's 2*Mp values: component i / Mp, observation k0 + i % Mp
    const int64_t comp = i >= Mp ? 1 : 0;
    const double v = fabs(r[comp * M + k0 + (i - comp * Mp)]);
    const unsigned long long key = (unsigned long long)__double_as_longlong(v);
    const bool match = hs >= 64 ? true : ((key >> hs) == (pre >> hs));
    if (match) atomicAdd(&sh[(unsigned int)(key >> shift) & dmask], 1u);
  }
  __syncthreads();
  unsigned int* g = hist + (int64_t)p * kSelBins;
  for (int i = threadIdx.x; i < kSelBins; i += blockDim.x)
    if (sh[i]) atomicAdd(&g[i], sh[i]);
}

__global__ void __launch_bounds__(256) k_select_pick(unsigned long long* __restrict__ prefix,
                                                     unsigned long long* __restrict__ rank, int shift,
                                                     unsigned int* __restrict__ hist, int last,
                                                     double* __restrict__ c_obs, const int64_t* __restrict__ obs_off) {
  __shared__ unsigned int part[256];
  const int p = blockIdx.x;
  unsigned int* g = hist + (int64_t)p * kSelBins;
  unsigned int loc[8];
  unsigned int s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) { loc[i] = g[threadIdx.x * 8 + i]; s += loc[i]; g[threadIdx.x * 8 + i] = 0; }
  part[threadIdx.x] = s;
  __syncthreads();
  // exclusive prefix of the 256 partial sums (small; done redundantly by every thread's own walk)
  unsigned long long before = 0;
  for (int i = 0; i < threadIdx.x; i++) before += part[i];
  const unsigned long long rk = rank[p];
  __syncthreads();
  if (rk >= before && rk < before + s) {
    unsigned long long cum = before;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (rk < cum + loc[i]) {
        const unsigned long long np = prefix[p] | ((unsigned long long)(threadIdx.x * 8 + i) << shift);
        prefix[p] = np;
        rank[p] = rk - cum;
        if (last) c_obs[p] = __longlong_as_double((long long)np);
        break;
      }
      cum += loc[i];
    }
  }
  if (last && threadIdx.x == 0 && obs_off[p + 1] == obs_off[p]) c_obs[p] = __longlong_as_double(0x7ff8000000000000LL);
}

// ---------------------------------------------------------------------------------------------------------
// Small problems (2 M_p keys fit in shared memory, e.g. 20 000 keys = 160 KB): ONE CTA per problem keeps the keys
// resident and runs all six radix-select passes on chip: one pass over HBM instead of six plus 12 launches.
// ---------------------------------------------------------------------------------------------------------
constexpr int kSelThreads = 1024;
constexpr int kSelCand = 2048;      // candidate list: keys that share the digits found so far

__global__ void __launch_bounds__(kSelThreads) k_select_smem(int64_t M, const int64_t* __restrict__ obs_off,
                                                             const double* __restrict__ r, double* __restrict__ c_obs,
                                                             const int32_t* __restrict__ gate) {
  if (gate && *gate) return;      // speculative launch behind an LM loop that is not finished (batch.cu)
  extern __shared__ __align__(16) unsigned long long sel_keys[];      // [n] keys, then [kSelCand] candidates
  __shared__ unsigned int hist[kSelBins];
  __shared__ unsigned int wsum[32];
  __shared__ unsigned long long s_prefix, s_rank, s_diff;
  __shared__ unsigned int s_cnt, s_bincount;
  const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t k0 = obs_off[p], Mp = obs_off[p + 1] - k0;
  const int n = (int)(2 * Mp);
  if (n == 0) {
    if (tid == 0) c_obs[p] = __longlong_as_double(0x7ff8000000000000LL);
    return;
  }
  unsigned long long* cand = sel_keys + n;
  // keys = bit patterns of |r|, loaded straight into registers (eight independent loads in flight per thread), made
  // non-negative and parked in shared memory.  Bits that are identical in every key carry no information: the same
  // pass ORs (key ^ key0) so that the 11-bit digits can start at the highest differing bit.
  if (tid == 0) s_diff = 0ull;
  {
    const unsigned long long* r0 = reinterpret_cast<const unsigned long long*>(r) + k0;          // u residuals
    const unsigned long long* r1 = reinterpret_cast<const unsigned long long*>(r) + M + k0;      // v residuals
    const unsigned long long key0 = r0[0] & 0x7fffffffffffffffull;
    unsigned long long d = 0ull;
    constexpr int kLd = 8;
    for (int base = tid; base < n; base += kSelThreads * kLd) {
      unsigned long long k[kLd];
#pragma unroll
      for (int j = 0; j < kLd; j++) {
        const int i = base + j * kSelThreads;
        k[j] = i < n ? (i < Mp ? r0[i] : r1[i - Mp]) : key0;
      }
#pragma unroll
      for (int j = 0; j < kLd; j++) {
        const int i = base + j * kSelThreads;
        const unsigned long long key = k[j] & 0x7fffffffffffffffull;
        if (i < n) sel_keys[i] = key;
        d |= key ^ key0;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d |= __shfl_xor_sync(0xffffffffu, d, o);
    __syncthreads();                       // s_diff = 0 is visible
    if (lane == 0 && d) atomicOr(&s_diff, d);
  }
  __syncthreads();
  const unsigned long long diff = s_diff;
  int hi = diff ? 64 - __clzll((long long)diff) : 0;        // digits cover bits [0, hi)
  if (tid == 0) {
    s_prefix = hi >= 64 ? 0ull : ((sel_keys[0] >> hi) << hi);
    s_rank = (unsigned long long)((n - 1) / 2);
  }
  // After a pass the keys that matter are those in the selected bin.  If they fit the candidate list they are
  // compacted into it and the remaining passes scan only them (typically a few dozen keys instead of 2 M_p).
  const unsigned long long* list = sel_keys;
  int nlist = n;
  bool all_match = true;          // every key of `list` shares the prefix found so far
#pragma unroll 1
  while (hi > 0) {
    const int nb = min(11, hi), shift = hi - nb, hs = hi;
    for (int i = tid; i < kSelBins; i += kSelThreads) hist[i] = 0;
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    const unsigned long long pre = s_prefix;
    const unsigned int dmask = (1u << nb) - 1u;
    for (int i = tid; i < nlist; i += kSelThreads) {
      const unsigned long long key = list[i];
      const bool match = (all_match || hs >= 64) ? true : ((key >> hs) == (pre >> hs));
      if (match) atomicAdd(&hist[(unsigned int)(key >> shift) & dmask], 1u);
    }
    __syncthreads();
    // each thread owns bins 2*tid, 2*tid+1; block-wide exclusive scan of the pair sums
    const unsigned int h0 = hist[2 * tid], h1 = hist[2 * tid + 1];
    unsigned int sum2 = h0 + h1, incl = sum2;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      unsigned int w = wsum[lane], wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const unsigned int v = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += v; }
      wsum[lane] = wi - w;          // exclusive prefix of the warp sums
    }
    __syncthreads();
    const unsigned long long before = (unsigned long long)wsum[warp] + (incl - sum2);
    const unsigned long long rk = s_rank;
    __syncthreads();
    if (rk >= before && rk < before + sum2) {
      const int bin = (rk < before + h0) ? 2 * tid : 2 * tid + 1;
      s_prefix = pre | ((unsigned long long)bin << shift);
      s_rank = rk - (bin == 2 * tid ? before : before + h0);
      s_bincount = (bin == 2 * tid) ? h0 : h1;
    }
    __syncthreads();
    hi = shift;
    const unsigned int bc = s_bincount;
    if (hi > 0 && list == sel_keys && bc <= (unsigned int)kSelCand) {
      // compact the selected bin (order is irrelevant: the k-th smallest of a set does not depend on it)
      const unsigned long long npre = s_prefix;
      for (int i = tid; i < nlist; i += kSelThreads) {
        const unsigned long long key = list[i];
        if ((key >> shift) == (npre >> shift)) cand[atomicAdd(&s_cnt, 1u)] = key;
      }
      __syncthreads();
      // The selected bin holds a few dozen keys: finish by COUNTING instead of four more digit passes (each of which
      // costs six block-wide barriers): candidate t is the answer iff exactly `rank` candidates order before it
      // (ties broken by position, so equal keys get distinct ranks).
      const unsigned int rk2 = (unsigned int)s_rank;
      for (unsigned int t = tid; t < bc; t += kSelThreads) {
        const unsigned long long mine = cand[t];
        unsigned int below = 0;
        for (unsigned int j = 0; j < bc; j++) {
          const unsigned long long o = cand[j];
          below += (o < mine || (o == mine && j < t)) ? 1u : 0u;
        }
        if (below == rk2) s_prefix = mine;
      }
      __syncthreads();
      break;
    } else {
      all_match = false;
    }
  }
  if (tid == 0) c_obs[p] = __longlong_as_double((long long)s_prefix);
}

static const int kSelShift[6] = {53, 42, 31, 20, 9, 0};
static const int kSelBits[6] = {11, 11, 11, 11, 11, 9};

__global__ void k_select_set_rank(unsigned long long* __restrict__ prefix, unsigned long long* __restrict__ rank,
                                  unsigned long long r) {
  prefix[0] = 0ull;
  rank[0] = r;
}

int launch_select_begin(vinsat_batch* b, int64_t global_values) {
  vinsat_ctx* ctx = b->ctx;
  if (global_values < 0) {
    VS_LAUNCH(ctx, F_SELECT, k_select_init, ceil_div(b->P, 128), 128, 0, (int)b->P, b->d_obs_off, b->sel_prefix,
              b->sel_rank);
  } else {   // frame-window sharded single problem: the rank counts every rank's values
    VS_LAUNCH(ctx, F_SELECT, k_select_set_rank, 1, 1, 0, b->sel_prefix, b->sel_rank,
              (unsigned long long)(global_values > 0 ? (global_values - 1) / 2 : 0));
  }
  return VINSAT_OK;
}

int launch_select_hist(vinsat_batch* b, int pass) {
  vinsat_ctx* ctx = b->ctx;
  int64_t chunks = ceil_div(2 * b->max_obs_per_problem, 256 * 16);
  if (chunks < 1) chunks = 1;
  const int64_t cap = std::max<int64_t>(1, (int64_t)ctx->sm_count * 8 / std::max<int64_t>(1, b->P));
  if (chunks > cap) chunks = cap;
  dim3 grid((unsigned)chunks, (unsigned)b->P);
  VS_LAUNCH(ctx, F_SELECT, k_select_hist, grid, 256, 0, b->M, b->d_obs_off, b->r, b->sel_prefix, kSelShift[pass],
            kSelBits[pass], b->sel_hist);
  return VINSAT_OK;
}

int launch_select_pick(vinsat_batch* b, int pass) {
  vinsat_ctx* ctx = b->ctx;
  VS_LAUNCH(ctx, F_SELECT, k_select_pick, (unsigned)b->P, 256, 0, b->sel_prefix, b->sel_rank, kSelShift[pass],
            b->sel_hist, pass == 5 ? 1 : 0, b->c_obs, b->d_obs_off);
  return VINSAT_OK;
}

int launch_select_median(vinsat_batch* b) {
  if (b->P == 0) return VINSAT_OK;
  static const bool no_smem = getenv("VINSAT_SELECT_GLOBAL") != nullptr;
  const int64_t keys = 2 * b->max_obs_per_problem;
  if (!no_smem && keys > 0 && keys <= 24000 && !b->window) {
    vinsat_ctx* ctx = b->ctx;
    const int smem = (int)((keys + kSelCand) * sizeof(unsigned long long));
    VS_SMEM_OPTIN(ctx, SM_SELECT, k_select_smem, smem);
    VS_LAUNCH(ctx, F_SELECT, k_select_smem, (unsigned)b->P, kSelThreads, smem, b->M, b->d_obs_off, b->r, b->c_obs,
              b->gate_arg);
    return VINSAT_OK;
  }
  int rc = launch_select_begin(b, -1);
  for (int pass = 0; pass < 6 && rc == VINSAT_OK; pass++) {
    rc = launch_select_hist(b, pass);
    if (rc == VINSAT_OK) rc = launch_select_pick(b, pass);
  }
  return rc;
}

// ---------------------------------------------------------------------------------------------------------
// block-tridiagonal normal equations (BA_filtering.py:28-48).  One warp per frame: the frame's
// records are staged in shared memory, the 171 outputs are written as one contiguous record.
// srec[f]: D 81 (row-major, WITHOUT damping) | U 81 (block (f,f+1), row-major) | b 9
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int pv_index(int a) { return a < 3 ? a : (a >= 6 ? a - 3 : -1); }
__device__ __forceinline__ int sym_index(int a, int b) {   // upper triangle of 6x6, a <= b
  return a * 6 - (a * (a - 1)) / 2 + (b - a);
}

// 3 threads per frame (row blocks pos / rot / vel), 32 frames per CTA.  The frames' observation and dynamics
// records are staged with coalesced loads into shared memory; every thread then builds three complete rows of
// D, U and b in registers and writes them as contiguous runs.  Used by the partitioned / frame-window sharded
// solves and by the diagnostics; the Monte-Carlo path builds its columns inside the sweep (kernels_chain.cu).
constexpr int kSysFrames = 32;
constexpr int kSysStride = VS_GREC + VS_DREC + 7;     // 99 doubles: odd stride => conflict-free row access
constexpr int kSysOut = VS_SREC + 1;                  // 173: odd stride of the staged output records

__global__ void __launch_bounds__(96) k_system_rows(int64_t T, const int32_t* __restrict__ gap,
                                                    const int32_t* __restrict__ fprob,
                                                    const unsigned long long* __restrict__ wmax,
                                                    const double* __restrict__ grec, const double* __restrict__ drec,
                                                    int initialize, double Sigma, double vc,
                                                    double* __restrict__ srec) {
  extern __shared__ __align__(16) double sys_smem[];
  double* sm = sys_smem;                                   // [kSysFrames][kSysStride] inputs
  double* so = sys_smem + kSysFrames * kSysStride;         // [kSysFrames][kSysOut]    outputs
  const int tid = threadIdx.x;
  const int64_t f0 = (int64_t)blockIdx.x * kSysFrames;
  const int nf = (int)min((int64_t)kSysFrames, T - f0);
  for (int i = tid; i < nf * VS_GREC; i += 96) sm[(i / VS_GREC) * kSysStride + (i % VS_GREC)] = grec[f0 * VS_GREC + i];
  if (!initialize) {
    for (int i = tid; i < nf * VS_DREC; i += 96)
      sm[(i / VS_DREC) * kSysStride + VS_GREC + (i % VS_DREC)] = drec[f0 * VS_DREC + i];
    if (tid < 6) sm[VS_GREC + VS_DREC + tid] = (f0 > 0) ? drec[(f0 - 1) * VS_DREC + 36 + tid] : 0.0;   // r of the pair before the tile
  }
  __syncthreads();
  const int lf = tid / 3, q = tid % 3;          // local frame, row block (0 pos, 1 rot, 2 vel)
  if (lf < nf) {
  const int64_t f = f0 + lf;
  const double* G = sm + lf * kSysStride;
  const double* Dr = G + VS_GREC;
  const double* rp = (lf > 0) ? (G - kSysStride + VS_GREC + 36) : (sm + VS_GREC + VS_DREC);   // r6 of pair (f-1, f)
  const bool has_next = !initialize && gap[f] > 0;
  const bool has_prev = !initialize && f > 0 && gap[f - 1] > 0;
  const unsigned long long wb = wmax[fprob[f]];
  const double invw = wb ? 1.0 / __longlong_as_double((long long)wb) : 0.0;
  const double dv2[6] = {1.0, 1.0, 1.0, vc * vc, vc * vc, vc * vc};
  const double dv1[6] = {1.0, 1.0, 1.0, vc, vc, vc};
  double* out = so + lf * kSysOut;
  double Drow[3][9], Urow[3][9], brow[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    brow[i] = 0.0;
#pragma unroll
    for (int c = 0; c < 9; c++) { Drow[i][c] = 0.0; Urow[i][c] = 0.0; }
  }
  if (q == 1) {
    // rotation rows 3..5: observation block columns 0..5, Sigma*Hq_diag / Sigma*Hq_off in the rotation columns
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const int a = 3 + i;
#pragma unroll
      for (int c = 0; c < 6; c++) Drow[i][c] = invw * G[a <= c ? sym_index(a, c) : sym_index(c, a)];
      brow[i] = invw * G[21 + a];
      if (!initialize) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
          Drow[i][3 + c] += Sigma * Dr[46 + i * 3 + c];
          if (has_next) Urow[i][3 + c] = Sigma * Dr[55 + i * 3 + c];
        }
        brow[i] -= Sigma * Dr[43 + i];
      }
    }
  } else {
    // position rows 0..2 (q = 0) or velocity rows 6..8 (q = 2): pv index pa = 3*(q/2) + i
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const int a = (q == 0) ? i : 6 + i;
      const int pa = (q == 0) ? i : 3 + i;
      if (q == 0) {
#pragma unroll
        for (int c = 0; c < 6; c++) Drow[i][c] = invw * G[a <= c ? sym_index(a, c) : sym_index(c, a)];
        brow[i] = invw * G[21 + a];
      }
      if (has_next) {
        double col[6];      // dv2[k] * Phi[k][pa]
#pragma unroll
        for (int k = 0; k < 6; k++) col[k] = dv2[k] * Dr[k * 6 + pa];
#pragma unroll
        for (int pc = 0; pc < 6; pc++) {
          double s = 0.0;
#pragma unroll
          for (int k = 0; k < 6; k++) s = fma(col[k], Dr[k * 6 + pc], s);
          Drow[i][pc < 3 ? pc : pc + 3] += Sigma * s;
          Urow[i][pc < 3 ? pc : pc + 3] = -Sigma * Dr[pc * 6 + pa] * dv2[pc];
        }
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 6; k++) s = fma(Dr[k * 6 + pa] * dv1[k], Dr[36 + k], s);
        brow[i] -= Sigma * s;
      }
      if (has_prev) {
        if (q == 0) Drow[i][i] += Sigma * dv2[i];
        else Drow[i][6 + i] += Sigma * dv2[3 + i];
        brow[i] += Sigma * dv1[pa] * rp[pa];
      }
    }
  }
  const int r0 = 3 * q;
#pragma unroll
  for (int i = 0; i < 3; i++) {
#pragma unroll
    for (int c = 0; c < 9; c++) {
      out[(r0 + i) * 9 + c] = Drow[i][c];
      out[81 + (r0 + i) * 9 + c] = Urow[i][c];
    }
    out[162 + r0 + i] = brow[i];
  }
  if (q == 0) out[171] = 0.0;
  }
  __syncthreads();
  // coalesced write-out of the staged records
  double* dst = srec + f0 * VS_SREC;
  for (int i = tid; i < nf * VS_SREC; i += 96) dst[i] = so[(i / VS_SREC) * kSysOut + (i % VS_SREC)];
}

int launch_system_build(vinsat_batch* b, int initialize, double Sigma, double vel_coeff) {
  vinsat_ctx* ctx = b->ctx;
  if (b->T == 0) return VINSAT_OK;
  const int smem = kSysFrames * (kSysStride + kSysOut) * (int)sizeof(double);
  VS_SMEM_OPTIN(ctx, SM_SYSROWS, k_system_rows, smem);
  VS_LAUNCH(ctx, F_SYSTEM, k_system_rows, ceil_div(b->T, kSysFrames), 96, smem, b->T, b->gap, b->fprob, b->wmax,
            b->grec, b->drec, initialize, Sigma, vel_coeff, b->srec);
  return VINSAT_OK;
}

// ---------------------------------------------------------------------------------------------------------
// per-problem reductions: one warp per problem, lane-strided over the problem's frames, shuffle tree
// (fixed order => deterministic accept decisions)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over the kPT threads that share a problem: one warp (kPT = 32, several problems per CTA) or the whole CTA
// (kPT = blockDim.x = 1024, one problem per CTA: a 100 000-frame arc summed by one warp took 1.9 ms per call).
// Fixed shape => deterministic.  Result valid in the group's first thread.
template <int kPT>
__device__ __forceinline__ double group_total(double v, double* s_part) {
  v = warp_sum(v);
  if (kPT == 32) return v;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();                       // s_part may still be read from the previous call
  if (lane == 0) s_part[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (warp == 0) {
    t = (lane < kPT / 32) ? s_part[lane] : 0.0;
    t = warp_sum(t);
  }
  return t;
}

// Long arcs (few problems, millions of frames each): a per-problem sum by ONE CTA is a latency-bound walk (5 ms for 2.4 M
// frames).  Two stages instead: stage 1 sums fixed chunks of kSumChunk frames, one CTA each, into part[chunk][3]; stage 2
// (k_init_residual / k_accept below, or k_la_sums_final of the sharded arc) adds a problem's chunk sums.  Chunk c of problem
// p covers frames [frame_off[p] + (c - chunk_off[p]) * kSumChunk, ...) -- fixed shapes, fixed order => still deterministic.
//   MODE 0: linearisation sums  { sum grec[f][27], sum_{pairs} |r_pred| (7 components), sum e_prior }
//   MODE 1: trial sums          { sum e_obs[f],    sum_{pairs} e_dyn[f],                sum e_prior }
// `lo`/`hi` (hi > lo) restrict the sums to the OWNED frames of a window batch (one problem).
template <int MODE>
__global__ void __launch_bounds__(kSumThreads) k_sum_partials(int P, const int32_t* __restrict__ chunk_off,
                                                              const int64_t* __restrict__ frame_off, int64_t lo, int64_t hi,
                                                              const int32_t* __restrict__ gap,
                                                              const double* __restrict__ grec,
                                                              const double* __restrict__ drec,
                                                              const double* __restrict__ e_obs,
                                                              const double* __restrict__ e_dyn,
                                                              const double* __restrict__ e_prior, int initialize,
                                                              const int32_t* __restrict__ gate, double* __restrict__ part) {
  if (gate && *gate) return;
  __shared__ double s_part[3][kSumThreads / 32];
  const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int p = 0;
  {                                       // problem of the chunk: last p with chunk_off[p] <= c
    int a = 0, b = P;
    while (b - a > 1) { const int m = (a + b) >> 1; if (chunk_off[m] <= c) a = m; else b = m; }
    p = a;
  }
  const int64_t pf1 = frame_off[p + 1];
  int64_t f0 = frame_off[p], f1 = pf1;
  const bool window = hi > lo;
  if (window) { f0 = lo; f1 = hi; }
  const int64_t c0 = f0 + (int64_t)(c - chunk_off[p]) * kSumChunk, c1 = min(c0 + kSumChunk, f1);
  double so = 0.0, sd = 0.0, sp = 0.0;
  for (int64_t f = c0 + tid; f < c1; f += kSumThreads) {
    if (e_prior) sp += e_prior[f];
    // same pair rule as the one-stage kernels this replaces (k_accept: f + 1 < f1; the others: gap[f] > 0)
    const bool pair = !initialize && ((MODE == 1 && !window) ? (f + 1 < pf1) : (gap[f] > 0));
    if (MODE == 0) {
      so += grec[f * VS_GREC + 27];
      if (pair) {
        const double* d = drec + f * VS_DREC + 36;
        sd += fabs(d[0]) + fabs(d[1]) + fabs(d[2]) + fabs(d[3]) + fabs(d[4]) + fabs(d[5]) + fabs(d[6]);
      }
    } else {
      so += e_obs[f];
      if (pair) sd += e_dyn[f];
    }
  }
  so = warp_sum(so); sd = warp_sum(sd); sp = warp_sum(sp);
  if (lane == 0) { s_part[0][warp] = so; s_part[1][warp] = sd; s_part[2][warp] = sp; }
  __syncthreads();
  if (tid < 3) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kSumThreads / 32; w++) t += s_part[tid][w];
    part[(int64_t)c * 3 + tid] = t;
  }
}

int launch_sum_partials(vinsat_batch* b, int mode, int initialize, const double* e_prior, int64_t lo, int64_t hi) {
  vinsat_ctx* ctx = b->ctx;
  const unsigned n = (unsigned)((hi > lo) ? ceil_div(hi - lo, kSumChunk) : b->sum_chunks);
  if (n == 0) return VINSAT_OK;
  if (mode == 0)
    VS_LAUNCH(ctx, F_ACCEPT, k_sum_partials<0>, n, kSumThreads, 0, (int)b->P, b->sum_chunk_off, b->d_frame_off, lo, hi, b->gap,
              b->grec, b->drec, b->e_obs, b->e_dyn, e_prior, initialize, b->gate_arg, b->sum_part);
  else
    VS_LAUNCH(ctx, F_ACCEPT, k_sum_partials<1>, n, kSumThreads, 0, (int)b->P, b->sum_chunk_off, b->d_frame_off, lo, hi, b->gap,
              b->grec, b->drec, b->e_obs, b->e_dyn, e_prior, initialize, b->gate_arg, b->sum_part);
  return VINSAT_OK;
}

// init_residual = mean |[r_obs ; sqrt(Sigma) r_pred]| (BA_filtering.py:51); also arms the LM loop.
// `part` != null: the sums come from the chunk sums of k_sum_partials (long arcs).
template <int kPT>
__global__ void __launch_bounds__(kPT == 32 ? 128 : kPT) k_init_residual(int P, const int64_t* __restrict__ frame_off,
                                                       const int64_t* __restrict__ obs_off,
                                                       const int32_t* __restrict__ gap,
                                                       const double* __restrict__ grec,
                                                       const double* __restrict__ drec, int initialize,
                                                       double sqrt_sigma, const double* __restrict__ lam_in,
                                                       double* __restrict__ lam, double* __restrict__ init_res,
                                                       int32_t* __restrict__ active, int32_t* __restrict__ ntrials,
                                                       const int32_t* __restrict__ gate,
                                                       const double* __restrict__ e_prior,
                                                       const double* __restrict__ part,
                                                       const int32_t* __restrict__ chunk_off) {
  if (gate && *gate) return;      // speculative launch behind an LM loop that is not finished (batch.cu)
  __shared__ double s_part[32];
  const int p = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / kPT);
  const int lane = threadIdx.x % kPT;
  if (p >= P) return;                    // whole groups leave together (kPT divides blockDim.x)
  const int64_t f0 = frame_off[p], f1 = frame_off[p + 1];
  double so = 0.0, sd = 0.0, sp = 0.0;
  if (part) {
    for (int c = chunk_off[p] + lane; c < chunk_off[p + 1]; c += kPT) {
      so += part[(int64_t)c * 3]; sd += part[(int64_t)c * 3 + 1]; sp += part[(int64_t)c * 3 + 2];
    }
  } else
  for (int64_t f = f0 + lane; f < f1; f += kPT) {
    if (e_prior) sp += e_prior[f];        // BA_reg: r_prior enters unscaled, 7 components per frame (BA_filtering.py:163)
    so += grec[f * VS_GREC + 27];
    if (!initialize && gap[f] > 0) {
      const double* d = drec + f * VS_DREC + 36;
      sd += fabs(d[0]) + fabs(d[1]) + fabs(d[2]) + fabs(d[3]) + fabs(d[4]) + fabs(d[5]) + fabs(d[6]);
    }
  }
  so = group_total<kPT>(so, s_part);
  sd = group_total<kPT>(sd, s_part);
  if (e_prior) sp = group_total<kPT>(sp, s_part);
  if (lane == 0) {
    const double n = 2.0 * (double)(obs_off[p + 1] - obs_off[p]) + (initialize ? 6.0 : 7.0) * (double)max((long long)(f1 - f0 - 1), 0ll)
                     + (e_prior ? 7.0 * (double)(f1 - f0) : 0.0);
    init_res[p] = (so + sqrt_sigma * sd + sp) / n;
    lam[p] = lam_in[p];
    active[p] = (f1 > f0) ? 1 : 0;
    ntrials[p] = 0;
  }
}

int launch_init_residual(vinsat_batch* b, int initialize, double Sigma, double, const double* d_lam_in, const double* e_prior) {
  vinsat_ctx* ctx = b->ctx;
  if (b->P == 0) return VINSAT_OK;
  if (b->T > 4096 * b->P) {              // long arcs: chunk sums by many CTAs, then one CTA per problem adds them
    if (int rc = launch_sum_partials(b, 0, initialize, e_prior, 0, 0)) return rc;
    VS_LAUNCH(ctx, F_ACCEPT, k_init_residual<1024>, (unsigned)b->P, 1024, 0, (int)b->P, b->d_frame_off,
              b->d_obs_off, b->gap, b->grec, b->drec, initialize, sqrt(Sigma), d_lam_in, b->lam, b->init_res,
              b->active, b->ntrials, b->gate_arg, e_prior, b->sum_part, b->sum_chunk_off);
  } else {
    VS_LAUNCH(ctx, F_ACCEPT, k_init_residual<32>, ceil_div(b->P * 32, 128), 128, 0, (int)b->P, b->d_frame_off,
              b->d_obs_off, b->gap, b->grec, b->drec, initialize, sqrt(Sigma), d_lam_in, b->lam, b->init_res,
              b->active, b->ntrials, b->gate_arg, e_prior, nullptr, nullptr);
  }
  return VINSAT_OK;
}

// accept test (BA_filtering.py:66-79)
template <int kPT>
__global__ void __launch_bounds__(kPT == 32 ? 128 : kPT) k_accept(int P, const int64_t* __restrict__ frame_off,
                                                const int64_t* __restrict__ obs_off,
                                                const unsigned long long* __restrict__ wmax,
                                                const double* __restrict__ e_obs, const double* __restrict__ e_dyn,
                                                int initialize, double sqrt_sigma,
                                                const double* __restrict__ init_res, double* __restrict__ lam,
                                                double* __restrict__ lam_next, int32_t* __restrict__ active,
                                                int32_t* __restrict__ ntrials, int32_t* __restrict__ flags,
                                                const int32_t* __restrict__ gate, const double* __restrict__ e_prior,
                                                const double* __restrict__ part, const int32_t* __restrict__ chunk_off) {
  if (gate && *gate) return;      // speculative launch behind an LM loop that is not finished (batch.cu)
  __shared__ double s_part[32];
  const int p = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / kPT);
  const int lane = threadIdx.x % kPT;
  if (p >= P) return;
  if (!active[p]) return;                // uniform over the group
  const int64_t f0 = frame_off[p], f1 = frame_off[p + 1];
  double so = 0.0, sd = 0.0, sp = 0.0;
  if (part) {
    for (int c = chunk_off[p] + lane; c < chunk_off[p + 1]; c += kPT) {
      so += part[(int64_t)c * 3]; sd += part[(int64_t)c * 3 + 1]; sp += part[(int64_t)c * 3 + 2];
    }
  } else
  for (int64_t f = f0 + lane; f < f1; f += kPT) {
    if (e_prior) sp += e_prior[f];
    so += e_obs[f];
    if (!initialize && f + 1 < f1) sd += e_dyn[f];
  }
  so = group_total<kPT>(so, s_part);
  sd = group_total<kPT>(sd, s_part);
  if (e_prior) sp = group_total<kPT>(sp, s_part);
  if (lane == 0) {
    const unsigned long long wb = wmax[p];
    const double invw = wb ? 1.0 / __longlong_as_double((long long)wb) : 0.0;
    const double n = 2.0 * (double)(obs_off[p + 1] - obs_off[p]) + (initialize ? 6.0 : 7.0) * (double)max((long long)(f1 - f0 - 1), 0ll)
                     + (e_prior ? 7.0 * (double)(f1 - f0) : 0.0);
    const double residual = (invw * so + sqrt_sigma * sd + sp) / n;
    const double l = lam[p] * 10.0;                       // :72
    lam[p] = l;
    ntrials[p] += 1;
    const bool done = (residual < init_res[p]) || (l > 1e4);   // :73-77
    if (done) {
      active[p] = 0;
      lam_next[p] = fmax(fmin(1e-1, l * 0.01), 1e-4);     // :79
    } else {
      atomicAdd(&flags[0], 1);
    }
  }
}

int launch_accept(vinsat_batch* b, int initialize, double Sigma, const double* e_prior) {
  vinsat_ctx* ctx = b->ctx;
  if (b->P == 0) return VINSAT_OK;
  if (b->T > 4096 * b->P) {
    if (int rc = launch_sum_partials(b, 1, initialize, e_prior, 0, 0)) return rc;
    VS_LAUNCH(ctx, F_ACCEPT, k_accept<1024>, (unsigned)b->P, 1024, 0, (int)b->P, b->d_frame_off, b->d_obs_off,
              b->wmax, b->e_obs, b->e_dyn, initialize, sqrt(Sigma), b->init_res, b->lam, b->lam_next, b->active,
              b->ntrials, b->flags, b->gate_arg, e_prior, b->sum_part, b->sum_chunk_off);
  } else {
    VS_LAUNCH(ctx, F_ACCEPT, k_accept<32>, ceil_div(b->P * 32, 128), 128, 0, (int)b->P, b->d_frame_off, b->d_obs_off,
              b->wmax, b->e_obs, b->e_dyn, initialize, sqrt(Sigma), b->init_res, b->lam, b->lam_next, b->active,
              b->ntrials, b->flags, b->gate_arg, e_prior, nullptr, nullptr);
  }
  return VINSAT_OK;
}

// initialize phase (BA_utils.py:463-466 => no dynamics terms): the system is block diagonal, every frame is
// an independent SPD 6x6 pose block (+ lam I) and three decoupled velocity rows with zero right-hand side.
// One thread per frame, Cholesky in registers, straight from the observation record (no srec round trip).
__global__ void __launch_bounds__(128) k_solve_init(int64_t T, const int32_t* __restrict__ fprob,
                                                    const int32_t* __restrict__ active,
                                                    const double* __restrict__ lam,
                                                    const unsigned long long* __restrict__ wmax,
                                                    const double* __restrict__ grec, double* __restrict__ delta,
                                                    double* __restrict__ lam32_last, const int32_t* __restrict__ gate) {
  if (gate && *gate) return;      // speculative launch behind an LM loop that is not finished (batch.cu)
  const int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (f >= T) return;
  const int p = fprob[f];
  if (!active[p]) return;
  const double lam32 = (double)(float)lam[p];
  lam32_last[p] = lam32;     // same value from every frame of the problem
  const unsigned long long wb = wmax[p];
  const double invw = wb ? 1.0 / __longlong_as_double((long long)wb) : 0.0;
  const double* g = grec + f * VS_GREC;
  double L[6][6], y[6];
  {
    int idx = 0;
#pragma unroll
    for (int a = 0; a < 6; a++)
#pragma unroll
      for (int b = a; b < 6; b++) { L[b][a] = invw * g[idx++] + (a == b ? lam32 : 0.0); }
#pragma unroll
    for (int a = 0; a < 6; a++) y[a] = invw * g[21 + a];
  }
  // in-place Cholesky of the lower triangle
#pragma unroll
  for (int j = 0; j < 6; j++) {
    double d = L[j][j];
#pragma unroll
    for (int k = 0; k < j; k++) d = fma(-L[j][k], L[j][k], d);
    d = sqrt(d);
    L[j][j] = d;
    const double inv = 1.0 / d;
#pragma unroll
    for (int i = j + 1; i < 6; i++) {
      double s = L[i][j];
#pragma unroll
      for (int k = 0; k < j; k++) s = fma(-L[i][k], L[j][k], s);
      L[i][j] = s * inv;
    }
  }
#pragma unroll
  for (int i = 0; i < 6; i++) {
    double s = y[i];
#pragma unroll
    for (int k = 0; k < i; k++) s = fma(-L[i][k], y[k], s);
    y[i] = s / L[i][i];
  }
#pragma unroll
  for (int i = 5; i >= 0; i--) {
    double s = y[i];
#pragma unroll
    for (int k = i + 1; k < 6; k++) s = fma(-L[k][i], y[k], s);
    y[i] = s / L[i][i];
  }
  double* d = delta + f * 9;
#pragma unroll
  for (int i = 0; i < 6; i++) d[i] = y[i];
  d[6] = 0.0; d[7] = 0.0; d[8] = 0.0;
}

// retraction (BA_filtering.py:56-60): p + dp, normalize(q (x) exp(dtheta)), v + dv; one thread per frame, rows
// staged through shared memory so that the 80 B / 72 B records move as coalesced runs
__global__ void __launch_bounds__(128) k_retract(int64_t T, const int32_t* __restrict__ fprob,
                                                 const int32_t* __restrict__ active,
                                                 const double* __restrict__ st, const double* __restrict__ delta,
                                                 double* __restrict__ st_new, const int32_t* __restrict__ gate) {
  if (gate && *gate) return;      // speculative launch behind an LM loop that is not finished (batch.cu)
  __shared__ double s_st[128 * 10 + 1];
  __shared__ double s_d[128 * 9];
  __shared__ int s_act[128];
  const int tid = threadIdx.x;
  const int64_t f0 = (int64_t)blockIdx.x * 128;
  const int nf = (int)min((int64_t)128, T - f0);
  s_act[tid] = (tid < nf) ? active[fprob[f0 + tid]] : 0;
  for (int i = tid; i < nf * 10; i += 128) s_st[i] = st[f0 * 10 + i];
  for (int i = tid; i < nf * 9; i += 128) s_d[i] = delta[f0 * 9 + i];
  __syncthreads();
  if (tid < nf && s_act[tid]) {
    double* s = s_st + tid * 10;
    const double* d = s_d + tid * 9;
    const Quat q = {s[3], s[4], s[5], s[6]};
    const Quat n = qmul(q, qexp(d[3], d[4], d[5]));
    const double nn = sqrt(n.x * n.x + n.y * n.y + n.z * n.z + n.w * n.w);
    s[0] += d[0]; s[1] += d[1]; s[2] += d[2];
    s[7] += d[6]; s[8] += d[7]; s[9] += d[8];
    s[3] = n.x / nn; s[4] = n.y / nn; s[5] = n.z / nn; s[6] = n.w / nn;
  }
  __syncthreads();
  for (int i = tid; i < nf * 10; i += 128)
    if (s_act[i / 10]) st_new[f0 * 10 + i] = s_st[i];
}

int launch_solve_init_only(vinsat_batch* b) {
  vinsat_ctx* ctx = b->ctx;
  if (b->T == 0) return VINSAT_OK;
  VS_LAUNCH(ctx, F_SOLVE_INIT, k_solve_init, ceil_div(b->T, 128), 128, 0, b->T, b->fprob, b->active, b->lam, b->wmax,
            b->grec, b->delta, b->lam32_last, b->gate_arg);
  return VINSAT_OK;
}

int launch_retract_only(vinsat_batch* b) {
  vinsat_ctx* ctx = b->ctx;
  if (b->T == 0) return VINSAT_OK;
  VS_LAUNCH(ctx, F_RETRACT, k_retract, ceil_div(b->T, 128), 128, 0, b->T, b->fprob, b->active, b->st, b->delta,
            b->st_new, b->gate_arg);
  return VINSAT_OK;
}

int launch_solve_retract(vinsat_batch* b, int initialize) {
  vinsat_ctx* ctx = b->ctx;
  if (b->P == 0 || b->T == 0) return VINSAT_OK;
  if (initialize) {
    VS_LAUNCH(ctx, F_SOLVE_INIT, k_solve_init, ceil_div(b->T, 128), 128, 0, b->T, b->fprob, b->active, b->lam, b->wmax,
              b->grec, b->delta, b->lam32_last, b->gate_arg);
  } else {
    int rc = launch_chain_solve(b);
    if (rc != VINSAT_OK) return rc;
  }
  VS_LAUNCH(ctx, F_RETRACT, k_retract, ceil_div(b->T, 128), 128, 0, b->T, b->fprob, b->active, b->st, b->delta,
            b->st_new, b->gate_arg);
  return VINSAT_OK;
}

// Zeroing that respects the speculation gate (a cudaMemsetAsync cannot be skipped on the device).
__global__ void k_gated_zero(unsigned long long* __restrict__ p64, int64_t n64, int32_t* __restrict__ p32, int n32,
                             const int32_t* __restrict__ gate) {
  if (gate && *gate) return;
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n64) p64[i] = 0ull;
  if (i < n32) p32[i] = 0;
}

int launch_gated_zero(vinsat_batch* b, unsigned long long* p64, int64_t n64, int32_t* p32, int n32) {
  vinsat_ctx* ctx = b->ctx;
  const int64_t n = n64 > n32 ? n64 : n32;
  if (n <= 0) return VINSAT_OK;
  VS_LAUNCH(ctx, F_ACCEPT, k_gated_zero, ceil_div(n, 256), 256, 0, p64, n64, p32, n32, b->gate_arg);
  return VINSAT_OK;
}

// After the first trial of a speculatively pipelined iteration: gate <- number of problems whose LM loop goes on.
// While it is non-zero every kernel launched behind it returns at once (the host then finishes that loop).
__global__ void k_gate_publish(int32_t* __restrict__ gate, const int32_t* __restrict__ flags) {
  if (*gate == 0) *gate = flags[0];
}

int launch_gate_publish(vinsat_batch* b) {
  vinsat_ctx* ctx = b->ctx;
  VS_LAUNCH(ctx, F_ACCEPT, k_gate_publish, 1, 1, 0, b->gate, b->flags);
  return VINSAT_OK;
}


// ---------------------------------------------------------------------------------------------------------
// JTwJ[:, -9:, -9:] of every problem in one pass (BA_filtering.py:97): D block of the problem's last frame plus the
// float32 damping on the diagonal, gathered into out[P][81] on the device (one D2H copy instead of one per problem).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_gather_last_hessian(int64_t P, const int64_t* __restrict__ frame_off,
                                                             const double* __restrict__ srec,
                                                             const double* __restrict__ lam32, double* __restrict__ out) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t p = t / 81;
  const int k = (int)(t % 81);
  if (p >= P) return;
  const int64_t fl = frame_off[p + 1] - 1;
  double v = 0.0;
  if (fl >= frame_off[p]) v = srec[fl * VS_SREC + k] + ((k % 10 == 0) ? lam32[p] : 0.0);
  out[p * 81 + k] = v;
}

int launch_gather_last_hessian(vinsat_batch* b, double* out_dev) {
  vinsat_ctx* ctx = b->ctx;
  VS_LAUNCH(ctx, F_LAYOUT, k_gather_last_hessian, ceil_div(b->P * 81, 128), 128, 0, b->P, b->d_frame_off, b->srec,
            b->lam32_last, out_dev);
  return VINSAT_OK;
}

// Streaming windows (od_pipe.py:1011-1019): frame f of the new window takes the state the chain reached at its time,
// states_prop[:, time_idx_prop - time_idx_prop[0]]; chain row k is the state k seconds after the chain start t0.
__global__ void __launch_bounds__(128) k_stream_gather(int64_t n, const double* __restrict__ chain,
                                                       const int64_t* __restrict__ time_idx, int64_t f0, int64_t t0,
                                                       double* __restrict__ st, double* __restrict__ vel,
                                                       double* __restrict__ seed) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= n) return;
  const double* c = chain + (time_idx[f0 + j] - t0) * 10;
#pragma unroll
  for (int k = 0; k < 10; k++) {
    const double v = c[k];
    if (st) st[(f0 + j) * 10 + k] = v;
    seed[(f0 + j) * 10 + k] = v;
  }
  vel[(f0 + j) * 3] = c[7]; vel[(f0 + j) * 3 + 1] = c[8]; vel[(f0 + j) * 3 + 2] = c[9];
}

int launch_stream_gather(vinsat_ctx* ctx, int64_t n, const double* chain, const int64_t* time_idx, int64_t f0, int64_t t0,
                         double* st, double* vel, double* seed) {
  if (n <= 0) return VINSAT_OK;
  VS_LAUNCH(ctx, F_LAYOUT, k_stream_gather, ceil_div(n, 128), 128, 0, n, chain, time_idx, f0, t0, st, vel, seed);
  return VINSAT_OK;
}

}  // namespace vs
