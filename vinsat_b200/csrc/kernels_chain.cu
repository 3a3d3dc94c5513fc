// Batched block-tridiagonal LU solve (a7 of SURVEY.md section 8; BA_filtering.py:54-55), partitioned.
//
// A problem's frames form a block-tridiagonal system with 9x9 blocks D_i, U_i = A(i,i+1) and
// Lo_i = A(i+1,i) (= U_i^T for the BA system, explicit for reduced systems).  The T sequential 9x9 eliminations
// of a problem are a latency chain, so each problem is cut into SEGMENTS: the last frame of a segment is a
// separator, the interior frames a..b-1 of all segments are eliminated in parallel with a "spike" block
// that carries the dependence on the left separator,
//     x_i = y_i - W_i x_{i+1} - Z_i x_left ,          [W_i | y_i | Z_i] = S_i^-1 [U_i | b~_i | Z~_i],
//     S_{i+1} = D_{i+1} - Lo_i W_i,  b~_{i+1} = b_{i+1} - Lo_i y_i,  Z~_{i+1} = -Lo_i Z_i,  Z~_a = Lo_{a-1}.
// A backward recurrence then gives x_a = yh - Wh x_right - Zh x_left in closed form, and the separators obey a
// small block-tridiagonal REDUCED system (one 9x9 row per segment) which the same elimination solves (plain
// mode, explicit lower blocks).  A last pass back-substitutes the interiors.
//
// Mapping: 8 lanes per chain, 4 chains per warp.  The augmented block [S | U | b | Z] (9 x 28) is held by
// COLUMNS: lane l of a group owns (S_l, U_l, Z_l) and lanes 0..3 own one extra column each (S_8, U_8, Z_8, b).
// Gauss-Jordan without pivoting (the symmetric part of every pivot block is positive definite):
// the pivot column is broadcast with group shuffles, every lane then updates its 3-4 columns (24-32
// independent DFMAs per step), so all 32 lanes do useful FP64 work and the kernel is FP64-pipe / latency
// bound instead of issue bound.
#include "common.cuh"
#include "launch.h"

namespace vs {

constexpr int kGL = 8;            // lanes per chain
constexpr int kCPW = 32 / kGL;    // chains per warp

__device__ __forceinline__ double fast_rcp_c(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}

__device__ __forceinline__ double gshfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src, kGL); }

// out[r] = sum_k M[k*9 + r] * v[k]   (M in shared memory, broadcast reads)
__device__ __forceinline__ void matvec9(const double* __restrict__ M, const double* v, double* out) {
#pragma unroll
  for (int r = 0; r < 9; r++) out[r] = 0.0;
#pragma unroll
  for (int k = 0; k < 9; k++) {
    const double vk = v[k];
#pragma unroll
    for (int r = 0; r < 9; r++) out[r] = fma(M[k * 9 + r], vk, out[r]);
  }
}

struct ChainArgs {
  int n_chains;
  const int32_t* ch_a;       // first element
  const int32_t* ch_b;       // SEG: separator element (interior [a,b)); PLAIN: one past the last element
  const int32_t* ch_left;    // SEG: left separator element or -1
  const int32_t* ch_prob;    // problem of the chain
  const int32_t* active;     // [P] or null
  const double* lam;         // [P] damping (float32-rounded inside) or null (no damping: reduced systems)
  const double* rec;         // [n][VS_SREC]  D | U | b
  const double* lrec;        // explicit lower blocks: lrec[i] = A(i+1,i) row-major, or null (=> U_i^T)
  double* wrec;              // [n][VS_WREC]  W (col-major) | y | Z (col-major)
  double* redrec;            // SEG: [n_chains][VS_RREC] contributions to the reduced system
  double* delta;             // PLAIN: solution rows; SEG back-substitution: in/out
  const int32_t* out_index;  // PLAIN: element -> row of delta (null = identity)
  double* lam32_last;        // [P] or null
};

// ---------------------------------------------------------------------------------------------------------
// forward elimination of one chain per 8-lane group.  SPIKE=true: segment mode.
// ---------------------------------------------------------------------------------------------------------
template <bool SPIKE>
__global__ void __launch_bounds__(32) k_chain(ChainArgs A) {
  __shared__ double s_M[kCPW][81];
  __shared__ double s_W[kCPW][81];
  const int lane = threadIdx.x & 31;
  const int g = lane / kGL, gl = lane % kGL;
  const int ch = blockIdx.x * kCPW + g;
  const bool valid = ch < A.n_chains;
  int a = 0, e = 0, left = -1, prob = 0;
  bool live = false;
  if (valid) {
    a = A.ch_a[ch]; e = A.ch_b[ch]; prob = A.ch_prob[ch];
    if (SPIKE) left = A.ch_left[ch];
    live = A.active ? (A.active[prob] != 0) : true;
  }
  if (!live) { a = 0; e = 0; left = -1; }
  double lam32 = 0.0;
  if (live && A.lam) {
    lam32 = (double)(float)A.lam[prob];           // torch.eye(n)*lamda is float32
    if (A.lam32_last && gl == 0) A.lam32_last[prob] = lam32;
  }
  // the groups of a warp run in lock step over max(len)
  int len = e - a;
  int maxlen = len;
#pragma unroll
  for (int o = 16; o >= kGL; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
  double* M = s_M[g];
  double* Wst = s_W[g];
  const bool has_x = gl < 4;                      // owns an extra column: 0:S_8 1:U_8 2:Z_8 3:b
  double c0[9], c1[9], c2[9], c3[9];
  double n0[9], n1[9], n3[9];
#pragma unroll
  for (int r = 0; r < 9; r++) { c0[r] = c1[r] = c2[r] = c3[r] = 0.0; n0[r] = n1[r] = n3[r] = 0.0; }

  auto load_cols = [&](int i, double* d0, double* d1, double* d3) {
    const double* rec = A.rec + (int64_t)i * VS_SREC;
#pragma unroll
    for (int r = 0; r < 9; r++) {
      d0[r] = rec[r * 9 + gl];
      d1[r] = rec[81 + r * 9 + gl];
      double x = 0.0;
      if (gl == 0) x = rec[r * 9 + 8];
      else if (gl == 1) x = rec[81 + r * 9 + 8];
      else if (gl == 3) x = rec[162 + r];
      d3[r] = x;
    }
  };

  if (len > 0) {
    load_cols(a, n0, n1, n3);
    if (SPIKE && left >= 0) {
      // Z~_a = Lo_left : column c of Lo_left.  lrec: Lo[r][c] ; base: Lo[r][c] = U_left[c][r]
      if (A.lrec) {
        const double* L = A.lrec + (int64_t)left * 81;
#pragma unroll
        for (int r = 0; r < 9; r++) { c2[r] = L[r * 9 + gl]; if (gl == 2) c3[r] = L[r * 9 + 8]; }
      } else {
        const double* U = A.rec + (int64_t)left * VS_SREC + 81;
#pragma unroll
        for (int r = 0; r < 9; r++) { c2[r] = U[gl * 9 + r]; if (gl == 2) c3[r] = U[8 * 9 + r]; }
      }
    }
  }
  double v0[9], v2[9], v3[9];
#pragma unroll
  for (int r = 0; r < 9; r++) { v0[r] = v2[r] = v3[r] = 0.0; }
  bool first = true;
  double w8[9];
#pragma unroll
  for (int r = 0; r < 9; r++) w8[r] = 0.0;

  for (int t = 0; t < maxlen; t++) {
    const int i = a + t;
    const bool on = t < len;
    if (on) {
      // assemble the augmented columns of element i
#pragma unroll
      for (int r = 0; r < 9; r++) {
        c0[r] = n0[r] + (r == gl ? lam32 : 0.0) - v0[r];
        c1[r] = n1[r];
        if (!first) c2[r] = -v2[r];
        double x = n3[r];
        if (gl == 0) x = n3[r] + (r == 8 ? lam32 : 0.0) - v3[r];
        else if (gl == 2) x = first ? c3[r] : -v3[r];
        else if (gl == 3) x = n3[r] - v3[r];
        c3[r] = x;
      }
      first = false;
      if (t + 1 < len) load_cols(i + 1, n0, n1, n3);
      // lower block Lo_i (for the update of element i+1), staged as M[k*9+r] = Lo_i[r][k]
      if (A.lrec) {
        const double* L = A.lrec + (int64_t)i * 81;
        for (int idx = gl; idx < 81; idx += kGL) { const int r = idx / 9, k = idx % 9; M[k * 9 + r] = L[idx]; }
      } else {
#pragma unroll
        for (int r = 0; r < 9; r++) { M[r * 9 + gl] = c1[r]; if (gl == 1) M[r * 9 + 8] = c3[r]; }
      }
    }
    // Gauss-Jordan on [S | U | b | Z]
#pragma unroll
    for (int k = 0; k < 9; k++) {
      double pk[9];
#pragma unroll
      for (int r = 0; r < 9; r++) pk[r] = gshfl(k < 8 ? c0[r] : c3[r], k < 8 ? k : 0);
      if (on) {
        const double inv = fast_rcp_c(pk[k]);
        const double p0 = c0[k] * inv, p1 = c1[k] * inv, p2 = c2[k] * inv, p3 = c3[k] * inv;
#pragma unroll
        for (int r = 0; r < 9; r++) {
          if (r == k) { c0[r] = p0; c1[r] = p1; c2[r] = p2; c3[r] = p3; }
          else {
            c0[r] = fma(-pk[r], p0, c0[r]);
            c1[r] = fma(-pk[r], p1, c1[r]);
            if (SPIKE) c2[r] = fma(-pk[r], p2, c2[r]);
            c3[r] = fma(-pk[r], p3, c3[r]);
          }
        }
      }
    }
    // W_8 lives in lane 1; the S_8 update (lane 0) needs it
#pragma unroll
    for (int r = 0; r < 9; r++) w8[r] = gshfl(c3[r], 1);
    __syncwarp();      // M complete
    if (on) {
      double* w = A.wrec + (int64_t)i * VS_WREC;
#pragma unroll
      for (int r = 0; r < 9; r++) {
        w[gl * 9 + r] = c1[r];
        if (SPIKE) w[90 + gl * 9 + r] = c2[r];
        if (gl == 1) w[8 * 9 + r] = c3[r];
        else if (gl == 2) { if (SPIKE) w[90 + 8 * 9 + r] = c3[r]; }
        else if (gl == 3) w[81 + r] = c3[r];
      }
      // products with Lo_i: for element i+1, or (segment mode, last interior) for the separator's row
      if (t + 1 < len || SPIKE) {
        matvec9(M, c1, v0);
        if (SPIKE) matvec9(M, c2, v2);
        if (has_x && gl != 1) matvec9(M, gl == 0 ? w8 : c3, v3);
      }
    }
    __syncwarp();      // before M is overwritten
  }

  if (!SPIKE) {
    // ---- plain mode: own backward substitution x_i = y_i - W_i x_{i+1}; lane l computes rows l (and 8) ----
    double x[9];
#pragma unroll
    for (int r = 0; r < 9; r++) x[r] = 0.0;
    for (int t = maxlen - 1; t >= 0; t--) {
      const bool on = t < len;
      const int i = a + t;
      double xa = 0.0, xb = 0.0;
      if (on) {
        const double* w = A.wrec + (int64_t)i * VS_WREC;
        xa = w[81 + gl];
        xb = w[81 + 8];
        if (t + 1 < len) {
#pragma unroll
          for (int c = 0; c < 9; c++) { xa = fma(-w[c * 9 + gl], x[c], xa); xb = fma(-w[c * 9 + 8], x[c], xb); }
        }
      }
#pragma unroll
      for (int c = 0; c < 8; c++) { const double v = gshfl(xa, c); if (on) x[c] = v; }
      if (on) {
        x[8] = xb;
        const int64_t row = A.out_index ? A.out_index[i] : i;
        A.delta[row * 9 + gl] = xa;
        if (gl == 0) A.delta[row * 9 + 8] = xb;
      }
    }
    return;
  }

  // ---- segment mode: contributions to the reduced system --------------------------------------------------
  double* rr = valid ? A.redrec + (int64_t)ch * VS_RREC : nullptr;
  // left part (row of this segment's separator b): Dl = -Lo_{b-1} W_{b-1}, bl = -Lo_{b-1} y_{b-1},
  // Ll = -Lo_{b-1} Z_{b-1}  (coefficient of x_left); empty interior: Ll = Lo_left, Dl = bl = 0.
  if (live) {
    if (len > 0) {
#pragma unroll
      for (int r = 0; r < 9; r++) {
        rr[r * 9 + gl] = -v0[r];                      // Dl[r][gl]
        rr[81 + r * 9 + gl] = -v2[r];                 // Ll[r][gl]
        if (gl == 0) rr[r * 9 + 8] = -v3[r];
        else if (gl == 2) rr[81 + r * 9 + 8] = -v3[r];
        else if (gl == 3) rr[162 + r] = -v3[r];
      }
    } else {
#pragma unroll
      for (int r = 0; r < 9; r++) {
        double l0 = 0.0, l8 = 0.0;
        if (left >= 0) {
          if (A.lrec) { l0 = A.lrec[(int64_t)left * 81 + r * 9 + gl]; l8 = A.lrec[(int64_t)left * 81 + r * 9 + 8]; }
          else { l0 = A.rec[(int64_t)left * VS_SREC + 81 + gl * 9 + r]; l8 = A.rec[(int64_t)left * VS_SREC + 81 + 8 * 9 + r]; }
        }
        rr[r * 9 + gl] = 0.0;
        rr[81 + r * 9 + gl] = l0;
        if (gl == 0) rr[r * 9 + 8] = 0.0;
        else if (gl == 2) rr[81 + r * 9 + 8] = l8;
        else if (gl == 3) rr[162 + r] = 0.0;
      }
    }
  }
  // right part (row of the LEFT separator): closed form x_a = yh - Wh x_b - Zh x_left by the backward recurrence
  //   Wh_i = -W_i Wh_{i+1}, Zh_i = Z_i - W_i Zh_{i+1}, yh_i = y_i - W_i yh_{i+1}, started at the last interior.
  // h1 = Wh column gl, h2 = Zh column gl, h3 = extra (lane0: Wh_8, lane2: Zh_8, lane3: yh)
  double h1[9], h2[9], h3[9];
#pragma unroll
  for (int r = 0; r < 9; r++) { h1[r] = c1[r]; h2[r] = c2[r]; h3[r] = (gl == 0) ? w8[r] : c3[r]; }
  const bool need_right = live && left >= 0;
  // segments of different length: a group starts its recurrence when the common counter reaches its own end
  for (int t = maxlen - 2; t >= 0; t--) {
    const bool on = need_right && (t < len - 1);
    const int i = a + t;
    if (on) {
      const double* w = A.wrec + (int64_t)i * VS_WREC;
      for (int idx = gl; idx < 81; idx += kGL) Wst[idx] = w[idx];
    }
    __syncwarp();
    if (on) {
      const double* w = A.wrec + (int64_t)i * VS_WREC;
      double o1[9], o2[9], o3[9];
      matvec9(Wst, h1, o1);
      matvec9(Wst, h2, o2);
      matvec9(Wst, h3, o3);
#pragma unroll
      for (int r = 0; r < 9; r++) {
        h1[r] = -o1[r];
        h2[r] = w[90 + gl * 9 + r] - o2[r];
        double own = 0.0;
        if (gl == 2) own = w[90 + 8 * 9 + r];
        else if (gl == 3) own = w[81 + r];
        h3[r] = own - o3[r];
      }
    }
    __syncwarp();
  }
  if (need_right) {
    // stage U_left transposed: Wst[k*9+r] = U_left[r][k]   (explicit mode: U of the reduced chain is in rec too)
    const double* U = A.rec + (int64_t)left * VS_SREC + 81;
    for (int idx = gl; idx < 81; idx += kGL) { const int r = idx / 9, k = idx % 9; Wst[k * 9 + r] = U[idx]; }
  }
  __syncwarp();
  if (need_right) {
    double* rq = rr + 171;
    if (len > 0) {
      double o1[9], o2[9], o3[9];
      matvec9(Wst, h1, o1);      // U_left Wh  -> Ur = -(...)
      matvec9(Wst, h2, o2);      // U_left Zh  -> Dr = -(...)
      matvec9(Wst, h3, o3);
#pragma unroll
      for (int r = 0; r < 9; r++) {
        rq[r * 9 + gl] = -o2[r];                       // Dr[r][gl]
        rq[81 + r * 9 + gl] = -o1[r];                  // Ur[r][gl]
        if (gl == 0) rq[81 + r * 9 + 8] = -o3[r];      // Ur col 8 (from Wh_8)
        else if (gl == 2) rq[r * 9 + 8] = -o3[r];      // Dr col 8 (from Zh_8)
        else if (gl == 3) rq[162 + r] = -o3[r];        // br
      }
    } else {
      const double* U = A.rec + (int64_t)left * VS_SREC + 81;
#pragma unroll
      for (int r = 0; r < 9; r++) {
        rq[r * 9 + gl] = 0.0;
        rq[81 + r * 9 + gl] = U[r * 9 + gl];
        if (gl == 0) rq[81 + r * 9 + 8] = U[r * 9 + 8];
        else if (gl == 2) rq[r * 9 + 8] = 0.0;
        else if (gl == 3) rq[162 + r] = 0.0;
      }
    }
  } else if (live) {
    double* rq = rr + 171;
    for (int idx = gl; idx < 171; idx += kGL) rq[idx] = 0.0;
  }
}

// ---------------------------------------------------------------------------------------------------------
// reduced system rows: one thread per (separator, element)
//   D~_s = D_b + lam I + Dl_s + Dr_{s+1};  U~_s = Ur_{s+1};  b~_s = b_b + bl_s + br_{s+1};  Lo~_{s-1} = Ll_s
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_reduced_build(int n_seg, const int32_t* __restrict__ ch_b,
                                                       const int32_t* __restrict__ ch_left,
                                                       const int32_t* __restrict__ ch_prob,
                                                       const int32_t* __restrict__ seg_has_next,
                                                       const int32_t* __restrict__ active,
                                                       const double* __restrict__ lam, const double* __restrict__ rec,
                                                       const double* __restrict__ redrec, double* __restrict__ rsys,
                                                       double* __restrict__ rlow) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int s = (int)(t / 192);
  const int e = (int)(t % 192);
  if (s >= n_seg || e >= 171) return;
  const int prob = ch_prob[s];
  if (active && !active[prob]) return;
  const int b = ch_b[s];
  const double* mine = redrec + (int64_t)s * VS_RREC;
  const bool nx = seg_has_next[s] != 0;
  const double* next = redrec + (int64_t)(s + 1) * VS_RREC + 171;
  const double* fr = rec + (int64_t)b * VS_SREC;
  double v;
  if (e < 81) {
    v = fr[e] + mine[e] + (nx ? next[e] : 0.0);
    if (e / 9 == e % 9) v += (double)(float)lam[prob];
  } else if (e < 162) {
    v = nx ? next[e] : 0.0;
  } else {
    v = fr[e] + mine[e] + (nx ? next[e] : 0.0);
  }
  rsys[(int64_t)s * VS_SREC + e] = v;
  // Ll_s = A~(s, s-1) is the lower block stored at reduced element s-1
  if (e < 81 && ch_left[s] >= 0) rlow[(int64_t)(s - 1) * 81 + e] = mine[81 + e];
}

// ---------------------------------------------------------------------------------------------------------
// interior back-substitution of a segment: x_i = y_i - W_i x_{i+1} - Z_i x_left, i = b-1 .. a
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_seg_backsub(int n_chains, const int32_t* __restrict__ ch_a,
                                                    const int32_t* __restrict__ ch_b,
                                                    const int32_t* __restrict__ ch_left,
                                                    const int32_t* __restrict__ ch_prob,
                                                    const int32_t* __restrict__ active,
                                                    const double* __restrict__ wrec, double* __restrict__ delta) {
  const int lane = threadIdx.x & 31;
  const int g = lane / kGL, gl = lane % kGL;
  const int ch = blockIdx.x * kCPW + g;
  int a = 0, b = 0, left = -1;
  bool live = false;
  if (ch < n_chains) {
    a = ch_a[ch]; b = ch_b[ch]; left = ch_left[ch];
    live = active ? (active[ch_prob[ch]] != 0) : true;
  }
  if (!live) { a = 0; b = 0; }
  const int len = b - a;
  int maxlen = len;
#pragma unroll
  for (int o = 16; o >= kGL; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
  double x[9], xl[9];
#pragma unroll
  for (int r = 0; r < 9; r++) {
    x[r] = live ? delta[(int64_t)b * 9 + r] : 0.0;
    xl[r] = (live && left >= 0) ? delta[(int64_t)left * 9 + r] : 0.0;
  }
  for (int t = maxlen - 1; t >= 0; t--) {
    const bool on = t < len;
    const int i = a + t;
    double xa = 0.0, xb = 0.0;
    if (on) {
      const double* w = wrec + (int64_t)i * VS_WREC;
      xa = w[81 + gl];
      xb = w[81 + 8];
#pragma unroll
      for (int c = 0; c < 9; c++) {
        xa = fma(-w[c * 9 + gl], x[c], xa);
        xb = fma(-w[c * 9 + 8], x[c], xb);
        xa = fma(-w[90 + c * 9 + gl], xl[c], xa);
        xb = fma(-w[90 + c * 9 + 8], xl[c], xb);
      }
    }
#pragma unroll
    for (int c = 0; c < 8; c++) { const double v = gshfl(xa, c); if (on) x[c] = v; }
    if (on) {
      x[8] = xb;
      delta[(int64_t)i * 9 + gl] = xa;
      if (gl == 0) delta[(int64_t)i * 9 + 8] = xb;
    }
  }
}

int launch_chain_solve(vinsat_batch* b) {
  vinsat_ctx* ctx = b->ctx;
  if (b->P == 0 || b->T == 0) return VINSAT_OK;
  ChainArgs A;
  A.active = b->active;
  A.lam = b->lam;
  A.rec = b->srec;
  A.lrec = nullptr;
  A.wrec = b->wrec;
  A.delta = b->delta;
  A.out_index = nullptr;
  A.lam32_last = b->lam32_last;
  A.redrec = b->redrec;
  if (!b->partitioned) {
    A.n_chains = (int)b->P;
    A.ch_a = b->pl_a; A.ch_b = b->pl_b; A.ch_left = nullptr; A.ch_prob = b->pl_prob;
    VS_LAUNCH(ctx, F_SOLVE, k_chain<false>, ceil_div(A.n_chains, kCPW), 32, 0, A);
    return VINSAT_OK;
  }
  A.n_chains = (int)b->n_seg;
  A.ch_a = b->seg_a; A.ch_b = b->seg_b; A.ch_left = b->seg_left; A.ch_prob = b->seg_prob;
  VS_LAUNCH(ctx, F_SOLVE, k_chain<true>, ceil_div(A.n_chains, kCPW), 32, 0, A);
  VS_LAUNCH(ctx, F_SOLVE, k_reduced_build, ceil_div((int64_t)b->n_seg * 192, 256), 256, 0, (int)b->n_seg, b->seg_b,
            b->seg_left, b->seg_prob, b->seg_has_next, b->active, b->lam, b->srec, b->redrec, b->rsys, b->rlow);
  ChainArgs R;
  R.n_chains = (int)b->P;
  R.ch_a = b->red_a; R.ch_b = b->red_b; R.ch_left = nullptr; R.ch_prob = b->pl_prob;
  R.active = b->active;
  R.lam = nullptr;
  R.rec = b->rsys;
  R.lrec = b->rlow;
  R.wrec = b->rwrec;
  R.redrec = nullptr;
  R.delta = b->delta;
  R.out_index = b->seg_b;          // separator s -> frame b_s
  R.lam32_last = nullptr;
  VS_LAUNCH(ctx, F_SOLVE, k_chain<false>, ceil_div(R.n_chains, kCPW), 32, 0, R);
  VS_LAUNCH(ctx, F_SOLVE, k_seg_backsub, ceil_div((int64_t)b->n_seg, kCPW), 32, 0, (int)b->n_seg, b->seg_a, b->seg_b,
            b->seg_left, b->seg_prob, b->active, b->wrec, b->delta);
  return VINSAT_OK;
}

}  // namespace vs
