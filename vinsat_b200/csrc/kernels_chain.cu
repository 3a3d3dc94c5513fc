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
// The Monte-Carlo path (many problems) needs no segments: every problem is swept from BOTH ends towards its middle
// frame (two chains per problem, same flops as a one-sided sweep, half the sequential length), the middle 9x9
// system couples the halves, and the back-substitution runs outward from it.
//
// Mapping: THREE CHAINS PER WARP, 10 lanes per chain; a lane holds one column of EACH of [S | b], [U] and (segment
// mode) [Z], 9 registers per column, so control flow is uniform inside a warp.  Block Gauss-Jordan with 3x3 pivots
// and no pivoting (the symmetric part of every pivot block is positive definite); the pivot columns
// are broadcast through shared memory.  On the Monte-Carlo path the columns are assembled on the fly from the
// per-frame assembly records (fused system build).
#include <cuda_pipeline.h>

#include "common.cuh"
#include "launch.h"

namespace vs {

#define VS_TRY(expr)            \
  do {                          \
    int _rc = (expr);           \
    if (_rc != VINSAT_OK) return _rc; \
  } while (0)

constexpr int kMS = 10;   // padded row stride of 9x9 blocks in shared memory (16-byte aligned rows)

__device__ __forceinline__ double fast_rcp_c(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}

// out[r] = sum_k M[k*kMS + r] * v[k]   (M in shared memory, all lanes read the same addresses)
__device__ __forceinline__ void matvec9(const double* __restrict__ M, const double* v, double* out) {
#pragma unroll
  for (int r = 0; r < 9; r++) out[r] = 0.0;
#pragma unroll
  for (int k = 0; k < 9; k++) {
    const double vk = v[k];
    const double2* row = reinterpret_cast<const double2*>(M + k * kMS);
#pragma unroll
    for (int r2 = 0; r2 < 4; r2++) {
      const double2 m = row[r2];
      out[2 * r2] = fma(m.x, vk, out[2 * r2]);
      out[2 * r2 + 1] = fma(m.y, vk, out[2 * r2 + 1]);
    }
    out[8] = fma(M[k * kMS + 8], vk, out[8]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Fused system build: the normal-equation blocks of frame f are assembled on the fly, one column per lane, from
// the per-frame records the assembly kernels leave behind (BA_filtering.py:28-48) -- no [D | U | b]
// record is written to / read back from HBM (1376 B + 1368 B per frame):
//   grec[f]  sum_k w J^T J (upper triangle 21) | sum_k w J^T r (6) | sum |r|          (k_obs_assemble)
//   drec[f]  Phi 36 | r 6 | rho | qgrad 3 | Hq_diag 9 | Hq_off 9                      (k_dynamics_stm, k_quat_terms)
//   mrec[f]  Phi^T D^2 Phi 36 | Phi^T D r 6                                           (k_dynamics_stm)
// ---------------------------------------------------------------------------------------------------------
struct FusedSrc {
  const double* grec = nullptr;
  const double* drec = nullptr;
  const double* mrec = nullptr;
  const int32_t* gap = nullptr;
  const unsigned long long* wmax = nullptr;   // [P] bit pattern of the largest raw weight
  const double* zeros = nullptr;              // >= 64 zero doubles: what a lane reads where it has no term
  double Sigma = 0.0, vc = 100.0;
};

__device__ __forceinline__ int sym6(int a, int b) { return a * 6 - (a * (a - 1)) / 2 + (b - a); }   // a <= b

// column l of D_f without damping (l < 9) or the right-hand side b_f (l == 9)
__device__ __forceinline__ void fused_col0(const FusedSrc& S, const int64_t f, const int l, const double invw,
                                           double (&out)[9]) {
  const double* G = S.grec + f * VS_GREC;
  const double* Dr = S.drec + f * VS_DREC;
  const double* Mm = S.mrec + f * VS_MREC;
  const bool isb = l == 9, pos = l < 3, rot = l >= 3 && l < 6, pv = (l < 9) && !rot;
  const int pl = pos ? l : l - 3;
  const bool hn = S.gap[f] > 0;
  const bool hp = f > 0 && S.gap[f - 1] > 0;
  const double vc2 = S.vc * S.vc;
  const double sN = hn ? (isb ? -S.Sigma : S.Sigma) : 0.0;
  double g[6], m[6], h[3], e[6];
#pragma unroll
  for (int j = 0; j < 6; j++) {
    const int gidx = isb ? 21 + j : (j <= l ? sym6(j, l) : sym6(l, j));
    g[j] = (l < 6 || isb) ? invw * G[gidx] : 0.0;
    m[j] = (pv || isb) ? sN * Mm[(isb ? 6 : pl) * 6 + j] : 0.0;
    e[j] = 0.0;
  }
#pragma unroll
  for (int j = 0; j < 3; j++) h[j] = rot ? S.Sigma * Dr[46 + j * 3 + (l - 3)] : (isb ? -S.Sigma * Dr[43 + j] : 0.0);
  if (isb && hp) {
    const double* rp = S.drec + (f - 1) * VS_DREC + 36;
#pragma unroll
    for (int j = 0; j < 6; j++) e[j] = S.Sigma * (j < 3 ? 1.0 : S.vc) * rp[j];
  }
#pragma unroll
  for (int j = 0; j < 3; j++) {
    out[j] = g[j] + m[j] + e[j];
    out[3 + j] = g[3 + j] + h[j];
    out[6 + j] = m[3 + j] + e[3 + j];
  }
  if (pv && hp) {
    const double dd = S.Sigma * (pos ? 1.0 : vc2);
#pragma unroll
    for (int r = 0; r < 9; r++) out[r] += (r == l) ? dd : 0.0;
  }
}

// The sweep kernel splits the build in two so that no arithmetic waits on a load inside the latency chain:
// fused_load() only ISSUES the (predicated, lane-specific) loads of element i+1 before the elimination of element
// i, fused_combine() turns the raw values into the two columns one iteration later, in registers.
struct FusedRaw {
  double g[6];    // grec: column l of the observation block (l < 6) or J^T W r (b lane)
  double m[6];    // mrec: column pl of Phi^T D^2 Phi (pos / vel lanes) or Phi^T D r (b lane)
  double h[3];    // drec: column of Hq_diag (rot lanes) or qgrad (b lane)
  double x[6];    // drec: Phi entries of the coupling block (pos / vel), Hq_off column (rot), r of the pair before (b)
  int gn, gp;     // gap[f], gap[f-1] (0 at the first frame)
};

// A/B switches of the round-2 experiments on this kernel (B200, P=1024: baseline 9.59 ms per step of sweeps):
//   VS_SWEEP_SPARSE_MATVEC=1  skip the 36 structurally zero products of Lo x [W | y]                      -> 9.03 ms (default)
//   VS_SWEEP_UNPRED_LOADS=1   unpredicated record loads through a zero page (no CS2R / predicates)        -> 11.6 ms (rejected)
//   (one merged select chain for the diagonal additions instead of two: 10.9 ms, rejected and removed)
// The kernel is a single dependent-issue stream per scheduler; what ptxas makes of a change decides more than the
// instruction count does.
#ifndef VS_SWEEP_UNPRED_LOADS
#define VS_SWEEP_UNPRED_LOADS 0
#endif
#ifndef VS_SWEEP_SPARSE_MATVEC
#define VS_SWEEP_SPARSE_MATVEC 1
#endif
#if VS_SWEEP_UNPRED_LOADS
struct FusedLane {   // loop-invariant lane constants
  // Every lane reads the same NUMBER of values per frame; a lane that has no term of a kind reads zeros instead
  // (base = the zero page, stride 0), so that the loop body carries no predicates and no register zero-fills.
  const double *gb, *mb, *hb, *xb;     // bases of the four sources (record array or zero page), offsets folded in
  int gs, ms, hs_rec, xs_rec;          // record strides (0 for the zero page)
  int gi[6];                           // grec indices of the lane's column
  int hs, xs;                          // element strides inside a record
  bool isb, pos, rot, pv, xprev;       // xprev: the coupling source is the PREVIOUS frame's record
};

__device__ __forceinline__ FusedLane fused_lane(const FusedSrc& S, const int l, const int dir) {
  FusedLane L;
  L.isb = l == 9; L.pos = l < 3; L.rot = l >= 3 && l < 6; L.pv = (l < 9) && !L.rot;
  const int pl = L.pos ? l : l - 3, cl = l - 3;
  const bool pg = l < 6 || L.isb, pm = L.pv || L.isb, ph = L.rot || L.isb;
#pragma unroll
  for (int j = 0; j < 6; j++) L.gi[j] = pg ? (L.isb ? 21 + j : (j <= l ? sym6(j, l) : sym6(l, j))) : j;
  L.gb = pg ? S.grec : S.zeros;  L.gs = pg ? VS_GREC : 0;
  L.mb = pm ? S.mrec + (L.isb ? 36 : pl * 6) : S.zeros;  L.ms = pm ? VS_MREC : 0;
  L.hb = ph ? S.drec + (L.rot ? 46 + cl : 43) : S.zeros;  L.hs_rec = ph ? VS_DREC : 0;
  L.hs = ph ? (L.rot ? 3 : 1) : 1;
  // coupling block: Phi entries (pos / vel), Hq_off column (rot), r of the pair before (b)
  int xo;
  if (L.isb) { xo = 36; L.xs = 1; }
  else if (L.rot) { xo = dir > 0 ? 55 + cl : 55 + cl * 3; L.xs = dir > 0 ? 3 : 1; }
  else { xo = dir > 0 ? pl * 6 : pl; L.xs = dir > 0 ? 1 : 6; }
  L.xb = S.drec + xo;  L.xs_rec = VS_DREC;
  L.xprev = L.isb || dir < 0;
  return L;
}

__device__ __forceinline__ void fused_load(const FusedSrc& S, const FusedLane& L, const int64_t f, const int dir,
                                           FusedRaw& R) {
  const double* G = L.gb + f * L.gs;
  const double* Dr = L.hb + f * L.hs_rec;
  const double* Mm = L.mb + f * L.ms;
  // coupling block: this frame's record (forward) or the previous frame's (reverse; also r of the pair before).  At
  // the first frame there is no previous pair: frame 0's own record is read instead and multiplied by a zero
  // coefficient in fused_combine (gp == 0 there), so no predicate is needed.
  const int64_t fx = L.xprev ? (f > 0 ? f - 1 : 0) : f;
  const double* Dx = L.xb + fx * L.xs_rec;
  R.gn = S.gap[f];
  R.gp = f > 0 ? S.gap[f - 1] : 0;
#pragma unroll
  for (int j = 0; j < 6; j++) {
    R.g[j] = G[L.gi[j]];
    R.m[j] = Mm[j];
  }
  // rot lanes have three coupling values; their x[3..5] would be multiplied by zero: not loaded
#pragma unroll
  for (int j = 0; j < 3; j++) {
    R.x[j] = Dx[j * L.xs];
    R.h[j] = Dr[j * L.hs];
  }
  if (!L.rot) {
#pragma unroll
    for (int j = 3; j < 6; j++) R.x[j] = Dx[j * L.xs];
  } else {
    R.x[3] = 0.0; R.x[4] = 0.0; R.x[5] = 0.0;
  }
}

#else
struct FusedLane {   // loop-invariant lane constants
  int gi[6];
  int mo, ho, hs, xo, xs, xn;
  bool pg, pm, ph, isb, pos, rot, pv;
};

__device__ __forceinline__ FusedLane fused_lane(const FusedSrc&, const int l, const int dir) {
  FusedLane L;
  L.isb = l == 9; L.pos = l < 3; L.rot = l >= 3 && l < 6; L.pv = (l < 9) && !L.rot;
  const int pl = L.pos ? l : l - 3, cl = l - 3;
#pragma unroll
  for (int j = 0; j < 6; j++) L.gi[j] = L.isb ? 21 + j : (l < 6 ? (j <= l ? sym6(j, l) : sym6(l, j)) : 0);
  L.pg = l < 6 || L.isb;
  L.pm = L.pv || L.isb;
  L.mo = L.isb ? 36 : (L.pv ? pl * 6 : 0);
  L.ph = L.rot || L.isb;
  L.ho = L.rot ? 46 + cl : 43; L.hs = L.rot ? 3 : 1;
  if (L.isb) { L.xo = 36; L.xs = 1; L.xn = 6; }
  else if (L.rot) { L.xo = dir > 0 ? 55 + cl : 55 + cl * 3; L.xs = dir > 0 ? 3 : 1; L.xn = 3; }
  else { L.xo = dir > 0 ? pl * 6 : pl; L.xs = dir > 0 ? 1 : 6; L.xn = 6; }
  return L;
}

__device__ __forceinline__ void fused_load(const FusedSrc& S, const FusedLane& L, const int64_t f, const int dir,
                                           FusedRaw& R) {
  const double* G = S.grec + f * VS_GREC;
  const double* Dr = S.drec + f * VS_DREC;
  const double* Mm = S.mrec + f * VS_MREC;
  // coupling block: this frame's record (forward) or the previous frame's (reverse; also r of the pair before)
  const int64_t fx = (L.isb || dir < 0) ? f - 1 : f;
  const double* Dx = S.drec + fx * VS_DREC;
  R.gn = S.gap[f];
  R.gp = f > 0 ? S.gap[f - 1] : 0;
#pragma unroll
  for (int j = 0; j < 6; j++) {
    R.g[j] = L.pg ? G[L.gi[j]] : 0.0;
    R.m[j] = L.pm ? Mm[L.mo + j] : 0.0;
    R.x[j] = (fx >= 0 && j < L.xn) ? Dx[L.xo + j * L.xs] : 0.0;
  }
#pragma unroll
  for (int j = 0; j < 3; j++) R.h[j] = L.ph ? Dr[L.ho + j * L.hs] : 0.0;
}

#endif
__device__ __forceinline__ double fused_combine(const FusedSrc& S, const FusedLane& L, const FusedRaw& R, const int l,
                                                const int dir, const double invw, double (&out0)[9], double (&out1)[9]) {
  const bool hn = R.gn > 0, hp = R.gp > 0;
  const double vc2 = S.vc * S.vc;
  const double sN = hn ? (L.isb ? -S.Sigma : S.Sigma) : 0.0;
  const double ch = L.rot ? S.Sigma : -S.Sigma;                       // h is zero outside rot / b lanes
  const double ep = (L.isb && hp) ? S.Sigma : 0.0, ev = ep * S.vc;
#pragma unroll
  for (int j = 0; j < 3; j++) {
    out0[j] = invw * R.g[j] + sN * R.m[j] + ep * R.x[j];
    out0[3 + j] = invw * R.g[3 + j] + ch * R.h[j];
    out0[6 + j] = sN * R.m[3 + j] + ev * R.x[3 + j];
  }
  const double dd = (L.pv && hp) ? S.Sigma * (L.pos ? 1.0 : vc2) : 0.0;
#pragma unroll
  for (int r = 0; r < 9; r++) out0[r] += (r == l) ? dd : 0.0;
  const bool on = dir > 0 ? hn : hp;
  const double sg = (on && !L.isb) ? S.Sigma : 0.0;
  // forward: U[r][l] = -Sigma dv2[pl] Phi[pl][pr];  reverse: U_{i-1}[l][r] = -Sigma dv2[pr] Phi[pr][pl]
  const double cp = L.pv ? -sg * (dir > 0 ? (L.pos ? 1.0 : vc2) : 1.0) : 0.0;
  const double cv = L.pv ? -sg * (dir > 0 ? (L.pos ? 1.0 : vc2) : vc2) : 0.0;
  const double cr = L.rot ? sg : 0.0;
#pragma unroll
  for (int j = 0; j < 3; j++) {
    out1[j] = cp * R.x[j];
    out1[3 + j] = cr * R.x[j];
    out1[6 + j] = cv * R.x[3 + j];
  }
  return 0.0;
}

// corr = Lo x v for the BA system's lower blocks: Lo = U^T couples position / velocity rows only with position /
// velocity columns and rotation rows only with rotation columns, so 45 of the 81 products are zero.
// M[k*kMS + r] = Lo[r][k] as in matvec9.
__device__ __forceinline__ void matvec9_pv_rot(const double* __restrict__ M, const double* v, double* out) {
#pragma unroll
  for (int r = 0; r < 9; r++) out[r] = 0.0;
#pragma unroll
  for (int k = 0; k < 9; k++) {
    const double vk = v[k];
    const double* row = M + k * kMS;
    if (k >= 3 && k < 6) {
      const double2 a = *reinterpret_cast<const double2*>(row + 2);      // [2], [3]
      const double2 b = *reinterpret_cast<const double2*>(row + 4);      // [4], [5]
      out[3] = fma(a.y, vk, out[3]);
      out[4] = fma(b.x, vk, out[4]);
      out[5] = fma(b.y, vk, out[5]);
    } else {
      const double2 a = *reinterpret_cast<const double2*>(row);          // [0], [1]
      const double2 b = *reinterpret_cast<const double2*>(row + 6);      // [6], [7]
      out[0] = fma(a.x, vk, out[0]);
      out[1] = fma(a.y, vk, out[1]);
      out[2] = fma(row[2], vk, out[2]);
      out[6] = fma(b.x, vk, out[6]);
      out[7] = fma(b.y, vk, out[7]);
      out[8] = fma(row[8], vk, out[8]);
    }
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
  double* delta;             // solution rows
  const int32_t* out_index;  // PLAIN: element -> row of delta (null = identity)
  double* lam32_last;        // [P] or null
  // two-sided sweep (PLAIN only): chains 2p (top, dir +1) and 2p+1 (bottom, dir -1) of problem p meet at ch_mid
  const int32_t* ch_dir = nullptr;     // direction of travel per chain, or null (= +1); ch_b is the exclusive end sentinel
  const int32_t* ch_mid = nullptr;     // middle element per chain (or -1: empty problem), or null
  double* mid = nullptr;               // [n_chains][VS_MIDREC]  Lo W (row-major 81) | Lo y (9) of the chain's last element
  int fused = 0;                       // 1: rec is null, columns come from `fs` (fused system build)
  const int32_t* gate = nullptr;       // speculation gate: the kernel returns at once while *gate != 0
  FusedSrc fs;
};
constexpr int VS_MIDREC = 96;


// Block Gauss-Jordan on the augmented columns [S | ...] (one column per lane, 9 registers) with 3x3 pivot blocks
// (position, rotation, velocity): three dependent pivot steps per element instead of nine.  The three pivot
// columns are broadcast through shared memory, every lane inverts the 3x3 pivot block in closed form (adjugate,
// one reciprocal) and updates its own column.  Lanes 0..8 must hold the columns of S.
__device__ __forceinline__ void gj_block3(double (&a_)[9], const int c, double (*colk3)[3][kMS]) {
#pragma unroll
  for (int kb = 0; kb < 3; kb++) {
    if (c >= 3 * kb && c < 3 * kb + 3) {
      double* dst = colk3[kb & 1][c - 3 * kb];
#pragma unroll
      for (int r = 0; r < 9; r++) dst[r] = a_[r];
    }
    __syncwarp();
    double pc[3][9];
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const double2* p2 = reinterpret_cast<const double2*>(colk3[kb & 1][j]);
#pragma unroll
      for (int r2 = 0; r2 < 4; r2++) { const double2 v = p2[r2]; pc[j][2 * r2] = v.x; pc[j][2 * r2 + 1] = v.y; }
      pc[j][8] = colk3[kb & 1][j][8];
    }
    const int o = 3 * kb;
    const double p00 = pc[0][o], p10 = pc[0][o + 1], p20 = pc[0][o + 2];
    const double p01 = pc[1][o], p11 = pc[1][o + 1], p21 = pc[1][o + 2];
    const double p02 = pc[2][o], p12 = pc[2][o + 1], p22 = pc[2][o + 2];
    const double c00 = p11 * p22 - p12 * p21, c01 = p12 * p20 - p10 * p22, c02 = p10 * p21 - p11 * p20;
    const double inv = fast_rcp_c(p00 * c00 + p01 * c01 + p02 * c02);
    const double i00 = c00 * inv, i01 = (p02 * p21 - p01 * p22) * inv, i02 = (p01 * p12 - p02 * p11) * inv;
    const double i10 = c01 * inv, i11 = (p00 * p22 - p02 * p20) * inv, i12 = (p02 * p10 - p00 * p12) * inv;
    const double i20 = c02 * inv, i21 = (p01 * p20 - p00 * p21) * inv, i22 = (p00 * p11 - p01 * p10) * inv;
    const double b0 = a_[o], b1 = a_[o + 1], b2 = a_[o + 2];
    const double t0 = i00 * b0 + i01 * b1 + i02 * b2;
    const double t1 = i10 * b0 + i11 * b1 + i12 * b2;
    const double t2 = i20 * b0 + i21 * b1 + i22 * b2;
#pragma unroll
    for (int r = 0; r < 9; r++) {
      if (r == o) a_[r] = t0;
      else if (r == o + 1) a_[r] = t1;
      else if (r == o + 2) a_[r] = t2;
      else a_[r] = fma(-pc[2][r], t2, fma(-pc[1][r], t1, fma(-pc[0][r], t0, a_[r])));
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Forward elimination, THREE CHAINS PER WARP.  A one-column-per-lane version (one warp per chain) was bound by the
// L1/shared-memory data pipe (ncu: 238 shared wavefronts per frame for the pivot / Lo broadcasts, 75-95 % busy),
// not by latency, so more resident chains do not help it.  Here a group of 10 lanes owns a chain and every lane
// holds one column of EACH of [S | b], [U | -] (and [Z | -] in segment mode): a broadcast shared-memory read
// now serves three chains, the 3x3 pivot inverse is shared by the lane's 2-3 columns, 29/30 column slots do
// useful work (19/32 before), and W_l = S^-1 U_l stays in the lane that needs it for Lo W (no W round trip).
// ---------------------------------------------------------------------------------------------------------
constexpr int kCPW = 3;      // chains per warp
constexpr int kGL = 10;      // lanes per chain

template <int NS>
__device__ __forceinline__ void gj_block3_slots(double (&a)[NS][9], const int l, double (*colk3)[3][kMS]) {
#pragma unroll
  for (int kb = 0; kb < 3; kb++) {
    if (l >= 3 * kb && l < 3 * kb + 3) {
      double* dst = colk3[kb & 1][l - 3 * kb];
#pragma unroll
      for (int r = 0; r < 9; r++) dst[r] = a[0][r];
    }
    __syncwarp();
    double pc[3][9];
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const double2* p2 = reinterpret_cast<const double2*>(colk3[kb & 1][j]);
#pragma unroll
      for (int r2 = 0; r2 < 4; r2++) { const double2 v = p2[r2]; pc[j][2 * r2] = v.x; pc[j][2 * r2 + 1] = v.y; }
      pc[j][8] = colk3[kb & 1][j][8];
    }
    const int o = 3 * kb;
    const double p00 = pc[0][o], p10 = pc[0][o + 1], p20 = pc[0][o + 2];
    const double p01 = pc[1][o], p11 = pc[1][o + 1], p21 = pc[1][o + 2];
    const double p02 = pc[2][o], p12 = pc[2][o + 1], p22 = pc[2][o + 2];
    const double c00 = p11 * p22 - p12 * p21, c01 = p12 * p20 - p10 * p22, c02 = p10 * p21 - p11 * p20;
    const double inv = fast_rcp_c(p00 * c00 + p01 * c01 + p02 * c02);
    const double i00 = c00 * inv, i01 = (p02 * p21 - p01 * p22) * inv, i02 = (p01 * p12 - p02 * p11) * inv;
    const double i10 = c01 * inv, i11 = (p00 * p22 - p02 * p20) * inv, i12 = (p02 * p10 - p00 * p12) * inv;
    const double i20 = c02 * inv, i21 = (p01 * p20 - p00 * p21) * inv, i22 = (p00 * p11 - p01 * p10) * inv;
#pragma unroll
    for (int sl = 0; sl < NS; sl++) {
      const double b0 = a[sl][o], b1 = a[sl][o + 1], b2 = a[sl][o + 2];
      const double t0 = i00 * b0 + i01 * b1 + i02 * b2;
      const double t1 = i10 * b0 + i11 * b1 + i12 * b2;
      const double t2 = i20 * b0 + i21 * b1 + i22 * b2;
#pragma unroll
      for (int r = 0; r < 9; r++) {
        if (r == o) a[sl][r] = t0;
        else if (r == o + 1) a[sl][r] = t1;
        else if (r == o + 2) a[sl][r] = t2;
        else a[sl][r] = fma(-pc[2][r], t2, fma(-pc[1][r], t1, fma(-pc[0][r], t0, a[sl][r])));
      }
    }
  }
}

constexpr int kF3Warps = 1;

// MINB = resident CTAs per SM the register allocation must allow: 8 lets ptxas take 244 registers, 12 caps it at 168
// (68 bytes of spills in the fused variant) so that three warps share a scheduler when there are enough chains.
template <bool SPIKE, bool FUSED, int MINB = 8>
__global__ void __launch_bounds__(kF3Warps * 32, MINB) k_chain_forward3(ChainArgs A) {
  constexpr int NS = SPIKE ? 3 : 2;
  if (A.gate && *A.gate) return;
  __shared__ __align__(16) double s_col[kF3Warps][kCPW][2][3][kMS];
  __shared__ __align__(16) double s_M[kF3Warps][kCPW][9 * kMS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = min(lane / kGL, kCPW - 1);
  const bool spare = lane >= kCPW * kGL;              // lanes 30, 31 shadow lane 29 and never write
  const int l = spare ? kGL - 1 : lane - g * kGL;
  const int ch = (blockIdx.x * kF3Warps + warp) * kCPW + g;
  bool valid = ch < A.n_chains;
  int prob = 0, a = 0, e = 0, dir = 1, left = -1;
  if (valid) {
    prob = A.ch_prob[ch];
    if (A.active && !A.active[prob]) valid = false;
  }
  if (valid) {
    a = A.ch_a[ch]; e = A.ch_b[ch];
    dir = (!SPIKE && A.ch_dir) ? A.ch_dir[ch] : 1;
    left = SPIKE ? A.ch_left[ch] : -1;
  }
  const bool writer = valid && !spare;
  double lam32 = 0.0;
  if (valid && A.lam) {
    lam32 = (double)(float)A.lam[prob];           // torch.eye(n)*lamda is float32
    if (A.lam32_last && writer && l == 0) A.lam32_last[prob] = lam32;
  }
  const bool isS = l < 9;                         // slot 0: S column l (l < 9) or b (l == 9); slot 1: U column l; slot 2: Z column l
  const bool rev = dir < 0;
  // slot 0 / slot 1 offsets inside a system record.  A bottom chain (dir -1) couples element i to i-1 through
  // A(i, i-1) = U_{i-1}^T: its U column is a (contiguous) row of the record of element i-1.
  const int base0 = isS ? l : 162, rs0 = isS ? 9 : 1;
  const int base1 = rev ? 81 + l * 9 : 81 + l, rs1 = rev ? 1 : 9, joff1 = rev ? -1 : 0;
  double (*colk3)[3][kMS] = s_col[warp][g];
  double* M = s_M[warp][g];
  const int len = valid ? (e - a) * dir : 0;
  double* rr = (SPIKE && valid) ? A.redrec + (int64_t)ch * VS_RREC : nullptr;
  double* midrec = (!SPIKE && A.mid && valid) ? A.mid + (int64_t)ch * VS_MIDREC : nullptr;

  if (writer && len <= 0) {
    if (midrec) {
      for (int idx = l; idx < 90; idx += kGL) midrec[idx] = 0.0;
    }
    if (SPIKE && len == 0) {
      // no interior: the separator couples directly to the left separator.  Ll = Lo_left, Dl = bl = 0.
      for (int idx = l; idx < 171; idx += kGL) {
        double v = 0.0;
        if (idx >= 81 && idx < 162 && left >= 0) {
          const int r = (idx - 81) / 9, k = (idx - 81) % 9;
          v = A.lrec ? A.lrec[(int64_t)left * 81 + r * 9 + k] : A.rec[(int64_t)left * VS_SREC + 81 + k * 9 + r];
        }
        rr[idx] = v;
      }
    }
  }
  int maxlen = max(len, 0);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
  if (maxlen == 0) return;

  double av[NS][9], nxt0[9], nxt1[9], corr0[9], corr2[9];
#pragma unroll
  for (int r = 0; r < 9; r++) {
    corr0[r] = 0.0; corr2[r] = 0.0; nxt0[r] = 0.0; nxt1[r] = 0.0;
#pragma unroll
    for (int sl = 0; sl < NS; sl++) av[sl][r] = (sl == 0 && r == l) ? 1.0 : 0.0;    // idle groups: harmless identity pivots
  }
  double invw = 0.0;
  if (FUSED && valid) {
    const unsigned long long wb = A.fs.wmax[prob];
    invw = wb ? 1.0 / __longlong_as_double((long long)wb) : 0.0;
  }
  FusedLane FL;
  FusedRaw FR;
  if (FUSED) FL = fused_lane(A.fs, l, dir);
  if (len > 0) {
    if (FUSED) {
      fused_load(A.fs, FL, a, dir, FR);
    } else {
      const double* rec0 = A.rec + (int64_t)a * VS_SREC;
      const double* rec1 = A.rec + (int64_t)(a + joff1) * VS_SREC;
#pragma unroll
      for (int r = 0; r < 9; r++) {
        nxt0[r] = rec0[base0 + r * rs0];
        nxt1[r] = isS ? rec1[base1 + r * rs1] : 0.0;
      }
    }
    if (SPIKE && isS && left >= 0) {
      // Z~_a = Lo_left, column l; kept NEGATED in corr2 because the assembly below uses a2 = -corr2
      if (A.lrec) {
#pragma unroll
        for (int r = 0; r < 9; r++) corr2[r] = -A.lrec[(int64_t)left * 81 + r * 9 + l];
      } else {
#pragma unroll
        for (int r = 0; r < 9; r++) corr2[r] = -A.rec[(int64_t)left * VS_SREC + 81 + l * 9 + r];
      }
    }
  }

  for (int s = 0; s < maxlen; s++) {
    const bool on = s < len;
    const bool more = s + 1 < len;
    const int i = a + dir * s;
    if (FUSED && on) fused_combine(A.fs, FL, FR, l, dir, invw, nxt0, nxt1);
    if (on) {
      // assemble the columns of element i:  S: D + lam I - Lo W;  b: b - Lo y;  U: fresh;  Z: -Lo Z
#pragma unroll
      for (int r = 0; r < 9; r++) {
        av[0][r] = nxt0[r] + ((isS && r == l) ? lam32 : 0.0) - corr0[r];
        av[1][r] = nxt1[r];
        if (SPIKE) av[NS - 1][r] = -corr2[r];
      }
    }
    if (more) {
      if (FUSED) {
        fused_load(A.fs, FL, i + dir, dir, FR);
      } else {
        const double* rec0 = A.rec + (int64_t)(i + dir) * VS_SREC;
        const double* rec1 = A.rec + (int64_t)(i + dir + joff1) * VS_SREC;
#pragma unroll
        for (int r = 0; r < 9; r++) {
          nxt0[r] = rec0[base0 + r * rs0];
          if (isS) nxt1[r] = rec1[base1 + r * rs1];
        }
      }
    }
    // lower block Lo_i staged as M[k*kMS + r] = Lo_i[r][k]
    if (on) {
      if (A.lrec) {
        if (!spare) {
          const double* L = A.lrec + (int64_t)i * 81;
          for (int idx = l; idx < 81; idx += kGL) { const int r = idx / 9, k = idx % 9; M[k * kMS + r] = L[idx]; }
        }
      } else if (isS) {
#pragma unroll
        for (int r = 0; r < 9; r++) M[r * kMS + l] = av[1][r];       // M[k][r'] = U[k][r'] = Lo[r'][k]
      }
    }
    gj_block3_slots<NS>(av, l, colk3);
    if (on && writer) {
      double* w = A.wrec + (int64_t)i * VS_WREC;
      if (isS) {
#pragma unroll
        for (int r = 0; r < 9; r++) w[l * 9 + r] = av[1][r];
        if (SPIKE) {
#pragma unroll
          for (int r = 0; r < 9; r++) w[90 + l * 9 + r] = av[NS - 1][r];
        }
      } else {
#pragma unroll
        for (int r = 0; r < 9; r++) w[81 + r] = av[0][r];
      }
    }
    __syncwarp();      // M complete (and the pivot buffers free) before the products below
    // corr = Lo_i x (W_l | y | Z_l): for element i+1 and, after the last element, for the middle / separator row
    if (on && (more || SPIKE || midrec)) {
      double v[9];
#pragma unroll
      for (int k = 0; k < 9; k++) v[k] = isS ? av[1][k] : av[0][k];
      if (FUSED && VS_SWEEP_SPARSE_MATVEC) matvec9_pv_rot(M, v, corr0);
      else matvec9(M, v, corr0);
      if (SPIKE) matvec9(M, av[NS - 1], corr2);
      if (!more && writer) {
        if (midrec) {
          // contribution of this half to the middle element's row: Lo W (9x9) and Lo y
#pragma unroll
          for (int r = 0; r < 9; r++) {
            if (isS) midrec[r * 9 + l] = corr0[r];
            else midrec[81 + r] = corr0[r];
          }
        }
        if (SPIKE) {
          // left part of the separator's row: Dl = -Lo W, bl = -Lo y, Ll = -Lo Z (from the last interior element)
#pragma unroll
          for (int r = 0; r < 9; r++) {
            if (isS) { rr[r * 9 + l] = -corr0[r]; rr[81 + r * 9 + l] = -corr2[r]; }
            else rr[162 + r] = -corr0[r];
          }
        }
      }
    }
    __syncwarp();
  }
}

template <bool SPIKE>
static int launch_forward(vinsat_ctx* ctx, const ChainArgs& A) {
  if (A.fused) {
    static const int minb = getenv("VINSAT_SWEEP_MINB") ? atoi(getenv("VINSAT_SWEEP_MINB")) : 8;
    if (minb >= 16)
      VS_LAUNCH(ctx, F_SOLVE, (k_chain_forward3<false, true, 16>), ceil_div(A.n_chains, kF3Warps * kCPW), kF3Warps * 32, 0, A);
    else if (minb >= 12)
      VS_LAUNCH(ctx, F_SOLVE, (k_chain_forward3<false, true, 12>), ceil_div(A.n_chains, kF3Warps * kCPW), kF3Warps * 32, 0, A);
    else
      VS_LAUNCH(ctx, F_SOLVE, (k_chain_forward3<false, true, 8>), ceil_div(A.n_chains, kF3Warps * kCPW), kF3Warps * 32, 0, A);
  } else {
    VS_LAUNCH(ctx, F_SOLVE, (k_chain_forward3<SPIKE, false>), ceil_div(A.n_chains, kF3Warps * kCPW), kF3Warps * 32, 0, A);
  }
  return VINSAT_OK;
}

// ---------------------------------------------------------------------------------------------------------
// plain chains: backward substitution x_i = y_i - W_i x_{i+1}; lane r < 9 owns row r.  The per-frame work is
// ~150 cycles but a W record comes from HBM (~800 cycles), so records are streamed through a shared-memory ring
// with cp.async, kRing frames ahead.
// ---------------------------------------------------------------------------------------------------------
constexpr int kRing = 8;

__global__ void __launch_bounds__(128) k_chain_backward(ChainArgs A) {
  if (A.gate && *A.gate) return;
  __shared__ __align__(16) double s_ring[4][kRing][96];
  __shared__ __align__(16) double s_col[4][2][3][kMS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ch = blockIdx.x * 4 + warp;
  if (ch >= A.n_chains) return;
  const int prob = A.ch_prob[ch];
  if (A.active && !A.active[prob]) return;
  const int a = A.ch_a[ch], e = A.ch_b[ch];
  const int dir = A.ch_dir ? A.ch_dir[ch] : 1;
  const int len = (e - a) * dir;
  const int mid = A.ch_mid ? A.ch_mid[ch] : -1;
  if (len <= 0 && mid < 0) return;
  double (*ring)[96] = s_ring[warp];
  // elements are visited from the chain's last one (next to the sentinel e) back to its first: i_s = e - dir - dir*s
  auto issue = [&](int s, int slot) {          // W (81) | y (9) of the s-th visited element -> ring[slot][0..90)
    if (s < len) {
      const double* w = A.wrec + (int64_t)(e - dir - dir * s) * VS_WREC;
      for (int idx = lane; idx < 90; idx += 32) __pipeline_memcpy_async(&ring[slot][idx], w + idx, 8);
    }
    __pipeline_commit();
  };
#pragma unroll
  for (int d = 0; d < kRing; d++) issue(d, d);
  double dn[9];
#pragma unroll
  for (int k = 0; k < 9; k++) dn[k] = 0.0;
  bool have_dn = false;
  if (mid >= 0) {
    // two-sided sweep: both halves solve the middle element's 9x9 system (identical arithmetic in both warps)
    //   (D_m + lam I - Lo W|top - Lo W|bottom) x_m = b_m - Lo y|top - Lo y|bottom
    const double* mt = A.mid + (int64_t)(ch & ~1) * VS_MIDREC;
    const double* mb = A.mid + (int64_t)(ch | 1) * VS_MIDREC;
    const double lam32 = A.lam ? (double)(float)A.lam[prob] : 0.0;
    double col[9], own[9];
#pragma unroll
    for (int r = 0; r < 9; r++) own[r] = 0.0;
    if (A.fused) {
      const unsigned long long wb = A.fs.wmax[prob];
      const double invw = wb ? 1.0 / __longlong_as_double((long long)wb) : 0.0;
      if (lane < 10) fused_col0(A.fs, mid, lane, invw, own);
    } else {
      const double* rec = A.rec + (int64_t)mid * VS_SREC;
#pragma unroll
      for (int r = 0; r < 9; r++) own[r] = lane < 9 ? rec[r * 9 + lane] : (lane == 9 ? rec[162 + r] : 0.0);
    }
#pragma unroll
    for (int r = 0; r < 9; r++) {
      double v = 0.0;
      if (lane < 9) v = own[r] + (r == lane ? lam32 : 0.0) - mt[r * 9 + lane] - mb[r * 9 + lane];
      else if (lane == 9) v = own[r] - mt[81 + r] - mb[81 + r];
      else if (lane < 18) v = (r == lane - 9) ? 1.0 : 0.0;     // idle lanes carry harmless finite columns
      col[r] = v;
    }
    gj_block3(col, lane, s_col[warp]);
#pragma unroll
    for (int k = 0; k < 9; k++) dn[k] = __shfl_sync(0xffffffffu, col[k], 9);
    have_dn = true;
    if ((ch & 1) == 0 && lane < 9) {
      const int64_t row = A.out_index ? A.out_index[mid] : mid;
      double xm = 0.0;
#pragma unroll
      for (int k = 0; k < 9; k++) xm = (lane == k) ? dn[k] : xm;
      A.delta[row * 9 + lane] = xm;
    }
  }
  int slot = 0;
  for (int s = 0; s < len; s++) {
    const int i = e - dir - dir * s;
    __pipeline_wait_prior(kRing - 1);
    __syncwarp();
    double dr = 0.0;
    if (lane < 9) {
      const double* w = ring[slot];
      dr = w[81 + lane];
      if (have_dn) {
#pragma unroll
        for (int k = 0; k < 9; k++) dr = fma(-w[k * 9 + lane], dn[k], dr);
      }
    }
    have_dn = true;
    __syncwarp();
    issue(s + kRing, slot);
#pragma unroll
    for (int k = 0; k < 9; k++) dn[k] = __shfl_sync(0xffffffffu, dr, k);
    const int64_t row = A.out_index ? A.out_index[i] : i;
    if (lane < 9) A.delta[row * 9 + lane] = dr;
    slot = (slot + 1 == kRing) ? 0 : slot + 1;
  }
}

// ---------------------------------------------------------------------------------------------------------
// segment mode, second pass: closed form x_a = yh - Wh x_b - Zh x_left by the backward recurrence
//   Wh_i = -W_i Wh_{i+1},  Zh_i = Z_i - W_i Zh_{i+1},  yh_i = y_i - W_i yh_{i+1}   (started at the last interior)
// and the RIGHT part of the left separator's row: Dr = -U_left Zh_a, Ur = -U_left Wh_a, br = -U_left yh_a.
// Lanes 0-8 own Wh columns, lane 9 owns yh, lanes 10-18 own Zh columns.
// ---------------------------------------------------------------------------------------------------------
// The records are streamed through a shared-memory ring with cp.async, kBrRing elements ahead (a step is ~600 cycles of
// work, an HBM round trip ~1000: loading each record when it is needed made this pass slower than the elimination itself,
// 2.9 ms against 2.1 ms on a 2.4 M-frame arc).
constexpr int kBrRing = 4;
constexpr int kBrRow = 184;     // W re-strided to 9 x kMS (90) | y (9) at 90 | Z col-major (81) at 99 | pad

__global__ void __launch_bounds__(128) k_seg_backrec(ChainArgs A) {
  __shared__ __align__(16) double s_ring[4][kBrRing][kBrRow];
  __shared__ __align__(16) double s_M[4][9 * kMS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ch = blockIdx.x * 4 + warp;
  if (ch >= A.n_chains) return;
  if (A.active && !A.active[A.ch_prob[ch]]) return;
  const int a = A.ch_a[ch], e = A.ch_b[ch], left = A.ch_left[ch];
  double* rq = A.redrec + (int64_t)ch * VS_RREC + 171;
  if (left < 0) {
    for (int idx = lane; idx < 171; idx += 32) rq[idx] = 0.0;
    return;
  }
  const double* Ul = A.rec + (int64_t)left * VS_SREC + 81;
  if (e <= a) {      // no interior: Ur = U_left, Dr = br = 0
    for (int idx = lane; idx < 171; idx += 32) rq[idx] = (idx >= 81 && idx < 162) ? Ul[idx - 81] : 0.0;
    return;
  }
  double (*ring)[kBrRow] = s_ring[warp];
  auto issue = [&](int i, int slot) {          // wrec of interior element i -> ring[slot]
    if (i >= a) {
      const double* w = A.wrec + (int64_t)i * VS_WREC;
      double* dst = ring[slot];
      for (int idx = lane; idx < 171; idx += 32) {
        const int d = idx < 81 ? (idx / 9) * kMS + (idx % 9) : idx + 9;
        __pipeline_memcpy_async(dst + d, w + idx, 8);
      }
    }
    __pipeline_commit();
  };
#pragma unroll
  for (int d = 0; d < kBrRing; d++) issue(e - 2 - d, d);
  double* M = s_M[warp];
  const bool isW = lane < 9, isY = lane == 9, isZ = lane >= 10 && lane < 19;
  const int cc = isW ? lane : (isZ ? lane - 10 : 0);
  // own column of element i inside a wrec record: W col cc | y | Z col cc
  const int off = isW ? cc * 9 : (isY ? 81 : 90 + cc * 9);
  const int roff = isY ? 90 : 99 + cc * 9;             // the same column inside a ring slot (y / Z only)
  const bool act = lane < 19;
  double h[9], own[9];
  {
    const double* w = A.wrec + (int64_t)(e - 1) * VS_WREC;
#pragma unroll
    for (int r = 0; r < 9; r++) h[r] = act ? w[off + r] : 0.0;
  }
  int slot = 0;
  for (int i = e - 2; i >= a; i--) {
    __pipeline_wait_prior(kBrRing - 1);
    __syncwarp();
    const double* R = ring[slot];              // R[c*kMS + r] = W_i[r][c]
#pragma unroll
    for (int r = 0; r < 9; r++) own[r] = (act && !isW) ? R[roff + r] : 0.0;
    double o[9];
    matvec9(R, h, o);
#pragma unroll
    for (int r = 0; r < 9; r++) h[r] = own[r] - o[r];
    __syncwarp();
    issue(i - kBrRing, slot);
    slot = (slot + 1 == kBrRing) ? 0 : slot + 1;
  }
  // M[k*kMS + r] = U_left[r][k]
  for (int idx = lane; idx < 81; idx += 32) { const int r = idx / 9, k = idx % 9; M[k * kMS + r] = Ul[idx]; }
  __syncwarp();
  double o[9];
  matvec9(M, h, o);
#pragma unroll
  for (int r = 0; r < 9; r++) {
    if (isZ) rq[r * 9 + cc] = -o[r];            // Dr
    else if (isW) rq[81 + r * 9 + cc] = -o[r];  // Ur
    else if (isY) rq[162 + r] = -o[r];          // br
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
    if (lam && e / 9 == e % 9) v += (double)(float)lam[prob];        // lam == null: level 2 (already damped at level 1)
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
// interior back-substitution of a segment: x_i = y_i - W_i x_{i+1} - Z_i x_left, i = b-1 .. a.
// Lane r < 9 owns row r; records are streamed through a cp.async ring like k_chain_backward.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_seg_backsub(int n_chains, const int32_t* __restrict__ ch_a,
                                                     const int32_t* __restrict__ ch_b,
                                                     const int32_t* __restrict__ ch_left,
                                                     const int32_t* __restrict__ ch_prob,
                                                     const int32_t* __restrict__ active,
                                                     const double* __restrict__ wrec, double* __restrict__ delta) {
  constexpr int kRing = 4;       // (shadows the 8 of k_chain_backward) 22.5 KB per CTA: every CTA of a 3552-segment arc resident
  __shared__ __align__(16) double s_ring[4][kRing][176];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ch = blockIdx.x * 4 + warp;
  if (ch >= n_chains) return;
  if (active && !active[ch_prob[ch]]) return;
  const int a = ch_a[ch], b = ch_b[ch], left = ch_left[ch];
  if (b <= a) return;
  double (*ring)[176] = s_ring[warp];
  auto issue = [&](int i, int slot) {
    if (i >= a) {
      const double* w = wrec + (int64_t)i * VS_WREC;
      for (int idx = lane; idx < 171; idx += 32) __pipeline_memcpy_async(&ring[slot][idx], w + idx, 8);
    }
    __pipeline_commit();
  };
#pragma unroll
  for (int d = 0; d < kRing; d++) issue(b - 1 - d, d);
  double x[9], xl[9];
#pragma unroll
  for (int r = 0; r < 9; r++) {
    x[r] = delta[(int64_t)b * 9 + r];
    xl[r] = left >= 0 ? delta[(int64_t)left * 9 + r] : 0.0;
  }
  int slot = 0;
  for (int i = b - 1; i >= a; i--) {
    __pipeline_wait_prior(kRing - 1);
    __syncwarp();
    double dr = 0.0;
    if (lane < 9) {
      const double* w = ring[slot];
      double zr = 0.0;
#pragma unroll
      for (int k = 0; k < 9; k++) zr = fma(w[90 + k * 9 + lane], xl[k], zr);     // off the critical path
      dr = w[81 + lane] - zr;
#pragma unroll
      for (int k = 0; k < 9; k++) dr = fma(-w[k * 9 + lane], x[k], dr);
    }
    __syncwarp();
    issue(i - kRing, slot);
#pragma unroll
    for (int k = 0; k < 9; k++) x[k] = __shfl_sync(0xffffffffu, dr, k);
    if (lane < 9) delta[(int64_t)i * 9 + lane] = dr;
    slot = (slot + 1 == kRing) ? 0 : slot + 1;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Bulk-copy rings.  The two backward passes of the partitioned solve stream one contiguous 1376-byte W|y|Z record per
// step through a per-warp shared-memory ring.  Fed with 8-byte cp.async that is 171 LDGSTS per record = 6 per lane per
// step at 8 LSU cycles each -- with 24-32 warps per SM the load/store unit, not the arithmetic, set the step time
// (~2000 cycles).  Here ONE lane issues ONE cp.async.bulk per record (the copy engine moves it, completion is counted
// in bytes on the slot's mbarrier) and every lane waits on the barrier's phase.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}

constexpr int kBulkRing = 4;
constexpr uint32_t kWrecBytes = VS_WREC * sizeof(double);      // 1376: a multiple of 16, records are 16-byte aligned
static_assert(kWrecBytes % 16 == 0, "cp.async.bulk moves multiples of 16 bytes");

// out[r] = sum_k M[k*9 + r] * v[k] for a 9x9 block stored with stride 9 (as it lies in a wrec record).  Rows start
// 16-byte aligned for even k and 8 bytes past that for odd k; the 16-byte loads are placed accordingly.
__device__ __forceinline__ void matvec9_s9(const double* __restrict__ M, const double* v, double* out) {
#pragma unroll
  for (int r = 0; r < 9; r++) out[r] = 0.0;
#pragma unroll
  for (int k = 0; k < 9; k++) {
    const double vk = v[k];
    const double* row = M + k * 9;
    if ((k & 1) == 0) {
#pragma unroll
      for (int r2 = 0; r2 < 4; r2++) {
        const double2 m = *reinterpret_cast<const double2*>(row + 2 * r2);
        out[2 * r2] = fma(m.x, vk, out[2 * r2]);
        out[2 * r2 + 1] = fma(m.y, vk, out[2 * r2 + 1]);
      }
      out[8] = fma(row[8], vk, out[8]);
    } else {
      out[0] = fma(row[0], vk, out[0]);
#pragma unroll
      for (int r2 = 0; r2 < 4; r2++) {
        const double2 m = *reinterpret_cast<const double2*>(row + 1 + 2 * r2);
        out[1 + 2 * r2] = fma(m.x, vk, out[1 + 2 * r2]);
        out[2 + 2 * r2] = fma(m.y, vk, out[2 + 2 * r2]);
      }
    }
  }
}

// k_seg_backrec with the bulk ring (same arithmetic, same order).
__global__ void __launch_bounds__(128) k_seg_backrec_bulk(ChainArgs A) {
  __shared__ __align__(128) double s_ring[4][kBulkRing][VS_WREC];
  __shared__ __align__(16) double s_M[4][9 * kMS];
  __shared__ __align__(8) unsigned long long s_bar[4][kBulkRing];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ch = blockIdx.x * 4 + warp;
  if (ch >= A.n_chains) return;
  if (A.active && !A.active[A.ch_prob[ch]]) return;
  const int a = A.ch_a[ch], e = A.ch_b[ch], left = A.ch_left[ch];
  double* rq = A.redrec + (int64_t)ch * VS_RREC + 171;
  if (left < 0) {
    for (int idx = lane; idx < 171; idx += 32) rq[idx] = 0.0;
    return;
  }
  const double* Ul = A.rec + (int64_t)left * VS_SREC + 81;
  if (e <= a) {      // no interior: Ur = U_left, Dr = br = 0
    for (int idx = lane; idx < 171; idx += 32) rq[idx] = (idx >= 81 && idx < 162) ? Ul[idx - 81] : 0.0;
    return;
  }
  double (*ring)[VS_WREC] = s_ring[warp];
  const uint32_t bar0 = smem_u32(&s_bar[warp][0]), ring0 = smem_u32(&ring[0][0]);
  if (lane == 0) {
#pragma unroll
    for (int d = 0; d < kBulkRing; d++) mbar_init(bar0 + 8 * d, 1);
    mbar_fence_init();
  }
  __syncwarp();
  auto issue = [&](int i, int slot) {          // wrec of interior element i -> ring[slot], one bulk copy
    if (lane == 0 && i >= a) {
      mbar_expect_tx(bar0 + 8 * slot, kWrecBytes);
      bulk_g2s(ring0 + slot * kWrecBytes, A.wrec + (int64_t)i * VS_WREC, kWrecBytes, bar0 + 8 * slot);
    }
  };
#pragma unroll
  for (int d = 0; d < kBulkRing; d++) issue(e - 2 - d, d);
  double* M = s_M[warp];
  const bool isW = lane < 9, isY = lane == 9, isZ = lane >= 10 && lane < 19;
  const int cc = isW ? lane : (isZ ? lane - 10 : 0);
  const int off = isW ? cc * 9 : (isY ? 81 : 90 + cc * 9);      // own column inside a wrec record: W col cc | y | Z col cc
  const bool act = lane < 19;
  double h[9], own[9];
  {
    const double* w = A.wrec + (int64_t)(e - 1) * VS_WREC;
#pragma unroll
    for (int r = 0; r < 9; r++) h[r] = act ? w[off + r] : 0.0;
  }
  uint32_t phases = 0;
  int slot = 0;
  for (int i = e - 2; i >= a; i--) {
    mbar_wait(bar0 + 8 * slot, (phases >> slot) & 1u);
    phases ^= 1u << slot;
    const double* R = ring[slot];              // R[c*9 + r] = W_i[r][c]
#pragma unroll
    for (int r = 0; r < 9; r++) own[r] = (act && !isW) ? R[off + r] : 0.0;
    double o[9];
    matvec9_s9(R, h, o);
#pragma unroll
    for (int r = 0; r < 9; r++) h[r] = own[r] - o[r];
    __syncwarp();                              // every lane has read the slot before it is refilled
    issue(i - kBulkRing, slot);
    slot = (slot + 1 == kBulkRing) ? 0 : slot + 1;
  }
  // M[k*kMS + r] = U_left[r][k]
  for (int idx = lane; idx < 81; idx += 32) { const int r = idx / 9, k = idx % 9; M[k * kMS + r] = Ul[idx]; }
  __syncwarp();
  double o[9];
  matvec9(M, h, o);
#pragma unroll
  for (int r = 0; r < 9; r++) {
    if (isZ) rq[r * 9 + cc] = -o[r];            // Dr
    else if (isW) rq[81 + r * 9 + cc] = -o[r];  // Ur
    else if (isY) rq[162 + r] = -o[r];          // br
  }
}

// k_seg_backsub with the bulk ring (same arithmetic, same order).
__global__ void __launch_bounds__(128) k_seg_backsub_bulk(int n_chains, const int32_t* __restrict__ ch_a,
                                                          const int32_t* __restrict__ ch_b,
                                                          const int32_t* __restrict__ ch_left,
                                                          const int32_t* __restrict__ ch_prob,
                                                          const int32_t* __restrict__ active,
                                                          const double* __restrict__ wrec, double* __restrict__ delta) {
  __shared__ __align__(128) double s_ring[4][kBulkRing][VS_WREC];
  __shared__ __align__(8) unsigned long long s_bar[4][kBulkRing];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ch = blockIdx.x * 4 + warp;
  if (ch >= n_chains) return;
  if (active && !active[ch_prob[ch]]) return;
  const int a = ch_a[ch], b = ch_b[ch], left = ch_left[ch];
  if (b <= a) return;
  double (*ring)[VS_WREC] = s_ring[warp];
  const uint32_t bar0 = smem_u32(&s_bar[warp][0]), ring0 = smem_u32(&ring[0][0]);
  if (lane == 0) {
#pragma unroll
    for (int d = 0; d < kBulkRing; d++) mbar_init(bar0 + 8 * d, 1);
    mbar_fence_init();
  }
  __syncwarp();
  auto issue = [&](int i, int slot) {
    if (lane == 0 && i >= a) {
      mbar_expect_tx(bar0 + 8 * slot, kWrecBytes);
      bulk_g2s(ring0 + slot * kWrecBytes, wrec + (int64_t)i * VS_WREC, kWrecBytes, bar0 + 8 * slot);
    }
  };
#pragma unroll
  for (int d = 0; d < kBulkRing; d++) issue(b - 1 - d, d);
  double x[9], xl[9];
#pragma unroll
  for (int r = 0; r < 9; r++) {
    x[r] = delta[(int64_t)b * 9 + r];
    xl[r] = left >= 0 ? delta[(int64_t)left * 9 + r] : 0.0;
  }
  uint32_t phases = 0;
  int slot = 0;
  for (int i = b - 1; i >= a; i--) {
    mbar_wait(bar0 + 8 * slot, (phases >> slot) & 1u);
    phases ^= 1u << slot;
    double dr = 0.0;
    if (lane < 9) {
      const double* w = ring[slot];
      double zr = 0.0;
#pragma unroll
      for (int k = 0; k < 9; k++) zr = fma(w[90 + k * 9 + lane], xl[k], zr);     // off the critical path
      dr = w[81 + lane] - zr;
#pragma unroll
      for (int k = 0; k < 9; k++) dr = fma(-w[k * 9 + lane], x[k], dr);
    }
    __syncwarp();                              // the slot has been read before it is refilled
    issue(i - kBulkRing, slot);
#pragma unroll
    for (int k = 0; k < 9; k++) x[k] = __shfl_sync(0xffffffffu, dr, k);
    if (lane < 9) delta[(int64_t)i * 9 + lane] = dr;
    slot = (slot + 1 == kBulkRing) ? 0 : slot + 1;
  }
}

// VINSAT_NO_BULK_RING=1 selects the cp.async (LDGSTS) rings instead.
static bool use_bulk_ring() {
  static const bool on = getenv("VINSAT_NO_BULK_RING") == nullptr;
  return on;
}

static int launch_backrec(vinsat_ctx* ctx, const ChainArgs& A) {
  if (use_bulk_ring()) VS_LAUNCH(ctx, F_SOLVE, k_seg_backrec_bulk, ceil_div(A.n_chains, 4), 128, 0, A);
  else VS_LAUNCH(ctx, F_SOLVE, k_seg_backrec, ceil_div(A.n_chains, 4), 128, 0, A);
  return VINSAT_OK;
}

static int launch_backsub(vinsat_ctx* ctx, int64_t n, const int32_t* a, const int32_t* b, const int32_t* left,
                          const int32_t* prob, const int32_t* active, const double* wrec, double* delta) {
  // measured on a 2.4 M-frame arc (profiles/r02_longarc_2400k_kernels_bulk.json vs ..._level2.json): with 3552 chains in flight
  // the bulk ring is SLOWER here (0.90 vs 0.76 ms: a step is short, four bulk copies ahead do not cover the copy engine's
  // latency) while the backward recurrence gains (0.92 vs 1.28 ms); with 60 chains (second level, latency bound) it is
  // faster (21 vs 32 us).  Hence: bulk ring for fewer than 1024 chains, LDGSTS ring above.
  if (use_bulk_ring() && n < 1024)
    VS_LAUNCH(ctx, F_SOLVE, k_seg_backsub_bulk, ceil_div(n, 4), 128, 0, (int)n, a, b, left, prob, active, wrec, delta);
  else
    VS_LAUNCH(ctx, F_SOLVE, k_seg_backsub, ceil_div(n, 4), 128, 0, (int)n, a, b, left, prob, active, wrec, delta);
  return VINSAT_OK;
}

// reduced system of a frame-window sharded arc from the all-gathered per-segment packs
//   pack[s] = { redrec (VS_RREC) | system record of the separator frame (VS_SREC) }, one problem, global order
__global__ void __launch_bounds__(256) k_reduced_build_packed(int n_seg, const double* __restrict__ pack,
                                                              const double* __restrict__ lam,
                                                              double* __restrict__ rsys, double* __restrict__ rlow) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int s = (int)(t / 192);
  const int e = (int)(t % 192);
  if (s >= n_seg || e >= 171) return;
  constexpr int kPack = VS_RREC + VS_SREC;
  const double* mine = pack + (int64_t)s * kPack;
  const double* fr = mine + VS_RREC;
  const bool nx = s + 1 < n_seg;
  const double* next = pack + (int64_t)(s + 1) * kPack + 171;
  double v;
  if (e < 81) {
    v = fr[e] + mine[e] + (nx ? next[e] : 0.0);
    if (e / 9 == e % 9) v += (double)(float)lam[0];
  } else if (e < 162) {
    v = nx ? next[e] : 0.0;
  } else {
    v = fr[e] + mine[e] + (nx ? next[e] : 0.0);
  }
  rsys[(int64_t)s * VS_SREC + e] = v;
  if (e < 81 && s > 0) rlow[(int64_t)(s - 1) * 81 + e] = mine[81 + e];
}

static ChainArgs seg_args(vinsat_batch* b) {
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
  A.n_chains = (int)b->n_seg;
  A.ch_a = b->seg_a; A.ch_b = b->seg_b; A.ch_left = b->seg_left; A.ch_prob = b->seg_prob;
  return A;
}

int launch_seg_forward(vinsat_batch* b) {
  vinsat_ctx* ctx = b->ctx;
  if (b->n_seg == 0) return VINSAT_OK;
  ChainArgs A = seg_args(b);
  VS_TRY(launch_forward<true>(ctx, A));
  VS_TRY(launch_backrec(ctx, A));
  return VINSAT_OK;
}

int launch_seg_backsub(vinsat_batch* b) {
  vinsat_ctx* ctx = b->ctx;
  if (b->n_seg == 0) return VINSAT_OK;
  VS_TRY(launch_backsub(ctx, b->n_seg, b->seg_a, b->seg_b, b->seg_left, b->seg_prob, b->active, b->wrec, b->delta));
  return VINSAT_OK;
}

// separator solutions -> rows of the frames they belong to
__global__ void __launch_bounds__(128) k_sep_scatter(int n_seg, const int32_t* __restrict__ seg_b,
                                                     const int32_t* __restrict__ seg_prob, const int32_t* __restrict__ active,
                                                     const double* __restrict__ xsep, double* __restrict__ delta) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int s = (int)(t / 9), r = (int)(t % 9);
  if (s >= n_seg) return;
  if (active && !active[seg_prob[s]]) return;
  delta[(int64_t)seg_b[s] * 9 + r] = xsep[t];
}

// Second partition level: solves the reduced chains {rsys | rlow} (one element per level-1 separator, explicit lower blocks,
// already damped) by the same partitioned algorithm -- spike elimination of the level-2 interiors, backward recurrence,
// a (short) reduced-2 chain per problem, interior back-substitution -- and leaves the level-1 separator solutions in xsep.
int launch_level2_solve(vinsat_ctx* ctx, const Level2& L, const int32_t* active, const double* rsys, const double* rlow,
                        double* rwrec, double* xsep) {
  ChainArgs A;
  A.n_chains = (int)L.n;
  A.ch_a = L.a; A.ch_b = L.b; A.ch_left = L.left; A.ch_prob = L.prob;
  A.active = active;
  A.lam = nullptr;
  A.rec = rsys;
  A.lrec = rlow;
  A.wrec = rwrec;
  A.redrec = L.redrec;
  A.delta = xsep;
  A.out_index = nullptr;
  A.lam32_last = nullptr;
  VS_TRY(launch_forward<true>(ctx, A));
  VS_TRY(launch_backrec(ctx, A));
  VS_LAUNCH(ctx, F_SOLVE, k_reduced_build, ceil_div(L.n * 192, 256), 256, 0, (int)L.n, L.b, L.left, L.prob, L.has_next, active,
            (const double*)nullptr, rsys, L.redrec, L.rsys, L.rlow);
  ChainArgs R;
  R.n_chains = (int)L.n_chains;
  R.ch_a = L.red_a; R.ch_b = L.red_b; R.ch_left = nullptr; R.ch_prob = L.red_prob;
  R.active = active;
  R.lam = nullptr;
  R.rec = L.rsys;
  R.lrec = L.rlow;
  R.wrec = L.rwrec;
  R.redrec = nullptr;
  R.delta = xsep;
  R.out_index = L.b;               // level-2 separator -> level-1 separator it is
  R.lam32_last = nullptr;
  VS_TRY(launch_forward<false>(ctx, R));
  VS_LAUNCH(ctx, F_SOLVE, k_chain_backward, ceil_div(R.n_chains, 4), 128, 0, R);
  VS_TRY(launch_backsub(ctx, L.n, L.a, L.b, L.left, L.prob, active, rwrec, xsep));
  return VINSAT_OK;
}

int launch_reduced_packed(vinsat_batch* b, int64_t S_total, const double* pack, double* rsys, double* rlow,
                          double* rwrec, double* xsep, const int32_t* one_chain) {
  vinsat_ctx* ctx = b->ctx;
  VS_LAUNCH(ctx, F_SOLVE, k_reduced_build_packed, ceil_div(S_total * 192, 256), 256, 0, (int)S_total, pack, b->lam, rsys,
            rlow);
  if (b->la_l2.n > 0) return launch_level2_solve(ctx, b->la_l2, nullptr, rsys, rlow, rwrec, xsep);
  ChainArgs R;
  R.n_chains = 1;
  R.ch_a = one_chain; R.ch_b = one_chain + 1; R.ch_left = nullptr; R.ch_prob = one_chain + 2;
  R.active = nullptr;
  R.lam = nullptr;
  R.rec = rsys;
  R.lrec = rlow;
  R.wrec = rwrec;
  R.redrec = nullptr;
  R.delta = xsep;
  R.out_index = nullptr;
  R.lam32_last = nullptr;
  VS_TRY(launch_forward<false>(ctx, R));
  VS_LAUNCH(ctx, F_SOLVE, k_chain_backward, 1, 128, 0, R);
  return VINSAT_OK;
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
    static const bool one_sided = getenv("VINSAT_ONE_SIDED_SWEEP") != nullptr;
    if (one_sided) {
      A.n_chains = (int)b->P;
      A.ch_a = b->pl_a; A.ch_b = b->pl_b; A.ch_left = nullptr; A.ch_prob = b->pl_prob;
    } else {
      // two-sided sweep: a top chain and a bottom chain per problem eliminate towards the middle frame
      // (same flops as the one-sided sweep, twice the independent chains, half the sequential length)
      A.n_chains = (int)(2 * b->P);
      A.ch_a = b->bb_a; A.ch_b = b->bb_e; A.ch_left = nullptr; A.ch_prob = b->bb_prob;
      A.ch_dir = b->bb_dir; A.ch_mid = b->bb_mid; A.mid = b->midrec;
      A.gate = b->gate_arg;
      if (b->fused_system) {
        A.fused = 1; A.rec = nullptr;
        A.fs.grec = b->grec; A.fs.drec = b->drec; A.fs.mrec = b->mrec; A.fs.gap = b->gap; A.fs.wmax = b->wmax; A.fs.zeros = b->zeros;
        A.fs.Sigma = b->cur_sigma; A.fs.vc = b->cur_vc;
      }
    }
    VS_TRY(launch_forward<false>(ctx, A));
    VS_LAUNCH(ctx, F_SOLVE_BWD, k_chain_backward, ceil_div(A.n_chains, 4), 128, 0, A);
    return VINSAT_OK;
  }
  A.n_chains = (int)b->n_seg;
  A.ch_a = b->seg_a; A.ch_b = b->seg_b; A.ch_left = b->seg_left; A.ch_prob = b->seg_prob;
  VS_TRY(launch_forward<true>(ctx, A));
  VS_TRY(launch_backrec(ctx, A));
  VS_LAUNCH(ctx, F_SOLVE, k_reduced_build, ceil_div((int64_t)b->n_seg * 192, 256), 256, 0, (int)b->n_seg, b->seg_b,
            b->seg_left, b->seg_prob, b->seg_has_next, b->active, b->lam, b->srec, b->redrec, b->rsys, b->rlow);
  if (b->l2.n > 0) {
    VS_TRY(launch_level2_solve(ctx, b->l2, b->active, b->rsys, b->rlow, b->rwrec, b->xsep));
    VS_LAUNCH(ctx, F_SOLVE, k_sep_scatter, ceil_div(b->n_seg * 9, 128), 128, 0, (int)b->n_seg, b->seg_b, b->seg_prob, b->active,
              b->xsep, b->delta);
    VS_TRY(launch_backsub(ctx, b->n_seg, b->seg_a, b->seg_b, b->seg_left, b->seg_prob, b->active, b->wrec, b->delta));
    return VINSAT_OK;
  }
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
  VS_TRY(launch_forward<false>(ctx, R));
  VS_LAUNCH(ctx, F_SOLVE, k_chain_backward, ceil_div(R.n_chains, 4), 128, 0, R);
  VS_TRY(launch_backsub(ctx, b->n_seg, b->seg_a, b->seg_b, b->seg_left, b->seg_prob, b->active, b->wrec, b->delta));
  return VINSAT_OK;
}

}  // namespace vs
