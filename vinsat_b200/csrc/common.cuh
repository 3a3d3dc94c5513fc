// Device math shared by the kernels: quaternion algebra, pinhole projection with its closed-form
// Jacobian, two-body+J2 acceleration with its gradient, RK4 with the variational equation.
// Citations: path:line under <reference>/estimation.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vs {

constexpr double kMu = 398600.4418;    // BA/BA_utils.py:883
constexpr double kJ2 = 1.75553e10;     // BA/BA_utils.py:883

struct Quat { double x, y, z, w; };
struct Vec3 { double x, y, z; };

// ---- exact (unfused) arithmetic: every op rounds separately, like NumPy / ATen on the CPU ----------
__device__ __forceinline__ double xmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double xsub(double a, double b) { return __dsub_rn(a, b); }

// BA_utils.py:992-1000, evaluated left to right without contraction.
__device__ __forceinline__ Quat qmul_exact(const Quat& a, const Quat& b) {
  Quat r;
  r.w = xsub(xsub(xsub(xmul(a.w, b.w), xmul(a.x, b.x)), xmul(a.y, b.y)), xmul(a.z, b.z));
  r.x = xsub(xadd(xadd(xmul(a.w, b.x), xmul(a.x, b.w)), xmul(a.y, b.z)), xmul(a.z, b.y));
  r.y = xadd(xadd(xsub(xmul(a.w, b.y), xmul(a.x, b.z)), xmul(a.y, b.w)), xmul(a.z, b.x));
  r.z = xadd(xsub(xadd(xmul(a.w, b.z), xmul(a.x, b.y)), xmul(a.y, b.x)), xmul(a.z, b.w));
  return r;
}

// Same product, contraction allowed (used where 1e-9 relative is the bar).
__device__ __forceinline__ Quat qmul(const Quat& a, const Quat& b) {
  Quat r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
  r.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
  return r;
}

__device__ __forceinline__ double qnorm_exact(const Quat& q) {
  return sqrt(xadd(xadd(xadd(xmul(q.x, q.x), xmul(q.y, q.y)), xmul(q.z, q.z)), xmul(q.w, q.w)));
}

// BA_utils.py:970-985
__device__ __forceinline__ Quat qexp(double dx, double dy, double dz) {
  double t = sqrt(dx * dx + dy * dy + dz * dz);
  Quat q;
  if (t < 1e-16) { q.x = 0; q.y = 0; q.z = 0; q.w = 1; return q; }
  double s, c;
  sincos(0.5 * t, &s, &c);
  double k = s / (t + 1e-16);
  q.x = dx * k; q.y = dy * k; q.z = dz * k; q.w = c;
  return q;
}

// ---- a1: pinhole projection ---------------------------------------------------------------------
struct ProjOut {
  double u, v;          // pixel estimate
  double Xc, Yc, Zc;    // camera-frame point
  double d;             // 1/max(Zc, 0.1)
  Quat qn;              // normalised quaternion
};

// Forward projection, bit-compatible with the reference op order:
//   point = X - p; qn = q/|q|; p_c = conj(qn) (x) ((point,0) (x) qn)    (BA_utils.py:1052-1069)
//   d = 1/clamp(Z, 0.1); u = fx*(d*X)+cx; v = fy*(d*Y)+cy                (BA_utils.py:7-17)
__device__ __forceinline__ ProjOut project_exact(double px, double py, double pz, Quat q, double X, double Y,
                                                 double Z, double fx, double fy, double cx, double cy) {
  ProjOut o;
  double n = qnorm_exact(q);
  Quat qn = {q.x / n, q.y / n, q.z / n, q.w / n};
  Quat v = {xsub(X, px), xsub(Y, py), xsub(Z, pz), 0.0};
  Quat qc = {-qn.x, -qn.y, -qn.z, qn.w};
  Quat t = qmul_exact(v, qn);
  Quat r = qmul_exact(qc, t);
  o.Xc = r.x; o.Yc = r.y; o.Zc = r.z; o.qn = qn;
  double zc = r.z < 0.1 ? 0.1 : r.z;    // clamp(min=0.1); NaN propagates like torch
  if (r.z != r.z) zc = r.z;
  o.d = 1.0 / zc;
  o.u = xadd(xmul(fx, xmul(o.d, r.x)), cx);
  o.v = xadd(xmul(fy, xmul(o.d, r.y)), cy);
  return o;
}

// Closed-form Jacobian rows: Jg = [-Pi R^T | 2 Pi hat(p_c)], 2x6 nonzero columns.
// ju[0..5], jv[0..5].
__device__ __forceinline__ void project_jacobian(const ProjOut& o, double fx, double fy, double* ju, double* jv) {
  const double live = (o.Zc >= 0.1) ? 1.0 : 0.0;
  const double a = fx * o.d, b = fy * o.d;
  const double c = -fx * o.Xc * o.d * o.d * live;
  const double e = -fy * o.Yc * o.d * o.d * live;
  const double x = o.qn.x, y = o.qn.y, z = o.qn.z, w = o.qn.w;
  // R^T rows (R^T v = conj(q) v q)
  const double r00 = 1 - 2 * (y * y + z * z), r01 = 2 * (x * y + z * w), r02 = 2 * (x * z - y * w);
  const double r10 = 2 * (x * y - z * w), r11 = 1 - 2 * (x * x + z * z), r12 = 2 * (y * z + x * w);
  const double r20 = 2 * (x * z + y * w), r21 = 2 * (y * z - x * w), r22 = 1 - 2 * (x * x + y * y);
  ju[0] = -(a * r00 + c * r20); ju[1] = -(a * r01 + c * r21); ju[2] = -(a * r02 + c * r22);
  jv[0] = -(b * r10 + e * r20); jv[1] = -(b * r11 + e * r21); jv[2] = -(b * r12 + e * r22);
  // 2 Pi hat(p_c)
  ju[3] = 2 * (-c * o.Yc);            ju[4] = 2 * (c * o.Xc - a * o.Zc); ju[5] = 2 * (a * o.Yc);
  jv[3] = 2 * (b * o.Zc - e * o.Yc);  jv[4] = 2 * (e * o.Xc);            jv[5] = 2 * (-b * o.Xc);
}

// ---- a3: orbit dynamics (BA_utils.py:883-899) ------------------------------------------------------
// accel = -mu r/|r|^3 + (J2/|r|^7) (M r^2) (.) r,  M = [[6,-1.5,-1.5],[6,-1.5,-1.5],[3,-4.5,-4.5]].
__device__ __forceinline__ void accel(double x, double y, double z, double& ax, double& ay, double& az) {
  const double x2 = x * x, y2 = y * y, z2 = z * z;
  const double n2 = x2 + y2 + z2;
  const double n = sqrt(n2);
  const double n3 = n2 * n;
  const double n7 = n3 * n2 * n2;
  const double k = -kMu / n3, j = kJ2 / n7;
  const double sx = 6.0 * x2 - 1.5 * y2 - 1.5 * z2;
  const double sz = 3.0 * x2 - 4.5 * y2 - 4.5 * z2;
  ax = k * x + j * sx * x;
  ay = k * y + j * sx * y;
  az = k * z + j * sz * z;
}

// accel and G = d accel / d r (row-major g[9]; not symmetric because M is not).
__device__ __forceinline__ void accel_grad(double x, double y, double z, double& ax, double& ay, double& az,
                                           double* g) {
  const double x2 = x * x, y2 = y * y, z2 = z * z;
  const double n2 = x2 + y2 + z2;
  const double inv_n = rsqrt(n2);
  const double inv_n2 = inv_n * inv_n;      // one rsqrt, no division (2 ulp; the STM tolerance is 1e-9 relative)
  const double inv_n3 = inv_n * inv_n2;
  const double inv_n5 = inv_n3 * inv_n2;
  const double inv_n7 = inv_n5 * inv_n2;
  const double inv_n9 = inv_n7 * inv_n2;
  const double k = -kMu * inv_n3, j = kJ2 * inv_n7;
  const double sx = 6.0 * x2 - 1.5 * y2 - 1.5 * z2;
  const double sz = 3.0 * x2 - 4.5 * y2 - 4.5 * z2;
  // accel_i = r_i (k + j s_i);  G_ic = r_i r_c (3 mu/n^5 - 7 J2 s_i/n^9 + 2 j M_ic) + delta_ic (k + j s_i), factored
  // so that every entry is one product of r_i r_c with one of four row/column coefficients
  const double dxy = k + j * sx, dz = k + j * sz;
  ax = dxy * x;
  ay = dxy * y;
  az = dz * z;
  const double t3 = 3.0 * kMu * inv_n5, t7 = -7.0 * kJ2 * inv_n9, j2 = 2.0 * j;
  const double cx = t3 + t7 * sx, cz = t3 + t7 * sz;
  const double cxa = cx + 6.0 * j2, cxb = cx - 1.5 * j2, cza = cz + 3.0 * j2, czb = cz - 4.5 * j2;
  const double xy = x * y, xz = x * z, yz = y * z;
  g[0] = x2 * cxa + dxy; g[1] = xy * cxb;       g[2] = xz * cxb;
  g[3] = xy * cxa;       g[4] = y2 * cxb + dxy; g[5] = yz * cxb;
  g[6] = xz * cza;       g[7] = yz * czb;       g[8] = z2 * czb + dz;
}

// Same acceleration with the arithmetic of accel_grad (one rsqrt, no division, no square root): used by the LM
// trial residual, whose trajectory then follows the one the Jacobian kernel integrates.
__device__ __forceinline__ void accel_fast(double x, double y, double z, double& ax, double& ay, double& az) {
  const double x2 = x * x, y2 = y * y, z2 = z * z;
  const double n2 = x2 + y2 + z2;
  const double inv_n = rsqrt(n2);
  const double inv_n2 = inv_n * inv_n;
  const double inv_n3 = inv_n * inv_n2;
  const double inv_n7 = inv_n3 * inv_n2 * inv_n2;
  const double k = -kMu * inv_n3, j = kJ2 * inv_n7;
  const double sx = 6.0 * x2 - 1.5 * y2 - 1.5 * z2;
  const double sz = 3.0 * x2 - 4.5 * y2 - 4.5 * z2;
  const double dxy = k + j * sx, dz = k + j * sz;
  ax = dxy * x;
  ay = dxy * y;
  az = dz * z;
}

// One classic RK4 step of the 6-state (BA_utils.py:901-912).
template <bool FAST = false>
__device__ __forceinline__ void rk4_step(double* x, double h) {
  double a1x, a1y, a1z, a2x, a2y, a2z, a3x, a3y, a3z, a4x, a4y, a4z;
  const double hh = 0.5 * h;
  auto accel = [](double px, double py, double pz, double& ox, double& oy, double& oz) {
    if (FAST) accel_fast(px, py, pz, ox, oy, oz);
    else vs::accel(px, py, pz, ox, oy, oz);
  };
  accel(x[0], x[1], x[2], a1x, a1y, a1z);
  double p2x = x[0] + hh * x[3], p2y = x[1] + hh * x[4], p2z = x[2] + hh * x[5];
  double v2x = x[3] + hh * a1x, v2y = x[4] + hh * a1y, v2z = x[5] + hh * a1z;
  accel(p2x, p2y, p2z, a2x, a2y, a2z);
  double p3x = x[0] + hh * v2x, p3y = x[1] + hh * v2y, p3z = x[2] + hh * v2z;
  double v3x = x[3] + hh * a2x, v3y = x[4] + hh * a2y, v3z = x[5] + hh * a2z;
  accel(p3x, p3y, p3z, a3x, a3y, a3z);
  double p4x = x[0] + h * v3x, p4y = x[1] + h * v3y, p4z = x[2] + h * v3z;
  double v4x = x[3] + h * a3x, v4y = x[4] + h * a3y, v4z = x[5] + h * a3z;
  accel(p4x, p4y, p4z, a4x, a4y, a4z);
  const double h6 = h / 6.0;
  x[0] += h6 * (x[3] + 2 * v2x + 2 * v3x + v4x);
  x[1] += h6 * (x[4] + 2 * v2y + 2 * v3y + v4y);
  x[2] += h6 * (x[5] + 2 * v2z + 2 * v3z + v4z);
  x[3] += h6 * (a1x + 2 * a2x + 2 * a3x + a4x);
  x[4] += h6 * (a1y + 2 * a2y + 2 * a3y + a4y);
  x[5] += h6 * (a1z + 2 * a2z + 2 * a3z + a4z);
}

// RK4 step of the state AND of NC columns of the state-transition matrix (variational equation through
// the same stages = exact Jacobian of the discrete map).  phi[c][6] are columns.
template <int NC>
__device__ __forceinline__ void rk4_step_stm(double* x, double (*phi)[6], double h) {
  const double hh = 0.5 * h, h6 = h / 6.0;
  double acc[NC][6];   // weighted sum of stage derivatives
  double Y[NC][6];     // stage argument of the variational system
  double g[9];
  double ax, ay, az;
  double xs[6];        // stage state
  double xa[6];        // accumulated state derivative
  // stage 1
  accel_grad(x[0], x[1], x[2], ax, ay, az, g);
  double f[6] = {x[3], x[4], x[5], ax, ay, az};
#pragma unroll
  for (int k = 0; k < 6; k++) { xa[k] = f[k]; xs[k] = x[k] + hh * f[k]; }
#pragma unroll
  for (int c = 0; c < NC; c++) {
    double d[6] = {phi[c][3], phi[c][4], phi[c][5],
                   g[0] * phi[c][0] + g[1] * phi[c][1] + g[2] * phi[c][2],
                   g[3] * phi[c][0] + g[4] * phi[c][1] + g[5] * phi[c][2],
                   g[6] * phi[c][0] + g[7] * phi[c][1] + g[8] * phi[c][2]};
#pragma unroll
    for (int k = 0; k < 6; k++) { acc[c][k] = d[k]; Y[c][k] = phi[c][k] + hh * d[k]; }
  }
  // stages 2, 3
#pragma unroll
  for (int st = 0; st < 2; st++) {
    const double cn = (st == 0) ? hh : h;   // coefficient of the NEXT stage argument
    accel_grad(xs[0], xs[1], xs[2], ax, ay, az, g);
    double f2[6] = {xs[3], xs[4], xs[5], ax, ay, az};
#pragma unroll
    for (int k = 0; k < 6; k++) { xa[k] += 2.0 * f2[k]; xs[k] = x[k] + cn * f2[k]; }
#pragma unroll
    for (int c = 0; c < NC; c++) {
      double d[6] = {Y[c][3], Y[c][4], Y[c][5],
                     g[0] * Y[c][0] + g[1] * Y[c][1] + g[2] * Y[c][2],
                     g[3] * Y[c][0] + g[4] * Y[c][1] + g[5] * Y[c][2],
                     g[6] * Y[c][0] + g[7] * Y[c][1] + g[8] * Y[c][2]};
#pragma unroll
      for (int k = 0; k < 6; k++) { acc[c][k] += 2.0 * d[k]; Y[c][k] = phi[c][k] + cn * d[k]; }
    }
  }
  // stage 4
  accel_grad(xs[0], xs[1], xs[2], ax, ay, az, g);
  {
    double f4[6] = {xs[3], xs[4], xs[5], ax, ay, az};
#pragma unroll
    for (int k = 0; k < 6; k++) x[k] += h6 * (xa[k] + f4[k]);
#pragma unroll
    for (int c = 0; c < NC; c++) {
      double d[6] = {Y[c][3], Y[c][4], Y[c][5],
                     g[0] * Y[c][0] + g[1] * Y[c][1] + g[2] * Y[c][2],
                     g[3] * Y[c][0] + g[4] * Y[c][1] + g[5] * Y[c][2],
                     g[6] * Y[c][0] + g[7] * Y[c][1] + g[8] * Y[c][2]};
#pragma unroll
      for (int k = 0; k < 6; k++) phi[c][k] += h6 * (acc[c][k] + d[k]);
    }
  }
}

// Step size of hop k for a frame gap in the two propagator modes.
//   STEP1S : gap steps of 1 s.   SKIP100: gap/100 steps of 100 s, then one step of gap%100 (possibly 0).
__device__ __forceinline__ int num_hops(int gap, int mode) { return mode == 0 ? gap : gap / 100 + 1; }
__device__ __forceinline__ double hop_size(int gap, int mode, int k) {
  if (mode == 0) return 1.0;
  return (k < gap / 100) ? 100.0 : (double)(gap % 100);
}

}  // namespace vs
