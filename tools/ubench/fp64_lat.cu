// Micro-benchmark (measurement tool, not part of the library): dependent-issue latencies of the FP64 pipe, the
// shared-memory broadcast round trip and warp shuffles on one warp of one SM, in SM clock cycles.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/fp64_lat tools/ubench/fp64_lat.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}

template <int MODE>
__global__ void k(double* out, long long* cyc, int n, double a, double b) {
  __shared__ double s[64];
  double x = a + threadIdx.x * 1e-9, y = b;
  double x1 = x + 1, x2 = x + 2, x3 = x + 3, x4 = x + 4, x5 = x + 5, x6 = x + 6, x7 = x + 7;
  s[threadIdx.x] = x;
  __syncwarp();
  long long t0 = clock64();
  for (int i = 0; i < n; i++) {
    if (MODE == 0) { x = fma(x, y, y); }                                  // dependent DFMA
    if (MODE == 1) { x = x * y; }                                         // dependent DMUL
    if (MODE == 2) { x = x + y; }                                         // dependent DADD
    if (MODE == 3) { x = fast_rcp(x) + y; }                               // rcp.approx + 2 Newton steps (+1 DADD)
    if (MODE == 4) { x = 1.0 / x + y; }                                   // IEEE division (+1 DADD)
    if (MODE == 5) { x = sqrt(x) + y; }
    if (MODE == 6) { x = rsqrt(x) + y; }
    if (MODE == 7) {                                                      // STS -> syncwarp -> LDS (broadcast read)
      s[threadIdx.x] = x;
      __syncwarp();
      x = s[(threadIdx.x + 1) & 31] + y;
      __syncwarp();
    }
    if (MODE == 8) { x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31) + y; }   // double shuffle (+1 DADD)
    if (MODE == 9) {                                                      // 8 independent DFMA chains: issue interval
      x = fma(x, y, y); x1 = fma(x1, y, y); x2 = fma(x2, y, y); x3 = fma(x3, y, y);
      x4 = fma(x4, y, y); x5 = fma(x5, y, y); x6 = fma(x6, y, y); x7 = fma(x7, y, y);
    }
  }
  long long t1 = clock64();
  out[threadIdx.x + blockIdx.x * blockDim.x] = x + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
double run(int warps, int n) {
  double* out; long long* cyc;
  cudaMalloc(&out, 4096 * sizeof(double)); cudaMalloc(&cyc, 8);
  k<MODE><<<1, 32 * warps>>>(out, cyc, n, 1.000001, 0.999999);
  k<MODE><<<1, 32 * warps>>>(out, cyc, n, 1.000001, 0.999999);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  cudaFree(out); cudaFree(cyc);
  return (double)h / n;
}

int main() {
  const int n = 4096;
  const char* names[] = {"dep DFMA", "dep DMUL", "dep DADD", "fast_rcp + DADD", "1.0/x + DADD", "sqrt + DADD",
                         "rsqrt + DADD", "STS+syncwarp+LDS+DADD+syncwarp", "SHFL(double) + DADD", "8 indep DFMA (per 8)"};
  for (int w : {1, 2, 4}) {
    double v[10] = {run<0>(w, n), run<1>(w, n), run<2>(w, n), run<3>(w, n), run<4>(w, n), run<5>(w, n), run<6>(w, n),
                    run<7>(w, n), run<8>(w, n), run<9>(w, n)};
    for (int i = 0; i < 10; i++) printf("{\"warps_per_cta\": %d, \"op\": \"%s\", \"cycles_per_iter\": %.2f}\n", w, names[i], v[i]);
  }
  return 0;
}
