// Internal declarations shared by the translation units of libvinsat_b200.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/vinsat_b200.h"

namespace vs {

// Kernel families for the launch counter / per-family device timing.
enum Family {
  F_PROJECT = 0,      // a1 stand-alone + headline residual/Jacobian kernel
  F_OBS_RESID,        // residuals for the robust scale
  F_SELECT,           // radix select (median) + weights max
  F_OBS_ASSEMBLE,     // fused projection + weights + per-frame JtWJ / JtWr
  F_DYNAMICS,         // RK4 + STM
  F_QUAT,             // quaternion smoothness gradient / Hessian blocks
  F_SYSTEM,           // block-tridiagonal system build
  F_SOLVE,            // block-tridiagonal LU solve: forward elimination (+ partition / reduced-system kernels)
  F_SOLVE_BWD,        // block-tridiagonal LU solve: back-substitution of the whole-problem chains
  F_SOLVE_INIT,       // initialize phase: per-frame Cholesky
  F_RETRACT,          // retraction
  F_TRIAL,            // trial residual evaluation (obs + dynamics)
  F_ACCEPT,           // per-problem reductions, LM accept test
  F_LAYOUT,           // AoS<->SoA conversion, gathers
  F_SIM,              // orbit propagation / chain
  F_SATCAM,           // SatCam projection / visibility
  F_PEAK,             // peak microbenchmarks
  F_COUNT
};
extern const char* const kFamilyNames[F_COUNT];

struct TimedLaunch { int family; cudaEvent_t e0, e1; };

}  // namespace vs

struct vinsat_ctx {
  int device = 0;
  int sm_count = 148;
  int max_smem_optin = 48 * 1024;      // cudaDevAttrMaxSharedMemoryPerBlockOptin of `device`
  bool smem_optin_done[8] = {false};
  int smem_optin_cap[8] = {0};   // per kernel (vs::SmemSlot): the opt-in is per device, so it is tracked per context
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  std::string err;
  int64_t launches = 0;
  bool timing = false;
  double fam_ms[vs::F_COUNT] = {0};
  int64_t fam_launches[vs::F_COUNT] = {0};
  std::vector<vs::TimedLaunch> pending;
  std::vector<cudaEvent_t> event_pool;
  // scratch (grown on demand)
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
  void* pinned = nullptr;
  size_t pinned_bytes = 0;
  // bump arena for batches that are created and destroyed in a loop (vinsat_stream_solve): one cudaMalloc for all windows
  char* arena = nullptr;
  size_t arena_bytes = 0, arena_off = 0;
  bool arena_on = false;
};

namespace vs {

int set_error(vinsat_ctx* ctx, int code, const char* fmt, ...);
void set_global_error(const char* msg);
void timing_begin(vinsat_ctx* ctx, int family);
void timing_end(vinsat_ctx* ctx);
void timing_resolve(vinsat_ctx* ctx);
void* ctx_scratch(vinsat_ctx* ctx, size_t bytes);   // device scratch, valid until the next call
void* ctx_pinned(vinsat_ctx* ctx, size_t bytes);    // pinned host scratch

#define VS_CUDA(ctx, expr)                                                                      \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return vs::set_error((ctx), VINSAT_ECUDA, "%s failed: %s (%s:%d)", #expr,                 \
                           cudaGetErrorString(_e), __FILE__, __LINE__);                         \
  } while (0)

#define VS_CHECK_ARG(ctx, cond)                                                                 \
  do {                                                                                          \
    if (!(cond)) return vs::set_error((ctx), VINSAT_EINVAL, "invalid argument: %s", #cond);     \
  } while (0)

// Launch wrapper: counts the launch, optionally brackets it with events, checks the launch error.
#define VS_LAUNCH(ctx, family, kernel, grid, block, smem, ...)                                  \
  do {                                                                                          \
    vs::timing_begin((ctx), (family));                                                          \
    kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                            \
    vs::timing_end((ctx));                                                                      \
    cudaError_t _e = cudaGetLastError();                                                        \
    if (_e != cudaSuccess)                                                                      \
      return vs::set_error((ctx), VINSAT_ECUDA, "launch %s failed: %s (%s:%d)", #kernel,        \
                           cudaGetErrorString(_e), __FILE__, __LINE__);                         \
  } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Kernels that need more than 48 KB of dynamic shared memory.  cudaFuncSetAttribute applies to the CURRENT device only,
// so the opt-in is remembered per context (one context = one device), never in a process-wide static; it is always
// raised to the device maximum so that two contexts on one device cannot lower each other's limit.
enum SmemSlot { SM_SELECT = 0, SM_ASM_STAGED, SM_ASM_2PASS, SM_SYSROWS, SM_TRIAL, SM_SPARE0, SM_SPARE1, SM_SPARE2 };
int smem_optin(vinsat_ctx* ctx, int slot, const void* func, const char* name, int bytes);
#define VS_SMEM_OPTIN(ctx, slot, kernel, bytes)                                                              \
  do {                                                                                                       \
    const int _rc = vs::smem_optin((ctx), (slot), (const void*)(kernel), #kernel, (int)(bytes));             \
    if (_rc != VINSAT_OK) return _rc;                                                                        \
  } while (0)

// RAII device buffer for the stand-alone entry points.
template <typename T>
struct DevBuf {
  T* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t n) { return cudaMalloc((void**)&p, (n ? n : 1) * sizeof(T)); }
};

}  // namespace vs
