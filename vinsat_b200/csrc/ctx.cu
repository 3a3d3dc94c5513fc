// Context, error reporting, launch accounting, scratch memory, peak microbenchmarks.
#include <stdarg.h>
#include <string.h>

#include <mutex>

#include "internal.h"

namespace vs {

const char* const kFamilyNames[F_COUNT] = {
    "project_resjac", "obs_residual", "select_median", "obs_assemble", "dynamics_stm", "quat_terms", "system_build",
    "blocktridiag_solve", "blocktridiag_backsub", "solve_init", "retract", "trial_residual", "accept_reduce", "layout", "orbit_sim", "satcam", "peak"};

static std::mutex g_err_mu;
static std::string g_err;

void set_global_error(const char* msg) {
  std::lock_guard<std::mutex> lk(g_err_mu);
  g_err = msg;
}

int set_error(vinsat_ctx* ctx, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  set_global_error(buf);
  return code;
}

static cudaEvent_t get_event(vinsat_ctx* ctx) {
  if (!ctx->event_pool.empty()) {
    cudaEvent_t e = ctx->event_pool.back();
    ctx->event_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

void timing_begin(vinsat_ctx* ctx, int family) {
  ctx->launches++;
  if (!ctx->timing) return;
  TimedLaunch t;
  t.family = family;
  t.e0 = get_event(ctx);
  t.e1 = get_event(ctx);
  cudaEventRecord(t.e0, ctx->stream);
  ctx->pending.push_back(t);
}

void timing_end(vinsat_ctx* ctx) {
  if (!ctx->timing) return;
  cudaEventRecord(ctx->pending.back().e1, ctx->stream);
  if (ctx->pending.size() >= 4096) timing_resolve(ctx);
}

void timing_resolve(vinsat_ctx* ctx) {
  if (ctx->pending.empty()) return;
  cudaStreamSynchronize(ctx->stream);
  for (auto& t : ctx->pending) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, t.e0, t.e1) == cudaSuccess) {
      ctx->fam_ms[t.family] += ms;
      ctx->fam_launches[t.family] += 1;
    }
    ctx->event_pool.push_back(t.e0);
    ctx->event_pool.push_back(t.e1);
  }
  ctx->pending.clear();
}

int smem_optin(vinsat_ctx* ctx, int slot, const void* func, const char* name, int bytes) {
  if (!ctx->smem_optin_done[slot]) {
    // dynamic limit = device opt-in maximum minus the kernel's static shared memory
    cudaFuncAttributes fa;
    VS_CUDA(ctx, cudaFuncGetAttributes(&fa, func));
    ctx->smem_optin_cap[slot] = ctx->max_smem_optin - (int)fa.sharedSizeBytes;
    VS_CUDA(ctx, cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->smem_optin_cap[slot]));
    ctx->smem_optin_done[slot] = true;
  }
  if (bytes > ctx->smem_optin_cap[slot])
    return set_error(ctx, VINSAT_EINVAL, "%s needs %d B of dynamic shared memory, device allows %d", name, bytes,
                     ctx->smem_optin_cap[slot]);
  return VINSAT_OK;
}

void* ctx_scratch(vinsat_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->scratch_bytes) return ctx->scratch;
  if (ctx->scratch) {
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->scratch);
    ctx->scratch = nullptr;
    ctx->scratch_bytes = 0;
  }
  size_t want = bytes + bytes / 4 + 4096;
  if (cudaMalloc(&ctx->scratch, want) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  ctx->scratch_bytes = want;
  return ctx->scratch;
}

void* ctx_pinned(vinsat_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->pinned_bytes) return ctx->pinned;
  if (ctx->pinned) {
    cudaStreamSynchronize(ctx->stream);
    cudaFreeHost(ctx->pinned);
    ctx->pinned = nullptr;
    ctx->pinned_bytes = 0;
  }
  size_t want = bytes + bytes / 4 + 4096;
  if (cudaMallocHost(&ctx->pinned, want) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  ctx->pinned_bytes = want;
  return ctx->pinned;
}

// ---- peak microbenchmarks ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, int iters, double seed) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
         a7 = a0 + 7;
  const double m = 0.999999, c = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == 123.456) out[0] = s;   // never true; keeps the loop alive
}

__global__ void __launch_bounds__(256) k_copy(const double2* __restrict__ a, double2* __restrict__ b, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) b[i] = a[i];
}

}  // namespace vs

using namespace vs;

extern "C" {

int vinsat_abi_version(void) { return VINSAT_ABI_VERSION; }

int vinsat_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

const char* vinsat_last_error(const vinsat_ctx* ctx) {
  if (ctx) return ctx->err.c_str();
  static thread_local std::string copy;
  std::lock_guard<std::mutex> lk(g_err_mu);
  copy = g_err;
  return copy.c_str();
}

int vinsat_ctx_create(int device, vinsat_ctx** out) {
  if (!out) return set_error(nullptr, VINSAT_EINVAL, "vinsat_ctx_create: out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return set_error(nullptr, VINSAT_ENODEV, "no CUDA device available (%s); this library has no CPU fallback",
                     e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  }
  if (device < 0 || device >= n) return set_error(nullptr, VINSAT_EINVAL, "device %d out of range [0,%d)", device, n);
  vinsat_ctx* ctx = new vinsat_ctx();
  ctx->device = device;
  VS_CUDA(ctx, cudaSetDevice(device));
  cudaDeviceProp prop;
  VS_CUDA(ctx, cudaGetDeviceProperties(&prop, device));
  ctx->sm_count = prop.multiProcessorCount;
  ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  VS_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
  ctx->stream = ctx->own_stream;
  *out = ctx;
  return VINSAT_OK;
}

int vinsat_ctx_destroy(vinsat_ctx* ctx) {
  if (!ctx) return VINSAT_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  timing_resolve(ctx);
  for (auto e : ctx->event_pool) cudaEventDestroy(e);
  if (ctx->scratch) cudaFree(ctx->scratch);
  if (ctx->arena) cudaFree(ctx->arena);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
  return VINSAT_OK;
}

int vinsat_ctx_set_stream(vinsat_ctx* ctx, void* cuda_stream) {
  if (!ctx) return VINSAT_EINVAL;
  cudaStreamSynchronize(ctx->stream);
  timing_resolve(ctx);
  ctx->stream = (cudaStream_t)cuda_stream;
  return VINSAT_OK;
}

int vinsat_ctx_reset_stream(vinsat_ctx* ctx) {
  if (!ctx) return VINSAT_EINVAL;
  cudaStreamSynchronize(ctx->stream);
  timing_resolve(ctx);
  ctx->stream = ctx->own_stream;
  return VINSAT_OK;
}

int vinsat_ctx_synchronize(vinsat_ctx* ctx) {
  if (!ctx) return VINSAT_EINVAL;
  VS_CUDA(ctx, cudaSetDevice(ctx->device));
  VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return VINSAT_OK;
}

int vinsat_ctx_enable_timing(vinsat_ctx* ctx, int on) {
  if (!ctx) return VINSAT_EINVAL;
  timing_resolve(ctx);
  ctx->timing = on != 0;
  return VINSAT_OK;
}

int vinsat_ctx_reset_timing(vinsat_ctx* ctx) {
  if (!ctx) return VINSAT_EINVAL;
  timing_resolve(ctx);
  for (int i = 0; i < F_COUNT; i++) {
    ctx->fam_ms[i] = 0;
    ctx->fam_launches[i] = 0;
  }
  return VINSAT_OK;
}

int vinsat_ctx_get_timing(vinsat_ctx* ctx, int cap, const char** names_out, double* ms_out, int64_t* launches_out) {
  if (!ctx) return VINSAT_EINVAL;
  timing_resolve(ctx);
  int n = 0;
  for (int i = 0; i < F_COUNT && n < cap; i++) {
    if (names_out) names_out[n] = kFamilyNames[i];
    if (ms_out) ms_out[n] = ctx->fam_ms[i];
    if (launches_out) launches_out[n] = ctx->fam_launches[i];
    n++;
  }
  return n;
}

int64_t vinsat_ctx_launch_count(const vinsat_ctx* ctx) { return ctx ? ctx->launches : 0; }

int vinsat_measure_fp64_peak(vinsat_ctx* ctx, double* tflops_out) {
  VS_CHECK_ARG(ctx, ctx && tflops_out);
  VS_CUDA(ctx, cudaSetDevice(ctx->device));
  double* d = (double*)ctx_scratch(ctx, 64);
  if (!d) return set_error(ctx, VINSAT_ENOMEM, "scratch alloc failed");
  const int iters = 20000, threads = 256, blocks = ctx->sm_count * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0, ctx->stream);
    VS_LAUNCH(ctx, F_PEAK, k_dfma_peak, blocks, threads, 0, d, iters, 1.0);
    cudaEventRecord(e1, ctx->stream);
    VS_CUDA(ctx, cudaEventSynchronize(e1));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 8.0 * iters * (double)threads * blocks;
    double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tflops_out = best;
  return VINSAT_OK;
}

int vinsat_measure_copy_bw(vinsat_ctx* ctx, int64_t bytes, double* gbs_out) {
  VS_CHECK_ARG(ctx, ctx && gbs_out && bytes >= 1024);
  VS_CUDA(ctx, cudaSetDevice(ctx->device));
  DevBuf<char> a, b;
  VS_CUDA(ctx, a.alloc(bytes));
  VS_CUDA(ctx, b.alloc(bytes));
  VS_CUDA(ctx, cudaMemsetAsync(a.p, 1, bytes, ctx->stream));
  int64_t n = bytes / sizeof(double2);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 6; rep++) {
    cudaEventRecord(e0, ctx->stream);
    VS_LAUNCH(ctx, F_PEAK, k_copy, ctx->sm_count * 16, 256, 0, (const double2*)a.p, (double2*)b.p, n);
    cudaEventRecord(e1, ctx->stream);
    VS_CUDA(ctx, cudaEventSynchronize(e1));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double gbs = 2.0 * (double)n * sizeof(double2) / (ms * 1e-3) / 1e9;
    if (rep > 0 && gbs > best) best = gbs;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *gbs_out = best;
  return VINSAT_OK;
}

}  // extern "C"
