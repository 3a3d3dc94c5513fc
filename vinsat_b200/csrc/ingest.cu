// (f)2: the integer half of the ingest path as device kernels -- observation indexing of read_detections
// (od_pipe.py:214-247: unique frames, knot frames at multiples of 1000 s, obs -> frame index `ii`) and the re-indexing
// of remove_elems (:253-288: frames kept = frames with a surviving observation or knots, ii_new = ii_old - #dropped
// frames below).  Everything is int64 / uint8 arithmetic, bit-exact against the reference's own outputs
// (tests/golden/seq_*.npz).  The sizes are 1e4..1e6, so one 1024-thread CTA streams each array with a tiled scan.
#include "common.cuh"
#include "launch.h"

using namespace vs;

namespace {

#define VS_TRY(expr)            \
  do {                          \
    int _rc = (expr);           \
    if (_rc != VINSAT_OK) return _rc; \
  } while (0)

constexpr int kScanThreads = 1024;

// exclusive prefix sum of one value per thread across the CTA; returns the thread's offset and the CTA total
__device__ __forceinline__ int64_t block_exclusive_scan(int64_t v, int64_t& total) {
  __shared__ int64_t s_warp[32];
  __shared__ int64_t s_total;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int64_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int64_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) s_warp[warp] = x;
  __syncthreads();
  if (warp == 0) {
    int64_t w = s_warp[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += y;
    }
    s_warp[lane] = w;
    if (lane == 31) s_total = w;
  }
  __syncthreads();
  const int64_t off = (warp ? s_warp[warp - 1] : 0) + x - v;
  total = s_total;
  __syncthreads();
  return off;
}

// uid[k] = index of obs k's frame among the unique frames (frames sorted non-decreasing, as read_detections assumes,
// SURVEY B.1); uniq[j] = j-th unique frame as int64; *n_uniq.  flags[0] |= 1 when the input is not sorted.
__global__ void __launch_bounds__(kScanThreads) k_unique_frames(int64_t n, const double* __restrict__ frames,
                                                                int64_t* __restrict__ uid, int64_t* __restrict__ uniq,
                                                                int64_t* __restrict__ n_uniq, int32_t* __restrict__ flags) {
  int64_t carry = 0;
  for (int64_t base = 0; base < n; base += kScanThreads) {
    const int64_t k = base + threadIdx.x;
    int64_t head = 0;
    if (k < n) {
      head = (k == 0 || frames[k] != frames[k - 1]) ? 1 : 0;
      if (k > 0 && frames[k] < frames[k - 1]) atomicOr(flags, 1);
    }
    int64_t total;
    const int64_t off = block_exclusive_scan(head, total);
    if (k < n) {
      const int64_t j = carry + off + head - 1;
      uid[k] = j;
      if (head) uniq[j] = (int64_t)frames[k];          // np.unique(...).astype(np.int64): truncation
    }
    carry += total;
  }
  if (threadIdx.x == 0) *n_uniq = carry;
}

// The knot walk of od_pipe.py:216-226,242-245 over the unique frames.  It carries one counter (`filler_idx`) whose
// update at frame j depends on the update at frame j-1, and m <= 1e5: one thread walks it (~10 cycles per frame).
// slot[j] = position of unique frame j in the new time_idx (= j + knots inserted before it).
__global__ void k_knot_walk(const int64_t* __restrict__ uniq, const int64_t* __restrict__ n_uniq, int64_t n_orbit,
                            int64_t cap, int64_t* __restrict__ time_idx, int64_t* __restrict__ slot,
                            int64_t* __restrict__ n_frames, int32_t* __restrict__ flags) {
  if (threadIdx.x || blockIdx.x) return;
  const int64_t m = *n_uniq;
  if (m == 0) { *n_frames = 0; return; }
  // python floor division for the first frame (frames are >= 0 in practice)
  int64_t filler = (uniq[0] >= 0 ? uniq[0] / 1000 : -((-uniq[0] + 999) / 1000)) + 1;
  int64_t out = 0;
  bool overflow = false;
  auto push = [&](int64_t v) { if (out < cap) time_idx[out] = v; else overflow = true; out++; };
  for (int64_t j = 0; j < m; j++) {
    const int64_t t = uniq[j];
    if (t == filler * 1000) filler++;                    // a detection exactly on a knot: the counter just advances (:220-221)
    while (t > filler * 1000) { push(filler * 1000); filler++; }
    slot[j] = out;
    push(t);
  }
  if (uniq[m - 1] < n_orbit)                             // trailing knots up to (n_orbit // 1000) * 1000 (:242-245)
    while (filler * 1000 < (n_orbit / 1000) * 1000 + 1) { push(filler * 1000); filler++; }
  *n_frames = out;
  if (overflow) atomicOr(flags, 2);
}

__global__ void __launch_bounds__(256) k_gather_slots(int64_t n, const int64_t* __restrict__ uid,
                                                      const int64_t* __restrict__ slot, int64_t* __restrict__ ii) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k < n) ii[k] = slot[uid[k]];
}

// ---- remove_elems ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_mark_frames(int64_t n, int64_t T, const uint8_t* __restrict__ mask,
                                                     const int64_t* __restrict__ ii, int32_t* __restrict__ has_obs,
                                                     unsigned long long* __restrict__ ii_max, int32_t* __restrict__ flags) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= n || !mask[k]) return;
  const int64_t f = ii[k];
  if (f < 0 || f >= T) { atomicOr(flags, 4); return; }
  has_obs[f] = 1;
  atomicMax(ii_max, (unsigned long long)(f + 1));        // stores max+1 so that 0 means "no surviving observation"
}

// keep[f] = has_obs[f] || time_idx[f] % 1000 == 0 (:259-263); shift[f] = #dropped frames strictly below f, counting
// only frames <= max surviving ii (the reference's loop stops there, :272); new_pos[f] = #kept frames below f
__global__ void __launch_bounds__(kScanThreads) k_frame_scan(int64_t T, const int32_t* __restrict__ has_obs,
                                                             const int64_t* __restrict__ time_idx,
                                                             const unsigned long long* __restrict__ ii_max,
                                                             uint8_t* __restrict__ keep, int64_t* __restrict__ shift,
                                                             int64_t* __restrict__ time_idx_out, int64_t* __restrict__ n_kept) {
  const int64_t last = (int64_t)*ii_max - 1;
  int64_t carry_d = 0, carry_k = 0;
  for (int64_t base = 0; base < T; base += kScanThreads) {
    const int64_t f = base + threadIdx.x;
    int64_t kp = 0, dr = 0;
    if (f < T) {
      const int64_t t = time_idx[f];
      kp = (has_obs[f] || (((t % 1000) + 1000) % 1000 == 0)) ? 1 : 0;      // python modulo
      dr = (!kp && f <= last) ? 1 : 0;
    }
    int64_t td, tk;
    const int64_t od = block_exclusive_scan(dr, td);
    const int64_t ok = block_exclusive_scan(kp, tk);
    if (f < T) {
      keep[f] = (uint8_t)kp;
      shift[f] = carry_d + od;
      if (kp) time_idx_out[carry_k + ok] = time_idx[f];
    }
    carry_d += td;
    carry_k += tk;
  }
  if (threadIdx.x == 0) *n_kept = carry_k;
}

__global__ void __launch_bounds__(kScanThreads) k_obs_compact(int64_t n, const uint8_t* __restrict__ mask,
                                                              const int64_t* __restrict__ ii,
                                                              const int64_t* __restrict__ shift,
                                                              int64_t* __restrict__ ii_out, int64_t* __restrict__ n_kept) {
  int64_t carry = 0;
  for (int64_t base = 0; base < n; base += kScanThreads) {
    const int64_t k = base + threadIdx.x;
    const int64_t m = (k < n && mask[k]) ? 1 : 0;
    int64_t total;
    const int64_t off = block_exclusive_scan(m, total);
    if (m) ii_out[carry + off] = ii[k] - shift[ii[k]];
    carry += total;
  }
  if (threadIdx.x == 0) *n_kept = carry;
}

template <typename T>
struct Buf {
  DevBuf<T> own;
  T* p = nullptr;
  T* host = nullptr;
  int64_t n = 0;
  int in(vinsat_ctx* ctx, int mem, const T* src, int64_t n_) {
    n = n_;
    if (mem == VINSAT_MEM_DEVICE) { p = const_cast<T*>(src); return VINSAT_OK; }
    if (own.alloc(n) != cudaSuccess) { cudaGetLastError(); return set_error(ctx, VINSAT_ENOMEM, "cudaMalloc failed"); }
    p = own.p;
    if (n) VS_CUDA(ctx, cudaMemcpyAsync(p, src, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return VINSAT_OK;
  }
  int out(vinsat_ctx* ctx, int mem, T* dst, int64_t n_) {
    n = n_;
    if (mem == VINSAT_MEM_DEVICE && dst) { p = dst; return VINSAT_OK; }
    if (own.alloc(n) != cudaSuccess) { cudaGetLastError(); return set_error(ctx, VINSAT_ENOMEM, "cudaMalloc failed"); }
    p = own.p;
    host = (mem == VINSAT_MEM_DEVICE) ? nullptr : dst;
    return VINSAT_OK;
  }
  int back(vinsat_ctx* ctx, int64_t count) {
    if (host && count > 0)
      VS_CUDA(ctx, cudaMemcpyAsync(host, p, count * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    return VINSAT_OK;
  }
};

}  // namespace

extern "C" {

int vinsat_index_detections(vinsat_ctx* ctx, int mem, int64_t n_det, const double* frames, int64_t n_orbit,
                            int64_t cap_frames, int64_t* time_idx_out, int64_t* ii_out, int64_t* n_frames_out) {
  VS_CHECK_ARG(ctx, ctx != nullptr);
  VS_CHECK_ARG(ctx, n_det >= 1 && frames && n_orbit >= 0 && cap_frames >= 1 && time_idx_out && ii_out && n_frames_out);
  VS_CUDA(ctx, cudaSetDevice(ctx->device));
  Buf<double> fr;
  Buf<int64_t> uid, uniq, slot, tix, ii, cnt;
  DevBuf<int32_t> flags;
  VS_TRY(fr.in(ctx, mem, frames, n_det));
  VS_TRY(uid.out(ctx, VINSAT_MEM_DEVICE, nullptr, n_det));
  VS_TRY(uniq.out(ctx, VINSAT_MEM_DEVICE, nullptr, n_det));
  VS_TRY(slot.out(ctx, VINSAT_MEM_DEVICE, nullptr, n_det));
  VS_TRY(cnt.out(ctx, VINSAT_MEM_DEVICE, nullptr, 2));
  VS_TRY(tix.out(ctx, mem, time_idx_out, cap_frames));
  VS_TRY(ii.out(ctx, mem, ii_out, n_det));
  VS_CUDA(ctx, flags.alloc(1));
  VS_CUDA(ctx, cudaMemsetAsync(flags.p, 0, sizeof(int32_t), ctx->stream));
  VS_LAUNCH(ctx, F_LAYOUT, k_unique_frames, 1, kScanThreads, 0, n_det, fr.p, uid.p, uniq.p, cnt.p, flags.p);
  VS_LAUNCH(ctx, F_LAYOUT, k_knot_walk, 1, 32, 0, uniq.p, cnt.p, n_orbit, cap_frames, tix.p, slot.p, cnt.p + 1, flags.p);
  VS_LAUNCH(ctx, F_LAYOUT, k_gather_slots, ceil_div(n_det, 256), 256, 0, n_det, uid.p, slot.p, ii.p);
  int64_t h_cnt[2] = {0, 0};
  int32_t h_flags = 0;
  VS_CUDA(ctx, cudaMemcpyAsync(h_cnt, cnt.p, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  VS_CUDA(ctx, cudaMemcpyAsync(&h_flags, flags.p, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (h_flags & 1) return set_error(ctx, VINSAT_EINVAL, "detections must be sorted by frame (read_detections assumes it)");
  if (h_flags & 2) return set_error(ctx, VINSAT_EINVAL, "cap_frames = %lld too small for %lld frames", (long long)cap_frames,
                                    (long long)h_cnt[1]);
  *n_frames_out = h_cnt[1];
  VS_TRY(tix.back(ctx, h_cnt[1]));
  VS_TRY(ii.back(ctx, n_det));
  VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return VINSAT_OK;
}

int vinsat_remove_elems_index(vinsat_ctx* ctx, int mem, int64_t n_det, int64_t n_frames, const uint8_t* mask,
                              const int64_t* ii, const int64_t* time_idx, int64_t* ii_out, int64_t* time_idx_out,
                              uint8_t* frame_keep_out, int64_t* counts_out) {
  VS_CHECK_ARG(ctx, ctx != nullptr);
  VS_CHECK_ARG(ctx, n_det >= 0 && n_frames >= 1 && time_idx && ii_out && time_idx_out && frame_keep_out && counts_out);
  VS_CHECK_ARG(ctx, n_det == 0 || (mask && ii));
  VS_CUDA(ctx, cudaSetDevice(ctx->device));
  Buf<uint8_t> mk, keep;
  Buf<int64_t> iin, tin, shift, iio, tio, cnt;
  DevBuf<int32_t> has_obs, flags;
  DevBuf<unsigned long long> iimax;
  VS_TRY(mk.in(ctx, mem, mask, n_det));
  VS_TRY(iin.in(ctx, mem, ii, n_det));
  VS_TRY(tin.in(ctx, mem, time_idx, n_frames));
  VS_TRY(shift.out(ctx, VINSAT_MEM_DEVICE, nullptr, n_frames));
  VS_TRY(cnt.out(ctx, VINSAT_MEM_DEVICE, nullptr, 2));
  VS_TRY(iio.out(ctx, mem, ii_out, std::max<int64_t>(n_det, 1)));
  VS_TRY(tio.out(ctx, mem, time_idx_out, n_frames));
  VS_TRY(keep.out(ctx, mem, frame_keep_out, n_frames));
  VS_CUDA(ctx, has_obs.alloc(n_frames));
  VS_CUDA(ctx, flags.alloc(1));
  VS_CUDA(ctx, iimax.alloc(1));
  VS_CUDA(ctx, cudaMemsetAsync(has_obs.p, 0, n_frames * sizeof(int32_t), ctx->stream));
  VS_CUDA(ctx, cudaMemsetAsync(flags.p, 0, sizeof(int32_t), ctx->stream));
  VS_CUDA(ctx, cudaMemsetAsync(iimax.p, 0, sizeof(unsigned long long), ctx->stream));
  VS_CUDA(ctx, cudaMemsetAsync(cnt.p, 0, 2 * sizeof(int64_t), ctx->stream));
  if (n_det > 0)
    VS_LAUNCH(ctx, F_LAYOUT, k_mark_frames, ceil_div(n_det, 256), 256, 0, n_det, n_frames, mk.p, iin.p, has_obs.p, iimax.p,
              flags.p);
  VS_LAUNCH(ctx, F_LAYOUT, k_frame_scan, 1, kScanThreads, 0, n_frames, has_obs.p, tin.p, iimax.p, keep.p, shift.p, tio.p,
            cnt.p + 1);
  if (n_det > 0)
    VS_LAUNCH(ctx, F_LAYOUT, k_obs_compact, 1, kScanThreads, 0, n_det, mk.p, iin.p, shift.p, iio.p, cnt.p);
  int64_t h_cnt[2] = {0, 0};
  int32_t h_flags = 0;
  VS_CUDA(ctx, cudaMemcpyAsync(h_cnt, cnt.p, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  VS_CUDA(ctx, cudaMemcpyAsync(&h_flags, flags.p, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (h_flags & 4) return set_error(ctx, VINSAT_EINVAL, "ii out of range [0, n_frames)");
  counts_out[0] = h_cnt[0];
  counts_out[1] = h_cnt[1];
  VS_TRY(iio.back(ctx, h_cnt[0]));
  VS_TRY(tio.back(ctx, h_cnt[1]));
  VS_TRY(keep.back(ctx, n_frames));
  VS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return VINSAT_OK;
}

}  // extern "C"
