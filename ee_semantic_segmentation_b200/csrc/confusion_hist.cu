// Confusion-matrix histogram: cm[n][t'][p] (int64) from logits (argmax in-kernel) or a class map.
// Reference contract: SegMetric._compute_basics (seg_metrics.py:13-28) — see include/eeseg.h.
//
// HBM-bound integer work. Layout: logits are NCHW planes, lanes walk consecutive pixels of a plane
// (coalesced 128 B / 64 B requests), targets are int64 (256 B per warp request). Counters are
// privatised per warp in shared memory (32-bit), warp-aggregated with match.any so that the
// blocky label maps of segmentation data do not serialise on one bank, then flushed once per block
// with 64-bit global atomics.
#include "common.cuh"

namespace eeseg {

__device__ __forceinline__ void hist_add(unsigned* h, int key) {
  // key < 0: lane has no pixel. Lanes with equal keys elect one leader that adds the group size.
  unsigned peers = __match_any_sync(0xffffffffu, key);
  if (key >= 0) {
    int leader = __ffs(peers) - 1;
    if ((int)(threadIdx.x & 31) == leader) atomicAdd(h + key, (unsigned)__popc(peers));
  }
}

__device__ __forceinline__ void hist_flush(const unsigned* hist, int copies, int bins,
                                           unsigned long long* cm_n) {
  __syncthreads();
  for (int i = threadIdx.x; i < bins; i += blockDim.x) {
    unsigned long long s = 0;
    for (int k = 0; k < copies; ++k) s += hist[k * bins + i];
    if (s) atomicAdd(cm_n + i, s);
  }
}

template <typename T, int PIX>
__global__ void __launch_bounds__(256) cm_from_logits_kernel(const T* __restrict__ logits,
                                                              const int64_t* __restrict__ targets,
                                                              int C, int64_t HW, int copies,
                                                              unsigned long long* __restrict__ cm) {
  extern __shared__ unsigned hist[];
  const int bins = (C + 1) * C;
  for (int i = threadIdx.x; i < copies * bins; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  unsigned* h = hist + ((threadIdx.x >> 5) % copies) * bins;
  const int n = blockIdx.y;
  const T* base = logits + (int64_t)n * C * HW;
  const int64_t* tg = targets + (int64_t)n * HW;
  const int64_t step = (int64_t)gridDim.x * blockDim.x * PIX;
  // all lanes of a warp iterate together (the loop bound is rounded to the warp) for match.any
  for (int64_t p0 = (int64_t)blockIdx.x * blockDim.x * PIX + (threadIdx.x & ~31); p0 < HW;
       p0 += step) {
    float best[PIX];
    int arg[PIX];
    int64_t pp[PIX];
#pragma unroll
    for (int j = 0; j < PIX; ++j) {
      pp[j] = p0 + (threadIdx.x & 31) + (int64_t)j * blockDim.x;
      best[j] = -INFINITY;
      arg[j] = 0;
    }
    int64_t t[PIX];
#pragma unroll
    for (int j = 0; j < PIX; ++j) t[j] = pp[j] < HW ? __ldg(tg + pp[j]) : -1;
#pragma unroll 7
    for (int c = 0; c < C; ++c) {
#pragma unroll
      for (int j = 0; j < PIX; ++j) {
        float v = pp[j] < HW ? ldf_stream(base + (int64_t)c * HW + pp[j]) : 0.f;
        // torch.argmax: first maximal index, NaN counts as the maximum
        bool take = (v > best[j]) || (v != v && best[j] == best[j]) || (c == 0);
        if (take) { best[j] = v; arg[j] = c; }
      }
    }
#pragma unroll
    for (int j = 0; j < PIX; ++j) {
      int tt = (t[j] >= 0 && t[j] < C) ? (int)t[j] : C;
      hist_add(h, pp[j] < HW ? tt * C + arg[j] : -1);
    }
  }
  hist_flush(hist, copies, bins, cm + (int64_t)n * bins);
}

// torch.argmax treats NaN as the maximum and returns the first one. Pixels with a NaN (or inf - inf)
// logit are redone here, out of line and re-reading the pixel, so that the hot loop carries no NaN
// bookkeeping beyond one add per class.
template <typename T>
__device__ __noinline__ int argmax_with_nan(const T* px, int C, uint32_t plane_bytes) {
  float best = ldf(px);
  int arg = 0;
  if (best != best) return 0;
  for (int c = 1; c < C; ++c) {
    const float v = ldf(plane_ptr(px, (uint32_t)c, plane_bytes));
    if (v != v) return c;
    if (v > best) { best = v; arg = c; }
  }
  return arg;
}

// Default for logits: thread = PIX pixels (strided by the block), all C class values of those pixels
// loaded into registers before the first compare (PIX x C independent plane-coalesced loads in flight
// per thread, no staging). tools/ub_stream.cu: 79 % of the HBM copy peak with PIX = 2, against 44 %
// for the bulk-TMA shared-memory ring it replaces and 71 % / 60 % for PIX = 1 / 4.
// CMAX 32 = register capacity with a runtime class count, otherwise the exact count (19, 21).
template <typename T, int CMAX, int PIX, int THREADS>
__global__ void __launch_bounds__(THREADS) cm_from_logits_direct_kernel(
    const T* __restrict__ logits, const int64_t* __restrict__ targets, int C, int64_t HW, int copies,
    unsigned long long* __restrict__ cm) {
  if (CMAX != 32) C = CMAX;
  extern __shared__ unsigned hist[];
  const int bins = (C + 1) * C;
  for (int i = threadIdx.x; i < copies * bins; i += THREADS) hist[i] = 0;
  unsigned* h = hist + ((threadIdx.x >> 5) % copies) * bins;
  const int n = blockIdx.y;
  const T* base = logits + (int64_t)n * C * HW;
  const int64_t p0 = (int64_t)blockIdx.x * (THREADS * PIX) + threadIdx.x;
  const uint32_t pb = (uint32_t)HW * (uint32_t)sizeof(T);   // plane stride in bytes (< 2^32, checked on the host)
  float v[PIX][CMAX];
  int64_t t[PIX];
#pragma unroll
  for (int j = 0; j < PIX; ++j) {
    const int64_t p = p0 + j * THREADS;
    t[j] = p < HW ? __ldg(targets + (int64_t)n * HW + p) : -1;
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) v[j][c] = p < HW ? ldf_stream(plane_ptr(base + p, (uint32_t)c, pb)) : 0.f;
  }
  __syncthreads();   // histogram zeroed (the loads above are already in flight)
#pragma unroll
  for (int j = 0; j < PIX; ++j) {
    const int64_t p = p0 + j * THREADS;
    // torch.argmax: first maximal index, NaN counts as the maximum. Hot loop: one compare + two selects
    // per class; the running sum (FMA pipe, off the compare/select pipe) turns NaN if any logit is NaN
    // (or for inf - inf) and sends that rare pixel to the exact out-of-line path.
    float best = v[j][0], nan_probe = v[j][0];
    int arg = 0;
#pragma unroll
    for (int c = 1; c < CMAX; ++c)
      if (c < C) {
        const bool take = v[j][c] > best;
        best = take ? v[j][c] : best;
        arg = take ? c : arg;
        nan_probe += v[j][c];
      }
    if (nan_probe != nan_probe && p < HW) arg = argmax_with_nan(base + p, C, pb);
    const int tt = (t[j] >= 0 && t[j] < C) ? (int)t[j] : C;
    hist_add(h, p < HW ? tt * C + arg : -1);   // whole warps reach this together (match.any)
  }
  hist_flush(hist, copies, bins, cm + (int64_t)n * bins);
}

template <typename P, int PIX>
__global__ void __launch_bounds__(256) cm_from_map_kernel(const P* __restrict__ pred,
                                                           const int64_t* __restrict__ targets,
                                                           int C, int64_t HW, int copies,
                                                           unsigned long long* __restrict__ cm) {
  extern __shared__ unsigned hist[];
  const int bins = (C + 1) * C;
  for (int i = threadIdx.x; i < copies * bins; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  unsigned* h = hist + ((threadIdx.x >> 5) % copies) * bins;
  const int n = blockIdx.y;
  const P* pr = pred + (int64_t)n * HW;
  const int64_t* tg = targets + (int64_t)n * HW;
  const int64_t step = (int64_t)gridDim.x * blockDim.x * PIX;
  for (int64_t p0 = (int64_t)blockIdx.x * blockDim.x * PIX + (threadIdx.x & ~31); p0 < HW;
       p0 += step) {
    int64_t t[PIX];
    int a[PIX];
#pragma unroll
    for (int j = 0; j < PIX; ++j) {
      int64_t pp = p0 + (threadIdx.x & 31) + (int64_t)j * blockDim.x;
      bool ok = pp < HW;
      t[j] = ok ? __ldg(tg + pp) : -1;
      a[j] = ok ? (int)__ldg(pr + pp) : -1;
    }
#pragma unroll
    for (int j = 0; j < PIX; ++j) {
      int key = -1;
      if (a[j] >= 0 && a[j] < C) {  // a class map value outside [0,C) is not counted
        int tt = (t[j] >= 0 && t[j] < C) ? (int)t[j] : C;
        key = tt * C + a[j];
      }
      hist_add(h, key);
    }
  }
  hist_flush(hist, copies, bins, cm + (int64_t)n * bins);
}

// End-of-batch accumulation of the early-exit evaluators (eval_br_ent.py:57-70): for image n the
// argmax map of the exit it took (amax_all[exit][n]) is histogrammed against the target and added to
// the accumulators of that exit and of the global slot E; exit counters are bumped; still-active
// images (exit_idx < 0) are assigned the final exit. One launch instead of a dozen tiny library ops.
template <int PIX, typename TG>
__global__ void __launch_bounds__(256) exit_accumulate_kernel(
    const uint8_t* __restrict__ amax_all, const TG* __restrict__ targets, int32_t* __restrict__ exit_idx,
    int E, int N, int C, int64_t HW, int copies, unsigned long long* __restrict__ cm_acc,
    unsigned long long* __restrict__ counts, uint8_t* __restrict__ pred) {
  extern __shared__ unsigned hist[];
  const int bins = (C + 1) * C;
  for (int i = threadIdx.x; i < copies * bins; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  unsigned* h = hist + ((threadIdx.x >> 5) % copies) * bins;
  const int n = blockIdx.y;
  int e = exit_idx[n];
  if (e < 0 || e >= E) e = E - 1;
  const uint8_t* pr = amax_all + ((int64_t)e * N + n) * HW;
  const TG* tg = targets + (int64_t)n * HW;
  const int64_t step = (int64_t)gridDim.x * blockDim.x * PIX;
  for (int64_t p0 = (int64_t)blockIdx.x * blockDim.x * PIX + (threadIdx.x & ~31); p0 < HW; p0 += step) {
#pragma unroll
    for (int j = 0; j < PIX; ++j) {
      const int64_t pp = p0 + (threadIdx.x & 31) + (int64_t)j * blockDim.x;
      int key = -1;
      if (pp < HW) {
        const int a = (int)__ldg(pr + pp);
        const int64_t t = (int64_t)__ldg(tg + pp);     // uint8 labels widen here: 0..C-1 classes, >= C void
        if (pred) pred[(int64_t)n * HW + pp] = (uint8_t)a;
        if (a < C) key = ((t >= 0 && t < C) ? (int)t : C) * C + a;
      }
      hist_add(h, key);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < bins; i += blockDim.x) {
    unsigned long long sum = 0;
    for (int k = 0; k < copies; ++k) sum += hist[k * bins + i];
    if (sum) {
      atomicAdd(cm_acc + (int64_t)e * bins + i, sum);
      atomicAdd(cm_acc + (int64_t)E * bins + i, sum);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    exit_idx[n] = e;
    atomicAdd(counts + e, 1ull);
    atomicAdd(counts + E, 1ull);
  }
}

// Fallback for very large C (histogram does not fit shared memory): global atomics.
template <typename T>
__global__ void cm_from_logits_global_kernel(const T* __restrict__ logits,
                                             const int64_t* __restrict__ targets, int C,
                                             int64_t HW, unsigned long long* __restrict__ cm) {
  const int n = blockIdx.y;
  const T* base = logits + (int64_t)n * C * HW;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < HW;
       p += (int64_t)gridDim.x * blockDim.x) {
    float best = ldf_stream(base + p);
    int arg = 0;
    for (int c = 1; c < C; ++c) {
      float v = ldf_stream(base + (int64_t)c * HW + p);
      if ((v > best) || (v != v && best == best)) { best = v; arg = c; }
    }
    int64_t t = targets[(int64_t)n * HW + p];
    int tt = (t >= 0 && t < C) ? (int)t : C;
    atomicAdd(cm + ((int64_t)n * (C + 1) + tt) * C + arg, 1ull);
  }
}

}  // namespace eeseg

using namespace eeseg;

extern "C" int eeseg_confusion_hist(const void* pred, int pred_kind, int dtype,
                                    const int64_t* targets, int N, int C, int64_t HW, int64_t* cm,
                                    int accumulate, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  EESEG_REQUIRE(pred && targets && cm, "confusion_hist: null pointer");
  EESEG_REQUIRE(N >= 0 && C >= 1 && HW >= 0, "confusion_hist: bad sizes N=%d C=%d HW=%lld", N, C,
                (long long)HW);
  EESEG_REQUIRE(pred_kind >= 0 && pred_kind <= 2, "confusion_hist: pred_kind %d", pred_kind);
  EESEG_REQUIRE(pred_kind != 1 || C <= 256, "confusion_hist: uint8 map needs C <= 256");
  const int64_t bins = (int64_t)(C + 1) * C;
  if (!accumulate) EESEG_CUDA(cudaMemsetAsync(cm, 0, sizeof(int64_t) * bins * (N > 0 ? N : 0), stream));
  if (N == 0 || HW == 0) return EESEG_OK;
  auto* cmu = reinterpret_cast<unsigned long long*>(cm);
  constexpr int kThreads = 256, kPix = 2;
  const size_t bytes1 = (size_t)bins * sizeof(unsigned);
  int copies = (int)((48 * 1024) / bytes1);
  if (copies > kThreads / 32) copies = kThreads / 32;
  int64_t want = (HW + kThreads * kPix - 1) / (kThreads * kPix);
  int64_t cap = (kNumSMs * 8 + N - 1) / N;
  dim3 grid((unsigned)(want < cap ? want : cap), (unsigned)N);
  if (pred_kind == 0) {
    EESEG_REQUIRE(dtype == EESEG_F32 || dtype == EESEG_BF16, "confusion_hist: dtype %d", dtype);
    if (copies == 0) {
      dim3 g2((unsigned)((HW + 255) / 256 < cap ? (HW + 255) / 256 : cap), (unsigned)N);
      if (dtype == EESEG_F32)
        cm_from_logits_global_kernel<float><<<g2, 256, 0, stream>>>((const float*)pred, targets, C, HW, cmu);
      else
        cm_from_logits_global_kernel<__nv_bfloat16><<<g2, 256, 0, stream>>>((const __nv_bfloat16*)pred, targets, C, HW, cmu);
      return check_launch("cm_from_logits_global_kernel");
    }
    if (C <= 32 && HW < (1ll << 29)) {
      constexpr int kT = 256, kP = 2;
      int dcopies = copies < kT / 32 ? copies : kT / 32;
      dim3 dgrid((unsigned)((HW + kT * kP - 1) / (kT * kP)), (unsigned)N);
      const size_t sm = (size_t)dcopies * bytes1;
#define EESEG_HIST_CASE(T, CM)                                                                             \
  cm_from_logits_direct_kernel<T, CM, kP, kT><<<dgrid, kT, sm, stream>>>((const T*)pred, targets, C, HW, dcopies, cmu)
      if (dtype == EESEG_F32) {
        if (C == 21) EESEG_HIST_CASE(float, 21);
        else if (C == 19) EESEG_HIST_CASE(float, 19);
        else EESEG_HIST_CASE(float, 32);
      } else {
        if (C == 21) EESEG_HIST_CASE(__nv_bfloat16, 21);
        else if (C == 19) EESEG_HIST_CASE(__nv_bfloat16, 19);
        else EESEG_HIST_CASE(__nv_bfloat16, 32);
      }
#undef EESEG_HIST_CASE
      return check_launch("cm_from_logits_direct_kernel");
    }
    if (dtype == EESEG_F32)
      cm_from_logits_kernel<float, kPix><<<grid, kThreads, copies * bytes1, stream>>>(
          (const float*)pred, targets, C, HW, copies, cmu);
    else
      cm_from_logits_kernel<__nv_bfloat16, kPix><<<grid, kThreads, copies * bytes1, stream>>>(
          (const __nv_bfloat16*)pred, targets, C, HW, copies, cmu);
    return check_launch("cm_from_logits_kernel");
  }
  EESEG_REQUIRE(copies > 0, "confusion_hist: C=%d too large for the class-map path", C);
  constexpr int kPixMap = 4;
  want = (HW + kThreads * kPixMap - 1) / (kThreads * kPixMap);
  grid.x = (unsigned)(want < cap ? want : cap);
  if (pred_kind == 1)
    cm_from_map_kernel<uint8_t, kPixMap><<<grid, kThreads, copies * bytes1, stream>>>(
        (const uint8_t*)pred, targets, C, HW, copies, cmu);
  else
    cm_from_map_kernel<int64_t, kPixMap><<<grid, kThreads, copies * bytes1, stream>>>(
        (const int64_t*)pred, targets, C, HW, copies, cmu);
  return check_launch("cm_from_map_kernel");
}

extern "C" int eeseg_exit_accumulate_u8(const uint8_t* amax_all, const uint8_t* targets, int32_t* exit_idx, int E,
                                        int N, int C, int64_t HW, int64_t* cm_acc, int64_t* counts, uint8_t* pred,
                                        void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  EESEG_REQUIRE(amax_all && targets && exit_idx && cm_acc && counts, "exit_accumulate: null pointer");
  EESEG_REQUIRE(E >= 1 && C >= 1 && C <= 255 && HW >= 1, "exit_accumulate: bad sizes (uint8 labels: C <= 255)");
  if (N <= 0) return EESEG_OK;
  const size_t bytes1 = (size_t)(C + 1) * C * sizeof(unsigned);
  int copies = (int)((48 * 1024) / bytes1);
  EESEG_REQUIRE(copies > 0, "exit_accumulate: C=%d too large", C);
  if (copies > 8) copies = 8;
  constexpr int kPix = 4;
  int64_t want = (HW + 256 * kPix - 1) / (256 * kPix);
  int64_t cap = (kNumSMs * 8 + N - 1) / N;
  dim3 grid((unsigned)(want < cap ? want : cap), (unsigned)N);
  exit_accumulate_kernel<kPix, uint8_t><<<grid, 256, copies * bytes1, stream>>>(
      amax_all, targets, exit_idx, E, N, C, HW, copies, reinterpret_cast<unsigned long long*>(cm_acc),
      reinterpret_cast<unsigned long long*>(counts), pred);
  return check_launch("exit_accumulate_kernel");
}

extern "C" int eeseg_exit_accumulate(const uint8_t* amax_all, const int64_t* targets, int32_t* exit_idx, int E,
                                     int N, int C, int64_t HW, int64_t* cm_acc, int64_t* counts, uint8_t* pred,
                                     void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  EESEG_REQUIRE(amax_all && targets && exit_idx && cm_acc && counts, "exit_accumulate: null pointer");
  EESEG_REQUIRE(E >= 1 && C >= 1 && C <= 256 && HW >= 1, "exit_accumulate: bad sizes");
  if (N <= 0) return EESEG_OK;
  const size_t bytes1 = (size_t)(C + 1) * C * sizeof(unsigned);
  int copies = (int)((48 * 1024) / bytes1);
  EESEG_REQUIRE(copies > 0, "exit_accumulate: C=%d too large", C);
  if (copies > 8) copies = 8;
  constexpr int kPix = 4;
  int64_t want = (HW + 256 * kPix - 1) / (256 * kPix);
  int64_t cap = (kNumSMs * 8 + N - 1) / N;
  dim3 grid((unsigned)(want < cap ? want : cap), (unsigned)N);
  exit_accumulate_kernel<kPix, int64_t><<<grid, 256, copies * bytes1, stream>>>(
      amax_all, targets, exit_idx, E, N, C, HW, copies, reinterpret_cast<unsigned long long*>(cm_acc),
      reinterpret_cast<unsigned long long*>(counts), pred);
  return check_launch("exit_accumulate_kernel");
}
