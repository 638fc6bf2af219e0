// Confusion-matrix histogram: cm[n][t'][p] (int64) from logits (argmax in-kernel) or a class map.
// Reference contract: SegMetric._compute_basics (seg_metrics.py:13-28) — see include/eeseg.h.
//
// HBM-bound integer work. Layout: logits are NCHW planes, lanes walk consecutive pixels of a plane
// (coalesced 128 B / 64 B requests), targets are int64 (256 B per warp request). Counters are
// privatised per warp in shared memory (32-bit), warp-aggregated with match.any so that the
// blocky label maps of segmentation data do not serialise on one bank, then flushed once per block
// with 64-bit global atomics.
#include "common.cuh"
#include "plane_stream.cuh"

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

// Streaming variant (the default for logits): class planes staged through shared memory with
// bulk-TMA copies (plane_stream.cuh), one persistent CTA per SM, several tiles in flight.
template <typename T, int TILE>
__global__ void __launch_bounds__(TILE, 2) cm_from_logits_stream_kernel(
    const T* __restrict__ logits, const int64_t* __restrict__ targets, int C, int64_t HW, int copies,
    unsigned long long* __restrict__ cm, const uint8_t* __restrict__ limit_logits,
    const uint8_t* __restrict__ limit_targets, int stages) {
  extern __shared__ __align__(128) uint8_t hs_smem[];
  constexpr int ES = (int)sizeof(T);
  constexpr int rb = ps::row_bytes(TILE, ES), rbt = ps::row_bytes(TILE, 8);
  const int stage_bytes = C * rb + rbt;
  uint64_t* full = reinterpret_cast<uint64_t*>(hs_smem + (size_t)stages * stage_bytes);
  uint64_t* done = full + 4;   // every thread has read its pixel out of a stage
  unsigned* hist = reinterpret_cast<unsigned*>(full + 8);
  constexpr int NW = TILE / 32;   // row r is owned by warp r % NW, lane r / NW
  const int my_row = (int)(threadIdx.x & 31) * NW + (int)(threadIdx.x >> 5);
  const int bins = (C + 1) * C;
  for (int i = threadIdx.x; i < copies * bins; i += TILE) hist[i] = 0;
  unsigned* h = hist + ((threadIdx.x >> 5) % copies) * bins;

  const int n = blockIdx.y;
  const uint8_t* base_b = reinterpret_cast<const uint8_t*>(logits + (int64_t)n * C * HW);
  const uint8_t* tg_b = reinterpret_cast<const uint8_t*>(targets + (int64_t)n * HW);
  const int num_tiles = (int)((HW + TILE - 1) / TILE);
  const int my_count = (int)blockIdx.x < num_tiles ? (num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  auto issue = [&](int k) {   // the owner of row r issues it (r < C: class plane, r == C: targets)
    const int r = my_row;
    if (r > C) return;
    const int s = k % stages;
    const int64_t p0 = ((int64_t)blockIdx.x + (int64_t)k * gridDim.x) * TILE;
    const int count = (int)min((int64_t)TILE, HW - p0);
    uint8_t* st = hs_smem + (size_t)s * stage_bytes;
    if (r < C) ps::issue_tile<ES>(st + (size_t)r * rb, rb, full + s, base_b + (int64_t)r * HW * ES, 0, 1, p0, count, limit_logits);
    else ps::issue_tile<8>(st + (size_t)C * rb, rbt, full + s, tg_b, 0, 1, p0, count, limit_targets);
  };
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      ps::mbar_init(full + s, C + 1);
      ps::mbar_init(done + s, TILE);
    }
    ps::fence_barrier_init();
  }
  __syncthreads();
  for (int k = 0; k < stages && k < my_count; ++k) issue(k);
  const uint32_t delta = (uint32_t)(((uint64_t)HW * ES) & 15);
  for (int k = 0; k < my_count; ++k) {
    const int s = k % stages;
    const int64_t p0 = ((int64_t)blockIdx.x + (int64_t)k * gridDim.x) * TILE;
    const int64_t p = p0 + threadIdx.x;
    const uint8_t* st = hs_smem + (size_t)s * stage_bytes;
    ps::mbar_wait(full + s, (uint32_t)(k / stages) & 1u);
    const uint32_t a0 = (uint32_t)((uintptr_t)(base_b + p0 * ES) & 15);
    const uint32_t t0 = (uint32_t)((uintptr_t)(tg_b + p0 * 8) & 15);
    const uint32_t stu = ps::smem_u32(st);
    constexpr int P = 16 / ES;   // the plane misalignment pattern repeats every P classes
    uint32_t rowbase[P];
#pragma unroll
    for (int r = 0; r < P; ++r) rowbase[r] = stu + ((a0 + (uint32_t)r * delta) & 15u) + threadIdx.x * ES;
    int key = -1;
    if (p < HW) {
      int64_t t;
      asm volatile("ld.shared.b64 %0, [%1];" : "=l"(t) : "r"(stu + (uint32_t)C * rb + t0 + threadIdx.x * 8));
      float best = -INFINITY;
      int arg = 0;
      for (int c0 = 0; c0 < C; c0 += P) {
#pragma unroll
        for (int r = 0; r < P; ++r) {
          const int c = c0 + r;
          if (c < C) {
            float v;
            if constexpr (ES == 4) {
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(rowbase[r] + (uint32_t)c * rb));
            } else {
              unsigned short u;
              asm volatile("ld.shared.u16 %0, [%1];" : "=h"(u) : "r"(rowbase[r] + (uint32_t)c * rb));
              v = __uint_as_float(((uint32_t)u) << 16);
            }
            // torch.argmax: first maximal index, NaN counts as the maximum
            const bool take = (v > best) || (v != v && best == best) || (c == 0);
            if (take) { best = v; arg = c; }
          }
        }
      }
      const int tt = (t >= 0 && t < C) ? (int)t : C;
      key = tt * C + arg;
    }
    ps::mbar_arrive(done + s);  // this thread has read its pixel
    if (my_row <= C && k + stages < my_count) {   // only the row owners wait before refilling the stage
      ps::mbar_wait(done + s, (uint32_t)(k / stages) & 1u);
      issue(k + stages);
    }
    hist_add(h, key);
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
template <int PIX>
__global__ void __launch_bounds__(256) exit_accumulate_kernel(
    const uint8_t* __restrict__ amax_all, const int64_t* __restrict__ targets, int32_t* __restrict__ exit_idx,
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
  const int64_t* tg = targets + (int64_t)n * HW;
  const int64_t step = (int64_t)gridDim.x * blockDim.x * PIX;
  for (int64_t p0 = (int64_t)blockIdx.x * blockDim.x * PIX + (threadIdx.x & ~31); p0 < HW; p0 += step) {
#pragma unroll
    for (int j = 0; j < PIX; ++j) {
      const int64_t pp = p0 + (threadIdx.x & 31) + (int64_t)j * blockDim.x;
      int key = -1;
      if (pp < HW) {
        const int a = (int)__ldg(pr + pp);
        const int64_t t = __ldg(tg + pp);
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
    {
      // streaming path: one persistent CTA per SM, bulk-TMA staged planes
      constexpr int kTile = 512;
      const int es = dtype == EESEG_F32 ? 4 : 2;
      const size_t stage_bytes = (size_t)C * ps::row_bytes(kTile, es) + ps::row_bytes(kTile, 8);
      int hcopies = copies < 4 ? copies : 4;
      const size_t fixed = 8 * sizeof(uint64_t) + (size_t)hcopies * bytes1;
      int stages = (int)((112 * 1024 - fixed) / stage_bytes);   // two CTAs per SM
      if (stages > 3) stages = 3;
      if (stages >= 2 && C + 1 <= kTile && stages <= 4) {
        const size_t smem = stages * stage_bytes + fixed;
        int64_t per_img = 2 * kNumSMs / N;
        if (per_img < 1) per_img = 1;
        const int64_t tiles = (HW + kTile - 1) / kTile;
        dim3 sgrid((unsigned)(per_img < tiles ? per_img : tiles), (unsigned)N);
        const uintptr_t end_l = ((uintptr_t)pred + (size_t)N * C * HW * es + 15) & ~(uintptr_t)15;
        const uintptr_t end_t = ((uintptr_t)(targets + (int64_t)N * HW) + 15) & ~(uintptr_t)15;
        if (dtype == EESEG_F32) {
          auto kern = cm_from_logits_stream_kernel<float, kTile>;
          EESEG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          kern<<<sgrid, kTile, smem, stream>>>((const float*)pred, targets, C, HW, hcopies, cmu,
                                               (const uint8_t*)end_l, (const uint8_t*)end_t, stages);
        } else {
          auto kern = cm_from_logits_stream_kernel<__nv_bfloat16, kTile>;
          EESEG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          kern<<<sgrid, kTile, smem, stream>>>((const __nv_bfloat16*)pred, targets, C, HW, hcopies, cmu,
                                               (const uint8_t*)end_l, (const uint8_t*)end_t, stages);
        }
        return check_launch("cm_from_logits_stream_kernel");
      }
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
  exit_accumulate_kernel<kPix><<<grid, 256, copies * bytes1, stream>>>(
      amax_all, targets, exit_idx, E, N, C, HW, copies, reinterpret_cast<unsigned long long*>(cm_acc),
      reinterpret_cast<unsigned long long*>(counts), pred);
  return check_launch("exit_accumulate_kernel");
}
