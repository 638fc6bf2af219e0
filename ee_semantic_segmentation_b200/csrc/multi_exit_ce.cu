// Weighted multi-exit pixelwise cross-entropy, forward (+ fused gradient) and backward.
// Reference contract: BrXEntropyLoss (my_pixelwise_xentropy.py:19-46) over
// torch.nn.CrossEntropyLoss(reduction='mean', ignore_index) per exit — see include/eeseg.h.
//
// HBM-bound: one pass over logits [E][N][C][HW] (read) and, when the gradient is wanted, one write
// of the same size — the reference does E separate log_softmax + nll passes forward and E backward.
// One thread per pixel (lanes = consecutive pixels of a class plane -> coalesced requests), the C
// class values live in registers, per-block loss partials are written to a scratch array and
// reduced in a fixed order (bit-reproducible loss).
//
// v2: the class planes are staged through shared memory with bulk-TMA copies (plane_stream.cuh),
// 3-4 tiles in flight per CTA, one persistent CTA per SM; a thread still owns one pixel and keeps the
// C values in registers while computing, and writes its gradient with coalesced plane stores.
#include "common.cuh"
#include "plane_stream.cuh"

namespace eeseg {


// two persistent CTAs per SM, split evenly over the E*N (exit, image) pairs
static inline int ce_grid_x(int E, int N, int64_t HW) {
  int64_t per_pair = 2 * kNumSMs / ((int64_t)E * N);
  if (per_pair < 1) per_pair = 1;
  int64_t tiles = (HW + 255) / 256;
  return (int)(per_pair < tiles ? per_pair : tiles);
}

__global__ void count_valid_kernel(const int64_t* __restrict__ targets, int64_t total, int C,
                                   int64_t ignore, unsigned long long* __restrict__ out) {
  int cnt = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = __ldg(targets + i);
    cnt += (t != ignore && t >= 0 && t < C) ? 1 : 0;
  }
  cnt = warp_sum(cnt);
  __shared__ int s[32];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s[i];
    if (t) atomicAdd(out, t);
  }
}

__device__ __forceinline__ float ce_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ce_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <typename T>
__device__ __forceinline__ void sts_f(uint32_t addr, float v) {
  if constexpr (sizeof(T) == 4) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
  } else {
    __nv_bfloat16 b = __float2bfloat16_rn(v);
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(*reinterpret_cast<unsigned short*>(&b)) : "memory");
  }
}
template <typename T>
__device__ __forceinline__ float lds_f(uint32_t addr) {
  if constexpr (sizeof(T) == 4) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
  } else {
    unsigned short u;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(u) : "r"(addr));
    return __uint_as_float(((uint32_t)u) << 16);
  }
}

// The kernel is instruction-issue bound before it is HBM bound (ncu on v1: 836 thread instructions
// per pixel-exit, issue slots 50 % busy at 36 % of DRAM peak), so the per-class work is kept to
// ~10 instructions: one LDS with a compile-time offset (plane misalignments repeat with period
// 16/sizeof(T), so only that many row bases are computed per tile), one exp2 (kept for the
// gradient), base-2 log-sum-exp, the target logit fetched by one dynamic LDS.
template <typename T, int CMAX, int TILE, int STAGES>
__global__ void __launch_bounds__(TILE, 2) ce_kernel(
    const T* __restrict__ logits, int64_t exit_stride, const int64_t* __restrict__ targets, int N,
    int C, int64_t HW, int64_t ignore, const float* __restrict__ coef,
    const int64_t* __restrict__ valid_count, T* __restrict__ dlogits, double* __restrict__ part,
    const uint8_t* __restrict__ limit_logits, const uint8_t* __restrict__ limit_targets) {
  // CMAX 32/64 = register capacity with a runtime class count; any other CMAX is the exact count
  // (19 Cityscapes, 21 VOC) and the per-class guards fold away
  if (CMAX != 32 && CMAX != 64) C = CMAX;
  extern __shared__ __align__(128) uint8_t ce_smem[];
  constexpr int ES = (int)sizeof(T);
  constexpr int P = 16 / ES;                       // period of the plane misalignment pattern
  constexpr int rb = ps::row_bytes(TILE, ES), rbt = ps::row_bytes(TILE, 8);
  const int stage_bytes = C * rb + rbt;
  uint64_t* full = reinterpret_cast<uint64_t*>(ce_smem + (size_t)STAGES * stage_bytes);
  uint64_t* done = full + STAGES;   // all TILE threads have consumed (and re-filled with gradients) a stage
  // row ownership is spread over the warps (row r -> warp r % NW, lane r / NW) so that no single warp
  // carries all the bulk-copy bookkeeping
  constexpr int NW = TILE / 32;
  const int my_row = (int)(threadIdx.x & 31) * NW + (int)(threadIdx.x >> 5);

  const int e = blockIdx.y / N, n = blockIdx.y % N;
  const T* base = logits + (int64_t)e * exit_stride + (int64_t)n * C * HW;
  T* gbase = dlogits ? dlogits + (int64_t)e * exit_stride + (int64_t)n * C * HW : nullptr;
  const uint8_t* base_b = reinterpret_cast<const uint8_t*>(base);
  const uint8_t* tg_b = reinterpret_cast<const uint8_t*>(targets + (int64_t)n * HW);

  const int num_tiles = (int)((HW + TILE - 1) / TILE);
  const int my_count = (int)blockIdx.x < num_tiles ? (num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  // row r of tile k is issued by thread r (r < C: class plane r, r == C: the int64 targets)
  auto issue = [&](int k) {
    const int r = my_row;
    if (r > C) return;
    const int s = k % STAGES;
    const int64_t p0 = ((int64_t)blockIdx.x + (int64_t)k * gridDim.x) * TILE;
    const int count = (int)min((int64_t)TILE, HW - p0);
    uint8_t* st = ce_smem + (size_t)s * stage_bytes;
    if (r < C) ps::issue_tile<ES>(st + (size_t)r * rb, rb, full + s, base_b + (int64_t)r * HW * ES, 0, 1, p0, count, limit_logits);
    else ps::issue_tile<8>(st + (size_t)C * rb, rbt, full + s, tg_b, 0, 1, p0, count, limit_targets);
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ps::mbar_init(full + s, C + 1);  // one arming arrival per row
      ps::mbar_init(done + s, TILE);
    }
    ps::fence_barrier_init();
  }
  __syncthreads();
  for (int k = 0; k < STAGES && k < my_count; ++k) issue(k);

  float gscale = 0.f;
  if (gbase) gscale = (coef ? coef[e] : 1.f) / (float)(*valid_count);  // valid == 0 -> NaN like torch
  const uint32_t delta = (uint32_t)(((uint64_t)HW * ES) & 15);
  constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
  float loss_acc = 0.f;

  for (int k = 0; k < my_count; ++k) {
    const int s = k % STAGES;
    const int64_t p0 = ((int64_t)blockIdx.x + (int64_t)k * gridDim.x) * TILE;
    const int64_t p = p0 + threadIdx.x;
    const uint32_t st = ps::smem_u32(ce_smem + (size_t)s * stage_bytes);
    ps::mbar_wait(full + s, (uint32_t)(k / STAGES) & 1u);
    const uint32_t a0 = (uint32_t)((uintptr_t)(base_b + p0 * ES) & 15);
    const uint32_t t0 = (uint32_t)((uintptr_t)(tg_b + p0 * 8) & 15);
    uint32_t rowbase[P];                              // shared address of this thread's pixel in row c, minus c*rb
#pragma unroll
    for (int r = 0; r < P; ++r) rowbase[r] = st + ((a0 + (uint32_t)r * delta) & 15u) + threadIdx.x * ES;
    float v[CMAX];
    int64_t tt = ignore;
    float vt = 0.f;
    if (p < HW) {
      asm volatile("ld.shared.b64 %0, [%1];" : "=l"(tt) : "r"(st + (uint32_t)C * rb + t0 + threadIdx.x * 8));
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) v[c] = lds_f<T>(rowbase[c % P] + (uint32_t)c * rb);
      if (tt >= 0 && tt < C) {
        const uint32_t c = (uint32_t)tt;
        vt = lds_f<T>(st + ((a0 + c * delta) & 15u) + threadIdx.x * ES + c * rb);   // target logit
      }
    }
    const bool ok = p < HW && tt != ignore && tt >= 0 && tt < C;
    if (ok) {
      const int t = (int)tt;
      float m = v[0];
#pragma unroll
      for (int c = 1; c < CMAX; ++c)
        if (c < C) m = fmaxf(m, v[c]);
      const float m2 = m * kLog2e;
      float S = 0.f;
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) {
          v[c] = ce_ex2(fmaf(v[c], kLog2e, -m2));   // exp(v - m), kept for the gradient
          S += v[c];
        }
      loss_acc += ce_lg2(S) * kLn2 - (vt - m);
      if (gbase) {
        // gradient goes back IN PLACE into the staged rows (same misalignment as its destination
        // plane), then leaves with one bulk store per row
        const float inv = gscale / S;
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
          if (c < C) sts_f<T>(rowbase[c % P] + (uint32_t)c * rb, fmaf(v[c], inv, c == t ? -gscale : 0.f));
      }
    } else if (gbase && p < HW) {
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) sts_f<T>(rowbase[c % P] + (uint32_t)c * rb, 0.f);
    }
    if (gbase) ps::fence_proxy_async();      // generic-proxy smem writes -> visible to the bulk engine
    ps::mbar_arrive(done + s);               // this thread is finished with stage s
    if (my_row <= C) {                       // the row's owner stores it and refills it; nobody else waits
      const int r = my_row;
      ps::mbar_wait(done + s, (uint32_t)(k / STAGES) & 1u);
      if (gbase && r < C) {
        const int count = (int)min((int64_t)TILE, HW - p0);
        ps::store_row<T>(ce_smem + (size_t)s * stage_bytes + (size_t)r * rb, gbase + (int64_t)r * HW + p0, count);
        ps::bulk_commit();
        if (k + STAGES < my_count) ps::bulk_wait_read0();   // the row must be read out before it is refilled
      }
      if (k + STAGES < my_count) issue(k + STAGES);
    }
  }
  if (gbase && my_row < C) ps::bulk_wait_read0();
  if (part) {
    __shared__ double sred[32];
    double ws = warp_sum((double)loss_acc);
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = ws;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int i = 0; i < TILE / 32; ++i) t += sred[i];
      part[((int64_t)e * N + n) * gridDim.x + blockIdx.x] = t;  // slot order: [e][n][blockIdx.x]
    }
  }
}

__global__ void ce_finalize_kernel(const double* __restrict__ part, int per_exit_parts,
                                   const int64_t* __restrict__ valid_count,
                                   float* __restrict__ per_exit) {
  const int e = blockIdx.x;
  __shared__ double s[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < per_exit_parts; i += blockDim.x)
    acc += part[(int64_t)e * per_exit_parts + i];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) per_exit[e] = (float)(s[0] / (double)(*valid_count));
}

template <typename T>
__global__ void scale_exits_kernel(T* __restrict__ d, int64_t exit_stride, int64_t elems,
                                   const float* __restrict__ g, const float* __restrict__ coef) {
  const int e = blockIdx.y;
  const float ge = g[e], ce = coef ? coef[e] : 1.f;
  if (ge == ce) return;  // the common case (upstream gradient == what the forward assumed)
  const float r = ge / ce;
  T* p = d + (int64_t)e * exit_stride;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < elems;
       i += (int64_t)gridDim.x * blockDim.x)
    stf(p + i, ldf(p + i) * r);
}

template <typename T, int CMAX, int TILE>
static int launch_ce_cfg(const T* logits, int64_t exit_stride, const int64_t* targets, int E, int N,
                         int C, int64_t HW, int64_t ignore, const float* coef,
                         const int64_t* valid_count, T* dlogits, double* part, cudaStream_t stream) {
  constexpr int kStages = 3;   // two CTAs per SM x three tiles each in flight
  const int rb = ps::row_bytes(TILE, (int)sizeof(T)), rbt = ps::row_bytes(TILE, 8);
  const size_t stage_bytes = (size_t)C * rb + rbt;
  const size_t smem = kStages * stage_bytes + 2 * kStages * sizeof(uint64_t);
  if (smem > 113 * 1024) { set_error("multi_exit_ce: C=%d does not fit the staging buffers", C); return EESEG_ERR_UNSUPPORTED; }
  if (dlogits && (((uintptr_t)dlogits ^ (uintptr_t)logits) & 15)) {
    set_error("multi_exit_ce: logits and dlogits must have the same 16-byte misalignment");
    return EESEG_ERR_ARG;
  }
  auto kern = ce_kernel<T, CMAX, TILE, kStages>;
  EESEG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const uintptr_t end_l = (uintptr_t)(logits + (int64_t)(E - 1) * exit_stride + (int64_t)N * C * HW);
  const uintptr_t end_t = (uintptr_t)(targets + (int64_t)N * HW);
  dim3 grid(ce_grid_x(E, N, HW), E * N);
  kern<<<grid, TILE, smem, stream>>>(logits, exit_stride, targets, N, C, HW, ignore, coef, valid_count,
                                     dlogits, part, (const uint8_t*)((end_l + 15) & ~(uintptr_t)15),
                                     (const uint8_t*)((end_t + 15) & ~(uintptr_t)15));
  return check_launch("ce_kernel");
}

template <typename T>
static int launch_ce(const T* logits, int64_t exit_stride, const int64_t* targets, int E, int N,
                     int C, int64_t HW, int64_t ignore, const float* coef,
                     const int64_t* valid_count, T* dlogits, double* part, cudaStream_t stream) {
  if (C == 21) return launch_ce_cfg<T, 21, 256>(logits, exit_stride, targets, E, N, C, HW, ignore, coef, valid_count, dlogits, part, stream);
  if (C == 19) return launch_ce_cfg<T, 19, 256>(logits, exit_stride, targets, E, N, C, HW, ignore, coef, valid_count, dlogits, part, stream);
  if (C <= 32) return launch_ce_cfg<T, 32, 256>(logits, exit_stride, targets, E, N, C, HW, ignore, coef, valid_count, dlogits, part, stream);
  if (C <= 64) return launch_ce_cfg<T, 64, 128>(logits, exit_stride, targets, E, N, C, HW, ignore, coef, valid_count, dlogits, part, stream);
  set_error("multi_exit_ce: C=%d > 64 classes is not supported by this build", C);
  return EESEG_ERR_UNSUPPORTED;
}

}  // namespace eeseg

using namespace eeseg;

extern "C" size_t eeseg_multi_exit_ce_workspace_bytes(int E, int N, int64_t HW) {
  if (E <= 0 || N <= 0) return 256;
  return (size_t)E * N * ce_grid_x(E, N, HW) * sizeof(double) + 256;
}

extern "C" int eeseg_multi_exit_ce_fwd(const void* logits, int dtype, int64_t exit_stride,
                                       const int64_t* targets, int E, int N, int C, int64_t HW,
                                       int64_t ignore_index, const float* coef, float* per_exit,
                                       int64_t* valid_count, void* dlogits, void* workspace,
                                       void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  EESEG_REQUIRE(logits && targets && per_exit && valid_count && workspace, "multi_exit_ce_fwd: null pointer");
  EESEG_REQUIRE(E >= 1 && N >= 1 && C >= 1 && HW >= 1, "multi_exit_ce_fwd: bad sizes");
  EESEG_REQUIRE((int64_t)E * N <= 65535, "multi_exit_ce_fwd: E*N too large");
  EESEG_REQUIRE(dtype == EESEG_F32 || dtype == EESEG_BF16, "multi_exit_ce_fwd: dtype %d", dtype);
  EESEG_CUDA(cudaMemsetAsync(valid_count, 0, sizeof(int64_t), stream));
  const int64_t total = (int64_t)N * HW;
  int cblocks = (int)((total + 1023) / 1024 < kNumSMs * 4 ? (total + 1023) / 1024 : kNumSMs * 4);
  count_valid_kernel<<<cblocks, 256, 0, stream>>>(targets, total, C, ignore_index,
                                                  reinterpret_cast<unsigned long long*>(valid_count));
  int rc = check_launch("count_valid_kernel");
  if (rc) return rc;
  double* part = reinterpret_cast<double*>(workspace);
  if (dtype == EESEG_F32)
    rc = launch_ce<float>((const float*)logits, exit_stride, targets, E, N, C, HW, ignore_index,
                          coef, valid_count, (float*)dlogits, part, stream);
  else
    rc = launch_ce<__nv_bfloat16>((const __nv_bfloat16*)logits, exit_stride, targets, E, N, C, HW,
                                  ignore_index, coef, valid_count, (__nv_bfloat16*)dlogits, part,
                                  stream);
  if (rc) return rc;
  ce_finalize_kernel<<<E, 256, 0, stream>>>(part, N * ce_grid_x(E, N, HW), valid_count, per_exit);
  return check_launch("ce_finalize_kernel");
}

extern "C" int eeseg_multi_exit_ce_bwd(const void* logits, int dtype, int64_t exit_stride,
                                       const int64_t* targets, int E, int N, int C, int64_t HW,
                                       int64_t ignore_index, const float* g,
                                       const int64_t* valid_count, void* dlogits, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  EESEG_REQUIRE(logits && targets && g && valid_count && dlogits, "multi_exit_ce_bwd: null pointer");
  EESEG_REQUIRE(E >= 1 && N >= 1 && C >= 1 && HW >= 1, "multi_exit_ce_bwd: bad sizes");
  EESEG_REQUIRE(dtype == EESEG_F32 || dtype == EESEG_BF16, "multi_exit_ce_bwd: dtype %d", dtype);
  if (dtype == EESEG_F32)
    return launch_ce<float>((const float*)logits, exit_stride, targets, E, N, C, HW, ignore_index, g,
                            valid_count, (float*)dlogits, nullptr, stream);
  return launch_ce<__nv_bfloat16>((const __nv_bfloat16*)logits, exit_stride, targets, E, N, C, HW,
                                  ignore_index, g, valid_count, (__nv_bfloat16*)dlogits, nullptr,
                                  stream);
}

extern "C" int eeseg_scale_exits(void* dlogits, int dtype, int64_t exit_stride, int E,
                                 int64_t per_exit_elems, const float* g, const float* coef,
                                 void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  EESEG_REQUIRE(dlogits && g, "scale_exits: null pointer");
  EESEG_REQUIRE(dtype == EESEG_F32 || dtype == EESEG_BF16, "scale_exits: dtype %d", dtype);
  if (E <= 0 || per_exit_elems <= 0) return EESEG_OK;
  dim3 grid(kNumSMs * 4, E);
  if (dtype == EESEG_F32)
    scale_exits_kernel<float><<<grid, 256, 0, stream>>>((float*)dlogits, exit_stride, per_exit_elems, g, coef);
  else
    scale_exits_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((__nv_bfloat16*)dlogits, exit_stride, per_exit_elems, g, coef);
  return check_launch("scale_exits_kernel");
}
