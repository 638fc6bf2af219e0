// Weighted multi-exit pixelwise cross-entropy, forward (+ fused gradient) and backward.
// Reference contract: BrXEntropyLoss (my_pixelwise_xentropy.py:19-46) over
// torch.nn.CrossEntropyLoss(reduction='mean', ignore_index) per exit — see include/eeseg.h.
//
// HBM-bound: one pass over logits [E][N][C][HW] (read) and, when the gradient is wanted, one write
// of the same size — the reference does E separate log_softmax + nll passes forward and E backward.
//
// v3 (what the measurements chose, tools/ub_stream.cu): thread = one pixel, ALL exits of that pixel
// in flight at once: up to 3 x C independent, plane-coalesced 4 B loads per thread are issued before
// the first use, so a 256-thread CTA keeps ~64 KB of reads outstanding without any staging — 87 % of
// the measured HBM copy peak for fp32, against 52 % for the bulk-TMA shared-memory ring of v2, whose
// 1 KB row copies and per-stage barriers left the SM waiting (ncu: 61 % of warp stalls at barriers).
// 513 x 513 planes are odd-sized, so no wider-than-element vector access is aligned; lanes walk
// consecutive pixels of a plane, every request is one (misaligned) 128 B / 64 B line pair that L2 merges.
// Per-block loss partials go to a scratch array and are reduced in a fixed order (bit-reproducible).
#include "common.cuh"

namespace eeseg {

constexpr int kCeThreads = 256;

// blocks per image (== loss partial slots per (exit, image))
static inline int ce_grid_x(int E, int N, int64_t HW) {
  (void)E; (void)N;
  return (int)((HW + kCeThreads - 1) / kCeThreads);
}

__global__ void count_valid_kernel(const int64_t* __restrict__ targets, int64_t total, int C,
                                   int64_t ignore, unsigned long long* __restrict__ out) {
  // programmatic dependent launch: the CE kernel behind us may start loading logits right away; it
  // waits (griddepcontrol.wait) for this grid to finish before it reads the count
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
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

// CMAX 32/64 = register capacity with a runtime class count; any other CMAX is the exact count
// (19 Cityscapes, 21 VOC) and the per-class guards fold away. EB = exits handled by one thread
// (their loads are all issued before the first use); blockIdx.z selects the group of EB exits.
template <typename T, int CMAX, int EB>
__global__ void __launch_bounds__(kCeThreads) ce_kernel(
    const T* __restrict__ logits, int64_t exit_stride, const int64_t* __restrict__ targets, int e0, int N,
    int C, int64_t HW, int64_t ignore, const float* __restrict__ coef,
    const int64_t* __restrict__ valid_count, T* __restrict__ dlogits, double* __restrict__ part) {
  if (CMAX != 32 && CMAX != 64) C = CMAX;
  const int n = blockIdx.y;
  const int eg = e0 + (int)blockIdx.z * EB;
  const int64_t p = (int64_t)blockIdx.x * kCeThreads + threadIdx.x;
  const bool live = p < HW;
  const uint32_t pb = (uint32_t)HW * (uint32_t)sizeof(T);   // plane stride in bytes (host: HW*sizeof(T) < 2^32)
  constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;

  float v[EB][CMAX];
  float vt[EB];
  int64_t tt = ignore;
  if (live) {
    tt = __ldg(targets + (int64_t)n * HW + p);
#pragma unroll
    for (int j = 0; j < EB; ++j) {
      const T* base = logits + (int64_t)(eg + j) * exit_stride + (int64_t)n * C * HW + p;
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) v[j][c] = ldf(plane_ptr(base, (uint32_t)c, pb));
    }
  }
  const bool ok = live && tt != ignore && tt >= 0 && tt < C;
  const uint32_t t = ok ? (uint32_t)tt : 0u;
#pragma unroll
  for (int j = 0; j < EB; ++j) {   // the target logit: one gathered load (L1/L2 hit: the line was just read)
    vt[j] = 0.f;
    if (ok) vt[j] = ldf(plane_ptr(logits + (int64_t)(eg + j) * exit_stride + (int64_t)n * C * HW + p, t, pb));
  }
  // everything above touched only tensors that were complete before the preceding kernel in the
  // stream (count_valid_kernel) started; its result is needed from here on
  asm volatile("griddepcontrol.wait;" ::: "memory");
  float vinv = 0.f;
  if (dlogits) vinv = 1.f / (float)(*valid_count);   // valid == 0 -> inf; only ok pixels use it
  float loss[EB];
#pragma unroll
  for (int j = 0; j < EB; ++j) {
    loss[j] = 0.f;
    if (live) {
      float m = v[j][0];
#pragma unroll
      for (int c = 1; c < CMAX; ++c)
        if (c < C) m = fmaxf(m, v[j][c]);
      const float m2 = m * kLog2e;
      float S = 0.f;
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) {
          v[j][c] = ce_ex2(fmaf(v[j][c], kLog2e, -m2));   // exp(v - m), kept for the gradient
          S += v[j][c];
        }
      if (ok) loss[j] = ce_lg2(S) * kLn2 - (vt[j] - m);
      if (dlogits) {
        // softmax * g for every class, then the target class once more with its -g (same thread, same
        // address, program order); void pixels get zeros
        const float gscale = ok ? (coef ? coef[eg + j] : 1.f) * vinv : 0.f;
        const float inv = ok ? gscale / S : 0.f;
        T* gb = dlogits + (int64_t)(eg + j) * exit_stride + (int64_t)n * C * HW + p;
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
          if (c < C) stf(plane_ptr(gb, (uint32_t)c, pb), v[j][c] * inv);
        if (ok) stf(plane_ptr(gb, t, pb), fmaf(ce_ex2(fmaf(vt[j], kLog2e, -m2)), inv, -gscale));
      }
    }
  }
  if (part) {
    __shared__ double sred[EB][kCeThreads / 32];
#pragma unroll
    for (int j = 0; j < EB; ++j) {
      const double ws = warp_sum((double)loss[j]);
      if ((threadIdx.x & 31) == 0) sred[j][threadIdx.x >> 5] = ws;
    }
    __syncthreads();
    if (threadIdx.x < EB) {
      double tsum = 0.0;
      for (int i = 0; i < kCeThreads / 32; ++i) tsum += sred[threadIdx.x][i];
      part[((int64_t)(eg + threadIdx.x) * N + n) * gridDim.x + blockIdx.x] = tsum;   // slot order: [e][n][block]
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

template <typename T, int CMAX, int EB>
static int launch_ce_one(dim3 grid, bool pdl, const T* logits, int64_t exit_stride, const int64_t* targets, int e0,
                         int N, int C, int64_t HW, int64_t ignore, const float* coef, const int64_t* valid_count,
                         T* dlogits, double* part, cudaStream_t stream) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kCeThreads);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  EESEG_CUDA(cudaLaunchKernelEx(&cfg, ce_kernel<T, CMAX, EB>, logits, exit_stride, targets, e0, N, C, HW, ignore, coef,
                                valid_count, dlogits, part));
  return check_launch("ce_kernel");
}

// `pdl`: the launch directly follows count_valid_kernel in the stream and may overlap it
template <typename T, int CMAX, int EBMAX>
static int launch_ce_cfg(const T* logits, int64_t exit_stride, const int64_t* targets, int E, int N,
                         int C, int64_t HW, int64_t ignore, const float* coef,
                         const int64_t* valid_count, T* dlogits, double* part, bool pdl, cudaStream_t stream) {
  // exits in groups of 3 (2, 1 for the remainder): a thread holds the logits of all exits of a group
  const unsigned gx = (unsigned)ce_grid_x(E, N, HW);
  int e0 = 0;
  if (EBMAX >= 3 && E >= 3) {
    int rc = launch_ce_one<T, CMAX, (EBMAX >= 3 ? 3 : 1)>(dim3(gx, N, E / 3), pdl, logits, exit_stride, targets, 0, N, C, HW,
                                                         ignore, coef, valid_count, dlogits, part, stream);
    if (rc) return rc;
    e0 = E / 3 * 3;
    pdl = false;
  }
  if (EBMAX >= 2 && E - e0 == 2)
    return launch_ce_one<T, CMAX, (EBMAX >= 2 ? 2 : 1)>(dim3(gx, N, 1), pdl, logits, exit_stride, targets, e0, N, C, HW, ignore,
                                                       coef, valid_count, dlogits, part, stream);
  if (E - e0 >= 1)
    return launch_ce_one<T, CMAX, 1>(dim3(gx, N, E - e0), pdl, logits, exit_stride, targets, e0, N, C, HW, ignore, coef,
                                     valid_count, dlogits, part, stream);
  return EESEG_OK;
}

template <typename T>
static int launch_ce(const T* logits, int64_t exit_stride, const int64_t* targets, int E, int N,
                     int C, int64_t HW, int64_t ignore, const float* coef,
                     const int64_t* valid_count, T* dlogits, double* part, bool pdl, cudaStream_t stream) {
  if (C == 21) return launch_ce_cfg<T, 21, 3>(logits, exit_stride, targets, E, N, C, HW, ignore, coef, valid_count, dlogits, part, pdl, stream);
  if (C == 19) return launch_ce_cfg<T, 19, 3>(logits, exit_stride, targets, E, N, C, HW, ignore, coef, valid_count, dlogits, part, pdl, stream);
  if (C <= 32) return launch_ce_cfg<T, 32, 2>(logits, exit_stride, targets, E, N, C, HW, ignore, coef, valid_count, dlogits, part, pdl, stream);
  if (C <= 64) return launch_ce_cfg<T, 64, 1>(logits, exit_stride, targets, E, N, C, HW, ignore, coef, valid_count, dlogits, part, pdl, stream);
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
  EESEG_REQUIRE(N <= 65535 && E <= 65535, "multi_exit_ce_fwd: E, N must be <= 65535");
  EESEG_REQUIRE(HW < (1ll << 29), "multi_exit_ce_fwd: HW must be < 2^29 pixels per image");
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
                          coef, valid_count, (float*)dlogits, part, true, stream);
  else
    rc = launch_ce<__nv_bfloat16>((const __nv_bfloat16*)logits, exit_stride, targets, E, N, C, HW,
                                  ignore_index, coef, valid_count, (__nv_bfloat16*)dlogits, part,
                                  true, stream);
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
  EESEG_REQUIRE(N <= 65535 && E <= 65535, "multi_exit_ce_bwd: E, N must be <= 65535");
  EESEG_REQUIRE(HW < (1ll << 29), "multi_exit_ce_bwd: HW must be < 2^29 pixels per image");
  EESEG_REQUIRE(dtype == EESEG_F32 || dtype == EESEG_BF16, "multi_exit_ce_bwd: dtype %d", dtype);
  if (dtype == EESEG_F32)
    return launch_ce<float>((const float*)logits, exit_stride, targets, E, N, C, HW, ignore_index, g,
                            valid_count, (float*)dlogits, nullptr, false, stream);
  return launch_ce<__nv_bfloat16>((const __nv_bfloat16*)logits, exit_stride, targets, E, N, C, HW,
                                  ignore_index, g, valid_count, (__nv_bfloat16*)dlogits, nullptr,
                                  false, stream);
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
