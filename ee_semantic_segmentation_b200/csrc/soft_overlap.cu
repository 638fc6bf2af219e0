// Soft-overlap sums of the branchy Dice / Jaccard losses and their gradient.
// Reference contract: BSL.DiceLoss._compute_loss (branchy_seg_losses.py:40-48) and BSL.JaccardLoss._compute_loss
// (:50-77) under BrSegLoss.forward (:24-38): per exit, image and class, over the pixels of the image,
//     S_pt[e,n,c] = sum softmax(y[e,n])[c] * [t == c]     ("intersection")
//     S_p [e,n,c] = sum softmax(y[e,n])[c]
//     S_t [n,c]   = sum [t == c]                           (targets outside [0,C) — void — match no class)
// Every loss of that family is a small formula on these [E,N,C] tensors (the Python mirror keeps the reference's
// formula verbatim), so its gradient reaches the logits through two coefficient tensors A = dL/dS_pt, B = dL/dS_p:
//     dL/dz_c = p_c * (g_c - sum_k p_k g_k),   g_c = B_c + A_c [t == c].
// The reference materialises softmax, a one-hot int64 tensor [N,HW,C] and their products per exit; here both
// directions are one streaming pass over the logits (HBM-bound, the access pattern of multi_exit_ce.cu: thread =
// pixel, the C class values of the pixel in registers, plane-coalesced element loads). Block partials are summed
// in a fixed order (bit-reproducible).
#include "common.cuh"

namespace eeseg {

constexpr int kSoThreads = 256;

__device__ __forceinline__ float so_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// softmax of one pixel in registers; returns the target class (or -1)
template <typename T, int CMAX>
__device__ __forceinline__ int so_softmax(const T* __restrict__ base, const int64_t* __restrict__ tgt, int64_t p, int C,
                                          uint32_t pb, bool live, float (&v)[CMAX]) {
  constexpr float kLog2e = 1.4426950408889634f;
  int64_t tt = -1;
  if (live) {
    tt = __ldg(tgt + p);
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) v[c] = ldf_stream(plane_ptr(base + p, (uint32_t)c, pb));
  } else {
#pragma unroll
    for (int c = 0; c < CMAX; ++c) v[c] = 0.f;
  }
  float m = v[0];
#pragma unroll
  for (int c = 1; c < CMAX; ++c)
    if (c < C) m = fmaxf(m, v[c]);
  const float m2 = m * kLog2e;
  float S = 0.f;
#pragma unroll
  for (int c = 0; c < CMAX; ++c)
    if (c < C) {
      v[c] = so_ex2(fmaf(v[c], kLog2e, -m2));
      S += v[c];
    }
  const float inv = live ? 1.f / S : 0.f;
#pragma unroll
  for (int c = 0; c < CMAX; ++c)
    if (c < C) v[c] *= inv;
  return (tt >= 0 && tt < C) ? (int)tt : -1;
}

// grid (blocks per image, N, E); part[(((e*N + n)*gridDim.x + blk)*3 + q)*C + c]
template <typename T, int CMAX>
__global__ void __launch_bounds__(kSoThreads) soft_sums_kernel(const T* __restrict__ logits, int64_t exit_stride,
                                                               const int64_t* __restrict__ targets, int N, int C, int64_t HW,
                                                               float* __restrict__ part) {
  if (CMAX != 32 && CMAX != 64) C = CMAX;
  const int e = blockIdx.z, n = blockIdx.y;
  const int64_t p = (int64_t)blockIdx.x * kSoThreads + threadIdx.x;
  const bool live = p < HW;
  const uint32_t pb = (uint32_t)HW * (uint32_t)sizeof(T);
  float v[CMAX];
  const int t = so_softmax<T, CMAX>(logits + (int64_t)e * exit_stride + (int64_t)n * C * HW, targets + (int64_t)n * HW, p, C, pb,
                                    live, v);
  __shared__ float sm[kSoThreads / 32][3][CMAX];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int c = 0; c < CMAX; ++c)
    if (c < C) {
      const bool hit = t == c;
      const float a = warp_sum(hit ? v[c] : 0.f), b = warp_sum(v[c]);
      const unsigned cnt = __ballot_sync(0xffffffffu, hit);
      if (lane == 0) { sm[warp][0][c] = a; sm[warp][1][c] = b; sm[warp][2][c] = (float)__popc(cnt); }
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * C; i += kSoThreads) {
    const int q = i / C, c = i - q * C;
    float s = 0.f;
    for (int w = 0; w < kSoThreads / 32; ++w) s += sm[w][q][c];   // fixed order
    part[((((int64_t)e * N + n) * gridDim.x + blockIdx.x) * 3 + q) * C + c] = s;
  }
}

// one warp per (e, n, q, c): ordered fp64 sum over the blocks of the image
__global__ void soft_sums_finalize_kernel(const float* __restrict__ part, int blocks, int C, int64_t total /* E*N*3*C */,
                                          float* __restrict__ sums) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= total) return;
  const int64_t en = w / (3 * C);
  const int qc = (int)(w - en * 3 * C);
  double s = 0.0;
  for (int i = (int)(threadIdx.x & 31); i < blocks; i += 32) s += (double)part[(en * blocks + i) * 3 * C + qc];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sums[w] = (float)s;
}

template <typename T, int CMAX>
__global__ void __launch_bounds__(kSoThreads) soft_bwd_kernel(const T* __restrict__ logits, int64_t exit_stride,
                                                              const int64_t* __restrict__ targets, int N, int C, int64_t HW,
                                                              const float* __restrict__ A, const float* __restrict__ B,
                                                              T* __restrict__ dlogits) {
  if (CMAX != 32 && CMAX != 64) C = CMAX;
  const int e = blockIdx.z, n = blockIdx.y;
  const int64_t p = (int64_t)blockIdx.x * kSoThreads + threadIdx.x;
  const bool live = p < HW;
  const uint32_t pb = (uint32_t)HW * (uint32_t)sizeof(T);
  __shared__ float sA[CMAX], sB[CMAX];
  if ((int)threadIdx.x < C) {
    sA[threadIdx.x] = A[((int64_t)e * N + n) * C + threadIdx.x];
    sB[threadIdx.x] = B[((int64_t)e * N + n) * C + threadIdx.x];
  }
  float v[CMAX];
  const int64_t img = (int64_t)e * exit_stride + (int64_t)n * C * HW;
  const int t = so_softmax<T, CMAX>(logits + img, targets + (int64_t)n * HW, p, C, pb, live, v);
  __syncthreads();
  if (!live) return;
  // p_t * A_t by a select chain (a runtime index into v would move the array to local memory)
  float hit = 0.f;
#pragma unroll
  for (int c = 0; c < CMAX; ++c)
    if (c < C) hit = (c == t) ? v[c] * sA[c] : hit;
  float dot = hit;
#pragma unroll
  for (int c = 0; c < CMAX; ++c)
    if (c < C) dot = fmaf(v[c], sB[c], dot);
  T* gb = dlogits + img + p;
#pragma unroll
  for (int c = 0; c < CMAX; ++c)
    if (c < C) stf(plane_ptr(gb, (uint32_t)c, pb), v[c] * (sB[c] - dot) + (c == t ? hit : 0.f));
}

template <typename T>
static int launch_sums(const T* logits, int64_t exit_stride, const int64_t* targets, int E, int N, int C, int64_t HW,
                       float* part, cudaStream_t stream) {
  dim3 grid((unsigned)((HW + kSoThreads - 1) / kSoThreads), N, E);
  if (C == 21) soft_sums_kernel<T, 21><<<grid, kSoThreads, 0, stream>>>(logits, exit_stride, targets, N, C, HW, part);
  else if (C == 19) soft_sums_kernel<T, 19><<<grid, kSoThreads, 0, stream>>>(logits, exit_stride, targets, N, C, HW, part);
  else if (C <= 32) soft_sums_kernel<T, 32><<<grid, kSoThreads, 0, stream>>>(logits, exit_stride, targets, N, C, HW, part);
  else soft_sums_kernel<T, 64><<<grid, kSoThreads, 0, stream>>>(logits, exit_stride, targets, N, C, HW, part);
  return check_launch("soft_sums_kernel");
}

template <typename T>
static int launch_bwd(const T* logits, int64_t exit_stride, const int64_t* targets, int E, int N, int C, int64_t HW,
                      const float* A, const float* B, T* dlogits, cudaStream_t stream) {
  dim3 grid((unsigned)((HW + kSoThreads - 1) / kSoThreads), N, E);
  if (C == 21) soft_bwd_kernel<T, 21><<<grid, kSoThreads, 0, stream>>>(logits, exit_stride, targets, N, C, HW, A, B, dlogits);
  else if (C == 19) soft_bwd_kernel<T, 19><<<grid, kSoThreads, 0, stream>>>(logits, exit_stride, targets, N, C, HW, A, B, dlogits);
  else if (C <= 32) soft_bwd_kernel<T, 32><<<grid, kSoThreads, 0, stream>>>(logits, exit_stride, targets, N, C, HW, A, B, dlogits);
  else soft_bwd_kernel<T, 64><<<grid, kSoThreads, 0, stream>>>(logits, exit_stride, targets, N, C, HW, A, B, dlogits);
  return check_launch("soft_bwd_kernel");
}

}  // namespace eeseg

using namespace eeseg;

extern "C" size_t eeseg_soft_overlap_workspace_bytes(int E, int N, int C, int64_t HW) {
  if (E <= 0 || N <= 0 || C <= 0 || HW <= 0) return 256;
  return (size_t)E * N * ((HW + kSoThreads - 1) / kSoThreads) * 3 * C * sizeof(float) + 256;
}

extern "C" int eeseg_soft_overlap_fwd(const void* logits, int dtype, int64_t exit_stride, const int64_t* targets, int E, int N,
                                      int C, int64_t HW, float* sums, void* workspace, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  EESEG_REQUIRE(logits && targets && sums && workspace, "soft_overlap_fwd: null pointer");
  EESEG_REQUIRE(E >= 1 && N >= 1 && C >= 1 && C <= 64 && HW >= 1, "soft_overlap_fwd: bad sizes (C <= 64)");
  EESEG_REQUIRE(N <= 65535 && E <= 65535 && HW < (1ll << 29), "soft_overlap_fwd: E, N <= 65535 and HW < 2^29");
  EESEG_REQUIRE(dtype == EESEG_F32 || dtype == EESEG_BF16, "soft_overlap_fwd: dtype %d", dtype);
  float* part = reinterpret_cast<float*>(workspace);
  int rc = dtype == EESEG_F32
               ? launch_sums<float>((const float*)logits, exit_stride, targets, E, N, C, HW, part, stream)
               : launch_sums<__nv_bfloat16>((const __nv_bfloat16*)logits, exit_stride, targets, E, N, C, HW, part, stream);
  if (rc) return rc;
  const int blocks = (int)((HW + kSoThreads - 1) / kSoThreads);
  const int64_t total = (int64_t)E * N * 3 * C;
  soft_sums_finalize_kernel<<<(unsigned)((total * 32 + 255) / 256), 256, 0, stream>>>(part, blocks, C, total, sums);
  return check_launch("soft_sums_finalize_kernel");
}

extern "C" int eeseg_soft_overlap_bwd(const void* logits, int dtype, int64_t exit_stride, const int64_t* targets, int E, int N,
                                      int C, int64_t HW, const float* dsum_pt, const float* dsum_p, void* dlogits,
                                      void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  EESEG_REQUIRE(logits && targets && dsum_pt && dsum_p && dlogits, "soft_overlap_bwd: null pointer");
  EESEG_REQUIRE(E >= 1 && N >= 1 && C >= 1 && C <= 64 && HW >= 1, "soft_overlap_bwd: bad sizes (C <= 64)");
  EESEG_REQUIRE(N <= 65535 && E <= 65535 && HW < (1ll << 29), "soft_overlap_bwd: E, N <= 65535 and HW < 2^29");
  EESEG_REQUIRE(dtype == EESEG_F32 || dtype == EESEG_BF16, "soft_overlap_bwd: dtype %d", dtype);
  if (dtype == EESEG_F32)
    return launch_bwd<float>((const float*)logits, exit_stride, targets, E, N, C, HW, dsum_pt, dsum_p, (float*)dlogits, stream);
  return launch_bwd<__nv_bfloat16>((const __nv_bfloat16*)logits, exit_stride, targets, E, N, C, HW, dsum_pt, dsum_p,
                                   (__nv_bfloat16*)dlogits, stream);
}
