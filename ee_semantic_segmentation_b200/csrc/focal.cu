// Branchy focal loss, forward (+ fused gradient).
// Reference contract: BSL.FocalLoss._compute_loss (branchy_seg_losses.py:122-131) under BrSegLoss.forward (:24-38):
//     loss[e,n,px] = -alpha[t] * w[e,n,px] * (1 - p_t)^gamma * log p_t,   p = softmax over C,   t = targets[n,px]
// reduced per exit by mean / sum over (n, px) and combined with the exit weights. The reference gathers with the
// raw target (labels outside [0,C) are an index error there; the Python mirror rejects them), so every pixel counts.
// w is an optional per-pixel weight with its own exit / image strides (0 = broadcast): it carries the reference's
// `loss * alpha[targets]` product, which broadcasts [N,H,W] against [N,1,H,W] (a pixel is weighted by the sum of the
// alphas of ALL images at that position), and the upstream gradient map of reduction='none'.
// One streaming pass in the access pattern of multi_exit_ce.cu (thread = pixel, the C class values in registers):
// per-block fp64 partial sums of the loss, the optional per-pixel loss map (reduction 'none'), and the gradient
//     dL/dz_c = g * alpha_t * F'(lp) * ([c == t] - p_c),   lp = log p_t,  F(lp) = -(1 - e^lp)^gamma * lp,
//     F'(lp) = gamma * (1 - p_t)^(gamma-1) * p_t * lp - (1 - p_t)^gamma
// with g = coef[e] (the caller folds 1/(N*HW) of a mean reduction into coef).
#include "common.cuh"

namespace eeseg {

constexpr int kFocalThreads = 256;

__device__ __forceinline__ float fo_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <typename T, int CMAX>
__global__ void __launch_bounds__(kFocalThreads) focal_kernel(const T* __restrict__ logits, int64_t exit_stride,
                                                              const int64_t* __restrict__ targets, int N, int C, int64_t HW,
                                                              float gamma, const float* __restrict__ alpha,
                                                              const float* __restrict__ pixw, int64_t pw_exit_stride,
                                                              int64_t pw_image_stride, const float* __restrict__ coef,
                                                              float* __restrict__ loss_map,
                                                              T* __restrict__ dlogits, double* __restrict__ part) {
  if (CMAX != 32 && CMAX != 64) C = CMAX;
  const int e = blockIdx.z, n = blockIdx.y;
  const int64_t p = (int64_t)blockIdx.x * kFocalThreads + threadIdx.x;
  const bool live = p < HW;
  const uint32_t pb = (uint32_t)HW * (uint32_t)sizeof(T);
  constexpr float kLog2e = 1.4426950408889634f;
  const int64_t img = (int64_t)e * exit_stride + (int64_t)n * C * HW;
  float loss = 0.f;
  if (live) {
    float v[CMAX];
    const int64_t tt = __ldg(targets + (int64_t)n * HW + p);
    const T* base = logits + img + p;
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) v[c] = ldf_stream(plane_ptr(base, (uint32_t)c, pb));
    const uint32_t t = (tt >= 0 && tt < C) ? (uint32_t)tt : 0u;   // validated on the host side of the API
    const float vt = ldf(plane_ptr(base, t, pb));
    float m = v[0];
#pragma unroll
    for (int c = 1; c < CMAX; ++c)
      if (c < C) m = fmaxf(m, v[c]);
    const float m2 = m * kLog2e;
    float S = 0.f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) {
        v[c] = fo_ex2(fmaf(v[c], kLog2e, -m2));
        S += v[c];
      }
    const float lp = (vt - m) - logf(S);          // log p_t
    const float pt = expf(lp);
    const float om = fmaxf(1.f - pt, 0.f);
    float a = alpha ? alpha[t] : 1.f;
    if (pixw) a *= __ldg(pixw + (int64_t)e * pw_exit_stride + (int64_t)n * pw_image_stride + p);
    const float omg = gamma == 0.f ? 1.f : powf(om, gamma);
    loss = -a * omg * lp;
    if (loss_map) loss_map[((int64_t)e * N + n) * HW + p] = loss;
    if (dlogits) {
      // F'(lp); (1-pt)^(gamma-1) * pt -> 0 as pt -> 1 for gamma >= 1, guarded for gamma < 1
      const float omg1 = gamma == 0.f ? 0.f : (om > 0.f ? gamma * powf(om, gamma - 1.f) * pt * lp : 0.f);
      const float k = (coef ? coef[e] : 1.f) * a * (omg1 - omg);     // g * alpha_t * F'(lp)
      const float inv = 1.f / S;
      T* gb = dlogits + img + p;
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) stf(plane_ptr(gb, (uint32_t)c, pb), -k * v[c] * inv);
      stf(plane_ptr(gb, t, pb), k * (1.f - pt));   // the target class once more: k * (1 - p_t)
    }
  }
  if (part) {
    __shared__ double sred[kFocalThreads / 32];
    const double ws = warp_sum((double)loss);
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = ws;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int i = 0; i < kFocalThreads / 32; ++i) t += sred[i];
      part[((int64_t)e * N + n) * gridDim.x + blockIdx.x] = t;
    }
  }
}

__global__ void focal_finalize_kernel(const double* __restrict__ part, int per_exit_parts, float* __restrict__ per_exit) {
  const int e = blockIdx.x;
  __shared__ double s[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < per_exit_parts; i += blockDim.x) acc += part[(int64_t)e * per_exit_parts + i];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) per_exit[e] = (float)s[0];
}

template <typename T>
static int launch_focal(const T* logits, int64_t exit_stride, const int64_t* targets, int E, int N, int C, int64_t HW, float gamma,
                        const float* alpha, const float* pixw, int64_t pw_es, int64_t pw_is, const float* coef, float* loss_map,
                        T* dlogits, double* part, cudaStream_t stream) {
  dim3 grid((unsigned)((HW + kFocalThreads - 1) / kFocalThreads), N, E);
#define EESEG_FOCAL(CM) focal_kernel<T, CM><<<grid, kFocalThreads, 0, stream>>>(logits, exit_stride, targets, N, C, HW, gamma, alpha, \
                                                                                 pixw, pw_es, pw_is, coef, loss_map, dlogits, part)
  if (C == 21) EESEG_FOCAL(21);
  else if (C == 19) EESEG_FOCAL(19);
  else if (C <= 32) EESEG_FOCAL(32);
  else EESEG_FOCAL(64);
#undef EESEG_FOCAL
  return check_launch("focal_kernel");
}

}  // namespace eeseg

using namespace eeseg;

extern "C" size_t eeseg_focal_workspace_bytes(int E, int N, int64_t HW) {
  if (E <= 0 || N <= 0 || HW <= 0) return 256;
  return (size_t)E * N * ((HW + kFocalThreads - 1) / kFocalThreads) * sizeof(double) + 256;
}

extern "C" int eeseg_focal_fwd(const void* logits, int dtype, int64_t exit_stride, const int64_t* targets, int E, int N, int C,
                               int64_t HW, float gamma, const float* alpha, const float* pixel_weight,
                               int64_t pw_exit_stride, int64_t pw_image_stride, const float* coef, float* per_exit_sum,
                               float* loss_map, void* dlogits, void* workspace, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  EESEG_REQUIRE(logits && targets && per_exit_sum && workspace, "focal_fwd: null pointer");
  EESEG_REQUIRE(E >= 1 && N >= 1 && C >= 1 && C <= 64 && HW >= 1, "focal_fwd: bad sizes (C <= 64)");
  EESEG_REQUIRE(N <= 65535 && E <= 65535 && HW < (1ll << 29), "focal_fwd: E, N <= 65535 and HW < 2^29");
  EESEG_REQUIRE(dtype == EESEG_F32 || dtype == EESEG_BF16, "focal_fwd: dtype %d", dtype);
  EESEG_REQUIRE(gamma >= 0.f, "focal_fwd: gamma must be >= 0");
  double* part = reinterpret_cast<double*>(workspace);
  int rc = dtype == EESEG_F32
               ? launch_focal<float>((const float*)logits, exit_stride, targets, E, N, C, HW, gamma, alpha, pixel_weight,
                                     pw_exit_stride, pw_image_stride, coef, loss_map, (float*)dlogits, part, stream)
               : launch_focal<__nv_bfloat16>((const __nv_bfloat16*)logits, exit_stride, targets, E, N, C, HW, gamma, alpha,
                                             pixel_weight, pw_exit_stride, pw_image_stride, coef, loss_map,
                                             (__nv_bfloat16*)dlogits, part, stream);
  if (rc) return rc;
  focal_finalize_kernel<<<E, 256, 0, stream>>>(part, N * (int)((HW + kFocalThreads - 1) / kFocalThreads), per_exit_sum);
  return check_launch("focal_finalize_kernel");
}
