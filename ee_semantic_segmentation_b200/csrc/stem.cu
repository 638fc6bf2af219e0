// ResNet stem helpers: space-to-depth + horizontal tap unrolling of the input image (so the 7x7/s2 conv
// runs on the tcgen05 implicit-GEMM kernel as a 4x1/s1 conv over 48 of 64 channels: 4 K blocks) and the 3x3/s2 max-pool on NHWC
// bf16. Both are pure streaming kernels (coalesced 16 B accesses). Reference: torchvision resnet
// conv1/bn1/relu/maxpool inside base_model[0] (from_deepv3_new.py:75-79,146).
#include "common.cuh"

namespace eeseg {

// image element -> float: fp32 / bf16 as stored; uint8 as ToTensor + Normalize would produce it,
// (u/255 - mean[c]) / std[c] = u * a[c] + b[c] (a, b folded on the host side of the launcher)
__device__ __forceinline__ float img_ld(const float* p, float, float) { return __ldg(p); }
__device__ __forceinline__ float img_ld(const __nv_bfloat16* p, float, float) { return __bfloat162float(__ldg(p)); }
__device__ __forceinline__ float img_ld(const uint8_t* p, float a, float b) { return fmaf((float)__ldg(p), a, b); }

struct StemNorm {
  float a[3], b[3];
};

template <typename T>
__global__ void __launch_bounds__(256) stem_s2d_kernel(const T* __restrict__ x, int N, int H, int W,
                                                        int H2, int W2, const StemNorm nm,
                                                        __nv_bfloat16* __restrict__ out) {
  // one thread per output pixel (n, Y, X): 4 horizontal taps x (2x2 space-to-depth x 3 channels) = 48
  // values from input rows 2Y, 2Y+1 and columns 2(X-2) .. 2(X+1)+1, then 16 zero channels
  const int64_t total = (int64_t)N * H2 * W2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int X = (int)(i % W2);
    const int Y = (int)((i / W2) % H2);
    const int n = (int)(i / ((int64_t)W2 * H2));
    uint32_t w[32];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float v[12];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const int yy = 2 * Y + a, xx = 2 * (X + u - 2) + b;
            v[(a * 2 + b) * 3 + c] =
                (yy < H && xx >= 0 && xx < W) ? img_ld(x + (((int64_t)n * 3 + c) * H + yy) * W + xx, nm.a[c], nm.b[c]) : 0.f;
          }
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        __nv_bfloat162 b2 = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
        w[u * 6 + k] = *reinterpret_cast<uint32_t*>(&b2);
      }
    }
#pragma unroll
    for (int k = 24; k < 32; ++k) w[k] = 0;
    uint4* o = reinterpret_cast<uint4*>(out + i * 64);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = make_uint4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
  }
}

__global__ void __launch_bounds__(256) maxpool3x3s2_kernel(const __nv_bfloat16* __restrict__ x, int N, int h,
                                                            int w, int C, int ho, int wo,
                                                            __nv_bfloat16* __restrict__ out) {
  // one thread per 16-byte chunk (8 channels) of an output pixel; grid = (chunks of an output row / 256, Y, n)
  const int cv = C / 8;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= wo * cv) return;
  const int X = i / cv, c8 = i - X * cv;
  const int Y = blockIdx.y, n = blockIdx.z;
  float m[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) m[k] = -INFINITY;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int yy = 2 * Y + dy, xx = 2 * X + dx;
      if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + (((int64_t)n * h + yy) * w + xx) * C) + c8);
        const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          m[2 * k] = fmaxf(m[2 * k], __uint_as_float(w4[k] << 16));
          m[2 * k + 1] = fmaxf(m[2 * k + 1], __uint_as_float(w4[k] & 0xffff0000u));
        }
      }
    }
  uint32_t o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    __nv_bfloat162 b2 = __floats2bfloat162_rn(m[2 * k], m[2 * k + 1]);
    o[k] = *reinterpret_cast<uint32_t*>(&b2);
  }
  reinterpret_cast<uint4*>(out + (((int64_t)n * ho + Y) * wo + X) * C)[c8] = make_uint4(o[0], o[1], o[2], o[3]);
}

// Training max-pool: forward also records which of the 9 window taps won (first maximum in row-major window order,
// ATen's rule: a later tap wins only if strictly greater), backward gathers: an input pixel belongs to at most 2x2
// windows and receives dOut of those whose recorded tap points at it — fixed order, no atomics.
__global__ void __launch_bounds__(256) maxpool3x3s2_idx_kernel(const __nv_bfloat16* __restrict__ x, int N, int h,
                                                                int w, int C, int ho, int wo,
                                                                __nv_bfloat16* __restrict__ out,
                                                                uint8_t* __restrict__ idx) {
  const int cv = C / 8;
  const int64_t total = (int64_t)N * ho * wo * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cv);
    const int X = (int)((i / cv) % wo);
    const int Y = (int)((i / ((int64_t)cv * wo)) % ho);
    const int n = (int)(i / ((int64_t)cv * wo * ho));
    float m[8];
    uint32_t bits[8];
    int arg[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { m[k] = -INFINITY; bits[k] = 0xff80u; arg[k] = 0; }
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int yy = 2 * Y + dy, xx = 2 * X + dx;
        if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
          const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + (((int64_t)n * h + yy) * w + xx) * C) + c8);
          const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
          const int tap = (dy + 1) * 3 + (dx + 1);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint32_t b = (k & 1) ? (w4[k >> 1] >> 16) : (w4[k >> 1] & 0xffffu);
            const float v = __uint_as_float(b << 16);
            if (v > m[k] || v != v) { m[k] = v; bits[k] = b; arg[k] = tap; }
          }
        }
      }
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = bits[2 * k] | (bits[2 * k + 1] << 16);
    const int64_t opix = ((int64_t)n * ho + Y) * wo + X;
    reinterpret_cast<uint4*>(out + opix * C)[c8] = make_uint4(o[0], o[1], o[2], o[3]);
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { lo |= (uint32_t)arg[k] << (8 * k); hi |= (uint32_t)arg[4 + k] << (8 * k); }
    reinterpret_cast<uint2*>(idx + opix * C)[c8] = make_uint2(lo, hi);
  }
}

__global__ void __launch_bounds__(256) maxpool3x3s2_bwd_kernel(const __nv_bfloat16* __restrict__ dout,
                                                                const uint8_t* __restrict__ idx, int N, int h, int w,
                                                                int C, int ho, int wo, __nv_bfloat16* __restrict__ dx) {
  const int cv = C / 8;
  const int64_t total = (int64_t)N * h * w * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cv);
    const int x = (int)((i / cv) % w);
    const int y = (int)((i / ((int64_t)cv * w)) % h);
    const int n = (int)(i / ((int64_t)cv * w * h));
    float g[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) g[k] = 0.f;
    // windows (Y, X) with |2Y - y| <= 1: Y in [ceil((y-1)/2), floor((y+1)/2)]
    const int Y0 = y >> 1, Y1 = (y + 1) >> 1, X0 = x >> 1, X1 = (x + 1) >> 1;   // equal when y (x) is even
    for (int Y = Y0; Y <= Y1; ++Y) {
      if (Y >= ho) continue;
      for (int X = X0; X <= X1; ++X) {
        if (X >= wo) continue;
        const int tap = (y - 2 * Y + 1) * 3 + (x - 2 * X + 1);
        const int64_t opix = ((int64_t)n * ho + Y) * wo + X;
        const uint2 a = __ldg(reinterpret_cast<const uint2*>(idx + opix * C) + c8);
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(dout + opix * C) + c8);
        const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int ak = (int)(((k < 4 ? a.x : a.y) >> (8 * (k & 3))) & 0xffu);
          const uint32_t b = (k & 1) ? (w4[k >> 1] & 0xffff0000u) : (w4[k >> 1] << 16);
          if (ak == tap) g[k] += __uint_as_float(b);
        }
      }
    }
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      __nv_bfloat162 b2 = __floats2bfloat162_rn(g[2 * k], g[2 * k + 1]);
      o[k] = *reinterpret_cast<uint32_t*>(&b2);
    }
    reinterpret_cast<uint4*>(dx + (((int64_t)n * h + y) * w + x) * C)[c8] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

}  // namespace eeseg

using namespace eeseg;

extern "C" int eeseg_maxpool3x3s2_nhwc_train(const void* x, int N, int h, int w, int C, void* out, void* idx, void* stream) {
  EESEG_REQUIRE(x && out && idx, "maxpool3x3s2_train: null pointer");
  EESEG_REQUIRE(C % 8 == 0 && (((uintptr_t)x | (uintptr_t)out) & 15) == 0 && ((uintptr_t)idx & 7) == 0,
                "maxpool3x3s2_train: C %% 8 == 0 and aligned pointers required");
  if (N <= 0) return EESEG_OK;
  const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
  const int64_t total = (int64_t)N * ho * wo * (C / 8);
  const int blocks = (int)((total + 255) / 256 < kNumSMs * 8 ? (total + 255) / 256 : kNumSMs * 8);
  maxpool3x3s2_idx_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, N, h, w, C, ho, wo,
                                                                      (__nv_bfloat16*)out, (uint8_t*)idx);
  return check_launch("maxpool3x3s2_idx_kernel");
}

extern "C" int eeseg_maxpool3x3s2_nhwc_bwd(const void* dout, const void* idx, int N, int h, int w, int C, void* dx,
                                           void* stream) {
  EESEG_REQUIRE(dout && idx && dx, "maxpool3x3s2_bwd: null pointer");
  EESEG_REQUIRE(C % 8 == 0 && (((uintptr_t)dout | (uintptr_t)dx) & 15) == 0 && ((uintptr_t)idx & 7) == 0,
                "maxpool3x3s2_bwd: C %% 8 == 0 and aligned pointers required");
  if (N <= 0) return EESEG_OK;
  const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
  const int64_t total = (int64_t)N * h * w * (C / 8);
  const int blocks = (int)((total + 255) / 256 < kNumSMs * 16 ? (total + 255) / 256 : kNumSMs * 16);
  maxpool3x3s2_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dout, (const uint8_t*)idx, N, h, w,
                                                                      C, ho, wo, (__nv_bfloat16*)dx);
  return check_launch("maxpool3x3s2_bwd_kernel");
}

extern "C" int eeseg_stem_space_to_depth_any(const void* x, int x_kind, const float* mean, const float* std_, int N, int H,
                                             int W, void* out, void* stream) {
  EESEG_REQUIRE(x && out, "stem_space_to_depth: null pointer");
  EESEG_REQUIRE(((uintptr_t)out & 15) == 0, "stem_space_to_depth: output must be 16-byte aligned");
  EESEG_REQUIRE(x_kind == EESEG_F32 || x_kind == EESEG_BF16 || x_kind == EESEG_U8, "stem_space_to_depth: image dtype %d", x_kind);
  if (N <= 0) return EESEG_OK;
  const int H2 = (H + 1) / 2, W2 = (W + 1) / 2;
  const int64_t total = (int64_t)N * H2 * W2;
  const int blocks = (int)((total + 255) / 256 < kNumSMs * 8 ? (total + 255) / 256 : kNumSMs * 8);
  StemNorm nm;
  for (int c = 0; c < 3; ++c) {   // host arrays (3 floats each); NULL = plain u/255
    const float m = mean ? mean[c] : 0.f, sd = std_ ? std_[c] : 1.f;
    EESEG_REQUIRE(sd != 0.f, "stem_space_to_depth: zero std");
    nm.a[c] = 1.f / (255.f * sd);
    nm.b[c] = -m / sd;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (x_kind == EESEG_F32)
    stem_s2d_kernel<float><<<blocks, 256, 0, st>>>((const float*)x, N, H, W, H2, W2, nm, (__nv_bfloat16*)out);
  else if (x_kind == EESEG_BF16)
    stem_s2d_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)x, N, H, W, H2, W2, nm, (__nv_bfloat16*)out);
  else
    stem_s2d_kernel<uint8_t><<<blocks, 256, 0, st>>>((const uint8_t*)x, N, H, W, H2, W2, nm, (__nv_bfloat16*)out);
  return check_launch("stem_s2d_kernel");
}

extern "C" int eeseg_stem_space_to_depth(const float* x, int N, int H, int W, void* out, void* stream) {
  return eeseg_stem_space_to_depth_any(x, EESEG_F32, nullptr, nullptr, N, H, W, out, stream);
}

extern "C" int eeseg_maxpool3x3s2_nhwc(const void* x, int N, int h, int w, int C, void* out, void* stream) {
  EESEG_REQUIRE(x && out, "maxpool3x3s2: null pointer");
  EESEG_REQUIRE(C % 8 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)out & 15) == 0,
                "maxpool3x3s2: C %% 8 == 0 and 16-byte aligned pointers required");
  if (N <= 0) return EESEG_OK;
  const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
  EESEG_REQUIRE(ho <= 65535 && N <= 65535, "maxpool3x3s2: map too tall / batch too large for the grid");
  const dim3 blocks((unsigned)((wo * (C / 8) + 255) / 256), (unsigned)ho, (unsigned)N);
  maxpool3x3s2_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, N, h, w, C, ho, wo,
                                                                  (__nv_bfloat16*)out);
  return check_launch("maxpool3x3s2_kernel");
}

// ---- small dense layer for the ASPP pooled branch -----------------------------------------------
// y[n][o] = act( (sum_k x[n][k] * W[o][k]) * scale[o] + shift[o] ), one warp per output element.
// (AdaptiveAvgPool -> 1x1 conv -> BN -> ReLU, and its share of the ASPP projection: two launches
// instead of a dozen tiny library kernels.)
namespace eeseg {
__global__ void __launch_bounds__(256) dense_bn_act_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                            const float* __restrict__ scale,
                                                            const float* __restrict__ shift, int N, int K, int O,
                                                            int relu, float* __restrict__ y) {
  const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (wid >= N * O) return;
  const int n = wid / O, o = wid % O;
  const float* xr = x + (int64_t)n * K;
  const float* wr = W + (int64_t)o * K;
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) acc = fmaf(__ldg(xr + k), __ldg(wr + k), acc);
  acc = warp_sum(acc);
  if (lane == 0) {
    float v = acc * (scale ? scale[o] : 1.f) + (shift ? shift[o] : 0.f);
    y[(int64_t)n * O + o] = relu ? fmaxf(v, 0.f) : v;
  }
}
}  // namespace eeseg

extern "C" int eeseg_dense_bn_act(const float* x, const float* W, const float* scale, const float* shift, int N,
                                  int K, int O, int relu, float* y, void* stream) {
  EESEG_REQUIRE(x && W && y, "dense_bn_act: null pointer");
  if (N <= 0 || O <= 0) return EESEG_OK;
  const int warps = N * O;
  eeseg::dense_bn_act_kernel<<<(warps + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, W, scale, shift, N, K, O, relu, y);
  return eeseg::check_launch("dense_bn_act_kernel");
}
