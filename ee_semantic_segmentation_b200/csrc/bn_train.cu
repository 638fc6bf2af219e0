// Training-mode BatchNorm2d (+ residual add) (+ ReLU) on NHWC bf16 activations, forward and backward.
// Reference contract: the nn.BatchNorm2d / ReLU / `out += identity` modules of torchvision's DeepLabHead, ASPP and
// ResNet Bottleneck as they run under net.train() inside train_epoch (train_funcs.py:12-33): batch statistics
// over N*h*w per channel, biased variance for the normalisation, running statistics updated with `momentum`
// and the unbiased variance, eps inside the square root.
//
// HBM-bound: x is [P = N*h*w pixels][C channels] bf16 with the channels contiguous, so a thread owns 8
// channels (one 16 B load) of a pixel and a warp reads 4 pixels x 128 B. Four kernels, each one streaming pass:
//   bn_stats      : per-channel sum / sum of squares (fp32 per thread, ordered fp64 finalize) -> mean, invstd,
//                   running-stat update, and the folded a = gamma*invstd, b = beta - mean*a
//   bn_apply      : y = act(a*x + b (+ residual))                         read 2(+2) B, write 2 B per element
//   bn_bwd_reduce : dbeta = sum dy', dgamma = sum dy'*xhat, dy' = dy * [y > 0]
//   bn_bwd_apply  : dx = a*(dy' - dbeta/P - xhat*dgamma/P), dres = dy'
// Fixed summation order everywhere (bit-reproducible), no atomics.
#include "common.cuh"

namespace eeseg {

constexpr int kBnThreads = 256;      // 8 channel groups (8 channels each) x 32 pixel lanes
constexpr int kBnMaxSplits = 512;

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    f[2 * k] = __uint_as_float(w[k] << 16);
    f[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    __nv_bfloat162 b = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
    w[k] = *reinterpret_cast<uint32_t*>(&b);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ uint4 ld16_stream(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// Per-channel reduction of two quantities over a slice of the pixels. MODE 0: (x, x*x). MODE 1: (dy', dy'*xhat).
// grid (C/64, splits); partial[(split*2 + q)*C + c].
template <int MODE>
__global__ void __launch_bounds__(kBnThreads) bn_reduce_kernel(const __nv_bfloat16* __restrict__ x,
                                                               const __nv_bfloat16* __restrict__ dy,
                                                               const __nv_bfloat16* __restrict__ y, int64_t P, int C,
                                                               const float* __restrict__ mean, const float* __restrict__ invstd,
                                                               int relu, float* __restrict__ partial) {
  const int cg = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int c = blockIdx.x * 64 + cg * 8;
  const int splits = gridDim.y;
  const int64_t per = (P + splits - 1) / splits;
  const int64_t p0 = (int64_t)blockIdx.y * per, p1 = min(p0 + per, P);
  float s0[8], s1[8], m[8], is[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { s0[k] = 0.f; s1[k] = 0.f; }
  if (MODE == 1) {
#pragma unroll
    for (int k = 0; k < 8; ++k) { m[k] = mean[c + k]; is[k] = invstd[c + k]; }
  }
#pragma unroll 2
  for (int64_t p = p0 + pl; p < p1; p += 32) {
    float xv[8];
    unpack8(ld16_stream(x + p * C + c), xv);
    if (MODE == 0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) { s0[k] += xv[k]; s1[k] = fmaf(xv[k], xv[k], s1[k]); }
    } else {
      float g[8], yv[8];
      unpack8(ld16_stream(dy + p * C + c), g);
      if (relu) unpack8(ld16_stream(y + p * C + c), yv);   // mask from the forward output
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float gg = (!relu || yv[k] > 0.f) ? g[k] : 0.f;
        s0[k] += gg;
        s1[k] = fmaf(gg, (xv[k] - m[k]) * is[k], s1[k]);
      }
    }
  }
  __shared__ float sm[2][32][65];
#pragma unroll
  for (int k = 0; k < 8; ++k) { sm[0][pl][cg * 8 + k] = s0[k]; sm[1][pl][cg * 8 + k] = s1[k]; }
  __syncthreads();
  if (threadIdx.x < 128) {
    const int q = threadIdx.x >> 6, cc = threadIdx.x & 63;
    float t = 0.f;
    for (int i = 0; i < 32; ++i) t += sm[q][i][cc];   // fixed order
    partial[((int64_t)blockIdx.y * 2 + q) * C + blockIdx.x * 64 + cc] = t;
  }
}

// Sum of the `splits` partials of quantity q for channel c by ONE WARP in a fixed order: lane l adds partials
// l, l+32, ... in fp64, then a fixed shuffle tree. Every lane returns the total.
__device__ __forceinline__ double bn_warp_total(const float* __restrict__ partial, int splits, int C, int c, int q) {
  double t = 0.0;
  for (int i = (int)(threadIdx.x & 31); i < splits; i += 32) t += (double)partial[((int64_t)i * 2 + q) * C + c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;
}

// mean / invstd / folded scale-shift / running statistics from the partial sums (one warp per channel)
__global__ void bn_stats_finalize_kernel(const float* __restrict__ partial, int splits, int C, int64_t P,
                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                         float* __restrict__ running_mean, float* __restrict__ running_var, float momentum,
                                         float eps, float* __restrict__ save_mean, float* __restrict__ save_invstd,
                                         float* __restrict__ fa, float* __restrict__ fb) {
  const int c = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (c >= C) return;
  const double s = bn_warp_total(partial, splits, C, c, 0), ss = bn_warp_total(partial, splits, C, c, 1);
  if ((threadIdx.x & 31) != 0) return;
  const double mean = s / (double)P;
  double var = ss / (double)P - mean * mean;
  if (var < 0.0) var = 0.0;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  save_mean[c] = (float)mean;
  save_invstd[c] = invstd;
  const float a = (gamma ? gamma[c] : 1.f) * invstd;
  fa[c] = a;
  fb[c] = (beta ? beta[c] : 0.f) - (float)mean * a;
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
  if (running_var) {
    const double unbiased = P > 1 ? var * (double)P / (double)(P - 1) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

__global__ void __launch_bounds__(256) bn_apply_kernel(const __nv_bfloat16* __restrict__ x,
                                                       const __nv_bfloat16* __restrict__ res, int64_t total8, int C,
                                                       const float* __restrict__ fa, const float* __restrict__ fb, int relu,
                                                       __nv_bfloat16* __restrict__ y) {
  // total8 = P*C/8 groups of 8 channels. The host sizes the grid so that the grid stride is a multiple of C/8:
  // a thread then keeps the same 8 channels for its whole loop and their constants stay in registers
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int c8 = C >> 3;
  const int c = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) % c8) * 8;
  const float4 a0 = __ldg(reinterpret_cast<const float4*>(fa + c)), a1 = __ldg(reinterpret_cast<const float4*>(fa + c + 4));
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(fb + c)), b1 = __ldg(reinterpret_cast<const float4*>(fb + c + 4));
  const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
  const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll 2
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += stride) {
    float xv[8], rv[8], o[8];
    unpack8(ld16_stream(x + i * 8), xv);
    if (res) unpack8(ld16_stream(res + i * 8), rv);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float v = fmaf(av[k], xv[k], bv[k]);
      if (res) v += rv[k];
      o[k] = relu ? fmaxf(v, 0.f) : v;
    }
    *reinterpret_cast<uint4*>(y + i * 8) = pack8(o);
  }
}

// dgamma / dbeta and the two per-channel constants of the input gradient
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partial, int splits, int C, int64_t P,
                                       const float* __restrict__ gamma, const float* __restrict__ mean,
                                       const float* __restrict__ invstd, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ c1,
                                       float* __restrict__ c2, float* __restrict__ fa, int accumulate) {
  const int c = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (c >= C) return;
  const double s = bn_warp_total(partial, splits, C, c, 0), ss = bn_warp_total(partial, splits, C, c, 1);
  if ((threadIdx.x & 31) != 0) return;
  // accumulate: dgamma / dbeta ARE the parameters' .grad tensors (one launch less per BatchNorm than AccumulateGrad's add_)
  if (dbeta) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + (float)s;
  if (dgamma) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + (float)ss;
  // dx = a*(dy' - s/P - xhat*ss/P) = a*dy' + k1 + k2*x with xhat = (x - mean)*invstd
  const double a = (double)(gamma ? gamma[c] : 1.f) * (double)invstd[c];
  const double k2 = -a * (ss / (double)P) * (double)invstd[c];
  fa[c] = (float)a;
  c1[c] = (float)(-a * (s / (double)P) - k2 * (double)mean[c]);
  c2[c] = (float)k2;
}

__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy,
                                                           const __nv_bfloat16* __restrict__ x,
                                                           const __nv_bfloat16* __restrict__ y, int64_t total8, int C,
                                                           const float* __restrict__ fa, const float* __restrict__ k1,
                                                           const float* __restrict__ k2, int relu,
                                                           __nv_bfloat16* __restrict__ dx, __nv_bfloat16* __restrict__ dres) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;   // a multiple of C/8 (see bn_apply_kernel)
  const int c8 = C >> 3;
  const int c = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) % c8) * 8;
  const float4 a0 = __ldg(reinterpret_cast<const float4*>(fa + c)), a1 = __ldg(reinterpret_cast<const float4*>(fa + c + 4));
  const float4 p0 = __ldg(reinterpret_cast<const float4*>(k1 + c)), p1 = __ldg(reinterpret_cast<const float4*>(k1 + c + 4));
  const float4 q0 = __ldg(reinterpret_cast<const float4*>(k2 + c)), q1 = __ldg(reinterpret_cast<const float4*>(k2 + c + 4));
  const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
  const float pv[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
  const float qv[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll 2
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += stride) {
    float g[8], xv[8], yv[8], o[8], r[8];
    unpack8(ld16_stream(dy + i * 8), g);
    unpack8(ld16_stream(x + i * 8), xv);
    if (relu) unpack8(ld16_stream(y + i * 8), yv);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float gg = (!relu || yv[k] > 0.f) ? g[k] : 0.f;
      o[k] = fmaf(av[k], gg, fmaf(qv[k], xv[k], pv[k]));
      r[k] = gg;
    }
    *reinterpret_cast<uint4*>(dx + i * 8) = pack8(o);
    if (dres) *reinterpret_cast<uint4*>(dres + i * 8) = pack8(r);
  }
}

// number of 256-thread blocks of the apply kernels: about 16 per SM, and blocks*256 a multiple of C/8
static unsigned bn_apply_blocks(int64_t total8, int C) {
  const int c8 = C >> 3;
  int64_t unit = 1;                       // smallest block count with (unit*256) % c8 == 0
  while ((unit * 256) % c8) ++unit;
  int64_t blocks = (total8 + 255) / 256;
  const int64_t cap = (int64_t)kNumSMs * 16;
  if (blocks > cap) blocks = cap;
  blocks = (blocks + unit - 1) / unit * unit;
  return (unsigned)blocks;
}

static int bn_splits(int64_t P, int C) {
  int64_t s = (int64_t)kNumSMs * 8 / (C / 64);
  if (s < 1) s = 1;
  if (s > kBnMaxSplits) s = kBnMaxSplits;
  const int64_t cap = (P + 127) / 128;   // at least 128 pixels per split
  if (s > cap) s = cap;
  return (int)(s < 1 ? 1 : s);
}

}  // namespace eeseg

using namespace eeseg;

// workspace: partial sums [kBnMaxSplits][2][C] + fa, fb, c1, c2 [C] each
extern "C" size_t eeseg_bn_train_workspace_bytes(int C) {
  if (C <= 0) return 256;
  return ((size_t)kBnMaxSplits * 2 + 4) * (size_t)C * sizeof(float) + 256;
}

static void bn_ws(void* workspace, int C, float*& partial, float*& fa, float*& fb, float*& c1, float*& c2) {
  partial = reinterpret_cast<float*>(workspace);
  fa = partial + (size_t)kBnMaxSplits * 2 * C;
  fb = fa + C; c1 = fb + C; c2 = c1 + C;
}

extern "C" int eeseg_bn_train_fwd(const void* x, int64_t P, int C, const float* gamma, const float* beta, float* running_mean,
                                  float* running_var, float momentum, float eps, int relu, const void* residual, void* y,
                                  float* save_mean, float* save_invstd, void* workspace, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  EESEG_REQUIRE(x && y && save_mean && save_invstd && workspace, "bn_train_fwd: null pointer");
  EESEG_REQUIRE(P >= 1 && C >= 64 && C % 64 == 0, "bn_train_fwd: C=%d must be a positive multiple of 64", C);
  EESEG_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)residual & 15) == 0 &&
                ((uintptr_t)workspace & 15) == 0, "bn_train_fwd: 16-byte aligned tensors required");
  float *partial, *fa, *fb, *c1, *c2;
  bn_ws(workspace, C, partial, fa, fb, c1, c2);
  const int splits = bn_splits(P, C);
  bn_reduce_kernel<0><<<dim3(C / 64, splits), kBnThreads, 0, stream>>>((const __nv_bfloat16*)x, nullptr, nullptr, P, C, nullptr,
                                                                       nullptr, 0, partial);
  int rc = check_launch("bn_reduce_kernel<stats>");
  if (rc) return rc;
  bn_stats_finalize_kernel<<<(C + 7) / 8, 256, 0, stream>>>(partial, splits, C, P, gamma, beta, running_mean, running_var,
                                                              momentum, eps, save_mean, save_invstd, fa, fb);
  rc = check_launch("bn_stats_finalize_kernel");
  if (rc) return rc;
  const int64_t total8 = P * C / 8;
  bn_apply_kernel<<<bn_apply_blocks(total8, C), 256, 0, stream>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)residual, total8, C, fa, fb,
                                                       relu, (__nv_bfloat16*)y);
  return check_launch("bn_apply_kernel");
}

static int bn_bwd_launch(const void* dy, const void* x, const void* y, int64_t P, int C, const float* gamma,
                         const float* save_mean, const float* save_invstd, int relu, void* dx, void* dres, float* dgamma,
                         float* dbeta, int accumulate, void* workspace, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  EESEG_REQUIRE(dy && x && save_mean && save_invstd && dx && workspace, "bn_train_bwd: null pointer");
  EESEG_REQUIRE(P >= 1 && C >= 64 && C % 64 == 0, "bn_train_bwd: C=%d must be a positive multiple of 64", C);
  EESEG_REQUIRE(!relu || y, "bn_train_bwd: the ReLU mask needs the forward output y");
  EESEG_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)dy & 15) == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)dx & 15) == 0 &&
                ((uintptr_t)dres & 15) == 0 && ((uintptr_t)workspace & 15) == 0, "bn_train_bwd: 16-byte aligned tensors required");
  float *partial, *fa, *fb, *c1, *c2;
  bn_ws(workspace, C, partial, fa, fb, c1, c2);
  const int splits = bn_splits(P, C);
  bn_reduce_kernel<1><<<dim3(C / 64, splits), kBnThreads, 0, stream>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy,
                                                                       (const __nv_bfloat16*)y, P, C, save_mean, save_invstd,
                                                                       relu, partial);
  int rc = check_launch("bn_reduce_kernel<bwd>");
  if (rc) return rc;
  bn_bwd_finalize_kernel<<<(C + 7) / 8, 256, 0, stream>>>(partial, splits, C, P, gamma, save_mean, save_invstd, dgamma, dbeta,
                                                              c1, c2, fa, accumulate);
  rc = check_launch("bn_bwd_finalize_kernel");
  if (rc) return rc;
  const int64_t total8 = P * C / 8;
  bn_bwd_apply_kernel<<<bn_apply_blocks(total8, C), 256, 0, stream>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x,
                                                           (const __nv_bfloat16*)y, total8, C, fa, c1, c2, relu,
                                                           (__nv_bfloat16*)dx, (__nv_bfloat16*)dres);
  return check_launch("bn_bwd_apply_kernel");
}

extern "C" int eeseg_bn_train_bwd(const void* dy, const void* x, const void* y, int64_t P, int C, const float* gamma,
                                  const float* save_mean, const float* save_invstd, int relu, void* dx, void* dres,
                                  float* dgamma, float* dbeta, void* workspace, void* stream_) {
  return bn_bwd_launch(dy, x, y, P, C, gamma, save_mean, save_invstd, relu, dx, dres, dgamma, dbeta, 0, workspace, stream_);
}

extern "C" int eeseg_bn_train_bwd_acc(const void* dy, const void* x, const void* y, int64_t P, int C, const float* gamma,
                                      const float* save_mean, const float* save_invstd, int relu, void* dx, void* dres,
                                      float* gamma_grad, float* beta_grad, void* workspace, void* stream_) {
  EESEG_REQUIRE(gamma_grad && beta_grad, "bn_train_bwd_acc: null gradient tensors");
  return bn_bwd_launch(dy, x, y, P, C, gamma, save_mean, save_invstd, relu, dx, dres, gamma_grad, beta_grad, 1, workspace, stream_);
}
