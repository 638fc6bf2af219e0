// Small training kernels that complete the step on eeseg code (the reference's step is train_funcs.py:12-33 around
// torchvision's DeepLabHead and torch.optim.SGD, deepv3_funcs.py:74-101):
//   * dropout (ASPP.project's Dropout(0.5), torchvision deeplabv3.py:101) with a counter-based generator whose seed and
//     offset live in DEVICE memory, so a captured CUDA graph draws a fresh mask on every replay;
//   * the multi-tensor SGD update (momentum, weight decay; torch.optim.SGD semantics) in ONE launch over a table of
//     parameter chunks, learning rates read from device memory (a scheduler never invalidates the captured graph);
//   * dense helpers of the pooled ASPP branch's backward (outer-product weight gradient, input gradient) and per-image
//     channel sums of an NHWC tensor (bias / shift gradients).
// All streaming or tiny; no atomics; fixed summation order.
#include "common.cuh"

namespace eeseg {

// ---- counter-based RNG: 2 rounds of a 64-bit mix (splitmix64 finaliser) over (seed, element index / 8) --------------
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

// x, y: bf16, 8 elements (16 B) per thread; mask: one byte per 8 elements (bit k = element k kept).
// keep probability = 1 - p as a 8-bit threshold on 8 independent bytes of one 64-bit draw (p = 0.5 is exact).
__global__ void __launch_bounds__(256) dropout_fwd_kernel(const uint4* __restrict__ x, int64_t n8, uint32_t keep_thr,
                                                           float scale, const uint64_t* __restrict__ state,
                                                           uint4* __restrict__ y, uint8_t* __restrict__ mask) {
  const uint64_t seed = state[0], offset = state[1];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t r = mix64(mix64(seed ^ 0x9e3779b97f4a7c15ull) + (offset + (uint64_t)i) * 0xd1342543de82ef95ull);
    const uint4 v = __ldg(x + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[4];
    uint8_t m = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const bool k0 = ((r >> (16 * k)) & 0xff) < keep_thr, k1 = ((r >> (16 * k + 8)) & 0xff) < keep_thr;
      m |= (uint8_t)((k0 ? 1 : 0) << (2 * k)) | (uint8_t)((k1 ? 1 : 0) << (2 * k + 1));
      const float a = k0 ? __uint_as_float(w[k] << 16) * scale : 0.f;
      const float b = k1 ? __uint_as_float(w[k] & 0xffff0000u) * scale : 0.f;
      __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
      o[k] = *reinterpret_cast<uint32_t*>(&p);
    }
    y[i] = make_uint4(o[0], o[1], o[2], o[3]);
    mask[i] = m;
  }
}

__global__ void dropout_advance_kernel(uint64_t* state, uint64_t n8) { state[1] += n8; }

__global__ void __launch_bounds__(256) dropout_bwd_kernel(const uint4* __restrict__ dy, const uint8_t* __restrict__ mask,
                                                           int64_t n8, float scale, uint4* __restrict__ dx) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 v = __ldg(dy + i);
    const uint8_t m = __ldg(mask + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float a = ((m >> (2 * k)) & 1) ? __uint_as_float(w[k] << 16) * scale : 0.f;
      const float b = ((m >> (2 * k + 1)) & 1) ? __uint_as_float(w[k] & 0xffff0000u) * scale : 0.f;
      __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
      o[k] = *reinterpret_cast<uint32_t*>(&p);
    }
    dx[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ---- multi-tensor SGD ------------------------------------------------------------------------------------------
struct SgdChunk {          // one block's work: `count` consecutive elements of one tensor
  float* p;
  const float* g;
  float* buf;
  int32_t count;
  int32_t group;           // index into the learning-rate array
};

// torch.optim.SGD (dampening 0, no nesterov): d = g + wd * p; buf = momentum * buf + d; p -= lr[group] * buf.
// A zero-initialised buffer makes the first step buf = d, as torch's clone() does.
__global__ void __launch_bounds__(256) sgd_multi_kernel(const SgdChunk* __restrict__ chunks, const float* __restrict__ lrs,
                                                         float momentum, float wd) {
  const SgdChunk c = chunks[blockIdx.x];
  const float lr = lrs[c.group];
  const int n4 = c.count >> 2;
  const bool vec = ((((uintptr_t)c.p | (uintptr_t)c.g | (uintptr_t)c.buf) & 15) == 0);
  if (vec) {
    float4* p4 = reinterpret_cast<float4*>(c.p);
    const float4* g4 = reinterpret_cast<const float4*>(c.g);
    float4* b4 = reinterpret_cast<float4*>(c.buf);
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      float4 p = p4[i], b = b4[i];
      const float4 g = g4[i];
      b.x = fmaf(momentum, b.x, fmaf(wd, p.x, g.x)); p.x = fmaf(-lr, b.x, p.x);
      b.y = fmaf(momentum, b.y, fmaf(wd, p.y, g.y)); p.y = fmaf(-lr, b.y, p.y);
      b.z = fmaf(momentum, b.z, fmaf(wd, p.z, g.z)); p.z = fmaf(-lr, b.z, p.z);
      b.w = fmaf(momentum, b.w, fmaf(wd, p.w, g.w)); p.w = fmaf(-lr, b.w, p.w);
      p4[i] = p; b4[i] = b;
    }
  }
  for (int i = (vec ? n4 * 4 : 0) + threadIdx.x; i < c.count; i += blockDim.x) {
    const float b = fmaf(momentum, c.buf[i], fmaf(wd, c.p[i], c.g[i]));
    c.buf[i] = b;
    c.p[i] = fmaf(-lr, b, c.p[i]);
  }
}

// ---- dense backward (pooled ASPP branch: y[n][o] = x[n] . W[o]) ---------------------------------------------------
// dW[o][k] = sum_n dy[n][o] * x[n][k]  (N is the batch: a handful of terms, fixed order)
__global__ void __launch_bounds__(256) dense_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ x, int N,
                                                           int K, int O, float* __restrict__ dW) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)O * K) return;
  const int o = (int)(i / K), k = (int)(i % K);
  float acc = 0.f;
  for (int n = 0; n < N; ++n) acc = fmaf(__ldg(dy + (int64_t)n * O + o), __ldg(x + (int64_t)n * K + k), acc);
  dW[i] = acc;
}
// dx[n][k] = sum_o dy[n][o] * W[o][k]
__global__ void __launch_bounds__(256) dense_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ W, int N,
                                                           int K, int O, float* __restrict__ dx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)N * K) return;
  const int n = (int)(i / K), k = (int)(i % K);
  float acc = 0.f;
  for (int o = 0; o < O; ++o) acc = fmaf(__ldg(dy + (int64_t)n * O + o), __ldg(W + (int64_t)o * K + k), acc);
  dx[i] = acc;
}

// out[n][h][w][c] = bf16(v[n][c] * s): the broadcast of a per-image vector over the map (ASPPPooling's "bilinear"
// up-sampling of a 1x1 map is this constant; its backward with s = 1/HW is the global average pool's)
__global__ void __launch_bounds__(256) broadcast_rows_kernel(const float* __restrict__ v, int64_t hw, int C, float s,
                                                              __nv_bfloat16* __restrict__ out, int64_t total8) {
  const int c8n = C >> 3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % c8n);
    const int64_t n = i / ((int64_t)c8n * hw);
    const float* src = v + n * C + c8 * 8;
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      __nv_bfloat162 p = __floats2bfloat162_rn(__ldg(src + 2 * k) * s, __ldg(src + 2 * k + 1) * s);
      o[k] = *reinterpret_cast<uint32_t*>(&p);
    }
    reinterpret_cast<uint4*>(out)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ---- BatchNorm over a handful of fp32 row vectors (the pooled ASPP branch: batch statistics over the N pooled
// vectors, nn.BatchNorm2d on an [N,C,1,1] tensor) + ReLU. One thread per channel, fp32 throughout: with N = 2..8 samples
// the normalised values sit near +-1 and the input gradient is a small difference of large terms — bf16 inputs lose it.
__global__ void __launch_bounds__(256) bn_rows_fwd_kernel(const float* __restrict__ x, int N, int C, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float* __restrict__ running_mean,
                                                           float* __restrict__ running_var, float momentum, float eps, int relu,
                                                           float* __restrict__ y, float* __restrict__ save_mean,
                                                           float* __restrict__ save_invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float m = 0.f;
  for (int n = 0; n < N; ++n) m += x[(int64_t)n * C + c];
  m /= (float)N;
  float v = 0.f;
  for (int n = 0; n < N; ++n) {
    const float d = x[(int64_t)n * C + c] - m;
    v = fmaf(d, d, v);
  }
  const float var = v / (float)N;
  const float invstd = rsqrtf(var + eps);
  const float a = gamma[c] * invstd, b = beta[c] - m * a;
  for (int n = 0; n < N; ++n) {
    const float o = fmaf(a, x[(int64_t)n * C + c], b);
    y[(int64_t)n * C + c] = relu ? fmaxf(o, 0.f) : o;
  }
  save_mean[c] = m;
  save_invstd[c] = invstd;
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * m;
  if (running_var) running_var[c] = (1.f - momentum) * running_var[c] + momentum * (N > 1 ? v / (float)(N - 1) : var);
}

__global__ void __launch_bounds__(256) bn_rows_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                           const float* __restrict__ y, int N, int C,
                                                           const float* __restrict__ gamma, const float* __restrict__ mean,
                                                           const float* __restrict__ invstd, int relu, float* __restrict__ dx,
                                                           float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float m = mean[c], is = invstd[c];
  float sb = 0.f, sg = 0.f;
  for (int n = 0; n < N; ++n) {
    const int64_t i = (int64_t)n * C + c;
    const float g = (relu && y[i] <= 0.f) ? 0.f : dy[i];
    sb += g;
    sg = fmaf(g, (x[i] - m) * is, sg);
  }
  dgamma[c] = sg;
  dbeta[c] = sb;
  const float k = gamma[c] * is / (float)N;
  for (int n = 0; n < N; ++n) {
    const int64_t i = (int64_t)n * C + c;
    const float g = (relu && y[i] <= 0.f) ? 0.f : dy[i];
    dx[i] = k * ((float)N * g - sb - (x[i] - m) * is * sg);
  }
}

// ---- per-step weight preparation for ALL convolutions in one launch ---------------------------------------------------
// A training step needs every conv weight twice in bf16: [Cout][R][S][Cin] (forward, weight gradient) and the 180-degree
// rotated transpose [Cin][R][S][Cout] (input gradient = the forward kernel on it). Doing that per layer costs three small
// launches per convolution per step (layout/precision copy, rotation, unit scale/shift fill); this kernel walks a table of
// 32 x 32 (co, ci) tiles over all layers: coalesced fp32 reads of the parameter [Cout][Cin][R][S], both bf16 layouts
// written as 64-byte runs through one shared-memory tile.
struct WeightPrepTile {
  const float* src;          // parameter [Cout][Cin][RS]
  __nv_bfloat16* krsc;       // [Cout][RS][Cin]
  __nv_bfloat16* rot;        // [Cin][RS][Cout], taps reversed
  int32_t Cout, Cin, RS, co0, ci0, pad;
};

__global__ void __launch_bounds__(256) weight_prep_kernel(const WeightPrepTile* __restrict__ tiles) {
  extern __shared__ float wp_tile[];                 // [32][32 * RS + 1]
  const WeightPrepTile t = tiles[blockIdx.x];
  const int RS = t.RS, row = 32 * RS, pitch = row + 1;
  for (int j = threadIdx.x >> 5; j < 32; j += 8) {
    const float* s = t.src + ((int64_t)(t.co0 + j) * t.Cin + t.ci0) * RS;
    for (int i = threadIdx.x & 31; i < row; i += 32) wp_tile[j * pitch + i] = __ldg(s + i);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  // krsc[(co*RS + rs)*Cin + ci0 + lane] = w[co][ci0 + lane][rs]
  for (int q = threadIdx.x >> 5; q < 32 * RS; q += 8) {
    const int j = q / RS, rs = q - j * RS;
    t.krsc[((int64_t)(t.co0 + j) * RS + rs) * t.Cin + t.ci0 + lane] = __float2bfloat16_rn(wp_tile[j * pitch + lane * RS + rs]);
  }
  // rot[(ci*RS + RS-1-rs)*Cout + co0 + lane] = w[co0 + lane][ci][rs]
  for (int q = threadIdx.x >> 5; q < 32 * RS; q += 8) {
    const int c = q / RS, rs = q - c * RS;
    t.rot[((int64_t)(t.ci0 + c) * RS + (RS - 1 - rs)) * t.Cout + t.co0 + lane] = __float2bfloat16_rn(wp_tile[lane * pitch + c * RS + rs]);
  }
}

}  // namespace eeseg

using namespace eeseg;

static inline int stream_blocks(int64_t items) {
  const int64_t b = (items + 255) / 256;
  return (int)(b < kNumSMs * 8 ? (b > 0 ? b : 1) : kNumSMs * 8);
}

extern "C" int eeseg_dropout_fwd(const void* x, int64_t n, float p, void* rng_state, void* y, void* mask, void* stream_) {
  EESEG_REQUIRE(x && y && mask && rng_state, "dropout_fwd: null pointer");
  EESEG_REQUIRE(n % 8 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0, "dropout_fwd: n %% 8 == 0 and 16-byte alignment required");
  EESEG_REQUIRE(p >= 0.f && p < 1.f, "dropout_fwd: p = %f", (double)p);
  if (n == 0) return EESEG_OK;
  cudaStream_t st = (cudaStream_t)stream_;
  const uint32_t keep_thr = (uint32_t)((1.0 - (double)p) * 256.0 + 0.5);
  EESEG_REQUIRE(keep_thr >= 1 && keep_thr <= 256, "dropout_fwd: p = %f is not representable", (double)p);
  const float scale = 256.f / (float)keep_thr;      // 1 / realised keep probability
  dropout_fwd_kernel<<<stream_blocks(n / 8), 256, 0, st>>>((const uint4*)x, n / 8, keep_thr, scale, (const uint64_t*)rng_state,
                                                          (uint4*)y, (uint8_t*)mask);
  int rc = check_launch("dropout_fwd_kernel");
  if (rc) return rc;
  dropout_advance_kernel<<<1, 1, 0, st>>>((uint64_t*)rng_state, (uint64_t)(n / 8));
  return check_launch("dropout_advance_kernel");
}

extern "C" int eeseg_dropout_bwd(const void* dy, const void* mask, int64_t n, float p, void* dx, void* stream_) {
  EESEG_REQUIRE(dy && dx && mask, "dropout_bwd: null pointer");
  EESEG_REQUIRE(n % 8 == 0 && ((uintptr_t)dy & 15) == 0 && ((uintptr_t)dx & 15) == 0, "dropout_bwd: n %% 8 == 0 and 16-byte alignment required");
  if (n == 0) return EESEG_OK;
  const uint32_t keep_thr = (uint32_t)((1.0 - (double)p) * 256.0 + 0.5);
  EESEG_REQUIRE(keep_thr >= 1 && keep_thr <= 256, "dropout_bwd: p = %f", (double)p);
  dropout_bwd_kernel<<<stream_blocks(n / 8), 256, 0, (cudaStream_t)stream_>>>((const uint4*)dy, (const uint8_t*)mask, n / 8,
                                                                             256.f / (float)keep_thr, (uint4*)dx);
  return check_launch("dropout_bwd_kernel");
}

extern "C" size_t eeseg_sgd_chunk_bytes(void) { return sizeof(SgdChunk); }

extern "C" int eeseg_sgd_multi(const void* chunks, int n_chunks, const float* lrs, float momentum, float weight_decay,
                               void* stream_) {
  EESEG_REQUIRE(chunks && lrs, "sgd_multi: null pointer");
  if (n_chunks <= 0) return EESEG_OK;
  sgd_multi_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream_>>>((const SgdChunk*)chunks, lrs, momentum, weight_decay);
  return check_launch("sgd_multi_kernel");
}

extern "C" int eeseg_dense_bwd(const float* dy, const float* x, const float* W, int N, int K, int O, float* dW, float* dx,
                               void* stream_) {
  EESEG_REQUIRE(dy && x && W, "dense_bwd: null pointer");
  if (N <= 0 || K <= 0 || O <= 0) return EESEG_OK;
  cudaStream_t st = (cudaStream_t)stream_;
  int rc = EESEG_OK;
  if (dW) {
    dense_wgrad_kernel<<<(unsigned)(((int64_t)O * K + 255) / 256), 256, 0, st>>>(dy, x, N, K, O, dW);
    rc = check_launch("dense_wgrad_kernel");
    if (rc) return rc;
  }
  if (dx) {
    dense_dgrad_kernel<<<(unsigned)(((int64_t)N * K + 255) / 256), 256, 0, st>>>(dy, W, N, K, O, dx);
    rc = check_launch("dense_dgrad_kernel");
  }
  return rc;
}

extern "C" int eeseg_broadcast_rows_nhwc(const float* v, int N, int64_t hw, int C, float scale, void* out, void* stream_) {
  EESEG_REQUIRE(v && out, "broadcast_rows: null pointer");
  EESEG_REQUIRE(C % 8 == 0 && ((uintptr_t)out & 15) == 0, "broadcast_rows: C %% 8 == 0 and a 16-byte aligned output required");
  if (N <= 0 || hw <= 0) return EESEG_OK;
  const int64_t total8 = (int64_t)N * hw * (C / 8);
  broadcast_rows_kernel<<<stream_blocks(total8), 256, 0, (cudaStream_t)stream_>>>(v, hw, C, scale, (__nv_bfloat16*)out, total8);
  return check_launch("broadcast_rows_kernel");
}

extern "C" int eeseg_bn_rows_fwd(const float* x, int N, int C, const float* gamma, const float* beta, float* running_mean,
                                 float* running_var, float momentum, float eps, int relu, float* y, float* save_mean,
                                 float* save_invstd, void* stream_) {
  EESEG_REQUIRE(x && gamma && beta && y && save_mean && save_invstd, "bn_rows_fwd: null pointer");
  if (N <= 0 || C <= 0) return EESEG_OK;
  bn_rows_fwd_kernel<<<(C + 255) / 256, 256, 0, (cudaStream_t)stream_>>>(x, N, C, gamma, beta, running_mean, running_var,
                                                                          momentum, eps, relu, y, save_mean, save_invstd);
  return check_launch("bn_rows_fwd_kernel");
}

extern "C" int eeseg_bn_rows_bwd(const float* dy, const float* x, const float* y, int N, int C, const float* gamma,
                                 const float* save_mean, const float* save_invstd, int relu, float* dx, float* dgamma,
                                 float* dbeta, void* stream_) {
  EESEG_REQUIRE(dy && x && y && gamma && save_mean && save_invstd && dx && dgamma && dbeta, "bn_rows_bwd: null pointer");
  if (N <= 0 || C <= 0) return EESEG_OK;
  bn_rows_bwd_kernel<<<(C + 255) / 256, 256, 0, (cudaStream_t)stream_>>>(dy, x, y, N, C, gamma, save_mean, save_invstd, relu, dx,
                                                                          dgamma, dbeta);
  return check_launch("bn_rows_bwd_kernel");
}

extern "C" size_t eeseg_weight_prep_tile_bytes(void) { return sizeof(WeightPrepTile); }

extern "C" int eeseg_weight_prep_multi(const void* tiles, int n_tiles, int max_rs, void* stream_) {
  EESEG_REQUIRE(tiles, "weight_prep_multi: null pointer");
  EESEG_REQUIRE(max_rs >= 1 && max_rs <= 32, "weight_prep_multi: at most 32 taps");
  if (n_tiles <= 0) return EESEG_OK;
  const size_t smem = (size_t)32 * (32 * max_rs + 1) * sizeof(float);
  EESEG_CUDA(ensure_max_smem(weight_prep_kernel, 160 * 1024));
  weight_prep_kernel<<<n_tiles, 256, smem, (cudaStream_t)stream_>>>((const WeightPrepTile*)tiles);
  return check_launch("weight_prep_kernel");
}
