// Weighted multi-exit pixelwise cross-entropy, forward (+ fused gradient) and backward.
// Reference contract: BrXEntropyLoss (my_pixelwise_xentropy.py:19-46) over
// torch.nn.CrossEntropyLoss(reduction='mean', ignore_index) per exit — see include/eeseg.h.
//
// HBM-bound: one pass over logits [E][N][C][HW] (read) and, when the gradient is wanted, one write
// of the same size — the reference does E separate log_softmax + nll passes forward and E backward.
// One thread per pixel (lanes = consecutive pixels of a class plane -> coalesced requests), the C
// class values live in registers, per-block loss partials are written to a scratch array and
// reduced in a fixed order (bit-reproducible loss).
#include "common.cuh"

namespace eeseg {

constexpr int kCeThreads = 256;
constexpr int kCePix = 2;

static inline int ce_grid_x(int E, int N, int64_t HW) {
  int64_t want = (HW + kCeThreads * kCePix - 1) / (kCeThreads * kCePix);
  int64_t cap = (kNumSMs * 8 + (int64_t)E * N - 1) / ((int64_t)E * N);
  if (cap < 1) cap = 1;
  return (int)(want < cap ? want : cap);
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

template <typename T, int CMAX>
__global__ void __launch_bounds__(kCeThreads) ce_kernel(
    const T* __restrict__ logits, int64_t exit_stride, const int64_t* __restrict__ targets, int N,
    int C, int64_t HW, int64_t ignore, const float* __restrict__ coef,
    const int64_t* __restrict__ valid_count, T* __restrict__ dlogits, double* __restrict__ part) {
  const int e = blockIdx.y / N, n = blockIdx.y % N;
  const T* base = logits + (int64_t)e * exit_stride + (int64_t)n * C * HW;
  T* gbase = dlogits ? dlogits + (int64_t)e * exit_stride + (int64_t)n * C * HW : nullptr;
  const int64_t* tg = targets + (int64_t)n * HW;
  float gscale = 0.f;
  if (gbase) {
    const float valid = (float)(*valid_count);
    gscale = (coef ? coef[e] : 1.f) / valid;  // valid == 0 -> inf/NaN like torch's 0/0
  }
  float loss_acc = 0.f;
  const int64_t step = (int64_t)gridDim.x * kCeThreads * kCePix;
  for (int64_t p0 = (int64_t)blockIdx.x * kCeThreads * kCePix + threadIdx.x; p0 < HW; p0 += step) {
    float v[kCePix][CMAX];
    int t[kCePix];
    bool ok[kCePix];
#pragma unroll
    for (int j = 0; j < kCePix; ++j) {
      const int64_t p = p0 + (int64_t)j * kCeThreads;
      int64_t tt = p < HW ? __ldg(tg + p) : ignore;
      ok[j] = p < HW && tt != ignore && tt >= 0 && tt < C;
      t[j] = (int)tt;
    }
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
#pragma unroll
      for (int j = 0; j < kCePix; ++j)
        if (c < C && ok[j]) v[j][c] = ldf_stream(base + (int64_t)c * HW + p0 + (int64_t)j * kCeThreads);
#pragma unroll
    for (int j = 0; j < kCePix; ++j) {
      const int64_t p = p0 + (int64_t)j * kCeThreads;
      if (ok[j]) {
        float m = v[j][0];
#pragma unroll
        for (int c = 1; c < CMAX; ++c)
          if (c < C) m = fmaxf(m, v[j][c]);
        float S = 0.f, picked = 0.f;
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
          if (c < C) {
            float z = v[j][c] - m;
            v[j][c] = z;
            S += exp2f(z * 1.4426950408889634f);
            picked = (c == t[j]) ? z : picked;
          }
        const float lse = logf(S);
        loss_acc += lse - picked;
        if (gbase) {
#pragma unroll
          for (int c = 0; c < CMAX; ++c)
            if (c < C) {
              float sm = exp2f((v[j][c] - lse) * 1.4426950408889634f);
              stf(gbase + (int64_t)c * HW + p, gscale * (sm - (c == t[j] ? 1.f : 0.f)));
            }
        }
      } else if (gbase && p < HW) {
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
          if (c < C) stf(gbase + (int64_t)c * HW + p, 0.f);
      }
    }
  }
  if (part) {
    __shared__ double s[kCeThreads / 32];
    double ws = warp_sum((double)loss_acc);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = ws;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int i = 0; i < kCeThreads / 32; ++i) t += s[i];
      // slot order: [e][n][blockIdx.x]
      part[((int64_t)e * N + n) * gridDim.x + blockIdx.x] = t;
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

template <typename T>
static int launch_ce(const T* logits, int64_t exit_stride, const int64_t* targets, int E, int N,
                     int C, int64_t HW, int64_t ignore, const float* coef,
                     const int64_t* valid_count, T* dlogits, double* part, cudaStream_t stream) {
  dim3 grid(ce_grid_x(E, N, HW), E * N);
#define EESEG_CE_CASE(CM)                                                                    \
  if (C <= CM) {                                                                              \
    ce_kernel<T, CM><<<grid, kCeThreads, 0, stream>>>(logits, exit_stride, targets, N, C, HW, \
                                                      ignore, coef, valid_count, dlogits, part); \
    return check_launch("ce_kernel");                                                         \
  }
  EESEG_CE_CASE(24)
  EESEG_CE_CASE(32)
  EESEG_CE_CASE(64)
#undef EESEG_CE_CASE
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
