// Micro-benchmark (tuning tool, not part of the product): access-pattern variants for the HBM-bound
// plane-streaming kernels (multi-exit CE, confusion histogram) on [E][N][C][HW] planes with odd HW.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/ub_stream tools/ub_stream.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ld_stream(const float* p) {
  float v; asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v;
}

// ---- CE direct: thread = PIX pixels (strided by the block), loop over exits ----------------------
template <int C, int PIX, int THREADS>
__global__ void __launch_bounds__(THREADS) ce_direct(const float* __restrict__ logits, int64_t exit_stride,
                                                     const int64_t* __restrict__ targets, int E, int N, int64_t HW,
                                                     int64_t ignore, float gscale, float* __restrict__ dlogits,
                                                     float* __restrict__ part) {
  const int n = blockIdx.y;
  const int64_t p0 = (int64_t)blockIdx.x * (THREADS * PIX) + threadIdx.x;
  int64_t tt[PIX];
#pragma unroll
  for (int j = 0; j < PIX; ++j) {
    const int64_t p = p0 + j * THREADS;
    tt[j] = p < HW ? __ldg(targets + (int64_t)n * HW + p) : ignore;
  }
  constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
  for (int e = 0; e < E; ++e) {
    const float* base = logits + (int64_t)e * exit_stride + (int64_t)n * C * HW;
    float* gb = dlogits + (int64_t)e * exit_stride + (int64_t)n * C * HW;
    float v[PIX][C];
#pragma unroll
    for (int j = 0; j < PIX; ++j) {
      const int64_t p = p0 + j * THREADS;
#pragma unroll
      for (int c = 0; c < C; ++c) v[j][c] = p < HW ? ld_stream(base + (int64_t)c * HW + p) : 0.f;
    }
    float loss = 0.f;
#pragma unroll
    for (int j = 0; j < PIX; ++j) {
      const int64_t p = p0 + j * THREADS;
      if (p >= HW) continue;
      const bool ok = tt[j] != ignore && tt[j] >= 0 && tt[j] < C;
      const int t = ok ? (int)tt[j] : -1;
      float m = v[j][0];
#pragma unroll
      for (int c = 1; c < C; ++c) m = fmaxf(m, v[j][c]);
      const float m2 = m * kLog2e;
      float S = 0.f, vt = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        vt = c == t ? v[j][c] : vt;
        v[j][c] = ex2(fmaf(v[j][c], kLog2e, -m2));
        S += v[j][c];
      }
      if (ok) loss += lg2(S) * kLn2 - (vt - m);
      const float inv = ok ? gscale / S : 0.f;
      const float gs = ok ? gscale : 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) __stcs(gb + (int64_t)c * HW + p, fmaf(v[j][c], inv, c == t ? -gs : 0.f));
    }
    // block partial (not reduced here: the benchmark only needs the traffic); keep the value alive
    if (loss == 123.456f) part[0] = loss;
  }
}

// ---- CE direct, all exits of a pixel in flight at once (E compile-time) --------------------------
template <int C, int E, int THREADS>
__global__ void __launch_bounds__(THREADS) ce_direct_allE(const float* __restrict__ logits, int64_t exit_stride,
                                                          const int64_t* __restrict__ targets, int N, int64_t HW,
                                                          int64_t ignore, float gscale, float* __restrict__ dlogits,
                                                          float* __restrict__ part) {
  const int n = blockIdx.y;
  const int64_t p = (int64_t)blockIdx.x * THREADS + threadIdx.x;
  if (p >= HW) return;
  const int64_t tt = __ldg(targets + (int64_t)n * HW + p);
  float v[E][C];
#pragma unroll
  for (int e = 0; e < E; ++e)
#pragma unroll
    for (int c = 0; c < C; ++c) v[e][c] = ld_stream(logits + (int64_t)e * exit_stride + ((int64_t)n * C + c) * HW + p);
  constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
  const bool ok = tt != ignore && tt >= 0 && tt < C;
  const int t = ok ? (int)tt : -1;
  float loss = 0.f;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    float* gb = dlogits + (int64_t)e * exit_stride + (int64_t)n * C * HW;
    float m = v[e][0];
#pragma unroll
    for (int c = 1; c < C; ++c) m = fmaxf(m, v[e][c]);
    const float m2 = m * kLog2e;
    float S = 0.f, vt = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      vt = c == t ? v[e][c] : vt;
      v[e][c] = ex2(fmaf(v[e][c], kLog2e, -m2));
      S += v[e][c];
    }
    if (ok) loss += lg2(S) * kLn2 - (vt - m);
    const float inv = ok ? gscale / S : 0.f;
    const float gs = ok ? gscale : 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) __stcs(gb + (int64_t)c * HW + p, fmaf(v[e][c], inv, c == t ? -gs : 0.f));
  }
  if (loss == 123.456f) part[0] = loss;
}

// ---- plain copy with the same access pattern (ceiling for scalar plane accesses) -----------------
template <int C, int THREADS>
__global__ void __launch_bounds__(THREADS) plane_copy(const float* __restrict__ in, float* __restrict__ out, int64_t HW) {
  const int64_t plane0 = (int64_t)blockIdx.y * C * HW;
  const int64_t p = (int64_t)blockIdx.x * THREADS + threadIdx.x;
  if (p >= HW) return;
  float v[C];
#pragma unroll
  for (int c = 0; c < C; ++c) v[c] = ld_stream(in + plane0 + (int64_t)c * HW + p);
#pragma unroll
  for (int c = 0; c < C; ++c) __stcs(out + plane0 + (int64_t)c * HW + p, v[c] * 1.0001f);
}

// flat float4 copy (the MEASURED_PEAKS-style ceiling)
__global__ void flat_copy(const float4* __restrict__ in, float4* __restrict__ out, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) out[i] = in[i];
}

// ---- histogram direct ----------------------------------------------------------------------------
__device__ __forceinline__ void hist_add(unsigned* h, int key) {
  unsigned peers = __match_any_sync(0xffffffffu, key);
  if (key >= 0) {
    int leader = __ffs(peers) - 1;
    if ((int)(threadIdx.x & 31) == leader) atomicAdd(h + key, (unsigned)__popc(peers));
  }
}
template <int C, int PIX, int THREADS>
__global__ void __launch_bounds__(THREADS) hist_direct(const float* __restrict__ logits, const int64_t* __restrict__ targets,
                                                       int64_t HW, unsigned long long* __restrict__ cm) {
  constexpr int bins = (C + 1) * C, NW = THREADS / 32;
  __shared__ unsigned hist[NW * bins > 12000 ? 12000 / bins * bins : NW * bins];
  constexpr int copies = (NW * bins > 12000 ? 12000 / bins : NW);
  for (int i = threadIdx.x; i < copies * bins; i += THREADS) hist[i] = 0;
  __syncthreads();
  unsigned* h = hist + ((threadIdx.x >> 5) % copies) * bins;
  const int n = blockIdx.y;
  const float* base = logits + (int64_t)n * C * HW;
  const int64_t p0 = (int64_t)blockIdx.x * (THREADS * PIX) + threadIdx.x;
  float v[PIX][C];
  int64_t t[PIX];
#pragma unroll
  for (int j = 0; j < PIX; ++j) {
    const int64_t p = p0 + j * THREADS;
    t[j] = p < HW ? __ldg(targets + (int64_t)n * HW + p) : -1;
#pragma unroll
    for (int c = 0; c < C; ++c) v[j][c] = p < HW ? ld_stream(base + (int64_t)c * HW + p) : 0.f;
  }
#pragma unroll
  for (int j = 0; j < PIX; ++j) {
    const int64_t p = p0 + j * THREADS;
    float best = v[j][0];
    int arg = 0;
#pragma unroll
    for (int c = 1; c < C; ++c) {
      const bool take = (v[j][c] > best) || (v[j][c] != v[j][c] && best == best);
      best = take ? v[j][c] : best;
      arg = take ? c : arg;
    }
    const int tt = (t[j] >= 0 && t[j] < C) ? (int)t[j] : C;
    hist_add(h, p < HW ? tt * C + arg : -1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < bins; i += THREADS) {
    unsigned long long s = 0;
    for (int k = 0; k < copies; ++k) s += hist[k * bins + i];
    if (s) atomicAdd(cm + (int64_t)n * bins + i, s);
  }
}

template <typename F>
static float time_it(F f, int iters = 20) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; ++i) f();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < iters; ++i) f();
  CK(cudaEventRecord(b));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, a, b));
  CK(cudaGetLastError());
  return ms / iters;
}

extern "C" int eeseg_confusion_hist(const void* pred, int pred_kind, int dtype, const int64_t* targets, int N, int C, int64_t HW,
                                    int64_t* cm, int accumulate, void* stream);

int main() {
  constexpr int E = 3, N = 4, C = 21;
  const int64_t HW = 513 * 513;
  const int64_t elems = (int64_t)E * N * C * HW;
  float *x, *g; int64_t* tg; float* part; unsigned long long* cm;
  CK(cudaMalloc(&x, elems * 4)); CK(cudaMalloc(&g, elems * 4)); CK(cudaMalloc(&tg, N * HW * 8)); CK(cudaMalloc(&part, 1 << 20));
  CK(cudaMalloc(&cm, N * (C + 1) * C * 8));
  CK(cudaMemset(cm, 0, N * (C + 1) * C * 8));
  {  // deterministic junk
    float* hx = (float*)malloc(elems * 4);
    uint32_t s = 12345u;
    for (int64_t i = 0; i < elems; ++i) { s = s * 1664525u + 1013904223u; hx[i] = ((s >> 8) & 0xffff) / 65536.f * 6.f - 3.f; }
    CK(cudaMemcpy(x, hx, elems * 4, cudaMemcpyHostToDevice));
    int64_t* ht = (int64_t*)malloc(N * HW * 8);
    for (int64_t i = 0; i < N * HW; ++i) ht[i] = ((i / 16) % 513 / 16 + (i / 16 / 513)) % 22;
    CK(cudaMemcpy(tg, ht, N * HW * 8, cudaMemcpyHostToDevice));
    free(hx); free(ht);
  }
  const double ce_bytes = 2.0 * elems * 4 + 2.0 * N * HW * 8;
  const double hist_bytes = (double)N * HW * (C * 4 + 8);
  auto rep = [&](const char* name, float ms, double bytes) {
    printf("%-44s %8.1f us  %7.1f GB/s  %5.1f %% of 6555\n", name, ms * 1e3, bytes / ms / 1e6, bytes / ms / 1e6 / 6555.2 * 100);
  };
  {
    const int64_t n4 = elems / 4;
    rep("flat float4 copy (same bytes as CE)", time_it([&] { flat_copy<<<148 * 16, 512>>>((const float4*)x, (float4*)g, n4); }), 2.0 * elems * 4);
  }
#define PCOPY(T) rep("plane_copy threads=" #T, time_it([&] { plane_copy<C, T><<<dim3((HW + T - 1) / T, E * N), T>>>(x, g, HW); }), 2.0 * elems * 4)
  PCOPY(128); PCOPY(256); PCOPY(512);
#define CED(PIX, T) rep("ce_direct PIX=" #PIX " threads=" #T, time_it([&] { ce_direct<C, PIX, T><<<dim3((HW + T * PIX - 1) / (T * PIX), N), T>>>(x, (int64_t)N * C * HW, tg, E, N, HW, 21, 1e-6f, g, part); }), ce_bytes)
  CED(1, 128); CED(1, 256); CED(1, 512); CED(2, 128); CED(2, 256); CED(4, 128);
#define CEA(T) rep("ce_direct_allE threads=" #T, time_it([&] { ce_direct_allE<C, E, T><<<dim3((HW + T - 1) / T, N), T>>>(x, (int64_t)N * C * HW, tg, N, HW, 21, 1e-6f, g, part); }), ce_bytes)
  CEA(128); CEA(256);
#define HD(PIX, T) rep("hist_direct PIX=" #PIX " threads=" #T, time_it([&] { hist_direct<C, PIX, T><<<dim3((HW + T * PIX - 1) / (T * PIX), N), T>>>(x, tg, HW, cm); }), hist_bytes)
  // histogram input = exit 0 only (88 MB < L2 126 MB: alternate exits to defeat the cache)
  {
    int k = 0;
    auto hrun = [&](auto kern, dim3 grid, int T) { kern<<<grid, T>>>(x + (int64_t)(k++ % E) * N * C * HW, tg, HW, cm); };
#define HDR(PIX, T) rep("hist_direct PIX=" #PIX " threads=" #T, time_it([&] { hrun(hist_direct<C, PIX, T>, dim3((HW + T * PIX - 1) / (T * PIX), N), T); }), hist_bytes)
    HDR(1, 128); HDR(1, 256); HDR(2, 128); HDR(2, 256); HDR(4, 128); HDR(4, 256);
    // the product kernel through the C ABI, same buffers and timing loop
    int kk = 0;
    rep("product eeseg_confusion_hist (accumulate)", time_it([&] {
          eeseg_confusion_hist(x + (int64_t)(kk++ % E) * N * C * HW, 0, 0, tg, N, C, HW, (int64_t*)cm, 1, nullptr); }), hist_bytes);
    rep("product eeseg_confusion_hist (memset + kernel)", time_it([&] {
          eeseg_confusion_hist(x + (int64_t)(kk++ % E) * N * C * HW, 0, 0, tg, N, C, HW, (int64_t*)cm, 0, nullptr); }), hist_bytes);
  }
  return 0;
}
