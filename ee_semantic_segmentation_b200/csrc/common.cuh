// Shared helpers for libeeseg_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <mutex>

#include "../../include/eeseg.h"

namespace eeseg {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs (grids are sized in multiples of this)

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

inline int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return EESEG_ERR_CUDA;
  }
  return EESEG_OK;
}

#define EESEG_REQUIRE(cond, ...)      \
  do {                                \
    if (!(cond)) {                    \
      ::eeseg::set_error(__VA_ARGS__); \
      return EESEG_ERR_ARG;           \
    }                                 \
  } while (0)

#define EESEG_CUDA(call)                                                       \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess) {                                                  \
      ::eeseg::set_error("%s: %s", #call, cudaGetErrorString(e__));            \
      return EESEG_ERR_CUDA;                                                   \
    }                                                                          \
  } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device), race-free (the attribute is per device;
// callers may launch from several host threads). Keyed by the kernel's ADDRESS: two instantiations of one template share
// their function type.
template <typename K>
inline cudaError_t ensure_max_smem(K kernel, int bytes) {
  static std::mutex mu;
  static const void* done_fn[64];
  static int done_dev[64];
  static int n_done = 0;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const void* fn = reinterpret_cast<const void*>(kernel);
  std::lock_guard<std::mutex> lock(mu);
  for (int i = 0; i < n_done; ++i)
    if (done_fn[i] == fn && done_dev[i] == dev) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && n_done < 64) {
    done_fn[n_done] = fn;
    done_dev[n_done] = dev;
    ++n_done;
  }
  return e;
}

// ---- typed scalar access ------------------------------------------------------------------------
__device__ __forceinline__ float ldf(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ldf(const __nv_bfloat16* p) {
  return __bfloat162float(__ldg(p));
}
// streaming (read-once) variants: bypass L1 allocation
__device__ __forceinline__ float ldf_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float ldf_stream(const __nv_bfloat16* p) {
  unsigned short u;
  asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(u) : "l"(p));
  return __uint_as_float(((unsigned)u) << 16);
}
__device__ __forceinline__ void stf(float* p, float v) { __stcs(p, v); }
__device__ __forceinline__ void stf(__nv_bfloat16* p, float v) {
  __nv_bfloat16 b = __float2bfloat16_rn(v);
  __stcs(reinterpret_cast<unsigned short*>(p), *reinterpret_cast<unsigned short*>(&b));
}

// Plane addressing: element (c, p) of an image lives at base + c*HW + p. With the plane stride as a
// 32-bit BYTE count and c a compile-time constant the address is ONE widening multiply-add
// (IMAD.WIDE.U32) per access; 64-bit HW arithmetic costs 3-4 instructions per access and made v3.0 of
// this kernel half issue-bound (ncu: 578 warp instructions per pixel-exit, issue slots 50 % busy).
template <typename T>
__device__ __forceinline__ const T* plane_ptr(const T* base, uint32_t c, uint32_t plane_bytes) {
  return reinterpret_cast<const T*>(reinterpret_cast<const char*>(base) + (uint64_t)c * plane_bytes);
}
template <typename T>
__device__ __forceinline__ T* plane_ptr(T* base, uint32_t c, uint32_t plane_bytes) {
  return reinterpret_cast<T*>(reinterpret_cast<char*>(base) + (uint64_t)c * plane_bytes);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace eeseg
