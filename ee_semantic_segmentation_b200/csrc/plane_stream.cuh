// Bulk-TMA staging of NCHW class planes for the streaming (HBM-bound) kernels.
//
// A tile = the same TILE consecutive pixels of all C planes of one image ([C][HW] planes, HW often
// odd — 513*513 — so plane starts are not 16 B aligned and per-lane vector loads are impossible).
// One elected thread issues one `cp.async.bulk` (UBLKCP) per plane: the 16 B-aligned superset of
// the wanted bytes is copied into a shared-memory row and completes on an mbarrier; consumers read
// their pixel at row[shift_c + i] where shift_c is the plane's misalignment in elements. With 3-4
// stages in flight per CTA this keeps > 100 KB of reads outstanding per SM without holding them
// in registers, which is what a 6.5 TB/s stream needs (Little's law at ~1-2 us loaded latency).
#pragma once
#include "common.cuh"

namespace eeseg {
namespace ps {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// Bytes of one staged row (TILE elements of ESIZE bytes + up to 15 B of leading and trailing slack).
__host__ __device__ constexpr int row_bytes(int tile, int esize) { return ((tile * esize + 15 + 15) / 16) * 16 + 16; }

// Issued by ONE thread: copies rows [p0, p0+count) of `rows` planes (plane r starts at
// base + r*plane_stride elements) into smem rows of `rbytes` bytes each, completing on `bar`.
// `limit` = one-past-the-end byte address that may be read (end of the tensor, rounded up to 16).
template <int ESIZE>
__device__ __forceinline__ void issue_tile(uint8_t* smem, int rbytes, uint64_t* bar, const uint8_t* base,
                                           int64_t plane_stride, int rows, int64_t p0, int count,
                                           const uint8_t* limit) {
  uint32_t total = 0;
  // first pass: byte counts (expect_tx must be armed before or with the copies; arm first)
  for (int r = 0; r < rows; ++r) {
    const uintptr_t a = (uintptr_t)(base + ((int64_t)r * plane_stride + p0) * ESIZE);
    const uintptr_t a0 = a & ~(uintptr_t)15;
    uintptr_t a1 = (a + (uintptr_t)count * ESIZE + 15) & ~(uintptr_t)15;
    if (a1 > (uintptr_t)limit) a1 = (uintptr_t)limit;
    total += (uint32_t)(a1 - a0);
  }
  mbar_expect_tx(bar, total);
  for (int r = 0; r < rows; ++r) {
    const uintptr_t a = (uintptr_t)(base + ((int64_t)r * plane_stride + p0) * ESIZE);
    const uintptr_t a0 = a & ~(uintptr_t)15;
    uintptr_t a1 = (a + (uintptr_t)count * ESIZE + 15) & ~(uintptr_t)15;
    if (a1 > (uintptr_t)limit) a1 = (uintptr_t)limit;
    bulk_g2s(smem + (size_t)r * rbytes, reinterpret_cast<const void*>(a0), (uint32_t)(a1 - a0), bar);
  }
}

__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Issued by ONE thread: writes `count` elements of one staged row (element i at row + (dst & 15) +
// i*ESIZE, i.e. the same layout issue_tile produces) to global memory at dst: the 16 B-aligned
// interior with one bulk store (full lines to L2), the < 16 B head and tail with scalar stores.
template <typename T>
__device__ __forceinline__ void store_row(const uint8_t* row, T* dst, int count) {
  constexpr int ES = (int)sizeof(T);
  const uintptr_t a = (uintptr_t)dst, b = a + (uintptr_t)count * ES;
  uintptr_t a1 = (a + 15) & ~(uintptr_t)15, b1 = b & ~(uintptr_t)15;
  const uint8_t* src = row + (a & 15);          // element 0
  if (b1 <= a1) { a1 = b; b1 = b; }             // shorter than one aligned chunk: all scalar
  for (uintptr_t q = a; q < a1; q += ES) *reinterpret_cast<T*>(q) = *reinterpret_cast<const T*>(src + (q - a));
  for (uintptr_t q = b1; q < b; q += ES) *reinterpret_cast<T*>(q) = *reinterpret_cast<const T*>(src + (q - a));
  if (b1 > a1) bulk_s2g(reinterpret_cast<void*>(a1), src + (a1 - a), (uint32_t)(b1 - a1));
}

// Element offset of pixel p0 inside the staged row of plane r.
template <int ESIZE>
__device__ __forceinline__ int row_shift(const uint8_t* base, int64_t plane_stride, int r, int64_t p0) {
  const uintptr_t a = (uintptr_t)(base + ((int64_t)r * plane_stride + p0) * ESIZE);
  return (int)((a & 15) / ESIZE);
}

}  // namespace ps
}  // namespace eeseg
