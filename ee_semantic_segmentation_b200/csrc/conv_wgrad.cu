// Weight gradient of the exit-head convolutions (training: DeepLabHead / ASPP convs called at
// from_deepv3_new.py:147,151 under train_funcs.py:22-27) as an implicit GEMM on tcgen05:
//
//   dW[co][r][s][ci] = sum over pixels (n,y,x) of dY[n,y,x,co] * X[n, y + r*dil - pad, x + s*dil - pad, ci]
//
// GEMM view per tap (r,s): D[M = co][N = ci] += A[M][K] * B[K][N] with K = output pixels. In NHWC both
// operands have the pixel index as their SLOW dimension, i.e. they are "MN-major" for the tensor core:
// * A tile: 128 co x 128 pixels = two 4-D TMA boxes {64 co, BW, BH, 1 image} of dY (SW128); the box lands as
//   [pixel][64 channels = 128 B], which is exactly the canonical MN-major SWIZZLE_128B layout
//   ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16 B units: k rows 128 B apart, 8-row groups SBO = 1024 B apart,
//   64-channel blocks LBO = 16 KB apart.
// * B tile: BN ci x 128 pixels = BN/64 boxes of X shifted by the tap; TMA's out-of-bounds zero fill is the
//   padding, as in the forward kernel. Pixel rectangles are BW x BH <= 128 (65x65 maps: 11x11 = 121): smem
//   rows 121..127 of every box are zeroed once and never written again (0 x 0 adds nothing to the K sum).
// * One CTA = one (tap, 128-co block, BN-ci block, K split): it walks its pixel tiles through a TMA ring
//   (warp 0), issues 8 UMMA 128xBNx16 per tile (warp 1, fp32 accumulator in TMEM), and warps 2-5 write the
//   accumulator to dW in fp32 — or, when the K dimension is split across CTAs to fill the 148 SMs, to its own
//   partial buffer, summed afterwards in a fixed order (fp32 atomics cost 36 of 55 us on a 1x1 Cin=2048 layer and
//   made dW irreproducible).
// Pixel tiles for which the shifted window lies entirely in the padding are skipped.
#include <cuda.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace eeseg {

constexpr int kWgThreads = 192;   // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
constexpr int kWgBox = 128 * 128; // bytes of one {64 ch, <=128 pixels} box in shared memory
constexpr int kWgMaxStages = 4;

struct WgradParams {
  int N, h, w, Cin, Cout;
  int R, S, dil, pad;
  int BW, BH, tiles_x, tiles_y;
  int BN, nblocks, mblocks, ksplit;
  int stages;
  int co_off;       // first channel of this conv inside the dY tensor map
  float* dw;        // ksplit == 1: dW itself; else the partial buffer [ksplit][Cout][R*S][Cin]
  int64_t part_stride;   // elements between the partial buffers of consecutive K splits
};

__host__ __device__ inline uint32_t tmem_cols_for_wg(int bn) { return bn <= 32 ? 32u : bn <= 64 ? 64u : bn <= 128 ? 128u : 256u; }

// MN-major, SWIZZLE_128B: LBO = distance between 64-element blocks along M/N, SBO = distance between
// 8-row groups along K (sm_100 descriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version 1
// [46,48), layout_type [61,64) = 2)
__device__ __forceinline__ uint64_t make_sw128_mn_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffff) >> 4);
  d |= (uint64_t)(kWgBox >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ bool wg_tile_live(const WgradParams& p, int dy, int dx, int y0, int x0) {
  const int y_hi = min(y0 + p.BH, p.h), x_hi = min(x0 + p.BW, p.w);
  return (y_hi - 1 + dy >= 0) && (y0 + dy < p.h) && (x_hi - 1 + dx >= 0) && (x0 + dx < p.w);
}

__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x,
                  const WgradParams p) {
  extern __shared__ __align__(1024) uint8_t wg_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(wg_smem_raw) + 1023) & ~(uintptr_t)1023);
  const int nb_boxes = p.BN >> 6;
  const uint32_t a_bytes = 2 * kWgBox;
  const uint32_t stage_bytes = a_bytes + (uint32_t)nb_boxes * kWgBox;
  uint8_t* tail = smem + (size_t)p.stages * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + kWgMaxStages;
  uint64_t* acc_bar = empty_bar + kWgMaxStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // work item
  int item = blockIdx.x;
  const int ks = item % p.ksplit; item /= p.ksplit;
  const int nb = item % p.nblocks; item /= p.nblocks;
  const int mb = item % p.mblocks;
  const int tap = item / p.mblocks;
  const int dy = (tap / p.S) * p.dil - (p.R > 1 ? p.pad : 0), dx = (tap % p.S) * p.dil - (p.S > 1 ? p.pad : 0);
  const int tiles_img = p.tiles_x * p.tiles_y;
  const int total_pt = p.N * tiles_img;
  const uint32_t ncols = tmem_cols_for_wg(p.BN);

  // zero the ring once: the rows past BW*BH of every box are never written by TMA
  for (uint32_t i = threadIdx.x; i < (uint32_t)p.stages * stage_bytes / 16; i += kWgThreads)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_dy);
    prefetch_tmap(&tmap_x);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
    }
    mbar_init(acc_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // number of live pixel tiles of this CTA (identical in every role): the lanes of each warp split the tiles (a serial
  // scan is up to 144 iterations with two divisions each — ~10 us in front of a 40-140 us kernel when ksplit == 1)
  int n_live = 0;
  for (int pt = ks + lane * p.ksplit; pt < total_pt; pt += 32 * p.ksplit) {
    const int trem = pt % tiles_img;
    n_live += wg_tile_live(p, dy, dx, (trem / p.tiles_x) * p.BH, (trem % p.tiles_x) * p.BW) ? 1 : 0;
  }
  n_live = __reduce_add_sync(0xffffffffu, n_live);

  if (warp == 0) {
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t box_tx = (uint32_t)(p.BW * p.BH * 128);
      // (image, tile row, tile column) of pixel tile pt, advanced by ksplit per iteration without divisions
      int n_img = ks / tiles_img, ty = (ks % tiles_img) / p.tiles_x, tx = (ks % tiles_img) % p.tiles_x;
      for (int pt = ks; pt < total_pt; pt += p.ksplit) {
        if (pt != ks) {
          tx += p.ksplit;
          while (tx >= p.tiles_x) { tx -= p.tiles_x; ++ty; }
          while (ty >= p.tiles_y) { ty -= p.tiles_y; ++n_img; }
        }
        const int y0 = ty * p.BH, x0 = tx * p.BW;
        if (!wg_tile_live(p, dy, dx, y0, x0)) continue;
        mbar_wait(empty_bar + s, ph ^ 1u);
        uint8_t* sa = smem + (size_t)s * stage_bytes;
        mbar_expect_tx(full_bar + s, box_tx * (uint32_t)(2 + nb_boxes));
        for (int j = 0; j < 2; ++j)
          tma_load_4d(sa + j * kWgBox, &tmap_dy, full_bar + s, p.co_off + mb * 128 + j * 64, x0, y0, n_img);
        for (int j = 0; j < nb_boxes; ++j)
          tma_load_4d(sa + a_bytes + j * kWgBox, &tmap_x, full_bar + s, nb * p.BN + j * 64, x0 + dx, y0 + dy, n_img);
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // instruction descriptor: D fp32, A/B bf16, both MN-major (bits 15, 16), N >> 3 at 17, M >> 4 at 24
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                           ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = 0; t < n_live; ++t) {
        mbar_wait(full_bar + s, ph);
        tcgen05_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
#pragma unroll
        for (int k = 0; k < 8; ++k) {   // 16 pixel rows per MMA = two 8-row groups = 2 KB
          const uint64_t adesc = make_sw128_mn_desc(sa + (uint32_t)k * 2048u);
          const uint64_t bdesc = make_sw128_mn_desc(sa + a_bytes + (uint32_t)k * 2048u);
          umma_bf16(tmem_base, adesc, bdesc, idesc, (t > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(empty_bar + s);
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
      if (n_live > 0) umma_commit(acc_bar);
    }
    __syncwarp();
  } else {
    const int quad = warp & 3;
    const int m = quad * 32 + lane;
    const int co = mb * 128 + m;
    float* row = p.dw + (int64_t)ks * p.part_stride + ((int64_t)co * (p.R * p.S) + tap) * p.Cin + (int64_t)nb * p.BN;
    const bool store = co < p.Cout;   // rows past Cout (64-channel layers) hold zeros and have no destination
    if (n_live > 0) {
      mbar_wait(acc_bar, 0);
      tcgen05_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
      // 16-column chunks over two register sets: the next tcgen05.ld is in flight while the current chunk is stored
      // (BN is 64, 128 or 256)
      auto put = [&](const uint32_t (&v)[16], int col) {
        if (!store) return;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          reinterpret_cast<float4*>(row + col)[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                                 __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
      };
      uint32_t va[16], vb[16];
      tmem_ld16(trow, va);
      tmem_ld_wait();
      for (int col = 0; col < p.BN; col += 32) {
        tmem_ld16(trow + (uint32_t)(col + 16), vb);
        put(va, col);
        tmem_ld_wait();
        if (col + 32 < p.BN) tmem_ld16(trow + (uint32_t)(col + 32), va);
        put(vb, col + 16);
        if (col + 32 < p.BN) tmem_ld_wait();
      }
    } else if (store) {
      for (int col = 0; col < p.BN; col += 4) *reinterpret_cast<float4*>(row + col) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
  }
}

// [Cout][R][S][Cin] bf16 -> [Cin][R][S][Cout] bf16 with the taps rotated by 180 degrees: the weights of the
// input-gradient convolution (dX = conv(dY, W')). Small (<= 9.4 MB), once per step.
__global__ void weight_rot180_t_kernel(const __nv_bfloat16* __restrict__ w, int Cout, int RS, int Cin,
                                       __nv_bfloat16* __restrict__ out) {
  __shared__ __nv_bfloat16 tile[32][33];
  const int tap = blockIdx.z;
  const int ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int co = co0 + j, ci = ci0 + threadIdx.x;
    if (co < Cout && ci < Cin) tile[j][threadIdx.x] = w[((int64_t)co * RS + tap) * Cin + ci];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int ci = ci0 + j, co = co0 + threadIdx.x;
    if (co < Cout && ci < Cin) out[((int64_t)ci * RS + (RS - 1 - tap)) * Cout + co] = tile[threadIdx.x][j];
  }
}

// dW = sum over K splits of the partial buffers, fixed order (deterministic)
__global__ void wgrad_reduce_kernel(const float4* __restrict__ part, int ksplit, int64_t n4, float4* __restrict__ dw,
                                    int accumulate) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 a = part[i];
    for (int k = 1; k < ksplit; ++k) {
      const float4 b = part[(int64_t)k * n4 + i];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    if (accumulate) {
      const float4 o = dw[i];
      a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
    }
    dw[i] = a;
  }
}

// The same sum written straight into the PARAMETER's gradient, which lives in the parameter's own layout
// [Cout][Cin][R][S] (nn.Conv2d.weight): grad[co][ci][rs] (+)= sum_k part[k][co][rs][ci]. For 1x1 kernels the two layouts
// coincide (wgrad_reduce_kernel with the accumulate flag); otherwise one block per (co, chunk of up to 256 ci): RS coalesced
// rows in, transposed through shared memory, chunk*RS contiguous floats out. Replaces, per layer, autograd's AccumulateGrad
// `grad.add_(dW.permute(...))` (a strided read-modify-write launch).
__global__ void __launch_bounds__(256) wgrad_reduce_to_param_kernel(const float* __restrict__ part, int ksplit, int64_t part_stride,
                                                                     int Cout, int RS, int Cin, int accumulate,
                                                                     float* __restrict__ grad) {
  extern __shared__ float tile[];            // [RS][257]
  const int co = blockIdx.y, ci0 = blockIdx.x * 256, t = threadIdx.x;
  const int nci = min(256, Cin - ci0);
  if (t < nci) {
    for (int rs = 0; rs < RS; ++rs) {
      const int64_t src = ((int64_t)co * RS + rs) * Cin + ci0 + t;
      float a = part[src];
      for (int k = 1; k < ksplit; ++k) a += part[(int64_t)k * part_stride + src];   // fixed order
      tile[rs * 257 + t] = a;
    }
  }
  __syncthreads();
  float* dst = grad + ((int64_t)co * Cin + ci0) * RS;
  for (int i = t; i < nci * RS; i += 256) {
    const int ci = i / RS, rs = i - ci * RS;
    const float v = tile[rs * 257 + ci];
    dst[i] = accumulate ? dst[i] + v : v;
  }
}

__global__ void fill_scale_shift_kernel(float* scale, float* shift, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { scale[i] = 1.f; shift[i] = 0.f; }
}

}  // namespace eeseg

using namespace eeseg;

static int wgrad_geometry(int N, int h, int w, int Cin, int Cout, int R, int S, WgradParams& p) {
  pick_tile(h, w, p.BW, p.BH);
  p.tiles_x = (w + p.BW - 1) / p.BW;
  p.tiles_y = (h + p.BH - 1) / p.BH;
  p.BN = Cin % 256 == 0 ? 256 : (Cin % 128 == 0 ? 128 : 64);
  p.nblocks = Cin / p.BN;
  p.mblocks = (Cout + 127) / 128;   // Cout = 64: the upper half of the 128-row tile reads channels past the end of dY
                                    // (TMA zero fill) and is not stored
  const int base_items = R * S * p.nblocks * p.mblocks;
  const int total_pt = N * p.tiles_x * p.tiles_y;
  int ksplit = kNumSMs / base_items;
  if (ksplit < 1) ksplit = 1;
  if (ksplit > total_pt) ksplit = total_pt;
  p.ksplit = ksplit;
  return base_items;
}

extern "C" size_t eeseg_conv_igemm_wgrad_workspace_bytes(int N, int h, int w, int Cin, int Cout, int R, int S) {
  if (N < 1 || h < 1 || w < 1 || Cin < 64 || Cout < 64 || R < 1 || S < 1) return 256;
  WgradParams p;
  wgrad_geometry(N, h, w, Cin, Cout, R, S, p);
  return (p.ksplit > 1 ? (size_t)p.ksplit * Cout * R * S * Cin * sizeof(float) : 0) + 256;
}

// to_param: 0 = dw is fp32 [Cout][R][S][Cin], overwritten; 1 / 2 = dw is the parameter's gradient [Cout][Cin][R][S],
// overwritten / accumulated into (always through the partial buffers: the workspace holds max(ksplit, 1) of them)
static int wgrad_launch(const void* x, const void* dy, int64_t ldy, int dy_channels, int co_off, int N, int h, int w, int Cin,
                        int Cout, int R, int S, int dilation, float* dw, void* workspace, int to_param, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  EESEG_REQUIRE(x && dy && dw, "conv_wgrad: null pointer");
  EESEG_REQUIRE(N >= 1 && h >= 1 && w >= 1, "conv_wgrad: bad sizes");
  EESEG_REQUIRE(Cin % 64 == 0, "conv_wgrad: Cin=%d must be a multiple of 64", Cin);
  EESEG_REQUIRE(Cout % 64 == 0, "conv_wgrad: Cout=%d must be a multiple of 64", Cout);
  // even kernel sizes (the space-to-depth'ed 4x1 stem) use the same rule pad = dilation * (R / 2), i.e. the forward's pad = 2
  EESEG_REQUIRE(R >= 1 && S >= 1 && R * S <= 32, "conv_wgrad: at most 32 taps");
  EESEG_REQUIRE(dilation >= 1, "conv_wgrad: dilation %d", dilation);
  EESEG_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)dy & 15) == 0 && ((uintptr_t)dw & 15) == 0 && (ldy % 8) == 0 &&
                ((uintptr_t)workspace & 15) == 0, "conv_wgrad: pointers and the dY pixel stride must be 16-byte aligned");
  EESEG_REQUIRE(co_off >= 0 && co_off + Cout <= dy_channels && dy_channels <= ldy, "conv_wgrad: channel window outside dY");
  EESEG_REQUIRE(R == S || R == 1 || S == 1, "conv_wgrad: square or 1-D kernels");
  EncodeTiledFn encode = get_encode();
  if (!encode) { set_error("conv_wgrad: cuTensorMapEncodeTiled unavailable"); return EESEG_ERR_CUDA; }
  WgradParams p;
  p.N = N; p.h = h; p.w = w; p.Cin = Cin; p.Cout = Cout; p.R = R; p.S = S; p.dil = dilation;
  p.pad = dilation * (R / 2);   // 'same'
  const int base_items = wgrad_geometry(N, h, w, Cin, Cout, R, S, p);
  const int64_t dw_elems = (int64_t)Cout * R * S * Cin;
  EESEG_REQUIRE((p.ksplit == 1 && !to_param) || workspace, "conv_wgrad: this shape splits the pixels over %d CTAs and needs the workspace", p.ksplit);
  p.co_off = co_off;
  p.dw = (p.ksplit > 1 || to_param) ? reinterpret_cast<float*>(workspace) : dw;
  p.part_stride = p.ksplit > 1 ? dw_elems : 0;
  const size_t stage_bytes = (size_t)(2 + p.BN / 64) * kWgBox;
  int stages = (int)((227 * 1024 - 1024 - 256) / stage_bytes);
  if (stages > kWgMaxStages) stages = kWgMaxStages;
  EESEG_REQUIRE(stages >= 2, "conv_wgrad: tile does not fit shared memory");
  p.stages = stages;
  const size_t smem_bytes = 1024 + stages * stage_bytes + 256;
  CUtensorMap tmdy, tmx;
  int rc = encode_act_map(encode, &tmdy, dy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dy_channels, w, h, N, ldy, 64, p.BW, p.BH,
                          1, true, "dy");
  if (rc) return rc;
  rc = encode_act_map(encode, &tmx, x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Cin, w, h, N, Cin, 64, p.BW, p.BH, 1, true, "x");
  if (rc) return rc;
  EESEG_CUDA(ensure_max_smem(conv_wgrad_kernel, 227 * 1024));
  conv_wgrad_kernel<<<base_items * p.ksplit, kWgThreads, smem_bytes, stream>>>(tmdy, tmx, p);
  rc = check_launch("conv_wgrad_kernel");
  if (rc) return rc;
  if (to_param && R * S > 1) {
    wgrad_reduce_to_param_kernel<<<dim3((unsigned)((Cin + 255) / 256), (unsigned)Cout), 256, (size_t)R * S * 257 * sizeof(float),
                                   stream>>>(reinterpret_cast<const float*>(workspace), p.ksplit, dw_elems, Cout, R * S, Cin,
                                             to_param == 2 ? 1 : 0, dw);
    return check_launch("wgrad_reduce_to_param_kernel");
  }
  if (p.ksplit == 1 && !to_param) return rc;
  const int64_t n4 = dw_elems / 4;
  int64_t blocks = (n4 + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  // 1x1 kernels: [Cout][1][1][Cin] is the parameter's layout already
  wgrad_reduce_kernel<<<(unsigned)blocks, 256, 0, stream>>>(reinterpret_cast<const float4*>(workspace), p.ksplit, n4,
                                                           reinterpret_cast<float4*>(dw), to_param == 2 ? 1 : 0);
  return check_launch("wgrad_reduce_kernel");
}

extern "C" int eeseg_conv_igemm_wgrad(const void* x, const void* dy, int64_t ldy, int dy_channels, int co_off, int N, int h,
                                      int w, int Cin, int Cout, int R, int S, int dilation, float* dw, void* workspace,
                                      void* stream_) {
  return wgrad_launch(x, dy, ldy, dy_channels, co_off, N, h, w, Cin, Cout, R, S, dilation, dw, workspace, 0, stream_);
}

extern "C" size_t eeseg_conv_igemm_wgrad_to_param_workspace_bytes(int N, int h, int w, int Cin, int Cout, int R, int S) {
  if (N < 1 || h < 1 || w < 1 || Cin < 64 || Cout < 64 || R < 1 || S < 1) return 256;
  WgradParams p;
  wgrad_geometry(N, h, w, Cin, Cout, R, S, p);
  return (size_t)(p.ksplit > 1 ? p.ksplit : 1) * Cout * R * S * Cin * sizeof(float) + 256;
}

extern "C" int eeseg_conv_igemm_wgrad_to_param(const void* x, const void* dy, int64_t ldy, int dy_channels, int co_off, int N,
                                               int h, int w, int Cin, int Cout, int R, int S, int dilation, float* grad,
                                               int accumulate, void* workspace, void* stream_) {
  EESEG_REQUIRE(workspace, "conv_wgrad_to_param: null workspace");
  return wgrad_launch(x, dy, ldy, dy_channels, co_off, N, h, w, Cin, Cout, R, S, dilation, grad, workspace, accumulate ? 2 : 1,
                      stream_);
}

extern "C" int eeseg_conv_weight_rot180_t(const void* w, int Cout, int R, int S, int Cin, void* out, void* stream_) {
  EESEG_REQUIRE(w && out, "conv_weight_rot180_t: null pointer");
  EESEG_REQUIRE(Cout >= 1 && Cin >= 1 && R >= 1 && S >= 1, "conv_weight_rot180_t: bad sizes");
  dim3 grid((Cin + 31) / 32, (Cout + 31) / 32, R * S);
  weight_rot180_t_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream_>>>((const __nv_bfloat16*)w, Cout, R * S, Cin,
                                                                         (__nv_bfloat16*)out);
  return check_launch("weight_rot180_t_kernel");
}

// ---- input gradient: dX = conv(dY, rot180(W)^T), same dilation, 'same' padding, on the forward kernel ----
static size_t dgrad_w_bytes(int Cin, int Cout, int R, int S) {
  return (((size_t)Cin * R * S * Cout * 2) + 255) & ~(size_t)255;
}

extern "C" size_t eeseg_conv_igemm_dgrad_workspace_bytes(int Cin, int Cout, int R, int S) {
  if (Cin <= 0 || Cout <= 0 || R <= 0 || S <= 0) return 256;
  return dgrad_w_bytes(Cin, Cout, R, S) + 2 * (((size_t)Cin * 4 + 255) & ~(size_t)255) + 256;
}

extern "C" int eeseg_conv_igemm_dgrad(const void* dy, const void* wt, int N, int h, int w, int Cin, int Cout, int R, int S,
                                      int dilation, void* dx, int dx_dtype, int64_t lddx, void* workspace, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  EESEG_REQUIRE(dy && wt && dx && workspace, "conv_dgrad: null pointer");
  EESEG_REQUIRE(Cout % 64 == 0, "conv_dgrad: Cout=%d must be a multiple of 64 (it is the contraction dimension)", Cout);
  EESEG_REQUIRE(Cin % 16 == 0, "conv_dgrad: Cin=%d must be a multiple of 16", Cin);
  EESEG_REQUIRE((R & 1) && (S & 1), "conv_dgrad: odd kernel sizes ('same' padding)");
  EESEG_REQUIRE(((uintptr_t)workspace & 255) == 0, "conv_dgrad: workspace must be 256-byte aligned");
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  void* wT = ws;
  float* scale = reinterpret_cast<float*>(ws + dgrad_w_bytes(Cin, Cout, R, S));
  float* shift = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(scale) + (((size_t)Cin * 4 + 255) & ~(size_t)255));
  int rc = eeseg_conv_weight_rot180_t(wt, Cout, R, S, Cin, wT, stream_);
  if (rc) return rc;
  fill_scale_shift_kernel<<<(Cin + 255) / 256, 256, 0, stream>>>(scale, shift, Cin);
  rc = check_launch("fill_scale_shift_kernel");
  if (rc) return rc;
  // the forward kernel with the roles of the channel dimensions swapped: input dY [N,h,w,Cout], weights
  // [Cin][R][S][Cout], output dX [N,h,w,Cin]
  return eeseg_conv_igemm_fwd(dy, wT, scale, shift, 0, N, h, w, Cout, Cin, R, S, dilation, 1, -1, 0, nullptr, 0, dx, dx_dtype,
                              lddx, stream_);
}
