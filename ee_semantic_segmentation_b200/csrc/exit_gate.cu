// Exit gate: fused bilinear up-sample + per-pixel softmax + normalised entropy + argmax + threshold,
// with ordered per-block partial sums for the per-image score, and the per-image exit decision +
// active-image compaction. Reference contract: from_deepv3_new.py:149-152 (interpolate),
// eval_br_ent.py:19-36 (img_norm_entropy), eval_br_ent.py:57-64 / ee_dnn_op_ne.py:80-87 (gate).
//
// Mapping: one thread per output column X (lanes = consecutive X, so every plane store is a
// coalesced 128 B / 64 B request), each thread walks a strip of kRowsPerStrip output rows keeping
// the horizontally interpolated low-res rows (top / bottom) in registers: the four-tap lerp costs
// 2*C cached loads per low-res row step instead of 4*C per pixel. Low-res logits (<= 1 MB per
// image) stay L1/L2 resident; HBM traffic is the optional full-res output planes.
#include "common.cuh"

namespace eeseg {

constexpr int kGateThreads = 128;
constexpr int kRowsPerStrip = 8;

struct GateParams {
  const void* in;
  int64_t in_sn, in_sc, in_sy, in_sx;
  int N, C, h, w, H, W;
  int in_kind;  // 0 logits, 1 probabilities
  float tau;
  float scale_y, scale_x;
  void* up;
  int64_t up_sn;
  float* ent;
  uint8_t* amax;
  uint8_t* mask;
  double* part_sum;
  int32_t* part_cnt;
  int stats;  // any of ent/amax/mask/part_* requested
  int need_ent;  // entropy needed (ent/mask/part_*); argmax-only calls skip the softmax work
};

__device__ __forceinline__ void src_index(int dst, float scale, int in_size, int& i0, int& i1,
                                          float& l0, float& l1) {
  // ATen area_pixel_compute_source_index(align_corners=False): max(scale*(dst+0.5)-0.5, 0)
  float src = fmaxf(scale * ((float)dst + 0.5f) - 0.5f, 0.f);
  i0 = min((int)src, in_size - 1);
  i1 = min(i0 + 1, in_size - 1);
  l1 = src - (float)i0;
  l0 = 1.f - l1;
}

// CMAX values 24/32/64 are register-array capacities with a runtime class count (guards in the
// unrolled loops); any other CMAX (19 = Cityscapes, 21 = VOC) is the exact class count, so every
// guard folds away at compile time.
__host__ __device__ constexpr bool cmax_is_exact(int cmax) { return cmax != 24 && cmax != 32 && cmax != 64; }

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Horizontal lerp of one low-res row into registers: out[c] = lx0*L[x0][c] + lx1*L[x1][c].
// NHWC_VEC: channels contiguous and 16 B aligned (the conv kernel's [N,h,w,Cp] fp32 output):
// 128-bit loads; otherwise scalar loads with a runtime channel stride.
template <typename TI, int CMAX, bool NHWC_VEC>
__device__ __forceinline__ void lerp_row(const TI* __restrict__ row, int off0, int off1, int sc, int C,
                                         float lx0, float lx1, float (&out)[CMAX]) {
  if constexpr (NHWC_VEC) {
    const float4* a = reinterpret_cast<const float4*>(row + off0);
    const float4* b = reinterpret_cast<const float4*>(row + off1);
#pragma unroll
    for (int q = 0; q < (CMAX + 3) / 4; ++q) {
      if (q * 4 < C) {
        const float4 u = __ldg(a + q), v = __ldg(b + q);
        if (4 * q + 0 < CMAX) out[4 * q + 0] = lx0 * u.x + lx1 * v.x;
        if (4 * q + 1 < CMAX) out[4 * q + 1] = lx0 * u.y + lx1 * v.y;
        if (4 * q + 2 < CMAX) out[4 * q + 2] = lx0 * u.z + lx1 * v.z;
        if (4 * q + 3 < CMAX) out[4 * q + 3] = lx0 * u.w + lx1 * v.w;
      }
    }
  } else {
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) out[c] = lx0 * ldf(row + off0 + c * sc) + lx1 * ldf(row + off1 + c * sc);
  }
}

template <typename TI, typename TO, int CMAX, bool INTERP, bool NHWC_VEC>
#ifndef EESEG_GATE_MINB
#define EESEG_GATE_MINB 4
#endif
__global__ void __launch_bounds__(kGateThreads, (CMAX <= 24 ? EESEG_GATE_MINB : (CMAX <= 32 ? 3 : 1))) gate_kernel(const GateParams p) {
  // Work item = one WARP: 32 consecutive columns x kRowsPerStrip rows of one image. Items are dealt to
  // warps from a flat index, so a 513-wide image costs 17 warps per strip (not 5 blocks of 4 warps with
  // 3 of the last block's warps parked at a barrier — ncu: 19 % of stall samples), and there is no
  // block barrier: every warp writes its own ordered partial.
  const int xgroups = (p.W + 31) >> 5, strips = (p.H + kRowsPerStrip - 1) / kRowsPerStrip;
  const int item = blockIdx.x * (kGateThreads / 32) + (int)(threadIdx.x >> 5);
  if (item >= xgroups * strips * p.N) return;   // whole warp
  const int n = item / (xgroups * strips);
  const int irem = item - n * (xgroups * strips);   // partial slot inside the image: strip-major
  const int X = (irem % xgroups) * 32 + (int)(threadIdx.x & 31);
  const int Y0 = (irem / xgroups) * kRowsPerStrip;
  const int Y1 = min(Y0 + kRowsPerStrip, p.H);
  const bool live = X < p.W;
  const int C = cmax_is_exact(CMAX) ? CMAX : p.C;
  const TI* in = reinterpret_cast<const TI*>(p.in) + (int64_t)n * p.in_sn;
  TO* up = p.up ? reinterpret_cast<TO*>(p.up) + (int64_t)n * p.up_sn : nullptr;
  const int64_t HW = (int64_t)p.H * p.W;
  const uint32_t upb = (uint32_t)HW * (uint32_t)sizeof(TO);   // output plane stride in bytes (< 2^32, host-checked)
  const float inv_lnC = C > 1 ? 1.f / logf((float)C) : 0.f;
  constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;

  float acc_ent = 0.f;   // <= kRowsPerStrip values in [0,1]: exact enough in fp32, widened once below
  int acc_cnt = 0;

  float top[CMAX], bot[CMAX];
  int x0 = 0, x1 = 0, cur_y0 = -1, cur_y1 = -1;
  float lx0 = 1.f, lx1 = 0.f;
  if (INTERP && live) src_index(X, p.scale_x, p.w, x0, x1, lx0, lx1);
  // the low-res tensor of one image is far below 2^31 elements (checked on the host)
  const int sx = (int)p.in_sx, sy = (int)p.in_sy, sc = (int)p.in_sc;
  const int off0 = x0 * sx, off1 = x1 * sx;

  if (live) {
    for (int Y = Y0; Y < Y1; ++Y) {
      float v[CMAX];
      if (INTERP) {
        int y0, y1;
        float ly0, ly1;
        src_index(Y, p.scale_y, p.h, y0, y1, ly0, ly1);
        if (y0 != cur_y0 || y1 != cur_y1) {
          if (y0 == cur_y1 && cur_y1 >= 0) {
#pragma unroll
            for (int c = 0; c < CMAX; ++c) top[c] = bot[c];
          } else {
            lerp_row<TI, CMAX, NHWC_VEC>(in + y0 * sy, off0, off1, sc, C, lx0, lx1, top);
          }
          if (y1 == y0) {
#pragma unroll
            for (int c = 0; c < CMAX; ++c) bot[c] = top[c];
          } else {
            lerp_row<TI, CMAX, NHWC_VEC>(in + y1 * sy, off0, off1, sc, C, lx0, lx1, bot);
          }
          cur_y0 = y0;
          cur_y1 = y1;
        }
#pragma unroll
        for (int c = 0; c < CMAX; ++c) v[c] = ly0 * top[c] + ly1 * bot[c];
      } else {
        const TI* r = in + (int64_t)Y * p.in_sy + (int64_t)X * p.in_sx;
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
          if (c < C) v[c] = ldf_stream(r + (int64_t)c * p.in_sc);
      }
      const int64_t pix = (int64_t)Y * p.W + X;
      if (up) {   // one widening multiply-add per plane address (see plane_ptr)
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
          if (c < C) stf(plane_ptr(up + pix, (uint32_t)c, upb), v[c]);
      }
      if (p.stats) {
        // max with 3-input FMNMX, then the first index that equals it (2 instructions per class instead
        // of the 3 of a compare-select-select chain). A NaN logit poisons the entropy anyway.
        float m = v[0];
#pragma unroll
        for (int c = 1; c < CMAX; ++c)
          if (c < C) m = fmaxf(m, v[c]);
        int am = 0;
#pragma unroll
        for (int c = CMAX - 1; c >= 1; --c)
          if (c < C) am = (v[c] == m) ? c : am;
        am = (v[0] == m) ? 0 : am;
        if (!p.need_ent) {   // argmax only (the final exit): no softmax / entropy
          if (p.amax) p.amax[(int64_t)n * HW + pix] = (uint8_t)am;
          continue;
        }
        float hn;
        if (p.in_kind == 0) {
          // H = ln S - sum e_c z_c / S with z = v - max, e = exp(z); computed in base 2
          const float m2 = m * kLog2e;
          float Sa[3] = {0.f, 0.f, 0.f}, Ta[3] = {0.f, 0.f, 0.f};   // three interleaved chains (ILP)
#pragma unroll
          for (int c = 0; c < CMAX; ++c)
            if (c < C) {
              const float z2 = fmaf(v[c], kLog2e, -m2);   // (v - m) * log2(e) <= 0
              const float e = ex2_approx(z2);
              Sa[c % 3] += e;
              Ta[c % 3] = fmaf(e, z2, Ta[c % 3]);          // e == 0 contributes 0 (entr(0) = 0)
            }
          const float S = (Sa[0] + Sa[1]) + Sa[2], T = (Ta[0] + Ta[1]) + Ta[2];
          hn = (__log2f(S) - __fdividef(T, S)) * (kLn2 * inv_lnC);
        } else {
          float S = 0.f;
#pragma unroll
          for (int c = 0; c < CMAX; ++c)
            if (c < C) S += v[c];
          float invS = 1.f / S, T = 0.f;
#pragma unroll
          for (int c = 0; c < CMAX; ++c)
            if (c < C) {
              float q = v[c] * invS;
              T -= q > 0.f ? q * __log2f(q) : 0.f;
            }
          hn = T * (kLn2 * inv_lnC);
        }
        hn = fmaxf(hn, 0.f) + 0.f * hn;  // clamp tiny negatives, keep NaN
        if (p.ent) __stcs(p.ent + (int64_t)n * HW + pix, hn);
        if (p.amax) p.amax[(int64_t)n * HW + pix] = (uint8_t)am;
        const bool below = hn < p.tau;
        if (p.mask) p.mask[(int64_t)n * HW + pix] = below ? 1 : 0;
        acc_ent += hn;
        acc_cnt += below ? 1 : 0;
      }
    }
  }

  if (p.part_sum || p.part_cnt) {
    const double ws = warp_sum((double)acc_ent);
    const int wc = warp_sum(acc_cnt);
    if ((threadIdx.x & 31) == 0) {
      const int64_t slot = (int64_t)n * (xgroups * strips) + irem;
      if (p.part_sum) p.part_sum[slot] = ws;
      if (p.part_cnt) p.part_cnt[slot] = wc;
    }
  }
}

__global__ void pool_mean_kernel(const float* __restrict__ ent, int H, int W, int s, int mode,
                                 float* __restrict__ score) {
  const int n = blockIdx.x;
  const float* e = ent + (int64_t)n * H * W;
  const int bh = (H + s - 1) / s, bw = (W + s - 1) / s;
  double acc = 0.0;
  for (int b = threadIdx.x; b < bh * bw; b += blockDim.x) {
    const int by = b / bw, bx = b % bw;
    // skimage pads the end with 0: ragged blocks see zeros (matters for min; entropies are >= 0)
    const bool ragged = (by + 1) * s > H || (bx + 1) * s > W;
    float r = mode == 0 ? 0.f : (ragged ? 0.f : INFINITY);
    bool nan = false;
    for (int yy = by * s; yy < min((by + 1) * s, H); ++yy)
      for (int xx = bx * s; xx < min((bx + 1) * s, W); ++xx) {
        float v = e[(int64_t)yy * W + xx];
        nan |= (v != v);
        r = mode == 0 ? fmaxf(r, v) : fminf(r, v);
      }
    acc += nan ? (double)NAN : (double)r;
  }
  __shared__ double sm[32];
  double ws = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = ws;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sm[i];
    score[n] = (float)(t / (double)((int64_t)bh * bw));
  }
}

// Stage 2a: per-image score = ordered sum of the block partials / HW (one block per image; every
// thread adds a fixed strided subset in order, then a fixed tree -> bit-reproducible).
__global__ void score_kernel(const double* __restrict__ part_sum, const int32_t* __restrict__ part_cnt,
                             int num_partials, int64_t HW, float* __restrict__ score,
                             int64_t* __restrict__ exited_px) {
  const int n = blockIdx.x;
  double t = 0.0;
  long long k = 0;
  for (int i = threadIdx.x; i < num_partials; i += blockDim.x) {
    if (part_sum) t += part_sum[(int64_t)n * num_partials + i];
    if (part_cnt) k += part_cnt[(int64_t)n * num_partials + i];
  }
  __shared__ double st[128];
  __shared__ long long sk[128];
  st[threadIdx.x] = t;
  sk[threadIdx.x] = k;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) { st[threadIdx.x] += st[threadIdx.x + o]; sk[threadIdx.x] += sk[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (score) score[n] = (float)(st[0] / (double)HW);
    if (exited_px) exited_px[n] = sk[0];
  }
}

// Stage 2b: decision rule + compaction of the still-active images.
__global__ void decide_kernel(const float* __restrict__ score, int N, float tau, int less_than,
                              int exit_id, int32_t* __restrict__ exit_idx,
                              int32_t* __restrict__ active_list, int32_t* __restrict__ active_count) {
  if (exit_idx) {
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
      const float sc = score[n];
      const bool conf = less_than ? (sc < tau) : (sc > tau);
      if (exit_idx[n] < 0 && conf) exit_idx[n] = exit_id;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && active_list && active_count && exit_idx) {
    int k = 0;
    for (int n = 0; n < N; ++n)
      if (exit_idx[n] < 0) active_list[k++] = n;
    *active_count = k;
  }
}

// Stage commit of the compute-skipping engine: everything that follows a gate's scores for the n images still in flight,
// in one launch — the decision rule (eval_br_ent.py:57-64), the results of the images that leave here scattered to their
// batch positions (score, exit index, arg-max map), the ascending list of survivors with its count, and the survivors' batch
// positions for the next stage. grid = (chunks of the map, n); block (0,0) does the O(n) bookkeeping.
__global__ void __launch_bounds__(256) stage_commit_kernel(const float* __restrict__ score, float tau, int less_than,
                                                           int exit_id, int take_all, const int64_t* __restrict__ act,
                                                           const uint8_t* __restrict__ amax, int n, int64_t HW,
                                                           float* __restrict__ scores_row, int32_t* __restrict__ exit_out,
                                                           uint8_t* __restrict__ pred, const int64_t* __restrict__ px_in,
                                                           int64_t* __restrict__ px_acc, int32_t* __restrict__ keep,
                                                           int32_t* __restrict__ count, int64_t* __restrict__ act_next) {
  const int j = blockIdx.y;
  const float sc = score ? score[j] : 0.f;
  const bool took = take_all || (less_than ? (sc < tau) : (sc > tau));
  const int64_t dst_img = act[j];
  if (took) {
    const uint8_t* s = amax + (int64_t)j * HW;
    uint8_t* d = pred + dst_img * HW;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < HW; i += (int64_t)gridDim.x * 256) d[i] = s[i];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (scores_row) scores_row[dst_img] = sc;
    exit_out[dst_img] = took ? exit_id : -1;
    if (j == 0) {
      int k = 0;
      long long px = 0;
      for (int q = 0; q < n; ++q) {
        const float sq = score ? score[q] : 0.f;
        const bool tq = take_all || (less_than ? (sq < tau) : (sq > tau));
        if (!tq) {
          if (keep) keep[k] = q;
          if (act_next) act_next[k] = act[q];
          ++k;
        }
        if (px_in) px += px_in[q];
      }
      if (count) *count = k;
      if (px_acc && px_in) *px_acc += px;
    }
  }
}

// Batch compaction after a gate: dst[j] = src[active_list[j]] for j < *active_count (all n rows when the count pointer is
// null) — the still-active images' activations moved to the front of the next backbone section's input. Rows are whole
// images (tens of MB): 16-byte vectors, grid.y = destination row, grid-stride along the row; rows past the count cost nothing.
__global__ void __launch_bounds__(256) compact_rows_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst,
                                                           const int32_t* __restrict__ list,
                                                           const int32_t* __restrict__ count, int n_src, int64_t row_vecs) {
  const int j = blockIdx.y;
  if (count && j >= *count) return;
  const int sidx = list[j];
  if (sidx < 0 || sidx >= n_src) return;
  const uint4* s = src + (int64_t)sidx * row_vecs;
  uint4* d = dst + (int64_t)j * row_vecs;
  const int64_t step = (int64_t)gridDim.x * 256 * 4;
  for (int64_t v = (int64_t)blockIdx.x * 1024 + threadIdx.x; v < row_vecs; v += step) {
    uint4 r[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (v + u * 256 < row_vecs) r[u] = __ldcs(s + v + u * 256);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (v + u * 256 < row_vecs) d[v + u * 256] = r[u];
  }
}

// Adjoint of the bilinear up-sampling (F.interpolate backward, from_deepv3_new.py:149,152 under autograd):
//   dlow[n,c,y,x] = sum_{Y,X} wy(Y,y) * wx(X,x) * dout[n,c,Y,X]
// in GATHER form: one thread per low-res pixel walks the (at most ~2*scale+2)^2 output pixels that
// interpolate from it, in a fixed order — no atomics, bit-reproducible (ATen's kernel scatters with atomics).
// Lanes = consecutive low-res x, so a warp reads contiguous row segments of the dout plane.
__device__ __forceinline__ float up_weight(int dst, float scale, int in_size, int src_idx) {
  int i0, i1;
  float l0, l1;
  src_index(dst, scale, in_size, i0, i1, l0, l1);
  return (i0 == src_idx ? l0 : 0.f) + (i1 == src_idx ? l1 : 0.f);   // both when the window is clamped to the edge
}

__device__ __forceinline__ void up_range(int src_idx, float scale, int in_size, int out_size, int& lo, int& hi) {
  // outputs whose source coordinate lies in (src_idx-1, src_idx+1), widened by one (weights are re-checked)
  const float inv = 1.f / scale;
  lo = max((int)floorf(((float)src_idx - 1.f + 0.5f) * inv - 0.5f) - 1, 0);
  hi = min((int)ceilf(((float)src_idx + 1.f + 0.5f) * inv - 0.5f) + 1, out_size - 1);
  if (src_idx == in_size - 1) hi = out_size - 1;   // clamped tail
  if (src_idx == 0) lo = 0;
}

template <typename T>
__global__ void __launch_bounds__(128) upsample_bwd_kernel(const T* __restrict__ dout, int64_t planes, int h, int w, int H,
                                                           int W, float scale_y, float scale_x, float* __restrict__ dlow) {
  const int x = blockIdx.x * 128 + threadIdx.x;
  const int y = blockIdx.y;
  const int64_t plane = blockIdx.z;
  if (x >= w) return;
  int Ylo, Yhi, Xlo, Xhi;
  up_range(y, scale_y, h, H, Ylo, Yhi);
  up_range(x, scale_x, w, W, Xlo, Xhi);
  const T* g = dout + plane * (int64_t)H * W;
  float acc = 0.f;
  // horizontal weights once per thread (the window is <= 2*scale + 4 wide: 20 for the 8x up-sampling of the heads)
  constexpr int kMaxWin = 24;
  float wxs[kMaxWin];
  const int nx = Xhi - Xlo + 1;
  const bool cached = nx <= kMaxWin;
  if (cached) {
#pragma unroll
    for (int i = 0; i < kMaxWin; ++i) wxs[i] = i < nx ? up_weight(Xlo + i, scale_x, w, x) : 0.f;
  }
  for (int Y = Ylo; Y <= Yhi; ++Y) {
    const float wy = up_weight(Y, scale_y, h, y);
    if (wy == 0.f) continue;
    const T* grow = g + (int64_t)Y * W + Xlo;
    float row = 0.f;
    if (cached) {
#pragma unroll
      for (int i = 0; i < kMaxWin; ++i)
        if (i < nx) row = fmaf(wxs[i], ldf(grow + i), row);
    } else {
      for (int X = Xlo; X <= Xhi; ++X) {
        const float wx = up_weight(X, scale_x, w, x);
        if (wx != 0.f) row = fmaf(wx, ldf(grow + (X - Xlo)), row);
      }
    }
    acc = fmaf(wy, row, acc);
  }
  dlow[(plane * h + y) * (int64_t)w + x] = acc;
}

template <typename TI, typename TO>
static int launch_gate(const GateParams& p, bool interp, bool vec, cudaStream_t stream) {
  const int64_t items = (int64_t)((p.W + 31) / 32) * ((p.H + kRowsPerStrip - 1) / kRowsPerStrip) * p.N;
  dim3 grid((unsigned)((items + kGateThreads / 32 - 1) / (kGateThreads / 32)));
#define EESEG_GATE_CASE(CM)                                                                        \
  if (p.C <= CM) {                                                                                  \
    if (interp && vec) gate_kernel<TI, TO, CM, true, true><<<grid, kGateThreads, 0, stream>>>(p);   \
    else if (interp) gate_kernel<TI, TO, CM, true, false><<<grid, kGateThreads, 0, stream>>>(p);    \
    else gate_kernel<TI, TO, CM, false, false><<<grid, kGateThreads, 0, stream>>>(p);               \
    return check_launch("gate_kernel");                                                             \
  }
  if (p.C == 21) { EESEG_GATE_CASE(21) }
  if (p.C == 19) { EESEG_GATE_CASE(19) }
  EESEG_GATE_CASE(24)
  EESEG_GATE_CASE(32)
  EESEG_GATE_CASE(64)
#undef EESEG_GATE_CASE
  set_error("exit_gate: C=%d > 64 classes is not supported by this build", p.C);
  return EESEG_ERR_UNSUPPORTED;
}

static int dispatch_gate(const GateParams& p, int in_dtype, int up_dtype, cudaStream_t stream) {
  const bool interp = !(p.h == p.H && p.w == p.W);
  EESEG_REQUIRE(in_dtype == EESEG_F32 || in_dtype == EESEG_BF16, "exit_gate: in_dtype %d", in_dtype);
  EESEG_REQUIRE(!p.up || up_dtype == EESEG_F32 || up_dtype == EESEG_BF16, "exit_gate: up_dtype %d", up_dtype);
  if (in_dtype == EESEG_F32) {
    // 128-bit path: fp32 channels contiguous, every pixel 16 B aligned and padded to a multiple of 4
    const int cpad = (p.C + 3) / 4 * 4;
    const bool vec = interp && p.in_sc == 1 && p.in_sx >= cpad && (p.in_sx % 4) == 0 && (p.in_sy % 4) == 0 &&
                     (p.in_sn % 4) == 0 && ((uintptr_t)p.in % 16) == 0;
    if (up_dtype == EESEG_BF16 && p.up) return launch_gate<float, __nv_bfloat16>(p, interp, vec, stream);
    return launch_gate<float, float>(p, interp, vec, stream);
  }
  if (up_dtype == EESEG_BF16 || !p.up) return launch_gate<__nv_bfloat16, __nv_bfloat16>(p, interp, false, stream);
  return launch_gate<__nv_bfloat16, float>(p, interp, false, stream);
}

}  // namespace eeseg

using namespace eeseg;

extern "C" int eeseg_exit_gate_num_partials(int H, int W) {
  return ((W + 31) / 32) * ((H + kRowsPerStrip - 1) / kRowsPerStrip);
}

extern "C" int eeseg_exit_gate_pixels(const void* in, int in_dtype, int in_kind, int64_t in_sn,
                                      int64_t in_sc, int64_t in_sy, int64_t in_sx, int N, int C,
                                      int h, int w, int H, int W, float tau, void* up_logits,
                                      int up_dtype, int64_t up_sn, float* ent, uint8_t* amax,
                                      uint8_t* mask, double* part_sum, int32_t* part_cnt,
                                      void* stream) {
  EESEG_REQUIRE(in, "exit_gate: null input");
  EESEG_REQUIRE(N >= 0 && C >= 1 && h >= 1 && w >= 1 && H >= 1 && W >= 1, "exit_gate: bad sizes");
  EESEG_REQUIRE(N <= 65535, "exit_gate: N=%d > 65535", N);
  EESEG_REQUIRE(in_kind == 0 || in_kind == 1, "exit_gate: in_kind %d", in_kind);
  EESEG_REQUIRE(!amax || C <= 256, "exit_gate: uint8 argmax needs C <= 256");
  EESEG_REQUIRE((int64_t)h * in_sy < (1ll << 31) && (int64_t)w * in_sx < (1ll << 31) && (int64_t)C * in_sc < (1ll << 31),
                "exit_gate: one low-res image must span fewer than 2^31 elements");
  EESEG_REQUIRE((int64_t)H * W < (1ll << 29), "exit_gate: H*W must be < 2^29 pixels");
  if (N == 0) return EESEG_OK;
  GateParams p;
  p.in = in; p.in_sn = in_sn; p.in_sc = in_sc; p.in_sy = in_sy; p.in_sx = in_sx;
  p.N = N; p.C = C; p.h = h; p.w = w; p.H = H; p.W = W; p.in_kind = in_kind; p.tau = tau;
  p.scale_y = (float)h / (float)H; p.scale_x = (float)w / (float)W;
  p.up = up_logits; p.up_sn = up_sn; p.ent = ent; p.amax = amax; p.mask = mask;
  p.part_sum = part_sum; p.part_cnt = part_cnt;
  p.stats = (ent || amax || mask || part_sum || part_cnt) ? 1 : 0;
  p.need_ent = (ent || mask || part_sum || part_cnt) ? 1 : 0;
  return dispatch_gate(p, in_dtype, up_dtype, (cudaStream_t)stream);
}

extern "C" int eeseg_upsample_bilinear(const void* in, int in_dtype, int64_t in_sn, int64_t in_sc,
                                       int64_t in_sy, int64_t in_sx, int N, int C, int h, int w,
                                       int H, int W, void* out, int out_dtype, int64_t out_sn,
                                       void* stream) {
  EESEG_REQUIRE(out, "upsample: null output");
  return eeseg_exit_gate_pixels(in, in_dtype, 0, in_sn, in_sc, in_sy, in_sx, N, C, h, w, H, W, 0.f,
                                out, out_dtype, out_sn, nullptr, nullptr, nullptr, nullptr, nullptr,
                                stream);
}

extern "C" int eeseg_entropy_pool_mean(const float* ent, int N, int H, int W, int s, int mode,
                                       float* score, void* stream) {
  EESEG_REQUIRE(ent && score, "entropy_pool_mean: null pointer");
  EESEG_REQUIRE(s >= 1 && (mode == 0 || mode == 1), "entropy_pool_mean: s=%d mode=%d", s, mode);
  if (N == 0) return EESEG_OK;
  pool_mean_kernel<<<N, 1024, 0, (cudaStream_t)stream>>>(ent, H, W, s, mode, score);
  return check_launch("pool_mean_kernel");
}

extern "C" int eeseg_exit_gate_decide(const double* part_sum, const int32_t* part_cnt,
                                      int num_partials, const float* score_in, int N, int64_t HW,
                                      float tau, int less_than, int exit_id, int32_t* exit_idx,
                                      float* score_out, int64_t* exited_px, int32_t* active_list,
                                      int32_t* active_count, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  EESEG_REQUIRE(part_sum || score_in, "exit_gate_decide: need part_sum or score_in");
  EESEG_REQUIRE(!part_sum || score_out || !exit_idx, "exit_gate_decide: score_out is required with part_sum when deciding");
  if (N == 0) return EESEG_OK;
  const float* score = score_in;
  if (part_sum) {
    score_kernel<<<N, 128, 0, stream>>>(part_sum, part_cnt, num_partials, HW, score_out, exited_px);
    int rc = check_launch("score_kernel");
    if (rc) return rc;
    score = score_out;
  }
  if (exit_idx) {
    decide_kernel<<<1, 256, 0, stream>>>(score, N, tau, less_than, exit_id, exit_idx, active_list, active_count);
    return check_launch("decide_kernel");
  }
  return EESEG_OK;
}

extern "C" int eeseg_exit_stage_commit(const float* score, float tau, int less_than, int exit_id, int take_all,
                                       const int64_t* positions, const uint8_t* amax, int n, int64_t HW, float* scores_row,
                                       int32_t* exit_idx, uint8_t* pred, const int64_t* exited_px_in, int64_t* exited_px_acc,
                                       int32_t* active_list, int32_t* active_count, int64_t* next_positions, void* stream) {
  EESEG_REQUIRE(positions && amax && exit_idx && pred, "exit_stage_commit: null pointer");
  EESEG_REQUIRE(score || take_all, "exit_stage_commit: scores are required unless every image leaves");
  EESEG_REQUIRE(n >= 0 && n <= 65535 && HW >= 1, "exit_stage_commit: bad sizes");
  if (n == 0) {
    if (active_count) EESEG_CUDA(cudaMemsetAsync(active_count, 0, sizeof(int32_t), (cudaStream_t)stream));
    return EESEG_OK;
  }
  int bx = (int)((HW + 4095) / 4096);
  if (bx > 64) bx = 64;
  stage_commit_kernel<<<dim3((unsigned)bx, (unsigned)n), 256, 0, (cudaStream_t)stream>>>(
      score, tau, less_than, exit_id, take_all, positions, amax, n, HW, scores_row, exit_idx, pred, exited_px_in,
      exited_px_acc, active_list, active_count, next_positions);
  return check_launch("stage_commit_kernel");
}

extern "C" int eeseg_compact_rows(const void* src, void* dst, const int32_t* active_list, const int32_t* active_count,
                                  int n_src, int n_dst, int64_t row_bytes, void* stream) {
  EESEG_REQUIRE(src && dst && active_list, "compact_rows: null pointer");
  EESEG_REQUIRE(n_src >= 0 && n_dst >= 0 && n_dst <= 65535 && row_bytes >= 0, "compact_rows: bad sizes");
  EESEG_REQUIRE((row_bytes & 15) == 0 && (((uintptr_t)src | (uintptr_t)dst) & 15) == 0,
                "compact_rows: rows must be 16-byte multiples and 16-byte aligned");
  if (n_dst == 0 || row_bytes == 0) return EESEG_OK;
  const int64_t row_vecs = row_bytes >> 4;
  int64_t bx = (row_vecs + 1023) / 1024;
  const int64_t cap = (8 * kNumSMs + n_dst - 1) / n_dst;      // ~8 blocks per SM over all rows
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  compact_rows_kernel<<<dim3((unsigned)bx, (unsigned)n_dst), 256, 0, (cudaStream_t)stream>>>(
      (const uint4*)src, (uint4*)dst, active_list, active_count, n_src, row_vecs);
  return check_launch("compact_rows_kernel");
}

extern "C" int eeseg_upsample_bilinear_bwd(const void* dout, int dtype, int64_t planes, int h, int w, int H, int W,
                                           float* dlow, void* stream) {
  EESEG_REQUIRE(dout && dlow, "upsample_bwd: null pointer");
  EESEG_REQUIRE(planes >= 0 && h >= 1 && w >= 1 && H >= 1 && W >= 1 && h <= 65535 && planes <= 0x7fffffff,
                "upsample_bwd: bad sizes");
  EESEG_REQUIRE(dtype == EESEG_F32 || dtype == EESEG_BF16, "upsample_bwd: dtype %d", dtype);
  if (planes == 0) return EESEG_OK;
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  // grid.z carries the planes in chunks of 65535
  for (int64_t p0 = 0; p0 < planes; p0 += 65535) {
    const int64_t np = planes - p0 < 65535 ? planes - p0 : 65535;
    dim3 grid((w + 127) / 128, h, (unsigned)np);
    if (dtype == EESEG_F32)
      upsample_bwd_kernel<float><<<grid, 128, 0, (cudaStream_t)stream>>>((const float*)dout + p0 * (int64_t)H * W, np, h, w, H, W,
                                                                        sy, sx, dlow + p0 * (int64_t)h * w);
    else
      upsample_bwd_kernel<__nv_bfloat16><<<grid, 128, 0, (cudaStream_t)stream>>>(
          (const __nv_bfloat16*)dout + p0 * (int64_t)H * W, np, h, w, H, W, sy, sx, dlow + p0 * (int64_t)h * w);
    int rc = check_launch("upsample_bwd_kernel");
    if (rc) return rc;
  }
  return EESEG_OK;
}
