// Lovasz-softmax, forward + gradient, for all classes of one exit at a time.
// Reference contract: lovasz_softmax / lovasz_softmax_flat / flatten_probas / lovasz_grad
// (lovaszsoftmax.py:154-219, 19-31), looped over exits by BSL.LovaszSoftmax.forward
// (branchy_seg_losses.py:151-159) — see include/eeseg.h.
//
// Per (group g, class c) segment ("group" = whole batch, or one image when per_image):
//   err_i = |[label_i == c] - p_i,c|   (void pixels take err = 0, fg = 0: they sort last and add 0)
//   sort err descending, F_k = #foreground among the first k+1, G = #foreground,
//   J_k = 1 - (G - F_k) / (G + k + 1 - F_k),  g_k = J_k - J_{k-1},  loss_c = sum_k err_k g_k
//   d loss_c / d p_i = sign(p_i - fg_i) g_rank(i)
// The reference loops classes in Python with one torch.sort each (and a host sync per class); here
// all C segments are sorted together by a hand-written stable LSD radix sort (4 passes of 8 bits on
// the bit pattern of err; payload = pixel index | fg << 31 | sign << 30), followed by a segmented
// scan of the fg bit and one apply kernel that forms g_k from exact integer counts, accumulates
// err*g in fp64 per tile (ordered reduction) and scatters the gradient.
// HBM-bound, sort-dominated: 4 passes x (8 B hist read + 8 B read + 8 B write) per element; the
// scatter re-orders each tile in shared memory first so the global writes are whole runs.
#include "common.cuh"

namespace eeseg {

constexpr int kRsThreads = 256;
constexpr int kRsWarps = kRsThreads / 32;
constexpr int kRsItems = 16;
constexpr int kRsTile = kRsThreads * kRsItems;  // 4096 elements per block

struct LovaszDims {
  int N, C, G;     // G = number of groups (1, or N when per_image)
  int64_t HW, L;   // L = elements per segment (N*HW or HW)
  int T;           // tiles per segment
};

// ---- label statistics -------------------------------------------------------------------------
__global__ void lv_label_hist_kernel(const int64_t* __restrict__ labels, LovaszDims d,
                                     int has_ignore, int64_t ignore, int* __restrict__ counts) {
  // counts[g][c] = #valid pixels of group g with label c
  extern __shared__ unsigned sh[];
  const int g = blockIdx.y;
  for (int i = threadIdx.x; i < d.C; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const int64_t* lab = labels + (d.G == 1 ? 0 : (int64_t)g * d.HW);
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31); i0 < d.L;
       i0 += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = i0 + (threadIdx.x & 31);
    int key = -1;
    if (i < d.L) {
      int64_t t = __ldg(lab + i);
      if (!(has_ignore && t == ignore) && t >= 0 && t < d.C) key = (int)t;
    }
    unsigned peers = __match_any_sync(0xffffffffu, key);
    if (key >= 0 && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(sh + key, (unsigned)__popc(peers));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < d.C; i += blockDim.x)
    if (sh[i]) atomicAdd(counts + g * d.C + i, (int)sh[i]);
}

__global__ void lv_segments_kernel(const int* __restrict__ counts, int G, int C, int classes_mode,
                                   int* __restrict__ skip, int* __restrict__ npresent) {
  // skip[g*C+c] = 1 when the class is absent and classes == 'present'; npresent[g] = #classes averaged
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    int k = 0;
    for (int c = 0; c < C; ++c) {
      const bool use = classes_mode == 1 || counts[g * C + c] > 0;
      skip[g * C + c] = use ? 0 : 1;
      k += use ? 1 : 0;
    }
    npresent[g] = k;
  }
}

// ---- key generation ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) lv_keygen_kernel(const T* __restrict__ probas,
                                                         const int64_t* __restrict__ labels,
                                                         LovaszDims d, int has_ignore,
                                                         int64_t ignore, const int* __restrict__ skip,
                                                         uint2* __restrict__ kv,
                                                         T* __restrict__ dprobas) {
  // thread per pixel of the batch; loops classes so the int64 label is read once
  const int64_t P = (int64_t)d.N * d.HW;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < P;
       q += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(q / d.HW);
    const int64_t pix = q - (int64_t)n * d.HW;
    const int64_t lab = __ldg(labels + q);
    const bool valid = !(has_ignore && lab == ignore);
    const int g = d.G == 1 ? 0 : n;
    const int64_t i = d.G == 1 ? q : pix;  // index inside the segment
    // classes in chunks of 4: the chunk's loads are issued before the first use (the loop is latency bound otherwise)
    for (int c0 = 0; c0 < d.C; c0 += 4) {
      float pv[4];
      bool sk[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = c0 + u;
        sk[u] = c >= d.C || skip[g * d.C + c] != 0;
        pv[u] = sk[u] ? 0.f : ldf_stream(probas + ((int64_t)n * d.C + c) * d.HW + pix);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = c0 + u;
        if (c >= d.C) break;
        const int64_t src = ((int64_t)n * d.C + c) * d.HW + pix;
        if (sk[u]) {
          if (dprobas) stf(dprobas + src, 0.f);
          continue;
        }
        const float fg = (lab == c) ? 1.f : 0.f;
        const float diff = pv[u] - fg;
        const float err = valid ? fabsf(diff) : 0.f;
        const int64_t dst = ((int64_t)g * d.C + c) * d.L + i;
        // key: ascending on ~bits == descending on err (err >= 0); value: index | fg << 31 | sign << 30
        kv[dst] = make_uint2(~__float_as_uint(err),
                             (uint32_t)i | (fg != 0.f ? 0x80000000u : 0u) | (diff > 0.f ? 0x40000000u : 0u));
      }
    }
  }
}

// ---- LSD radix sort, one 8-bit pass = hist + scan (2 small kernels) + scatter ----------------------
// Lanes holding the same 8-bit digit, in constant time (8 ballots) whatever the digit distribution
// (MATCH.ANY serialises over the distinct values, i.e. it is slowest exactly on the random low bytes).
__device__ __forceinline__ unsigned match_digit8(int dg, bool ok) {
  unsigned peers = __ballot_sync(0xffffffffu, ok);
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const bool bit = (dg >> b) & 1;
    const unsigned bal = __ballot_sync(0xffffffffu, bit);
    peers &= bit ? bal : ~bal;
  }
  return ok ? peers : 0u;
}

// Tile histograms. Counting goes through per-warp private shared-memory histograms with plain atomics: random digits
// (the mantissa bytes) spread over the banks and cost ~1 instruction per element, where digit matching by ballots made the
// kernel ALU bound (ncu: math-pipe throttle the top stall, 31 % of HBM peak). Concentrated digits (kTop: the sign/exponent
// byte takes a handful of values, so the atomics would serialise 32-way) are aggregated with MATCH.ANY first, which is
// cheap exactly when there are few distinct values; a warp whose 32 digits coincide adds them with one atomic either way.
template <bool kTop>
__global__ void __launch_bounds__(kRsThreads) rs_hist_kernel(const uint2* __restrict__ kv,
                                                              LovaszDims d, int shift,
                                                              const int* __restrict__ skip,
                                                              uint32_t* __restrict__ hist) {
  // hist[seg][digit][tile]
  const int seg = blockIdx.y, tile = blockIdx.x;
  if (skip[seg]) return;
  __shared__ unsigned h[kRsWarps][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int w = 0; w < kRsWarps; ++w) h[w][threadIdx.x] = 0;
  __syncthreads();
  const uint2* k = kv + (int64_t)seg * d.L;
  const int64_t base = (int64_t)tile * kRsTile;
  int dg[kRsItems];
#pragma unroll
  for (int r = 0; r < kRsItems; ++r) {   // all loads first (16 independent loads in flight per thread)
    const int64_t i = base + r * kRsThreads + threadIdx.x;
    dg[r] = i < d.L ? (int)((__ldg(k + i).x >> shift) & 255u) : 256;
  }
  unsigned* hw = h[warp];
#pragma unroll
  for (int r = 0; r < kRsItems; ++r) {
    if (kTop) {
      const unsigned peers = __match_any_sync(0xffffffffu, dg[r]);
      if (dg[r] < 256 && lane == __ffs(peers) - 1) atomicAdd(hw + dg[r], (unsigned)__popc(peers));
    } else {
      int same;
      __match_all_sync(0xffffffffu, dg[r], &same);
      if (same) {
        if (lane == 0 && dg[r] < 256) atomicAdd(hw + dg[r], 32u);
      } else if (dg[r] < 256) {
        atomicAdd(hw + dg[r], 1u);
      }
    }
  }
  __syncthreads();
  unsigned t = 0;
#pragma unroll
  for (int w = 0; w < kRsWarps; ++w) t += h[w][threadIdx.x];
  hist[((int64_t)seg * 256 + threadIdx.x) * d.T + tile] = t;
}

// one warp per (segment, digit): exclusive scan of that digit's counts over the tiles
__global__ void __launch_bounds__(256) rs_scan_tiles_kernel(uint32_t* __restrict__ hist, int T,
                                                             const int* __restrict__ skip,
                                                             uint32_t* __restrict__ digit_total) {
  const int w = blockIdx.x * 8 + (threadIdx.x >> 5);   // = seg * 256 + digit
  const int seg = w >> 8, lane = threadIdx.x & 31;
  if (skip[seg]) return;
  uint32_t* row = hist + (int64_t)w * T;
  unsigned running = 0;
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int t = t0 + lane;
    unsigned v = t < T ? row[t] : 0u, inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    if (t < T) row[t] = running + inc - v;
    running += __shfl_sync(0xffffffffu, inc, 31);
  }
  if (lane == 0) digit_total[w] = running;
}

// one block per segment: exclusive scan of the 256 digit totals
__global__ void __launch_bounds__(256) rs_scan_digits_kernel(const uint32_t* __restrict__ digit_total,
                                                              const int* __restrict__ skip,
                                                              uint32_t* __restrict__ digit_base) {
  const int seg = blockIdx.x;
  if (skip[seg]) return;
  __shared__ unsigned wsum[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned v = digit_total[seg * 256 + threadIdx.x];
  unsigned inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned u = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += u;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  unsigned off = 0;
  for (int i = 0; i < warp; ++i) off += wsum[i];
  digit_base[seg * 256 + threadIdx.x] = off + inc - v;
}

// Scatter: stable ranking inside the tile (warp-blocked order, match.any), then the tile is
// re-ordered by digit in shared memory and written out run by run, so consecutive threads write
// consecutive (key, value) pairs: full lines instead of one 32 B sector per element.
template <bool kMatchAny>
__global__ void __launch_bounds__(kRsThreads, 3) rs_scatter_kernel(
    const uint2* __restrict__ in, uint2* __restrict__ out, LovaszDims d, int shift,
    const int* __restrict__ skip, const uint32_t* __restrict__ hist,
    const uint32_t* __restrict__ digit_base) {
  const int seg = blockIdx.y, tile = blockIdx.x;
  if (skip[seg]) return;
  __shared__ unsigned cnt[kRsWarps][256];     // per-warp digit counts -> exclusive prefix over warps
  __shared__ unsigned gl_off[256];            // global position of the tile's first element of a digit
  __shared__ unsigned wsum[8];
  __shared__ uint2 stage[kRsTile];            // 32 KB
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kRsWarps * 256; i += kRsThreads) (&cnt[0][0])[i] = 0;
  __syncthreads();
  const int64_t seg_off = (int64_t)seg * d.L;
  // warp-blocked order keeps the sort stable: warp w owns a contiguous run, round r is contiguous
  const int64_t wbase = (int64_t)tile * kRsTile + (int64_t)warp * (32 * kRsItems);
  uint2 kv[kRsItems];
  unsigned lrank[kRsItems];
#pragma unroll
  for (int r = 0; r < kRsItems; ++r) {
    const int64_t i = wbase + r * 32 + lane;
    kv[r] = i < d.L ? __ldg(in + seg_off + i) : make_uint2(0, 0);
  }
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < kRsItems; ++r) {
    const bool ok = wbase + r * 32 + lane < d.L;
    const int dg = ok ? (int)((kv[r].x >> shift) & 255u) : 0;
    const unsigned peers = kMatchAny ? (__match_any_sync(0xffffffffu, ok ? dg : 256) & (ok ? 0xffffffffu : 0u))
                                     : match_digit8(dg, ok);
    const int leader = ok ? __ffs(peers) - 1 : lane;
    unsigned old = 0;
    if (ok && lane == leader) {
      old = cnt[warp][dg];
      cnt[warp][dg] = old + __popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    lrank[r] = old + __popc(peers & lt);
    __syncwarp();
  }
  __syncthreads();
  {
    const int dg = threadIdx.x;  // kRsThreads == 256 digits
    unsigned run = 0;
#pragma unroll
    for (int w = 0; w < kRsWarps; ++w) {
      unsigned t = cnt[w][dg];
      cnt[w][dg] = run;
      run += t;
    }
    // exclusive prefix of the tile's digit totals (block scan over 256 threads)
    unsigned inc = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    unsigned off = 0;
    for (int i = 0; i < warp; ++i) off += wsum[i];
    // one table lookup per element in each of the two phases below: cnt[w][dg] becomes the position of warp w's first
    // element of the digit inside the staged tile, gl_off[dg] the global position of staged slot 0 of the digit's run
    const unsigned tp = off + inc - run;
#pragma unroll
    for (int w = 0; w < kRsWarps; ++w) cnt[w][dg] += tp;
    gl_off[dg] = digit_base[seg * 256 + dg] + hist[((int64_t)seg * 256 + dg) * d.T + tile] - tp;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kRsItems; ++r) {
    if (wbase + r * 32 + lane < d.L) {
      const int dg = (int)((kv[r].x >> shift) & 255u);
      stage[cnt[warp][dg] + lrank[r]] = kv[r];
    }
  }
  __syncthreads();
  const int64_t remaining = d.L - (int64_t)tile * kRsTile;
  const int n_valid = (int)(remaining < kRsTile ? remaining : kRsTile);
  uint2* o = out + seg_off;
#pragma unroll 4
  for (int i = threadIdx.x; i < n_valid; i += kRsThreads) {
    const uint2 e = stage[i];
    const int dg = (int)((e.x >> shift) & 255u);
    o[gl_off[dg] + (unsigned)i] = e;      // unsigned wrap-around: gl_off may be "negative" by less than the tile size
  }
}

// ---- segmented scan of the fg bit + Jaccard gradient + dot + gradient scatter --------------------
__global__ void __launch_bounds__(kRsThreads) lv_fgcount_kernel(const uint2* __restrict__ kv,
                                                                 LovaszDims d,
                                                                 const int* __restrict__ skip,
                                                                 uint32_t* __restrict__ tile_fg) {
  const int seg = blockIdx.y, tile = blockIdx.x;
  if (skip[seg]) return;
  const uint2* v = kv + (int64_t)seg * d.L;
  int c = 0;
  for (int r = 0; r < kRsItems; ++r) {
    const int64_t i = (int64_t)tile * kRsTile + r * kRsThreads + threadIdx.x;
    c += (i < d.L) ? (int)(__ldg(v + i).y >> 31) : 0;
  }
  c = warp_sum(c);
  __shared__ int s[kRsWarps];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < kRsWarps; ++i) t += s[i];
    tile_fg[(int64_t)seg * d.T + tile] = (uint32_t)t;
  }
}

__global__ void lv_fgscan_kernel(uint32_t* __restrict__ tile_fg, int T, const int* __restrict__ skip) {
  const int seg = blockIdx.x, lane = threadIdx.x;  // one warp per segment
  if (skip[seg]) return;
  uint32_t* row = tile_fg + (int64_t)seg * T;
  unsigned running = 0;
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int t = t0 + lane;
    unsigned v = t < T ? row[t] : 0u, inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    if (t < T) row[t] = running + inc - v;
    running += __shfl_sync(0xffffffffu, inc, 31);
  }
}

template <typename T>
__global__ void __launch_bounds__(kRsThreads) lv_apply_kernel(
    const uint2* __restrict__ kv, LovaszDims d,
    const int* __restrict__ skip, const int* __restrict__ counts, const int* __restrict__ npresent,
    const uint32_t* __restrict__ tile_fg, double* __restrict__ partial, T* __restrict__ dprobas) {
  const int seg = blockIdx.y, tile = blockIdx.x;
  if (skip[seg]) return;
  const int g = seg / d.C, c = seg % d.C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t seg_off = (int64_t)seg * d.L;
  const int64_t wbase = (int64_t)tile * kRsTile + (int64_t)warp * (32 * kRsItems);
  uint32_t key[kRsItems], val[kRsItems];
  unsigned ballots[kRsItems];
  int wtotal = 0;
#pragma unroll
  for (int r = 0; r < kRsItems; ++r) {
    const int64_t i = wbase + r * 32 + lane;
    const bool ok = i < d.L;
    const uint2 e = ok ? __ldg(kv + seg_off + i) : make_uint2(0xffffffffu, 0u);
    key[r] = e.x;
    val[r] = e.y;
    ballots[r] = __ballot_sync(0xffffffffu, ok && (val[r] >> 31));
    wtotal += __popc(ballots[r]);
  }
  __shared__ int wsum[kRsWarps];
  __shared__ double psum[kRsWarps];
  if (lane == 0) wsum[warp] = wtotal;
  __syncthreads();
  int64_t F = tile_fg[(int64_t)seg * d.T + tile];  // fg before this tile
  for (int w = 0; w < warp; ++w) F += wsum[w];
  const int64_t Gi = counts[g * d.C + c];
  const int np = npresent[g];
  const float inv_norm = np > 0 ? 1.f / ((float)np * (float)d.G) : 0.f;
  const unsigned le = (lane == 31) ? 0xffffffffu : ((2u << lane) - 1u);
  double acc = 0.0;
#pragma unroll
  for (int r = 0; r < kRsItems; ++r) {
    const int64_t k = wbase + r * 32 + lane;
    if (k < d.L) {
      const int64_t fgbit = val[r] >> 31;
      const int64_t Fk = F + __popc(ballots[r] & le);  // inclusive foreground count (exact)
      const int64_t Fp = Fk - fgbit;
      // lovasz_grad (lovaszsoftmax.py:24-30): jaccard = 1 - (gts - cumsum fg) / (gts + cumsum(1-fg))
      const float Jk = 1.f - (float)(Gi - Fk) / (float)(Gi + (k + 1) - Fk);
      const float Jp = k == 0 ? 0.f : 1.f - (float)(Gi - Fp) / (float)(Gi + k - Fp);
      const float gk = Jk - Jp;
      const float err = __uint_as_float(~key[r]);
      acc += (double)err * (double)gk;
      if (dprobas) {
        const int64_t i = (int64_t)(val[r] & 0x3fffffffu);
        const int n = d.G == 1 ? (int)(i / d.HW) : g;
        const int64_t pix = d.G == 1 ? i - (int64_t)n * d.HW : i;
        const float sgn = err == 0.f ? 0.f : ((val[r] & 0x40000000u) ? 1.f : -1.f);
        stf(dprobas + ((int64_t)n * d.C + c) * d.HW + pix, sgn * gk * inv_norm);
      }
    }
    F += __popc(ballots[r]);
  }
  acc = warp_sum(acc);
  if (lane == 0) psum[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kRsWarps; ++w) t += psum[w];
    partial[(int64_t)seg * d.T + tile] = t;
  }
}

__global__ void lv_final_kernel(const double* __restrict__ partial, LovaszDims d,
                                const int* __restrict__ skip, const int* __restrict__ npresent,
                                float* __restrict__ out) {
  // out = mean_g ( mean_{c used} loss[g][c] ), summed in a fixed order
  __shared__ double s[256];
  double acc = 0.0;
  const int S = d.G * d.C;
  for (int seg = 0; seg < S; ++seg) {
    if (skip[seg]) continue;
    const int np = npresent[seg / d.C];
    double a = 0.0;
    for (int t = threadIdx.x; t < d.T; t += blockDim.x) a += partial[(int64_t)seg * d.T + t];
    acc += a / ((double)np * (double)d.G);
  }
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = (float)s[0];
}

// ---- workspace layout ---------------------------------------------------------------------------
struct LovaszWs {
  uint2* kv[2];
  uint32_t *hist, *digit_base, *digit_total, *tile_fg;
  double* partial;
  int *counts, *skip, *npresent;
  size_t bytes;
};

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static LovaszWs carve(void* base, const LovaszDims& d) {
  LovaszWs w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? (char*)base + off : nullptr;
    off += align256(bytes);
    return p;
  };
  const size_t S = (size_t)d.G * d.C, SL = S * (size_t)d.L;
  w.kv[0] = (uint2*)take(SL * 8);
  w.kv[1] = (uint2*)take(SL * 8);
  w.hist = (uint32_t*)take(S * 256 * (size_t)d.T * 4);
  w.digit_base = (uint32_t*)take(S * 256 * 4);
  w.digit_total = (uint32_t*)take(S * 256 * 4);
  w.tile_fg = (uint32_t*)take(S * (size_t)d.T * 4);
  w.partial = (double*)take(S * (size_t)d.T * 8);
  w.counts = (int*)take(S * 4);
  w.skip = (int*)take(S * 4);
  w.npresent = (int*)take((size_t)d.G * 4);
  w.bytes = off;
  return w;
}

static LovaszDims make_dims(int N, int C, int64_t HW, int per_image) {
  LovaszDims d;
  d.N = N; d.C = C; d.HW = HW;
  d.G = per_image ? N : 1;
  d.L = per_image ? HW : (int64_t)N * HW;
  d.T = (int)((d.L + kRsTile - 1) / kRsTile);
  return d;
}

template <typename T>
static int run_exit(const T* probas, const int64_t* labels, const LovaszDims& d, int has_ignore,
                    int64_t ignore, const LovaszWs& w, float* out, T* dprobas, cudaStream_t stream) {
  const int S = d.G * d.C;
  const int64_t P = (int64_t)d.N * d.HW;
  int kb = (int)((P + 255) / 256 < kNumSMs * 8 ? (P + 255) / 256 : kNumSMs * 8);
  lv_keygen_kernel<T><<<kb, 256, 0, stream>>>(probas, labels, d, has_ignore, ignore, w.skip,
                                              w.kv[0], dprobas);
  int rc = check_launch("lv_keygen_kernel");
  if (rc) return rc;
  dim3 tiles(d.T, S);
  int cur = 0;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = pass * 8;
    if (pass == 3)
      rs_hist_kernel<true><<<tiles, kRsThreads, 0, stream>>>(w.kv[cur], d, shift, w.skip, w.hist);
    else
      rs_hist_kernel<false><<<tiles, kRsThreads, 0, stream>>>(w.kv[cur], d, shift, w.skip, w.hist);
    if ((rc = check_launch("rs_hist_kernel"))) return rc;
    rs_scan_tiles_kernel<<<S * 32, 256, 0, stream>>>(w.hist, d.T, w.skip, w.digit_total);
    if ((rc = check_launch("rs_scan_tiles_kernel"))) return rc;
    rs_scan_digits_kernel<<<S, 256, 0, stream>>>(w.digit_total, w.skip, w.digit_base);
    if ((rc = check_launch("rs_scan_digits_kernel"))) return rc;
    if (pass == 3)   // the sign/exponent byte: few distinct digits per warp, MATCH.ANY beats the 8-ballot match there (only)
      rs_scatter_kernel<true><<<tiles, kRsThreads, 0, stream>>>(w.kv[cur], w.kv[cur ^ 1], d, shift, w.skip, w.hist,
                                                                w.digit_base);
    else
      rs_scatter_kernel<false><<<tiles, kRsThreads, 0, stream>>>(w.kv[cur], w.kv[cur ^ 1], d, shift, w.skip, w.hist,
                                                                 w.digit_base);
    if ((rc = check_launch("rs_scatter_kernel"))) return rc;
    cur ^= 1;
  }
  lv_fgcount_kernel<<<tiles, kRsThreads, 0, stream>>>(w.kv[cur], d, w.skip, w.tile_fg);
  if ((rc = check_launch("lv_fgcount_kernel"))) return rc;
  lv_fgscan_kernel<<<S, 32, 0, stream>>>(w.tile_fg, d.T, w.skip);
  if ((rc = check_launch("lv_fgscan_kernel"))) return rc;
  lv_apply_kernel<T><<<tiles, kRsThreads, 0, stream>>>(w.kv[cur], d, w.skip, w.counts,
                                                       w.npresent, w.tile_fg, w.partial, dprobas);
  if ((rc = check_launch("lv_apply_kernel"))) return rc;
  lv_final_kernel<<<1, 256, 0, stream>>>(w.partial, d, w.skip, w.npresent, out);
  return check_launch("lv_final_kernel");
}

}  // namespace eeseg

using namespace eeseg;

extern "C" size_t eeseg_lovasz_workspace_bytes(int E, int N, int C, int64_t HW) {
  (void)E;  // exits are processed one at a time and share the scratch
  if (N <= 0 || C <= 0 || HW <= 0) return 256;
  // per_image changes T slightly; take the larger of the two layouts
  size_t a = carve(nullptr, make_dims(N, C, HW, 0)).bytes;
  size_t b = carve(nullptr, make_dims(N, C, HW, 1)).bytes;
  return (a > b ? a : b) + 256;
}

extern "C" int eeseg_lovasz_fwd_bwd(const void* probas, int dtype, int64_t exit_stride,
                                    const int64_t* labels, int E, int N, int C, int64_t HW,
                                    int has_ignore, int64_t ignore, int classes_mode, int per_image,
                                    float* per_exit, void* dprobas, void* workspace,
                                    size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  EESEG_REQUIRE(probas && labels && per_exit && workspace, "lovasz: null pointer");
  EESEG_REQUIRE(E >= 1 && N >= 1 && C >= 1 && HW >= 1, "lovasz: bad sizes");
  EESEG_REQUIRE(dtype == EESEG_F32 || dtype == EESEG_BF16, "lovasz: dtype %d", dtype);
  EESEG_REQUIRE(classes_mode == 0 || classes_mode == 1, "lovasz: classes_mode %d", classes_mode);
  const LovaszDims d = make_dims(N, C, HW, per_image);
  EESEG_REQUIRE(d.L < (1ll << 30), "lovasz: %lld pixels per segment exceed the 30-bit index", (long long)d.L);
  EESEG_REQUIRE((int64_t)d.G * C <= 65535, "lovasz: too many segments");
  const LovaszWs w = carve(workspace, d);
  EESEG_REQUIRE(w.bytes <= workspace_bytes, "lovasz: workspace too small (%zu < %zu)", workspace_bytes, w.bytes);
  const int S = d.G * C;
  EESEG_CUDA(cudaMemsetAsync(w.counts, 0, sizeof(int) * S, stream));
  int hb = (int)((d.L + 255) / 256 < kNumSMs * 4 ? (d.L + 255) / 256 : kNumSMs * 4);
  lv_label_hist_kernel<<<dim3(hb, d.G), 256, C * sizeof(unsigned), stream>>>(labels, d, has_ignore, ignore, w.counts);
  int rc = check_launch("lv_label_hist_kernel");
  if (rc) return rc;
  lv_segments_kernel<<<1, 128, 0, stream>>>(w.counts, d.G, C, classes_mode, w.skip, w.npresent);
  if ((rc = check_launch("lv_segments_kernel"))) return rc;
  for (int e = 0; e < E; ++e) {
    if (dtype == EESEG_F32)
      rc = run_exit<float>((const float*)probas + (int64_t)e * exit_stride, labels, d, has_ignore,
                           ignore, w, per_exit + e,
                           dprobas ? (float*)dprobas + (int64_t)e * exit_stride : nullptr, stream);
    else
      rc = run_exit<__nv_bfloat16>((const __nv_bfloat16*)probas + (int64_t)e * exit_stride, labels, d,
                                   has_ignore, ignore, w, per_exit + e,
                                   dprobas ? (__nv_bfloat16*)dprobas + (int64_t)e * exit_stride : nullptr,
                                   stream);
    if (rc) return rc;
  }
  return EESEG_OK;
}
