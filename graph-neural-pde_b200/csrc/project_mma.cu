// Q/K projection on the tensor cores with the 3xTF32 split:  Q = X Wq^T + bq,  K = X Wk^T + bk.
// Replaces the two nn.Linear calls of reference src/function_transformer_attention.py:174-175.
//
// The projection is HBM-bound (reads X once, writes 2*ad floats per row) but the fp32 FFMA version (project.cu) is
// limited by shared-memory bandwidth and issue slots, not by HBM.  A single TF32 pass cannot meet the 1e-5 parity bar
// (10-bit mantissa), so every operand is split  a = a_hi + a_lo  (both TF32) and the product is taken as
//     a_lo*b_hi + a_hi*b_lo + a_hi*b_hi          (a_lo*b_lo ~ 2^-22 relative is dropped)
// with fp32 accumulation: ~2^-21 relative error per product, i.e. fp32-class accuracy, at 3 mma.sync per tile.
// mma.sync.m16n8k8.tf32 (legacy warp-level path; SASS HMMA) on purpose: the contraction is tiny (K = d <= 256,
// N = 2*ad <= 128) and the kernel stays HBM-bound -- tcgen05/TMEM would buy nothing here.
//
// One warp owns a 16-row tile of X staged in its private shared-memory slice; the split weights stay resident in shared
// memory for the lifetime of a persistent CTA.  Leading dimensions are = 4 mod 32 so that every fragment load is
// bank-conflict free.  Each warp double-buffers its tiles with cp.async (LDGSTS, zero-fill for the padding): the next
// tile's 8 KB are in flight while the current one is multiplied, which is what keeps enough bytes outstanding per SM for
// an HBM-bound stream (the synchronous version had ~1-2 loads per lane in flight and ran at 2.0 TB/s).
#include <stdlib.h>

#include "common.cuh"

namespace grand {

constexpr int kMmaThreads = 128;
constexpr int kMmaWarps = kMmaThreads / kWarp;
constexpr int kMmaRows = 16;      // rows per warp tile
constexpr int kMmaMaxNT = 16;     // n-tiles of 8 outputs: 2*ad <= 128

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return u;
}

// 16-byte LDGSTS; bytes beyond src_bytes (0 or 16 here) are written as zero
__device__ __forceinline__ void cp_async16_zfill(float* dst_smem, const float* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)),
               "l"(src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// Split weights  W = W_hi + W_lo  (both TF32) into shared memory, rows [Wq ; Wk] of `ld` floats, zero padded from d to
// dpad columns; loads are batched ahead of the conversions so that a CTA pays the L2 latency a few times, not once per
// element.
__device__ __forceinline__ void stage_split_weights(float* __restrict__ w_hi, float* __restrict__ w_lo, int ld, int d,
                                                    int d8, int out, int ad, const float* __restrict__ wq,
                                                    const float* __restrict__ wk, int tid, int nthreads) {
  if ((d & 3) == 0) {
    const int c4n = d >> 2, total = out * c4n;
    constexpr int kBatch = 8;
    for (int base = tid; base < total; base += nthreads * kBatch) {
      float4 v[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int idx = base + u * nthreads;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < total) {
          const int o = idx / c4n, c4 = idx - o * c4n;
          v[u] = ldg4(((o < ad) ? wq + (size_t)o * d : wk + (size_t)(o - ad) * d) + 4 * c4);
        }
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int idx = base + u * nthreads;
        if (idx < total) {
          const int o = idx / c4n, c4 = idx - o * c4n;
          const float vv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float hi = __uint_as_float(to_tf32(vv[e]));
            w_hi[o * ld + 4 * c4 + e] = hi;
            w_lo[o * ld + 4 * c4 + e] = __uint_as_float(to_tf32(vv[e] - hi));
          }
        }
      }
    }
    for (int idx = tid; idx < out * (d8 - d); idx += nthreads) {   // zero the pad columns
      const int o = idx / (d8 - d), c = d + idx - o * (d8 - d);
      w_hi[o * ld + c] = 0.f;
      w_lo[o * ld + c] = 0.f;
    }
  } else {
    for (int idx = tid; idx < out * d8; idx += nthreads) {
      const int o = idx / d8, c = idx - o * d8;
      float v = 0.f;
      if (c < d) v = (o < ad) ? __ldg(wq + (size_t)o * d + c) : __ldg(wk + (size_t)(o - ad) * d + c);
      const float hi = __uint_as_float(to_tf32(v));
      w_hi[o * ld + c] = hi;
      w_lo[o * ld + c] = __uint_as_float(to_tf32(v - hi));
    }
  }
}

template <int NT>
__global__ void __launch_bounds__(kMmaThreads)
project_qk_mma_kernel(const float* __restrict__ x, int n_rows, int n_cols, int d, int d8, int ld,
                      const float* __restrict__ wq, const float* __restrict__ bq, const float* __restrict__ wk,
                      const float* __restrict__ bk, int ad, int n_tiles, float* __restrict__ q,
                      float* __restrict__ k) {
  extern __shared__ __align__(16) float smem[];
  const int out = 2 * ad;                       // == NT * 8
  float* w_hi = smem;                           // [out][ld]
  float* w_lo = w_hi + (size_t)out * ld;        // [out][ld]
  float* xs_all = w_lo + (size_t)out * ld;      // [warps][nbuf][16][ld], nbuf = 2 when rows are 16-byte multiples
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = lane >> 2, tig = lane & 3;

  stage_split_weights(w_hi, w_lo, ld, d, d8, out, ad, wq, wk, tid, kMmaThreads);
  __syncthreads();

  const bool vec_ok = (d & 3) == 0;
  const int nbuf = vec_ok ? 2 : 1;
  float* xs_warp = xs_all + (size_t)warp * nbuf * kMmaRows * ld;
  const int c4n = d8 >> 2;
  // asynchronous staging of one 16-row tile (coalesced 16-byte pieces, zero filled beyond d columns / n_cols rows)
  auto stage_async = [&](int tile, float* dst) {
    const int row0 = tile * kMmaRows;
    for (int idx = lane; idx < kMmaRows * c4n; idx += kWarp) {
      const int r = idx / c4n, c4 = idx - r * c4n;
      const bool ok = row0 + r < n_cols && 4 * c4 < d;
      cp_async16_zfill(dst + r * ld + 4 * c4, ok ? x + (size_t)(row0 + r) * d + 4 * c4 : x, ok ? 16 : 0);
    }
  };
  const int tile_step = gridDim.x * kMmaWarps;
  int buf = 0;
  if (vec_ok) {
    const int first = blockIdx.x * kMmaWarps + warp;
    if (first < n_tiles) stage_async(first, xs_warp);
    cp_async_commit();
  }
  for (int tile = blockIdx.x * kMmaWarps + warp; tile < n_tiles; tile += tile_step) {
    const int row0 = tile * kMmaRows;
    float* xs = xs_warp + (size_t)buf * kMmaRows * ld;
    if (vec_ok) {
      // the other buffer was last read one iteration ago (warp-synchronous): refill it with the next tile, then wait
      // for this tile's group
      __syncwarp();
      if (tile + tile_step < n_tiles) stage_async(tile + tile_step, xs_warp + (size_t)(buf ^ 1) * kMmaRows * ld);
      cp_async_commit();
      cp_async_wait_group<1>();
      __syncwarp();
      buf ^= 1;
    } else {
      __syncwarp();
      for (int idx = lane; idx < kMmaRows * d8; idx += kWarp) {
        const int r = idx / d8, c = idx - r * d8;
        xs[r * ld + c] = (row0 + r < n_cols && c < d) ? x[(size_t)(row0 + r) * d + c] : 0.f;
      }
      __syncwarp();
    }

    float acc[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;

    for (int k0 = 0; k0 < d8; k0 += 8) {
      // A fragment (16 x 8, row major): a0 (grp, tig) a1 (grp+8, tig) a2 (grp, tig+4) a3 (grp+8, tig+4)
      const float f0 = xs[grp * ld + k0 + tig], f1 = xs[(grp + 8) * ld + k0 + tig];
      const float f2 = xs[grp * ld + k0 + tig + 4], f3 = xs[(grp + 8) * ld + k0 + tig + 4];
      const uint32_t h0 = to_tf32(f0), h1 = to_tf32(f1), h2 = to_tf32(f2), h3 = to_tf32(f3);
      const uint32_t l0 = to_tf32(f0 - __uint_as_float(h0)), l1 = to_tf32(f1 - __uint_as_float(h1));
      const uint32_t l2 = to_tf32(f2 - __uint_as_float(h2)), l3 = to_tf32(f3 - __uint_as_float(h3));
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        // B fragment (8 x 8, "col"): b0 (k = tig, n = grp) b1 (k = tig+4, n = grp), B[k][n] = W[n][k]
        const int wrow = (nt * 8 + grp) * ld + k0 + tig;
        const uint32_t bh0 = __float_as_uint(w_hi[wrow]), bh1 = __float_as_uint(w_hi[wrow + 4]);
        const uint32_t bl0 = __float_as_uint(w_lo[wrow]), bl1 = __float_as_uint(w_lo[wrow + 4]);
        mma_tf32(acc[nt], l0, l1, l2, l3, bh0, bh1);   // small terms first
        mma_tf32(acc[nt], h0, h1, h2, h3, bl0, bl1);
        mma_tf32(acc[nt], h0, h1, h2, h3, bh0, bh1);
      }
    }

    // C fragment: c0 (grp, 2 tig) c1 (grp, 2 tig + 1) c2 (grp+8, 2 tig) c3 (grp+8, 2 tig + 1)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int o = nt * 8 + 2 * tig;              // even; o and o+1 are on the same side of the Q|K boundary (ad even?)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int r = row0 + grp + half * 8;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int oo = o + e;
          const float v = acc[nt][half * 2 + e];
          if (oo < ad) {
            if (r < n_rows) q[(size_t)r * ad + oo] = v + __ldg(bq + oo);
          } else {
            if (r < n_cols) k[(size_t)r * ad + (oo - ad)] = v + __ldg(bk + oo - ad);
          }
        }
      }
    }
  }
}


// ---- rows that are 16-byte multiples: A fragments straight from global memory -------------------------------------------
// The contraction index of an mma can be permuted freely as long as A and B agree.  Thread (grp, tig) loads ONE float4
// of row grp (and one of row grp + 8) at columns 16 kb + 4 tig .. + 3 and feeds it to two k-steps:
//     step 1:  k = tig <-> column 4 tig,      k = tig + 4 <-> column 4 tig + 1
//     step 2:  k = tig <-> column 4 tig + 2,  k = tig + 4 <-> column 4 tig + 3
// so X never touches shared memory (a warp's load instruction covers 8 rows x 64 contiguous bytes, every sector fully
// used) and the matching B fragments of both steps are one LDS.128 per split half (ld = 16 mod 32 floats keeps each
// quarter-warp on 32 distinct banks).  Without the per-warp X tiles a CTA needs only the weights (37 KB for
// ogbn-arxiv), so 24 warps per SM are resident instead of 8-12 -- the staged kernel above was latency-bound at that
// occupancy (54 us per launch, 2.0 TB/s; double-buffering its tiles with cp.async changed nothing).
constexpr int kDirThreads = 256;
constexpr int kDirWarps = kDirThreads / kWarp;

__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
  hi = to_tf32(v);
  lo = to_tf32(v - __uint_as_float(hi));
}

template <int NT>
__global__ void __launch_bounds__(kDirThreads, NT <= 4 ? 3 : (NT <= 8 ? 2 : 1))
project_qk_mma_direct_kernel(const float* __restrict__ x, int n_rows, int n_cols, int d, int d16, int ld,
                             const float* __restrict__ wq, const float* __restrict__ bq,
                             const float* __restrict__ wk, const float* __restrict__ bk, int ad, int n_tiles,
                             float* __restrict__ q, float* __restrict__ k) {
  extern __shared__ __align__(16) float smem[];
  const int out = 2 * ad;                       // == NT * 8
  float* w_hi = smem;                           // [out][ld]
  float* w_lo = w_hi + (size_t)out * ld;        // [out][ld]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = lane >> 2, tig = lane & 3;
  stage_split_weights(w_hi, w_lo, ld, d, d16, out, ad, wq, wk, tid, kDirThreads);
  __syncthreads();

  const int n_kb = d16 >> 4;
  for (int tile = blockIdx.x * kDirWarps + warp; tile < n_tiles; tile += gridDim.x * kDirWarps) {
    const int row0 = tile * kMmaRows;
    const int ra = row0 + grp, rb = ra + 8;
    const bool oka = ra < n_cols, okb = rb < n_cols;
    const float* pa = x + (size_t)ra * d + 4 * tig;
    const float* pb = x + (size_t)rb * d + 4 * tig;

    float acc[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;

    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 a_nxt = (oka && 4 * tig < d) ? ld_stream4(pa) : zero4;
    float4 b_nxt = (okb && 4 * tig < d) ? ld_stream4(pb) : zero4;
    for (int kb = 0; kb < n_kb; ++kb) {
      const float4 a = a_nxt, b = b_nxt;
      const int cn = (kb + 1) * 16 + 4 * tig;          // next block's first column for this thread
      a_nxt = (oka && cn < d) ? ld_stream4(pa + (kb + 1) * 16) : zero4;
      b_nxt = (okb && cn < d) ? ld_stream4(pb + (kb + 1) * 16) : zero4;
      uint32_t ah[4], al[4], bh[4], bl[4];
      split_tf32(a.x, ah[0], al[0]); split_tf32(a.y, ah[1], al[1]);
      split_tf32(a.z, ah[2], al[2]); split_tf32(a.w, ah[3], al[3]);
      split_tf32(b.x, bh[0], bl[0]); split_tf32(b.y, bh[1], bl[1]);
      split_tf32(b.z, bh[2], bl[2]); split_tf32(b.w, bh[3], bl[3]);
      const int wcol = kb * 16 + 4 * tig;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const float4 wh = *reinterpret_cast<const float4*>(w_hi + (nt * 8 + grp) * ld + wcol);
        const float4 wl = *reinterpret_cast<const float4*>(w_lo + (nt * 8 + grp) * ld + wcol);
        const uint32_t h0 = __float_as_uint(wh.x), h1 = __float_as_uint(wh.y);
        const uint32_t h2 = __float_as_uint(wh.z), h3 = __float_as_uint(wh.w);
        const uint32_t l0 = __float_as_uint(wl.x), l1 = __float_as_uint(wl.y);
        const uint32_t l2 = __float_as_uint(wl.z), l3 = __float_as_uint(wl.w);
        // step 1 (columns 4 tig, 4 tig + 1): small terms first
        mma_tf32(acc[nt], al[0], bl[0], al[1], bl[1], h0, h1);
        mma_tf32(acc[nt], ah[0], bh[0], ah[1], bh[1], l0, l1);
        mma_tf32(acc[nt], ah[0], bh[0], ah[1], bh[1], h0, h1);
        // step 2 (columns 4 tig + 2, 4 tig + 3)
        mma_tf32(acc[nt], al[2], bl[2], al[3], bl[3], h2, h3);
        mma_tf32(acc[nt], ah[2], bh[2], ah[3], bh[3], l2, l3);
        mma_tf32(acc[nt], ah[2], bh[2], ah[3], bh[3], h2, h3);
      }
    }

    // C fragment: c0 (grp, 2 tig) c1 (grp, 2 tig + 1) c2 (grp+8, 2 tig) c3 (grp+8, 2 tig + 1): 8-byte stores
    // (2 tig is even and ad is a multiple of 4, so both outputs fall on the same side of the Q | K boundary)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int o = nt * 8 + 2 * tig;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int r = row0 + grp + half * 8;
        if (o < ad) {
          if (r < n_rows)
            *reinterpret_cast<float2*>(q + (size_t)r * ad + o) =
                make_float2(acc[nt][half * 2] + __ldg(bq + o), acc[nt][half * 2 + 1] + __ldg(bq + o + 1));
        } else {
          if (r < n_cols)
            *reinterpret_cast<float2*>(k + (size_t)r * ad + (o - ad)) =
                make_float2(acc[nt][half * 2] + __ldg(bk + o - ad), acc[nt][half * 2 + 1] + __ldg(bk + o - ad + 1));
        }
      }
    }
  }
}

// returns GRAND_EUNSUPPORTED (without setting an error) when the shape is outside what this kernel is built for; the
// caller then uses the FFMA kernel
int launch_project_qk_mma(const float* x, int n_rows, int n_cols, int d, const float* wq, const float* bq,
                          const float* wk, const float* bk, int ad, float* q, float* k, cudaStream_t stream) {
  const int out = 2 * ad;
  if ((out & 7) != 0 || out > 8 * kMmaMaxNT || d > 512) return GRAND_EUNSUPPORTED;
  if ((d & 3) == 0 && (ad & 3) == 0 && getenv("GRAND_PROJ_STAGED") == nullptr) {
    const int d16 = (d + 15) & ~15;
    const int ldd = d16 + ((16 - d16 % 32) + 32) % 32;            // smallest ld >= d16 with ld = 16 mod 32
    const size_t smem_d = (size_t)2 * out * ldd * sizeof(float);
    const int nt_d = out / 8;
    const int want = nt_d <= 4 ? 3 : (nt_d <= 8 ? 2 : 1);
    if (smem_d <= 200 * 1024 && (nt_d == 1 || nt_d == 2 || nt_d == 4 || nt_d == 8 || nt_d == 16)) {
      int per_sm = (int)((220 * 1024) / (smem_d + 1024));
      if (per_sm > want) per_sm = want;
      if (per_sm < 1) per_sm = 1;
      const int n_tiles_d = (n_cols + kMmaRows - 1) / kMmaRows;
      int grid_d = num_sms() * per_sm;
      const int max_grid_d = (n_tiles_d + kDirWarps - 1) / kDirWarps;
      if (grid_d > max_grid_d) grid_d = max_grid_d;
#define GRAND_LAUNCH_DIRECT(NTV)                                                                                    \
  {                                                                                                                 \
    auto kern = project_qk_mma_direct_kernel<NTV>;                                                                  \
    /* per device and cheap: set on every launch (no process-wide "already done" flag) */                           \
    GRAND_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_d));            \
    kern<<<grid_d, kDirThreads, smem_d, stream>>>(x, n_rows, n_cols, d, d16, ldd, wq, bq, wk, bk, ad, n_tiles_d, q, \
                                                  k);                                                               \
  }
      switch (nt_d) {
        case 1: GRAND_LAUNCH_DIRECT(1) break;
        case 2: GRAND_LAUNCH_DIRECT(2) break;
        case 4: GRAND_LAUNCH_DIRECT(4) break;
        case 8: GRAND_LAUNCH_DIRECT(8) break;
        default: GRAND_LAUNCH_DIRECT(16) break;
      }
#undef GRAND_LAUNCH_DIRECT
      GRAND_CUDA_OK(cudaGetLastError());
      count_launch();
      return GRAND_OK;
    }
  }
  const int d8 = (d + 7) & ~7;
  const int ld = d8 + 4;                                   // = 4 mod 8 ... and = 4 mod 32 when d8 % 32 == 0
  const int nbuf = (d & 3) == 0 ? 2 : 1;
  const size_t smem = ((size_t)2 * out + (size_t)kMmaWarps * nbuf * kMmaRows) * ld * sizeof(float);
  if (smem > 200 * 1024) return GRAND_EUNSUPPORTED;
  const int n_tiles = (n_cols + kMmaRows - 1) / kMmaRows;
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  int grid = num_sms() * per_sm;
  const int max_grid = (n_tiles + kMmaWarps - 1) / kMmaWarps;
  if (grid > max_grid) grid = max_grid;
  const int nt = out / 8;

#define GRAND_LAUNCH_MMA(NTV)                                                                                       \
  {                                                                                                                 \
    auto kern = project_qk_mma_kernel<NTV>;                                                                         \
    GRAND_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));              \
    kern<<<grid, kMmaThreads, smem, stream>>>(x, n_rows, n_cols, d, d8, ld, wq, bq, wk, bk, ad, n_tiles, q, k);     \
  }
  switch (nt) {
    case 1: GRAND_LAUNCH_MMA(1) break;
    case 2: GRAND_LAUNCH_MMA(2) break;
    case 4: GRAND_LAUNCH_MMA(4) break;
    case 8: GRAND_LAUNCH_MMA(8) break;
    case 16: GRAND_LAUNCH_MMA(16) break;
    default: return GRAND_EUNSUPPORTED;
  }
#undef GRAND_LAUNCH_MMA
  GRAND_CUDA_OK(cudaGetLastError());
  count_launch();
  return GRAND_OK;
}

}  // namespace grand
