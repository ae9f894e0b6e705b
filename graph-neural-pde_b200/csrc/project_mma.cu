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
// bank-conflict free.
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

__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
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
  float* xs_all = w_lo + (size_t)out * ld;      // [warps][16][ld]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = lane >> 2, tig = lane & 3;

  // split weights, zero padded to d8; loads are batched ahead of the conversions so that a CTA pays the L2 latency a
  // few times, not once per element
  if ((d & 3) == 0) {
    const int c4n = d >> 2, total = out * c4n;
    constexpr int kBatch = 8;
    for (int base = tid; base < total; base += kMmaThreads * kBatch) {
      float4 v[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int idx = base + u * kMmaThreads;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < total) {
          const int o = idx / c4n, c4 = idx - o * c4n;
          v[u] = ldg4(((o < ad) ? wq + (size_t)o * d : wk + (size_t)(o - ad) * d) + 4 * c4);
        }
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int idx = base + u * kMmaThreads;
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
    for (int idx = tid; idx < out * (d8 - d); idx += kMmaThreads) {   // zero the pad columns
      const int o = idx / (d8 - d), c = d + idx - o * (d8 - d);
      w_hi[o * ld + c] = 0.f;
      w_lo[o * ld + c] = 0.f;
    }
  } else {
    for (int idx = tid; idx < out * d8; idx += kMmaThreads) {
      const int o = idx / d8, c = idx - o * d8;
      float v = 0.f;
      if (c < d) v = (o < ad) ? __ldg(wq + (size_t)o * d + c) : __ldg(wk + (size_t)(o - ad) * d + c);
      const float hi = __uint_as_float(to_tf32(v));
      w_hi[o * ld + c] = hi;
      w_lo[o * ld + c] = __uint_as_float(to_tf32(v - hi));
    }
  }
  __syncthreads();

  float* xs = xs_all + (size_t)warp * kMmaRows * ld;
  const bool vec_ok = (d & 3) == 0;
  for (int tile = blockIdx.x * kMmaWarps + warp; tile < n_tiles; tile += gridDim.x * kMmaWarps) {
    const int row0 = tile * kMmaRows;
    __syncwarp();
    // stage 16 rows of X (coalesced), zero padded to d8 columns / n_cols rows
    if (vec_ok) {
      const int c4n = d8 >> 2;
      for (int idx = lane; idx < kMmaRows * c4n; idx += kWarp) {
        const int r = idx / c4n, c4 = idx - r * c4n;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < n_cols && 4 * c4 < d) v = ld_stream4(x + (size_t)(row0 + r) * d + 4 * c4);
        *reinterpret_cast<float4*>(xs + r * ld + 4 * c4) = v;
      }
    } else {
      for (int idx = lane; idx < kMmaRows * d8; idx += kWarp) {
        const int r = idx / d8, c = idx - r * d8;
        xs[r * ld + c] = (row0 + r < n_cols && c < d) ? x[(size_t)(row0 + r) * d + c] : 0.f;
      }
    }
    __syncwarp();

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

// returns GRAND_EUNSUPPORTED (without setting an error) when the shape is outside what this kernel is built for; the
// caller then uses the FFMA kernel
int launch_project_qk_mma(const float* x, int n_rows, int n_cols, int d, const float* wq, const float* bq,
                          const float* wk, const float* bk, int ad, float* q, float* k, cudaStream_t stream) {
  const int out = 2 * ad;
  if ((out & 7) != 0 || out > 8 * kMmaMaxNT || d > 512) return GRAND_EUNSUPPORTED;
  const int d8 = (d + 7) & ~7;
  const int ld = d8 + 4;                                   // = 4 mod 8 ... and = 4 mod 32 when d8 % 32 == 0
  const size_t smem = ((size_t)2 * out + (size_t)kMmaWarps * kMmaRows) * ld * sizeof(float);
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
    static size_t configured = 0;                                                                                   \
    if (smem > configured) {                                                                                        \
      GRAND_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));            \
      configured = smem;                                                                                            \
    }                                                                                                               \
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
