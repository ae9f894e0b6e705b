// The edge phase of the GRAND right-hand side.
//
//   rhs_window_kernel<MODE, VEC, R, STAGE>   one warp per WINDOW of <= GRAND_TILE (32) CSR edges, one edge per lane.
//     A window is either several whole query rows packed together (the common case: the average in-degree of the
//     reference's graphs is 5-30) or one 32-edge chunk of a long row.  Per window:
//       1. every lane reads its edge (row i, col j) and the neighbour rows X[col_e] of ALL edges of the window start
//          moving into shared memory at once -- STAGE_BULK: one cp.async.bulk (TMA bulk copy, completion on an
//          mbarrier) per lane; STAGE_CPASYNC: LDGSTS; STAGE_DIRECT: no staging, batched register loads;
//       2. MODE_FUSED: scores q_i.k_j / sqrt(d_k) lane-parallel, softmax over each row's lanes with segmented
//          shuffle reductions, head mean (reference src/function_transformer_attention.py:190-213, :34) -- nothing
//          edge-sized is ever written (the reference materialises [E,d_k,H] x2, ~10 [E,H] tensors and [E,d]);
//          MODE_PRECOMP: scores / segment statistics from the general pre-pass (attention_norm_idx = 1, squareplus,
//          non-default scorers, > 32 heads...); MODE_WEIGHT: GRAND-l, weights given (function_laplacian_diffusion.py:28-36);
//       3. per row: a_e-weighted sum of the staged neighbour rows (torch_sparse.spmm, :35), then the residual
//          alpha (ax - x) + beta x0 and the Runge-Kutta stage combine as the epilogue.
//     Chunks of long rows publish partial sums; the last chunk to finish adds them in slot order (deterministic) and
//     runs the epilogue.
//
//   split_tile_stats_kernel   (max, sum-exp) per head for every chunk of a long row, merged by the fused kernel.
//   edge_scores_kernel        prods[E,H] in CSR order (+ global max for squareplus)         reference :190-209
//   segment_stats_kernel      per segment (row or column) max / normaliser sum               reference :210-213
//   attention_out_kernel      attention[E,H] / prods[E,H] back in the caller's edge order, head-mean in CSR order
//
// HBM/L2-bound gather work; the design goal is bytes in flight per SM (all neighbour rows of a window are requested
// before any is consumed) with a short instruction stream per edge.
#include <stdlib.h>
#include <string.h>

#include "attention.cuh"

namespace grand {

constexpr int MODE_FUSED = 0;
constexpr int MODE_PRECOMP = 1;
constexpr int MODE_WEIGHT = 2;

constexpr int STAGE_DIRECT = 0;   // neighbour rows straight into registers, kUnroll at a time
constexpr int STAGE_CPASYNC = 1;  // cp.async (LDGSTS) into the warp's shared-memory window buffer
constexpr int STAGE_BULK = 2;     // cp.async.bulk (TMA bulk copy) + mbarrier into the window buffer

constexpr int kUnroll = 8;        // STAGE_DIRECT: neighbour rows in flight per warp
constexpr int kRhsThreads = 128;  // 4 warps per block
constexpr int kRhsWarps = kRhsThreads / kWarp;

struct GraphView {
  int n_rows, n_cols, n_edges, n_tiles, n_split_rows, n_split_tiles;
  const int *rowptr, *col, *erow, *eid, *tile_row, *tile_beg, *tile_nrows, *tile_split, *split_tile0, *colptr, *csc_pos;
};

struct EpiView {
  const float* alpha_train;
  const float* beta_train;
  const float* x0;
  int no_alpha_sigmoid;
  int n_terms;
  float* f_out;
  float* y_out;
  float coef_f;
  float coef[GRAND_MAX_TERMS];
  const float* term[GRAND_MAX_TERMS];
};

struct RhsArgs {
  GraphView g;
  AttnView a;          // MODE_FUSED / MODE_PRECOMP
  EpiView ep;
  int d;
  const float* x;      // [n_cols, d]
  const float* q;      // [n_rows, ad_eff]
  const float* k;      // [n_cols, ad_eff]
  const float* w_csr;  // MODE_WEIGHT
  const float* scores;       // MODE_PRECOMP [E, H]
  const float* seg_stats;    // MODE_PRECOMP [n_seg, H, 2]
  const float* gmax;         // squareplus global max
  const float* split_stats;  // [n_split_tiles, H, 2]
  float* split_partial;      // [n_split_tiles, d]
  int* split_counter;        // [n_split_rows]
};

// ----------------------------------------------------------------------------------------------------------
// small PTX wrappers: shared-memory async copies and mbarrier
// ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const float* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}

// segmented all-reduce over contiguous lane segments [lo, hi): every lane gets the reduction of its own segment
__device__ __forceinline__ float seg_max(float v, int lane, int lo, int hi) {
#pragma unroll
  for (int o = 1; o < kWarp; o <<= 1) {
    const float t = __shfl_down_sync(kFull, v, o);
    if (lane + o < hi) v = fmaxf(v, t);
  }
  return __shfl_sync(kFull, v, lo);
}
__device__ __forceinline__ float seg_sum(float v, int lane, int lo, int hi) {
#pragma unroll
  for (int o = 1; o < kWarp; o <<= 1) {
    const float t = __shfl_down_sync(kFull, v, o);
    if (lane + o < hi) v += t;
  }
  return __shfl_sync(kFull, v, lo);
}

// one lane-per-edge score of the fused path (reference's default scorer scaled_dot, :195-196)
__device__ __forceinline__ float fused_score(const AttnView& a, const float* __restrict__ qi,
                                             const float* __restrict__ kj, int h, bool vec4) {
  const float s = vec4 ? head_dot<true>(qi + h * a.dk, kj + h * a.dk, a.dk)
                       : head_dot<false>(qi + h * a.dk, kj + h * a.dk, a.dk);
  return s / a.sqrt_dk;
}

__global__ void __launch_bounds__(256)
split_tile_stats_kernel(GraphView g, AttnView a, const float* __restrict__ q, const float* __restrict__ k,
                        float* __restrict__ split_stats) {
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // split tiles are tiles [0, n_split_tiles)
  if (t >= g.n_split_tiles) return;
  const int row = __ldg(g.tile_row + t);
  const int beg = __ldg(g.tile_beg + t);
  const int end = min(beg + GRAND_TILE, __ldg(g.rowptr + row + 1));
  const int p = beg + lane;
  const bool valid = p < end;
  const float* kj = k + (size_t)(valid ? __ldg(g.col + p) : 0) * a.ad_eff;
  const float* qi = q + (size_t)row * a.ad_eff;
  const bool vec4 = ((a.dk & 3) == 0) && ((a.ad_eff & 3) == 0);
  const float ew = (valid && a.edge_weight_csr != nullptr) ? __ldg(a.edge_weight_csr + p) : 1.f;
  for (int h = 0; h < a.heads; ++h) {
    float s = -INFINITY;
    if (valid) s = fused_score(a, qi, kj, h, vec4) * ew;
    const float m = warp_max(s);
    const float l = warp_sum(valid ? expf(s - m) : 0.f);
    if (lane == 0) {
      split_stats[((size_t)t * a.heads + h) * 2 + 0] = m;
      split_stats[((size_t)t * a.heads + h) * 2 + 1] = l;
    }
  }
}

// ----------------------------------------------------------------------------------------------------------
// epilogue: residual + RK stage combine for one finished row
// ----------------------------------------------------------------------------------------------------------
template <int VEC, int R>
__device__ __forceinline__ void finish_row(const RhsArgs& A, int row, int lane, float alpha, float beta,
                                           float (&acc)[R][VEC]) {
  const int d = A.d;
  const size_t base = (size_t)row * d;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int c = (lane + r * kWarp) * VEC;
    if (c >= d) continue;
    float xi[VEC], f[VEC];
    if (VEC == 4) {
      const float4 v = ldg4(A.x + base + c);
      xi[0] = v.x; xi[1 % VEC] = v.y; xi[2 % VEC] = v.z; xi[3 % VEC] = v.w;
    } else {
      xi[0] = __ldg(A.x + base + c);
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) f[v] = alpha * (acc[r][v] - xi[v]);   // f = alpha (ax - x)
    if (A.ep.x0 != nullptr) {                                           // + beta x0   (add_source)
      if (VEC == 4) {
        const float4 s = ld_stream4(A.ep.x0 + base + c);
        f[0] += beta * s.x; f[1 % VEC] += beta * s.y; f[2 % VEC] += beta * s.z; f[3 % VEC] += beta * s.w;
      } else {
        f[0] += beta * __ldg(A.ep.x0 + base + c);
      }
    }
    if (A.ep.f_out != nullptr) {
      if (VEC == 4) st_stream4(A.ep.f_out + base + c, make_float4(f[0], f[1 % VEC], f[2 % VEC], f[3 % VEC]));
      else A.ep.f_out[base + c] = f[0];
    }
    if (A.ep.y_out != nullptr) {
      float y[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) y[v] = A.ep.coef_f * f[v];
#pragma unroll
      for (int t = 0; t < GRAND_MAX_TERMS; ++t) {
        if (t < A.ep.n_terms) {
          const float ct = A.ep.coef[t];
          if (VEC == 4) {
            const float4 s = ld_stream4(A.ep.term[t] + base + c);
            y[0] = fmaf(ct, s.x, y[0]); y[1 % VEC] = fmaf(ct, s.y, y[1 % VEC]);
            y[2 % VEC] = fmaf(ct, s.z, y[2 % VEC]); y[3 % VEC] = fmaf(ct, s.w, y[3 % VEC]);
          } else {
            y[0] = fmaf(ct, __ldg(A.ep.term[t] + base + c), y[0]);
          }
        }
      }
      if (VEC == 4) st_stream4(A.ep.y_out + base + c, make_float4(y[0], y[1 % VEC], y[2 % VEC], y[3 % VEC]));
      else A.ep.y_out[base + c] = y[0];
    }
  }
}

// ----------------------------------------------------------------------------------------------------------
// the RHS kernel
// ----------------------------------------------------------------------------------------------------------
template <int MODE, int VEC, int R, int STAGE>
__global__ void __launch_bounds__(kRhsThreads, STAGE == STAGE_DIRECT ? 4 : 3)
rhs_window_kernel(const __grid_constant__ RhsArgs A) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int t = blockIdx.x * kRhsWarps + warp;
  if (t >= A.g.n_tiles) return;
  const int d = A.d;
  const int row_bytes = d * (int)sizeof(float);
  const int row0 = __ldg(A.g.tile_row + t);
  const int beg = __ldg(A.g.tile_beg + t);
  const int nrows = __ldg(A.g.tile_nrows + t);
  const int split = __ldg(A.g.tile_split + t);
  // rowptr slice of this window: lane l holds rowptr[row0 + min(l, nrows)]
  const int rp = __ldg(A.g.rowptr + row0 + min(lane, nrows));
  const int end = split < 0 ? __shfl_sync(kFull, rp, nrows) : min(beg + GRAND_TILE, __shfl_sync(kFull, rp, 1));
  const int cnt = end - beg;          // <= GRAND_TILE
  const int p = beg + lane;
  const bool valid = lane < cnt;
  const int j = valid ? __ldg(A.g.col + p) : 0;

  // ---- 1. start moving every neighbour row of the window into shared memory ------------------------------
  float* stage = nullptr;
  uint32_t bar = 0;
  if (STAGE != STAGE_DIRECT) {
    const int warp_bytes = GRAND_TILE * row_bytes + 16;
    unsigned char* base = smem_raw + (size_t)warp * warp_bytes;
    stage = reinterpret_cast<float*>(base);
    if (STAGE == STAGE_BULK) {
      bar = smem_u32(base + GRAND_TILE * row_bytes);
      if (lane == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, (uint32_t)(cnt * row_bytes));
      }
      __syncwarp();
      if (valid) bulk_copy_g2s(smem_u32(stage) + lane * row_bytes, A.x + (size_t)j * d, (uint32_t)row_bytes, bar);
    } else {
      for (int e = 0; e < cnt; ++e) {
        const int je = __shfl_sync(kFull, j, e);
        const float* src = A.x + (size_t)je * d;
        const uint32_t dst = smem_u32(stage) + e * row_bytes;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int c = (lane + r * kWarp) * 4;
          if (c < d) cp_async16(dst + c * 4, src + c);
        }
      }
    }
  }

  // ---- 2. per-edge weight ab = head-mean attention (or the given weight) ----------------------------------
  // own row and its lane segment [lo, hi) inside the window
  int i = row0, lo = 0, hi = cnt;
  if (split < 0) {
    i = valid ? __ldg(A.g.erow + p) : row0;
    const int kk = i - row0;
    lo = __shfl_sync(kFull, rp, kk) - beg;
    hi = __shfl_sync(kFull, rp, kk + 1) - beg;
  }
  float ab = 0.f;
  if (MODE == MODE_WEIGHT) {
    if (valid) ab = __ldg(A.w_csr + p);
  } else if (MODE == MODE_PRECOMP) {
    const int H = A.a.heads;
    if (valid) {
      const float gmax = A.a.norm_type == GRAND_NORM_SQUAREPLUS ? __ldg(A.gmax) : 0.f;
      const int seg = A.a.norm_idx == 0 ? i : j;
      float a_sum = 0.f;
      for (int h = 0; h < H; ++h) {
        const float sc = __ldg(A.scores + (size_t)p * H + h);
        const float m = __ldg(A.seg_stats + ((size_t)seg * H + h) * 2 + 0);
        const float l = __ldg(A.seg_stats + ((size_t)seg * H + h) * 2 + 1);
        a_sum += norm_num(A.a.norm_type, sc, A.a.norm_type == GRAND_NORM_SOFTMAX ? m : gmax) / (l + 1e-16f);
      }
      ab = a_sum / (float)H;            // attention.mean(dim=1), reference :34
    }
  } else {  // MODE_FUSED: scaled_dot scores, softmax over the row (attention_norm_idx == 0)
    const bool vec4 = ((A.a.dk & 3) == 0) && ((A.a.ad_eff & 3) == 0);
    const float* qi = A.q + (size_t)i * A.a.ad_eff;
    const float* kj = A.k + (size_t)j * A.a.ad_eff;
    const float ew = (valid && A.a.edge_weight_csr != nullptr) ? __ldg(A.a.edge_weight_csr + p) : 1.f;
    int slot0 = 0, slot1 = 0;
    if (split >= 0) {
      slot0 = __ldg(A.g.split_tile0 + split);
      slot1 = __ldg(A.g.split_tile0 + split + 1);
    }
    for (int h = 0; h < A.a.heads; ++h) {
      float s = -INFINITY;
      if (valid) s = fused_score(A.a, qi, kj, h, vec4) * ew;     // (* ew: reweight_attention, reference :208-209)
      float m, l;
      if (split < 0) {
        m = seg_max(s, lane, lo, hi);
        l = seg_sum(valid ? expf(s - m) : 0.f, lane, lo, hi);
      } else {
        // merge the per-chunk (max, sum-exp) pairs of this row: l = sum_t l_t exp(m_t - m)
        m = -INFINITY;
        for (int u = slot0 + lane; u < slot1; u += kWarp)
          m = fmaxf(m, __ldg(A.split_stats + ((size_t)u * A.a.heads + h) * 2));
        m = warp_max(m);
        l = 0.f;
        for (int u = slot0 + lane; u < slot1; u += kWarp) {
          const float mt = __ldg(A.split_stats + ((size_t)u * A.a.heads + h) * 2);
          const float lt = __ldg(A.split_stats + ((size_t)u * A.a.heads + h) * 2 + 1);
          l += lt * expf(mt - m);
        }
        l = warp_sum(l);
      }
      if (valid) ab += expf(s - m) / (l + 1e-16f);
    }
    ab = ab / (float)A.a.heads;
  }

  // ---- 3. aggregation per row + epilogue -------------------------------------------------------------------
  float alpha = __ldg(A.ep.alpha_train);
  if (!A.ep.no_alpha_sigmoid) alpha = 1.f / (1.f + expf(-alpha));   // torch.sigmoid
  const float beta = A.ep.x0 != nullptr ? __ldg(A.ep.beta_train) : 0.f;

  if (STAGE == STAGE_BULK) mbar_wait(bar, 0);
  if (STAGE == STAGE_CPASYNC) {
    cp_async_wait_all();
    __syncwarp();
  }

  for (int kr = 0; kr < nrows; ++kr) {
    int e_lo = 0, e_hi = cnt;
    if (split < 0) {
      e_lo = __shfl_sync(kFull, rp, kr) - beg;
      e_hi = __shfl_sync(kFull, rp, kr + 1) - beg;
    }
    float acc[R][VEC];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[r][v] = 0.f;

    if (STAGE != STAGE_DIRECT) {
      for (int e = e_lo; e < e_hi; ++e) {
        const float a = __shfl_sync(kFull, ab, e);
        const float* srow = stage + (size_t)e * d;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int c = (lane + r * kWarp) * 4;
          if (c < d) {
            const float4 v = *reinterpret_cast<const float4*>(srow + c);
            acc[r][0] = fmaf(a, v.x, acc[r][0]); acc[r][1 % VEC] = fmaf(a, v.y, acc[r][1 % VEC]);
            acc[r][2 % VEC] = fmaf(a, v.z, acc[r][2 % VEC]); acc[r][3 % VEC] = fmaf(a, v.w, acc[r][3 % VEC]);
          }
        }
      }
    } else {
      constexpr int U = (kUnroll / R) < 2 ? 2 : (kUnroll / R);
      for (int e0 = e_lo; e0 < e_hi; e0 += U) {
        int jj[U];
        float aa[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int src = min(e0 + u, e_hi - 1);
          jj[u] = __shfl_sync(kFull, j, src);
          aa[u] = __shfl_sync(kFull, ab, src);
          if (e0 + u >= e_hi) aa[u] = 0.f;
        }
        float xv[U][R][VEC];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const float* xr = A.x + (size_t)jj[u] * d;
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const int c = (lane + r * kWarp) * VEC;
            if (VEC == 4) {
              float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
              if (c < d) v = ldg4(xr + c);
              xv[u][r][0] = v.x; xv[u][r][1 % VEC] = v.y; xv[u][r][2 % VEC] = v.z; xv[u][r][3 % VEC] = v.w;
            } else {
              xv[u][r][0] = c < d ? __ldg(xr + c) : 0.f;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int r = 0; r < R; ++r)
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[r][v] = fmaf(aa[u], xv[u][r][v], acc[r][v]);
      }
    }

    if (split < 0) {
      finish_row<VEC, R>(A, row0 + kr, lane, alpha, beta, acc);
      continue;
    }

    // ---- chunk of a long row: publish the partial sum, the last chunk of the row reduces and finishes --------
    float* part = A.split_partial + (size_t)t * d;   // split tiles are tiles [0, n_split_tiles): slot == t
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int c = (lane + r * kWarp) * VEC;
      if (c >= d) continue;
      if (VEC == 4) __stcg(reinterpret_cast<float4*>(part + c),
                           make_float4(acc[r][0], acc[r][1 % VEC], acc[r][2 % VEC], acc[r][3 % VEC]));
      else __stcg(part + c, acc[r][0]);
    }
    __threadfence();
    __syncwarp();
    const int slot0 = __ldg(A.g.split_tile0 + split);
    const int slot1 = __ldg(A.g.split_tile0 + split + 1);
    int old = 0;
    if (lane == 0) old = atomicAdd(A.split_counter + split, 1);
    old = __shfl_sync(kFull, old, 0);
    if (old != slot1 - slot0 - 1) return;
    __threadfence();
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[r][v] = 0.f;
    for (int u = slot0; u < slot1; ++u) {
      const float* pu = A.split_partial + (size_t)u * d;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int c = (lane + r * kWarp) * VEC;
        if (c >= d) continue;
        if (VEC == 4) {
          const float4 v = __ldcg(reinterpret_cast<const float4*>(pu + c));
          acc[r][0] += v.x; acc[r][1 % VEC] += v.y; acc[r][2 % VEC] += v.z; acc[r][3 % VEC] += v.w;
        } else {
          acc[r][0] += __ldcg(pu + c);
        }
      }
    }
    if (lane == 0) A.split_counter[split] = 0;   // leave the counter ready for the next launch
    finish_row<VEC, R>(A, row0, lane, alpha, beta, acc);
  }
}

// ----------------------------------------------------------------------------------------------------------
// general-path pre-passes
// ----------------------------------------------------------------------------------------------------------
__global__ void init_gmax_kernel(float* gmax) { *reinterpret_cast<int*>(gmax) = float_to_ordered(-INFINITY); }
__global__ void decode_gmax_kernel(float* gmax) { *gmax = ordered_to_float(*reinterpret_cast<int*>(gmax)); }

template <bool kVec4>
__global__ void __launch_bounds__(256)
edge_scores_kernel(GraphView g, AttnView a, const float* __restrict__ q, const float* __restrict__ k,
                   float* __restrict__ scores, float* gmax) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  float local_max = -INFINITY;
  if (p < g.n_edges) {
    const float* qi = q + (size_t)__ldg(g.erow + p) * a.ad_eff;
    const float* kj = k + (size_t)__ldg(g.col + p) * a.ad_eff;
    const float ew = a.edge_weight_csr != nullptr ? __ldg(a.edge_weight_csr + p) : 1.f;
    for (int h = 0; h < a.heads; ++h) {
      float v = edge_score<kVec4>(a, qi, kj, h);
      if (a.edge_weight_csr != nullptr) v *= ew;
      scores[(size_t)p * a.heads + h] = v;
      local_max = fmaxf(local_max, v);
    }
  }
  if (gmax != nullptr) {   // squareplus: src.max() over the whole [E,H] tensor, reference src/utils.py:196
    local_max = warp_max(local_max);
    if ((threadIdx.x & 31) == 0 && local_max != -INFINITY)
      atomicMax(reinterpret_cast<int*>(gmax), float_to_ordered(local_max));
  }
}

// one warp per segment; segments are CSR rows (norm_idx 0) or CSC columns (norm_idx 1)
__global__ void __launch_bounds__(256)
segment_stats_kernel(GraphView g, AttnView a, const float* __restrict__ scores, const float* __restrict__ gmax,
                     float* __restrict__ seg_stats) {
  const int lane = threadIdx.x & 31;
  const int seg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int n_seg = a.norm_idx == 0 ? g.n_rows : g.n_cols;
  if (seg >= n_seg) return;
  const int* ptr = a.norm_idx == 0 ? g.rowptr : g.colptr;
  const int beg = __ldg(ptr + seg), end = __ldg(ptr + seg + 1);
  const int H = a.heads;
  const float gm = a.norm_type == GRAND_NORM_SQUAREPLUS ? *gmax : 0.f;
  for (int h = 0; h < H; ++h) {
    float m = -INFINITY;
    if (a.norm_type == GRAND_NORM_SOFTMAX) {
      for (int c = beg + lane; c < end; c += kWarp) {
        const int p = a.norm_idx == 0 ? c : __ldg(g.csc_pos + c);
        m = fmaxf(m, __ldg(scores + (size_t)p * H + h));
      }
      m = warp_max(m);
    } else {
      m = gm;
    }
    float l = 0.f;
    for (int c = beg + lane; c < end; c += kWarp) {
      const int p = a.norm_idx == 0 ? c : __ldg(g.csc_pos + c);
      l += norm_num(a.norm_type, __ldg(scores + (size_t)p * H + h), m);
    }
    l = warp_sum(l);
    if (lane == 0) {
      seg_stats[((size_t)seg * H + h) * 2 + 0] = m;
      seg_stats[((size_t)seg * H + h) * 2 + 1] = l;
    }
  }
}

__global__ void __launch_bounds__(256)
attention_out_kernel(GraphView g, AttnView a, const float* __restrict__ scores, const float* __restrict__ seg_stats,
                     const float* __restrict__ gmax, float* __restrict__ attention_out,
                     float* __restrict__ prods_out, float* __restrict__ mean_csr_out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= g.n_edges) return;
  const int H = a.heads;
  const int seg = a.norm_idx == 0 ? __ldg(g.erow + p) : __ldg(g.col + p);
  const int e = __ldg(g.eid + p);
  const float gm = a.norm_type == GRAND_NORM_SQUAREPLUS ? *gmax : 0.f;
  float a_sum = 0.f;
  for (int h = 0; h < H; ++h) {
    const float sc = __ldg(scores + (size_t)p * H + h);
    const float m = a.norm_type == GRAND_NORM_SOFTMAX ? __ldg(seg_stats + ((size_t)seg * H + h) * 2) : gm;
    const float l = __ldg(seg_stats + ((size_t)seg * H + h) * 2 + 1);
    const float att = norm_num(a.norm_type, sc, m) / (l + 1e-16f);
    a_sum += att;
    if (attention_out != nullptr) attention_out[(size_t)e * H + h] = att;
    if (prods_out != nullptr) prods_out[(size_t)e * H + h] = sc;
  }
  if (mean_csr_out != nullptr) mean_csr_out[p] = a_sum / (float)H;
}

__global__ void __launch_bounds__(256)
edge_values_to_csr_kernel(GraphView g, const float* __restrict__ w, int heads, float* __restrict__ w_csr) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= g.n_edges) return;
  const int e = __ldg(g.eid + p);
  if (heads == 1) {
    w_csr[p] = __ldg(w + e);
  } else {
    float s = 0.f;
    for (int h = 0; h < heads; ++h) s += __ldg(w + (size_t)e * heads + h);
    w_csr[p] = s / (float)heads;     // attention_weights.mean(dim=1), reference function_laplacian_diffusion.py:30
  }
}

// ----------------------------------------------------------------------------------------------------------
// host-side launchers
// ----------------------------------------------------------------------------------------------------------
static GraphView view(const grand_graph* g) {
  GraphView v;
  v.n_rows = g->n_rows; v.n_cols = g->n_cols; v.n_edges = g->n_edges; v.n_tiles = g->n_tiles;
  v.n_split_rows = g->n_split_rows; v.n_split_tiles = g->n_split_tiles;
  v.rowptr = g->rowptr; v.col = g->col; v.erow = g->erow; v.eid = g->eid; v.tile_row = g->tile_row;
  v.tile_beg = g->tile_beg; v.tile_nrows = g->tile_nrows; v.tile_split = g->tile_split;
  v.split_tile0 = g->split_tile0; v.colptr = g->colptr; v.csc_pos = g->csc_pos;
  return v;
}

int make_attn_view(const grand_attention* a, AttnView* out) {
  GRAND_REQUIRE(a->heads > 0 && a->ad > 0 && a->ad % a->heads == 0, GRAND_EINVAL,
                "Number of heads (%d) must be a factor of the dimension size (%d)", a->heads, a->ad);
  GRAND_REQUIRE(a->score_type >= 0 && a->score_type <= GRAND_SCORE_EXP_KERNEL2, GRAND_EINVAL,
                "unknown score_type %d", a->score_type);
  GRAND_REQUIRE(a->norm_type == GRAND_NORM_SOFTMAX || a->norm_type == GRAND_NORM_SQUAREPLUS, GRAND_EINVAL,
                "unknown norm_type %d", a->norm_type);
  GRAND_REQUIRE(a->norm_idx == 0 || a->norm_idx == 1, GRAND_EINVAL, "attention_norm_idx must be 0 or 1");
  const bool needs_scalars = a->score_type == GRAND_SCORE_EXP_KERNEL || a->score_type == GRAND_SCORE_EXP_KERNEL2;
  GRAND_REQUIRE(!needs_scalars || a->kernel_scalars != nullptr, GRAND_EINVAL, "exp_kernel needs kernel_scalars");
  out->ad = a->ad;
  out->ad_eff = a->score_type == GRAND_SCORE_EXP_KERNEL2 ? 2 * a->ad : a->ad;
  out->heads = a->heads;
  out->dk = a->ad / a->heads;
  out->score_type = a->score_type;
  out->norm_type = a->norm_type;
  out->norm_idx = a->norm_idx;
  out->sqrt_dk = (float)sqrt((double)out->dk);
  out->kernel_scalars = a->kernel_scalars;
  out->edge_weight_csr = a->edge_weight_csr;
  return GRAND_OK;
}

static EpiView view(const grand_epilogue* ep) {
  EpiView v;
  v.alpha_train = ep->alpha_train; v.beta_train = ep->beta_train; v.x0 = ep->x0;
  v.no_alpha_sigmoid = ep->no_alpha_sigmoid; v.n_terms = ep->n_terms; v.f_out = ep->f_out; v.y_out = ep->y_out;
  v.coef_f = ep->coef_f;
  for (int t = 0; t < GRAND_MAX_TERMS; ++t) {
    v.coef[t] = t < ep->n_terms ? ep->coef[t] : 0.f;
    v.term[t] = t < ep->n_terms ? ep->term[t] : nullptr;
  }
  return v;
}

static int check_epilogue(const grand_epilogue* ep) {
  GRAND_REQUIRE(ep != nullptr && ep->alpha_train != nullptr, GRAND_EINVAL, "epilogue: alpha_train is null");
  GRAND_REQUIRE(ep->x0 == nullptr || ep->beta_train != nullptr, GRAND_EINVAL, "epilogue: x0 without beta_train");
  GRAND_REQUIRE(ep->n_terms >= 0 && ep->n_terms <= GRAND_MAX_TERMS, GRAND_EINVAL, "epilogue: n_terms %d",
                ep->n_terms);
  GRAND_REQUIRE(ep->f_out != nullptr || ep->y_out != nullptr, GRAND_EINVAL, "epilogue: no output requested");
  for (int t = 0; t < ep->n_terms; ++t)
    GRAND_REQUIRE(ep->term[t] != nullptr, GRAND_EINVAL, "epilogue: term %d is null", t);
  return GRAND_OK;
}

static int check_graph(const grand_graph* g) {
  GRAND_REQUIRE(g != nullptr, GRAND_EINVAL, "graph is null");
  GRAND_REQUIRE(g->n_rows >= 0 && g->n_cols >= g->n_rows && g->n_edges >= 0 && g->n_tiles >= 0 &&
                    (g->n_tiles > 0 || g->n_rows == 0) && g->n_split_tiles <= g->n_tiles,
                GRAND_EINVAL, "graph: inconsistent sizes rows %d cols %d edges %d tiles %d", g->n_rows, g->n_cols,
                g->n_edges, g->n_tiles);
  GRAND_REQUIRE(g->n_tiles == 0 || (g->rowptr && g->tile_row && g->tile_beg && g->tile_nrows && g->tile_split),
                GRAND_EINVAL, "graph: null CSR / tile arrays");
  GRAND_REQUIRE(g->n_edges == 0 || (g->col && g->erow && g->eid), GRAND_EINVAL, "graph: null edge arrays");
  GRAND_REQUIRE(g->n_split_tiles == 0 || g->split_tile0 != nullptr, GRAND_EINVAL, "graph: null split_tile0");
  return GRAND_OK;
}

// Which staging the edge kernel uses for d % 4 == 0.  GRAND_STAGE=direct|cpasync|bulk overrides (A/B measurements).
static int stage_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("GRAND_STAGE");
    mode = STAGE_DIRECT;   // fastest of the three as measured in round 1 (profiles/r01b_*): occupancy beats staging
    if (e != nullptr) {
      if (!strcmp(e, "direct")) mode = STAGE_DIRECT;
      else if (!strcmp(e, "cpasync")) mode = STAGE_CPASYNC;
      else if (!strcmp(e, "bulk")) mode = STAGE_BULK;
    }
  }
  return mode;
}

template <int MODE, int VEC, int R, int STAGE>
static int launch_rhs_inst(const RhsArgs& A, cudaStream_t stream) {
  const int grid = (A.g.n_tiles + kRhsWarps - 1) / kRhsWarps;
  size_t smem = 0;
  if (STAGE != STAGE_DIRECT) smem = (size_t)kRhsWarps * ((size_t)GRAND_TILE * A.d * sizeof(float) + 16);
  auto kern = rhs_window_kernel<MODE, VEC, R, STAGE>;
  static size_t configured = 48 * 1024;
  if (smem > configured) {
    GRAND_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  kern<<<grid, kRhsThreads, smem, stream>>>(A);
  GRAND_CUDA_OK(cudaGetLastError());
  count_launch();
  return GRAND_OK;
}

template <int MODE, int VEC, int R>
static int launch_rhs_stage(const RhsArgs& A, cudaStream_t stream) {
  // staging needs 16-byte rows and a window buffer that leaves room for >= 2 blocks per SM
  const bool can_stage = VEC == 4 && (size_t)kRhsWarps * GRAND_TILE * A.d * sizeof(float) <= 100 * 1024;
  const int mode = can_stage ? stage_mode() : STAGE_DIRECT;
  if (VEC == 4 && mode == STAGE_BULK) return launch_rhs_inst<MODE, VEC, R, VEC == 4 ? STAGE_BULK : STAGE_DIRECT>(A, stream);
  if (VEC == 4 && mode == STAGE_CPASYNC)
    return launch_rhs_inst<MODE, VEC, R, VEC == 4 ? STAGE_CPASYNC : STAGE_DIRECT>(A, stream);
  return launch_rhs_inst<MODE, VEC, R, STAGE_DIRECT>(A, stream);
}

template <int MODE>
static int launch_rhs_mode(const RhsArgs& A, cudaStream_t stream) {
  const int d = A.d;
  if (A.g.n_tiles == 0) return GRAND_OK;
  if ((d & 3) == 0) {
    const int chunks = (d / 4 + kWarp - 1) / kWarp;
    if (chunks <= 1) return launch_rhs_stage<MODE, 4, 1>(A, stream);
    if (chunks <= 2) return launch_rhs_stage<MODE, 4, 2>(A, stream);
    if (chunks <= 4) return launch_rhs_stage<MODE, 4, 4>(A, stream);
    GRAND_REQUIRE(false, GRAND_EUNSUPPORTED, "state width %d > 512 is not built", d);
  }
  const int chunks = (d + kWarp - 1) / kWarp;
  if (chunks <= 1) return launch_rhs_stage<MODE, 1, 1>(A, stream);
  if (chunks <= 2) return launch_rhs_stage<MODE, 1, 2>(A, stream);
  if (chunks <= 4) return launch_rhs_stage<MODE, 1, 4>(A, stream);
  if (chunks <= 8) return launch_rhs_stage<MODE, 1, 8>(A, stream);
  GRAND_REQUIRE(false, GRAND_EUNSUPPORTED, "state width %d (not a multiple of 4) > 256 is not built", d);
  return GRAND_EUNSUPPORTED;
}

int launch_project_qk(const float* x, int n_rows, int n_cols, int d, const float* wq, const float* bq,
                      const float* wk, const float* bk, int ad, float* q, float* k, cudaStream_t stream);

static bool fused_ok(const AttnView& a) {
  return a.score_type == GRAND_SCORE_SCALED_DOT && a.norm_idx == 0 && a.norm_type == GRAND_NORM_SOFTMAX;
}

// scores -> (gmax) -> segment stats; leaves ws->scores / ws->seg_stats / ws->gmax filled
static int run_general_prepass(const GraphView& g, const AttnView& a, const grand_workspace* ws,
                               cudaStream_t stream) {
  GRAND_REQUIRE(ws->scores && ws->seg_stats && ws->gmax, GRAND_EINVAL,
                "general attention path needs workspace scores / seg_stats / gmax");
  GRAND_REQUIRE(a.norm_idx == 0 || (g.colptr && g.csc_pos), GRAND_EINVAL,
                "attention_norm_idx = 1 needs the CSC view (colptr / csc_pos)");
  const bool sp = a.norm_type == GRAND_NORM_SQUAREPLUS;
  if (sp) { init_gmax_kernel<<<1, 1, 0, stream>>>(ws->gmax); count_launch(); }
  if (g.n_edges > 0) {
    const int grid = (g.n_edges + 255) / 256;
    const bool vec4 = ((a.dk & 3) == 0) && ((a.ad_eff & 3) == 0);
    if (vec4) edge_scores_kernel<true><<<grid, 256, 0, stream>>>(g, a, ws->q, ws->k, ws->scores, sp ? ws->gmax : nullptr);
    else edge_scores_kernel<false><<<grid, 256, 0, stream>>>(g, a, ws->q, ws->k, ws->scores, sp ? ws->gmax : nullptr);
    count_launch();
  }
  if (sp) { decode_gmax_kernel<<<1, 1, 0, stream>>>(ws->gmax); count_launch(); }
  const int n_seg = a.norm_idx == 0 ? g.n_rows : g.n_cols;
  if (n_seg > 0) {
    segment_stats_kernel<<<(n_seg + 7) / 8, 256, 0, stream>>>(g, a, ws->scores, ws->gmax, ws->seg_stats);
    count_launch();
  }
  GRAND_CUDA_OK(cudaGetLastError());
  return GRAND_OK;
}

int rhs_attention_impl(const grand_graph* g, const grand_attention* a, const float* x, const grand_workspace* ws,
                       const grand_epilogue* ep, bool project, cudaStream_t stream) {
  int rc;
  if ((rc = check_graph(g)) != GRAND_OK) return rc;
  if ((rc = check_epilogue(ep)) != GRAND_OK) return rc;
  GRAND_REQUIRE(a != nullptr && x != nullptr && ws != nullptr, GRAND_EINVAL, "rhs_attention: null argument");
  GRAND_REQUIRE(ws->q && ws->k, GRAND_EINVAL, "rhs_attention: workspace q / k are null");
  RhsArgs A;
  A.g = view(g);
  if ((rc = make_attn_view(a, &A.a)) != GRAND_OK) return rc;
  A.ep = view(ep);
  A.d = a->d;
  A.x = x; A.q = ws->q; A.k = ws->k; A.w_csr = nullptr;
  A.scores = ws->scores; A.seg_stats = ws->seg_stats; A.gmax = ws->gmax;
  A.split_stats = ws->split_stats; A.split_partial = ws->split_partial; A.split_counter = ws->split_counter;
  GRAND_REQUIRE(g->n_split_tiles == 0 || (ws->split_partial && ws->split_counter), GRAND_EINVAL,
                "graph has split rows: workspace split_partial / split_counter are null");
  if (project) {
    rc = launch_project_qk(x, g->n_rows, g->n_cols, a->d, a->wq, a->bq, a->wk, a->bk, A.a.ad_eff, ws->q, ws->k,
                           stream);
    if (rc != GRAND_OK) return rc;
  }
  if (fused_ok(A.a)) {
    if (g->n_split_tiles > 0) {
      GRAND_REQUIRE(ws->split_stats != nullptr, GRAND_EINVAL, "graph has split rows: workspace split_stats is null");
      split_tile_stats_kernel<<<(g->n_split_tiles + 7) / 8, 256, 0, stream>>>(A.g, A.a, ws->q, ws->k, ws->split_stats);
      GRAND_CUDA_OK(cudaGetLastError());
      count_launch();
    }
    return launch_rhs_mode<MODE_FUSED>(A, stream);
  }
  if ((rc = run_general_prepass(A.g, A.a, ws, stream)) != GRAND_OK) return rc;
  return launch_rhs_mode<MODE_PRECOMP>(A, stream);
}

int rhs_laplacian_impl(const grand_graph* g, int d, const float* x, const float* w_csr, const grand_workspace* ws,
                       const grand_epilogue* ep, cudaStream_t stream) {
  int rc;
  if ((rc = check_graph(g)) != GRAND_OK) return rc;
  if ((rc = check_epilogue(ep)) != GRAND_OK) return rc;
  GRAND_REQUIRE(x != nullptr && (w_csr != nullptr || g->n_edges == 0), GRAND_EINVAL, "rhs_laplacian: null argument");
  GRAND_REQUIRE(d > 0, GRAND_EINVAL, "rhs_laplacian: d = %d", d);
  GRAND_REQUIRE(g->n_split_tiles == 0 || (ws && ws->split_partial && ws->split_counter), GRAND_EINVAL,
                "graph has split rows: workspace split_partial / split_counter are null");
  RhsArgs A;
  A.g = view(g);
  A.a = AttnView{};
  A.ep = view(ep);
  A.d = d;
  A.x = x; A.q = nullptr; A.k = nullptr; A.w_csr = w_csr;
  A.scores = nullptr; A.seg_stats = nullptr; A.gmax = nullptr; A.split_stats = nullptr;
  A.split_partial = ws ? ws->split_partial : nullptr;
  A.split_counter = ws ? ws->split_counter : nullptr;
  return launch_rhs_mode<MODE_WEIGHT>(A, stream);
}

int attention_weights_impl(const grand_graph* g, const grand_attention* a, const float* x, const grand_workspace* ws,
                           float* attention_out, float* prods_out, float* mean_csr_out, cudaStream_t stream) {
  int rc;
  if ((rc = check_graph(g)) != GRAND_OK) return rc;
  GRAND_REQUIRE(a != nullptr && x != nullptr && ws != nullptr && ws->q && ws->k, GRAND_EINVAL,
                "attention_weights: null argument");
  AttnView av;
  if ((rc = make_attn_view(a, &av)) != GRAND_OK) return rc;
  GraphView gv = view(g);
  rc = launch_project_qk(x, g->n_rows, g->n_cols, a->d, a->wq, a->bq, a->wk, a->bk, av.ad_eff, ws->q, ws->k, stream);
  if (rc != GRAND_OK) return rc;
  if ((rc = run_general_prepass(gv, av, ws, stream)) != GRAND_OK) return rc;
  if (g->n_edges > 0) {
    attention_out_kernel<<<(g->n_edges + 255) / 256, 256, 0, stream>>>(gv, av, ws->scores, ws->seg_stats, ws->gmax,
                                                                      attention_out, prods_out, mean_csr_out);
    GRAND_CUDA_OK(cudaGetLastError());
    count_launch();
  }
  return GRAND_OK;
}

int edge_values_to_csr_impl(const grand_graph* g, const float* w, int heads, float* w_csr, cudaStream_t stream) {
  int rc;
  if ((rc = check_graph(g)) != GRAND_OK) return rc;
  GRAND_REQUIRE(heads >= 1 && (g->n_edges == 0 || (w && w_csr)), GRAND_EINVAL, "edge_values_to_csr: bad argument");
  if (g->n_edges > 0) {
    edge_values_to_csr_kernel<<<(g->n_edges + 255) / 256, 256, 0, stream>>>(view(g), w, heads, w_csr);
    GRAND_CUDA_OK(cudaGetLastError());
    count_launch();
  }
  return GRAND_OK;
}

}  // namespace grand
