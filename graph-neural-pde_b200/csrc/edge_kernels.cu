// The edge phase of the GRAND right-hand side.
//
//   rhs_kernel<MODE, VEC, R>  one warp per row tile (<= GRAND_TILE in-edges of one query row):
//       MODE_FUSED    scores q_i.k_j, segment softmax over the row, head mean -- all in registers -- then the
//                     attention-weighted aggregation sum_e a_e X[col_e] and the residual / RK-combine epilogue.
//                     No edge-sized tensor is ever written (reference materialises [E,d_k,H] x2, [E,H] x~10, [E,d]).
//       MODE_PRECOMP  general pipeline (attention_norm_idx = 1, squareplus, > GRAND_MAX_HEADS heads): scores and
//                     per-segment statistics come from the two pre-pass kernels below.
//       MODE_WEIGHT   GRAND-l: per-edge weights given (reference src/function_laplacian_diffusion.py:28-36).
//   Rows with more than GRAND_TILE in-edges are cut into several tiles; each writes a partial row sum, the last
//   tile to finish adds the partials in slot order (deterministic) and runs the epilogue.
//
//   split_tile_stats_kernel   (max, sum-exp) per head for every tile of a split row, merged by the fused kernel.
//   edge_scores_kernel        prods[E,H] in CSR order (+ global max for squareplus)         reference :190-209
//   segment_stats_kernel      per segment (row or column) max / normaliser sum               reference :210-213
//   attention_out_kernel      attention[E,H] / prods[E,H] back in the caller's edge order, head-mean in CSR order
//
// HBM/L2-bound gather work: every load of a neighbour row is one fully coalesced 128-bit-per-lane warp access.
#include "attention.cuh"

namespace grand {

constexpr int MODE_FUSED = 0;
constexpr int MODE_PRECOMP = 1;
constexpr int MODE_WEIGHT = 2;

constexpr int kEPL = GRAND_TILE / kWarp;  // edge slots per lane
constexpr int kHMax = GRAND_MAX_HEADS;
constexpr int kWarpsPerBlock = 8;
constexpr int kUnroll = 4;                // neighbour rows in flight per warp in the aggregation loop

struct GraphView {
  int n_rows, n_cols, n_edges, n_tiles, n_split_rows, n_split_tiles;
  const int *rowptr, *col, *erow, *eid, *tile_row, *tile_beg, *tile_split, *split_tile0, *colptr, *csc_pos;
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
// lane-per-edge scores of one tile
// ----------------------------------------------------------------------------------------------------------
// The fused path is built for the reference's default scorer (scaled_dot, :195-196); the other scorers go through
// the general pipeline (edge_scores_kernel).
template <bool kVec4>
__device__ __forceinline__ void tile_scores(const AttnView& a, const float* __restrict__ q_row,
                                            const float* __restrict__ k, int beg, int end, int lane,
                                            const int (&jcol)[kEPL], float (&sc)[kEPL][kHMax]) {
#pragma unroll
  for (int s = 0; s < kEPL; ++s) {
    const int p = beg + s * kWarp + lane;
    const bool valid = p < end;
    const float* kj = k + (size_t)jcol[s] * a.ad_eff;
    float ew = 1.f;
    if (valid && a.edge_weight_csr != nullptr) ew = __ldg(a.edge_weight_csr + p);
#pragma unroll
    for (int h = 0; h < kHMax; ++h) {
      sc[s][h] = -INFINITY;
      if (h < a.heads && valid) {
        float v = head_dot<kVec4>(q_row + h * a.dk, kj + h * a.dk, a.dk) / a.sqrt_dk;
        if (a.edge_weight_csr != nullptr) v *= ew;   // reweight_attention, reference :208-209
        sc[s][h] = v;
      }
    }
  }
}

template <bool kVec4>
__global__ void __launch_bounds__(kWarpsPerBlock* kWarp, 3)
split_tile_stats_kernel(GraphView g, AttnView a, const float* __restrict__ q, const float* __restrict__ k,
                        float* __restrict__ split_stats) {
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);   // split tiles are tiles [0, n_split_tiles)
  if (t >= g.n_split_tiles) return;
  const int row = g.tile_row[t];
  const int beg = g.tile_beg[t];
  const int end = min(beg + GRAND_TILE, g.rowptr[row + 1]);
  int jcol[kEPL];
#pragma unroll
  for (int s = 0; s < kEPL; ++s) {
    const int p = beg + s * kWarp + lane;
    jcol[s] = p < end ? __ldg(g.col + p) : 0;
  }
  float sc[kEPL][kHMax];
  tile_scores<kVec4>(a, q + (size_t)row * a.ad_eff, k, beg, end, lane, jcol, sc);
#pragma unroll
  for (int h = 0; h < kHMax; ++h) {
    if (h < a.heads) {
      float m = -INFINITY;
#pragma unroll
      for (int s = 0; s < kEPL; ++s) m = fmaxf(m, sc[s][h]);
      m = warp_max(m);
      float l = 0.f;
#pragma unroll
      for (int s = 0; s < kEPL; ++s) l += (sc[s][h] == -INFINITY) ? 0.f : expf(sc[s][h] - m);
      l = warp_sum(l);
      if (lane == 0) {
        split_stats[((size_t)t * a.heads + h) * 2 + 0] = m;
        split_stats[((size_t)t * a.heads + h) * 2 + 1] = l;
      }
    }
  }
}

// ----------------------------------------------------------------------------------------------------------
// epilogue: residual + RK stage combine for one finished row
// ----------------------------------------------------------------------------------------------------------
template <int VEC, int R>
__device__ __forceinline__ void finish_row(const RhsArgs& A, int row, int lane, float (&acc)[R][VEC]) {
  const int d = A.d;
  float alpha = __ldg(A.ep.alpha_train);
  if (!A.ep.no_alpha_sigmoid) alpha = 1.f / (1.f + expf(-alpha));   // torch.sigmoid
  const float beta = A.ep.x0 != nullptr ? __ldg(A.ep.beta_train) : 0.f;
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
template <int MODE, int VEC, int R>
__global__ void __launch_bounds__(kWarpsPerBlock* kWarp, 3) rhs_kernel(const __grid_constant__ RhsArgs A) {
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (t >= A.g.n_tiles) return;
  const int d = A.d;
  const int row = __ldg(A.g.tile_row + t);
  const int beg = __ldg(A.g.tile_beg + t);
  const int end = min(beg + GRAND_TILE, __ldg(A.g.rowptr + row + 1));
  const int n = end - beg;
  const int split = __ldg(A.g.tile_split + t);

  int jcol[kEPL];
  float ab[kEPL];   // head-mean attention of the lane's edges (or the given weight)
#pragma unroll
  for (int s = 0; s < kEPL; ++s) {
    const int p = beg + s * kWarp + lane;
    jcol[s] = p < end ? __ldg(A.g.col + p) : 0;
    ab[s] = 0.f;
  }

  if (MODE == MODE_WEIGHT) {
#pragma unroll
    for (int s = 0; s < kEPL; ++s) {
      const int p = beg + s * kWarp + lane;
      if (p < end) ab[s] = __ldg(A.w_csr + p);
    }
  } else if (MODE == MODE_PRECOMP) {
    const int H = A.a.heads;
    const float gmax = A.a.norm_type == GRAND_NORM_SQUAREPLUS ? __ldg(A.gmax) : 0.f;
#pragma unroll
    for (int s = 0; s < kEPL; ++s) {
      const int p = beg + s * kWarp + lane;
      if (p < end) {
        const int seg = A.a.norm_idx == 0 ? row : jcol[s];
        float a_sum = 0.f;
        for (int h = 0; h < H; ++h) {
          const float sc = __ldg(A.scores + (size_t)p * H + h);
          const float m = __ldg(A.seg_stats + ((size_t)seg * H + h) * 2 + 0);
          const float l = __ldg(A.seg_stats + ((size_t)seg * H + h) * 2 + 1);
          const float num = norm_num(A.a.norm_type, sc, A.a.norm_type == GRAND_NORM_SOFTMAX ? m : gmax);
          a_sum += num / (l + 1e-16f);
        }
        ab[s] = a_sum / (float)H;          // attention.mean(dim=1), reference :34
      }
    }
  } else {  // MODE_FUSED: softmax over the row, attention_norm_idx == 0
    const bool vec4 = ((A.a.dk & 3) == 0) && ((A.a.ad_eff & 3) == 0);
    float sc[kEPL][kHMax];
    const float* q_row = A.q + (size_t)row * A.a.ad_eff;
    if (vec4) tile_scores<true>(A.a, q_row, A.k, beg, end, lane, jcol, sc);
    else tile_scores<false>(A.a, q_row, A.k, beg, end, lane, jcol, sc);
    int slot0 = 0, slot1 = 0;
    if (split >= 0) {
      slot0 = __ldg(A.g.split_tile0 + split);
      slot1 = __ldg(A.g.split_tile0 + split + 1);
    }
#pragma unroll
    for (int h = 0; h < kHMax; ++h) {
      if (h < A.a.heads) {
        float m, l;
        if (split < 0) {
          m = -INFINITY;
#pragma unroll
          for (int s = 0; s < kEPL; ++s) m = fmaxf(m, sc[s][h]);
          m = warp_max(m);
          l = 0.f;
#pragma unroll
          for (int s = 0; s < kEPL; ++s) l += (sc[s][h] == -INFINITY) ? 0.f : expf(sc[s][h] - m);
          l = warp_sum(l);
        } else {
          // merge the per-tile (max, sum-exp) pairs of this row: l = sum_t l_t exp(m_t - m)
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
        const float inv = l + 1e-16f;
#pragma unroll
        for (int s = 0; s < kEPL; ++s)
          if (sc[s][h] != -INFINITY) ab[s] += expf(sc[s][h] - m) / inv;
      }
    }
    const float hf = (float)A.a.heads;
#pragma unroll
    for (int s = 0; s < kEPL; ++s) ab[s] = ab[s] / hf;
  }

  // ---- aggregation: acc = sum_e ab_e * X[col_e]  (torch_sparse.spmm, reference :35) ----------------------
  float acc[R][VEC];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[r][v] = 0.f;

#pragma unroll
  for (int s = 0; s < kEPL; ++s) {
    const int cnt = min(kWarp, n - s * kWarp);
    for (int e0 = 0; e0 < cnt; e0 += kUnroll) {
      int j[kUnroll];
      float a[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int src = min(e0 + u, cnt - 1);
        j[u] = __shfl_sync(kFull, jcol[s], src);
        a[u] = __shfl_sync(kFull, ab[s], src);
        if (e0 + u >= cnt) a[u] = 0.f;
      }
      float xv[kUnroll][R][VEC];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const float* xr = A.x + (size_t)j[u] * d;
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
      for (int u = 0; u < kUnroll; ++u)
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int v = 0; v < VEC; ++v) acc[r][v] = fmaf(a[u], xv[u][r][v], acc[r][v]);
    }
  }

  if (split < 0) {
    finish_row<VEC, R>(A, row, lane, acc);
    return;
  }

  // ---- split row: publish the partial sum, the last tile of the row reduces and finishes -------------------
  {
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
    finish_row<VEC, R>(A, row, lane, acc);
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
  v.tile_beg = g->tile_beg; v.tile_split = g->tile_split; v.split_tile0 = g->split_tile0;
  v.colptr = g->colptr; v.csc_pos = g->csc_pos;
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
  GRAND_REQUIRE(g->n_rows >= 0 && g->n_cols >= g->n_rows && g->n_edges >= 0 && g->n_tiles >= g->n_rows, GRAND_EINVAL,
                "graph: inconsistent sizes rows %d cols %d edges %d tiles %d", g->n_rows, g->n_cols, g->n_edges,
                g->n_tiles);
  GRAND_REQUIRE(g->n_tiles == 0 || (g->rowptr && g->tile_row && g->tile_beg && g->tile_split), GRAND_EINVAL,
                "graph: null CSR / tile arrays");
  GRAND_REQUIRE(g->n_edges == 0 || (g->col && g->erow && g->eid), GRAND_EINVAL, "graph: null edge arrays");
  GRAND_REQUIRE(g->n_split_tiles == 0 || g->split_tile0 != nullptr, GRAND_EINVAL, "graph: null split_tile0");
  return GRAND_OK;
}

template <int MODE>
static int launch_rhs_mode(const RhsArgs& A, cudaStream_t stream) {
  const int d = A.d;
  const int grid = (A.g.n_tiles + kWarpsPerBlock - 1) / kWarpsPerBlock;
  if (grid == 0) return GRAND_OK;
  const dim3 block(kWarpsPerBlock * kWarp);
  if ((d & 3) == 0) {
    const int chunks = (d / 4 + kWarp - 1) / kWarp;
    if (chunks <= 1) rhs_kernel<MODE, 4, 1><<<grid, block, 0, stream>>>(A);
    else if (chunks <= 2) rhs_kernel<MODE, 4, 2><<<grid, block, 0, stream>>>(A);
    else if (chunks <= 4) rhs_kernel<MODE, 4, 4><<<grid, block, 0, stream>>>(A);
    else GRAND_REQUIRE(false, GRAND_EUNSUPPORTED, "state width %d > 512 is not built", d);
  } else {
    const int chunks = (d + kWarp - 1) / kWarp;
    if (chunks <= 1) rhs_kernel<MODE, 1, 1><<<grid, block, 0, stream>>>(A);
    else if (chunks <= 2) rhs_kernel<MODE, 1, 2><<<grid, block, 0, stream>>>(A);
    else if (chunks <= 4) rhs_kernel<MODE, 1, 4><<<grid, block, 0, stream>>>(A);
    else if (chunks <= 8) rhs_kernel<MODE, 1, 8><<<grid, block, 0, stream>>>(A);
    else GRAND_REQUIRE(false, GRAND_EUNSUPPORTED, "state width %d (not a multiple of 4) > 256 is not built", d);
  }
  GRAND_CUDA_OK(cudaGetLastError());
  count_launch();
  return GRAND_OK;
}

int launch_project_qk(const float* x, int n_rows, int n_cols, int d, const float* wq, const float* bq,
                      const float* wk, const float* bk, int ad, float* q, float* k, cudaStream_t stream);

static bool fused_ok(const AttnView& a) {
  return a.score_type == GRAND_SCORE_SCALED_DOT && a.norm_idx == 0 && a.norm_type == GRAND_NORM_SOFTMAX &&
         a.heads <= kHMax;
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
      const int grid = (g->n_split_tiles + kWarpsPerBlock - 1) / kWarpsPerBlock;
      const bool vec4 = ((A.a.dk & 3) == 0) && ((A.a.ad_eff & 3) == 0);
      if (vec4) split_tile_stats_kernel<true><<<grid, kWarpsPerBlock * kWarp, 0, stream>>>(A.g, A.a, ws->q, ws->k, ws->split_stats);
      else split_tile_stats_kernel<false><<<grid, kWarpsPerBlock * kWarp, 0, stream>>>(A.g, A.a, ws->q, ws->k, ws->split_stats);
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
