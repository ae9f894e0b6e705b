// The dominant kernel: one warp per WINDOW of <= GRAND_TILE (32) CSR edges, one edge per lane.
//
// A window is either several whole query rows packed together (average in-degree of the reference's graphs is 5-30) or
// one 32-edge chunk of a long row.  Everything a lane needs to know about its edge was laid out at plan time in ONE
// 8-byte record (column, row-in-window, the lane range of its row), so the dependent chain of global loads is
//     window header + lane record  ->  { K_j, Q_i, X_j gathers, L2 prefetch of the epilogue rows }  ->  arithmetic
// instead of tile -> rowptr -> col -> erow -> K/Q -> X.
//
//   MODE_FUSED    scores q_i.k_j / sqrt(d_k) lane-parallel, softmax over each row's lanes with segmented shuffle
//                 reductions, head mean (reference src/function_transformer_attention.py:190-213, :34); nothing
//                 edge-sized is written (the reference materialises [E,d_k,H] x2, ~10 [E,H] tensors and [E,d]).
//   MODE_PRECOMP  general pipeline (attention_norm_idx = 1, squareplus, cosine / pearson / exp-kernel scorers).
//   MODE_WEIGHT   GRAND-l, per-edge weights given (reference src/function_laplacian_diffusion.py:28-36).
//   then per row: a_e-weighted sum of the neighbour rows (torch_sparse.spmm, :35), residual alpha (ax - x) + beta x0
//   and the Runge-Kutta stage combine as the epilogue.
//   Neighbour rows: STAGE_BULK one cp.async.bulk (TMA bulk copy + mbarrier) per lane into the warp's shared-memory
//   window, issued BEFORE the score phase; STAGE_CPASYNC the same with LDGSTS; STAGE_DIRECT batches of register loads.
//   Chunks of long rows publish partial sums; the last chunk to finish adds them in slot order (deterministic).
#include <stdlib.h>
#include <string.h>

#include "edge_common.cuh"

namespace grand {

constexpr int kRhsThreads = 128;  // 4 warps per block
constexpr int kRhsWarps = kRhsThreads / kWarp;
// Register staging: neighbour rows in flight per warp and resident blocks per SM.  Measured on ogbn-arxiv (B200,
// profiles/r01f_*): batch 8 x 8 blocks (32 warps/SM, 64 regs) 337 us, batch 4 x 12 blocks (48 warps/SM, 42 regs) 327 us,
// batch 4 x 10 blocks 332 us, batch 8 x 6 blocks 392 us -- more resident windows beat deeper batches.
#ifndef GRAND_DIRECT_BATCH
#define GRAND_DIRECT_BATCH 4
#endif
#ifndef GRAND_DIRECT_MINBLOCKS
#define GRAND_DIRECT_MINBLOCKS 12
#endif
// (The A/B build switches of round 1 -- shuffle segment max, two-at-a-time epilogue terms, persistent loop, gather-first,
// no L2 prefetch, L1 epilogue prefetch -- were measured in round 2 (profiles/r02_edge_ab_runs.txt), none won; they are
// resolved to the validated defaults here: same SASS.)
constexpr int kDirectBatch = GRAND_DIRECT_BATCH;   // STAGE_DIRECT: neighbour rows in flight per warp (divided by R)

// q.k over one head, then the reference's division by np.sqrt(d_k) (:196); for d_k a power of 4 the division is an
// exact power-of-two scaling and is done as a multiplication (bit-identical).
__device__ __forceinline__ float fused_score(const AttnView& a, const float* __restrict__ qh,
                                             const float* __restrict__ kh, bool vec4) {
  const float s = vec4 ? head_dot<true>(qh, kh, a.dk) : head_dot<false>(qh, kh, a.dk);
  return a.inv_sqrt_dk_exact != 0.f ? s * a.inv_sqrt_dk_exact : s / a.sqrt_dk;
}

// ---- long rows, first step: (max, sum-exp) per head and chunk -------------------------------------------------------
__global__ void __launch_bounds__(256)
split_tile_stats_kernel(GraphView g, AttnView a, const float* __restrict__ q, const float* __restrict__ k,
                        float* __restrict__ split_stats) {
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // split tiles are tiles [0, n_split_tiles)
  if (t >= g.n_split_tiles) return;
  const int4 hdr = __ldg(g.win_hdr + t);
  const int2 rec = __ldg(g.win_rec + (size_t)t * kWarp + lane);
  const int row = hdr.x, cnt = hdr.z >> 8;
  const bool valid = lane < cnt;
  const float* kj = k + (size_t)rec.x * a.ad_eff;
  const float* qi = q + (size_t)row * a.ad_eff;
  const bool vec4 = ((a.dk & 3) == 0) && ((a.ad_eff & 3) == 0);
  const float ew = (valid && a.edge_weight_csr != nullptr) ? __ldg(a.edge_weight_csr + hdr.y + lane) : 1.f;
  const int H = a.heads;
  for (int h = 0; h < H; ++h) {
    float s = -INFINITY;
    if (valid) s = fused_score(a, qi + h * a.dk, kj + h * a.dk, vec4) * ew;
    const float m = warp_max(s);
    const float l = warp_sum(valid ? expf(s - m) : 0.f);
    if (lane == 0) {
      split_stats[((size_t)t * H + h) * 2 + 0] = m;
      split_stats[((size_t)t * H + h) * 2 + 1] = l;
    }
  }
}

// ---- long rows, second step: one warp per long row merges its chunks' pairs (no fences, no atomics: stream order
//      between the two small kernels is the synchronisation) ---------------------------------------------------
__global__ void __launch_bounds__(256)
split_row_merge_kernel(GraphView g, int H, const float* __restrict__ split_stats, float* __restrict__ row_stats) {
  const int lane = threadIdx.x & 31;
  const int split = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (split >= g.n_split_rows) return;
  const int slot0 = __ldg(g.split_tile0 + split), slot1 = __ldg(g.split_tile0 + split + 1);
  // last chunk of the row merges the per-chunk pairs: m = max_t m_t, l = sum_t l_t exp(m_t - m).  One pass with a
  // running (m, l) per lane over (chunk, head) items -- 8-byte loads, four in flight -- then a shuffle merge; the
  // longest row of ogbn-arxiv has ~330 chunks and this merge is the tail of the kernel.
  const float2* pairs = reinterpret_cast<const float2*>(split_stats);
  for (int h = 0; h < H; ++h) {
    float m = -INFINITY, l = 0.f;
    int u = slot0 + lane;
    for (; u + 3 * kWarp < slot1; u += 4 * kWarp) {
      float2 p[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) p[i] = __ldg(pairs + (size_t)(u + i * kWarp) * H + h);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float mn = fmaxf(m, p[i].x);
        l = l * expf(m - mn) + p[i].y * expf(p[i].x - mn);
        m = mn;
      }
    }
    for (; u < slot1; u += kWarp) {
      const float2 p = __ldg(pairs + (size_t)u * H + h);
      const float mn = fmaxf(m, p.x);
      l = l * expf(m - mn) + p.y * expf(p.x - mn);
      m = mn;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float mo = __shfl_xor_sync(kFull, m, o), lo2 = __shfl_xor_sync(kFull, l, o);
      const float mn = fmaxf(m, mo);
      // lanes that saw no chunk carry (m, l) = (-inf, 0): exp(-inf - mn) = 0 keeps them neutral (mn is finite once any
      // lane has data; two empty lanes give exp(nan)*0 -> guard)
      const float a = (m == -INFINITY) ? 0.f : l * expf(m - mn);
      const float b = (mo == -INFINITY) ? 0.f : lo2 * expf(mo - mn);
      l = a + b;
      m = mn;
    }
    if (lane == 0) {
      row_stats[((size_t)split * H + h) * 2 + 0] = m;
      row_stats[((size_t)split * H + h) * 2 + 1] = l;
    }
  }
}

// ---- epilogue: residual + RK stage combine for one finished row -----------------------------------------------------
template <int VEC, int R>
__device__ __forceinline__ void finish_row(const RhsArgs& A, int row, int lane, float alpha, float beta,
                                           uint64_t pol_stream, float (&acc)[R][VEC]) {
  const int d = A.d;
  const size_t base = (size_t)row * d;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int c = (lane + r * kWarp) * VEC;
    if (c >= d) continue;
    if (VEC == 4) {
      // issue every load of the row first, then the arithmetic
      const float4 xi = ldg4(A.x + base + c);
      float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (A.ep.x0 != nullptr) s0 = ld_stream4_hint(A.ep.x0 + base + c, pol_stream);
      constexpr int kTB = 4;   // stage-combine terms per batch of loads (rk4 needs <= 4, dopri5 up to 7)
      // every element initialised: a conditionally-written array ends up in LOCAL memory (ncu: more local than global
      // load requests in this kernel before the change), zero-initialised it stays in 16 registers
      float4 tv[kTB];
#pragma unroll
      for (int t = 0; t < kTB; ++t) tv[t] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (A.ep.y_out != nullptr) {
#pragma unroll
        for (int t = 0; t < kTB; ++t)
          if (t < A.ep.n_terms) tv[t] = ld_stream4_hint(A.ep.term[t] + base + c, pol_stream);
      }
      float4 f;
      f.x = alpha * (acc[r][0] - xi.x); f.y = alpha * (acc[r][1 % VEC] - xi.y);     // f = alpha (ax - x)
      f.z = alpha * (acc[r][2 % VEC] - xi.z); f.w = alpha * (acc[r][3 % VEC] - xi.w);
      if (A.ep.x0 != nullptr) {                                                     // + beta x0   (add_source)
        f.x += beta * s0.x; f.y += beta * s0.y; f.z += beta * s0.z; f.w += beta * s0.w;
      }
      if (A.ep.f_out != nullptr) st_stream4_hint(A.ep.f_out + base + c, f, pol_stream);
      if (A.ep.y_out != nullptr) {
        const float cf = epi_coef_f(A.ep);
        float4 y = make_float4(cf * f.x, cf * f.y, cf * f.z, cf * f.w);
#pragma unroll
        for (int t = 0; t < kTB; ++t) {
          if (t < A.ep.n_terms) {
            const float ct = epi_coef(A.ep, t);
            y.x = fmaf(ct, tv[t].x, y.x); y.y = fmaf(ct, tv[t].y, y.y);
            y.z = fmaf(ct, tv[t].z, y.z); y.w = fmaf(ct, tv[t].w, y.w);
          }
        }
        for (int t = kTB; t < A.ep.n_terms; ++t) {
          const float ct = epi_coef(A.ep, t);
          const float4 v = ld_stream4_hint(A.ep.term[t] + base + c, pol_stream);
          y.x = fmaf(ct, v.x, y.x); y.y = fmaf(ct, v.y, y.y); y.z = fmaf(ct, v.z, y.z); y.w = fmaf(ct, v.w, y.w);
        }
        // y_out is the next stage's gather source: keep it cacheable
        *reinterpret_cast<float4*>(A.ep.y_out + base + c) = y;
      }
    } else {
      const float xi = __ldg(A.x + base + c);
      float f = alpha * (acc[r][0] - xi);
      if (A.ep.x0 != nullptr) f += beta * __ldg(A.ep.x0 + base + c);
      if (A.ep.f_out != nullptr) A.ep.f_out[base + c] = f;
      if (A.ep.y_out != nullptr) {
        float y = epi_coef_f(A.ep) * f;
#pragma unroll
        for (int t = 0; t < GRAND_MAX_TERMS; ++t)
          if (t < A.ep.n_terms) y = fmaf(epi_coef(A.ep, t), __ldg(A.ep.term[t] + base + c), y);
        A.ep.y_out[base + c] = y;
      }
    }
  }
}

// ---- STAGE_DIRECT: NB neighbour rows of one query row in flight in registers, then accumulated ----------------------
template <int VEC, int R, int NB>
__device__ __forceinline__ void gather_batch(const float* __restrict__ xl, int d, int lane, bool full, int j, float ab,
                                             int e, uint64_t pol_keep, float (&acc)[R][VEC]) {
  float aa[NB];
  float xv[NB][R][VEC];
#pragma unroll
  for (int u = 0; u < NB; ++u) {
    const int ju = __shfl_sync(kFull, j, e + u);
    aa[u] = __shfl_sync(kFull, ab, e + u);
    const float* xr = xl + (size_t)ju * d;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int c = (lane + r * kWarp) * VEC;
      if (VEC == 4) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (full || c < d) v = ldg4_hint(xr + r * kWarp * VEC, pol_keep);
        xv[u][r][0] = v.x; xv[u][r][1 % VEC] = v.y; xv[u][r][2 % VEC] = v.z; xv[u][r][3 % VEC] = v.w;
      } else {
        xv[u][r][0] = c < d ? __ldg(xr + r * kWarp) : 0.f;
      }
    }
  }
#pragma unroll
  for (int u = 0; u < NB; ++u)
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[r][v] = fmaf(aa[u], xv[u][r][v], acc[r][v]);
}

// ---- the RHS kernel -----------------------------------------------------------------------------------------------------
template <int MODE, int VEC, int R, int STAGE>
__global__ void __launch_bounds__(kRhsThreads, (STAGE == STAGE_DIRECT || STAGE == STAGE_ROWBUF) ? GRAND_DIRECT_MINBLOCKS : 3)
rhs_window_kernel(const __grid_constant__ RhsArgs A) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int t = A.tile_begin + blockIdx.x * kRhsWarps + warp;
  if (t >= A.tile_end) return;
  const int d = A.d;
  const int row_bytes = d * (int)sizeof(float);

  // ---- level 1: window header (broadcast) + per-lane record (coalesced) ---------------------------------------------
  const int4 hdr = __ldg(A.g.win_hdr + t);
  const int2 rec = __ldg(A.g.win_rec + (size_t)t * kWarp + lane);
  const int row0 = hdr.x, beg = hdr.y, nrows = hdr.z & 0xff, cnt = hdr.z >> 8, split = hdr.w;
  const int j = rec.x;
  const int kk = rec.y & 0xff, lo = (rec.y >> 8) & 0x3f, hi = (rec.y >> 14) & 0x3f;
  const bool valid = lane < cnt;
  const int i = row0 + kk;
  const uint64_t pol_keep = (A.l2_hints & 1) ? policy_evict_last() : policy_evict_normal();
  const uint64_t pol_stream = (A.l2_hints & 2) ? policy_evict_first() : policy_evict_normal();

  // ---- level 2a: start moving every neighbour row of the window --------------------------------------------------
  float* stage = nullptr;
  uint32_t bar = 0;
  if (STAGE == STAGE_CPASYNC || STAGE == STAGE_BULK) {
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
      if (valid)
        bulk_copy_g2s(smem_u32(stage) + lane * row_bytes, A.x + (size_t)j * d, (uint32_t)row_bytes, bar, pol_keep);
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
  // epilogue inputs of the window's rows: contiguous rows, pull them towards L2 now (one 128-byte line per lane)
  if (VEC == 4) {
    const int lines = (nrows * row_bytes + 127) >> 7;
    for (int ln = lane; ln < lines; ln += kWarp) {
      const size_t off = (size_t)row0 * row_bytes + (size_t)ln * 128;
      if (A.ep.x0 != nullptr) prefetch_l2(reinterpret_cast<const char*>(A.ep.x0) + off);
      for (int tt = 0; tt < A.ep.n_terms; ++tt) prefetch_l2(reinterpret_cast<const char*>(A.ep.term[tt]) + off);
    }
  }

  // ---- level 2b: per-edge weight ab = head-mean attention (or the given weight) -------------------------------------
  float ab = 0.f;
  if (MODE == MODE_WEIGHT) {
    if (valid) ab = __ldg(A.w_csr + beg + lane);
  } else if (MODE == MODE_PRECOMP) {
    const int H = A.a.heads;
    if (valid) {
      const size_t p = (size_t)beg + lane;
      const float gmax = A.a.norm_type == GRAND_NORM_SQUAREPLUS ? __ldg(A.gmax) : 0.f;
      const int seg = A.a.norm_idx == 0 ? i : j;
      float a_sum = 0.f;
      for (int h = 0; h < H; ++h) {
        const float sc = __ldg(A.scores + p * H + h);
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
    const float ew = (valid && A.a.edge_weight_csr != nullptr) ? __ldg(A.a.edge_weight_csr + beg + lane) : 1.f;
    const int H = A.a.heads;
    const unsigned smask = seg_mask(lane, lo, hi, valid && split < 0);
    if (A.a.dk == 4 && vec4 && H <= 4) {
      // d_k = 4 (ogbn-arxiv: ad 16, 4 heads): a head is ONE float4 of the Q / K row.  All 2 H loads are issued before
      // the first score so the window pays one round trip for them, not one per head.
      float4 qv[4], kv[4];
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        qv[h] = make_float4(0.f, 0.f, 0.f, 0.f);
        kv[h] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (h < H && valid) {
          qv[h] = ldg4(qi + 4 * h);
          kv[h] = ldg4(kj + 4 * h);
        }
      }
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        if (h < H) {
          float s = -INFINITY;
          if (valid) {
            float t = 0.f;
            t = fmaf(qv[h].x, kv[h].x, t); t = fmaf(qv[h].y, kv[h].y, t);
            t = fmaf(qv[h].z, kv[h].z, t); t = fmaf(qv[h].w, kv[h].w, t);
            s = (t * 0.5f) * ew;              // / sqrt(4): exact; ew: reweight_attention :208-209
          }
          float m, l, e;
          if (split < 0) {
            m = seg_max_redux(s, smask);
            e = valid ? expf(s - m) : 0.f;
            l = seg_sum(e, lane, lo, hi);
          } else {
            m = __ldg(A.row_stats + ((size_t)split * H + h) * 2 + 0);
            l = __ldg(A.row_stats + ((size_t)split * H + h) * 2 + 1);
            e = valid ? expf(s - m) : 0.f;
          }
          ab += e / (l + 1e-16f);
        }
      }
    } else {
      for (int h = 0; h < H; ++h) {
        float s = -INFINITY;
        if (valid) s = fused_score(A.a, qi + h * A.a.dk, kj + h * A.a.dk, vec4) * ew;
        float m, l, e;
        if (split < 0) {
          m = seg_max_redux(s, smask);
          e = valid ? expf(s - m) : 0.f;
          l = seg_sum(e, lane, lo, hi);
        } else {
          m = __ldg(A.row_stats + ((size_t)split * H + h) * 2 + 0);
          l = __ldg(A.row_stats + ((size_t)split * H + h) * 2 + 1);
          e = valid ? expf(s - m) : 0.f;
        }
        ab += e / (l + 1e-16f);
      }
    }
    ab = ab / (float)H;
  }

  // ---- level 3: aggregation per row + epilogue ------------------------------------------------------------------------
  float alpha = __ldg(A.ep.alpha_train);
  if (!A.ep.no_alpha_sigmoid) alpha = 1.f / (1.f + expf(-alpha));   // torch.sigmoid
  const float beta = A.ep.x0 != nullptr ? __ldg(A.ep.beta_train) : 0.f;
  const bool full = (d == R * kWarp * VEC);     // every lane owns exactly R chunks: no column predicate

  if (STAGE == STAGE_BULK) mbar_wait(bar, 0);
  if (STAGE == STAGE_CPASYNC) {
    cp_async_wait_all();
    __syncwarp();
  }

  // ---- STAGE_ROWBUF: the window's edges are gathered in fixed batches that ignore row boundaries (ceil(cnt / NB)
  //      exposed latencies per window instead of one per power-of-two piece of every row); a finished row's sum waits
  //      in the warp's shared-memory slot (each lane re-reads only what it wrote: no barrier) for the epilogue ------
  if (STAGE == STAGE_ROWBUF && split < 0) {
    float* rowbuf = reinterpret_cast<float*>(smem_raw) + (size_t)warp * GRAND_MAX_PACKED_ROWS * d;
    const unsigned starts = __ballot_sync(kFull, valid && lo == lane);       // bit e: edge e opens a row
    const unsigned rowmask =                                                  // bit r: row r of the window has edges
        __ballot_sync(kFull, lane < nrows && ((rec.y >> 26) & 0x3f) > ((rec.y >> 20) & 0x3f));
    unsigned remaining = rowmask;
    int kr = -1;
    float acc[R][VEC];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[r][v] = 0.f;
    constexpr int NB = (kDirectBatch / R) < 2 ? 2 : (kDirectBatch / R);
    const float* xl = A.x + (size_t)lane * VEC;
    for (int e = 0; e < cnt; e += NB) {
      float aa[NB];
      float4 xv[NB][R];
#pragma unroll
      for (int u = 0; u < NB; ++u) {
        const int eu = e + u;
        const int ju = __shfl_sync(kFull, j, eu & 31);
        aa[u] = __shfl_sync(kFull, ab, eu & 31);
        const float* xr = xl + (size_t)ju * d;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          xv[u][r] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (eu < cnt && (full || (lane + r * kWarp) * VEC < d)) xv[u][r] = ldg4_hint(xr + r * kWarp * VEC, pol_keep);
        }
      }
#pragma unroll
      for (int u = 0; u < NB; ++u) {
        const int eu = e + u;
        if (eu < cnt) {
          if ((starts >> eu) & 1u) {
            if (kr >= 0) {
#pragma unroll
              for (int r = 0; r < R; ++r)
                if (full || (lane + r * kWarp) * VEC < d)
                  *reinterpret_cast<float4*>(rowbuf + (size_t)kr * d + (lane + r * kWarp) * VEC) =
                      make_float4(acc[r][0], acc[r][1 % VEC], acc[r][2 % VEC], acc[r][3 % VEC]);
            }
            kr = __ffs(remaining) - 1;
            remaining &= remaining - 1;
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
              for (int v = 0; v < VEC; ++v) acc[r][v] = 0.f;
          }
#pragma unroll
          for (int r = 0; r < R; ++r) {
            acc[r][0] = fmaf(aa[u], xv[u][r].x, acc[r][0]); acc[r][1 % VEC] = fmaf(aa[u], xv[u][r].y, acc[r][1 % VEC]);
            acc[r][2 % VEC] = fmaf(aa[u], xv[u][r].z, acc[r][2 % VEC]); acc[r][3 % VEC] = fmaf(aa[u], xv[u][r].w, acc[r][3 % VEC]);
          }
        }
      }
    }
    if (kr >= 0) {
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (full || (lane + r * kWarp) * VEC < d)
          *reinterpret_cast<float4*>(rowbuf + (size_t)kr * d + (lane + r * kWarp) * VEC) =
              make_float4(acc[r][0], acc[r][1 % VEC], acc[r][2 % VEC], acc[r][3 % VEC]);
    }
    for (int r0 = 0; r0 < nrows; ++r0) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (((rowmask >> r0) & 1u) && (full || (lane + r * kWarp) * VEC < d))
          v = *reinterpret_cast<const float4*>(rowbuf + (size_t)r0 * d + (lane + r * kWarp) * VEC);
        acc[r][0] = v.x; acc[r][1 % VEC] = v.y; acc[r][2 % VEC] = v.z; acc[r][3 % VEC] = v.w;
      }
      finish_row<VEC, R>(A, row0 + r0, lane, alpha, beta, pol_stream, acc);
    }
    return;
  }

  for (int kr = 0; kr < nrows; ++kr) {
    int e_lo = 0, e_hi = cnt;
    if (split < 0) {
      const int rr = __shfl_sync(kFull, rec.y, kr);
      e_lo = (rr >> 20) & 0x3f;
      e_hi = (rr >> 26) & 0x3f;
    }
    float acc[R][VEC];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[r][v] = 0.f;

    if (STAGE == STAGE_CPASYNC || STAGE == STAGE_BULK) {
      if (full) {
#pragma unroll 4
        for (int e = e_lo; e < e_hi; ++e) {
          const float a = __shfl_sync(kFull, ab, e);
          const float* srow = stage + (size_t)e * d + lane * 4;
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const float4 v = *reinterpret_cast<const float4*>(srow + r * kWarp * 4);
            acc[r][0] = fmaf(a, v.x, acc[r][0]); acc[r][1 % VEC] = fmaf(a, v.y, acc[r][1 % VEC]);
            acc[r][2 % VEC] = fmaf(a, v.z, acc[r][2 % VEC]); acc[r][3 % VEC] = fmaf(a, v.w, acc[r][3 % VEC]);
          }
        }
      } else {
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
      }
    } else {
      // exact decomposition of the row's edge count into batches of U, U/2, ..., 1 rows in flight: no padded loads
      constexpr int U = (kDirectBatch / R) < 2 ? 2 : (kDirectBatch / R);
      const float* xl = A.x + (size_t)lane * VEC;
      int e = e_lo;
      while (e_hi - e >= U) { gather_batch<VEC, R, U>(xl, d, lane, full, j, ab, e, pol_keep, acc); e += U; }
      if (U >= 8 && e_hi - e >= 4) { gather_batch<VEC, R, 4>(xl, d, lane, full, j, ab, e, pol_keep, acc); e += 4; }
      if (U >= 4 && e_hi - e >= 2) { gather_batch<VEC, R, 2>(xl, d, lane, full, j, ab, e, pol_keep, acc); e += 2; }
      if (e_hi - e >= 1) { gather_batch<VEC, R, 1>(xl, d, lane, full, j, ab, e, pol_keep, acc); e += 1; }
      if (U == 2 && e_hi - e >= 1) gather_batch<VEC, R, 1>(xl, d, lane, full, j, ab, e, pol_keep, acc);
    }

    if (split < 0) {
      finish_row<VEC, R>(A, row0 + kr, lane, alpha, beta, pol_stream, acc);
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
    // slot order is fixed => deterministic; eight partial rows are requested before the first is added (the longest
    // row of ogbn-arxiv has ~330 partials: a dependent load per iteration would serialise ~100 us of L2 latency)
    constexpr int kPB = (R == 1) ? 8 : 2;
    int u = slot0;
    for (; u + kPB <= slot1; u += kPB) {
      float pv[kPB][R][VEC];
#pragma unroll
      for (int b = 0; b < kPB; ++b) {
        const float* pu = A.split_partial + (size_t)(u + b) * d;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int c = (lane + r * kWarp) * VEC;
          if (VEC == 4) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < d) v = __ldcg(reinterpret_cast<const float4*>(pu + c));
            pv[b][r][0] = v.x; pv[b][r][1 % VEC] = v.y; pv[b][r][2 % VEC] = v.z; pv[b][r][3 % VEC] = v.w;
          } else {
            pv[b][r][0] = c < d ? __ldcg(pu + c) : 0.f;
          }
        }
      }
#pragma unroll
      for (int b = 0; b < kPB; ++b)
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int v = 0; v < VEC; ++v) acc[r][v] += pv[b][r][v];
    }
    for (; u < slot1; ++u) {
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
    finish_row<VEC, R>(A, row0, lane, alpha, beta, pol_stream, acc);
  }
}

// ---- launchers ------------------------------------------------------------------------------------------------------------
// Which staging the edge kernel uses for d % 4 == 0.  GRAND_STAGE=direct|cpasync|bulk overrides (A/B measurements).
static int stage_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("GRAND_STAGE");
    mode = STAGE_DIRECT;
    if (e != nullptr) {
      if (!strcmp(e, "direct")) mode = STAGE_DIRECT;
      else if (!strcmp(e, "cpasync")) mode = STAGE_CPASYNC;
      else if (!strcmp(e, "bulk")) mode = STAGE_BULK;
      else if (!strcmp(e, "rowbuf")) mode = STAGE_ROWBUF;
    }
  }
  return mode;
}

template <int MODE, int VEC, int R, int STAGE>
static int launch_rhs_inst(const RhsArgs& A, cudaStream_t stream) {
  int grid = (A.tile_end - A.tile_begin + kRhsWarps - 1) / kRhsWarps;
  if (grid <= 0) return GRAND_OK;
  size_t smem = 0;
  if (STAGE == STAGE_CPASYNC || STAGE == STAGE_BULK) smem = (size_t)kRhsWarps * ((size_t)GRAND_TILE * A.d * sizeof(float) + 16);
  if (STAGE == STAGE_ROWBUF) smem = (size_t)kRhsWarps * GRAND_MAX_PACKED_ROWS * A.d * sizeof(float);
  auto kern = rhs_window_kernel<MODE, VEC, R, STAGE>;
  if (smem > 48 * 1024)   // per device and cheap: set on every launch that needs it
    GRAND_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, kRhsThreads, smem, stream>>>(A);
  GRAND_CUDA_OK(cudaGetLastError());
  count_launch();
  return GRAND_OK;
}

template <int MODE, int VEC, int R>
static int launch_rhs_stage(const RhsArgs& A, cudaStream_t stream) {
  // staging needs 16-byte rows and a window buffer that leaves room for >= 2 blocks per SM
  const bool can_stage = VEC == 4 && (size_t)kRhsWarps * GRAND_TILE * A.d * sizeof(float) <= 100 * 1024;
  int mode = stage_mode();
  if (mode == STAGE_ROWBUF) {   // row buffer: 8 rows per warp; keep >= 6 blocks per SM resident
    if (VEC != 4 || (size_t)kRhsWarps * GRAND_MAX_PACKED_ROWS * A.d * sizeof(float) > 32 * 1024) mode = STAGE_DIRECT;
  } else if (!can_stage) {
    mode = STAGE_DIRECT;
  }
  if (VEC == 4 && mode == STAGE_ROWBUF)
    return launch_rhs_inst<MODE, VEC, R, VEC == 4 ? STAGE_ROWBUF : STAGE_DIRECT>(A, stream);
  if (VEC == 4 && mode == STAGE_BULK) return launch_rhs_inst<MODE, VEC, R, VEC == 4 ? STAGE_BULK : STAGE_DIRECT>(A, stream);
  if (VEC == 4 && mode == STAGE_CPASYNC)
    return launch_rhs_inst<MODE, VEC, R, VEC == 4 ? STAGE_CPASYNC : STAGE_DIRECT>(A, stream);
  return launch_rhs_inst<MODE, VEC, R, STAGE_DIRECT>(A, stream);
}

template <int MODE>
static int launch_rhs_mode_t(const RhsArgs& A, cudaStream_t stream) {
  const int d = A.d;
  if (A.tile_end <= A.tile_begin) return GRAND_OK;
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

int launch_rhs_mode(int mode, const RhsArgs& A, cudaStream_t stream) {
  if (mode == MODE_FUSED) return launch_rhs_mode_t<MODE_FUSED>(A, stream);
  if (mode == MODE_PRECOMP) return launch_rhs_mode_t<MODE_PRECOMP>(A, stream);
  return launch_rhs_mode_t<MODE_WEIGHT>(A, stream);
}

int launch_split_stats(const RhsArgs& A, cudaStream_t stream) {
  if (A.g.n_split_tiles == 0) return GRAND_OK;
  split_tile_stats_kernel<<<(A.g.n_split_tiles + 7) / 8, 256, 0, stream>>>(A.g, A.a, A.q, A.k, A.split_stats);
  GRAND_CUDA_OK(cudaGetLastError());
  count_launch();
  split_row_merge_kernel<<<(A.g.n_split_rows + 7) / 8, 256, 0, stream>>>(A.g, A.a.heads, A.split_stats, A.row_stats);
  GRAND_CUDA_OK(cudaGetLastError());
  count_launch();
  return GRAND_OK;
}

}  // namespace grand
