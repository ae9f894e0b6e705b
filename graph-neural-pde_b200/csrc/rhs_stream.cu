// The dominant kernel, second design: PERSISTENT blocks, every warp walks a list of work units.
//
//   unit kinds (graph_plan.py builds the list, longest first):
//     long    a segment of <= seg_edges CSR edges of a row with more than seg_edges in-edges; the row's softmax statistics
//             come from a one-block-per-row pre-pass, the segment publishes a partial sum, the last segment to arrive adds
//             the partials in slot order (deterministic) and runs the row's epilogue
//     medium  one whole row with GRAND_TILE < in-edges <= seg_edges: one warp, two passes over the row's K rows
//             (statistics, then weights + aggregation); no global scratch, no atomics
//     window  <= GRAND_TILE edges of several whole short rows, one edge per lane (the packed windows of the plan)
//   unit u goes to block u mod gridDim; inside a block the warps take the block's units in order from a shared-memory
//   counter, so the whole grid sweeps the (row-ordered) window list as one moving band: the streamed epilogue operands
//   of neighbouring windows are touched at about the same time and the warps of a block balance dynamically.
//
//   (csrc/rhs_ring.cu is this kernel with the neighbour rows staged through shared memory by cp.async.bulk; it takes the
//   attention modes for rows of <= 512 bytes, this file the rest and the fixed-weight mode.)
//   Per unit: per-edge weights abar_e (MODE_FUSED: scaled_dot scores, row softmax, head mean -- reference
//   src/function_transformer_attention.py:190-213, :34; MODE_PRECOMP / MODE_WEIGHT as rhs_window.cu), then per row ONE
//   round trip: the row's epilogue operands (x_i, x0_i, stage-combine terms) and up to 8 neighbour rows are requested
//   together, accumulated in CSR order (torch_sparse.spmm, :35), then residual alpha (ax - x) + beta x0 and the
//   Runge-Kutta stage combine.  The next unit's header and lane records are requested before the current unit is
//   processed.  Neighbour rows go straight from L2/L1 into registers: measured on B200 (profiles/r01*, r02*), staging
//   512-byte rows through shared memory doubles the load on the L1/shared data pipe and takes the L1 capacity the
//   gathers live on; what TMA is used for on this path is the contiguous tile traffic of the projection.
#include <stdlib.h>
#include <string.h>

#include "edge_common.cuh"

namespace grand {

#ifndef GRAND_EPI_COEF_ONCE
#define GRAND_EPI_COEF_ONCE 1   // 0: one test per coefficient (measured 2-4 us slower per launch, profiles/r02v_ab.txt)
#endif
#ifndef GRAND_STREAM_THREADS
#define GRAND_STREAM_THREADS 320
#endif
#ifndef GRAND_STREAM_MINBLOCKS
#define GRAND_STREAM_MINBLOCKS 2
#endif
#ifndef GRAND_STREAM_BATCH
#define GRAND_STREAM_BATCH 8      // neighbour rows in flight per warp (divided by R for wide states)
#endif
#ifndef GRAND_STREAM_FAST_DIV
#define GRAND_STREAM_FAST_DIV 1
#endif
#ifndef GRAND_STREAM_FAST_EXP
#define GRAND_STREAM_FAST_EXP 0
#endif
constexpr int kStreamThreads = GRAND_STREAM_THREADS;
constexpr int kStreamWarps = kStreamThreads / kWarp;
constexpr int kStatsThreads = 256;

__device__ __forceinline__ float stream_score(const AttnView& a, const float* __restrict__ qh,
                                              const float* __restrict__ kh, bool vec4) {
  const float s = vec4 ? head_dot<true>(qh, kh, a.dk) : head_dot<false>(qh, kh, a.dk);
  return a.inv_sqrt_dk_exact != 0.f ? s * a.inv_sqrt_dk_exact : s / a.sqrt_dk;
}

// a_eh = exp(s - max) / (sum + 1e-16)  (PyG softmax, reference :213).  The quotient is taken with the hardware reciprocal
// (2 ulp, measured 2e-7 of the output against the IEEE division and 11 us per launch faster on ogbn-arxiv); the
// exponential stays expf unless GRAND_STREAM_FAST_EXP asks for ex2.approx (absolute error <= 3e-7 of the row maximum).
// a 64-byte K / Q row (16 floats) as four float4: two 256-bit loads (sm_100a LDG.E.256) -- half the load instructions
// and, for a lane-per-edge gather that touches 32 different lines, half the L1 wavefronts of four 128-bit loads
__device__ __forceinline__ void load_row16(const float* p, float4 (&v)[4]) {
  const float8 a = ldg8(p), b = ldg8(p + 8);
  v[0] = a.lo; v[1] = a.hi; v[2] = b.lo; v[3] = b.hi;
}
__device__ __forceinline__ float attn_div(float e, float l) {
#if GRAND_STREAM_FAST_DIV
  return __fdividef(e, l + 1e-16f);
#else
  return e / (l + 1e-16f);
#endif
}
__device__ __forceinline__ float attn_exp(float x) {
#if GRAND_STREAM_FAST_EXP
  return __expf(x);
#else
  return expf(x);
#endif
}

// segmented all-reduce over contiguous lane segments; `room` = hi - lane (lanes of my segment at or above me)
__device__ __forceinline__ float sseg_max(float v, int room, int lo) {
#pragma unroll
  for (int o = 1; o < kWarp; o <<= 1) {
    const float t = __shfl_down_sync(kFull, v, o);
    if (o < room) v = fmaxf(v, t);
  }
  return __shfl_sync(kFull, v, lo);
}
__device__ __forceinline__ float sseg_sum(float v, int room, int lo) {
#pragma unroll
  for (int o = 1; o < kWarp; o <<= 1) {
    const float t = __shfl_down_sync(kFull, v, o);
    if (o < room) v += t;
  }
  return __shfl_sync(kFull, v, lo);
}

#ifndef GRAND_STREAM_DEVICE_ONLY
// ---- long rows: softmax statistics (max, sum-exp) per head of every long SEGMENT, one block per segment; the segments
//      of a row merge their row's pairs themselves (<= deg / seg_edges pairs per head) --------------------------------
__global__ void __launch_bounds__(kStatsThreads)
lseg_stats_kernel(GraphView g, AttnView a, const float* __restrict__ q, const float* __restrict__ k,
                  float* __restrict__ lseg_stats) {
  __shared__ float sm[kStatsThreads / kWarp][GRAND_MAX_HEADS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int unit = blockIdx.x;
  pdl_launch_dependents();
  const int4 desc = __ldg(g.seg_desc + unit);
  pdl_wait();                              // Q / K come from the projection launched just before
  const int row = desc.x, beg = desc.y, cnt = desc.z;
  const int H = a.heads;
  const bool vec4 = ((a.dk & 3) == 0) && ((a.ad_eff & 3) == 0);
  const float* qi = q + (size_t)row * a.ad_eff;
  float m[GRAND_MAX_HEADS], l[GRAND_MAX_HEADS];
#pragma unroll
  for (int h = 0; h < GRAND_MAX_HEADS; ++h) { m[h] = -INFINITY; l[h] = 0.f; }
  for (int c = threadIdx.x; c < cnt; c += kStatsThreads) {        // one trip for seg_edges <= kStatsThreads
    const int p = beg + c;
    const float* kj = k + (size_t)__ldg(g.col + p) * a.ad_eff;
    const float ew = a.edge_weight_csr != nullptr ? __ldg(a.edge_weight_csr + p) : 1.f;
#pragma unroll
    for (int h = 0; h < GRAND_MAX_HEADS; ++h) {
      if (h < H) {
        const float s = stream_score(a, qi + h * a.dk, kj + h * a.dk, vec4) * ew;
        const float mn = fmaxf(m[h], s);
        l[h] = l[h] * expf(m[h] - mn) + expf(s - mn);     // m = -inf, l = 0 on the first visit: exp(-inf) = 0
        m[h] = mn;
      }
    }
  }
  // block max per head, one rescale, block sum
#pragma unroll
  for (int h = 0; h < GRAND_MAX_HEADS; ++h) {
    if (h < H) {
      const float wm = warp_max(m[h]);
      if (lane == 0) sm[warp][h] = wm;
    }
  }
  __syncthreads();
  float mm[GRAND_MAX_HEADS];
#pragma unroll
  for (int h = 0; h < GRAND_MAX_HEADS; ++h) {
    mm[h] = -INFINITY;
    if (h < H) {
#pragma unroll
      for (int w = 0; w < kStatsThreads / kWarp; ++w) mm[h] = fmaxf(mm[h], sm[w][h]);
    }
  }
  __syncthreads();
#pragma unroll
  for (int h = 0; h < GRAND_MAX_HEADS; ++h) {
    if (h < H) {
      const float e = warp_sum(m[h] == -INFINITY ? 0.f : l[h] * expf(m[h] - mm[h]));
      if (lane == 0) sm[warp][h] = e;
    }
  }
  __syncthreads();
#pragma unroll
  for (int h = 0; h < GRAND_MAX_HEADS; ++h) {
    if (h < H && threadIdx.x == h) {
      float ll = 0.f;
#pragma unroll
      for (int w = 0; w < kStatsThreads / kWarp; ++w) ll += sm[w][h];
      lseg_stats[((size_t)unit * H + h) * 2 + 0] = mm[h];
      lseg_stats[((size_t)unit * H + h) * 2 + 1] = ll;
    }
  }
}

#endif  // GRAND_STREAM_DEVICE_ONLY

// ---- epilogue operands of one row, requested BEFORE the row's neighbours are gathered -----------------------------------
// Named members, every access statically indexed and written under the same warp-uniform conditions it is read under:
// the operands stay in registers without the zero-initialisation a conditionally written array needs.
template <int R>
struct EpiRegs {
  float4 xi[R], s0[R], t0[R], t1[R], t2[R], t3[R];
};

template <bool FULL, int R>
__device__ __forceinline__ void epi_load(const RhsArgs& A, unsigned rowoff, int lane, uint64_t pol_stream, EpiRegs<R>& E) {
  // rowoff = row * d (elements): fits 32 bits (checked by the launcher)
  const int nt = A.ep.y_out != nullptr ? A.ep.n_terms : 0;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int c = (lane + r * kWarp) * 4;
    if (FULL || c < A.d) {
      const unsigned o = rowoff + c;
      E.xi[r] = ldg4(A.x + o);
      if (A.ep.x0 != nullptr) E.s0[r] = ld_stream4_hint(A.ep.x0 + o, pol_stream);
      if (nt > 0) E.t0[r] = ld_stream4_hint(A.ep.term[0] + o, pol_stream);
      if (nt > 1) E.t1[r] = ld_stream4_hint(A.ep.term[1] + o, pol_stream);
      if (nt > 2) E.t2[r] = ld_stream4_hint(A.ep.term[2] + o, pol_stream);
      if (nt > 3) E.t3[r] = ld_stream4_hint(A.ep.term[3] + o, pol_stream);
    }
  }
}

#define GRAND_FMA4(y, c, v) \
  { fma2(y.x, y.y, c, v.x, v.y); fma2(y.z, y.w, c, v.z, v.w); }

// residual + RK stage combine (same arithmetic, in the same order, as finish_row of rhs_window.cu)
template <bool FULL, int R>
__device__ __forceinline__ void epi_finish(const RhsArgs& A, unsigned rowoff, int lane, float alpha, float beta,
                                           uint64_t pol_stream, const float (&acc)[R][4], const EpiRegs<R>& E) {
  const int nt = A.ep.y_out != nullptr ? A.ep.n_terms : 0;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int c = (lane + r * kWarp) * 4;
    if (!(FULL || c < A.d)) continue;
    const unsigned o = rowoff + c;
    float4 f;
    f.x = alpha * (acc[r][0] - E.xi[r].x); f.y = alpha * (acc[r][1] - E.xi[r].y);     // f = alpha (ax - x)
    f.z = alpha * (acc[r][2] - E.xi[r].z); f.w = alpha * (acc[r][3] - E.xi[r].w);
    if (A.ep.x0 != nullptr) {                                                         // + beta x0   (add_source)
      f.x += beta * E.s0[r].x; f.y += beta * E.s0[r].y; f.z += beta * E.s0[r].z; f.w += beta * E.s0[r].w;
    }
    if (A.ep.f_out != nullptr) st_stream4_hint(A.ep.f_out + o, f, pol_stream);
    if (A.ep.y_out != nullptr) {
#if GRAND_EPI_COEF_ONCE
      // ONE test of "coefficients on the device?" per row instead of one per coefficient
      float cf, c0, c1, c2, c3;
      if (A.ep.coef_dev != nullptr) {
        cf = __ldg(A.ep.coef_dev); c0 = __ldg(A.ep.coef_dev + 1); c1 = __ldg(A.ep.coef_dev + 2);
        c2 = __ldg(A.ep.coef_dev + 3); c3 = __ldg(A.ep.coef_dev + 4);
      } else {
        cf = A.ep.coef_f; c0 = A.ep.coef[0]; c1 = A.ep.coef[1]; c2 = A.ep.coef[2]; c3 = A.ep.coef[3];
      }
      float4 y = make_float4(cf * f.x, cf * f.y, cf * f.z, cf * f.w);
      if (nt > 0) GRAND_FMA4(y, c0, E.t0[r]);
      if (nt > 1) GRAND_FMA4(y, c1, E.t1[r]);
      if (nt > 2) GRAND_FMA4(y, c2, E.t2[r]);
      if (nt > 3) GRAND_FMA4(y, c3, E.t3[r]);
#else
      const float cf = epi_coef_f(A.ep);
      float4 y = make_float4(cf * f.x, cf * f.y, cf * f.z, cf * f.w);
      if (nt > 0) GRAND_FMA4(y, epi_coef(A.ep, 0), E.t0[r]);
      if (nt > 1) GRAND_FMA4(y, epi_coef(A.ep, 1), E.t1[r]);
      if (nt > 2) GRAND_FMA4(y, epi_coef(A.ep, 2), E.t2[r]);
      if (nt > 3) GRAND_FMA4(y, epi_coef(A.ep, 3), E.t3[r]);
#endif
      for (int t = 4; t < nt; ++t) {                               // dopri5 stages have up to 7 terms
        const float ct = epi_coef(A.ep, t);
        const float4 v = ld_stream4_hint(A.ep.term[t] + o, pol_stream);
        GRAND_FMA4(y, ct, v);
      }
      // y_out is the next stage's gather source: keep it cacheable
      *reinterpret_cast<float4*>(A.ep.y_out + o) = y;
      if (A.push.ptr != nullptr) {
        // node partition: the same row straight into the halo of every peer whose edges reference it (plain stores over
        // NVLink through the peer's arena mapping); the exchange rides along with the rest of the launch
        const int row = rowoff / (unsigned)A.d;
        const int pb = __ldg(A.push.ptr + row), pe = __ldg(A.push.ptr + row + 1);
        for (int t = pb; t < pe; ++t) {
          const int2 dst = __ldg(A.push.dst + t);
          *reinterpret_cast<float4*>(A.push.out[dst.x] + (size_t)dst.y * A.d + c) = y;
        }
      }
    }
  }
}

// ---- exactly N neighbour rows (edges e .. e+N-1 of the window, e + N <= 32) in flight, then accumulated in edge order ----
template <bool FULL, int R, int N>
__device__ __forceinline__ void gather_n(const float* __restrict__ xl, int d, int lane, unsigned joff, float ab, int e,
                                         uint64_t pol_keep, float (&acc)[R][4]) {
  float aa[N];
  float4 xv[N][R];
#pragma unroll
  for (int u = 0; u < N; ++u) {
    const unsigned off = __shfl_sync(kFull, joff, e + u);
    aa[u] = __shfl_sync(kFull, ab, e + u);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (FULL) {
        xv[u][r] = ldg4_hint(xl + off + r * kWarp * 4, pol_keep);
      } else {
        xv[u][r] = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((lane + r * kWarp) * 4 < d) xv[u][r] = ldg4_hint(xl + off + r * kWarp * 4, pol_keep);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < N; ++u)
#pragma unroll
    for (int r = 0; r < R; ++r) {
      fma2(acc[r][0], acc[r][1], aa[u], xv[u][r].x, xv[u][r].y);
      fma2(acc[r][2], acc[r][3], aa[u], xv[u][r].z, xv[u][r].w);
    }
}

// edges [e_lo, e_hi) of the window: full batches of NB, then ONE batch of exactly the remainder (no padded slots, one
// round trip per batch)
template <bool FULL, int R, int NB>
__device__ __forceinline__ void gather_range(const float* __restrict__ xl, int d, int lane, unsigned joff, float ab,
                                             int e_lo, int e_hi, uint64_t pol_keep, float (&acc)[R][4]) {
  int e = e_lo;
  for (; e + NB <= e_hi; e += NB) gather_n<FULL, R, NB>(xl, d, lane, joff, ab, e, pol_keep, acc);
  const int rem = e_hi - e;
  if (NB > 4 && rem >= 4) {
    if (rem == 4) gather_n<FULL, R, 4>(xl, d, lane, joff, ab, e, pol_keep, acc);
    else if (rem == 5) gather_n<FULL, R, 5>(xl, d, lane, joff, ab, e, pol_keep, acc);
    else if (rem == 6) gather_n<FULL, R, 6>(xl, d, lane, joff, ab, e, pol_keep, acc);
    else gather_n<FULL, R, (NB > 7 ? 7 : 1)>(xl, d, lane, joff, ab, e, pol_keep, acc);
  } else if (rem > 0) {
    if (rem == 1) gather_n<FULL, R, 1>(xl, d, lane, joff, ab, e, pol_keep, acc);
    else if (rem == 2) gather_n<FULL, R, 2>(xl, d, lane, joff, ab, e, pol_keep, acc);
    else gather_n<FULL, R, (NB > 3 ? 3 : 1)>(xl, d, lane, joff, ab, e, pol_keep, acc);
  }
}

// d_k = 4, 4 heads: scores of the lane's edge from its Q / K rows (four float4 each), row softmax over the lanes
// [lo, lo + room), head mean
__device__ __forceinline__ float dk4_weight(const float4 (&qv)[4], const float4 (&kv)[4], bool valid, float ew, int room,
                                            int lo) {
  float s[4], m[4], ex[4], l[4];
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    float t = 0.f;
    t = fmaf(qv[h].x, kv[h].x, t); t = fmaf(qv[h].y, kv[h].y, t);
    t = fmaf(qv[h].z, kv[h].z, t); t = fmaf(qv[h].w, kv[h].w, t);
    s[h] = valid ? (t * 0.5f) * ew : -INFINITY;   // / sqrt(4): exact; ew: reweight_attention :208-209
  }
#pragma unroll
  for (int h = 0; h < 4; ++h) m[h] = sseg_max(s[h], room, lo);
#pragma unroll
  for (int h = 0; h < 4; ++h) ex[h] = valid ? attn_exp(s[h] - m[h]) : 0.f;
#pragma unroll
  for (int h = 0; h < 4; ++h) l[h] = sseg_sum(ex[h], room, lo);
  float ab = 0.f;
#pragma unroll
  for (int h = 0; h < 4; ++h) ab += attn_div(ex[h], l[h]);
  return ab * 0.25f;                               // / H, exact
}

// ---- per-edge weight of the lane's edge inside a packed window ----------------------------------------------------------
template <int MODE>
__device__ __forceinline__ float window_weight(const RhsArgs& A, int lane, int i, int j, int p, bool valid, int lo, int hi) {
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
  } else {
    const bool vec4 = ((A.a.dk & 3) == 0) && ((A.a.ad_eff & 3) == 0);
    const float* qi = A.q + (size_t)i * A.a.ad_eff;
    const float* kj = A.k + (size_t)j * A.a.ad_eff;
    const float ew = (valid && A.a.edge_weight_csr != nullptr) ? __ldg(A.a.edge_weight_csr + p) : 1.f;
    const int H = A.a.heads;
    const int room = hi - lane;
    if (A.a.dk == 4 && vec4 && H == 4) {
      // d_k = 4, 4 heads (ogbn-arxiv: ad 16): a head is ONE float4 of the Q / K row; all 8 loads first
      float4 qv[4], kv[4];
      load_row16(qi, qv);                    // idle lanes read row0's Q row and K row 0: harmless, no predicates
      load_row16(kj, kv);
      ab = dk4_weight(qv, kv, valid, ew, room, lo);
    } else {
      for (int h = 0; h < H; ++h) {
        float s = -INFINITY;
        if (valid) s = stream_score(A.a, qi + h * A.a.dk, kj + h * A.a.dk, vec4) * ew;
        const float m = sseg_max(s, room, lo);
        const float e = valid ? attn_exp(s - m) : 0.f;
        const float l = sseg_sum(e, room, lo);
        ab += attn_div(e, l);
      }
      ab = ab / (float)H;
    }
  }
  return ab;
}

// ---- a packed window: several whole short rows, one edge per lane ---------------------------------------------------------
template <int MODE, bool FULL, int R>
__device__ __forceinline__ void process_window(const RhsArgs& A, const int4 hdr, const int2 rec, int lane, float alpha,
                                               float beta, uint64_t pol_keep, uint64_t pol_stream) {
  constexpr int NB = (GRAND_STREAM_BATCH / R) < 2 ? 2 : (GRAND_STREAM_BATCH / R);
  const int d = A.d;
  const int row0 = hdr.x, beg = hdr.y, nrows = hdr.z & 0xff, cnt = hdr.z >> 8;
  const int j = rec.x;
  const int kk = rec.y & 0xff, lo = (rec.y >> 8) & 0x3f, hi = (rec.y >> 14) & 0x3f;
  const bool valid = lane < cnt;
  const float ab = window_weight<MODE>(A, lane, row0 + kk, j, beg + lane, valid, lo, hi);
  const unsigned joff = (unsigned)j * (unsigned)d;
  const float* xl = A.x + lane * 4;
  unsigned rowoff = (unsigned)row0 * (unsigned)d;
  for (int kr = 0; kr < nrows; ++kr, rowoff += d) {
    const int rr = __shfl_sync(kFull, rec.y, kr);
    const int e_lo = (rr >> 20) & 0x3f, e_hi = (rr >> 26) & 0x3f;
    EpiRegs<R> E;
    epi_load<FULL, R>(A, rowoff, lane, pol_stream, E);
    float acc[R][4];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int v = 0; v < 4; ++v) acc[r][v] = 0.f;
    gather_range<FULL, R, NB>(xl, d, lane, joff, ab, e_lo, e_hi, pol_keep, acc);
    epi_finish<FULL, R>(A, rowoff, lane, alpha, beta, pol_stream, acc, E);
  }
}

// warp merge of per-lane running (max, sum-exp) pairs: max first, ONE rescale, then the sum
__device__ __forceinline__ void merge_pairs(float& m, float& l) {
  const float mm = warp_max(m);
  l = (m == -INFINITY) ? 0.f : l * expf(m - mm);
  l = warp_sum(l);
  m = mm;
}

// where the neighbour rows of a 32-edge chunk come from: NoRing = straight into registers (gather_range, this file);
// csrc/rhs_ring.cu passes its per-warp shared-memory stage fed by cp.async.bulk
struct NoRing {
  static constexpr bool enabled = false;
  __device__ __forceinline__ void issue(const float*, int, int, bool, int, int, uint64_t) {}
  __device__ __forceinline__ void wait() {}
  template <bool FULL>
  __device__ __forceinline__ void accumulate(float, int, int, int, int, float (&)[4]) {}
};

// ---- a whole medium row or a segment of a long row: cnt edges of ONE row from CSR position beg ----------------------------
template <int MODE, bool FULL, int R, int HM, bool STATS = false, class RING = NoRing>
__device__ __forceinline__ void process_segment(const RhsArgs& A, const int4 desc, int unit, int lane, float alpha,
                                                float beta, uint64_t pol_keep, uint64_t pol_stream, RING& ring) {
  constexpr int NB = (GRAND_STREAM_BATCH / R) < 2 ? 2 : (GRAND_STREAM_BATCH / R);
  const int d = A.d;
  const int row = desc.x, beg = desc.y, cnt = desc.z, split = desc.w;
  const int H = A.a.heads;
  const bool vec4 = ((A.a.dk & 3) == 0) && ((A.a.ad_eff & 3) == 0);
  const bool dk4 = MODE == MODE_FUSED && A.a.dk == 4 && vec4 && H == 4;
  const float* qi = MODE == MODE_FUSED ? A.q + (size_t)row * A.a.ad_eff : nullptr;
  float m[HM], l[HM];
#pragma unroll
  for (int h = 0; h < HM; ++h) { m[h] = -INFINITY; l[h] = 0.f; }
  if (MODE == MODE_FUSED) {
    if (split >= 0 && !STATS) {
      // the row's softmax statistics: merge the (max, sum-exp) pairs of all its segments (pre-pass lseg_stats_kernel)
      const int slot0 = __ldg(A.g.lrow_slot0 + split), slot1 = __ldg(A.g.lrow_slot0 + split + 1);
#pragma unroll
      for (int h = 0; h < HM; ++h) {
        if (h < H) {
          for (int u = slot0 + lane; u < slot1; u += kWarp) {
            // written by the pre-pass kernel of an EARLIER launch: the read-only path is fine (and keeps this kernel's
            // register allocation free of spills: with ld.global.cg here ptxas spills 128 bytes of the window path);
            // only the one-launch integration (rhs_coop.cu) writes these pairs inside the same launch
#ifdef GRAND_COHERENT_LOADS
            const float2 pr = __ldcg(reinterpret_cast<const float2*>(A.lrow_stats) + (size_t)u * H + h);
#else
            const float2 pr = __ldg(reinterpret_cast<const float2*>(A.lrow_stats) + (size_t)u * H + h);
#endif
            const float mn = fmaxf(m[h], pr.x);
            l[h] = l[h] * expf(m[h] - mn) + pr.y * expf(pr.x - mn);
            m[h] = mn;
          }
          merge_pairs(m[h], l[h]);
        }
      }
    } else {
      // pass 1 over the row's K rows: running (max, sum-exp) per lane, merged across the warp
      for (int c = lane; c < cnt; c += kWarp) {
        const int p = beg + c;
        const float* kj = A.k + (size_t)__ldg(A.g.col + p) * A.a.ad_eff;
        const float ew = A.a.edge_weight_csr != nullptr ? __ldg(A.a.edge_weight_csr + p) : 1.f;
        if (dk4) {
          float4 kv[4];
          load_row16(kj, kv);
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const float4 qh = ldg4(qi + 4 * h);                     // same address in every lane: one broadcast line
            float t = 0.f;
            t = fmaf(qh.x, kv[h].x, t); t = fmaf(qh.y, kv[h].y, t);
            t = fmaf(qh.z, kv[h].z, t); t = fmaf(qh.w, kv[h].w, t);
            const float s = (t * 0.5f) * ew;
            const float mn = fmaxf(m[h], s);
            l[h] = l[h] * expf(m[h] - mn) + expf(s - mn);
            m[h] = mn;
          }
        } else {
#pragma unroll
          for (int h = 0; h < HM; ++h) {
            if (h < H) {
              const float s = stream_score(A.a, qi + h * A.a.dk, kj + h * A.a.dk, vec4) * ew;
              const float mn = fmaxf(m[h], s);
              l[h] = l[h] * expf(m[h] - mn) + expf(s - mn);
              m[h] = mn;
            }
          }
        }
      }
#pragma unroll
      for (int h = 0; h < HM; ++h)
        if (h < H) merge_pairs(m[h], l[h]);
    }
    if (STATS) {   // statistics unit of a long segment (rhs_coop.cu, in-kernel pre-pass): publish (max, sum-exp), nothing else
      if (lane == 0) {
#pragma unroll
        for (int h = 0; h < HM; ++h) {
          if (h < H) {
            __stcg(A.lrow_stats + ((size_t)unit * H + h) * 2 + 0, m[h]);
            __stcg(A.lrow_stats + ((size_t)unit * H + h) * 2 + 1, l[h]);
          }
        }
      }
      return;
    }
  }

  float acc[R][4];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[r][v] = 0.f;
  EpiRegs<R> E;
  const float* xl = A.x + lane * 4;
  const unsigned rowoff = (unsigned)row * (unsigned)d;
  int jn = lane < cnt ? __ldg(A.g.col + beg + lane) : 0;            // column of the lane's edge in the next chunk
  for (int c0 = 0; c0 < cnt; c0 += kWarp) {
    const int nvalid = cnt - c0 < kWarp ? cnt - c0 : kWarp;
    const bool valid = lane < nvalid;
    const int p = beg + c0 + lane;
    const int j = jn;
    if (RING::enabled) ring.issue(A.x, d, j, valid, nvalid, lane, pol_keep);   // the chunk's rows fly while it is scored
    if (c0 + kWarp < cnt) jn = (c0 + kWarp + lane < cnt) ? __ldg(A.g.col + p + kWarp) : 0;
    float ab = 0.f;
    if (MODE == MODE_WEIGHT) {
      if (valid) ab = __ldg(A.w_csr + p);
    } else if (MODE == MODE_PRECOMP) {
      if (valid) {
        const float gmax = A.a.norm_type == GRAND_NORM_SQUAREPLUS ? __ldg(A.gmax) : 0.f;
        const int seg = A.a.norm_idx == 0 ? row : j;
        float a_sum = 0.f;
        for (int h = 0; h < H; ++h) {
          const float sc = __ldg(A.scores + (size_t)p * H + h);
          const float ms = __ldg(A.seg_stats + ((size_t)seg * H + h) * 2 + 0);
          const float ls = __ldg(A.seg_stats + ((size_t)seg * H + h) * 2 + 1);
          a_sum += norm_num(A.a.norm_type, sc, A.a.norm_type == GRAND_NORM_SOFTMAX ? ms : gmax) / (ls + 1e-16f);
        }
        ab = a_sum / (float)H;
      }
    } else {
      const float* kj = A.k + (size_t)j * A.a.ad_eff;
      const float ew = (valid && A.a.edge_weight_csr != nullptr) ? __ldg(A.a.edge_weight_csr + p) : 1.f;
      if (dk4) {
        float4 kv[4];
        load_row16(kj, kv);                                         // idle lanes read K row 0: harmless
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const float4 qh = ldg4(qi + 4 * h);
          float t = 0.f;
          t = fmaf(qh.x, kv[h].x, t); t = fmaf(qh.y, kv[h].y, t);
          t = fmaf(qh.z, kv[h].z, t); t = fmaf(qh.w, kv[h].w, t);
          ab += attn_div(attn_exp((t * 0.5f) * ew - m[h]), l[h]);
        }
        ab = valid ? ab * 0.25f : 0.f;
      } else if (valid) {
#pragma unroll
        for (int h = 0; h < HM; ++h) {
          if (h < H) {
            const float s = stream_score(A.a, qi + h * A.a.dk, kj + h * A.a.dk, vec4) * ew;
            ab += attn_div(attn_exp(s - m[h]), l[h]);
          }
        }
        ab = ab / (float)H;
      }
    }
    if (split < 0 && c0 + kWarp >= cnt) epi_load<FULL, R>(A, rowoff, lane, pol_stream, E);   // with the last chunk's gathers
    if (RING::enabled) {
      ring.wait();
      ring.template accumulate<FULL>(ab, 0, nvalid, d, lane, acc[0]);
    } else {
      const unsigned joff = (unsigned)j * (unsigned)d;
      gather_range<FULL, R, NB>(xl, d, lane, joff, ab, 0, nvalid, pol_keep, acc);
    }
  }

  if (split < 0) {
    epi_finish<FULL, R>(A, rowoff, lane, alpha, beta, pol_stream, acc, E);
    return;
  }
  // ---- segment of a long row: publish the partial sum; the last segment of the row to arrive reduces and finishes ----
  float* part = A.lpartial + (size_t)unit * d;          // long segments are units [0, n_lunits): slot == unit
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int c = (lane + r * kWarp) * 4;
    if (FULL || c < d) __stcg(reinterpret_cast<float4*>(part + c), make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]));
  }
  __threadfence();
  __syncwarp();
  const int slot0 = __ldg(A.g.lrow_slot0 + split), slot1 = __ldg(A.g.lrow_slot0 + split + 1);
  unsigned old = 0;
  if (lane == 0) old = atomicInc(reinterpret_cast<unsigned*>(A.lcounter) + split, (unsigned)(slot1 - slot0 - 1));
  old = __shfl_sync(kFull, old, 0);                      // atomicInc wraps to 0 after the last arrival: self-resetting
  if ((int)old != slot1 - slot0 - 1) return;
  __threadfence();
  epi_load<FULL, R>(A, rowoff, lane, pol_stream, E);
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[r][v] = 0.f;
  constexpr int kPB = (R == 1) ? 8 : 2;
  int u = slot0;
  for (; u + kPB <= slot1; u += kPB) {
    float4 pv[kPB][R];
#pragma unroll
    for (int b = 0; b < kPB; ++b) {
      const float* pu = A.lpartial + (size_t)(u + b) * d;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int c = (lane + r * kWarp) * 4;
        pv[b][r] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (FULL || c < d) pv[b][r] = __ldcg(reinterpret_cast<const float4*>(pu + c));
      }
    }
#pragma unroll
    for (int b = 0; b < kPB; ++b)
#pragma unroll
      for (int r = 0; r < R; ++r) {
        acc[r][0] += pv[b][r].x; acc[r][1] += pv[b][r].y; acc[r][2] += pv[b][r].z; acc[r][3] += pv[b][r].w;
      }
  }
  for (; u < slot1; ++u) {
    const float* pu = A.lpartial + (size_t)u * d;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int c = (lane + r * kWarp) * 4;
      if (FULL || c < d) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(pu + c));
        acc[r][0] += v.x; acc[r][1] += v.y; acc[r][2] += v.z; acc[r][3] += v.w;
      }
    }
  }
  epi_finish<FULL, R>(A, rowoff, lane, alpha, beta, pol_stream, acc, E);
}

#ifndef GRAND_STREAM_DEVICE_ONLY
// ---- the kernel -------------------------------------------------------------------------------------------------------
template <int MODE, bool FULL, int R, int HM>
__global__ void __launch_bounds__(kStreamThreads, R == 1 ? GRAND_STREAM_MINBLOCKS : 1)
rhs_stream_kernel(const __grid_constant__ RhsArgs A) {
  __shared__ int s_next;
  if (threadIdx.x == 0) s_next = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  float alpha = __ldg(A.ep.alpha_train);
  if (!A.ep.no_alpha_sigmoid) alpha = 1.f / (1.f + expf(-alpha));   // torch.sigmoid
  const float beta = A.ep.x0 != nullptr ? __ldg(A.ep.beta_train) : 0.f;
  const uint64_t pol_keep = (A.l2_hints & 1) ? policy_evict_last() : policy_evict_normal();
  const uint64_t pol_stream = (A.l2_hints & 2) ? policy_evict_first() : policy_evict_normal();
  const int n_seg = A.g.n_lunits + A.g.n_munits;
  const int n_units = n_seg + (A.g.n_tiles - A.g.n_split_tiles);

  auto grab = [&]() -> int {
    int v = 0;
    if (lane == 0) v = atomicAdd(&s_next, 1);
    v = __shfl_sync(kFull, v, 0);
    return v * (int)gridDim.x + (int)blockIdx.x;
  };
  auto fetch = [&](int u, int4& h4, int2& rec) {
    if (u < n_seg) {
      h4 = __ldg(A.g.seg_desc + u);
    } else if (u < n_units) {
      const int t = A.g.n_split_tiles + (u - n_seg);
      h4 = __ldg(A.g.win_hdr + t);
      rec = __ldg(A.g.win_rec + (size_t)t * kWarp + lane);
    }
  };

  NoRing no_ring;
  pdl_launch_dependents();
  int u = grab();
  int4 h4 = make_int4(0, 0, 0, 0);
  int2 rec = make_int2(0, 0);
  fetch(u, h4, rec);
  pdl_wait();
  while (u < n_units) {
    const int un = grab();
    int4 h4n = make_int4(0, 0, 0, 0);
    int2 recn = make_int2(0, 0);
    fetch(un, h4n, recn);
    if (u < n_seg) process_segment<MODE, FULL, R, HM>(A, h4, u, lane, alpha, beta, pol_keep, pol_stream, no_ring);
    else process_window<MODE, FULL, R>(A, h4, rec, lane, alpha, beta, pol_keep, pol_stream);
    u = un;
    h4 = h4n;
    rec = recn;
  }
  if (A.push.ptr != nullptr) __threadfence_system();     // every row this thread delivered to a peer is visible there
}

// ---- launchers ------------------------------------------------------------------------------------------------------------
template <int MODE, bool FULL, int R>
static int launch_stream_inst(const RhsArgs& A, cudaStream_t stream) {
  const int n_units = A.g.n_lunits + A.g.n_munits + (A.g.n_tiles - A.g.n_split_tiles);
  if (n_units <= 0) return GRAND_OK;
  int grid = (n_units + kStreamWarps - 1) / kStreamWarps;
  const int cap = num_sms() * (R == 1 ? GRAND_STREAM_MINBLOCKS : 1);
  if (grid > cap) grid = cap;
  if (MODE == MODE_FUSED && A.a.heads > 4)
    GRAND_CUDA_OK(launch_pdl(rhs_stream_kernel<MODE, FULL, R, GRAND_MAX_HEADS>, dim3(grid), dim3(kStreamThreads), 0, stream, A));
  else
    GRAND_CUDA_OK(launch_pdl(rhs_stream_kernel<MODE, FULL, R, 4>, dim3(grid), dim3(kStreamThreads), 0, stream, A));
  GRAND_CUDA_OK(cudaGetLastError());
  count_launch();
  return GRAND_OK;
}

template <int MODE>
static int launch_stream_mode_t(const RhsArgs& A, cudaStream_t stream) {
  const int chunks = (A.d / 4 + kWarp - 1) / kWarp;
  const bool full = A.d == chunks * kWarp * 4;
  if (chunks <= 1) return full ? launch_stream_inst<MODE, true, 1>(A, stream) : launch_stream_inst<MODE, false, 1>(A, stream);
  if (chunks <= 2) return full ? launch_stream_inst<MODE, true, 2>(A, stream) : launch_stream_inst<MODE, false, 2>(A, stream);
  GRAND_REQUIRE(chunks <= 4, GRAND_EUNSUPPORTED, "state width %d > 512 is not built", A.d);
  return launch_stream_inst<MODE, false, 4>(A, stream);
}

// true when the persistent kernel can take this launch (16-byte rows, 32-bit row offsets, the plan carries segments)
bool stream_kernel_ok(const RhsArgs& A) {
  static int off = -1;
  if (off < 0) {
    const char* e = getenv("GRAND_EDGE");                 // GRAND_EDGE=window: A/B against the round-1 kernel
    off = (e != nullptr && !strcmp(e, "window")) ? 1 : 0;
  }
  if (off) return false;
  if ((A.d & 3) != 0 || A.d > 512) return false;
  if ((long long)A.g.n_cols * A.d >= (1ll << 32)) return false;
  if (A.g.n_split_tiles > 0 && (A.g.seg_desc == nullptr || A.g.n_lunits + A.g.n_munits == 0)) return false;
  if (A.g.n_lunits > 0 && (A.lpartial == nullptr || A.lcounter == nullptr || A.g.lrow_slot0 == nullptr)) return false;
  return true;
}

bool ring_kernel_ok(int mode, const RhsArgs& A);      // csrc/rhs_ring.cu: the same work list with TMA-staged neighbour rows
int launch_ring_mode(int mode, const RhsArgs& A, cudaStream_t stream);

int launch_stream_mode(int mode, const RhsArgs& A, cudaStream_t stream) {
  const bool ring = ring_kernel_ok(mode, A);
  if (mode == MODE_FUSED) {
    if (A.g.n_lunits > 0) {
      GRAND_REQUIRE(A.lrow_stats != nullptr, GRAND_EINVAL, "graph has long rows: workspace lrow_stats is null");
      GRAND_CUDA_OK(launch_pdl(lseg_stats_kernel, dim3(A.g.n_lunits), dim3(kStatsThreads), 0, stream, A.g, A.a, A.q, A.k,
                               A.lrow_stats));
      GRAND_CUDA_OK(cudaGetLastError());
      count_launch();
    }
    return ring ? launch_ring_mode(mode, A, stream) : launch_stream_mode_t<MODE_FUSED>(A, stream);
  }
  if (ring) return launch_ring_mode(mode, A, stream);
  if (mode == MODE_PRECOMP) return launch_stream_mode_t<MODE_PRECOMP>(A, stream);
  return launch_stream_mode_t<MODE_WEIGHT>(A, stream);
}

#endif  // GRAND_STREAM_DEVICE_ONLY

}  // namespace grand
