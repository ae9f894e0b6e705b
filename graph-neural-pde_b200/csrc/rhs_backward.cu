// Vector-Jacobian product of the fused GRAND-nl right-hand side (scaled_dot scores, row softmax, head mean): the edge
// phase of the backward pass.  With u_i = alpha * grad_f_i,
//     c_e      = u_i . x_j                          dL/d(abar_e)
//     ds_eh    = a_eh (c_e - sum_{e' in row} a_e'h c_e') / H          softmax + head-mean backward
//     dQ_i^h  += sum_e ds_eh k_j^h / sqrt(d_k)       (row-local: segmented lane reduction, plain store)
//     dK_j^h  += ds_eh q_i^h / sqrt(d_k)             (scatter by column: fire-and-forget red.add)
//     abar_e   = mean_h a_eh                         written in the caller's edge order for the transposed aggregation
// The dense tails (dX += dQ Wq + dK Wk, dWq = dQ^T X, ...) are plain GEMMs on the host side (cuBLAS through torch), the
// transposed aggregation alpha (Abar^T g - g) is the forward MODE_WEIGHT kernel on the transposed plan.
// The reference gets all of this from autograd over its ~30 forward ops; nothing here has a reference
// line to cite beyond the forward it differentiates: src/function_transformer_attention.py:38-53, :174-213.
//
// One warp per window, one edge per lane, as in the forward kernel; neighbour rows x_j and the window's grad rows are
// staged in shared memory with TMA bulk copies so that every lane can take the full d-wide dot product of ITS edge
// (skewed 128-bit shared loads, conflict-free for d = 128).
#include "edge_common.cuh"

namespace grand {

constexpr int kBwdThreads = 128;
constexpr int kBwdWarps = kBwdThreads / kWarp;
constexpr int kBwdMaxRows = 8;    // whole rows per packed window (graph_plan.py max_rows)

struct BwdArgs {
  GraphView g;
  AttnView a;
  int d;
  const float* x;        // [n_cols, d]
  const float* q;        // [n_rows, ad]
  const float* k;        // [n_cols, ad]
  const float* grad_f;   // [n_rows, d]
  const float* alpha_train;
  int no_alpha_sigmoid;
  const float* row_stats;   // [n_split_rows, H, 2] merged softmax statistics of long rows (forward pre-pass)
  float* row_dot;           // [n_split_rows, H] sum_e a_eh c_e of long rows (zero on entry)
  int phase;                // long rows need two passes: 0 accumulate row_dot, 1 use it; short rows only in phase 1
  const float* c_in;        // or null.  Non-null: dL/d(abar_e) is GIVEN ([n_edges], CSR order; frozen attention,
                            // AttODEblock) instead of being formed as alpha grad_f_i . x_j: no staging, x / grad_f unused
  float* abar_out;          // [n_edges] caller order, or null
  float* dq;                // [n_rows, ad]  (zero on entry)
  float* dk;                // [n_cols, ad]  (zero on entry)
};

__device__ __forceinline__ float bwd_score(const AttnView& a, const float* __restrict__ qh,
                                           const float* __restrict__ kh, bool vec4) {
  const float s = vec4 ? head_dot<true>(qh, kh, a.dk) : head_dot<false>(qh, kh, a.dk);
  return a.inv_sqrt_dk_exact != 0.f ? s * a.inv_sqrt_dk_exact : s / a.sqrt_dk;
}

__global__ void __launch_bounds__(kBwdThreads, 2) rhs_attention_bwd_kernel(const __grid_constant__ BwdArgs A) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int d = A.d;
  const int row_bytes = d * (int)sizeof(float);
  // phase 0 runs over the chunk tiles only
  const int n_tiles = A.phase == 0 ? A.g.n_split_tiles : A.g.n_tiles;
  const int t = blockIdx.x * kBwdWarps + warp;
  if (t >= n_tiles) return;

  const int4 hdr = __ldg(A.g.win_hdr + t);
  const int2 rec = __ldg(A.g.win_rec + (size_t)t * kWarp + lane);
  const int row0 = hdr.x, beg = hdr.y, nrows = hdr.z & 0xff, cnt = hdr.z >> 8, split = hdr.w;
  const int j = rec.x;
  const int kk = rec.y & 0xff, lo = (rec.y >> 8) & 0x3f, hi = (rec.y >> 14) & 0x3f;
  const bool valid = lane < cnt;
  const int i = row0 + kk;

  // ---- stage x_j (one row per lane) and the grad rows of the window's rows -----------------------------------------
  const int warp_bytes = (GRAND_TILE + kBwdMaxRows) * row_bytes + 16;
  unsigned char* base = smem_raw + (size_t)warp * warp_bytes;
  float* xs = reinterpret_cast<float*>(base);
  float* gs = xs + (size_t)GRAND_TILE * d;
  const uint32_t bar = smem_u32(base + (GRAND_TILE + kBwdMaxRows) * row_bytes);
  const uint64_t pol = policy_evict_normal();
  if (A.c_in == nullptr) {
    if (lane == 0) {
      mbar_init(bar, 1);
      mbar_expect_tx(bar, (uint32_t)((cnt + nrows) * row_bytes));
    }
    __syncwarp();
    if (valid) bulk_copy_g2s(smem_u32(xs) + lane * row_bytes, A.x + (size_t)j * d, (uint32_t)row_bytes, bar, pol);
    if (lane < nrows)
      bulk_copy_g2s(smem_u32(gs) + lane * row_bytes, A.grad_f + (size_t)(row0 + lane) * d, (uint32_t)row_bytes, bar, pol);
  }

  // ---- attention a_eh of the lane's edge, as in the forward ----------------------------------------------------------
  float alpha = __ldg(A.alpha_train);
  if (!A.no_alpha_sigmoid) alpha = 1.f / (1.f + expf(-alpha));
  const bool vec4 = ((A.a.dk & 3) == 0) && ((A.a.ad_eff & 3) == 0);
  const float* qi = A.q + (size_t)i * A.a.ad_eff;
  const float* kj = A.k + (size_t)j * A.a.ad_eff;
  const float ew = (valid && A.a.edge_weight_csr != nullptr) ? __ldg(A.a.edge_weight_csr + beg + lane) : 1.f;
  const int H = A.a.heads;
  const int dk = A.a.dk;
  const float inv_scale = A.a.inv_sqrt_dk_exact != 0.f ? A.a.inv_sqrt_dk_exact : 1.f / A.a.sqrt_dk;

  // ---- c_e = alpha * grad_f_i . x_j : every lane its own edge, skewed chunks --------------------------------------
  float c_e = 0.f;
  if (A.c_in != nullptr) {
    if (valid) c_e = __ldg(A.c_in + beg + lane);
  } else {
    mbar_wait(bar, 0);
  }
  if (A.c_in == nullptr && valid) {
    const int chunks = d >> 2;
    const float* xr = xs + (size_t)lane * d;
    const float* gr = gs + (size_t)kk * d;
    for (int tt = 0; tt < chunks; ++tt) {
      int cidx = lane + tt;
      cidx = cidx >= chunks ? cidx - chunks : cidx;
      cidx = cidx >= chunks ? cidx % chunks : cidx;
      const float4 xv = *reinterpret_cast<const float4*>(xr + cidx * 4);
      const float4 gv = *reinterpret_cast<const float4*>(gr + cidx * 4);
      c_e = fmaf(xv.x, gv.x, c_e); c_e = fmaf(xv.y, gv.y, c_e);
      c_e = fmaf(xv.z, gv.z, c_e); c_e = fmaf(xv.w, gv.w, c_e);
    }
    c_e *= alpha;
  }

  float abar = 0.f;
  for (int h = 0; h < H; ++h) {
    float s = -INFINITY;
    if (valid) s = bwd_score(A.a, qi + h * dk, kj + h * dk, vec4) * ew;
    float a_eh, tdot;
    if (split < 0) {
      const float m = seg_max(s, lane, lo, hi);
      const float e = valid ? expf(s - m) : 0.f;
      const float l = seg_sum(e, lane, lo, hi);
      a_eh = e / (l + 1e-16f);
      tdot = seg_sum(a_eh * c_e, lane, lo, hi);
    } else {
      const float m = __ldg(A.row_stats + ((size_t)split * H + h) * 2 + 0);
      const float l = __ldg(A.row_stats + ((size_t)split * H + h) * 2 + 1);
      a_eh = valid ? expf(s - m) / (l + 1e-16f) : 0.f;
      if (A.phase == 0) {   // long row, first pass: only the row-wide sum_e a_eh c_e
        const float part = warp_sum(a_eh * c_e);
        if (lane == 0) atomicAdd(A.row_dot + (size_t)split * H + h, part);
        continue;
      }
      tdot = __ldcg(A.row_dot + (size_t)split * H + h);
    }
    abar += a_eh;
    // d(score): softmax backward, head mean 1/H, reweight_attention factor, 1/sqrt(d_k)
    const float ds = valid ? a_eh * (c_e - tdot) / (float)H * ew * inv_scale : 0.f;
    for (int c = 0; c < dk; ++c) {
      const float kv = valid ? __ldg(kj + h * dk + c) : 0.f;
      const float qv = valid ? __ldg(qi + h * dk + c) : 0.f;
      // dK_j: scatter by column
      if (valid) atomicAdd(A.dk + (size_t)j * A.a.ad_eff + h * dk + c, ds * qv);
      // dQ_i: reduce over the lanes of the row
      float r;
      if (split < 0) {
        r = seg_sum(ds * kv, lane, lo, hi);
        if (valid && lane == lo) A.dq[(size_t)i * A.a.ad_eff + h * dk + c] = r;
      } else {
        r = warp_sum(ds * kv);
        if (lane == 0) atomicAdd(A.dq + (size_t)i * A.a.ad_eff + h * dk + c, r);
      }
    }
  }
  if (A.phase == 1 && valid && A.abar_out != nullptr) A.abar_out[__ldg(A.g.eid + beg + lane)] = abar / (float)H;
}

int make_attn_view(const grand_attention* a, AttnView* out);

int rhs_attention_backward_impl(const grand_graph* g, const grand_attention* a, const float* x, const float* grad_f,
                                const grand_workspace* ws, const float* alpha_train, int no_alpha_sigmoid,
                                float* row_dot, float* abar_out, float* dq, float* dk, GraphView gv,
                                cudaStream_t stream, const float* c_in) {
  BwdArgs A;
  A.c_in = c_in;
  A.g = gv;
  int rc = make_attn_view(a, &A.a);
  if (rc != GRAND_OK) return rc;
  GRAND_REQUIRE(A.a.score_type == GRAND_SCORE_SCALED_DOT && A.a.norm_type == GRAND_NORM_SOFTMAX && A.a.norm_idx == 0,
                GRAND_EUNSUPPORTED, "rhs_attention_backward: only scaled_dot + softmax + attention_norm_idx 0 is built");
  GRAND_REQUIRE((a->d & 3) == 0, GRAND_EUNSUPPORTED, "rhs_attention_backward: state width %d is not a multiple of 4",
                a->d);
  A.d = a->d;
  A.x = x; A.q = ws->q; A.k = ws->k; A.grad_f = grad_f;
  A.alpha_train = alpha_train; A.no_alpha_sigmoid = no_alpha_sigmoid;
  A.row_stats = ws->row_stats; A.row_dot = row_dot;
  A.abar_out = abar_out; A.dq = dq; A.dk = dk;
  const size_t smem = c_in != nullptr ? 0 : (size_t)kBwdWarps * ((size_t)(GRAND_TILE + kBwdMaxRows) * a->d * sizeof(float) + 16);
  GRAND_REQUIRE(smem <= 227 * 1024, GRAND_EUNSUPPORTED, "rhs_attention_backward: state width %d needs %zu B of shared "
                "memory per block", a->d, smem);
  if (smem > 48 * 1024)   // per device and cheap: set on every launch that needs it (no process-wide "already done" flag)
    GRAND_CUDA_OK(cudaFuncSetAttribute(rhs_attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (g->n_split_tiles > 0) {
    GRAND_REQUIRE(row_dot != nullptr && ws->row_stats != nullptr, GRAND_EINVAL,
                  "rhs_attention_backward: graph has long rows, row_dot / row_stats are null");
    A.phase = 0;
    rhs_attention_bwd_kernel<<<(g->n_split_tiles + kBwdWarps - 1) / kBwdWarps, kBwdThreads, smem, stream>>>(A);
    GRAND_CUDA_OK(cudaGetLastError());
    count_launch();
  }
  if (g->n_tiles > 0) {
    A.phase = 1;
    rhs_attention_bwd_kernel<<<(g->n_tiles + kBwdWarps - 1) / kBwdWarps, kBwdThreads, smem, stream>>>(A);
    GRAND_CUDA_OK(cudaGetLastError());
    count_launch();
  }
  return GRAND_OK;
}

}  // namespace grand
