// General attention pipeline (any scorer / normaliser / norm index) and the host-side implementation of the edge-phase
// entry points.  The dominant fused kernel lives in rhs_window.cu.
//
//   edge_scores_kernel        prods[E,H] in CSR order (+ global max for squareplus)   reference
//                             src/function_transformer_attention.py:190-209
//   segment_stats_kernel      per segment (row or column) max / normaliser sum        reference :210-213, src/utils.py:196-208
//   attention_out_kernel      attention[E,H] / prods[E,H] back in the caller's edge order, head-mean in CSR order
//   edge_values_to_csr_kernel frozen attention / edge weights -> CSR order            reference
//                             src/function_laplacian_diffusion.py:29-35
#include <stdlib.h>
#include <string.h>

#include "edge_common.cuh"

namespace grand {

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
  v.rowptr = g->rowptr; v.col = g->col; v.erow = g->erow; v.eid = g->eid;
  v.split_tile0 = g->split_tile0; v.colptr = g->colptr; v.csc_pos = g->csc_pos;
  v.win_hdr = reinterpret_cast<const int4*>(g->win_hdr);
  v.win_rec = reinterpret_cast<const int2*>(g->win_rec);
  v.n_lunits = g->n_lunits; v.n_munits = g->n_munits; v.n_lrows = g->n_lrows;
  v.seg_desc = reinterpret_cast<const int4*>(g->seg_desc);
  v.lrow_slot0 = g->lrow_slot0; v.lrow_row = g->lrow_row;
  return v;
}

int make_attn_view(const grand_attention* a, AttnView* out) {
  GRAND_REQUIRE(a->heads > 0 && a->ad > 0 && a->ad % a->heads == 0, GRAND_EINVAL,
                "Number of heads (%d) must be a factor of the dimension size (%d)", a->heads, a->ad);
  GRAND_REQUIRE(a->score_type >= 0 && a->score_type <= GRAND_SCORE_GAT, GRAND_EINVAL,
                "unknown score_type %d", a->score_type);
  GRAND_REQUIRE(a->norm_type == GRAND_NORM_SOFTMAX || a->norm_type == GRAND_NORM_SQUAREPLUS, GRAND_EINVAL,
                "unknown norm_type %d", a->norm_type);
  GRAND_REQUIRE(a->norm_idx == 0 || a->norm_idx == 1, GRAND_EINVAL, "attention_norm_idx must be 0 or 1");
  const bool needs_scalars = a->score_type == GRAND_SCORE_EXP_KERNEL || a->score_type == GRAND_SCORE_EXP_KERNEL2 ||
                             a->score_type == GRAND_SCORE_GAT;
  GRAND_REQUIRE(!needs_scalars || a->kernel_scalars != nullptr, GRAND_EINVAL,
                "exp_kernel / GAT scorers need kernel_scalars");
  out->ad = a->ad;
  out->ad_eff = a->score_type == GRAND_SCORE_EXP_KERNEL2 ? 2 * a->ad : a->ad;
  out->heads = a->heads;
  out->dk = a->ad / a->heads;
  out->score_type = a->score_type;
  out->norm_type = a->norm_type;
  out->norm_idx = a->norm_idx;
  out->sqrt_dk = (float)sqrt((double)out->dk);
  // d_k a power of 4: sqrt(d_k) is a power of two and s / sqrt(d_k) == s * (1 / sqrt(d_k)) bit for bit
  out->inv_sqrt_dk_exact = 0.f;
  for (int p = 1; p <= 1024; p *= 2)
    if (p * p == out->dk) out->inv_sqrt_dk_exact = 1.f / (float)p;
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
  v.coef_dev = ep->coef_dev;
  return v;
}

// delivery of y_out rows to the peers: y_out as every peer maps it = peer arena + (y_out - own arena)
static int fill_push(const grand_epilogue* ep, PushView* out) {
  memset(out, 0, sizeof(*out));
  const grand_peers* pr = ep->peers;
  if (pr == nullptr || ep->y_out == nullptr) return GRAND_OK;
  GRAND_REQUIRE(pr->world >= 1 && pr->world <= GRAND_MAX_PEERS && pr->rank >= 0 && pr->rank < pr->world, GRAND_EINVAL,
                "peers: world %d rank %d", pr->world, pr->rank);
  GRAND_REQUIRE(pr->push_ptr != nullptr && pr->push_dst != nullptr, GRAND_EINVAL, "peers: null push lists");
  const ptrdiff_t off = reinterpret_cast<const char*>(ep->y_out) - reinterpret_cast<const char*>(pr->arena[pr->rank]);
  GRAND_REQUIRE(off >= 0, GRAND_EINVAL, "peers: y_out does not lie in the arena");
  out->ptr = pr->push_ptr;
  out->dst = reinterpret_cast<const int2*>(pr->push_dst);
  for (int p = 0; p < pr->world; ++p) out->out[p] = reinterpret_cast<float*>(reinterpret_cast<char*>(pr->arena[p]) + off);
  return GRAND_OK;
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
  GRAND_REQUIRE(g->n_tiles == 0 || (g->rowptr && g->win_hdr && g->win_rec), GRAND_EINVAL,
                "graph: null CSR / window arrays");
  GRAND_REQUIRE(g->n_edges == 0 || (g->col && g->erow && g->eid), GRAND_EINVAL, "graph: null edge arrays");
  GRAND_REQUIRE(g->n_split_tiles == 0 || g->split_tile0 != nullptr, GRAND_EINVAL, "graph: null split_tile0");
  return GRAND_OK;
}

int launch_rhs_mode(int mode, const RhsArgs& A, cudaStream_t stream);
int launch_stream_mode(int mode, const RhsArgs& A, cudaStream_t stream);
bool stream_kernel_ok(const RhsArgs& A);

// One non-blocking side stream + fork/join events per device, created on first use and kept for the life of the
// process (the only state this library owns; GRAND_FORK=1 enables the fork).
struct SideStream {
  cudaStream_t stream;
  cudaEvent_t fork, join;
};
static SideStream* side_stream() {
  static SideStream table[16];
  static bool ready[16] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  if (!ready[dev]) {
    // creation is not a stream operation; if a capture is in progress on the calling stream it is unaffected, but
    // the warm-up call every user makes before capturing creates these anyway
    if (cudaStreamCreateWithFlags(&table[dev].stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&table[dev].fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&table[dev].join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    ready[dev] = true;
  }
  return &table[dev];
}
// fork / join of the library's side stream around a pair of independent launches (captured into CUDA Graphs as two
// parallel branches)
int side_fork(cudaStream_t stream, cudaStream_t* side) {
  SideStream* ss = side_stream();
  GRAND_REQUIRE(ss != nullptr, GRAND_ECUDA, "could not create the side stream");
  GRAND_CUDA_OK(cudaEventRecord(ss->fork, stream));
  GRAND_CUDA_OK(cudaStreamWaitEvent(ss->stream, ss->fork, 0));
  *side = ss->stream;
  return GRAND_OK;
}
int side_join(cudaStream_t stream) {
  SideStream* ss = side_stream();
  GRAND_REQUIRE(ss != nullptr, GRAND_ECUDA, "could not create the side stream");
  GRAND_CUDA_OK(cudaEventRecord(ss->join, ss->stream));
  GRAND_CUDA_OK(cudaStreamWaitEvent(stream, ss->join, 0));
  return GRAND_OK;
}
static bool fork_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GRAND_FORK");
    v = (e != nullptr && atoi(e) != 0) ? 1 : 0;   // off by default: measured no gain on B200 (profiles/r01d), both
                                                   // kernels already fill every SM with resident warps
  }
  return v == 1;
}

// GRAND_L2_HINTS=0..3 (bit 0 evict_last gathers, bit 1 evict_first streams); default from round-1 measurements
static int l2_hints() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GRAND_L2_HINTS");
    v = e != nullptr ? atoi(e) & 3 : 3;
  }
  return v;
}
int launch_split_stats(const RhsArgs& A, cudaStream_t stream);

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
  A.l2_hints = l2_hints();
  A.split_stats = ws->split_stats; A.row_stats = ws->row_stats; A.split_partial = ws->split_partial;
  A.split_counter = ws->split_counter;
  A.lrow_stats = ws->lrow_stats; A.lpartial = ws->lpartial; A.lcounter = ws->lcounter;
  if ((rc = fill_push(ep, &A.push)) != GRAND_OK) return rc;
  A.tile_begin = 0;
  A.tile_end = g->n_tiles;
  const bool stream_ok = stream_kernel_ok(A);
  GRAND_REQUIRE(A.push.ptr == nullptr || stream_ok, GRAND_EUNSUPPORTED,
                "the peer push needs the persistent edge kernel (16-byte rows)");
  GRAND_REQUIRE(stream_ok || g->n_split_tiles == 0 || (ws->split_partial && ws->split_counter), GRAND_EINVAL,
                "graph has split rows: workspace split_partial / split_counter are null");
  if (project) {
    rc = launch_project_qk(x, g->n_rows, g->n_cols, a->d, a->wq, a->bq, a->wk, a->bk, A.a.ad_eff, ws->q, ws->k,
                           stream);
    if (rc != GRAND_OK) return rc;
  }
  if (stream_ok) {
    if (fused_ok(A.a)) return launch_stream_mode(MODE_FUSED, A, stream);
    if ((rc = run_general_prepass(A.g, A.a, ws, stream)) != GRAND_OK) return rc;
    return launch_stream_mode(MODE_PRECOMP, A, stream);
  }
  if (fused_ok(A.a)) {
    if (g->n_split_tiles == 0) return launch_rhs_mode(MODE_FUSED, A, stream);
    GRAND_REQUIRE(ws->split_stats != nullptr && ws->row_stats != nullptr, GRAND_EINVAL,
                  "graph has split rows: workspace split_stats / row_stats are null");
    // Fork: the packed windows (whole short rows) do not need the long-row statistics, so they run on a side stream
    // while this stream does  statistics pre-pass -> chunk tiles;  both kernels are latency-bound on their own and
    // overlap well.  The fork/join is plain event plumbing and is captured into CUDA Graphs like any other work.
    SideStream* ss = side_stream();
    const bool fork = ss != nullptr && fork_enabled() && g->n_tiles > g->n_split_tiles;
    if (fork) {
      GRAND_CUDA_OK(cudaEventRecord(ss->fork, stream));
      GRAND_CUDA_OK(cudaStreamWaitEvent(ss->stream, ss->fork, 0));
      RhsArgs P = A;
      P.tile_begin = g->n_split_tiles;
      if ((rc = launch_rhs_mode(MODE_FUSED, P, ss->stream)) != GRAND_OK) return rc;
      GRAND_CUDA_OK(cudaEventRecord(ss->join, ss->stream));
      A.tile_end = g->n_split_tiles;
    }
    if ((rc = launch_split_stats(A, stream)) != GRAND_OK) return rc;
    if ((rc = launch_rhs_mode(MODE_FUSED, A, stream)) != GRAND_OK) return rc;
    if (fork) GRAND_CUDA_OK(cudaStreamWaitEvent(stream, ss->join, 0));
    return GRAND_OK;
  }
  if ((rc = run_general_prepass(A.g, A.a, ws, stream)) != GRAND_OK) return rc;
  return launch_rhs_mode(MODE_PRECOMP, A, stream);
}

int integrate_fixed_coop(const RhsArgs& A, bool attention, const grand_attention* a, int method, const float* dt, int n_steps,
                         float* y, float* stage_buffers, int* sync, cudaStream_t stream);

// Small graphs: the whole fixed-grid integration as one cooperative launch (csrc/rhs_coop.cu).  GRAND_EUNSUPPORTED (no
// error message) = "not for this kernel", the caller falls back to one launch group per evaluation.
int integrate_fixed_coop_entry(const grand_graph* g, const grand_attention* a, const float* w_csr, int d,
                               const grand_workspace* ws, const grand_epilogue* residual, int method, const float* dt,
                               int n_steps, float* y, float* stage_buffers, cudaStream_t stream) {
  if (ws == nullptr || ws->coop_sync == nullptr || residual->peers != nullptr || residual->coef_dev != nullptr) return GRAND_EUNSUPPORTED;
  if (check_graph(g) != GRAND_OK) return GRAND_EUNSUPPORTED;
  RhsArgs A;
  memset(&A, 0, sizeof(A));
  A.g = view(g);
  A.d = d;
  if (a != nullptr) {
    if (make_attn_view(a, &A.a) != GRAND_OK || ws->q == nullptr || ws->k == nullptr) return GRAND_EUNSUPPORTED;
    if (a->edge_weight_csr != nullptr && false) return GRAND_EUNSUPPORTED;
    A.q = ws->q; A.k = ws->k;
  } else {
    if (w_csr == nullptr) return GRAND_EUNSUPPORTED;
    A.w_csr = w_csr;
  }
  grand_epilogue ep0;
  memset(&ep0, 0, sizeof(ep0));
  ep0.alpha_train = residual->alpha_train; ep0.beta_train = residual->beta_train; ep0.x0 = residual->x0;
  ep0.no_alpha_sigmoid = residual->no_alpha_sigmoid;
  if (ep0.alpha_train == nullptr || (ep0.x0 != nullptr && ep0.beta_train == nullptr)) return GRAND_EUNSUPPORTED;
  A.ep = view(&ep0);
  A.lrow_stats = ws->lrow_stats; A.lpartial = ws->lpartial; A.lcounter = ws->lcounter;
  A.l2_hints = 0;
  return integrate_fixed_coop(A, a != nullptr, a, method, dt, n_steps, y, stage_buffers, ws->coop_sync, stream);
}

int rhs_attention_backward_impl(const grand_graph* g, const grand_attention* a, const float* x, const float* grad_f,
                                const grand_workspace* ws, const float* alpha_train, int no_alpha_sigmoid,
                                float* row_dot, float* abar_out, float* dq, float* dk, GraphView gv,
                                cudaStream_t stream, const float* c_in);

int rhs_attention_backward_host(const grand_graph* g, const grand_attention* a, const float* x, const float* grad_f,
                                const grand_workspace* ws, const float* alpha_train, int no_alpha_sigmoid,
                                float* row_dot, float* abar_out, float* dq, float* dk, cudaStream_t stream) {
  int rc;
  if ((rc = check_graph(g)) != GRAND_OK) return rc;
  GRAND_REQUIRE(a && x && grad_f && ws && ws->q && ws->k && alpha_train && abar_out && dq && dk, GRAND_EINVAL,
                "rhs_attention_backward: null argument");
  GRAND_REQUIRE(g->n_cols == g->n_rows, GRAND_EUNSUPPORTED, "rhs_attention_backward: sharded graphs are not built");
  RhsArgs A;
  A.g = view(g);
  if ((rc = make_attn_view(a, &A.a)) != GRAND_OK) return rc;
  A.d = a->d;
  A.x = x; A.q = ws->q; A.k = ws->k;
  A.split_stats = ws->split_stats; A.row_stats = ws->row_stats; A.split_partial = ws->split_partial;
  A.split_counter = ws->split_counter;
  A.lrow_stats = nullptr; A.lpartial = nullptr; A.lcounter = nullptr;
  memset(&A.push, 0, sizeof(A.push));
  rc = launch_project_qk(x, g->n_rows, g->n_cols, a->d, a->wq, a->bq, a->wk, a->bk, A.a.ad_eff, ws->q, ws->k, stream);
  if (rc != GRAND_OK) return rc;
  if (g->n_split_tiles > 0) {
    GRAND_REQUIRE(ws->split_stats && ws->row_stats && ws->split_counter, GRAND_EINVAL,
                  "rhs_attention_backward: graph has long rows, workspace split arrays are null");
    if ((rc = launch_split_stats(A, stream)) != GRAND_OK) return rc;
  }
  return rhs_attention_backward_impl(g, a, x, grad_f, ws, alpha_train, no_alpha_sigmoid, row_dot, abar_out, dq, dk,
                                     A.g, stream, nullptr);
}

// ---- frozen attention (AttODEblock): dL/d(abar_e) accumulates over the whole integration, the softmax is differentiated
//      once at x(0) ----------------------------------------------------------------------------------------------------
// c[p] = alpha * grad_f[row_p] . x[col_p]  in CSR order: the gradient of one Laplacian RHS evaluation w.r.t. its per-edge
// weights (f_i = alpha (sum_e w_e x_j - x_i) + ...).  One warp per row, the lanes split the state width.
__global__ void __launch_bounds__(256)
edge_dot_kernel(GraphView g, int d, const float* __restrict__ x, const float* __restrict__ grad,
                const float* __restrict__ alpha_train, int no_alpha_sigmoid, float* __restrict__ c_out) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= g.n_rows) return;
  float alpha = __ldg(alpha_train);
  if (!no_alpha_sigmoid) alpha = 1.f / (1.f + expf(-alpha));
  const int beg = __ldg(g.rowptr + row), end = __ldg(g.rowptr + row + 1);
  const float* gi = grad + (size_t)row * d;
  for (int p = beg; p < end; ++p) {
    const float* xj = x + (size_t)__ldg(g.col + p) * d;
    float s = 0.f;
    if ((d & 3) == 0) {
      for (int c = lane * 4; c < d; c += 128) {
        const float4 a = ldg4(gi + c), b = ldg4(xj + c);
        s = fmaf(a.x, b.x, s); s = fmaf(a.y, b.y, s); s = fmaf(a.z, b.z, s); s = fmaf(a.w, b.w, s);
      }
    } else {
      for (int c = lane; c < d; c += 32) s = fmaf(__ldg(gi + c), __ldg(xj + c), s);
    }
    s = warp_sum(s);
    if (lane == 0) c_out[p] = alpha * s;
  }
}

int edge_dot_impl(const grand_graph* g, int d, const float* x, const float* grad, const float* alpha_train,
                  int no_alpha_sigmoid, float* c_out, cudaStream_t stream) {
  int rc;
  if ((rc = check_graph(g)) != GRAND_OK) return rc;
  GRAND_REQUIRE(d > 0 && x && grad && alpha_train && (c_out || g->n_edges == 0), GRAND_EINVAL, "edge_dot: null argument");
  if (g->n_rows > 0 && g->n_edges > 0) {
    edge_dot_kernel<<<(g->n_rows + 7) / 8, 256, 0, stream>>>(view(g), d, x, grad, alpha_train, no_alpha_sigmoid, c_out);
    GRAND_CUDA_OK(cudaGetLastError());
    count_launch();
  }
  return GRAND_OK;
}

// softmax + head-mean backward of SpGraphTransAttentionLayer.forward at x0, given c = dL/d(mean_h attention) in CSR order
int attention_weights_backward_host(const grand_graph* g, const grand_attention* a, const float* x0,
                                    const grand_workspace* ws, const float* c_csr, float* row_dot, float* dq, float* dk,
                                    cudaStream_t stream) {
  int rc;
  if ((rc = check_graph(g)) != GRAND_OK) return rc;
  GRAND_REQUIRE(a && x0 && ws && ws->q && ws->k && c_csr && dq && dk, GRAND_EINVAL, "attention_weights_backward: null argument");
  GRAND_REQUIRE(g->n_cols == g->n_rows, GRAND_EUNSUPPORTED, "attention_weights_backward: sharded graphs are not built");
  RhsArgs A;
  memset(&A, 0, sizeof(A));
  A.g = view(g);
  if ((rc = make_attn_view(a, &A.a)) != GRAND_OK) return rc;
  A.d = a->d;
  A.x = x0; A.q = ws->q; A.k = ws->k;
  A.split_stats = ws->split_stats; A.row_stats = ws->row_stats; A.split_partial = ws->split_partial;
  A.split_counter = ws->split_counter;
  rc = launch_project_qk(x0, g->n_rows, g->n_cols, a->d, a->wq, a->bq, a->wk, a->bk, A.a.ad_eff, ws->q, ws->k, stream);
  if (rc != GRAND_OK) return rc;
  if (g->n_split_tiles > 0) {
    GRAND_REQUIRE(ws->split_stats && ws->row_stats && row_dot, GRAND_EINVAL,
                  "attention_weights_backward: graph has long rows, workspace split arrays / row_dot are null");
    if ((rc = launch_split_stats(A, stream)) != GRAND_OK) return rc;
  }
  // alpha is already inside c; the kernel still loads *alpha_train, so hand it any valid device float
  return rhs_attention_backward_impl(g, a, x0, x0, ws, a->bq, 1, row_dot, nullptr, dq, dk, A.g, stream, c_csr);
}

int rhs_laplacian_impl(const grand_graph* g, int d, const float* x, const float* w_csr, const grand_workspace* ws,
                       const grand_epilogue* ep, cudaStream_t stream) {
  int rc;
  if ((rc = check_graph(g)) != GRAND_OK) return rc;
  if ((rc = check_epilogue(ep)) != GRAND_OK) return rc;
  GRAND_REQUIRE(x != nullptr && (w_csr != nullptr || g->n_edges == 0), GRAND_EINVAL, "rhs_laplacian: null argument");
  GRAND_REQUIRE(d > 0, GRAND_EINVAL, "rhs_laplacian: d = %d", d);
  RhsArgs A;
  A.g = view(g);
  A.a = AttnView{};
  A.ep = view(ep);
  A.d = d;
  A.x = x; A.q = nullptr; A.k = nullptr; A.w_csr = w_csr;
  A.scores = nullptr; A.seg_stats = nullptr; A.gmax = nullptr; A.split_stats = nullptr; A.row_stats = nullptr;
  A.l2_hints = l2_hints();
  A.tile_begin = 0;
  A.tile_end = g->n_tiles;
  A.split_partial = ws ? ws->split_partial : nullptr;
  A.split_counter = ws ? ws->split_counter : nullptr;
  A.lrow_stats = nullptr;
  if ((rc = fill_push(ep, &A.push)) != GRAND_OK) return rc;
  A.lpartial = ws ? ws->lpartial : nullptr;
  A.lcounter = ws ? ws->lcounter : nullptr;
  if (stream_kernel_ok(A)) return launch_stream_mode(MODE_WEIGHT, A, stream);
  GRAND_REQUIRE(g->n_split_tiles == 0 || (ws && ws->split_partial && ws->split_counter), GRAND_EINVAL,
                "graph has split rows: workspace split_partial / split_counter are null");
  return launch_rhs_mode(MODE_WEIGHT, A, stream);
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
