// extern "C" surface of libgrand_b200.so (include/grand_b200.h) + the fixed-grid stage driver and the small
// dense vector kernels the solvers need.
#include <stdarg.h>
#include <string.h>

#include <stdlib.h>

#include "common.cuh"

namespace grand {

static thread_local char g_error[512] = "";
bool pdl_enabled() {
  // Measured on B200 (profiles/r02p_pdl.txt): 7.17 vs 7.09 ms per ogbn-arxiv forward, Photo 121 vs 109 us per evaluation,
  // Cora 41.9 vs 42.2 -- the successor's early blocks take SM resources from the persistent predecessor's tail instead of
  // hiding a gap.  OFF by default; GRAND_PDL=1 turns it on (read once).
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GRAND_PDL");
    v = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

static int64_t g_launches = 0;   // diagnostic counter only (grand_launch_count); updated atomically

void count_launch(int n) { __atomic_fetch_add(&g_launches, (int64_t)n, __ATOMIC_RELAXED); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return GRAND_ECUDA;
}

int launch_linear(const float* x, int rows, int d_in, const float* w, const float* b, int d_out, float* y, int ldy,
                  cudaStream_t stream);
void set_projection_impl(int impl);
int launch_project_qk(const float* x, int n_rows, int n_cols, int d, const float* wq, const float* bq,
                      const float* wk, const float* bk, int ad, float* q, float* k, cudaStream_t stream);
int rhs_attention_impl(const grand_graph* g, const grand_attention* a, const float* x, const grand_workspace* ws,
                       const grand_epilogue* ep, bool project, cudaStream_t stream);
int rhs_laplacian_impl(const grand_graph* g, int d, const float* x, const float* w_csr, const grand_workspace* ws,
                       const grand_epilogue* ep, cudaStream_t stream);
int edge_dot_impl(const grand_graph* g, int d, const float* x, const float* grad, const float* alpha_train,
                  int no_alpha_sigmoid, float* c_out, cudaStream_t stream);
int attention_weights_backward_host(const grand_graph* g, const grand_attention* a, const float* x0,
                                    const grand_workspace* ws, const float* c_csr, float* row_dot, float* dq, float* dk,
                                    cudaStream_t stream);
int integrate_fixed_coop_entry(const grand_graph* g, const grand_attention* a, const float* w_csr, int d,
                               const grand_workspace* ws, const grand_epilogue* residual, int method, const float* dt,
                               int n_steps, float* y, float* stage_buffers, cudaStream_t stream);
int attention_weights_impl(const grand_graph* g, const grand_attention* a, const float* x, const grand_workspace* ws,
                           float* attention_out, float* prods_out, float* mean_csr_out, cudaStream_t stream);
int edge_values_to_csr_impl(const grand_graph* g, const float* w, int heads, float* w_csr, cudaStream_t stream);
int rhs_attention_backward_host(const grand_graph* g, const grand_attention* a, const float* x, const float* grad_f,
                                const grand_workspace* ws, const float* alpha_train, int no_alpha_sigmoid,
                                float* row_dot, float* abar_out, float* dq, float* dk, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------------------
// dense helpers
// ---------------------------------------------------------------------------------------------------------
struct LincombArgs {
  int n_terms;
  const float* term[GRAND_MAX_TERMS];
  float coef[GRAND_MAX_TERMS];
  const float* coef_dev;    // or null: device [n_terms] overriding coef[] (dopri5 controller)
};

__global__ void __launch_bounds__(256) lincomb_kernel(int64_t n, LincombArgs a, float* __restrict__ y) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  float c[GRAND_MAX_TERMS];
#pragma unroll
  for (int t = 0; t < GRAND_MAX_TERMS; ++t) c[t] = t < a.n_terms ? (a.coef_dev != nullptr ? __ldg(a.coef_dev + t) : a.coef[t]) : 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = 0.f;
#pragma unroll
    for (int t = 0; t < GRAND_MAX_TERMS; ++t)
      if (t < a.n_terms) v = fmaf(c[t], a.term[t][i], v);
    y[i] = v;
  }
}

struct ErrArgs {
  int n_terms;
  const float* k[GRAND_MAX_TERMS];
  float coef[GRAND_MAX_TERMS];
  const float* coef_dev;    // or null
  const float* tol_dev;     // or null: device (rtol, atol) pair... unused: tolerances are launch constants
};

__global__ void __launch_bounds__(256)
error_sumsq_kernel(int64_t n, ErrArgs a, const float* __restrict__ y0, const float* __restrict__ y1, float atol,
                   float rtol, double* __restrict__ sumsq) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double local = 0.0;
  float c[GRAND_MAX_TERMS];
#pragma unroll
  for (int t = 0; t < GRAND_MAX_TERMS; ++t) c[t] = t < a.n_terms ? (a.coef_dev != nullptr ? __ldg(a.coef_dev + t) : a.coef[t]) : 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float err = 0.f;
#pragma unroll
    for (int t = 0; t < GRAND_MAX_TERMS; ++t)
      if (t < a.n_terms) err = fmaf(c[t], a.k[t][i], err);
    const float tol = atol + rtol * fmaxf(fabsf(y0[i]), fabsf(y1[i]));
    const float r = err / tol;
    local += (double)r * (double)r;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(kFull, local, o);
  __shared__ double warp_part[8];
  if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += warp_part[w];
    atomicAdd(sumsq, s);
  }
}

__global__ void __launch_bounds__(256)
pack_rows_kernel(const float* __restrict__ x, const int* __restrict__ idx, int64_t n_idx, int d,
                 float* __restrict__ out) {
  // one warp per row, 128-bit lanes when the row width allows it
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n_idx) return;
  const float* src = x + (size_t)__ldg(idx + r) * d;
  float* dst = out + (size_t)r * d;
  if ((d & 3) == 0) {
    for (int c = lane * 4; c < d; c += 128) st_stream4(dst + c, ldg4(src + c));
  } else {
    for (int c = lane; c < d; c += 32) dst[c] = __ldg(src + c);
  }
}

// ---- node-partitioned multi-GPU: rows into the peers' halos, and the cross-GPU barrier -----------------------------------
struct PeerPtrs {
  float* out[GRAND_MAX_PEERS];
};

__global__ void __launch_bounds__(256)
peer_push_rows_kernel(const float* __restrict__ src, int n_push, int width, const int* __restrict__ push_src,
                      const int2* __restrict__ push_dst, PeerPtrs peers) {
  // one thread per (push entry, 16-byte piece): no dependent chain beyond entry -> source piece -> store
  const int pieces = width >> 2;
  const long long n_items = (long long)n_push * pieces;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < n_items; it += (long long)gridDim.x * blockDim.x) {
    const int entry = (int)(it / pieces), c = (int)(it - (long long)entry * pieces) * 4;
    const int row = __ldg(push_src + entry);
    const int2 dst = __ldg(push_dst + entry);
    const float4 v = ldg4(src + (size_t)row * width + c);
    *reinterpret_cast<float4*>(peers.out[dst.x] + (size_t)dst.y * width + c) = v;
  }
  __threadfence_system();
}

struct BarrierArgs {
  int world, rank;
  int* flags[GRAND_MAX_PEERS];     // flags[p]: rank p's arrival flags (int[GRAND_MAX_PEERS]); [rank] is the local one
  int* epoch;
  int* error;
};

__global__ void peer_barrier_kernel(BarrierArgs b) {
  __shared__ int s_epoch;
  if (threadIdx.x == 0) {
    s_epoch = *b.epoch + 1;
    *b.epoch = s_epoch;
  }
  __syncthreads();
  const int epoch = s_epoch, p = threadIdx.x;
  __threadfence_system();
  if (p < b.world && p != b.rank) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(b.flags[p] + b.rank), "r"(epoch) : "memory");
    const long long t0 = clock64();
    int seen;
    do {
      asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(seen) : "l"(b.flags[b.rank] + p) : "memory");
      if (seen - epoch >= 0) break;
      if (clock64() - t0 > 4000000000ll) {       // ~2 s at 2 GHz: a peer is gone; do not hang the GPU
        atomicExch(b.error, epoch);
        break;
      }
    } while (true);
  }
  __syncthreads();
}

static int elementwise_grid(int64_t n) {
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 8;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace grand

using namespace grand;

extern "C" {

int grand_abi_version(void) { return GRAND_ABI_VERSION; }

int64_t grand_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

const char* grand_last_error(void) { return g_error; }

int grand_workspace_sizes(const grand_graph* g, const grand_attention* a, int64_t* n_q, int64_t* n_k,
                          int64_t* n_scores, int64_t* n_seg_stats, int64_t* n_split_stats, int64_t* n_split_partial,
                          int64_t* n_split_counter, int64_t* n_lrow_stats, int64_t* n_lpartial, int64_t* n_lcounter) {
  GRAND_REQUIRE(g != nullptr, GRAND_EINVAL, "workspace_sizes: graph is null");
  const int64_t ad_eff = a ? (a->score_type == GRAND_SCORE_EXP_KERNEL2 ? 2 * a->ad : a->ad) : 0;
  const int64_t heads = a ? a->heads : 0;
  const int64_t d = a ? a->d : 0;
  if (n_q) *n_q = (int64_t)g->n_rows * ad_eff;
  if (n_k) *n_k = (int64_t)g->n_cols * ad_eff;
  if (n_scores) *n_scores = (int64_t)g->n_edges * heads;
  if (n_seg_stats) *n_seg_stats = (int64_t)g->n_cols * heads * 2;
  if (n_split_stats) *n_split_stats = (int64_t)g->n_split_tiles * heads * 2;
  if (n_split_partial) *n_split_partial = (int64_t)g->n_split_tiles * d;
  if (n_split_counter) *n_split_counter = g->n_split_rows;
  if (n_lrow_stats) *n_lrow_stats = (int64_t)g->n_lunits * heads * 2;
  if (n_lpartial) *n_lpartial = (int64_t)g->n_lunits * (heads > 0 ? heads : 1) * d;
  if (n_lcounter) *n_lcounter = g->n_lrows;
  return GRAND_OK;
}

int grand_project_qk(const float* x, int32_t n_rows, int32_t n_cols, const grand_attention* a, float* q, float* k,
                     void* stream) {
  GRAND_REQUIRE(x && a && q && k && a->wq && a->bq && a->wk && a->bk, GRAND_EINVAL, "project_qk: null argument");
  GRAND_REQUIRE(n_rows >= 0 && n_cols >= n_rows && a->d > 0 && a->ad > 0, GRAND_EINVAL, "project_qk: bad sizes");
  const int ad_eff = a->score_type == GRAND_SCORE_EXP_KERNEL2 ? 2 * a->ad : a->ad;
  return launch_project_qk(x, n_rows, n_cols, a->d, a->wq, a->bq, a->wk, a->bk, ad_eff, q, k, (cudaStream_t)stream);
}

int grand_linear(const float* x, int32_t n_rows, int32_t d_in, const float* w, const float* b, int32_t d_out, float* y,
                 int32_t ldy, void* stream) {
  GRAND_REQUIRE(x && w && y, GRAND_EINVAL, "linear: null argument");
  GRAND_REQUIRE(n_rows >= 0 && d_in > 0 && d_out > 0 && ldy >= d_out, GRAND_EINVAL, "linear: bad sizes");
  return launch_linear(x, n_rows, d_in, w, b, d_out, y, ldy, (cudaStream_t)stream);
}

int grand_set_projection_impl(int32_t impl) {
  GRAND_REQUIRE(impl >= -1 && impl <= 3, GRAND_EINVAL,
                "set_projection_impl: 0 (ffma), 1 (mma.sync), 2 (tcgen05, operands in shared memory), 3 (tcgen05, X in tensor "
                "memory) or -1 (back to the default / GRAND_PROJ)");
  set_projection_impl(impl);
  return GRAND_OK;
}

int grand_rhs_attention(const grand_graph* g, const grand_attention* a, const float* x, const grand_workspace* ws,
                        const grand_epilogue* ep, void* stream) {
  GRAND_REQUIRE(a && a->wq && a->bq && a->wk && a->bk, GRAND_EINVAL, "rhs_attention: null attention weights");
  return rhs_attention_impl(g, a, x, ws, ep, true, (cudaStream_t)stream);
}

int grand_rhs_attention_edges(const grand_graph* g, const grand_attention* a, const float* x, const grand_workspace* ws,
                              const grand_epilogue* ep, void* stream) {
  return rhs_attention_impl(g, a, x, ws, ep, false, (cudaStream_t)stream);
}

int grand_rhs_attention_backward(const grand_graph* g, const grand_attention* a, const float* x, const float* grad_f,
                                 const grand_workspace* ws, const float* alpha_train, int32_t no_alpha_sigmoid,
                                 float* row_dot, float* abar_out, float* dq, float* dk, void* stream) {
  GRAND_REQUIRE(a && a->wq && a->bq && a->wk && a->bk, GRAND_EINVAL, "rhs_attention_backward: null attention weights");
  return rhs_attention_backward_host(g, a, x, grad_f, ws, alpha_train, no_alpha_sigmoid, row_dot, abar_out, dq, dk,
                                     (cudaStream_t)stream);
}

int grand_edge_dot(const grand_graph* g, int32_t d, const float* x, const float* grad_f, const float* alpha_train,
                   int32_t no_alpha_sigmoid, float* c_csr_out, void* stream) {
  return edge_dot_impl(g, d, x, grad_f, alpha_train, no_alpha_sigmoid, c_csr_out, (cudaStream_t)stream);
}

int grand_attention_weights_backward(const grand_graph* g, const grand_attention* a, const float* x0,
                                     const grand_workspace* ws, const float* c_csr, float* row_dot, float* dq, float* dk,
                                     void* stream) {
  GRAND_REQUIRE(a && a->wq && a->bq && a->wk && a->bk, GRAND_EINVAL, "attention_weights_backward: null attention weights");
  return attention_weights_backward_host(g, a, x0, ws, c_csr, row_dot, dq, dk, (cudaStream_t)stream);
}

int grand_rhs_laplacian(const grand_graph* g, int32_t d, const float* x, const float* w_csr, const grand_workspace* ws,
                        const grand_epilogue* ep, void* stream) {
  return rhs_laplacian_impl(g, d, x, w_csr, ws, ep, (cudaStream_t)stream);
}

int grand_attention_weights(const grand_graph* g, const grand_attention* a, const float* x, const grand_workspace* ws,
                            float* attention_out, float* prods_out, float* mean_csr_out, void* stream) {
  GRAND_REQUIRE(a && a->wq && a->bq && a->wk && a->bk, GRAND_EINVAL, "attention_weights: null attention weights");
  return attention_weights_impl(g, a, x, ws, attention_out, prods_out, mean_csr_out, (cudaStream_t)stream);
}

int grand_edge_values_to_csr(const grand_graph* g, const float* w, int32_t heads, float* w_csr, void* stream) {
  return edge_values_to_csr_impl(g, w, heads, w_csr, (cudaStream_t)stream);
}

int grand_pack_rows(const float* x, const int32_t* idx, int64_t n_idx, int32_t d, float* out, void* stream) {
  GRAND_REQUIRE(n_idx >= 0 && d > 0 && (n_idx == 0 || (x && idx && out)), GRAND_EINVAL, "pack_rows: bad argument");
  if (n_idx == 0) return GRAND_OK;
  pack_rows_kernel<<<(unsigned)((n_idx + 7) / 8), 256, 0, (cudaStream_t)stream>>>(x, idx, n_idx, d, out);
  GRAND_CUDA_OK(cudaGetLastError());
  count_launch();
  return GRAND_OK;
}

int grand_lincomb(int64_t n, int32_t n_terms, const float* const* term, const float* coef, float* y_out,
                  void* stream) {
  GRAND_REQUIRE(n >= 0 && n_terms >= 1 && n_terms <= GRAND_MAX_TERMS && term && coef && y_out, GRAND_EINVAL,
                "lincomb: bad argument");
  if (n == 0) return GRAND_OK;
  LincombArgs a;
  a.n_terms = n_terms;
  a.coef_dev = nullptr;
  for (int t = 0; t < GRAND_MAX_TERMS; ++t) {
    a.term[t] = t < n_terms ? term[t] : nullptr;
    a.coef[t] = t < n_terms ? coef[t] : 0.f;
  }
  lincomb_kernel<<<elementwise_grid(n), 256, 0, (cudaStream_t)stream>>>(n, a, y_out);
  count_launch();
  GRAND_CUDA_OK(cudaGetLastError());
  return GRAND_OK;
}

int grand_error_sumsq(int64_t n, int32_t n_terms, const float* const* k, const float* coef, const float* y0,
                      const float* y1, float atol, float rtol, double* sumsq, void* stream) {
  GRAND_REQUIRE(n >= 0 && n_terms >= 1 && n_terms <= GRAND_MAX_TERMS && k && coef && y0 && y1 && sumsq, GRAND_EINVAL,
                "error_sumsq: bad argument");
  if (n == 0) return GRAND_OK;
  ErrArgs a;
  a.n_terms = n_terms;
  a.coef_dev = nullptr;
  a.tol_dev = nullptr;
  for (int t = 0; t < GRAND_MAX_TERMS; ++t) {
    a.k[t] = t < n_terms ? k[t] : nullptr;
    a.coef[t] = t < n_terms ? coef[t] : 0.f;
  }
  error_sumsq_kernel<<<elementwise_grid(n), 256, 0, (cudaStream_t)stream>>>(n, a, y0, y1, atol, rtol, sumsq);
  count_launch();
  GRAND_CUDA_OK(cudaGetLastError());
  return GRAND_OK;
}

// ---- peer arenas ------------------------------------------------------------------------------------------------------
int grand_peer_alloc(int64_t bytes, void** ptr) {
  GRAND_REQUIRE(bytes > 0 && ptr != nullptr, GRAND_EINVAL, "peer_alloc: bad argument");
  GRAND_CUDA_OK(cudaMalloc(ptr, (size_t)bytes));
  GRAND_CUDA_OK(cudaMemset(*ptr, 0, (size_t)bytes));
  GRAND_CUDA_OK(cudaDeviceSynchronize());
  return GRAND_OK;
}
int grand_peer_free(void* ptr) {
  if (ptr != nullptr) GRAND_CUDA_OK(cudaFree(ptr));
  return GRAND_OK;
}
int grand_peer_export(void* ptr, void* handle64) {
  GRAND_REQUIRE(ptr != nullptr && handle64 != nullptr, GRAND_EINVAL, "peer_export: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  GRAND_CUDA_OK(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), ptr));
  return GRAND_OK;
}
int grand_peer_open(const void* handle64, void** ptr) {
  GRAND_REQUIRE(ptr != nullptr && handle64 != nullptr, GRAND_EINVAL, "peer_open: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  GRAND_CUDA_OK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return GRAND_OK;
}
int grand_peer_close(void* ptr) {
  if (ptr != nullptr) GRAND_CUDA_OK(cudaIpcCloseMemHandle(ptr));
  return GRAND_OK;
}

static int check_peers(const grand_peers* pr) {
  GRAND_REQUIRE(pr != nullptr && pr->world >= 1 && pr->world <= GRAND_MAX_PEERS && pr->rank >= 0 && pr->rank < pr->world,
                GRAND_EINVAL, "peers: null or bad world / rank");
  for (int p = 0; p < pr->world; ++p) GRAND_REQUIRE(pr->arena[p] != nullptr, GRAND_EINVAL, "peers: arena %d is null", p);
  return GRAND_OK;
}

int grand_peer_push_rows(const grand_peers* pr, const float* src, int32_t n_rows, int32_t width, void* stream) {
  int rc;
  if ((rc = check_peers(pr)) != GRAND_OK) return rc;
  GRAND_REQUIRE(src != nullptr && n_rows >= 0 && width > 0 && (width & 3) == 0 && pr->push_ptr && pr->push_dst &&
                    pr->n_push >= 0 && (pr->n_push == 0 || pr->push_src), GRAND_EINVAL, "peer_push_rows: bad argument");
  if (n_rows == 0 || pr->world == 1 || pr->n_push == 0) return GRAND_OK;
  const ptrdiff_t off = reinterpret_cast<const char*>(src) - reinterpret_cast<const char*>(pr->arena[pr->rank]);
  GRAND_REQUIRE(off >= 0, GRAND_EINVAL, "peer_push_rows: src does not lie in the arena");
  PeerPtrs pp;
  for (int p = 0; p < GRAND_MAX_PEERS; ++p)
    pp.out[p] = p < pr->world ? reinterpret_cast<float*>(reinterpret_cast<char*>(pr->arena[p]) + off) : nullptr;
  const long long n_items = (long long)pr->n_push * (width >> 2);
  long long blocks = (n_items + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  peer_push_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      src, pr->n_push, width, pr->push_src, reinterpret_cast<const int2*>(pr->push_dst), pp);
  GRAND_CUDA_OK(cudaGetLastError());
  count_launch();
  return GRAND_OK;
}

int grand_peer_barrier(const grand_peers* pr, void* stream) {
  int rc;
  if ((rc = check_peers(pr)) != GRAND_OK) return rc;
  if (pr->world == 1) return GRAND_OK;
  BarrierArgs b;
  b.world = pr->world; b.rank = pr->rank;
  for (int p = 0; p < GRAND_MAX_PEERS; ++p)
    b.flags[p] = p < pr->world ? reinterpret_cast<int*>(reinterpret_cast<char*>(pr->arena[p]) + pr->flags_offset) : nullptr;
  b.epoch = b.flags[pr->rank] + GRAND_MAX_PEERS;
  b.error = b.epoch + 1;
  peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(b);
  GRAND_CUDA_OK(cudaGetLastError());
  count_launch();
  return GRAND_OK;
}

int grand_peer_error(const grand_peers* pr, int32_t* error_out) {
  int rc;
  if ((rc = check_peers(pr)) != GRAND_OK) return rc;
  GRAND_REQUIRE(error_out != nullptr, GRAND_EINVAL, "peer_error: null argument");
  const char* flags = reinterpret_cast<const char*>(pr->arena[pr->rank]) + pr->flags_offset;
  GRAND_CUDA_OK(cudaMemcpy(error_out, flags + (GRAND_MAX_PEERS + 1) * sizeof(int32_t), sizeof(int32_t), cudaMemcpyDeviceToHost));
  return GRAND_OK;
}

// Fixed-grid integration on a node partition: same stage algebra as grand_integrate_fixed below.
int grand_integrate_fixed_sharded(const grand_graph* g, const grand_attention* a, const grand_workspace* ws,
                                  const grand_epilogue* residual, int32_t method, const float* dt, int32_t n_steps,
                                  float* y, float* yb, float* yc, float* k1, float* k2, float* k3, float* const* k_bufs,
                                  int32_t* k_toggle, const grand_peers* pr, void* stream_) {
  int rc;
  if ((rc = check_peers(pr)) != GRAND_OK) return rc;
  GRAND_REQUIRE(g && a && ws && residual && dt && y && yb && yc && k_bufs && k_bufs[0] && k_bufs[1] && k_toggle && n_steps >= 0,
                GRAND_EINVAL, "integrate_fixed_sharded: null argument");
  GRAND_REQUIRE(method == GRAND_EULER || method == GRAND_MIDPOINT || method == GRAND_RK4, GRAND_EINVAL,
                "integrate_fixed_sharded: unknown method %d", method);
  GRAND_REQUIRE(method != GRAND_RK4 || (k1 && k2 && k3), GRAND_EINVAL, "integrate_fixed_sharded: rk4 needs k1..k3");
  cudaStream_t stream = (cudaStream_t)stream_;
  const int d = a->d;
  const int ad_eff = a->score_type == GRAND_SCORE_EXP_KERNEL2 ? 2 * a->ad : a->ad;
  grand_workspace wsx = *ws;

  // first stage input: own rows of y -> the peers' halos (every later stage input is delivered by the edge kernel's epilogue)
  if ((rc = grand_peer_push_rows(pr, y, g->n_rows, d, stream_)) != GRAND_OK) return rc;

  auto eval = [&](const float* x, grand_epilogue& ep) -> int {
    ep.alpha_train = residual->alpha_train;
    ep.beta_train = residual->beta_train;
    ep.x0 = residual->x0;
    ep.no_alpha_sigmoid = residual->no_alpha_sigmoid;
    ep.peers = pr;
    float* kb = k_bufs[*k_toggle & 1];
    *k_toggle ^= 1;
    wsx.k = kb;
    int r;
    if ((r = launch_project_qk(x, g->n_rows, g->n_rows, d, a->wq, a->bq, a->wk, a->bk, ad_eff, wsx.q, kb, stream)) != GRAND_OK)
      return r;
    if ((r = grand_peer_push_rows(pr, kb, g->n_rows, ad_eff, stream_)) != GRAND_OK) return r;
    if ((r = grand_peer_barrier(pr, stream_)) != GRAND_OK) return r;
    return rhs_attention_impl(g, a, x, &wsx, &ep, false, stream);
  };

  float* cur = y;
  float* nxt = yc;
  for (int s = 0; s < n_steps; ++s) {
    const float h = dt[s];
    grand_epilogue ep;
    memset(&ep, 0, sizeof(ep));
    if (method == GRAND_EULER) {
      ep.y_out = nxt; ep.coef_f = h; ep.n_terms = 1; ep.term[0] = cur; ep.coef[0] = 1.f;
      if ((rc = eval(cur, ep)) != GRAND_OK) return rc;
    } else if (method == GRAND_MIDPOINT) {
      ep.y_out = yb; ep.coef_f = 0.5f * h; ep.n_terms = 1; ep.term[0] = cur; ep.coef[0] = 1.f;
      if ((rc = eval(cur, ep)) != GRAND_OK) return rc;
      memset(&ep, 0, sizeof(ep));
      ep.y_out = nxt; ep.coef_f = h; ep.n_terms = 1; ep.term[0] = cur; ep.coef[0] = 1.f;
      if ((rc = eval(yb, ep)) != GRAND_OK) return rc;
    } else {
      const float third = 1.f / 3.f;
      ep.f_out = k1; ep.y_out = yb; ep.coef_f = h * third; ep.n_terms = 1; ep.term[0] = cur; ep.coef[0] = 1.f;
      if ((rc = eval(cur, ep)) != GRAND_OK) return rc;
      memset(&ep, 0, sizeof(ep));
      ep.f_out = k2; ep.y_out = nxt; ep.coef_f = h; ep.n_terms = 2;
      ep.term[0] = cur; ep.coef[0] = 1.f; ep.term[1] = k1; ep.coef[1] = -h * third;
      if ((rc = eval(yb, ep)) != GRAND_OK) return rc;
      memset(&ep, 0, sizeof(ep));
      ep.f_out = k3; ep.y_out = yb; ep.coef_f = h; ep.n_terms = 3;
      ep.term[0] = cur; ep.coef[0] = 1.f; ep.term[1] = k1; ep.coef[1] = h; ep.term[2] = k2; ep.coef[2] = -h;
      if ((rc = eval(nxt, ep)) != GRAND_OK) return rc;
      memset(&ep, 0, sizeof(ep));
      ep.y_out = nxt; ep.coef_f = h * 0.125f; ep.n_terms = 4;
      ep.term[0] = cur; ep.coef[0] = 1.f; ep.term[1] = k1; ep.coef[1] = h * 0.125f;
      ep.term[2] = k2; ep.coef[2] = 3.f * h * 0.125f; ep.term[3] = k3; ep.coef[3] = 3.f * h * 0.125f;
      if ((rc = eval(yb, ep)) != GRAND_OK) return rc;
    }
    float* tmp = cur; cur = nxt; nxt = tmp;
  }
  // every rank must be done READING the halos of this integration before anybody starts delivering the next one's
  if ((rc = grand_peer_barrier(pr, stream_)) != GRAND_OK) return rc;
  if (cur != y && n_steps > 0) {
    GRAND_CUDA_OK(cudaMemcpyAsync(y, cur, (size_t)g->n_rows * d * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    count_launch();
  }
  return GRAND_OK;
}

// Fixed-grid integration: one RHS launch per stage; the stage combine is the RHS kernel's epilogue.
//   euler     y += dt f(y)
//   midpoint  ym = y + dt/2 f(y);  y += dt f(ym)
//   rk4       3/8 rule (torchdiffeq rk4_alt_step_func):
//               k1 = f(y); k2 = f(y + dt k1/3); k3 = f(y + dt (k2 - k1/3)); k4 = f(y + dt (k1 - k2 + k3));
//               y += dt (k1 + 3 k2 + 3 k3 + k4) / 8
// stage_buffers: 5 x [n_rows_total, d] floats (ya, yb, k1, k2, k3); y must have n_cols rows when n_cols > n_rows.
int grand_integrate_fixed(const grand_graph* g, const grand_attention* a, const float* w_csr, int32_t d,
                          const grand_workspace* ws, const grand_epilogue* residual, int32_t method, const float* dt,
                          int32_t n_steps, float* y, float* stage_buffers, void* stream_) {
  GRAND_REQUIRE(g && residual && dt && y && stage_buffers && n_steps >= 0, GRAND_EINVAL,
                "integrate_fixed: null argument");
  GRAND_REQUIRE(method == GRAND_EULER || method == GRAND_MIDPOINT || method == GRAND_RK4, GRAND_EINVAL,
                "integrate_fixed: unknown method %d", method);
  GRAND_REQUIRE(a != nullptr || w_csr != nullptr || g->n_edges == 0, GRAND_EINVAL,
                "integrate_fixed: neither attention nor Laplacian weights given");
  GRAND_REQUIRE(g->n_cols == g->n_rows, GRAND_EUNSUPPORTED,
                "integrate_fixed: sharded graphs need a halo exchange between stages; drive the stages from the host");
  cudaStream_t stream = (cudaStream_t)stream_;
  {
    // small graphs: the whole integration in ONE cooperative launch (csrc/rhs_coop.cu); anything it is not built for
    // answers "unsupported" without an error and takes the per-evaluation path below
    const int rc_coop = integrate_fixed_coop_entry(g, a, w_csr, d, ws, residual, method, dt, n_steps, y, stage_buffers, stream);
    if (rc_coop != GRAND_EUNSUPPORTED) return rc_coop;
  }
  const size_t plane = (size_t)g->n_cols * d;
  float* yb = stage_buffers;            // next-stage input
  float* yc = stage_buffers + plane;    // second scratch state
  float* k1 = stage_buffers + 2 * plane;
  float* k2 = stage_buffers + 3 * plane;
  float* k3 = stage_buffers + 4 * plane;

  auto eval = [&](const float* x, grand_epilogue& ep) -> int {
    ep.alpha_train = residual->alpha_train;
    ep.beta_train = residual->beta_train;
    ep.x0 = residual->x0;
    ep.no_alpha_sigmoid = residual->no_alpha_sigmoid;
    if (a != nullptr) return rhs_attention_impl(g, a, x, ws, &ep, true, stream);
    return rhs_laplacian_impl(g, d, x, w_csr, ws, &ep, stream);
  };

  // y ping-pongs between `y` and `yc` so that the gather source of a stage is never its own output
  float* cur = y;
  float* nxt = yc;
  for (int s = 0; s < n_steps; ++s) {
    const float h = dt[s];
    int rc;
    grand_epilogue ep;
    memset(&ep, 0, sizeof(ep));
    if (method == GRAND_EULER) {
      ep.y_out = nxt; ep.coef_f = h; ep.n_terms = 1; ep.term[0] = cur; ep.coef[0] = 1.f;
      if ((rc = eval(cur, ep)) != GRAND_OK) return rc;
    } else if (method == GRAND_MIDPOINT) {
      ep.y_out = yb; ep.coef_f = 0.5f * h; ep.n_terms = 1; ep.term[0] = cur; ep.coef[0] = 1.f;
      if ((rc = eval(cur, ep)) != GRAND_OK) return rc;
      memset(&ep, 0, sizeof(ep));
      ep.y_out = nxt; ep.coef_f = h; ep.n_terms = 1; ep.term[0] = cur; ep.coef[0] = 1.f;
      if ((rc = eval(yb, ep)) != GRAND_OK) return rc;
    } else {
      const float third = 1.f / 3.f;
      // k1; y2 = y + h/3 k1
      ep.f_out = k1; ep.y_out = yb; ep.coef_f = h * third; ep.n_terms = 1; ep.term[0] = cur; ep.coef[0] = 1.f;
      if ((rc = eval(cur, ep)) != GRAND_OK) return rc;
      // k2; y3 = y + h (k2 - k1/3)      (input yb, output nxt used as scratch)
      memset(&ep, 0, sizeof(ep));
      ep.f_out = k2; ep.y_out = nxt; ep.coef_f = h; ep.n_terms = 2;
      ep.term[0] = cur; ep.coef[0] = 1.f; ep.term[1] = k1; ep.coef[1] = -h * third;
      if ((rc = eval(yb, ep)) != GRAND_OK) return rc;
      // k3; y4 = y + h (k1 - k2 + k3)   (input nxt, output yb)
      memset(&ep, 0, sizeof(ep));
      ep.f_out = k3; ep.y_out = yb; ep.coef_f = h; ep.n_terms = 3;
      ep.term[0] = cur; ep.coef[0] = 1.f; ep.term[1] = k1; ep.coef[1] = h; ep.term[2] = k2; ep.coef[2] = -h;
      if ((rc = eval(nxt, ep)) != GRAND_OK) return rc;
      // k4; y_next = y + h/8 (k1 + 3 k2 + 3 k3 + k4)   (input yb, output nxt)
      memset(&ep, 0, sizeof(ep));
      ep.y_out = nxt; ep.coef_f = h * 0.125f; ep.n_terms = 4;
      ep.term[0] = cur; ep.coef[0] = 1.f; ep.term[1] = k1; ep.coef[1] = h * 0.125f;
      ep.term[2] = k2; ep.coef[2] = 3.f * h * 0.125f; ep.term[3] = k3; ep.coef[3] = 3.f * h * 0.125f;
      if ((rc = eval(yb, ep)) != GRAND_OK) return rc;
    }
    float* tmp = cur; cur = nxt; nxt = tmp;
  }
  if (cur != y && n_steps > 0)
  {
    GRAND_CUDA_OK(cudaMemcpyAsync(y, cur, plane * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    count_launch();
  }
  return GRAND_OK;
}

}  // extern "C"
