// extern "C" surface of libgrand_b200.so (include/grand_b200.h) + the fixed-grid stage driver and the small
// dense vector kernels the solvers need.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace grand {

static thread_local char g_error[512] = "";
static int64_t g_launches = 0;

void count_launch(int n) { g_launches += n; }

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

int launch_project_qk(const float* x, int n_rows, int n_cols, int d, const float* wq, const float* bq,
                      const float* wk, const float* bk, int ad, float* q, float* k, cudaStream_t stream);
int rhs_attention_impl(const grand_graph* g, const grand_attention* a, const float* x, const grand_workspace* ws,
                       const grand_epilogue* ep, bool project, cudaStream_t stream);
int rhs_laplacian_impl(const grand_graph* g, int d, const float* x, const float* w_csr, const grand_workspace* ws,
                       const grand_epilogue* ep, cudaStream_t stream);
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
};

__global__ void __launch_bounds__(256) lincomb_kernel(int64_t n, LincombArgs a, float* __restrict__ y) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = 0.f;
#pragma unroll
    for (int t = 0; t < GRAND_MAX_TERMS; ++t)
      if (t < a.n_terms) v = fmaf(a.coef[t], a.term[t][i], v);
    y[i] = v;
  }
}

struct ErrArgs {
  int n_terms;
  const float* k[GRAND_MAX_TERMS];
  float coef[GRAND_MAX_TERMS];
};

__global__ void __launch_bounds__(256)
error_sumsq_kernel(int64_t n, ErrArgs a, const float* __restrict__ y0, const float* __restrict__ y1, float atol,
                   float rtol, double* __restrict__ sumsq) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double local = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float err = 0.f;
#pragma unroll
    for (int t = 0; t < GRAND_MAX_TERMS; ++t)
      if (t < a.n_terms) err = fmaf(a.coef[t], a.k[t][i], err);
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

static int elementwise_grid(int64_t n) {
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 8;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace grand

using namespace grand;

extern "C" {

int grand_abi_version(void) { return GRAND_ABI_VERSION; }

int64_t grand_launch_count(void) { return g_launches; }

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
  for (int t = 0; t < GRAND_MAX_TERMS; ++t) {
    a.k[t] = t < n_terms ? k[t] : nullptr;
    a.coef[t] = t < n_terms ? coef[t] : 0.f;
  }
  error_sumsq_kernel<<<elementwise_grid(n), 256, 0, (cudaStream_t)stream>>>(n, a, y0, y1, atol, rtol, sumsq);
  count_launch();
  GRAND_CUDA_OK(cudaGetLastError());
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
