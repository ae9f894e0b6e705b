// dopri5 with the controller on the device: grand_dopri5_create / launch / destroy (include/grand_b200.h).
//
// torchdiffeq's RKAdaptiveStepsizeODESolver (witness: reference src/early_stop_solver.py:30-86) as ONE CUDA Graph:
//
//   f0 = f(y0)                                   -+
//   d0, d1 norms -> h0 -> probe f(y0 + h0 f0)     |  _select_initial_step (Hairer), controller kernels `init_a`, `init_b`
//   d2 norm -> dt                                -+
//   WHILE (t < t1 and attempts < max)             <- conditional graph node; its body is one attempted step:
//       pre:     every dt-dependent coefficient of the attempt -> state->coef (fp32 products rounded as torchdiffeq's)
//       6 x RHS  stage combines fused into the RHS epilogues, coefficients read from state->coef by the kernels
//       error    sum of squares of err / (atol + rtol max(|y0|, |y1|))
//       post:    ratio, accept / reject, t += dt, new dt (safety 0.9, factors [0.2, 10]), loop condition
//       advance: on accept, ONE elementwise pass: yprev <- ycur, fprev <- k0, ymid, ycur <- y1, k0 <- k6 (FSAL)
//   final: dense-output weights at t1 -> out
//
// The host-driven twin (graph-neural-pde_b200/odeint.py:Dopri5) does the same arithmetic with one device->host read per
// attempt; it stays for the early-stop integrator, whose per-step read-out needs the host anyway.
#include <math.h>
#include <string.h>

#include "common.cuh"

namespace grand {

int rhs_attention_impl(const grand_graph* g, const grand_attention* a, const float* x, const grand_workspace* ws,
                       const grand_epilogue* ep, bool project, cudaStream_t stream);
int rhs_laplacian_impl(const grand_graph* g, int d, const float* x, const float* w_csr, const grand_workspace* ws,
                       const grand_epilogue* ep, cudaStream_t stream);

// Dormand-Prince 5(4): torchdiffeq/_impl/dopri5.py
__constant__ double kBeta[6][6] = {
    {1. / 5, 0, 0, 0, 0, 0},
    {3. / 40, 9. / 40, 0, 0, 0, 0},
    {44. / 45, -56. / 15, 32. / 9, 0, 0, 0},
    {19372. / 6561, -25360. / 2187, 64448. / 6561, -212. / 729, 0, 0},
    {9017. / 3168, -355. / 33, 46732. / 5247, 49. / 176, -5103. / 18656, 0},
    {35. / 384, 0, 500. / 1113, 125. / 192, -2187. / 6784, 11. / 84}};
__constant__ double kCErr[7] = {35. / 384 - 1951. / 21600, 0, 500. / 1113 - 22642. / 50085, 125. / 192 - 451. / 720,
                                -2187. / 6784 - -12231. / 42400, 11. / 84 - 649. / 6300, -1. / 60.};
__constant__ double kCMid[7] = {6025192743. / 30085553152. / 2, 0, 51252292925. / 65400821598. / 2,
                                -2691868925. / 45128329728. / 2, 187940372067. / 1594534317056. / 2,
                                -1776094331. / 19743644256. / 2, 11237099. / 235043384. / 2};

// layout of state->coef
constexpr int OFF_INIT = 0;      // [2]  ya = 1 * ycur + h0 * k0                      (initial-step probe)
constexpr int OFF_LIN0 = 2;      // [2]  ya = 1 * ycur + beta[0][0] h * k0            (input of stage 1)
constexpr int OFF_STAGE = 4;     // [5][9] per stage: coef_f, coef[0..8)              (epilogue of RHS launch s)
constexpr int OFF_ERR = 49;      // [7]  nonzero c_error terms, in k order
constexpr int OFF_MID = 56;      // [8]  1 * ycur, then the nonzero c_mid terms
constexpr int OFF_INTERP = 64;   // [5]  weights of (yprev, ycur, ymid, fprev, k0) at t_end
static_assert(OFF_INTERP + 5 <= GRAND_DOPRI5_COEFS, "coefficient table too small");

__device__ __forceinline__ double rms(double sumsq, int numel) { return sqrt(sumsq / (double)numel); }

__global__ void dopri5_reset_kernel(grand_dopri5_state* s, double t0, double t1, float rtol, float atol, int numel,
                                    int max_attempts) {
  s->t = t0; s->t_prev = t0; s->dt = 0.0; s->t_end = t1;
  s->acc[0] = s->acc[1] = s->acc[2] = 0.0;
  s->h0 = 0.0;
  s->rtol = rtol; s->atol = atol;
  s->numel = numel; s->attempts = 0; s->accepted = 0; s->nfe = 1; s->max_attempts = max_attempts;
  s->accept = 0; s->have_interp = 0; s->status = 0;
  for (int i = 0; i < GRAND_DOPRI5_COEFS; ++i) s->coef[i] = 0.f;
}

// after d0 = |y0|, d1 = |f0| (rms of the tolerance-scaled vectors): h0 and the probe combination
__global__ void dopri5_init_a_kernel(grand_dopri5_state* s) {
  const double d0 = rms(s->acc[0], s->numel), d1 = rms(s->acc[1], s->numel);
  double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
  h0 = (double)(float)fabs(h0);
  s->h0 = h0;
  s->coef[OFF_INIT + 0] = 1.f;
  s->coef[OFF_INIT + 1] = (float)h0;
  s->acc[2] = 0.0;
  s->nfe += 1;
}

// after d2 = |f(y0 + h0 f0) - f0| / h0: the first step size
__global__ void dopri5_init_b_kernel(grand_dopri5_state* s, cudaGraphConditionalHandle loop) {
  const double d1 = rms(s->acc[1], s->numel);
  const double d2 = rms(s->acc[2], s->numel) / s->h0;
  double h1;
  if (d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax(1e-6, s->h0 * 1e-3);
  else h1 = (double)(float)pow(0.01 / fmax(d1, d2), 1.0 / 5.0);
  s->dt = fmin(100.0 * s->h0, fabs(h1));
  cudaGraphSetConditional(loop, (s->t_end > s->t && s->attempts < s->max_attempts) ? 1u : 0u);
}

// start of an attempt: all coefficients that depend on dt; products of two fp32 numbers rounded once, like
// torch's python-scalar * fp32-tensor arithmetic in torchdiffeq
__global__ void dopri5_pre_kernel(grand_dopri5_state* s) {
  const float h = (float)s->dt;
  if (!(s->t + s->dt > s->t)) s->status = 1;        // torchdiffeq asserts here: "underflow in dt"
  float* c = s->coef;
  c[OFF_LIN0 + 0] = 1.f;
  c[OFF_LIN0 + 1] = __fmul_rn((float)kBeta[0][0], h);
  for (int st = 0; st < 5; ++st) {                  // RHS launch st evaluates k[st+1] and writes the input of stage st+2
    float* cs = c + OFF_STAGE + st * 9;
    cs[0] = __fmul_rn((float)kBeta[st + 1][st + 1], h);       // coef_f: weight of the k just computed
    int n = 0;
    cs[1 + n++] = 1.f;                                          // ycur
    for (int j = 0; j <= st; ++j)
      if (kBeta[st + 1][j] != 0.0) cs[1 + n++] = __fmul_rn((float)kBeta[st + 1][j], h);
  }
  int n = 0;
  for (int j = 0; j < 7; ++j)
    if (kCErr[j] != 0.0) c[OFF_ERR + n++] = __fmul_rn(h, (float)kCErr[j]);
  n = 0;
  c[OFF_MID + n++] = 1.f;
  for (int j = 0; j < 7; ++j)
    if (kCMid[j] != 0.0) c[OFF_MID + n++] = __fmul_rn(h, (float)kCMid[j]);
  s->acc[0] = 0.0;
}

__global__ void dopri5_post_kernel(grand_dopri5_state* s, cudaGraphConditionalHandle loop) {
  const double ratio = rms(s->acc[0], s->numel);
  const bool accept = ratio <= 1.0;
  s->attempts += 1;
  s->nfe += 6;
  s->accept = accept ? 1 : 0;
  const double dt = s->dt;
  if (accept) {
    s->accepted += 1;
    s->t_prev = s->t;
    s->t = s->t + dt;
    s->have_interp = 1;
    s->coef[OFF_INTERP + 0] = (float)dt;            // (float) h of the accepted step, for the dense output (slot reused below)
  }
  double factor;
  if (ratio == 0.0) {
    factor = 10.0;
  } else {
    const double dfactor = ratio < 1.0 ? 1.0 : 0.2;
    factor = fmin(10.0, fmax(0.9 / pow(ratio, 1.0 / 5.0), dfactor));
  }
  s->dt = dt * factor;
  const bool go = s->t_end > s->t && s->attempts < s->max_attempts && s->status == 0;
  cudaGraphSetConditional(loop, go ? 1u : 0u);
}

// on accept: yprev <- ycur, fprev <- k0, ymid = ycur + sum c_mid k, ycur <- y1, k0 <- k6   (one pass over [n])
struct AdvanceArgs {
  const grand_dopri5_state* s;
  float *ycur, *yprev, *fprev, *ymid;
  const float* y1;
  float* k[7];
};
__global__ void __launch_bounds__(256) dopri5_advance_kernel(int64_t n, AdvanceArgs a) {
  if (a.s->accept == 0) return;
  float cm[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) cm[t] = a.s->coef[OFF_MID + t];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float y0 = a.ycur[i], f0 = a.k[0][i];
    float m = 0.f;
    m = fmaf(cm[0], y0, m);
    int t = 1;
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      if (j == 1) continue;                          // c_mid[1] == 0: the term is not in the list
      m = fmaf(cm[t++], a.k[j][i], m);
    }
    a.yprev[i] = y0;
    a.fprev[i] = f0;
    a.ymid[i] = m;
    a.ycur[i] = a.y1[i];
    a.k[0][i] = a.k[6][i];
  }
}

// dense output at t_end: _interp_evaluate(_interp_fit(y0, y1, y_mid, f0, f1, dt), t0, t1, t) as five weights
__global__ void dopri5_final_kernel(grand_dopri5_state* s) {
  float* w = s->coef + OFF_INTERP;
  const double h = (double)w[0];                     // (float) dt of the last accepted step (stored by post)
  if (s->have_interp == 0 || s->t == s->t_end || !(s->t_end < s->t)) {
    // no accepted step, or the last step landed exactly on t_end, or the attempts ran out before t_end: the state itself
    w[0] = 0.f; w[1] = 1.f; w[2] = 0.f; w[3] = 0.f; w[4] = 0.f;
    return;
  }
  const double x = (double)(float)((s->t_end - s->t_prev) / (s->t - s->t_prev));
  const double x2 = x * x, x3 = x * x * x, x4 = x * x * x * x;
  w[0] = (float)(1.0 + x2 * (-11.0) + x3 * 18.0 + x4 * (-8.0));
  w[1] = (float)(x2 * (-5.0) + x3 * 14.0 + x4 * (-8.0));
  w[2] = (float)(x2 * 16.0 + x3 * (-32.0) + x4 * 16.0);
  w[3] = (float)(h * (x + x2 * (-4.0) + x3 * 5.0 + x4 * (-2.0)));
  w[4] = (float)(h * (x2 * 1.0 + x3 * (-3.0) + x4 * 2.0));
}

struct DevTerms {
  int n;
  const float* p[GRAND_MAX_TERMS];
};
// y = sum_t coef_dev[t] * p[t]
__global__ void __launch_bounds__(256) lincomb_dev_kernel(int64_t n, DevTerms a, const float* __restrict__ coef_dev,
                                                          float* __restrict__ y) {
  float c[GRAND_MAX_TERMS];
#pragma unroll
  for (int t = 0; t < GRAND_MAX_TERMS; ++t) c[t] = t < a.n ? coef_dev[t] : 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = 0.f;
#pragma unroll
    for (int t = 0; t < GRAND_MAX_TERMS; ++t)
      if (t < a.n) v = fmaf(c[t], a.p[t][i], v);
    y[i] = v;
  }
}
// *acc += sum ((sum_t c_t p_t) / (atol + rtol max(|y0|, |y1|)))^2, c from the device table or all given
__global__ void __launch_bounds__(256)
sumsq_dev_kernel(int64_t n, DevTerms a, const float* __restrict__ coef_dev, float c0, float c1,
                 const float* __restrict__ y0, const float* __restrict__ y1, float atol, float rtol,
                 double* __restrict__ acc) {
  float c[GRAND_MAX_TERMS];
#pragma unroll
  for (int t = 0; t < GRAND_MAX_TERMS; ++t) c[t] = t < a.n ? (coef_dev != nullptr ? coef_dev[t] : (t == 0 ? c0 : c1)) : 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double local = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float err = 0.f;
#pragma unroll
    for (int t = 0; t < GRAND_MAX_TERMS; ++t)
      if (t < a.n) err = fmaf(c[t], a.p[t][i], err);
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
    double sum = 0.0;
    for (int w = 0; w < 8; ++w) sum += warp_part[w];
    atomicAdd(acc, sum);
  }
}

static int ew_grid(int64_t n) {
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 8;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace grand

using namespace grand;

struct grand_dopri5_plan {
  cudaGraph_t graph;
  cudaGraphExec_t exec;
};

#define CU_OK(call)                                                                     \
  do {                                                                                  \
    cudaError_t _e = (call);                                                            \
    if (_e != cudaSuccess) {                                                            \
      rc = ::grand::cuda_fail(_e, #call);                                               \
      goto fail;                                                                        \
    }                                                                                   \
  } while (0)
#define RC_OK(call)            \
  do {                         \
    rc = (call);               \
    if (rc != GRAND_OK) goto fail; \
  } while (0)

extern "C" {

int grand_dopri5_create(const grand_graph* g, const grand_attention* a, const float* w_csr, int32_t d,
                        const grand_workspace* ws, const grand_epilogue* residual, const float* y0, float* out,
                        float* planes, grand_dopri5_state* state, double t0, double t1, float rtol, float atol,
                        int32_t max_attempts, grand_dopri5_plan** plan_out) {
  GRAND_REQUIRE(g && residual && y0 && out && planes && state && plan_out && d > 0 && max_attempts > 0, GRAND_EINVAL,
                "dopri5_create: null or bad argument");
  GRAND_REQUIRE(a != nullptr || w_csr != nullptr || g->n_edges == 0, GRAND_EINVAL,
                "dopri5_create: neither attention nor Laplacian weights given");
  GRAND_REQUIRE(g->n_cols == g->n_rows, GRAND_EUNSUPPORTED, "dopri5_create: sharded graphs are not built");
  GRAND_REQUIRE(t1 > t0, GRAND_EUNSUPPORTED, "dopri5_create: forward integration only (t1 > t0)");
  const int64_t n = (int64_t)g->n_rows * d;
  float* P[13];
  for (int i = 0; i < 13; ++i) P[i] = planes + (size_t)i * n;
  float *ycur = P[0], *ya = P[1], *yb = P[2], *ymid = P[3], *yprev = P[4], *fprev = P[5];
  float* k[7];
  for (int i = 0; i < 7; ++i) k[i] = P[6 + i];
  int rc = GRAND_OK;
  cudaStream_t cs = nullptr, bs = nullptr;
  cudaGraph_t graph = nullptr, body = nullptr;
  cudaGraphExec_t exec = nullptr;
  cudaGraphConditionalHandle loop;
  const int grid = ew_grid(n);

  auto rhs = [&](cudaStream_t st, const float* x, float* f_out, float* y_out, const float* coef_dev, int n_terms,
                 const float* const* terms) -> int {
    grand_epilogue ep;
    memset(&ep, 0, sizeof(ep));
    ep.alpha_train = residual->alpha_train; ep.beta_train = residual->beta_train; ep.x0 = residual->x0;
    ep.no_alpha_sigmoid = residual->no_alpha_sigmoid;
    ep.f_out = f_out; ep.y_out = y_out; ep.coef_dev = coef_dev; ep.n_terms = n_terms;
    for (int t = 0; t < n_terms; ++t) { ep.term[t] = terms[t]; ep.coef[t] = 0.f; }
    if (a != nullptr) return rhs_attention_impl(g, a, x, ws, &ep, true, st);
    return rhs_laplacian_impl(g, d, x, w_csr, ws, &ep, st);
  };
  auto sumsq = [&](cudaStream_t st, int nt, const float* p0, const float* p1, const float* coef_dev, float c0, float c1,
                   const float* yref0, const float* yref1, double* acc) {
    DevTerms tm;
    memset(&tm, 0, sizeof(tm));
    tm.n = nt; tm.p[0] = p0; tm.p[1] = p1;
    sumsq_dev_kernel<<<grid, 256, 0, st>>>(n, tm, coef_dev, c0, c1, yref0, yref1, atol, rtol, acc);
    count_launch();
  };

  CU_OK(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
  CU_OK(cudaStreamCreateWithFlags(&bs, cudaStreamNonBlocking));
  CU_OK(cudaGraphCreate(&graph, 0));
  CU_OK(cudaGraphConditionalHandleCreate(&loop, graph, 1, cudaGraphCondAssignDefault));

  // ---- segment 1: reset, f0, initial step selection ----
  CU_OK(cudaStreamBeginCaptureToGraph(cs, graph, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
  dopri5_reset_kernel<<<1, 1, 0, cs>>>(state, t0, t1, rtol, atol, (int)n, max_attempts);
  CU_OK(cudaMemcpyAsync(ycur, y0, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, cs));
  RC_OK(rhs(cs, ycur, k[0], nullptr, nullptr, 0, nullptr));
  sumsq(cs, 1, ycur, nullptr, nullptr, 1.f, 0.f, ycur, ycur, &state->acc[0]);           // d0
  sumsq(cs, 1, k[0], nullptr, nullptr, 1.f, 0.f, ycur, ycur, &state->acc[1]);           // d1
  dopri5_init_a_kernel<<<1, 1, 0, cs>>>(state);
  {
    DevTerms tm;
    memset(&tm, 0, sizeof(tm));
    tm.n = 2; tm.p[0] = ycur; tm.p[1] = k[0];
    lincomb_dev_kernel<<<grid, 256, 0, cs>>>(n, tm, state->coef + OFF_INIT, ya);       // ya = ycur + h0 f0
  }
  RC_OK(rhs(cs, ya, k[1], nullptr, nullptr, 0, nullptr));
  sumsq(cs, 2, k[1], k[0], nullptr, 1.f, -1.f, ycur, ycur, &state->acc[2]);             // d2 * h0
  dopri5_init_b_kernel<<<1, 1, 0, cs>>>(state, loop);
  count_launch(6);

  // ---- the WHILE node and its body ----
  {
    cudaStreamCaptureStatus status;
    const cudaGraphNode_t* deps = nullptr;
    size_t n_deps = 0;
    cudaGraph_t cap_graph = nullptr;
    CU_OK(cudaStreamGetCaptureInfo_v2(cs, &status, nullptr, &cap_graph, &deps, &n_deps));
    cudaGraphNodeParams np = {cudaGraphNodeTypeConditional};
    np.type = cudaGraphNodeTypeConditional;
    np.conditional.handle = loop;
    np.conditional.type = cudaGraphCondTypeWhile;
    np.conditional.size = 1;
    cudaGraphNode_t while_node;
    CU_OK(cudaGraphAddNode(&while_node, graph, deps, n_deps, &np));
    body = np.conditional.phGraph_out[0];
    CU_OK(cudaStreamUpdateCaptureDependencies(cs, &while_node, 1, cudaStreamSetCaptureDependencies));

    CU_OK(cudaStreamBeginCaptureToGraph(bs, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
    dopri5_pre_kernel<<<1, 1, 0, bs>>>(state);
    {
      DevTerms tm;
      memset(&tm, 0, sizeof(tm));
      tm.n = 2; tm.p[0] = ycur; tm.p[1] = k[0];
      lincomb_dev_kernel<<<grid, 256, 0, bs>>>(n, tm, state->coef + OFF_LIN0, ya);
    }
    // stages: RHS launch st reads `inp`, writes k[st+1] and (st < 5) the next stage's input
    {
      float* inp = ya;
      float* outp = yb;
      for (int st = 0; st < 6; ++st) {
        if (st < 5) {
          const float* terms[GRAND_MAX_TERMS];
          int nt = 0;
          terms[nt++] = ycur;
          const double beta_host[6][6] = {{1. / 5, 0, 0, 0, 0, 0},
                                          {3. / 40, 9. / 40, 0, 0, 0, 0},
                                          {44. / 45, -56. / 15, 32. / 9, 0, 0, 0},
                                          {19372. / 6561, -25360. / 2187, 64448. / 6561, -212. / 729, 0, 0},
                                          {9017. / 3168, -355. / 33, 46732. / 5247, 49. / 176, -5103. / 18656, 0},
                                          {35. / 384, 0, 500. / 1113, 125. / 192, -2187. / 6784, 11. / 84}};
          for (int j = 0; j <= st; ++j)
            if (beta_host[st + 1][j] != 0.0) terms[nt++] = k[j];
          RC_OK(rhs(bs, inp, k[st + 1], outp, state->coef + OFF_STAGE + st * 9, nt, terms));
          float* tmp = inp; inp = outp; outp = tmp;
        } else {
          RC_OK(rhs(bs, inp, k[6], nullptr, nullptr, 0, nullptr));
        }
      }
      // inp == yb: the 6th stage's input is the 5th-order solution y1
      DevTerms tm;
      memset(&tm, 0, sizeof(tm));
      tm.n = 0;
      for (int j = 0; j < 7; ++j)
        if (j != 1) tm.p[tm.n++] = k[j];               // c_error[1] == 0
      sumsq_dev_kernel<<<grid, 256, 0, bs>>>(n, tm, state->coef + OFF_ERR, 0.f, 0.f, ycur, inp, atol, rtol, &state->acc[0]);
      dopri5_post_kernel<<<1, 1, 0, bs>>>(state, loop);
      AdvanceArgs adv;
      adv.s = state; adv.ycur = ycur; adv.yprev = yprev; adv.fprev = fprev; adv.ymid = ymid; adv.y1 = inp;
      for (int j = 0; j < 7; ++j) adv.k[j] = k[j];
      dopri5_advance_kernel<<<grid, 256, 0, bs>>>(n, adv);
      count_launch(5);
    }
    CU_OK(cudaStreamEndCapture(bs, nullptr));
  }

  // ---- segment 3: dense output at t1 ----
  dopri5_final_kernel<<<1, 1, 0, cs>>>(state);
  {
    DevTerms tm;
    memset(&tm, 0, sizeof(tm));
    tm.n = 5; tm.p[0] = yprev; tm.p[1] = ycur; tm.p[2] = ymid; tm.p[3] = fprev; tm.p[4] = k[0];
    lincomb_dev_kernel<<<grid, 256, 0, cs>>>(n, tm, state->coef + OFF_INTERP, out);
    count_launch(2);
  }
  CU_OK(cudaStreamEndCapture(cs, nullptr));
  CU_OK(cudaGraphInstantiate(&exec, graph, 0));
  cudaStreamDestroy(cs);
  cudaStreamDestroy(bs);
  {
    grand_dopri5_plan* p = new grand_dopri5_plan;
    p->graph = graph;
    p->exec = exec;
    *plan_out = p;
  }
  return GRAND_OK;

fail:
  if (bs != nullptr) {
    cudaStreamCaptureStatus st2;
    if (cudaStreamIsCapturing(bs, &st2) == cudaSuccess && st2 == cudaStreamCaptureStatusActive) cudaStreamEndCapture(bs, nullptr);
    cudaStreamDestroy(bs);
  }
  if (cs != nullptr) {
    cudaStreamCaptureStatus st2;
    if (cudaStreamIsCapturing(cs, &st2) == cudaSuccess && st2 == cudaStreamCaptureStatusActive) cudaStreamEndCapture(cs, nullptr);
    cudaStreamDestroy(cs);
  }
  if (exec != nullptr) cudaGraphExecDestroy(exec);
  if (graph != nullptr) cudaGraphDestroy(graph);
  return rc;
}

int grand_dopri5_launch(grand_dopri5_plan* plan, void* stream) {
  GRAND_REQUIRE(plan != nullptr && plan->exec != nullptr, GRAND_EINVAL, "dopri5_launch: null plan");
  GRAND_CUDA_OK(cudaGraphLaunch(plan->exec, (cudaStream_t)stream));
  return GRAND_OK;
}

int grand_dopri5_destroy(grand_dopri5_plan* plan) {
  if (plan == nullptr) return GRAND_OK;
  if (plan->exec != nullptr) cudaGraphExecDestroy(plan->exec);
  if (plan->graph != nullptr) cudaGraphDestroy(plan->graph);
  delete plan;
  return GRAND_OK;
}

}  // extern "C"
