// Small graphs: a WHOLE fixed-grid integration in ONE cooperative launch (SURVEY.md section 7 step 5, section 8(d) caveat).
//
// On Cora / Photo-sized graphs an RHS evaluation is a few microseconds of work; three launches per evaluation (projection,
// long-row statistics, edge kernel), each as long as its longest dependent chain, made 45-112 us of it (profiles/README.md).
// Here every block stays resident for the whole integration and walks, per evaluation,
//     projection of its rows (FFMA, [Wq ; Wk] in shared memory for the lifetime of the launch)   | grid barrier
//     statistics units of the long segments                                                      | grid barrier
//     the work list of csrc/rhs_stream.cu (same device code: windows, medium rows, long segments) | grid barrier
// with the Runge-Kutta stage algebra of grand_integrate_fixed evaluated by the kernel itself (stage pointers ping-pong,
// coefficients from dt[]).  Inputs of a stage are outputs of the previous one, written by OTHER blocks: this translation
// unit compiles the shared device code with GRAND_COHERENT_LOADS (every state / Q / K load is ld.global.cg), and the grid
// barrier is release / acquire at gpu scope with a timeout that raises an error flag instead of hanging the GPU.  The
// launch is cooperative (co-residency is what makes the barrier legal); graphs that do not fit that way, or whose
// configuration this kernel is not built for, take the per-evaluation path.
#define GRAND_COHERENT_LOADS 1
#define GRAND_STREAM_DEVICE_ONLY 1
#include "rhs_stream.cu"

namespace grand {

constexpr int kCoopThreads = 256;
constexpr int kCoopMaxSteps = 128;

struct CoopArgs {
  RhsArgs base;               // graph, attention, residual parameters, workspace pointers; x / ep outputs are set per stage
  const float *wq, *bq, *wk, *bk;
  int method, n_steps;
  float dt[kCoopMaxSteps];
  float *y, *yb, *yc, *k1, *k2, *k3;
  unsigned* sync;             // [0] arrivals, [1] generation, [2] error
};

__device__ __forceinline__ void grid_barrier(unsigned* sync, unsigned& gen) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned arrived = atomicAdd(sync, 1u) + 1u;
    if (arrived == gridDim.x) {
      atomicExch(sync, 0u);
      __threadfence();
      atomicAdd(sync + 1, 1u);                     // release the generation
    } else {
      const long long t0 = clock64();
      unsigned g;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(g) : "l"(sync + 1) : "memory");
        if (g != gen) break;
        if (clock64() - t0 > 4000000000ll) {       // ~2 s: never on a co-resident grid; do not hang the GPU
          atomicExch(sync + 2, 1u);
          break;
        }
      } while (true);
    }
    __threadfence();
  }
  gen += 1;
  __syncthreads();
}

// q = Wq y + bq, k = Wk y + bk for one row held 4 R columns per lane; [Wq ; Wk] in shared memory.  32 outputs at a time:
// partial sums over the lane's own columns, then a 31-shuffle transposing butterfly leaves output o in lane o.
template <bool FULL, int R>
__device__ __forceinline__ void project_row(const RhsArgs& A, const CoopArgs& C, const float* __restrict__ wsm, int row,
                                            int lane, const float4 (&y)[R]) {
  const int d = A.d, ad = A.a.ad_eff, n_out = 2 * ad;
  for (int o0 = 0; o0 < n_out; o0 += kWarp) {
    float p[kWarp];
#pragma unroll
    for (int oo = 0; oo < kWarp; ++oo) {
      p[oo] = 0.f;
      if (o0 + oo < n_out) {
        const float* w = wsm + (size_t)(o0 + oo) * d + lane * 4;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (FULL || (lane + r * kWarp) * 4 < d) {
            const float4 wv = *reinterpret_cast<const float4*>(w + r * kWarp * 4);
            p[oo] = fmaf(wv.x, y[r].x, p[oo]); p[oo] = fmaf(wv.y, y[r].y, p[oo]);
            p[oo] = fmaf(wv.z, y[r].z, p[oo]); p[oo] = fmaf(wv.w, y[r].w, p[oo]);
          }
        }
      }
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
      const bool up = (lane & s) != 0;
#pragma unroll
      for (int i = 0; i < s; ++i) {
        const float send = up ? p[i] : p[i + s];
        const float keep = up ? p[i + s] : p[i];
        p[i] = keep + __shfl_xor_sync(kFull, send, s);
      }
    }
    const int o = o0 + lane;                       // p[0] is the full sum of output o
    if (o < ad) const_cast<float*>(A.q)[(size_t)row * ad + o] = p[0] + __ldg(C.bq + o);
    else if (o < n_out) const_cast<float*>(A.k)[(size_t)row * ad + (o - ad)] = p[0] + __ldg(C.bk + (o - ad));
  }
}

template <int MODE, bool FULL, int R, int HM>
__global__ void __launch_bounds__(kCoopThreads, 2) integrate_coop_kernel(const __grid_constant__ CoopArgs C) {
  extern __shared__ __align__(16) float wsm[];       // [2 ad, d]: Wq rows then Wk rows (MODE_FUSED)
  __shared__ RhsArgs sA;
  __shared__ int s_next;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  NoRing no_ring;
  const int warps_per_block = kCoopThreads / kWarp;
  if (MODE == MODE_FUSED) {
    const int nw = C.base.a.ad_eff * C.base.d;
    for (int i = threadIdx.x * 4; i < nw; i += kCoopThreads * 4) {
      *reinterpret_cast<float4*>(wsm + i) = __ldg(reinterpret_cast<const float4*>(C.wq + i));
      *reinterpret_cast<float4*>(wsm + nw + i) = __ldg(reinterpret_cast<const float4*>(C.wk + i));
    }
  }
  unsigned gen = 0;
  if (threadIdx.x == 0) {
    unsigned g;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(g) : "l"(C.sync + 1) : "memory");
    s_next = (int)g;
  }
  __syncthreads();
  gen = (unsigned)s_next;                            // nobody passes the first barrier before this block arrives at it
  __syncthreads();

  float alpha = __ldg(C.base.ep.alpha_train);
  if (!C.base.ep.no_alpha_sigmoid) alpha = 1.f / (1.f + expf(-alpha));
  const float beta = C.base.ep.x0 != nullptr ? __ldg(C.base.ep.beta_train) : 0.f;
  const uint64_t pol = policy_evict_normal();
  const int nL = C.base.g.n_lunits, n_seg = nL + C.base.g.n_munits;
  const int n_units = n_seg + (C.base.g.n_tiles - C.base.g.n_split_tiles);
  const int per_step = C.method == GRAND_EULER ? 1 : (C.method == GRAND_MIDPOINT ? 2 : 4);

  float* cur = C.y;
  float* nxt = C.yc;
  for (int step = 0; step < C.n_steps; ++step) {
    const float h = C.dt[step];
    for (int st = 0; st < per_step; ++st) {
      // ---- this evaluation's arguments (same stage algebra as grand_integrate_fixed) ----
      if (threadIdx.x == 0) {
        sA = C.base;
        EpiView& ep = sA.ep;
        ep.f_out = nullptr; ep.y_out = nullptr; ep.n_terms = 0; ep.coef_dev = nullptr;
        const float third = 1.f / 3.f;
        const float* x = cur;
        if (C.method == GRAND_EULER) {
          ep.y_out = nxt; ep.coef_f = h; ep.n_terms = 1; ep.term[0] = cur; ep.coef[0] = 1.f;
        } else if (C.method == GRAND_MIDPOINT) {
          if (st == 0) { ep.y_out = C.yb; ep.coef_f = 0.5f * h; ep.n_terms = 1; ep.term[0] = cur; ep.coef[0] = 1.f; }
          else { x = C.yb; ep.y_out = nxt; ep.coef_f = h; ep.n_terms = 1; ep.term[0] = cur; ep.coef[0] = 1.f; }
        } else if (st == 0) {
          ep.f_out = C.k1; ep.y_out = C.yb; ep.coef_f = h * third; ep.n_terms = 1; ep.term[0] = cur; ep.coef[0] = 1.f;
        } else if (st == 1) {
          x = C.yb; ep.f_out = C.k2; ep.y_out = nxt; ep.coef_f = h; ep.n_terms = 2;
          ep.term[0] = cur; ep.coef[0] = 1.f; ep.term[1] = C.k1; ep.coef[1] = -h * third;
        } else if (st == 2) {
          x = nxt; ep.f_out = C.k3; ep.y_out = C.yb; ep.coef_f = h; ep.n_terms = 3;
          ep.term[0] = cur; ep.coef[0] = 1.f; ep.term[1] = C.k1; ep.coef[1] = h; ep.term[2] = C.k2; ep.coef[2] = -h;
        } else {
          x = C.yb; ep.y_out = nxt; ep.coef_f = h * 0.125f; ep.n_terms = 4;
          ep.term[0] = cur; ep.coef[0] = 1.f; ep.term[1] = C.k1; ep.coef[1] = h * 0.125f;
          ep.term[2] = C.k2; ep.coef[2] = 3.f * h * 0.125f; ep.term[3] = C.k3; ep.coef[3] = 3.f * h * 0.125f;
        }
        sA.x = x;
        s_next = 0;
      }
      __syncthreads();
      const RhsArgs& A = sA;

      if (MODE == MODE_FUSED) {
        // ---- projection of this block's share of the rows ----
        const int gw = blockIdx.x * warps_per_block + warp, n_gw = gridDim.x * warps_per_block;
        for (int row = gw; row < A.g.n_rows; row += n_gw) {
          float4 yv[R];
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const int c = (lane + r * kWarp) * 4;
            yv[r] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (FULL || c < A.d) yv[r] = __ldcg(reinterpret_cast<const float4*>(A.x + (size_t)row * A.d + c));
          }
          project_row<FULL, R>(A, C, wsm, row, lane, yv);
        }
        grid_barrier(C.sync, gen);
        if (nL > 0) {
          // ---- statistics units of the long segments ----
          for (;;) {
            int v = 0;
            if (lane == 0) v = atomicAdd(&s_next, 1);
            v = __shfl_sync(kFull, v, 0);
            const int u = v * (int)gridDim.x + (int)blockIdx.x;
            if (u >= nL) break;
            const int4 desc = __ldg(A.g.seg_desc + u);
            process_segment<MODE, FULL, R, HM, true>(A, desc, u, lane, alpha, beta, pol, pol, no_ring);
          }
          __syncthreads();
          if (threadIdx.x == 0) s_next = 0;
          grid_barrier(C.sync, gen);
        }
      }

      // ---- the work list ----
      for (;;) {
        int v = 0;
        if (lane == 0) v = atomicAdd(&s_next, 1);
        v = __shfl_sync(kFull, v, 0);
        const int u = v * (int)gridDim.x + (int)blockIdx.x;
        if (u >= n_units) break;
        if (u < n_seg) {
          const int4 desc = __ldg(A.g.seg_desc + u);
          process_segment<MODE, FULL, R, HM, false>(A, desc, u, lane, alpha, beta, pol, pol, no_ring);
        } else {
          const int t = A.g.n_split_tiles + (u - n_seg);
          const int4 hdr = __ldg(A.g.win_hdr + t);
          const int2 rec = __ldg(A.g.win_rec + (size_t)t * kWarp + lane);
          process_window<MODE, FULL, R>(A, hdr, rec, lane, alpha, beta, pol, pol);
        }
      }
      grid_barrier(C.sync, gen);
    }
    float* tmp = cur; cur = nxt; nxt = tmp;
  }
}

template <int MODE, bool FULL, int R, int HM>
static int launch_coop_inst(const CoopArgs& C, cudaStream_t stream) {
  const void* kern = (const void*)integrate_coop_kernel<MODE, FULL, R, HM>;
  const size_t smem = MODE == MODE_FUSED ? (size_t)2 * C.base.a.ad_eff * C.base.d * sizeof(float) : 0;
  if (smem > 48 * 1024) GRAND_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  GRAND_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kCoopThreads, smem));
  if (per_sm < 1) return GRAND_EUNSUPPORTED;
  const int n_units = C.base.g.n_lunits + C.base.g.n_munits + (C.base.g.n_tiles - C.base.g.n_split_tiles);
  const int warps_per_block = kCoopThreads / kWarp;
  int want = (n_units + warps_per_block - 1) / warps_per_block;
  const int rows_blocks = (C.base.g.n_rows + warps_per_block - 1) / warps_per_block;
  if (MODE == MODE_FUSED && rows_blocks > want) want = rows_blocks;
  int grid = per_sm * num_sms();
  if (grid > want) grid = want;
  if (grid < 1) grid = 1;
  void* args[] = {(void*)&C};
  GRAND_CUDA_OK(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(kCoopThreads), args, smem, stream));
  count_launch();
  return GRAND_OK;
}

template <int MODE>
static int launch_coop_mode(const CoopArgs& C, cudaStream_t stream) {
  const int d = C.base.d;
  const int chunks = (d / 4 + kWarp - 1) / kWarp;
  const bool full = d == chunks * kWarp * 4;
  const bool h8 = MODE == MODE_FUSED && C.base.a.heads > 4;
  if (chunks > 1) return GRAND_EUNSUPPORTED;           // small graphs of the reference have d <= 128
  if (full) return h8 ? launch_coop_inst<MODE, true, 1, GRAND_MAX_HEADS>(C, stream) : launch_coop_inst<MODE, true, 1, 4>(C, stream);
  return h8 ? launch_coop_inst<MODE, false, 1, GRAND_MAX_HEADS>(C, stream) : launch_coop_inst<MODE, false, 1, 4>(C, stream);
}

// Returns GRAND_EUNSUPPORTED (without setting an error) when this configuration / size is not for the cooperative kernel;
// the caller then issues the per-evaluation launches.  `A` carries graph, attention view, residual parameters and workspace.
int integrate_fixed_coop(const RhsArgs& A, bool attention, const grand_attention* a, int method, const float* dt, int n_steps,
                         float* y, float* stage_buffers, int* sync, cudaStream_t stream) {
  static int enabled = -1;
  if (enabled < 0) {
    // OFF unless GRAND_COOP=1.  Measured on B200 (profiles/r02_small_graphs.txt): parity-green (the whole GPU suite ran through
    // it), but SLOWER than one launch group per evaluation -- Cora 80 vs 47 us per evaluation, Cora GRAND-l 20 vs 15: what a
    // small graph pays per evaluation is the dependent chain of its longest unit (a 64-edge row = ~10 sequential round
    // trips), not the launches; the barriers and the L2-only loads come on top.  Kept as the measured answer to "one
    // persistent kernel per integration" (SURVEY.md section 7 step 5).
    const char* e = getenv("GRAND_COOP");
    int coop = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    enabled = (coop != 0 && e != nullptr && atoi(e) == 1) ? 1 : 0;
  }
  if (!enabled || sync == nullptr || n_steps <= 0 || n_steps > kCoopMaxSteps) return GRAND_EUNSUPPORTED;
  if ((A.d & 3) != 0 || A.d > 128 || A.g.n_cols != A.g.n_rows) return GRAND_EUNSUPPORTED;
  const int n_units = A.g.n_lunits + A.g.n_munits + (A.g.n_tiles - A.g.n_split_tiles);
  if (n_units > 32768 || A.g.n_rows > 65536) return GRAND_EUNSUPPORTED;   // large graphs: bandwidth, not launches, is the issue
  if (A.g.n_split_tiles > 0 && A.g.seg_desc == nullptr) return GRAND_EUNSUPPORTED;
  if (A.g.n_lunits > 0 && (A.lpartial == nullptr || A.lcounter == nullptr || A.g.lrow_slot0 == nullptr)) return GRAND_EUNSUPPORTED;
  CoopArgs C;
  memset(&C, 0, sizeof(C));
  C.base = A;
  C.method = method; C.n_steps = n_steps;
  for (int s = 0; s < n_steps; ++s) C.dt[s] = dt[s];
  const size_t plane = (size_t)A.g.n_cols * A.d;
  C.y = y; C.yb = stage_buffers; C.yc = stage_buffers + plane;
  C.k1 = stage_buffers + 2 * plane; C.k2 = stage_buffers + 3 * plane; C.k3 = stage_buffers + 4 * plane;
  C.sync = reinterpret_cast<unsigned*>(sync);
  int rc;
  if (attention) {
    const bool vec4 = ((A.a.dk & 3) == 0) && ((A.a.ad_eff & 3) == 0);
    const bool fused = A.a.score_type == GRAND_SCORE_SCALED_DOT && A.a.norm_idx == 0 && A.a.norm_type == GRAND_NORM_SOFTMAX;
    if (!vec4 || !fused || (size_t)2 * A.a.ad_eff * A.d * sizeof(float) > 96 * 1024 || A.lrow_stats == nullptr && A.g.n_lunits > 0)
      return GRAND_EUNSUPPORTED;
    C.wq = a->wq; C.bq = a->bq; C.wk = a->wk; C.bk = a->bk;
    rc = launch_coop_mode<MODE_FUSED>(C, stream);
  } else {
    rc = launch_coop_mode<MODE_WEIGHT>(C, stream);
  }
  if (rc != GRAND_OK) return rc;
  // the result is in `cur` after n_steps swaps: y for an even number of steps, yc otherwise
  if ((n_steps & 1) != 0) {
    GRAND_CUDA_OK(cudaMemcpyAsync(y, C.yc, plane * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    count_launch();
  }
  return GRAND_OK;
}

}  // namespace grand
