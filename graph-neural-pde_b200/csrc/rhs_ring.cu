// The dominant kernel, third design: the persistent work-list kernel of rhs_stream.cu with the neighbour rows STAGED
// THROUGH SHARED MEMORY BY TMA BULK COPIES (cp.async.bulk, SASS UBLKCP) -- north_star's data path.
//
//   Every warp owns one stage of 32 row slots (16 KB at d = 128) and one mbarrier.  As soon as the column indices of a
//   32-edge chunk (a packed window, or 32 consecutive edges of a medium / long row) are known, every lane issues the bulk
//   copy of its edge's neighbour row; the copies fly while the warp loads the chunk's K / Q rows and the first rows'
//   epilogue operands into registers and computes scores, softmax and head mean.  Then ONE wait on the mbarrier and the
//   rows are accumulated from shared memory in CSR order (LDS.128 per lane, conflict-free), whatever the row boundaries
//   inside the chunk are.  Compared with the register gathers of rhs_stream.cu (one dependent round trip per batch of
//   <= 8 rows of ONE destination row, i.e. 4-6 trips per window on a power-law degree distribution), a chunk is ONE
//   round trip, and the bytes a warp keeps in flight are bounded by its shared-memory stage, not by its registers:
//   14 warps x 16 KB per SM (tools/tma_probe.cu: 14 TB/s of L2-resident 512-byte rows with this shape on B200, against
//   the 7.3 TB/s the register-gather kernel pulls).
//
//   Everything else -- work list, unit kinds, block-level dynamic unit grabbing, arithmetic and its order (so results are
//   bit-identical to rhs_stream.cu), fused epilogue, multi-GPU push -- is the shared device code of rhs_stream.cu.
//   Built for rows of at most 512 bytes (d <= 128, d % 4 == 0); wider states keep the register path.
#define GRAND_STREAM_DEVICE_ONLY 1
#include "rhs_stream.cu"

namespace grand {

#ifndef GRAND_RING_WARPS
#define GRAND_RING_WARPS 14         // ONE block per SM: 14 warps x 16 KB = 224 KB of stages, 128 registers per thread
#endif
#ifndef GRAND_RING_PIPELINE
#define GRAND_RING_PIPELINE 0       // 1: software-pipelined windows (half stages, next window's operands prefetched): needs 168
                                    // registers, i.e. 12 warps -- measured no faster than the plain version at 12 warps
#endif
#ifndef GRAND_RING_EB
#define GRAND_RING_EB 2             // rows whose epilogue operands are requested together (plain version)
#endif
// register budget: __launch_bounds__(threads, 1) lets ptxas pick what certainly launches (128 at 14 warps; 144 x 448
// threads = 63 K registers is refused by the launch: "too many resources"); GRAND_RING_MAXREG forces a cap instead
#ifdef GRAND_RING_MAXREG
#define GRAND_RING_BOUNDS __maxnreg__(GRAND_RING_MAXREG)
#else
#define GRAND_RING_BOUNDS __launch_bounds__(GRAND_RING_WARPS * 32, 1)
#endif
constexpr int kRingWarps = GRAND_RING_WARPS;
constexpr int kRingThreads = kRingWarps * kWarp;
constexpr int kRingStageBytes = kWarp * 512;

struct RingState {
  static constexpr bool enabled = true;
  const float4* slots;   // the warp's stage (generic address)
  uint32_t slots_s;      // the same, shared-window address for the async proxy
  uint32_t bar_s;        // whole stage: a 32-edge chunk of a medium / long row
  uint32_t bar_a, bar_b; // slots [0, 16) and [16, 32) of a packed window: the halves are refilled independently
  uint32_t ph_s, ph_a, ph_b;

  // rows x[j] of the lanes [lo, hi) -> their slots.  The __syncwarp orders every lane's reads of what the slots held
  // before the async-proxy writes.  One arrival (with the byte count) + the copies' complete_tx finish the phase.
  __device__ __forceinline__ void copy_rows(const float* x, int d, int j, int lane, int lo, int hi, uint32_t bar,
                                            uint64_t pol) {
    __syncwarp();
    const uint32_t row_bytes = (uint32_t)d * 4u;
    if (lane == lo) mbar_expect_tx(bar, (uint32_t)(hi - lo) * row_bytes);
    if (lane >= lo && lane < hi) bulk_copy_g2s(slots_s + (uint32_t)lane * row_bytes, x + (size_t)j * d, row_bytes, bar, pol);
  }
  // interface of process_segment (rhs_stream.cu): one chunk of n <= 32 edges per round trip
  __device__ __forceinline__ void issue(const float* x, int d, int j, bool, int n, int lane, uint64_t pol) {
    copy_rows(x, d, j, lane, 0, n, bar_s, pol);
  }
  __device__ __forceinline__ void wait() {
    mbar_wait(bar_s, ph_s);
    ph_s ^= 1u;
  }
  // acc += sum_{e in [e_lo, e_hi)} a_e * slot_e, in edge order (a_e lives in lane e)
  template <bool FULL>
  __device__ __forceinline__ void accumulate(float ab, int e_lo, int e_hi, int d, int lane, float (&acc)[4]) {
    const int q = d >> 2;
    const float4* s = slots + lane;
#pragma unroll 4
    for (int e = e_lo; e < e_hi; ++e) {
      const float a = __shfl_sync(kFull, ab, e);
      if (FULL || lane < q) {
        const float4 v = s[e * q];
        fma2(acc[0], acc[1], a, v.x, v.y);
        fma2(acc[2], acc[3], a, v.z, v.w);
      }
    }
  }
};

// what a window needs in registers before its scores can be computed
struct WinLoads {
  float4 qv[4], kv[4];   // MODE_FUSED, d_k = 4 x 4 heads: the lane's Q / K rows
  float w;               // MODE_WEIGHT
};

// ---- a packed window, one round trip: all of its neighbour rows (whole stage), K / Q and the first rows' epilogue
//      operands are requested together; later rows' operands are requested GRAND_RING_EB rows at a time -------------------
template <int MODE, bool FULL>
__device__ __forceinline__ void ring_window(const RhsArgs& A, const int4 hdr, const int2 rec, int lane, float alpha,
                                            float beta, uint64_t pol_keep, uint64_t pol_stream, RingState& ring) {
  const int d = A.d;
  const int row0 = hdr.x, beg = hdr.y, nrows = hdr.z & 0xff, cnt = hdr.z >> 8;
  const int j = rec.x;
  const int kk = rec.y & 0xff, lo = (rec.y >> 8) & 0x3f, hi = (rec.y >> 14) & 0x3f;
  const bool valid = lane < cnt;
  unsigned rowoff = (unsigned)row0 * (unsigned)d;
  // (requesting the K / Q rows before the copy loop instead of after it was measured slower: 274 vs 265 us)
  EpiRegs<1> E0;
  epi_load<FULL, 1>(A, rowoff, lane, pol_stream, E0);
#if GRAND_RING_EB > 1
  EpiRegs<1> E1;
  if (nrows > 1) epi_load<FULL, 1>(A, rowoff + d, lane, pol_stream, E1);
#endif
  ring.copy_rows(A.x, d, j, lane, 0, cnt, ring.bar_s, pol_keep);
  const float ab = window_weight<MODE>(A, lane, row0 + kk, j, beg + lane, valid, lo, hi);
  ring.wait();
  for (int kr = 0; kr < nrows; kr += GRAND_RING_EB, rowoff += GRAND_RING_EB * d) {
#if GRAND_RING_EB > 1
    const bool two = kr + 1 < nrows;
#endif
    if (kr > 0) {
      epi_load<FULL, 1>(A, rowoff, lane, pol_stream, E0);
#if GRAND_RING_EB > 1
      if (two) epi_load<FULL, 1>(A, rowoff + d, lane, pol_stream, E1);
#endif
    }
    const int r0 = __shfl_sync(kFull, rec.y, kr);
    float acc0[1][4] = {{0.f, 0.f, 0.f, 0.f}};
    ring.accumulate<FULL>(ab, (r0 >> 20) & 0x3f, (r0 >> 26) & 0x3f, d, lane, acc0[0]);
#if GRAND_RING_EB > 1
    const int r1 = __shfl_sync(kFull, rec.y, two ? kr + 1 : kr);
    float acc1[1][4] = {{0.f, 0.f, 0.f, 0.f}};
    if (two) ring.accumulate<FULL>(ab, (r1 >> 20) & 0x3f, (r1 >> 26) & 0x3f, d, lane, acc1[0]);
#endif
    epi_finish<FULL, 1>(A, rowoff, lane, alpha, beta, pol_stream, acc0, E0);
#if GRAND_RING_EB > 1
    if (two) epi_finish<FULL, 1>(A, rowoff + d, lane, alpha, beta, pol_stream, acc1, E1);
#endif
  }
}

// ---- first stage of a window, run while the PREVIOUS window is still being finished: request its Q / K rows (or weight)
//      and its first row's epilogue operands into registers, and the neighbour rows of its edges [0, 16) into the lower
//      half of the stage (free as soon as the previous window's edges [0, 16) are consumed) -------------------------------
template <int MODE, bool FULL>
__device__ __forceinline__ void win_stage1(const RhsArgs& A, const int4 hdr, const int2 rec, int lane, bool dk4,
                                           uint64_t pol_keep, uint64_t pol_stream, RingState& ring, WinLoads& L,
                                           EpiRegs<1>& E0) {
  const int d = A.d;
  const int row0 = hdr.x, beg = hdr.y, cnt = hdr.z >> 8;
  const int j = rec.x, kk = rec.y & 0xff;
  if (MODE == MODE_FUSED && dk4) {           // idle lanes read row0's Q row and K row 0: harmless, no predicates
    load_row16(A.q + (size_t)(row0 + kk) * A.a.ad_eff, L.qv);
    load_row16(A.k + (size_t)j * A.a.ad_eff, L.kv);
  }
  if (MODE == MODE_WEIGHT) L.w = lane < cnt ? __ldg(A.w_csr + beg + lane) : 0.f;
  epi_load<FULL, 1>(A, (unsigned)row0 * (unsigned)d, lane, pol_stream, E0);
  ring.copy_rows(A.x, d, j, lane, 0, cnt < 16 ? cnt : 16, ring.bar_a, pol_keep);
}

// ---- the packed windows of this warp, software-pipelined: while window w is scored and accumulated, the rows of w's
//      edges [16, 32) and (from the moment w's lower half is consumed) everything window w+1 needs first are in flight ----
template <int MODE, bool FULL, class Grab, class Fetch>
__device__ __forceinline__ void ring_windows(const RhsArgs& A, int4 hdr, int2 rec, int lane, float alpha, float beta,
                                             uint64_t pol_keep, uint64_t pol_stream, RingState& ring, Grab grab,
                                             Fetch fetch, int n_units) {
  const int d = A.d;
  const bool vec4 = ((A.a.dk & 3) == 0) && ((A.a.ad_eff & 3) == 0);
  const bool dk4 = MODE == MODE_FUSED && A.a.dk == 4 && vec4 && A.a.heads == 4;
  WinLoads L;
  EpiRegs<1> E0;
  win_stage1<MODE, FULL>(A, hdr, rec, lane, dk4, pol_keep, pol_stream, ring, L, E0);
  for (;;) {
    const int un = grab();
    int4 hdrn = make_int4(0, 0, 0, 0);
    int2 recn = make_int2(0, 0);
    fetch(un, hdrn, recn);
    const int row0 = hdr.x, beg = hdr.y, nrows = hdr.z & 0xff, cnt = hdr.z >> 8;
    const int j = rec.x;
    const int kk = rec.y & 0xff, lo = (rec.y >> 8) & 0x3f, hi = (rec.y >> 14) & 0x3f;
    const bool valid = lane < cnt;
    if (cnt > 16) ring.copy_rows(A.x, d, j, lane, 16, cnt, ring.bar_b, pol_keep);
    unsigned rowoff = (unsigned)row0 * (unsigned)d;
    EpiRegs<1> E1;
    if (nrows > 1) epi_load<FULL, 1>(A, rowoff + d, lane, pol_stream, E1);
    float ab;
    if (MODE == MODE_WEIGHT) {
      ab = L.w;
    } else if (MODE == MODE_FUSED && dk4) {
      const float ew = (valid && A.a.edge_weight_csr != nullptr) ? __ldg(A.a.edge_weight_csr + beg + lane) : 1.f;
      ab = dk4_weight(L.qv, L.kv, valid, ew, hi - lane, lo);
    } else {
      ab = window_weight<MODE>(A, lane, row0 + kk, j, beg + lane, valid, lo, hi);
    }
    mbar_wait(ring.bar_a, ring.ph_a);
    ring.ph_a ^= 1u;
    bool need_b = cnt > 16, pre = false;
    const int a_end = cnt < 16 ? cnt : 16;
    WinLoads Ln;
    EpiRegs<1> E0n;
    for (int kr = 0; kr < nrows; kr += 2, rowoff += 2 * d) {
      const bool two = kr + 1 < nrows;
      if (kr > 0) {
        epi_load<FULL, 1>(A, rowoff, lane, pol_stream, E0);
        if (two) epi_load<FULL, 1>(A, rowoff + d, lane, pol_stream, E1);
      }
      const int r0 = __shfl_sync(kFull, rec.y, kr);
      const int r1 = __shfl_sync(kFull, rec.y, two ? kr + 1 : kr);
      const int e_end = (r1 >> 26) & 0x3f;
      if (need_b && e_end > 16) {
        mbar_wait(ring.bar_b, ring.ph_b);
        ring.ph_b ^= 1u;
        need_b = false;
      }
      float acc0[1][4] = {{0.f, 0.f, 0.f, 0.f}}, acc1[1][4] = {{0.f, 0.f, 0.f, 0.f}};
      ring.accumulate<FULL>(ab, (r0 >> 20) & 0x3f, (r0 >> 26) & 0x3f, d, lane, acc0[0]);
      if (two) ring.accumulate<FULL>(ab, (r1 >> 20) & 0x3f, e_end, d, lane, acc1[0]);
      if (!pre && e_end >= a_end) {          // the lower half is consumed: hand it to the next window
        pre = true;
        if (un < n_units) win_stage1<MODE, FULL>(A, hdrn, recn, lane, dk4, pol_keep, pol_stream, ring, Ln, E0n);
      }
      epi_finish<FULL, 1>(A, rowoff, lane, alpha, beta, pol_stream, acc0, E0);
      if (two) epi_finish<FULL, 1>(A, rowoff + d, lane, alpha, beta, pol_stream, acc1, E1);
    }
    if (un >= n_units) break;
    hdr = hdrn;
    rec = recn;
    L = Ln;
    E0 = E0n;
  }
}

template <int MODE, bool FULL, int HM>
__global__ void GRAND_RING_BOUNDS rhs_ring_kernel(const __grid_constant__ RhsArgs A) {
  extern __shared__ __align__(128) unsigned char ring_smem[];
  __shared__ int s_next;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  RingState ring;
  ring.slots = reinterpret_cast<const float4*>(ring_smem + (size_t)warp * kRingStageBytes);
  ring.slots_s = smem_u32(ring.slots);
  ring.bar_s = smem_u32(ring_smem + (size_t)kRingWarps * kRingStageBytes + warp * 24);
  ring.bar_a = ring.bar_s + 8;
  ring.bar_b = ring.bar_s + 16;
  ring.ph_s = ring.ph_a = ring.ph_b = 0u;
  if (threadIdx.x == 0) s_next = 0;
  if (lane == 0) {
    mbar_init(ring.bar_s, 1);
    mbar_init(ring.bar_a, 1);
    mbar_init(ring.bar_b, 1);
  }
  __syncthreads();
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

  pdl_launch_dependents();
  int u = grab();
  int4 h4 = make_int4(0, 0, 0, 0);
  int2 rec = make_int2(0, 0);
  fetch(u, h4, rec);                       // the work list is not written by any kernel of the evaluation
  pdl_wait();                              // Q / K / statistics of this evaluation, and the previous stage's rows, are complete
  while (u < n_seg) {                      // medium rows and long-row segments come first in the list
    const int un = grab();
    int4 h4n = make_int4(0, 0, 0, 0);
    int2 recn = make_int2(0, 0);
    fetch(un, h4n, recn);
    process_segment<MODE, FULL, 1, HM, false, RingState>(A, h4, u, lane, alpha, beta, pol_keep, pol_stream, ring);
    u = un;
    h4 = h4n;
    rec = recn;
  }
#if GRAND_RING_PIPELINE
  if (u < n_units) ring_windows<MODE, FULL>(A, h4, rec, lane, alpha, beta, pol_keep, pol_stream, ring, grab, fetch, n_units);
#else
  while (u < n_units) {
    const int un = grab();
    int4 h4n = make_int4(0, 0, 0, 0);
    int2 recn = make_int2(0, 0);
    fetch(un, h4n, recn);
    ring_window<MODE, FULL>(A, h4, rec, lane, alpha, beta, pol_keep, pol_stream, ring);
    u = un;
    h4 = h4n;
    rec = recn;
  }
#endif
  if (A.push.ptr != nullptr) __threadfence_system();     // every row this thread delivered to a peer is visible there
}

// ---- launchers ------------------------------------------------------------------------------------------------------------
template <int MODE, bool FULL, int HM>
static int launch_ring_inst(const RhsArgs& A, cudaStream_t stream) {
  const int n_units = A.g.n_lunits + A.g.n_munits + (A.g.n_tiles - A.g.n_split_tiles);
  if (n_units <= 0) return GRAND_OK;
  int grid = (n_units + kRingWarps - 1) / kRingWarps;
  const int cap = num_sms();
  if (grid > cap) grid = cap;
  const int smem = kRingWarps * kRingStageBytes + kRingWarps * 24;
  // per device and cheap: set on every launch (no process-wide "already done" flag)
  GRAND_CUDA_OK(cudaFuncSetAttribute(rhs_ring_kernel<MODE, FULL, HM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  GRAND_CUDA_OK(cudaFuncSetAttribute(rhs_ring_kernel<MODE, FULL, HM>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  GRAND_CUDA_OK(launch_pdl(rhs_ring_kernel<MODE, FULL, HM>, dim3(grid), dim3(kRingThreads), (size_t)smem, stream, A));
  count_launch();
  return GRAND_OK;
}

template <int MODE>
static int launch_ring_mode_t(const RhsArgs& A, cudaStream_t stream) {
  const bool full = A.d == kWarp * 4;
  if (MODE == MODE_FUSED && A.a.heads > 4)
    return full ? launch_ring_inst<MODE, true, GRAND_MAX_HEADS>(A, stream) : launch_ring_inst<MODE, false, GRAND_MAX_HEADS>(A, stream);
  return full ? launch_ring_inst<MODE, true, 4>(A, stream) : launch_ring_inst<MODE, false, 4>(A, stream);
}

// the ring kernel takes the launch when the stream kernel would and a state row fits one 512-byte slot.
// GRAND_EDGE=stream / ring: A/B against the register-gather kernel / ring for every mode.
bool ring_kernel_ok(int mode, const RhsArgs& A) {
  static int sel = -1;                     // 0 default, 1 never (GRAND_EDGE=stream), 2 always (GRAND_EDGE=ring)
  if (sel < 0) {
    const char* e = getenv("GRAND_EDGE");
    sel = (e != nullptr && !strcmp(e, "stream")) ? 1 : (e != nullptr && !strcmp(e, "ring")) ? 2 : 0;
  }
  if (sel == 1) return false;
  // fixed weights (GRAND-l) have no scoring phase for the copies to overlap with: the register gathers are faster there
  // (ogbn-arxiv shape: 134 vs 144 us), the attention modes are faster on the ring (265 vs 288)
  if (sel == 0 && mode == MODE_WEIGHT) return false;
  // small graphs (everything L2-resident, a launch lasts as long as its longest unit): 2 x 10 register-gather warps per SM
  // finish sooner than 14 ring warps -- Photo 111 vs 139 us per evaluation, Cora 44.4 vs 45.5 (profiles/r02v_small.txt)
  if (sel == 0 && A.g.n_edges < (1 << 20)) return false;
  if (A.d > kWarp * 4 || (A.d & 3) != 0) return false;
  if ((reinterpret_cast<uintptr_t>(A.x) & 15) != 0) return false;
  return true;
}

int launch_ring_mode(int mode, const RhsArgs& A, cudaStream_t stream) {
  if (mode == MODE_FUSED) return launch_ring_mode_t<MODE_FUSED>(A, stream);
  if (mode == MODE_PRECOMP) return launch_ring_mode_t<MODE_PRECOMP>(A, stream);
  return launch_ring_mode_t<MODE_WEIGHT>(A, stream);
}

}  // namespace grand
