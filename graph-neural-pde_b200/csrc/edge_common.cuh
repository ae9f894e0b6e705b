// Views, PTX wrappers and warp helpers shared by the edge-phase translation units.
#pragma once
#include "attention.cuh"

namespace grand {

constexpr int MODE_FUSED = 0;     // scaled_dot scores + row softmax computed in the kernel
constexpr int MODE_PRECOMP = 1;   // scores / segment statistics from the general pre-pass
constexpr int MODE_WEIGHT = 2;    // GRAND-l: per-edge weights given

constexpr int STAGE_DIRECT = 0;   // neighbour rows straight into registers, a batch at a time
constexpr int STAGE_CPASYNC = 1;  // cp.async (LDGSTS) into the warp's shared-memory window buffer
constexpr int STAGE_BULK = 2;     // cp.async.bulk (TMA bulk copy) + mbarrier into the window buffer
constexpr int STAGE_ROWBUF = 3;   // register batches that ignore row boundaries; finished rows wait in shared memory

// window header written by the host plan (graph_plan.py): one int4 per window
//   .x first (or only) row   .y first CSR position   .z nrows | cnt << 8   .w split-row index or -1
// per-lane record: one int2 per lane
//   .x column j of the lane's edge (0 for idle lanes)
//   .y bits 0-7 row index inside the window, 8-13 segment lo, 14-19 segment hi (lanes of the lane's own row),
//      bits 20-25 / 26-31: lo / hi of row number `lane` (valid for lane < nrows)
struct GraphView {
  int n_rows, n_cols, n_edges, n_tiles, n_split_rows, n_split_tiles;
  const int *rowptr, *col, *erow, *eid, *split_tile0, *colptr, *csc_pos;
  const int4* win_hdr;
  const int2* win_rec;
  // row segments of the persistent edge kernel (rhs_stream.cu): rows with more than GRAND_TILE in-edges are handled
  // whole by one warp (<= seg_edges edges: "medium") or cut into segments of seg_edges edges ("long")
  int n_lunits, n_munits, n_lrows;
  const int4* seg_desc;      // [n_lunits + n_munits]: row, first CSR position, edge count, long-row index or -1
  const int* lrow_slot0;     // [n_lrows + 1] first segment (== partial-sum slot) of each long row
  const int* lrow_row;       // [n_lrows] row id of each long row
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
  const float* coef_dev;   // or null: device [1 + GRAND_MAX_TERMS] overriding coef_f / coef[] (device-side step control)
};
// stage-combine coefficients as the kernel sees them: host values, or the device table of the dopri5 controller
__device__ __forceinline__ float epi_coef_f(const EpiView& ep) { return ep.coef_dev != nullptr ? __ldg(ep.coef_dev) : ep.coef_f; }
__device__ __forceinline__ float epi_coef(const EpiView& ep, int t) {
  return ep.coef_dev != nullptr ? __ldg(ep.coef_dev + 1 + t) : ep.coef[t];
}

// delivery of finished y_out rows into the halo of the peers that reference them (node-partitioned multi-GPU)
struct PushView {
  const int* ptr;                    // [n_rows + 1] or null
  const int2* dst;                   // (peer, destination row)
  float* out[GRAND_MAX_PEERS];       // y_out as seen through every peer's arena mapping
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
  float* split_stats;        // [n_split_tiles, H, 2] per-chunk (max, sum-exp)
  float* row_stats;          // [n_split_rows, H, 2] merged per long row
  float* split_partial;      // [n_split_tiles, d]
  int* split_counter;        // [n_split_rows]
  float* lrow_stats;         // [n_lunits, H, 2] softmax (max, sum-exp) of every long-row segment (rhs_stream.cu pre-pass)
  float* lpartial;           // [n_lunits, max(H,1), d] partial sums of the long-row segments (per head in MODE_FUSED)
  int* lcounter;             // [n_lrows] arrival counters (atomicInc wraps them back to zero)
  PushView push;
  int l2_hints;              // bit 0: evict_last on the gathers, bit 1: evict_first on the streams
  int tile_begin, tile_end;  // window range of this launch (chunk tiles and packed windows may go to different streams)
};

// ---- shared-memory async copies and mbarrier -----------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const float* src, uint32_t bytes, uint32_t bar,
                                              uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
          "r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}

// ---- L2 residency: the gather sources (X, K, Q: ~108 MB on ogbn-arxiv) should stay in the 126 MB L2 across the
//      launch, the streamed stage buffers should not push them out --------------------------------------------
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_normal() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float4 ldg4_hint(const float* p, uint64_t policy) {
#ifdef GRAND_COHERENT_LOADS
  return __ldcg(reinterpret_cast<const float4*>(p));
#endif
  float4 r;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p), "l"(policy));
  return r;
}
__device__ __forceinline__ float4 ld_stream4_hint(const float* p, uint64_t policy) {
#ifdef GRAND_COHERENT_LOADS
  return __ldcg(reinterpret_cast<const float4*>(p));
#endif
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p), "l"(policy));
  return r;
}
__device__ __forceinline__ void st_stream4_hint(float* p, float4 v, uint64_t policy) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w), "l"(policy)
               : "memory");
}
// 256-bit read-only load (sm_100a: SASS LDG.E.256): a 64-byte K / Q row is two of these instead of four 128-bit loads --
// half the instructions and, for a lane-per-edge gather that touches 32 different lines, half the L1 wavefronts
struct float8 {
  float4 lo, hi;
};
__device__ __forceinline__ float8 ldg8(const float* p) {
  float8 r;
#ifdef GRAND_COHERENT_LOADS
  r.lo = __ldcg(reinterpret_cast<const float4*>(p));
  r.hi = __ldcg(reinterpret_cast<const float4*>(p) + 1);
  return r;
#endif
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z),
                 "=f"(r.hi.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void prefetch_l2_keep(const void* p) {
  asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(p));
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// ---- segmented all-reduce over contiguous lane segments [lo, hi): every lane gets its own segment's value ----
__device__ __forceinline__ float seg_max(float v, int lane, int lo, int hi) {
#pragma unroll
  for (int o = 1; o < kWarp; o <<= 1) {
    const float t = __shfl_down_sync(kFull, v, o);
    if (lane + o < hi) v = fmaxf(v, t);
  }
  return __shfl_sync(kFull, v, lo);
}
// the same maximum in ONE warp instruction: redux.sync over the segment's member mask on the order-preserving integer
// image of the float (every lane of a segment passes the same mask; lanes outside any segment pass their own bit)
__device__ __forceinline__ unsigned seg_mask(int lane, int lo, int hi, bool member) {
  const unsigned upto_hi = hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u);
  return member ? (upto_hi & ~((1u << lo) - 1u)) : (1u << lane);
}
__device__ __forceinline__ float seg_max_redux(float v, unsigned mask) {
  return ordered_to_float(__reduce_max_sync(mask, float_to_ordered(v)));
}
__device__ __forceinline__ float seg_sum(float v, int lane, int lo, int hi) {
#pragma unroll
  for (int o = 1; o < kWarp; o <<= 1) {
    const float t = __shfl_down_sync(kFull, v, o);
    if (lane + o < hi) v += t;
  }
  return __shfl_sync(kFull, v, lo);
}

}  // namespace grand
