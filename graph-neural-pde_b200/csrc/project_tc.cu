// Dense row projections on the 5th-generation tensor cores:  Y0 = X W0^T + b0,  Y1 = X W1^T + b1  (one pass over X).
//   * the Q / K projection of every RHS evaluation (reference src/function_transformer_attention.py:174-175):
//     W0 = Q.weight, W1 = K.weight;
//   * the encoder / decoder linears of the model shell (reference src/GNN.py:23-31, src/base_classes.py:110-127):
//     W1 absent.
//
// Data path (sm_100a only):  TMA (cp.async.bulk.tensor.2d, 128-byte swizzle) streams X in [128 rows x 32 columns]
// chunks into a ring of shared-memory stages; four converter warps split every chunk in place into its two TF32 halves
// (x = hi + lo); ONE thread issues tcgen05.mma.kind::tf32 (M = 128, N = N0 + N1, K = 8) three times per k-step --
// lo*W_hi + hi*W_lo + hi*W_hi, the 3xTF32 scheme of project_mma.cu, fp32-class accuracy -- accumulating in TENSOR MEMORY;
// tcgen05.commit hands the stage back to the TMA producer and, after the last chunk of a tile, the accumulator to the
// epilogue (tcgen05.ld 32x32b -> registers -> + bias -> global).  Accumulators are double buffered in TMEM, so the
// epilogue of tile t overlaps the loads, conversions and MMAs of tile t + 1.  The split weights are resident in shared
// memory for the lifetime of the persistent CTA, in the same K-major 128-byte-swizzled layout TMA produces.
//
//   warp 0      TMA producer (one lane)
//   warp 1      TMEM allocation (whole warp), MMA issue (one lane)
//   warps 2-5   converters, then epilogue of the previous tile (TMEM lane quadrant = warp index mod 4)
//
// Why tensor memory here: the mma.sync version (project_mma.cu) is bound by the legacy HMMA pipe (3 x 48 HMMA.1688 per
// 16-row tile plus 26 % cvt instructions: 34 us = 3.2 TB/s on ogbn-arxiv); here the conversions are 3 ALU instructions
// per element, the matrix work is 48 UMMA instructions per 128-row tile, and the kernel is a pure HBM stream.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace grand {

constexpr int kTcThreads = 192;
constexpr int kTcConvThreads = 128;
constexpr int kTcRows = 128;                    // rows per tile == UMMA M
constexpr int kTcChunkCols = 32;                // fp32 columns per chunk == one 128-byte swizzle row
constexpr int kTcTileBytes = kTcRows * 128;     // one [128 x 32] fp32 chunk
constexpr int kTcStageBytes = 2 * kTcTileBytes; // hi (in place of the raw chunk) + lo
constexpr int kTcMaxStages = 6;
constexpr int kTcMaxChunks = 4;                 // d <= 128
constexpr int kTcMaxN = 256;

struct TcArgs {
  int n_x_rows;        // rows of X to run through (tiles cover [0, n_x_rows))
  int d, nkc;          // contraction length, chunks of 32 columns
  int n0, n1;          // output widths (n1 may be 0); N = n0 + n1 is a multiple of 16
  int rows0, rows1;    // rows of Y0 / Y1 actually stored (<= n_x_rows)
  int ld0, ld1;        // leading dimensions of Y0 / Y1 (floats)
  int n_tiles, stages;
  int tmem_cols, acc_stride;
  int prefetch_tiles;  // tiles ahead of the ring that the producer requests into L2 (0 = off)
  const float* w0; const float* b0;
  const float* w1; const float* b1;
  float* y0; float* y1;
};

// ---- PTX wrappers ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tc_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tc_tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0),
               "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tc_tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
// K-major operand tile with 128-byte rows and the 128-byte swizzle: 8-row groups 1024 bytes apart (SBO), LBO unused (1),
// descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B; `addr` may point 32 k bytes into the first row (k-step)
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t addr) {
  return (uint64_t)((addr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand from TENSOR MEMORY (lane = row, one 32-bit column per k), B from shared memory
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 consecutive columns of this thread's TMEM lane <- 32 registers
__device__ __forceinline__ void tc_tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
      "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
      "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tc_tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tc_tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t tc_to_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return u;
}
__device__ __forceinline__ void tc_split4(const float4 v, float4& hi, float4& lo) {
  hi.x = __uint_as_float(tc_to_tf32(v.x)); lo.x = __uint_as_float(tc_to_tf32(v.x - hi.x));
  hi.y = __uint_as_float(tc_to_tf32(v.y)); lo.y = __uint_as_float(tc_to_tf32(v.y - hi.y));
  hi.z = __uint_as_float(tc_to_tf32(v.z)); lo.z = __uint_as_float(tc_to_tf32(v.z - hi.z));
  hi.w = __uint_as_float(tc_to_tf32(v.w)); lo.w = __uint_as_float(tc_to_tf32(v.w - hi.w));
}

// ---- the kernel -----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTcThreads, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ TcArgs A) {
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  // carve: [W_hi | W_lo | stages | barriers | tmem pointer | bias]; every tile base is a multiple of 1024 bytes
  // (aligned in the SHARED address space, which is what the swizzle and the descriptors see)
  unsigned char* base = tc_smem + ((1024u - (tc_smem_u32(tc_smem) & 1023u)) & 1023u);
  const int N = A.n0 + A.n1;
  const int w_chunk_bytes = N * 128;
  unsigned char* w_hi = base;
  unsigned char* w_lo = w_hi + (size_t)A.nkc * w_chunk_bytes;
  unsigned char* stage0 = w_lo + (size_t)A.nkc * w_chunk_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage0 + (size_t)A.stages * kTcStageBytes);
  uint64_t* full = bars;                          // TMA landed            (1 arrival + tx bytes)
  uint64_t* conv = bars + kTcMaxStages;           // chunk split into hi/lo (128 arrivals)
  uint64_t* empty = bars + 2 * kTcMaxStages;      // MMAs reading the stage retired (tcgen05.commit)
  uint64_t* acc_full = bars + 3 * kTcMaxStages;   // [2] accumulator complete (tcgen05.commit)
  uint64_t* acc_empty = acc_full + 2;             // [2] accumulator drained (128 arrivals)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* bias = reinterpret_cast<float*>(tmem_ptr + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time setup: barriers, TMEM, bias, split weights in the swizzled K-major layout --------------------------
  if (tid == 0) {
    for (int s = 0; s < A.stages; ++s) {
      tc_mbar_init(tc_smem_u32(full + s), 1);
      tc_mbar_init(tc_smem_u32(conv + s), kTcConvThreads);
      tc_mbar_init(tc_smem_u32(empty + s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc_mbar_init(tc_smem_u32(acc_full + b), 1);
      tc_mbar_init(tc_smem_u32(acc_empty + b), kTcConvThreads);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_ptr)),
                 "r"((uint32_t)A.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int o = tid; o < N; o += kTcThreads) {
    const float* b = o < A.n0 ? A.b0 : A.b1;
    bias[o] = b != nullptr ? __ldg(b + (o < A.n0 ? o : o - A.n0)) : 0.f;
  }
  {
    // element (n, k) of chunk c = k / 32 lives at  c * N * 128 + n * 128 + ((((k % 32) / 4) ^ (n & 7)) * 16) + (k % 4) * 4
    const int c4_per_row = A.nkc * 8;             // 16-byte pieces per weight row incl. the zero padding up to nkc * 32
    for (int idx = tid; idx < N * c4_per_row; idx += kTcThreads) {
      const int n = idx / c4_per_row, p = idx - n * c4_per_row;
      const int c = p >> 3, piece = p & 7, k = p * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < A.d && (n < A.n0 || A.w1 != nullptr))   // w1 == null with n1 > 0: zero rows padding N to a multiple of 16
        v = ldg4((n < A.n0 ? A.w0 + (size_t)n * A.d : A.w1 + (size_t)(n - A.n0) * A.d) + k);
      float4 hi, lo;
      tc_split4(v, hi, lo);
      const size_t off = (size_t)c * w_chunk_bytes + (size_t)n * 128 + (size_t)((piece ^ (n & 7)) * 16);
      *reinterpret_cast<float4*>(w_hi + off) = hi;
      *reinterpret_cast<float4*>(w_lo + off) = lo;
    }
  }
  tc_fence_async_smem();                           // generic-proxy writes (weights) -> visible to the tensor-core reads
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int first_tile = blockIdx.x, tile_step = gridDim.x;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      // The ring holds ~1.25 tiles, less than the bandwidth-delay product of an HBM stream at one CTA per SM: tiles further
      // ahead are requested into L2 (no shared memory needed), so that the ring's own loads see L2 latency
      for (int a = 1; a <= A.prefetch_tiles; ++a)
        for (int c = 0; c < A.nkc; ++c)
          if (first_tile + (a - 1) * tile_step < A.n_tiles)
            tc_tma_prefetch_2d(&xmap, c * kTcChunkCols, (first_tile + (a - 1) * tile_step) * kTcRows);
      for (int tile = first_tile; tile < A.n_tiles; tile += tile_step) {
        const int ahead = tile + A.prefetch_tiles * tile_step;
        if (A.prefetch_tiles > 0 && ahead < A.n_tiles)
          for (int c = 0; c < A.nkc; ++c) tc_tma_prefetch_2d(&xmap, c * kTcChunkCols, ahead * kTcRows);
        for (int c = 0; c < A.nkc; ++c) {
          tc_mbar_wait(tc_smem_u32(empty + s), ph ^ 1u);          // first pass over the ring: passes immediately
          tc_mbar_expect_tx(tc_smem_u32(full + s), kTcTileBytes);
          tc_tma_load_2d(tc_smem_u32(stage0 + (size_t)s * kTcStageBytes), &xmap, c * kTcChunkCols, tile * kTcRows,
                         tc_smem_u32(full + s));
          if (++s == A.stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      // instruction descriptor: D fp32 (bits 4-5 = 1), A and B TF32 (bits 7-9, 10-12 = 2), both K-major, N >> 3 at
      // bit 17, M >> 4 at bit 24
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kTcRows >> 4) << 24);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int tile = first_tile; tile < A.n_tiles; tile += tile_step, ++it) {
        const int b = it & 1;
        const uint32_t use = (uint32_t)(it >> 1) & 1u;
        tc_mbar_wait(tc_smem_u32(acc_empty + b), use ^ 1u);       // epilogue of tile it - 2 drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(b * A.acc_stride);
        for (int c = 0; c < A.nkc; ++c) {
          tc_mbar_wait(tc_smem_u32(conv + s), ph);
          tc_fence_after();
          const uint32_t a_hi = tc_smem_u32(stage0 + (size_t)s * kTcStageBytes);
          const uint32_t a_lo = a_hi + kTcTileBytes;
          const uint32_t b_hi = tc_smem_u32(w_hi + (size_t)c * w_chunk_bytes);
          const uint32_t b_lo = tc_smem_u32(w_lo + (size_t)c * w_chunk_bytes);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {                        // UMMA K = 8 tf32 = 32 bytes along the swizzled row
            const uint32_t ko = (uint32_t)k4 * 32u;
            tc_mma_tf32(tmem_d, tc_smem_desc(a_lo + ko), tc_smem_desc(b_hi + ko), idesc, (c | k4) != 0);   // small terms first
            tc_mma_tf32(tmem_d, tc_smem_desc(a_hi + ko), tc_smem_desc(b_lo + ko), idesc, 1u);
            tc_mma_tf32(tmem_d, tc_smem_desc(a_hi + ko), tc_smem_desc(b_hi + ko), idesc, 1u);
          }
          tc_commit(tc_smem_u32(empty + s));                      // stage free once these MMAs have read it
          if (c == A.nkc - 1) tc_commit(tc_smem_u32(acc_full + b));
          if (++s == A.stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else {
    // ===== converters + epilogue (128 threads) =====
    const int ct = tid - 64;                                     // 0 .. 127
    const int quad = warp & 3;                                   // TMEM lanes [32 quad, 32 quad + 32) belong to this warp
    int s = 0;
    uint32_t ph = 0;
    int it = 0;
    auto epilogue = [&](int tile, int eit) {
      const int b = eit & 1;
      const uint32_t use = (uint32_t)(eit >> 1) & 1u;
      tc_mbar_wait(tc_smem_u32(acc_full + b), use);
      tc_fence_after();
      const int row = tile * kTcRows + quad * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(b * A.acc_stride);
      for (int g = 0; g < N; g += 16) {
        uint32_t v[16];
        tc_tmem_ld16(taddr + (uint32_t)g, v);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int col = g + 4 * q4;
          float4 o;
          o.x = __uint_as_float(v[4 * q4 + 0]) + bias[col + 0];
          o.y = __uint_as_float(v[4 * q4 + 1]) + bias[col + 1];
          o.z = __uint_as_float(v[4 * q4 + 2]) + bias[col + 2];
          o.w = __uint_as_float(v[4 * q4 + 3]) + bias[col + 3];
          if (col < A.n0) {
            if (row < A.rows0) *reinterpret_cast<float4*>(A.y0 + (size_t)row * A.ld0 + col) = o;
          } else {
            if (row < A.rows1) *reinterpret_cast<float4*>(A.y1 + (size_t)row * A.ld1 + (col - A.n0)) = o;
          }
        }
      }
      tc_fence_before();
      tc_mbar_arrive(tc_smem_u32(acc_empty + b));
    };
    int prev_tile = -1;
    for (int tile = first_tile; tile < A.n_tiles; tile += tile_step, ++it) {
      for (int c = 0; c < A.nkc; ++c) {
        tc_mbar_wait(tc_smem_u32(full + s), ph);
        unsigned char* hi_t = stage0 + (size_t)s * kTcStageBytes;
        unsigned char* lo_t = hi_t + kTcTileBytes;
#pragma unroll
        for (int j = 0; j < kTcTileBytes / 16 / kTcConvThreads; ++j) {       // 8 x 16 bytes per thread
          const int off = (ct + j * kTcConvThreads) * 16;
          const float4 v = *reinterpret_cast<const float4*>(hi_t + off);
          float4 hi, lo;
          tc_split4(v, hi, lo);
          *reinterpret_cast<float4*>(hi_t + off) = hi;
          *reinterpret_cast<float4*>(lo_t + off) = lo;
        }
        tc_fence_async_smem();
        tc_mbar_arrive(tc_smem_u32(conv + s));
        if (++s == A.stages) { s = 0; ph ^= 1u; }
      }
      if (prev_tile >= 0) epilogue(prev_tile, it - 1);           // its MMAs ran while this tile was converted
      prev_tile = tile;
    }
    if (prev_tile >= 0) epilogue(prev_tile, it - 1);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)A.tmem_cols) : "memory");
  }
}

// ---- variant 2: the split X chunks live in TENSOR MEMORY, shared memory only holds what TMA lands ---------------------------
// The shared-memory version above moves every X byte through shared memory five times (TMA write, converter read, two
// converter writes, ~three tensor-core reads): ~0.5 MB per 128-row tile, which is what bounds it (36 us on ogbn-arxiv,
// no faster than mma.sync).  Here a converter thread owns ONE row of the tile (= one TMEM lane): it reads its 128 bytes
// of the landed chunk (8 LDS.128, conflict-free under the 128-byte swizzle), splits them and stores hi / lo straight
// into tensor memory (tcgen05.st 32x32b.x32), from where tcgen05.mma takes its A operand.  The landing buffer is free
// as soon as it has been read, so shared memory holds 10 x 16 KB of loads in flight instead of 5 stages; the tensor
// cores read only W from shared memory.
//   TMEM columns: [0, 2 acc_stride) two accumulators, then a_stages x (32 hi | 32 lo).
constexpr int kTmMaxRaw = 12;
constexpr int kTmConvThreads = 256;                 // 8 converter / epilogue warps
constexpr int kTmThreads = 64 + kTmConvThreads;
constexpr int kTmMaxA = 6;

struct TmArgs {
  TcArgs t;
  int raw_slots, a_stages, a_col0;
};

__global__ void __launch_bounds__(kTmThreads, 1)
linear_tm_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ TmArgs P) {
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  const TcArgs& A = P.t;
  unsigned char* base = tc_smem + ((1024u - (tc_smem_u32(tc_smem) & 1023u)) & 1023u);
  const int N = A.n0 + A.n1;
  const int w_chunk_bytes = N * 128;
  unsigned char* w_hi = base;
  unsigned char* w_lo = w_hi + (size_t)A.nkc * w_chunk_bytes;
  unsigned char* raw0 = w_lo + (size_t)A.nkc * w_chunk_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(raw0 + (size_t)P.raw_slots * kTcTileBytes);
  uint64_t* full = bars;                          // [raw_slots] TMA landed
  uint64_t* raw_free = full + kTmMaxRaw;          // [raw_slots] the 128 converter threads have read the slot
  uint64_t* a_ready = raw_free + kTmMaxRaw;       // [a_stages]  hi / lo of a chunk are in tensor memory (128 arrivals)
  uint64_t* a_free = a_ready + kTmMaxA;           // [a_stages]  the MMAs reading the stage retired (tcgen05.commit)
  uint64_t* acc_full = a_free + kTmMaxA;          // [2]
  uint64_t* acc_empty = acc_full + 2;             // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* bias = reinterpret_cast<float*>(tmem_ptr + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_launch_dependents();
  if (tid == 0) {
    for (int s = 0; s < P.raw_slots; ++s) {
      tc_mbar_init(tc_smem_u32(full + s), 1);
      tc_mbar_init(tc_smem_u32(raw_free + s), kTmConvThreads);
    }
    for (int s = 0; s < P.a_stages; ++s) {
      tc_mbar_init(tc_smem_u32(a_ready + s), kTmConvThreads);
      tc_mbar_init(tc_smem_u32(a_free + s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc_mbar_init(tc_smem_u32(acc_full + b), 1);
      tc_mbar_init(tc_smem_u32(acc_empty + b), kTmConvThreads);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_ptr)),
                 "r"((uint32_t)A.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int o = tid; o < N; o += kTmThreads) {
    const float* b = o < A.n0 ? A.b0 : A.b1;
    bias[o] = b != nullptr ? __ldg(b + (o < A.n0 ? o : o - A.n0)) : 0.f;
  }
  {
    const int c4_per_row = A.nkc * 8;
    for (int idx = tid; idx < N * c4_per_row; idx += kTmThreads) {
      const int n = idx / c4_per_row, p = idx - n * c4_per_row;
      const int c = p >> 3, piece = p & 7, k = p * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < A.d && (n < A.n0 || A.w1 != nullptr))
        v = ldg4((n < A.n0 ? A.w0 + (size_t)n * A.d : A.w1 + (size_t)(n - A.n0) * A.d) + k);
      float4 hi, lo;
      tc_split4(v, hi, lo);
      const size_t off = (size_t)c * w_chunk_bytes + (size_t)n * 128 + (size_t)((piece ^ (n & 7)) * 16);
      *reinterpret_cast<float4*>(w_hi + off) = hi;
      *reinterpret_cast<float4*>(w_lo + off) = lo;
    }
  }
  tc_fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int first_tile = blockIdx.x, tile_step = gridDim.x;
  // everything above (barriers, tensor memory, bias, split weights) is independent of the previous kernel in the stream;
  // X is read and Y written only below
  pdl_wait();

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = first_tile; tile < A.n_tiles; tile += tile_step) {
        for (int c = 0; c < A.nkc; ++c) {
          tc_mbar_wait(tc_smem_u32(raw_free + s), ph ^ 1u);
          tc_mbar_expect_tx(tc_smem_u32(full + s), kTcTileBytes);
          tc_tma_load_2d(tc_smem_u32(raw0 + (size_t)s * kTcTileBytes), &xmap, c * kTcChunkCols, tile * kTcRows,
                         tc_smem_u32(full + s));
          if (++s == P.raw_slots) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: A from tensor memory, B (split weights) from shared memory =====
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kTcRows >> 4) << 24);
      int as = 0;
      uint32_t aph = 0;
      int it = 0;
      for (int tile = first_tile; tile < A.n_tiles; tile += tile_step, ++it) {
        const int b = it & 1;
        const uint32_t use = (uint32_t)(it >> 1) & 1u;
        tc_mbar_wait(tc_smem_u32(acc_empty + b), use ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(b * A.acc_stride);
        for (int c = 0; c < A.nkc; ++c) {
          tc_mbar_wait(tc_smem_u32(a_ready + as), aph);
          tc_fence_after();
          const uint32_t a_hi = tmem_base + (uint32_t)(P.a_col0 + as * 64);
          const uint32_t a_lo = a_hi + 32u;
          const uint32_t b_hi = tc_smem_u32(w_hi + (size_t)c * w_chunk_bytes);
          const uint32_t b_lo = tc_smem_u32(w_lo + (size_t)c * w_chunk_bytes);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const uint32_t ko = (uint32_t)k4 * 32u, kc = (uint32_t)k4 * 8u;
            tc_mma_tf32_ts(tmem_d, a_lo + kc, tc_smem_desc(b_hi + ko), idesc, (c | k4) != 0);   // small terms first
            tc_mma_tf32_ts(tmem_d, a_hi + kc, tc_smem_desc(b_lo + ko), idesc, 1u);
            tc_mma_tf32_ts(tmem_d, a_hi + kc, tc_smem_desc(b_hi + ko), idesc, 1u);
          }
          tc_commit(tc_smem_u32(a_free + as));
          if (c == A.nkc - 1) tc_commit(tc_smem_u32(acc_full + b));
          if (++as == P.a_stages) { as = 0; aph ^= 1u; }
        }
      }
    }
  } else {
    // ===== converters + epilogue: 8 warps, two per TMEM lane quadrant; a thread owns HALF of a row's chunk (16 columns) =====
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;                            // warps 2-5: columns [0, 16) of a chunk, warps 6-9: [16, 32)
    const int m = quad * 32 + lane;                              // row of the tile == TMEM lane
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    int s = 0, as = 0;
    uint32_t ph = 0, aph = 0;
    int it = 0;
    auto epilogue = [&](int tile, int eit) {
      const int b = eit & 1;
      const uint32_t use = (uint32_t)(eit >> 1) & 1u;
      tc_mbar_wait(tc_smem_u32(acc_full + b), use);
      tc_fence_after();
      const int row = tile * kTcRows + m;
      const uint32_t taddr = tmem_base + lane_base + (uint32_t)(b * A.acc_stride);
      for (int g = 16 * half; g < N; g += 32) {                   // the two warps of a quadrant alternate 16-column groups
        uint32_t v[16];
        tc_tmem_ld16(taddr + (uint32_t)g, v);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int col = g + 4 * q4;
          float4 o;
          o.x = __uint_as_float(v[4 * q4 + 0]) + bias[col + 0];
          o.y = __uint_as_float(v[4 * q4 + 1]) + bias[col + 1];
          o.z = __uint_as_float(v[4 * q4 + 2]) + bias[col + 2];
          o.w = __uint_as_float(v[4 * q4 + 3]) + bias[col + 3];
          if (col < A.n0) {
            if (row < A.rows0) *reinterpret_cast<float4*>(A.y0 + (size_t)row * A.ld0 + col) = o;
          } else {
            if (row < A.rows1) *reinterpret_cast<float4*>(A.y1 + (size_t)row * A.ld1 + (col - A.n0)) = o;
          }
        }
      }
      tc_fence_before();
      tc_mbar_arrive(tc_smem_u32(acc_empty + b));
    };
    int prev_tile = -1;
    for (int tile = first_tile; tile < A.n_tiles; tile += tile_step, ++it) {
      for (int c = 0; c < A.nkc; ++c) {
        tc_mbar_wait(tc_smem_u32(full + s), ph);
        const unsigned char* rowp = raw0 + (size_t)s * kTcTileBytes + (size_t)m * 128;
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {                         // logical 16-byte piece j sits at physical piece j ^ (m & 7)
          const int j = 4 * half + jj;
          const float4 v = *reinterpret_cast<const float4*>(rowp + ((j ^ (m & 7)) * 16));
          float4 h, l;
          tc_split4(v, h, l);
          hi[4 * jj + 0] = __float_as_uint(h.x); hi[4 * jj + 1] = __float_as_uint(h.y);
          hi[4 * jj + 2] = __float_as_uint(h.z); hi[4 * jj + 3] = __float_as_uint(h.w);
          lo[4 * jj + 0] = __float_as_uint(l.x); lo[4 * jj + 1] = __float_as_uint(l.y);
          lo[4 * jj + 2] = __float_as_uint(l.z); lo[4 * jj + 3] = __float_as_uint(l.w);
        }
        tc_mbar_arrive(tc_smem_u32(raw_free + s));               // the slot's bytes are in registers: TMA may refill it
        tc_mbar_wait(tc_smem_u32(a_free + as), aph ^ 1u);        // the MMAs that read this A stage last time have retired
        tc_fence_after();
        const uint32_t acol = tmem_base + lane_base + (uint32_t)(P.a_col0 + as * 64);
        tc_tmem_st16(acol + (uint32_t)(16 * half), hi);
        tc_tmem_st16(acol + 32u + (uint32_t)(16 * half), lo);
        tc_tmem_wait_st();
        tc_fence_before();
        tc_mbar_arrive(tc_smem_u32(a_ready + as));
        if (++s == P.raw_slots) { s = 0; ph ^= 1u; }
        if (++as == P.a_stages) { as = 0; aph ^= 1u; }
      }
      if (prev_tile >= 0) epilogue(prev_tile, it - 1);
      prev_tile = tile;
    }
    if (prev_tile >= 0) epilogue(prev_tile, it - 1);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)A.tmem_cols) : "memory");
  }
}

// ---- host side --------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;     // resolved once per process through the runtime (no link against libcuda)
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// GRAND_EUNSUPPORTED (no error text) when the shape is outside what this kernel is built for: the caller falls back to
// the mma.sync kernel.
int launch_linear_tc(const float* x, int n_x_rows, int d, const float* w0, const float* b0, int n0, float* y0, int rows0,
                     int ld0, const float* w1, const float* b1, int n1, float* y1, int rows1, int ld1,
                     cudaStream_t stream, int variant) {
  const int N = n0 + n1;
  if (n_x_rows <= 0) return GRAND_OK;
  if ((d & 3) != 0 || d <= 0 || d > kTcMaxChunks * kTcChunkCols) return GRAND_EUNSUPPORTED;
  if ((N & 15) != 0 || N < 16 || N > kTcMaxN || (n0 & 3) != 0 || (n1 & 3) != 0) return GRAND_EUNSUPPORTED;
  if ((ld0 & 3) != 0 || (n1 > 0 && (ld1 & 3) != 0)) return GRAND_EUNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(x) & 15) != 0 || (reinterpret_cast<uintptr_t>(y0) & 15) != 0 ||
      (y1 != nullptr && (reinterpret_cast<uintptr_t>(y1) & 15) != 0) || (reinterpret_cast<uintptr_t>(w0) & 15) != 0 ||
      (w1 != nullptr && (reinterpret_cast<uintptr_t>(w1) & 15) != 0))
    return GRAND_EUNSUPPORTED;
  EncodeTiledFn encode = encode_tiled_fn();
  if (encode == nullptr) return GRAND_EUNSUPPORTED;

  TcArgs A;
  memset(&A, 0, sizeof(A));
  A.n_x_rows = n_x_rows; A.d = d; A.nkc = (d + kTcChunkCols - 1) / kTcChunkCols;
  A.n0 = n0; A.n1 = n1; A.rows0 = rows0; A.rows1 = rows1; A.ld0 = ld0; A.ld1 = ld1;
  A.n_tiles = (n_x_rows + kTcRows - 1) / kTcRows;
  A.w0 = w0; A.b0 = b0; A.w1 = w1; A.b1 = b1; A.y0 = y0; A.y1 = y1;
  int cols = 32;
  while (cols < 2 * N) cols <<= 1;
  A.tmem_cols = cols;
  A.acc_stride = cols / 2;
  const size_t w_bytes = (size_t)2 * A.nkc * N * 128;
  const size_t tail = 1024 /* alignment slack */ + 3 * kTcMaxStages * 8 + 4 * 8 + 16 + (size_t)N * 4;
  const size_t budget = 227 * 1024;
  if (w_bytes + tail + 2 * (size_t)kTcStageBytes > budget) return GRAND_EUNSUPPORTED;
  int stages = (int)((budget - w_bytes - tail) / kTcStageBytes);
  if (stages > kTcMaxStages) stages = kTcMaxStages;
  A.stages = stages;
  {
    static int pf = -1;                  // GRAND_TC_PREFETCH=<tiles>: measurement switch, read once
    if (pf < 0) {
      const char* e = getenv("GRAND_TC_PREFETCH");
      pf = e != nullptr ? atoi(e) : 0;   // measured on B200 (profiles/r02v_*): 0 -> 45.4 us, 2 -> 49.1, 4 -> 50.0: off
      if (pf < 0 || pf > 8) pf = 0;
    }
    A.prefetch_tiles = pf;
  }
  const size_t smem = w_bytes + (size_t)stages * kTcStageBytes + tail;

  CUtensorMap xmap;
  const cuuint64_t gdim[2] = {(cuuint64_t)d, (cuuint64_t)n_x_rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)d * 4};
  const cuuint32_t box[2] = {(cuuint32_t)kTcChunkCols, (cuuint32_t)kTcRows};
  const cuuint32_t estride[2] = {1, 1};
  const CUresult cr = encode(&xmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(x), gdim, gstride, box, estride,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GRAND_REQUIRE(cr == CUDA_SUCCESS, GRAND_ECUDA, "linear_tc: cuTensorMapEncodeTiled failed (%d) for [%d, %d]", (int)cr,
                n_x_rows, d);
  int grid = num_sms();
  if (grid > A.n_tiles) grid = A.n_tiles;
  // variant 3: split chunks in tensor memory (needs 2 accumulators + >= 2 A stages in the 512 columns)
  const int acc_stride_tm = (N + 31) & ~31;
  if (variant >= 3 && 2 * acc_stride_tm + 2 * 64 <= 512) {
    TmArgs P;
    P.t = A;
    P.t.tmem_cols = 512;
    P.t.acc_stride = acc_stride_tm;
    P.a_col0 = 2 * acc_stride_tm;
    P.a_stages = (512 - P.a_col0) / 64;
    if (P.a_stages > kTmMaxA) P.a_stages = kTmMaxA;
    const size_t tail_tm = 1024 + (2 * kTmMaxRaw + 2 * kTmMaxA + 4) * 8 + 16 + (size_t)N * 4;
    int raw = (int)((budget - w_bytes - tail_tm) / kTcTileBytes);
    if (raw > kTmMaxRaw) raw = kTmMaxRaw;
    if (raw >= 2) {
      P.raw_slots = raw;
      const size_t smem_tm = w_bytes + (size_t)raw * kTcTileBytes + tail_tm;
      GRAND_CUDA_OK(cudaFuncSetAttribute(linear_tm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tm));
      GRAND_CUDA_OK(launch_pdl(linear_tm_kernel, dim3(grid), dim3(kTmThreads), smem_tm, stream, xmap, P));
      count_launch();
      return GRAND_OK;
    }
  }
  // per device and cheap: set on every launch (no process-wide "already done" flag)
  GRAND_CUDA_OK(cudaFuncSetAttribute(linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  linear_tc_kernel<<<grid, kTcThreads, smem, stream>>>(xmap, A);
  GRAND_CUDA_OK(cudaGetLastError());
  count_launch();
  return GRAND_OK;
}

}  // namespace grand
