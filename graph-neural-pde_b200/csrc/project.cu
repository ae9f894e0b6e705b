// Q/K projection:  Q = X Wq^T + bq (rows < n_rows),  K = X Wk^T + bk (rows < n_cols).
// Replaces the two nn.Linear calls of reference src/function_transformer_attention.py:174-175.
//
// HBM-bound (reads X once, writes 2*ad floats per row); the contraction depth is d <= a few hundred, so this is a
// register-tiled fp32 FFMA kernel: the whole [2 ad, d] weight lives in shared memory for the lifetime of a
// persistent CTA, X row tiles stream through shared memory, every thread owns an RT x OT output block.
// fp32 FFMA on purpose: the parity bar is 1e-5 of the reference's fp32 GEMM, which a single TF32 pass cannot meet.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace grand {

#ifndef GRAND_PROJ_DEFAULT
#define GRAND_PROJ_DEFAULT 3   // tcgen05 with the split X chunks in tensor memory (validated on B200: profiles/r02w-y_*); shapes it
                               // does not cover fall through to 2 (operands in shared memory), 1 (mma.sync), 0 (FFMA)
#endif

constexpr int kProjThreads = 256;

template <int RT, int OT>
__global__ void __launch_bounds__(kProjThreads, 2)
project_qk_kernel(const float* __restrict__ x, int n_rows, int n_cols, int d, int d4, int ld,
                  const float* __restrict__ wq, const float* __restrict__ bq, const float* __restrict__ wk,
                  const float* __restrict__ bk, int ad, int tx_count, int rows_per_tile, int n_tiles,
                  float* __restrict__ q, float* __restrict__ k) {
  extern __shared__ __align__(16) float smem[];
  const int out = 2 * ad;
  float* ws = smem;                       // [out][ld]
  float* xs = smem + (size_t)out * ld;    // [rows_per_tile][ld]
  const int tid = threadIdx.x;
  const int nthreads = blockDim.x;   // 128 or 256
  const int c4n = d4 >> 2;

  // stage the weights once per CTA (zero padded to d4); loads are batched ahead of the shared stores so that a CTA
  // pays the L2 latency a few times, not once per element (this is what bounds the small graphs)
  if ((d & 3) == 0) {
    const int total = out * c4n;
    constexpr int kBatch = 8;
    for (int base = tid; base < total; base += nthreads * kBatch) {
      float4 v[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int idx = base + u * nthreads;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < total) {
          const int o = idx / c4n, c4 = idx - o * c4n;
          const float* src = (o < ad) ? wq + (size_t)o * d : wk + (size_t)(o - ad) * d;
          v[u] = ldg4(src + 4 * c4);
        }
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int idx = base + u * nthreads;
        if (idx < total) {
          const int o = idx / c4n, c4 = idx - o * c4n;
          *reinterpret_cast<float4*>(ws + o * ld + 4 * c4) = v[u];
        }
      }
    }
  } else {
    for (int idx = tid; idx < out * d4; idx += nthreads) {
      const int o = idx / d4, c = idx - o * d4;
      float v = 0.f;
      if (c < d) v = (o < ad) ? wq[(size_t)o * d + c] : wk[(size_t)(o - ad) * d + c];
      ws[o * ld + c] = v;
    }
  }

  const int tx = tid % tx_count;
  const int ty = tid / tx_count;
  const int ty_count = rows_per_tile / RT;
  const bool active = ty < ty_count;
  const bool vec_ok = (d & 3) == 0;

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int row0 = tile * rows_per_tile;
    __syncthreads();  // previous tile's compute is done (and the weights are visible on the first pass)
    if (vec_ok) {
      for (int idx = tid; idx < rows_per_tile * c4n; idx += nthreads) {
        const int r = idx / c4n, c4 = idx - r * c4n;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < n_cols) v = ld_stream4(x + (size_t)(row0 + r) * d + 4 * c4);
        *reinterpret_cast<float4*>(xs + r * ld + 4 * c4) = v;
      }
    } else {
      for (int idx = tid; idx < rows_per_tile * d4; idx += nthreads) {
        const int r = idx / d4, c = idx - r * d4;
        float v = 0.f;
        if (row0 + r < n_cols && c < d) v = x[(size_t)(row0 + r) * d + c];
        xs[r * ld + c] = v;
      }
    }
    __syncthreads();
    if (!active) continue;

    float acc[RT][OT];
#pragma unroll
    for (int ri = 0; ri < RT; ++ri)
#pragma unroll
      for (int oi = 0; oi < OT; ++oi) acc[ri][oi] = 0.f;

    const float* xrow = xs + (ty * RT) * ld;
    for (int c = 0; c < d4; c += 4) {
      float4 xv[RT], wv[OT];
#pragma unroll
      for (int ri = 0; ri < RT; ++ri) xv[ri] = *reinterpret_cast<const float4*>(xrow + ri * ld + c);
#pragma unroll
      for (int oi = 0; oi < OT; ++oi) {
        const int o = tx + oi * tx_count;
        wv[oi] = (o < out) ? *reinterpret_cast<const float4*>(ws + o * ld + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int ri = 0; ri < RT; ++ri)
#pragma unroll
        for (int oi = 0; oi < OT; ++oi) {
          acc[ri][oi] = fmaf(xv[ri].x, wv[oi].x, acc[ri][oi]);
          acc[ri][oi] = fmaf(xv[ri].y, wv[oi].y, acc[ri][oi]);
          acc[ri][oi] = fmaf(xv[ri].z, wv[oi].z, acc[ri][oi]);
          acc[ri][oi] = fmaf(xv[ri].w, wv[oi].w, acc[ri][oi]);
        }
    }
#pragma unroll
    for (int oi = 0; oi < OT; ++oi) {
      const int o = tx + oi * tx_count;
      if (o >= out) continue;
      const bool is_q = o < ad;
      const float bias = is_q ? bq[o] : bk[o - ad];
#pragma unroll
      for (int ri = 0; ri < RT; ++ri) {
        const int r = row0 + ty * RT + ri;
        const float v = acc[ri][oi] + bias;
        if (is_q) {
          if (r < n_rows) q[(size_t)r * ad + o] = v;
        } else {
          if (r < n_cols) k[(size_t)r * ad + (o - ad)] = v;
        }
      }
    }
  }
}

int launch_project_qk_mma(const float* x, int n_rows, int n_cols, int d, const float* wq, const float* bq,
                          const float* wk, const float* bk, int ad, float* q, float* k, cudaStream_t stream);

int launch_linear_tc(const float* x, int n_x_rows, int d, const float* w0, const float* b0, int n0, float* y0, int rows0,
                     int ld0, const float* w1, const float* b1, int n1, float* y1, int rows1, int ld1,
                     cudaStream_t stream, int variant);

// Which kernel computes the projections: 3 / 2 = tcgen05 / TMEM / TMA (project_tc.cu; 3 keeps the split X chunks in
// tensor memory), 1 = mma.sync 3xTF32 (project_mma.cu), 0 = fp32 FFMA.  A kernel that does not cover a shape hands over to the next one.  GRAND_PROJ=tm|tc|mma|ffma is read once;
// grand_set_projection_impl overrides it at run time (in-process A/B measurements).
static int g_proj_impl = -1;
int projection_impl() {
  if (g_proj_impl < 0) {
    const char* e = getenv("GRAND_PROJ");
    g_proj_impl = (e != nullptr && !strcmp(e, "ffma")) ? 0 : (e != nullptr && !strcmp(e, "mma")) ? 1
                : (e != nullptr && !strcmp(e, "tc")) ? 2 : (e != nullptr && !strcmp(e, "tm")) ? 3 : GRAND_PROJ_DEFAULT;
  }
  return g_proj_impl;
}
void set_projection_impl(int impl) { g_proj_impl = impl; }

int launch_project_qk(const float* x, int n_rows, int n_cols, int d, const float* wq, const float* bq,
                      const float* wk, const float* bk, int ad, float* q, float* k, cudaStream_t stream) {
  if (n_cols <= 0) return GRAND_OK;
  const int impl = projection_impl();
  if (impl >= 2) {
    const int rc = launch_linear_tc(x, n_cols, d, wq, bq, ad, q, n_rows, ad, wk, bk, ad, k, n_cols, ad, stream, impl);
    if (rc != GRAND_EUNSUPPORTED) return rc;
  }
  if (impl >= 1) {
    const int rc = launch_project_qk_mma(x, n_rows, n_cols, d, wq, bq, wk, bk, ad, q, k, stream);
    if (rc != GRAND_EUNSUPPORTED) return rc;
  }
  constexpr int RT = 4, OT = 4;
  const int out = 2 * ad;
  const int tx_count = (out + OT - 1) / OT;
  GRAND_REQUIRE(tx_count <= kProjThreads, GRAND_EUNSUPPORTED, "project_qk: attention_dim %d too wide (max %d)", ad,
                kProjThreads * OT / 2);
  const int d4 = (d + 3) & ~3;
  const int ld = ((d4 >> 2) & 1) ? d4 : d4 + 4;  // ld/4 odd -> conflict-free 128-bit shared loads across rows
  // narrow outputs (few thread columns): 128-thread CTAs with small row tiles, so that several CTAs per SM overlap one
  // CTA's tile load with another's arithmetic; wide outputs keep 256 threads
  const int threads = tx_count <= 16 ? 128 : kProjThreads;
  const int ty_count = threads / tx_count;
  const int rows_per_tile = ty_count * RT;
  const int n_tiles = (n_cols + rows_per_tile - 1) / rows_per_tile;
  const size_t smem = (size_t)(out + rows_per_tile) * ld * sizeof(float);
  GRAND_REQUIRE(smem <= 227 * 1024, GRAND_EUNSUPPORTED,
                "project_qk: [2*%d, %d] weights + row tile need %zu B of shared memory (> 227 KiB)", ad, d, smem);
  auto kern = project_qk_kernel<RT, OT>;
  // per device and cheap: set on every launch (no process-wide "already done" flag)
  GRAND_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 6) per_sm = 6;
  int grid = num_sms() * per_sm;
  if (grid > n_tiles) grid = n_tiles;
  kern<<<grid, threads, smem, stream>>>(x, n_rows, n_cols, d, d4, ld, wq, bq, wk, bk, ad, tx_count,
                                             rows_per_tile, n_tiles, q, k);
  GRAND_CUDA_OK(cudaGetLastError());
  count_launch();
  return GRAND_OK;
}

}  // namespace grand
