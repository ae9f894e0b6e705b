// y = x W^T + b for the model shell's dense layers (reference src/GNN.py:23-31: encoder m1 / m11 / m12, decoder m2;
// src/base_classes.py:110-127) -- SURVEY.md section 8 (f).3.  Shapes with a short contraction (d_in <= 128: the ogbn-arxiv
// encoder, every decoder, the gradient with respect to the decoder's input) run on the tcgen05 kernel of project_tc.cu;
// everything else (Cora's 1433 input features) on the exact-fp32 tiled kernel below.
#include "common.cuh"

namespace grand {

int launch_linear_tc(const float* x, int n_x_rows, int d, const float* w0, const float* b0, int n0, float* y0, int rows0,
                     int ld0, const float* w1, const float* b1, int n1, float* y1, int rows1, int ld1,
                     cudaStream_t stream, int variant);
int projection_impl();

constexpr int kLinThreads = 256;   // 16 x 16 threads, (BM / 16) x (BN / 16) outputs each

// BM x BN output tile per block, fp32 FFMA, k-slabs of BK staged through shared memory (transposed, padded: conflict-free
// column reads; a warp's global loads cover whole 128-byte row segments).  Sums run over k in increasing order:
// deterministic.  Small problems take 32 x 32 tiles so that the grid still covers the machine.
template <int BM, int BN, int BK>
__global__ void __launch_bounds__(kLinThreads)
linear_ffma_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                   float* __restrict__ y, int rows, int d_in, int d_out, int ldy) {
  constexpr int TM = BM / 16, TN = BN / 16;
  __shared__ float xs[BK][BM + 4];
  __shared__ float ws[BK][BN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int row0 = blockIdx.x * BM, col0 = blockIdx.y * BN;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < d_in; k0 += BK) {
    for (int idx = tid; idx < BM * BK; idx += kLinThreads) {
      const int r = idx / BK, k = idx - r * BK;
      xs[k][r] = (row0 + r < rows && k0 + k < d_in) ? __ldg(x + (size_t)(row0 + r) * d_in + k0 + k) : 0.f;
    }
    for (int idx = tid; idx < BN * BK; idx += kLinThreads) {
      const int c = idx / BK, k = idx - c * BK;
      ws[k][c] = (col0 + c < d_out && k0 + k < d_in) ? __ldg(w + (size_t)(col0 + c) * d_in + k0 + k) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < BK; ++k) {
      float xv[TM], wv[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) xv[i] = xs[k][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) wv[j] = ws[k][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(xv[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int r = row0 + ty * TM + i;
    if (r >= rows) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int c = col0 + tx * TN + j;
      if (c < d_out) y[(size_t)r * ldy + c] = acc[i][j] + (b != nullptr ? __ldg(b + c) : 0.f);
    }
  }
}

int launch_linear(const float* x, int rows, int d_in, const float* w, const float* b, int d_out, float* y, int ldy,
                  cudaStream_t stream) {
  if (rows <= 0 || d_out <= 0) return GRAND_OK;
  if (projection_impl() >= 2 && (d_out & 3) == 0) {
    // outputs are padded to the next multiple of 16 with zero weight rows that are never stored
    const int pad = (16 - (d_out & 15)) & 15;
    const int rc = launch_linear_tc(x, rows, d_in, w, b, d_out, y, rows, ldy, nullptr, nullptr, pad, nullptr, 0, 4, stream,
                                    projection_impl());
    if (rc != GRAND_EUNSUPPORTED) return rc;
  }
  const long blocks64 = (long)((rows + 63) / 64) * ((d_out + 63) / 64);
  if (blocks64 >= 2 * num_sms()) {
    dim3 grid((rows + 63) / 64, (d_out + 63) / 64);
    linear_ffma_kernel<64, 64, 32><<<grid, kLinThreads, 0, stream>>>(x, w, b, y, rows, d_in, d_out, ldy);
  } else {
    dim3 grid((rows + 31) / 32, (d_out + 31) / 32);
    linear_ffma_kernel<32, 32, 32><<<grid, kLinThreads, 0, stream>>>(x, w, b, y, rows, d_in, d_out, ldy);
  }
  GRAND_CUDA_OK(cudaGetLastError());
  count_launch();
  return GRAND_OK;
}

}  // namespace grand
