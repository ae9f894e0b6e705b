// What can the memory system of this GPU deliver to the SMs?  Microbenchmarks that bound the edge kernel's gather:
//   seq    every warp streams 512-byte rows of a buffer in order           (L2-resident when the buffer is < L2)
//   rand   every warp reads 512-byte rows at random row indices            (the SpMM gather pattern, no reuse)
// for buffer sizes from 16 MB (well inside the 126 MB L2) to 1 GB (DRAM), 8 rows in flight per warp, 48 warps / SM.
// Prints GB/s delivered into registers.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/l2_probe.cu -o l2_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

template <int NB>
__global__ void __launch_bounds__(256, 6) gather_kernel(const float4* __restrict__ x, const int* __restrict__ idx,
                                                          int64_t n_rows_to_read, float4* __restrict__ sink) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t r = warp * NB; r < n_rows_to_read; r += n_warps * NB) {
    int j[NB];
#pragma unroll
    for (int u = 0; u < NB; ++u) j[u] = __ldg(idx + r + u);
    float4 v[NB];
#pragma unroll
    for (int u = 0; u < NB; ++u) v[u] = ldg4(x + (size_t)j[u] * 32 + lane);     // 512-byte row = 32 float4
#pragma unroll
    for (int u = 0; u < NB; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
  }
  if (acc.x == 12345.678f) sink[0] = acc;
}

static float run(const float4* x, const int* idx, int64_t n_read, float4* sink, int sms) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) gather_kernel<8><<<sms * 6, 256>>>(x, idx, n_read, sink);
  cudaEventRecord(a);
  for (int i = 0; i < 5; ++i) gather_kernel<8><<<sms * 6, 256>>>(x, idx, n_read, sink);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, a, b);
  return ms / 5.f;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int64_t n_read = 4 << 20;                      // 4 Mi rows of 512 B = 2 GiB delivered per launch
  int* h_idx = (int*)malloc(n_read * sizeof(int));
  int* d_idx;
  float4 *d_x, *sink;
  cudaMalloc(&d_idx, n_read * sizeof(int));
  cudaMalloc(&sink, 64);
  const size_t max_bytes = (size_t)1 << 30;
  cudaMalloc(&d_x, max_bytes);
  cudaMemset(d_x, 0, max_bytes);
  printf("SMs %d; %lld rows of 512 B per launch\n", sms, (long long)n_read);
  const int mbs[] = {16, 32, 48, 64, 87, 128, 256, 1024};
  for (int mi = 0; mi < 8; ++mi) {
    const int64_t rows = (int64_t)mbs[mi] * 1024 * 1024 / 512;
    for (int mode = 0; mode < 2; ++mode) {
      uint64_t s = 88172645463325252ull;
      for (int64_t i = 0; i < n_read; ++i) {
        if (mode == 0) {
          h_idx[i] = (int)(i % rows);
        } else {
          s ^= s << 13; s ^= s >> 7; s ^= s << 17;
          h_idx[i] = (int)(s % (uint64_t)rows);
        }
      }
      cudaMemcpy(d_idx, h_idx, n_read * sizeof(int), cudaMemcpyHostToDevice);
      const float ms = run(d_x, d_idx, n_read, sink, sms);
      printf("buffer %5d MB  %-4s  %8.1f us  %8.1f GB/s\n", mbs[mi], mode == 0 ? "seq" : "rand", ms * 1e3,
             (double)n_read * 512 / (ms * 1e-3) / 1e9);
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
