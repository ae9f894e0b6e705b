// Can per-warp rings of 512-byte cp.async.bulk row copies feed an SM faster than register gathers?  Every warp owns
// STAGES buffers of ROWS rows; it keeps STAGES-1 chunks of random rows of a buffer in flight (one bulk copy per lane,
// one mbarrier per stage) while it adds up the rows of the oldest chunk from shared memory (LDS.128 per lane, 4 FMAs) --
// the data path of the edge kernel without its arithmetic.  Prints GB/s delivered for a few (warps/SM, ROWS, STAGES).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/tma_probe.cu -o tools/tma_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}

// CONSUME: 0 wait only, 1 read every row from shared memory and accumulate
template <int ROWS, int STAGES, int CONSUME>
__global__ void ring_kernel(const float4* __restrict__ x, const int* __restrict__ idx, int64_t n_chunks,
                            float4* __restrict__ sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  unsigned char* mine = smem + (size_t)warp * (STAGES * ROWS * 512);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)nwarps * (STAGES * ROWS * 512)) + warp * STAGES;
  if (lane == 0)
    for (int s = 0; s < STAGES; ++s) mbar_init(smem_u32(bars + s), 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const int64_t gw = (int64_t)blockIdx.x * nwarps + warp, ngw = (int64_t)gridDim.x * nwarps;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  auto issue = [&](int64_t chunk, int s) {
    if (chunk >= n_chunks) return;
    const uint32_t bar = smem_u32(bars + s);
    if (lane == 0) mbar_expect_tx(bar, ROWS * 512);
    __syncwarp();
    if (lane < ROWS) {
      const int j = __ldg(idx + chunk * ROWS + lane);
      bulk_copy_g2s(smem_u32(mine + (size_t)(s * ROWS + lane) * 512), x + (size_t)j * 32, 512, bar);
    }
  };
  int64_t c = gw;
  for (int s = 0; s < STAGES - 1; ++s) issue(c + (int64_t)s * ngw, s);
  int it = 0;
  for (; c < n_chunks; c += ngw, ++it) {
    const int s = it % STAGES;
    issue(c + (int64_t)(STAGES - 1) * ngw, (it + STAGES - 1) % STAGES);
    mbar_wait(smem_u32(bars + s), (it / STAGES) & 1);
    if (CONSUME) {
      const float4* rows = reinterpret_cast<const float4*>(mine + (size_t)s * ROWS * 512);
#pragma unroll 8
      for (int r = 0; r < ROWS; ++r) {
        const float4 v = rows[r * 32 + lane];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    __syncwarp();      // every lane has read stage s before it is refilled
  }
  if (acc.x == 12345.678f) sink[0] = acc;
}

template <int ROWS, int STAGES, int CONSUME>
static void run(const float4* x, const int* idx, int64_t n_rows_read, float4* sink, int sms, int warps, const char* what) {
  const size_t smem = (size_t)warps * (STAGES * ROWS * 512 + STAGES * 8);
  if (smem > 227 * 1024) { printf("%-28s warps/SM %2d  smem %zu too large\n", what, warps, smem); return; }
  cudaFuncSetAttribute(ring_kernel<ROWS, STAGES, CONSUME>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int64_t n_chunks = n_rows_read / ROWS;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  for (int i = 0; i < 2; ++i) ring_kernel<ROWS, STAGES, CONSUME><<<sms, warps * 32, smem>>>(x, idx, n_chunks, sink);
  cudaEventRecord(a);
  for (int i = 0; i < 5; ++i) ring_kernel<ROWS, STAGES, CONSUME><<<sms, warps * 32, smem>>>(x, idx, n_chunks, sink);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, a, b);
  ms /= 5.f;
  cudaError_t e = cudaGetLastError();
  printf("%-28s warps/SM %2d  smem/SM %6.1f KB  %8.1f us  %8.1f GB/s  %s\n", what, warps, smem / 1024.0, ms * 1e3,
         (double)n_chunks * ROWS * 512 / (ms * 1e-3) / 1e9, e == cudaSuccess ? "" : cudaGetErrorString(e));
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
  const int mbs[] = {87, 1024};
  cudaMalloc(&d_x, (size_t)1 << 30);
  cudaMemset(d_x, 0, (size_t)1 << 30);
  for (int mi = 0; mi < 2; ++mi) {
    const int64_t rows = (int64_t)mbs[mi] * 1024 * 1024 / 512;
    uint64_t s = 88172645463325252ull;
    for (int64_t i = 0; i < n_read; ++i) {
      s ^= s << 13; s ^= s >> 7; s ^= s << 17;
      h_idx[i] = (int)(s % (uint64_t)rows);
    }
    cudaMemcpy(d_idx, h_idx, n_read * sizeof(int), cudaMemcpyHostToDevice);
    printf("== random 512-byte rows of a %d MB buffer, %lld rows per launch\n", mbs[mi], (long long)n_read);
    run<32, 2, 1>(d_x, d_idx, n_read, sink, sms, 7, "32 rows x 2 stages, consume");
    run<32, 2, 0>(d_x, d_idx, n_read, sink, sms, 7, "32 rows x 2 stages, wait");
    run<32, 2, 1>(d_x, d_idx, n_read, sink, sms, 4, "32 rows x 2 stages, consume");
    run<16, 3, 1>(d_x, d_idx, n_read, sink, sms, 9, "16 rows x 3 stages, consume");
    run<16, 2, 1>(d_x, d_idx, n_read, sink, sms, 14, "16 rows x 2 stages, consume");
    run<16, 4, 1>(d_x, d_idx, n_read, sink, sms, 7, "16 rows x 4 stages, consume");
    run<8, 4, 1>(d_x, d_idx, n_read, sink, sms, 14, "8 rows x 4 stages, consume");
    run<8, 2, 1>(d_x, d_idx, n_read, sink, sms, 28, "8 rows x 2 stages, consume");
    run<32, 1, 1>(d_x, d_idx, n_read, sink, sms, 14, "32 rows x 1 stage, consume");
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
