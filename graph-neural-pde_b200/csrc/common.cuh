// Shared device/host helpers for the GRAND B200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/grand_b200.h"

namespace grand {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int cuda_fail(cudaError_t e, const char* what);

#define GRAND_CUDA_OK(call)                                    \
  do {                                                         \
    cudaError_t _e = (call);                                   \
    if (_e != cudaSuccess) return ::grand::cuda_fail(_e, #call); \
  } while (0)

#define GRAND_REQUIRE(cond, code, ...)    \
  do {                                    \
    if (!(cond)) {                        \
      ::grand::set_error(__VA_ARGS__);    \
      return (code);                      \
    }                                     \
  } while (0)

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------------------------
// The kernels of one RHS evaluation (projection -> long-row statistics -> edge kernel -> next projection ...) are launched
// with cudaLaunchAttributeProgrammaticStreamSerialization: a kernel calls pdl_launch_dependents() at its start, so its
// successor may be scheduled while this one still runs (on the SMs its persistent blocks have already left), and the
// successor runs its prologue -- barrier / TMEM set-up, weight staging, first work-list fetch -- before pdl_wait(), which
// returns once the predecessor grid has completed and its memory is visible.  NOTHING a predecessor writes (or still
// reads while this kernel would overwrite it) may be touched before pdl_wait().  Without the launch attribute both calls
// are no-ops.  Measured slower than plain launches (profiles/r02p_pdl.txt): off unless GRAND_PDL=1.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();

template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// order-preserving float <-> int encoding so that atomicMax on ints implements a float max
__device__ __forceinline__ int float_to_ordered(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

#ifdef GRAND_COHERENT_LOADS
// Translation units whose kernels REWRITE their inputs between grid-wide barriers (rhs_coop.cu: a whole integration in
// one launch) must not use the non-coherent path: every state / Q / K load goes to L2 (ld.global.cg), which is where
// the other blocks' stores of the previous stage are.
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld_stream4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
#else
// 128-bit read-only gather load.  X/K rows are re-read by many warps: keep them in L1 (default policy).
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// streaming accesses (touched once per kernel): do not pollute L1
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
#endif
// d{0,1} += a * b{0,1} as ONE packed instruction (sm_100 FFMA2: fma.rn.f32x2, each half an IEEE fma -- bit-identical to
// two fmaf).  GRAND_FFMA2=0 builds the scalar form (A/B).
#ifndef GRAND_FFMA2
#define GRAND_FFMA2 1
#endif
__device__ __forceinline__ void fma2(float& d0, float& d1, float a, float b0, float b1) {
#if GRAND_FFMA2
  unsigned long long pa, pb, pc;
  asm("mov.b64 %0, {%1, %2};" : "=l"(pa) : "f"(a), "f"(a));
  asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b0), "f"(b1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(pc) : "f"(d0), "f"(d1));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(pc) : "l"(pa), "l"(pb));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(pc));
#else
  d0 = fmaf(a, b0, d0);
  d1 = fmaf(a, b1, d1);
#endif
}
__device__ __forceinline__ void st_stream4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w));
}

}  // namespace grand
