// common.cuh -- shared helpers for libmyolo_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/myolo_b200.h"

namespace myolo {

void set_error(const char* fmt, ...);

#define MYOLO_CHECK_ARG(cond)                                                          \
  do {                                                                                 \
    if (!(cond)) {                                                                     \
      myolo::set_error("%s:%d: argument check failed: %s", __FILE__, __LINE__, #cond); \
      return MYOLO_ERR_ARG;                                                            \
    }                                                                                  \
  } while (0)

#define MYOLO_CHECK_LAUNCH()                                                              \
  do {                                                                                    \
    cudaError_t e__ = cudaGetLastError();                                                 \
    if (e__ != cudaSuccess) {                                                             \
      myolo::set_error("%s:%d: CUDA error: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return MYOLO_ERR_CUDA;                                                              \
    }                                                                                     \
  } while (0)

#define MYOLO_CUDA(call)                                                                  \
  do {                                                                                    \
    cudaError_t e__ = (call);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      myolo::set_error("%s:%d: CUDA error: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return MYOLO_ERR_CUDA;                                                              \
    }                                                                                     \
  } while (0)

static inline cudaStream_t as_stream(myolo_stream s) { return reinterpret_cast<cudaStream_t>(s); }
static inline long long ceil_div(long long a, long long b) { return (a + b - 1) / b; }

constexpr int kNumSMs = 148;  // B200

// ---- programmatic dependent launch (default; MYOLO_PDL=0 switches it off): the ~170 small kernels of the backbone form a serial chain on one
// stream; with the attribute below a kernel's CTAs are dispatched while its predecessor drains, and pdl_entry() -- the
// first statement of every kernel launched this way -- holds them until the predecessor has completed and its memory
// is visible.  Same results, same ordering; what is saved is the launch latency between dependent kernels.  Without the
// attribute (MYOLO_PDL=0) both instructions of pdl_entry() are no-ops.  Measured: +0.6 % on the step (profiles/r02_pdl_ab.txt).
int pdl_mode();
template <typename... KP, typename... A>
static inline void launch_k(void (*kernel)(KP...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_mode() ? 1 : 0;
  (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KP>(args)...);   // a failure is picked up by MYOLO_CHECK_LAUNCH
}
#define MYOLO_LAUNCH(kernel, grid, block, smem, st, ...) \
  myolo::launch_k(kernel, dim3(grid), dim3(block), (size_t)(smem), st, __VA_ARGS__)
__device__ __forceinline__ void pdl_entry() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// round-to-nearest fp32 -> tf32 (10-bit mantissa, low 13 bits zero).  tcgen05 kind::tf32 reads fp32
// words and ignores the low mantissa bits, so producers of tensor-core operands round here to keep
// the error unbiased.
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// `act` carries the activation in its low byte and MYOLO_ROUND_TF32 as a flag bit.
__device__ __forceinline__ float apply_act(float v, int act) {
  const int a = act & 0xff;
  if (a == MYOLO_ACT_RELU) v = fmaxf(v, 0.f);
  else if (a == MYOLO_ACT_RELU6) v = fminf(fmaxf(v, 0.f), 6.f);
  if (act & MYOLO_ROUND_TF32) v = round_tf32(v);
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// streaming 128-bit load that does not pollute L1 (data touched once)
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

}  // namespace myolo
