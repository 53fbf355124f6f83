// wgrad_h_tcgen05.cu -- filter gradients of the mask head on tcgen05 kind::f16 (IEEE-half operands, fp32 accumulation).
//
//   dW[t][k][n] += out_scale * sum_m A[m + shift_t, k] * D[m, n]
//
// A = the layer's input activation [rows][K] and D = the (loss-scaled) gradient of its output [rows][N], both stored
// as half with the CHANNEL index contiguous, i.e. both operands are MN-major (the reduction index is the row).
// Shared-memory tiles are TMA boxes of 64 channels (128 B) x RB rows in the plain 128B swizzle -- the canonical
// MN-major layout for 16-bit types: LBO = distance between 64-channel boxes, SBO = 1024 B between 8-row groups,
// one K=16 MMA spans two 8-row groups (2048 B).  (The tf32 flavour of this kernel, gemm_tcgen05.cu, needs the
// 32-byte-atom swizzle instead; half does not.)
// CTA = (row chunk, 128*NACC A-channels x BN D-channels, tap); NACC accumulators share every D stage; split-M
// partial sums are combined with fp32 atomics.  Warp roles as in gemm_tcgen05.cu.
// Replaces the Conv2DBackpropFilter nodes of the mask head (myolo/model.py:688-711) in the "h16" precision mode.
#include "tc_common.cuh"

namespace myolo {
namespace tc {

constexpr int kRBH = 64;  // reduction rows per stage
static int g_wgrad_sms = kNumSMs;   // CTAs per launch are sized for this many SMs (myolo_set_wgrad_sms)

template <int BN, int NACC>
__global__ void __launch_bounds__(kThreads)
tc_wgrad_h_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmD,
                  float* __restrict__ dW, long long M, int N, int K, TapShifts sh, long long chunk, int ntn,
                  int transpose_out, int stages, const float* __restrict__ out_scale) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * 8 + 1];
  __shared__ uint32_t tmem_slot;
  constexpr int RB = kRBH;
  constexpr uint32_t kBox = 64 * RB * 2;  // one 64-channel x RB-row half box
  constexpr int AM = BM * NACC;           // A-channels per CTA
  constexpr uint32_t kABytes = (AM / 64) * kBox, kDBytes = (BN / 64) * kBox, kStage = kABytes + kDBytes;
  constexpr uint32_t kCols = (BN * NACC) < 32 ? 32 : BN * NACC;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tap = blockIdx.z;
  const int k0 = (blockIdx.y / ntn) * AM;
  const int n0 = (blockIdx.y % ntn) * BN;
  const long long mbeg = (long long)blockIdx.x * chunk;
  const long long mend = min(M, mbeg + chunk);
  const int total = (int)((mend - mbeg + RB - 1) / RB);  // chunk is a multiple of RB; rows >= M are TMA zero fill
  const uint32_t bar0 = smem_u32(bars);
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (8 + s); };
  const uint32_t tfull = bar0 + 8u * 16;
  if (total <= 0) return;  // uniform per CTA

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmD);
    for (int s = 0; s < stages; ++s) {
      mbar_init(full(s), 1);
      mbar_init(empty(s), 1);
    }
    mbar_init(tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), kCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const int shift = sh.s[tap];
      for (int it = 0; it < total; ++it) {
        const int s = it % stages;
        const uint32_t ph = (uint32_t)(it / stages) & 1u;
        mbar_wait(empty(s), ph ^ 1u);
        mbar_expect_tx(full(s), kStage);
        const uint32_t sa = base + (uint32_t)s * kStage;
        const int row = (int)(mbeg + (long long)it * RB);
#pragma unroll
        for (int j = 0; j < AM / 64; ++j) tma_load_2d(sa + j * kBox, &tmA, full(s), k0 + 64 * j, row + shift);
#pragma unroll
        for (int j = 0; j < BN / 64; ++j) tma_load_2d(sa + kABytes + j * kBox, &tmD, full(s), n0 + 64 * j, row);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(BM, BN, 1, 1);
      for (int it = 0; it < total; ++it) {
        const int s = it % stages;
        const uint32_t ph = (uint32_t)(it / stages) & 1u;
        mbar_wait(full(s), ph);
        tc_fence_after();
        const uint32_t sa = base + (uint32_t)s * kStage;
        const uint64_t db = make_desc(sa + kABytes, kBox, 1024, 2);
#pragma unroll
        for (int acc = 0; acc < NACC; ++acc) {
          const uint64_t da = make_desc(sa + acc * (BM / 64) * kBox, kBox, 1024, 2);
#pragma unroll
          for (int k = 0; k < RB / 16; ++k)   // 16 reduction rows per MMA = 2048 B = 128 descriptor units
            umma_f16(tmem + (uint32_t)(acc * BN), da + (uint64_t)(k * 128), db + (uint64_t)(k * 128), idesc,
                     (it | k) != 0 ? 1u : 0u);
        }
        umma_commit(empty(s));
      }
      umma_commit(tfull);
    }
  } else {
    const int q = warp & 3;
    const float os = out_scale ? __ldg(out_scale) : 1.f;
    mbar_wait(tfull, 0);
    tc_fence_after();
    float* W = dW + (size_t)tap * K * N;
#pragma unroll 1
    for (int cc = 0; cc < NACC * BN; cc += 32) {
      const int acc = cc / BN, c0 = cc - acc * BN;
      const int k = k0 + acc * BM + q * 32 + lane;
      float v[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)cc, v);
      if (k < K) {
        if (transpose_out) {
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(W + (size_t)(n0 + c0 + j) * K + k, v[j] * os);
        } else {
          // this thread's 32 values are contiguous in memory: eight 128-bit vector reductions (sm_90+) instead of 32 scalar
          // ones -- a quarter of the L2 atomic operations, which bound the epilogue of the split-M filter gradients
          float4* wp = reinterpret_cast<float4*>(W + (size_t)k * N + n0 + c0);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            atomicAdd(wp + j, make_float4(v[4 * j] * os, v[4 * j + 1] * os, v[4 * j + 2] * os, v[4 * j + 3] * os));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, kCols);
}

template <int BN, int NACC>
static int launch_wgrad_h(const CUtensorMap& ta, const CUtensorMap& td, float* dW, long long M, int N, int K, int ntaps,
                          const TapShifts& sh, int transpose_out, const float* out_scale, cudaStream_t st) {
  const int ntk = (K + BM * NACC - 1) / (BM * NACC), ntn = N / BN;
  const long long tiles = (long long)ntk * ntn * ntaps;
  // g_wgrad_sms < 148 (myolo_set_wgrad_sms): the launch leaves SMs to the kernels of other streams
  long long nsplit = max(1LL, min(ceil_div(M, kRBH * 8), (long long)g_wgrad_sms / tiles));
  long long chunk = ceil_div(ceil_div(M, nsplit), kRBH) * kRBH;
  nsplit = ceil_div(M, chunk);
  const size_t per_stage = (size_t)(BM * NACC + BN) * kRBH * 2;
  const int stages = per_stage * 4 + 1024 <= 220 * 1024 ? 4 : 3;
  const size_t smem = (size_t)stages * per_stage + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    MYOLO_CUDA(cudaFuncSetAttribute(tc_wgrad_h_kernel<BN, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  dim3 grid((unsigned)nsplit, (unsigned)(ntk * ntn), (unsigned)ntaps);
  tc_wgrad_h_kernel<BN, NACC><<<grid, kThreads, smem, st>>>(ta, td, dW, M, N, K, sh, chunk, ntn, transpose_out, stages, out_scale);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

}  // namespace tc
}  // namespace myolo

using namespace myolo;
using namespace myolo::tc;

extern "C" int myolo_set_wgrad_sms(int n) {
  MYOLO_CHECK_ARG(n >= 32 && n <= kNumSMs);
  g_wgrad_sms = n;
  return MYOLO_OK;
}

extern "C" int myolo_gemm_taps_wgrad_h_supported(long long lda, long long ldd, long long M, int N, int K, int ntaps) {
  return M >= 64 && M < (1LL << 31) - 4096 && (K % 64) == 0 && (N % 64) == 0 && (lda % 8) == 0 && (ldd % 8) == 0 &&
         ntaps >= 1 && ntaps <= 32;
}

extern "C" int myolo_gemm_taps_wgrad_h(const void* A, long long lda, const void* D, long long ldd, float* dW, long long M,
                                       int N, int K, int ntaps, const int* shifts_host, int transpose_out,
                                       const float* out_scale, myolo_stream stream) {
  MYOLO_CHECK_ARG(A && D && dW && (((uintptr_t)A | (uintptr_t)D) & 15) == 0);
  MYOLO_CHECK_ARG(transpose_out || ((uintptr_t)dW & 15) == 0);     // 128-bit vector reductions into dW
  MYOLO_CHECK_ARG(myolo_gemm_taps_wgrad_h_supported(lda, ldd, M, N, K, ntaps));
  TapShifts sh;
  for (int t = 0; t < 32; ++t) sh.s[t] = (shifts_host && t < ntaps) ? shifts_host[t] : 0;
  CUtensorMap ta, td;
  int rc = get_map_h(A, M, K, lda, kRBH, 64, &ta);
  if (rc) return rc;
  rc = get_map_h(D, M, N, ldd, kRBH, 64, &td);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  const bool two = (K % (2 * BM)) == 0;
  if (N % 256 == 0) {
    if (two) return launch_wgrad_h<256, 2>(ta, td, dW, M, N, K, ntaps, sh, transpose_out, out_scale, st);
    return launch_wgrad_h<256, 1>(ta, td, dW, M, N, K, ntaps, sh, transpose_out, out_scale, st);
  }
  if (N % 128 == 0) {
    if (two) return launch_wgrad_h<128, 2>(ta, td, dW, M, N, K, ntaps, sh, transpose_out, out_scale, st);
    return launch_wgrad_h<128, 1>(ta, td, dW, M, N, K, ntaps, sh, transpose_out, out_scale, st);
  }
  return launch_wgrad_h<64, 1>(ta, td, dW, M, N, K, ntaps, sh, transpose_out, out_scale, st);
}
