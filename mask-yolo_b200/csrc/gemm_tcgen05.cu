// gemm_tcgen05.cu -- tensor-core tap-GEMM family for sm_100a: TMA -> 128B-swizzled shared memory ->
// tcgen05.mma (kind::tf32, fp32 accumulate in TMEM) -> tcgen05.ld epilogue.
//
//   forward / dgrad : C[m,n] = epi( sum_t sum_k A[m + shift_t, k] * Bt[t][n][k] )        (tc_gemm_kernel)
//   wgrad           : dW[t][k][n] += sum_m A[m + shift_t, k] * D[m, n]                   (tc_wgrad_kernel)
//
// These replace the Conv2D / Conv2DBackpropInput / Conv2DBackpropFilter call sites of the reference
// graph (myolo/model.py:271, 688-713, 848 and the keras_applications pointwise convs, SURVEY K3/K6/
// K8/K10).  A 3x3 SAME convolution on a padded-flat tensor is nine row-shifted
// GEMMs accumulated in one TMEM tile: the shift is just the TMA row coordinate, out-of-range rows
// are zero-filled by TMA, so there is no im2col buffer and no halo logic.
//
// Warp roles (192 threads): warp 0 = TMA producer (one lane), warp 1 = TMEM allocator + MMA issuer
// (one lane), warps 2..5 = epilogue (TMEM lane quarter = warp % 4).  smem ring of `stages` slots,
// full/empty mbarriers, tcgen05.commit releases a slot when the MMAs that read it retire.
//
// Operand layouts in shared memory
//   forward/dgrad: A and Bt are K-major: a TMA box of 32 fp32 (=128 B, one swizzle atom) x rows.
//                  UMMA descriptor: SWIZZLE_128B, SBO = 1024 B (8 rows), k-step of 8 tf32 = +32 B.
//   wgrad:         both operands are MN-major (the reduction index is the row index): boxes of
//                  32 channels x 32 rows; LBO = 4096 B between 32-channel chunks, SBO = 1024 B
//                  between 8-row groups, k-step of 8 rows = +1024 B.
#include <mutex>
#include <unordered_map>
#include "tc_common.cuh"

namespace myolo {
namespace tc {

// ------------------------------------------------------------------------------------------
// forward / dgrad kernel: one 128 x BN output tile per CTA
// ------------------------------------------------------------------------------------------
template <int BN>
__global__ void __launch_bounds__(kThreads)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ C,
               long long ldc, long long M, int N, int K, int ntaps, TapShifts sh, Epi ep, int stages) {
  pdl_entry();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * 8 + 1];
  __shared__ uint32_t tmem_slot;
  __shared__ float colacc[4][2][BN];   // epilogue statistics: per-column sum / sum of squares, one slot per epilogue warp
                                       // (summed in a fixed order afterwards: the statistics are reproducible run to run)
  __shared__ int s_last;
  constexpr uint32_t kABytes = BM * BK * 4, kBBytes = BN * BK * 4, kStage = kABytes + kBBytes;
  constexpr uint32_t kCols = BN < 32 ? 32 : BN;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int kblocks = K / BK;
  const int total = ntaps * kblocks;
  const uint32_t bar0 = smem_u32(bars);
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (8 + s); };
  const uint32_t tfull = bar0 + 8u * 16;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
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
      int it = 0;
      for (int t = 0; t < ntaps; ++t) {
        const int arow = (int)(m0 + sh.s[t]);
        const int brow = t * N + n0;
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % stages;
          const uint32_t ph = (uint32_t)(it / stages) & 1u;
          mbar_wait(empty(s), ph ^ 1u);
          mbar_expect_tx(full(s), kStage);
          const uint32_t sa = base + (uint32_t)s * kStage;
          tma_load_2d(sa, &tmA, full(s), kb * BK, arow);
          tma_load_2d(sa + kABytes, &tmB, full(s), kb * BK, brow);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN, 0, 0);
      for (int it = 0; it < total; ++it) {
        const int s = it % stages;
        const uint32_t ph = (uint32_t)(it / stages) & 1u;
        mbar_wait(full(s), ph);
        tc_fence_after();
        const uint32_t sa = base + (uint32_t)s * kStage;
        const uint64_t da = make_desc(sa, 16, 1024);
        const uint64_t db = make_desc(sa + kABytes, 16, 1024);
#pragma unroll
        for (int k = 0; k < BK / 8; ++k)
          umma_tf32(tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (it | k) != 0 ? 1u : 0u);
        umma_commit(empty(s));
      }
      umma_commit(tfull);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const long long m = m0 + row;
    const bool valid = (m < M) && pf_valid(m, ep.pf_w1, ep.pf_blk);
    mbar_wait(tfull, 0);
    tc_fence_after();
    float* crow = C + m * ldc + n0;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      float v[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      if (valid) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int n = n0 + c0 + j + e;
            float t = v[j + e];
            if (ep.bias) t += __ldg(ep.bias + n);
            if (ep.scale) t = fmaf(t, __ldg(ep.scale + n), __ldg(ep.shift + n));
            o[e] = apply_act(t, ep.act);
            v[j + e] = o[e];
          }
          float4* dst = reinterpret_cast<float4*>(crow + c0 + j);
          float4 r = make_float4(o[0], o[1], o[2], o[3]);
          if (ep.accumulate) {
            const float4 old = *dst;
            r.x += old.x; r.y += old.y; r.z += old.z; r.w += old.w;
          }
          *dst = r;
        }
      }
      if (ep.st_sums) {     // warp-uniform: column sums of the stored values over this warp's 32 rows (invalid rows count 0)
        float sq[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] = valid ? v[j] : 0.f;
          sq[j] = v[j] * v[j];
        }
        const float s0 = warp_colsum32(v, lane), s1 = warp_colsum32(sq, lane);
        colacc[q][0][c0 + lane] = s0;
        colacc[q][1][c0 + lane] = s1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, kCols);
  if (ep.st_sums == nullptr) return;
  for (int i = threadIdx.x; i < 2 * BN; i += blockDim.x) {
    const int which = i / BN, n = i - which * BN;
    atomicAdd(ep.st_sums + (size_t)which * N + n0 + n,
              ((double)colacc[0][which][n] + (double)colacc[1][which][n]) + ((double)colacc[2][which][n] + (double)colacc[3][which][n]));
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ep.st_ticket + blockIdx.y, 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int n = threadIdx.x; n < BN; n += blockDim.x) {
    const int c = n0 + n;
    const double S0 = *(volatile double*)(ep.st_sums + c), S1 = *(volatile double*)(ep.st_sums + N + c);
    const double mean = S0 * ep.st_inv;
    const double var = S1 * ep.st_inv - mean * mean;
    ep.st_mean[c] = (float)mean;
    ep.st_var[c] = (float)(var > 0.0 ? var : 0.0);
    ep.st_sums[c] = 0.0;
    ep.st_sums[N + c] = 0.0;
  }
  if (threadIdx.x == 0) ep.st_ticket[blockIdx.y] = 0;
}

// ------------------------------------------------------------------------------------------
// wgrad kernel: CTA = (row chunk, 128 A-channels x BN D-channels, tap); reduction over rows.
// ------------------------------------------------------------------------------------------
// NACC accumulators of 128 A-channels each share every D stage (halves the D traffic per FLOP).
template <int BN, int NACC>
__global__ void __launch_bounds__(kThreads)
tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmD,
                float* __restrict__ dW, long long M, int N, int K, TapShifts sh, long long chunk, int ntn,
                int transpose_out, int stages) {
  pdl_entry();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * 8 + 1];
  __shared__ uint32_t tmem_slot;
  constexpr int RB = 32;  // rows per stage
  constexpr uint32_t kBox = 32 * RB * 4;  // one 32-channel x 32-row box = 4 KB
  constexpr int AM = BM * NACC;  // A-channels per CTA
  constexpr uint32_t kABytes = (AM / 32) * kBox, kDBytes = (BN / 32) * kBox, kStage = kABytes + kDBytes;
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
        for (int j = 0; j < AM / 32; ++j) tma_load_2d(sa + j * kBox, &tmA, full(s), k0 + 32 * j, row + shift);
#pragma unroll
        for (int j = 0; j < BN / 32; ++j) tma_load_2d(sa + kABytes + j * kBox, &tmD, full(s), n0 + 32 * j, row);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN, 1, 1);
      for (int it = 0; it < total; ++it) {
        const int s = it % stages;
        const uint32_t ph = (uint32_t)(it / stages) & 1u;
        mbar_wait(full(s), ph);
        tc_fence_after();
        const uint32_t sa = base + (uint32_t)s * kStage;
        // MN-major tf32: 128B_BASE32B atoms of 4 reduction rows x 128 B; LBO = next 32-channel box,
        // SBO = next 4-row group (512 B); one K=8 MMA spans two atoms = 1024 B.
        const uint64_t db = make_desc(sa + kABytes, kBox, 512, 1);
#pragma unroll
        for (int acc = 0; acc < NACC; ++acc) {
          const uint64_t da = make_desc(sa + acc * (BM / 32) * kBox, kBox, 512, 1);
#pragma unroll
          for (int k = 0; k < RB / 8; ++k)
            umma_tf32(tmem + (uint32_t)(acc * BN), da + (uint64_t)(k * 64), db + (uint64_t)(k * 64), idesc,
                      (it | k) != 0 ? 1u : 0u);
        }
        umma_commit(empty(s));
      }
      umma_commit(tfull);
    }
  } else {
    const int q = warp & 3;
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
          for (int j = 0; j < 32; ++j) atomicAdd(W + (size_t)(n0 + c0 + j) * K + k, v[j]);
        } else {
          // this thread's 32 values are contiguous in memory: eight 128-bit vector reductions (sm_90+) instead of 32 scalar
          // ones -- a quarter of the L2 atomic operations, which bound the epilogue of the split-M filter gradients
          float4* wp = reinterpret_cast<float4*>(W + (size_t)k * N + n0 + c0);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            atomicAdd(wp + j, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, kCols);
}

// out[t][c][r] = f(in[t][r][c]) (transpose) or out = f(in); f = optional tf32 rounding
// round == 2: split staging for the 3xTF32 forward (hi = rna(v), lo = rna(v - hi)); three stacked
// copies [hi | hi | lo], each `total` floats, matching the tap triple (A_hi,B_hi) (A_lo,B_hi) (A_hi,B_lo).
__global__ void prep_weights_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols,
                                    int transpose, int round, size_t total) {
  pdl_entry();
  __shared__ float t[32][33];
  const float* ip = in + (size_t)blockIdx.z * rows * cols;
  float* op = out + (size_t)blockIdx.z * rows * cols;
  const int c = blockIdx.x * 32 + threadIdx.x;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int r = blockIdx.y * 32 + j;
    float v = (r < rows && c < cols) ? ip[(size_t)r * cols + c] : 0.f;
    if (round == 1) v = round_tf32(v);
    if (!transpose) {
      if (r < rows && c < cols) {
        if (round == 2) {
          const float hi = round_tf32(v);
          op[(size_t)r * cols + c] = hi;
          op[total + (size_t)r * cols + c] = hi;
          op[2 * total + (size_t)r * cols + c] = round_tf32(v - hi);
        } else {
          op[(size_t)r * cols + c] = v;
        }
      }
    } else {
      t[j][threadIdx.x] = v;
    }
  }
  if (!transpose) return;
  __syncthreads();
  const int r2 = blockIdx.y * 32 + threadIdx.x;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c2 = blockIdx.x * 32 + j;
    if (r2 < rows && c2 < cols) {
      const float v = t[threadIdx.x][j];
      if (round == 2) {
        const float hi = round_tf32(v);
        op[(size_t)c2 * rows + r2] = hi;
        op[total + (size_t)c2 * rows + r2] = hi;
        op[2 * total + (size_t)c2 * rows + r2] = round_tf32(v - hi);
      } else {
        op[(size_t)c2 * rows + r2] = v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// host side: tensor-map construction (driver entry point fetched at run time) + cache
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKey {
  const void* p;
  long long rows, pitch;
  int cols, box_rows, atom32;
  bool operator==(const MapKey& o) const {
    return p == o.p && rows == o.rows && pitch == o.pitch && cols == o.cols && box_rows == o.box_rows &&
           atom32 == o.atom32;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = std::hash<const void*>()(k.p);
    h ^= std::hash<long long>()(k.rows * 1315423911LL + k.pitch) + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2);
    h ^= std::hash<long long>()(((long long)k.cols << 20) ^ (k.box_rows << 1) ^ k.atom32) + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2);
    return h;
  }
};

// 2D fp32 tensor [rows][cols] with row pitch `pitch` elements; box = 32 columns x box_rows rows, 128B swizzle
// atom32 != 0 selects the 32-byte-chunk flavour of the 128B swizzle (MN-major tf32 operands).
int get_map(const float* p, long long rows, int cols, long long pitch, int box_rows, CUtensorMap* out, int atom32) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key{p, rows, pitch, cols, box_rows, atom32};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return MYOLO_OK;
    }
  }
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return MYOLO_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)pitch * sizeof(float)};
  cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(p), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for [%lld x %d] pitch %lld box_rows %d", (int)r, rows, cols, pitch, box_rows);
    return MYOLO_ERR_CUDA;
  }
  {
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 4096) cache.clear();
    cache[key] = m;
  }
  *out = m;
  return MYOLO_OK;
}

// IEEE-half flavour (operands and outputs of the kind::f16 kernels).
int get_map_h(const void* p, long long rows, int cols, long long pitch, int box_rows, int box_cols, CUtensorMap* out) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key{p, rows, pitch, cols, box_rows, 1000 + box_cols};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return MYOLO_OK;
    }
  }
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return MYOLO_ERR_CUDA;
  }
  if ((box_cols != 64 && box_cols != 32) || (pitch % 8) != 0 || ((uintptr_t)p & 15) != 0) {
    set_error("get_map_h: unsupported half map (box_cols %d, pitch %lld)", box_cols, pitch);
    return MYOLO_ERR_ARG;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)pitch * 2u};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(p), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (half) failed (%d) for [%lld x %d] pitch %lld box %d x %d", (int)r, rows, cols, pitch,
              box_rows, box_cols);
    return MYOLO_ERR_CUDA;
  }
  {
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 4096) cache.clear();
    cache[key] = m;
  }
  *out = m;
  return MYOLO_OK;
}

static int pick_bn(int N) {
  if (N % 256 == 0) return 256;
  if (N % 128 == 0) return 128;
  if (N % 64 == 0) return 64;
  if (N % 32 == 0) return 32;
  return 0;
}

template <int BN>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, float* C, long long ldc, long long M, int N, int K,
                       int ntaps, const TapShifts& sh, const Epi& ep, cudaStream_t st) {
  const int total = ntaps * (K / BK);
  const int stages = total < 4 ? total : 4;
  const size_t smem = (size_t)stages * (BM + BN) * BK * 4 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    MYOLO_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * (BM + BN) * BK * 4 + 1024));
    attr_set = true;
  }
  dim3 grid((unsigned)ceil_div(M, BM), (unsigned)(N / BN));
  MYOLO_LAUNCH(tc_gemm_kernel<BN>, grid, kThreads, smem, st, ta, tb, C, ldc, M, N, K, ntaps, sh, ep, stages);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

template <int BN, int NACC>
static int launch_wgrad_n(const CUtensorMap& ta, const CUtensorMap& td, float* dW, long long M, int N, int K, int ntaps,
                          const TapShifts& sh, int transpose_out, cudaStream_t st);

template <int BN>
static int launch_wgrad(const CUtensorMap& ta, const CUtensorMap& td, float* dW, long long M, int N, int K, int ntaps,
                        const TapShifts& sh, int transpose_out, cudaStream_t st) {
  if (BN * 2 <= 512 && (K % (2 * BM)) == 0) return launch_wgrad_n<BN, (BN * 2 <= 512 ? 2 : 1)>(ta, td, dW, M, N, K, ntaps, sh, transpose_out, st);
  return launch_wgrad_n<BN, 1>(ta, td, dW, M, N, K, ntaps, sh, transpose_out, st);
}

template <int BN, int NACC>
static int launch_wgrad_n(const CUtensorMap& ta, const CUtensorMap& td, float* dW, long long M, int N, int K, int ntaps,
                          const TapShifts& sh, int transpose_out, cudaStream_t st) {
  const int ntk = (K + BM * NACC - 1) / (BM * NACC), ntn = N / BN;
  const long long tiles = (long long)ntk * ntn * ntaps;
  long long nsplit = max(1LL, min(ceil_div(M, 32 * 8), (long long)kNumSMs / tiles));
  if (nsplit < 1) nsplit = 1;
  long long chunk = ceil_div(ceil_div(M, nsplit), 32) * 32;
  nsplit = ceil_div(M, chunk);
  const size_t per_stage = (size_t)(BM * NACC + BN) * BK * 4;
  const int stages = per_stage * 4 + 1024 <= 220 * 1024 ? 4 : 3;
  const size_t smem = (size_t)stages * per_stage + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    MYOLO_CUDA(cudaFuncSetAttribute(tc_wgrad_kernel<BN, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  dim3 grid((unsigned)nsplit, (unsigned)(ntk * ntn), (unsigned)ntaps);
  MYOLO_LAUNCH((tc_wgrad_kernel<BN, NACC>), grid, kThreads, smem, st, ta, td, dW, M, N, K, sh, chunk, ntn, transpose_out, stages);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

}  // namespace tc
}  // namespace myolo

using namespace myolo;
using namespace myolo::tc;

static bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

extern "C" int myolo_gemm_taps_tc_supported(long long lda, long long ldc, long long M, int N, int K, int ntaps,
                                            int accumulate) {
  (void)accumulate;
  return M >= 1 && M < (1LL << 31) - 4096 && (K % BK) == 0 && pick_bn(N) != 0 && (lda % 4) == 0 && (ldc % 4) == 0 &&
         ntaps >= 1 && ntaps <= 32 && (long long)ntaps * N < (1LL << 31);
}

static int gemm_taps_tc_impl(const float* A, long long lda, const float* Bt, float* C, long long ldc, long long M, int N, int K,
                             int ntaps, const int* shifts_host, const float* bias, const float* scale, const float* shift_c,
                             int act, int pf_w1, int pf_blk, int accumulate, float* st_mean, float* st_var, double* ws,
                             long long n_valid, myolo_stream stream) {
  MYOLO_CHECK_ARG(A && Bt && C && aligned16(A) && aligned16(Bt) && aligned16(C));
  MYOLO_CHECK_ARG(myolo_gemm_taps_tc_supported(lda, ldc, M, N, K, ntaps, accumulate));
  MYOLO_CHECK_ARG((scale == nullptr) == (shift_c == nullptr));
  MYOLO_CHECK_ARG(!(accumulate && (scale || (act & 0xff) != MYOLO_ACT_NONE)));
  MYOLO_CHECK_ARG(pf_w1 <= 0 || pf_blk > 0);
  TapShifts sh;
  for (int t = 0; t < 32; ++t) sh.s[t] = (shifts_host && t < ntaps) ? shifts_host[t] : 0;
  const int bn = pick_bn(N);
  CUtensorMap ta, tb;
  // rows below 0 are TMA zero fill; rows past M are real memory (zero guard rows of a padded-flat
  // tensor, or the low-part copy of a split operand), so the map extends over the largest shift
  long long a_rows = M;
  for (int t = 0; t < ntaps; ++t) a_rows = max(a_rows, M + (long long)sh.s[t]);
  int rc = get_map(A, a_rows, K, lda, BM, &ta);
  if (rc) return rc;
  rc = get_map(Bt, (long long)ntaps * N, K, K, bn, &tb);
  if (rc) return rc;
  Epi ep{bias, scale, shift_c, act, pf_w1, pf_blk, accumulate};
  if (st_mean) {      // statistics of the stored result in the epilogue: workspace layout of the BN family (bn.cu)
    MYOLO_CHECK_ARG(st_var && ws && N <= 1024 && n_valid > 0 && !accumulate);
    ep.st_sums = ws + 16;
    ep.st_ticket = reinterpret_cast<int*>(ws);
    ep.st_mean = st_mean;
    ep.st_var = st_var;
    ep.st_inv = 1.0 / (double)n_valid;
  }
  cudaStream_t st = as_stream(stream);
  switch (bn) {
    case 256: return launch_gemm<256>(ta, tb, C, ldc, M, N, K, ntaps, sh, ep, st);
    case 128: return launch_gemm<128>(ta, tb, C, ldc, M, N, K, ntaps, sh, ep, st);
    case 64: return launch_gemm<64>(ta, tb, C, ldc, M, N, K, ntaps, sh, ep, st);
    default: return launch_gemm<32>(ta, tb, C, ldc, M, N, K, ntaps, sh, ep, st);
  }
}

extern "C" int myolo_gemm_taps_tc(const float* A, long long lda, const float* Bt, float* C, long long ldc, long long M,
                                  int N, int K, int ntaps, const int* shifts_host, const float* bias,
                                  const float* scale, const float* shift_c, int act, int pf_w1, int pf_blk,
                                  int accumulate, myolo_stream stream) {
  return gemm_taps_tc_impl(A, lda, Bt, C, ldc, M, N, K, ntaps, shifts_host, bias, scale, shift_c, act, pf_w1, pf_blk, accumulate,
                           nullptr, nullptr, nullptr, 0, stream);
}

// the same GEMM with the batch statistics of its result (per output channel, over the valid rows: n_valid of them) taken
// in the epilogue -- the BatchNormalization that follows a pointwise convolution needs no statistics pass of its own
extern "C" int myolo_gemm_taps_tc_stats(const float* A, long long lda, const float* Bt, float* C, long long ldc, long long M,
                                        int N, int K, int ntaps, const int* shifts_host, int pf_w1, int pf_blk,
                                        float* mean, float* var, double* ws, long long n_valid, myolo_stream stream) {
  MYOLO_CHECK_ARG(mean && var && ws);
  return gemm_taps_tc_impl(A, lda, Bt, C, ldc, M, N, K, ntaps, shifts_host, nullptr, nullptr, nullptr, MYOLO_ACT_NONE, pf_w1,
                           pf_blk, 0, mean, var, ws, n_valid, stream);
}

extern "C" int myolo_gemm_taps_wgrad_tc_supported(long long lda, long long ldd, long long M, int N, int K, int ntaps) {
  // K (A channels) only needs whole 32-channel TMA boxes: channels past K inside the last 128-channel tile are
  // out-of-bounds boxes (zero fill) and their accumulator rows are never stored
  return M >= 32 && M < (1LL << 31) - 4096 && (K % 32) == 0 && pick_bn(N) != 0 && (lda % 4) == 0 && (ldd % 4) == 0 &&
         ntaps >= 1 && ntaps <= 32;
}

extern "C" int myolo_gemm_taps_wgrad_tc(const float* A, long long lda, const float* D, long long ldd, float* dW,
                                        long long M, int N, int K, int ntaps, const int* shifts_host,
                                        int transpose_out, myolo_stream stream) {
  MYOLO_CHECK_ARG(A && D && dW && aligned16(A) && aligned16(D));
  MYOLO_CHECK_ARG(transpose_out || aligned16(dW));     // 128-bit vector reductions into dW
  MYOLO_CHECK_ARG(myolo_gemm_taps_wgrad_tc_supported(lda, ldd, M, N, K, ntaps));
  TapShifts sh;
  for (int t = 0; t < 32; ++t) sh.s[t] = (shifts_host && t < ntaps) ? shifts_host[t] : 0;
  const int bn = pick_bn(N);
  CUtensorMap ta, td;
  int rc = get_map(A, M, K, lda, 32, &ta, 1);
  if (rc) return rc;
  rc = get_map(D, M, N, ldd, 32, &td, 1);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  switch (bn) {
    case 256: return launch_wgrad<256>(ta, td, dW, M, N, K, ntaps, sh, transpose_out, st);
    case 128: return launch_wgrad<128>(ta, td, dW, M, N, K, ntaps, sh, transpose_out, st);
    case 64: return launch_wgrad<64>(ta, td, dW, M, N, K, ntaps, sh, transpose_out, st);
    default: return launch_wgrad<32>(ta, td, dW, M, N, K, ntaps, sh, transpose_out, st);
  }
}

// ---- every weight staging job of a step in ONE launch ------------------------------------------------------------
// mode: 0 copy, 1 tf32 rounding, 2 3xTF32 split ([hi | hi | lo] stacked copies), 3 IEEE half.  A block handles one
// 32x32 tile of one tap of one job; tile_begin is the running sum of tiles over the jobs (binary search by block).
namespace myolo {
namespace tc {
struct PrepJob {
  const float* in;
  void* out;
  int ntaps, rows, cols, transpose, mode, tile_begin;
  int out_ld;      // leading dimension of the staged matrix (0 = dense: rows when transposing, cols otherwise); a larger
                   // value leaves zero padding columns / rows that the job never writes (conv_23: 27 -> 32 channels)
  int out_total;   // elements between the hi / hi / lo copies of mode 2 (0 = ntaps*rows*cols)
};
__global__ void prep_weights_batch_kernel(const PrepJob* __restrict__ jobs, int n_jobs) {
  pdl_entry();
  __shared__ float t[32][33];
  int lo = 0, hi = n_jobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].tile_begin <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const PrepJob jb = jobs[lo];
  const int tx = (jb.cols + 31) / 32, ty = (jb.rows + 31) / 32;
  int tile = (int)blockIdx.x - jb.tile_begin;
  const int tap = tile / (tx * ty);
  tile -= tap * tx * ty;
  const int by = tile / tx, bx = tile - by * tx;
  const size_t per_tap = (size_t)jb.rows * jb.cols, total = jb.out_total > 0 ? (size_t)jb.out_total : per_tap * jb.ntaps;
  const size_t ld_t = jb.out_ld > 0 ? (size_t)jb.out_ld : (size_t)jb.rows, ld_n = jb.out_ld > 0 ? (size_t)jb.out_ld : (size_t)jb.cols;
  const float* ip = jb.in + tap * per_tap;
  float* of = reinterpret_cast<float*>(jb.out) + tap * per_tap;
  uint16_t* oh = reinterpret_cast<uint16_t*>(jb.out) + tap * per_tap;
  auto emit = [&](size_t idx, float v) {
    if (jb.mode == 3) oh[idx] = f2h_sat(v);
    else if (jb.mode == 2) {
      const float h = round_tf32(v);
      of[idx] = h;
      of[total + idx] = h;
      of[2 * total + idx] = round_tf32(v - h);
    } else of[idx] = jb.mode == 1 ? round_tf32(v) : v;
  };
  const int c = bx * 32 + threadIdx.x;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int r = by * 32 + j;
    const float v = (r < jb.rows && c < jb.cols) ? ip[(size_t)r * jb.cols + c] : 0.f;
    if (!jb.transpose) {
      if (r < jb.rows && c < jb.cols) emit((size_t)r * ld_n + c, v);
    } else {
      t[j][threadIdx.x] = v;
    }
  }
  if (!jb.transpose) return;
  __syncthreads();
  const int r2 = by * 32 + threadIdx.x;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c2 = bx * 32 + j;
    if (r2 < jb.rows && c2 < jb.cols) emit((size_t)c2 * ld_t + r2, t[threadIdx.x][j]);
  }
}

// dst[r][0..cols) = src[r][0..cols) for two row pitches (columns of dst beyond `cols` are left alone)
__global__ void copy_cols_kernel(const float* __restrict__ src, long long src_ld, float* __restrict__ dst, long long dst_ld,
                                 long long rows, int cols) {
  pdl_entry();
  const long long total = rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = (int)(i - r * cols);
    dst[r * dst_ld + c] = src[r * src_ld + c];
  }
}
}  // namespace tc
}  // namespace myolo

namespace myolo {
namespace tc {
// tile i of the compact buffer <-> tile list[i] of the full buffer, 16 bytes per thread and step
__global__ void __launch_bounds__(256) copy_tiles_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst,
                                                         const int* __restrict__ list, long long tile_vec, int scatter) {
  pdl_entry();
  const long long t = __ldg(list + blockIdx.y);
  const uint4* s = src + (scatter ? (long long)blockIdx.y : t) * tile_vec;
  uint4* d = dst + (scatter ? t : (long long)blockIdx.y) * tile_vec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < tile_vec; i += (long long)gridDim.x * blockDim.x) d[i] = s[i];
}
}  // namespace tc
}  // namespace myolo

extern "C" int myolo_copy_tiles(const void* src, void* dst, const int* list, int n_list, long long tile_bytes, int scatter,
                                myolo_stream stream) {
  MYOLO_CHECK_ARG(src && dst && list && n_list > 0 && n_list <= 65535 && tile_bytes > 0 && (tile_bytes % 16) == 0);
  MYOLO_CHECK_ARG(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0);
  const long long tv = tile_bytes / 16;
  dim3 grid((unsigned)max(1LL, min(ceil_div(tv, 256), 64LL)), (unsigned)n_list);
  MYOLO_LAUNCH(copy_tiles_kernel, grid, 256, 0, as_stream(stream), reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst), list,
                                                         tv, scatter);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

extern "C" int myolo_copy_cols(const float* src, long long src_ld, float* dst, long long dst_ld, long long rows, int cols,
                               myolo_stream stream) {
  MYOLO_CHECK_ARG(src && dst && rows > 0 && cols > 0 && src_ld >= cols && dst_ld >= cols);
  const int blocks = (int)max(1LL, min(ceil_div(rows * cols, 256), (long long)kNumSMs * 8));
  MYOLO_LAUNCH(copy_cols_kernel, blocks, 256, 0, as_stream(stream), src, src_ld, dst, dst_ld, rows, cols);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

extern "C" int myolo_prep_weights_batch(const void* jobs_dev, int n_jobs, int total_tiles, myolo_stream stream) {
  MYOLO_CHECK_ARG(jobs_dev && n_jobs > 0 && total_tiles > 0);
  MYOLO_LAUNCH(prep_weights_batch_kernel, total_tiles, dim3(32, 8), 0, as_stream(stream), reinterpret_cast<const PrepJob*>(jobs_dev), n_jobs);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

namespace myolo {
namespace tc {
// half staging of a weight block: out[t][c][r] = half(in[t][r][c]) (transpose) or out = half(in)
__global__ void prep_weights_h_kernel(const float* __restrict__ in, uint16_t* __restrict__ out, int rows, int cols,
                                      int transpose) {
  pdl_entry();
  __shared__ float t[32][33];
  const float* ip = in + (size_t)blockIdx.z * rows * cols;
  uint16_t* op = out + (size_t)blockIdx.z * rows * cols;
  const int c = blockIdx.x * 32 + threadIdx.x;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int r = blockIdx.y * 32 + j;
    const float v = (r < rows && c < cols) ? ip[(size_t)r * cols + c] : 0.f;
    if (!transpose) {
      if (r < rows && c < cols) op[(size_t)r * cols + c] = f2h_sat(v);
    } else {
      t[j][threadIdx.x] = v;
    }
  }
  if (!transpose) return;
  __syncthreads();
  const int r2 = blockIdx.y * 32 + threadIdx.x;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c2 = blockIdx.x * 32 + j;
    if (r2 < rows && c2 < cols) op[(size_t)c2 * rows + r2] = f2h_sat(t[threadIdx.x][j]);
  }
}
}  // namespace tc
}  // namespace myolo

extern "C" int myolo_prep_weights_h(const float* in, void* out_half, int ntaps, int rows, int cols, int transpose,
                                    myolo_stream stream) {
  MYOLO_CHECK_ARG(in && out_half && (const void*)in != out_half && ntaps > 0 && rows > 0 && cols > 0);
  dim3 grid((cols + 31) / 32, (rows + 31) / 32, ntaps), block(32, 8);
  MYOLO_LAUNCH(prep_weights_h_kernel, grid, block, 0, as_stream(stream), in, reinterpret_cast<uint16_t*>(out_half), rows, cols, transpose);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

extern "C" int myolo_prep_weights(const float* in, float* out, int ntaps, int rows, int cols, int transpose,
                                  int round_tf32, myolo_stream stream) {
  MYOLO_CHECK_ARG(in && out && in != out && ntaps > 0 && rows > 0 && cols > 0 && round_tf32 >= 0 && round_tf32 <= 2);
  dim3 grid((cols + 31) / 32, (rows + 31) / 32, ntaps), block(32, 8);
  MYOLO_LAUNCH(prep_weights_kernel, grid, block, 0, as_stream(stream), in, out, rows, cols, transpose, round_tf32,
                                                             (size_t)ntaps * rows * cols);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}
