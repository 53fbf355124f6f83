// gemm_ffma.cu -- fp32 CUDA-core tap-GEMM family (exact-fp32 path).
//
//   forward / dgrad :  C[m,n] = epi( sum_t sum_k A[m + shift_t, k] * Bt[t][n][k] )
//   wgrad           :  dW_t[k,n] += sum_m A[m + shift_t, k] * D[m, n]
//
// A "tap" is one (dy,dx) offset of a 3x3 SAME convolution evaluated on a padded-flat tensor
// (DESIGN.md section 3): the spatial shift is a constant row offset, so a 3x3 conv is nine shifted
// GEMMs accumulated in registers and a 1x1 conv / transposed conv is the 1-tap special case.
// These kernels are the full-precision reference path of the library and the fallback for shapes
// the tcgen05 kernel (gemm_tcgen05.cu) does not take (tiny N, K not a multiple of 32).
// 128x128x8 block tile, 256 threads, 8x8 register tile (2x2 groups of 4x4 to keep shared-memory
// reads conflict-free), double-buffered shared memory with register prefetch.
#include "common.cuh"

namespace myolo {

struct TapShifts {
  int s[32];
};

constexpr int BM = 128, BN = 128, BK = 8;

__device__ __forceinline__ bool pf_valid(long long m, int pf_w1, int pf_blk) {
  if (pf_w1 <= 0) return true;
  const int r = (int)(m % pf_blk);
  return (r / pf_w1) >= 1 && (r % pf_w1) >= 1;
}

__global__ void __launch_bounds__(256)
sgemm_taps_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ B, float* __restrict__ C,
                  long long ldc, long long M, int N, int K, int ntaps, TapShifts sh, const float* __restrict__ bias,
                  const float* __restrict__ scale, const float* __restrict__ shift_c, int act, int pf_w1, int pf_blk,
                  int accumulate) {
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const bool nvec = ((N & 3) == 0);
  // load assignments: A row ar / Bt row ar (output channel n0+ar), k offset akq..akq+3
  const int ar = tid >> 1, akq = (tid & 1) * 4;
  const bool arow_ok = (m0 + ar) < M;
  const bool brow_ok = (n0 + ar) < N;
  const int kiters = (K + BK - 1) / BK;
  const int total = ntaps * kiters;
  // fast path: 128-bit loads need 16-byte aligned rows and a K that is a multiple of the k-tile
  const bool avec = ((lda & 3) == 0) && ((K & (BK - 1)) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
  const bool bvec = ((K & (BK - 1)) == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra, rb;
  auto gload = [&](int it) {
    const int t = it / kiters;
    const int k0 = (it - t * kiters) * BK;
    ra = make_float4(0.f, 0.f, 0.f, 0.f);
    if (arow_ok) {
      const float* ap = A + (m0 + ar + sh.s[t]) * lda + k0 + akq;
      if (avec) {
        ra = __ldg(reinterpret_cast<const float4*>(ap));
      } else {
        const int kk = k0 + akq;
        ra.x = (kk + 0 < K) ? __ldg(ap + 0) : 0.f;
        ra.y = (kk + 1 < K) ? __ldg(ap + 1) : 0.f;
        ra.z = (kk + 2 < K) ? __ldg(ap + 2) : 0.f;
        ra.w = (kk + 3 < K) ? __ldg(ap + 3) : 0.f;
      }
    }
    rb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (brow_ok) {
      const float* bp = B + ((size_t)t * N + n0 + ar) * K + k0 + akq;
      if (bvec) {
        rb = __ldg(reinterpret_cast<const float4*>(bp));
      } else {
        const int kk = k0 + akq;
        rb.x = (kk + 0 < K) ? __ldg(bp + 0) : 0.f;
        rb.y = (kk + 1 < K) ? __ldg(bp + 1) : 0.f;
        rb.z = (kk + 2 < K) ? __ldg(bp + 2) : 0.f;
        rb.w = (kk + 3 < K) ? __ldg(bp + 3) : 0.f;
      }
    }
  };
  auto sstore = [&](int buf) {
    As[buf][akq + 0][ar] = ra.x;
    As[buf][akq + 1][ar] = ra.y;
    As[buf][akq + 2][ar] = ra.z;
    As[buf][akq + 3][ar] = ra.w;
    Bs[buf][akq + 0][ar] = rb.x;
    Bs[buf][akq + 1][ar] = rb.y;
    Bs[buf][akq + 2][ar] = rb.z;
    Bs[buf][akq + 3][ar] = rb.w;
  };

  gload(0);
  sstore(0);
  __syncthreads();
  for (int it = 0; it < total; ++it) {
    const int buf = it & 1;
    if (it + 1 < total) gload(it + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (it + 1 < total) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M || !pf_valid(m, pf_w1, pf_blk)) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int nb = n0 + jh * 64 + tx * 4;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = nb + j;
        float t = acc[i][jh * 4 + j];
        if (n < N) {
          if (bias) t += __ldg(bias + n);
          if (scale) t = fmaf(t, __ldg(scale + n), __ldg(shift_c + n));
          t = apply_act(t, act);
        }
        v[j] = t;
      }
      float* cp = C + m * ldc + nb;
      if (nvec && ((ldc & 3) == 0) && nb + 3 < N) {
        float4 o = make_float4(v[0], v[1], v[2], v[3]);
        if (accumulate) {
          const float4 old = *reinterpret_cast<const float4*>(cp);
          o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
        }
        *reinterpret_cast<float4*>(cp) = o;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (nb + j < N) cp[j] = accumulate ? cp[j] + v[j] : v[j];
      }
    }
  }
}

// wgrad: tile = 128 (k of A) x 128 (n of D); the reduction dimension is the row index m, split
// across blockIdx.x.  fp32 atomics into dW.
__global__ void __launch_bounds__(256)
sgemm_taps_wgrad_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ D, long long ldd,
                        float* __restrict__ dW, long long M, int N, int K, TapShifts sh, long long chunk,
                        int ntn, int transpose_out) {
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Ds[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int tap = blockIdx.z;
  const int k0 = (blockIdx.y / ntn) * BM;
  const int n0 = (blockIdx.y % ntn) * BN;
  const long long mbeg = (long long)blockIdx.x * chunk;
  const long long mend = min(M, mbeg + chunk);
  if (mbeg >= mend) return;
  const int shift = sh.s[tap];
  const int lr = tid >> 5, lq = (tid & 31) * 4;
  const bool nvec = ((N & 3) == 0) && ((ldd & 3) == 0);

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra, rd;
  auto gload = [&](long long m) {
    ra = make_float4(0.f, 0.f, 0.f, 0.f);
    rd = ra;
    const long long r = m + lr;
    if (r < mend) {
      if (k0 + lq + 3 < K) {
        ra = __ldg(reinterpret_cast<const float4*>(A + (r + shift) * lda + k0 + lq));
      } else {
        const float* ap = A + (r + shift) * lda + k0 + lq;
        ra.x = (k0 + lq + 0 < K) ? __ldg(ap + 0) : 0.f;
        ra.y = (k0 + lq + 1 < K) ? __ldg(ap + 1) : 0.f;
        ra.z = (k0 + lq + 2 < K) ? __ldg(ap + 2) : 0.f;
        ra.w = (k0 + lq + 3 < K) ? __ldg(ap + 3) : 0.f;
      }
      const float* dp = D + r * ldd + n0 + lq;
      if (nvec && n0 + lq + 3 < N) {
        rd = __ldg(reinterpret_cast<const float4*>(dp));
      } else {
        rd.x = (n0 + lq + 0 < N) ? __ldg(dp + 0) : 0.f;
        rd.y = (n0 + lq + 1 < N) ? __ldg(dp + 1) : 0.f;
        rd.z = (n0 + lq + 2 < N) ? __ldg(dp + 2) : 0.f;
        rd.w = (n0 + lq + 3 < N) ? __ldg(dp + 3) : 0.f;
      }
    }
  };
  auto sstore = [&](int buf) {
    *reinterpret_cast<float4*>(&As[buf][lr][lq]) = ra;
    *reinterpret_cast<float4*>(&Ds[buf][lr][lq]) = rd;
  };

  gload(mbeg);
  sstore(0);
  __syncthreads();
  int buf = 0;
  for (long long m = mbeg; m < mend; m += BK) {
    const bool more = (m + BK) < mend;
    if (more) gload(m + BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Ds[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Ds[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      sstore(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }
  float* W = dW + (size_t)tap * K * N;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = k0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (k >= K) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= N) continue;
      if (transpose_out)
        atomicAdd(W + (size_t)n * K + k, acc[i][j]);
      else
        atomicAdd(W + (size_t)k * N + n, acc[i][j]);
    }
  }
}

}  // namespace myolo

using namespace myolo;

extern "C" int myolo_gemm_taps_ffma(const float* A, long long lda, const float* B, float* C, long long ldc, long long M,
                                    int N, int K, int ntaps, const int* shifts_host, const float* bias,
                                    const float* scale, const float* shift_c, int act, int pf_w1, int pf_blk,
                                    int accumulate, myolo_stream stream) {
  MYOLO_CHECK_ARG(A && B && C && M > 0 && N > 0 && K > 0 && lda >= K);
  MYOLO_CHECK_ARG(ntaps >= 1 && ntaps <= 32);
  MYOLO_CHECK_ARG((scale == nullptr) == (shift_c == nullptr));
  MYOLO_CHECK_ARG(!(accumulate && (scale || (act & 0xff) != MYOLO_ACT_NONE)));
  MYOLO_CHECK_ARG(pf_w1 <= 0 || pf_blk > 0);
  TapShifts sh;
  for (int t = 0; t < 32; ++t) sh.s[t] = (shifts_host && t < ntaps) ? shifts_host[t] : 0;
  dim3 grid((unsigned)ceil_div(M, BM), (unsigned)ceil_div(N, BN));
  sgemm_taps_kernel<<<grid, 256, 0, as_stream(stream)>>>(A, lda, B, C, ldc, M, N, K, ntaps, sh, bias, scale, shift_c,
                                                         act, pf_w1, pf_blk, accumulate);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

extern "C" int myolo_gemm_taps_wgrad_ffma(const float* A, long long lda, const float* D, long long ldd, float* dW,
                                     long long M, int N, int K, int ntaps, const int* shifts_host, int transpose_out,
                                     myolo_stream stream) {
  MYOLO_CHECK_ARG(A && D && dW && M > 0 && N > 0 && K > 0 && (K % 4) == 0 && (lda % 4) == 0);
  MYOLO_CHECK_ARG(ntaps >= 1 && ntaps <= 32);
  TapShifts sh;
  for (int t = 0; t < 32; ++t) sh.s[t] = (shifts_host && t < ntaps) ? shifts_host[t] : 0;
  const int ntk = (int)ceil_div(K, BM), ntn = (int)ceil_div(N, BN);
  const long long tiles = (long long)ntk * ntn * ntaps;
  long long nsplit = max(1LL, min(ceil_div(M, 256), (long long)(kNumSMs * 4) / tiles + 1));
  long long chunk = ceil_div(ceil_div(M, nsplit), BK) * BK;
  nsplit = ceil_div(M, chunk);
  dim3 grid((unsigned)nsplit, (unsigned)(ntk * ntn), (unsigned)ntaps);
  sgemm_taps_wgrad_kernel<<<grid, 256, 0, as_stream(stream)>>>(A, lda, D, ldd, dW, M, N, K, sh, chunk, ntn,
                                                               transpose_out);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}
