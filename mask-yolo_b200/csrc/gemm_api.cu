// gemm_api.cu -- precision mode, dispatch and the named conv wrappers of the tap-GEMM family.
// myolo_gemm_taps / myolo_gemm_taps_wgrad pick the tcgen05 kernel (gemm_tcgen05.cu) when the
// library is in TF32 mode and the shape qualifies, otherwise the exact-fp32 CUDA-core kernel
// (gemm_ffma.cu).  Both are sm_100a device code; there is no host or library fallback.
#include <stdlib.h>
#include "common.cuh"

extern "C" int myolo_gemm_taps_ffma(const float* A, long long lda, const float* Bt, float* C, long long ldc,
                                    long long M, int N, int K, int ntaps, const int* shifts_host, const float* bias,
                                    const float* scale, const float* shift_c, int act, int pf_w1, int pf_blk,
                                    int accumulate, myolo_stream stream);
extern "C" int myolo_gemm_taps_tc(const float* A, long long lda, const float* Bt, float* C, long long ldc, long long M,
                                  int N, int K, int ntaps, const int* shifts_host, const float* bias,
                                  const float* scale, const float* shift_c, int act, int pf_w1, int pf_blk,
                                  int accumulate, myolo_stream stream);
extern "C" int myolo_gemm_taps_tc_supported(long long lda, long long ldc, long long M, int N, int K, int ntaps, int accumulate);
extern "C" int myolo_gemm_taps_wgrad_ffma(const float* A, long long lda, const float* D, long long ldd, float* dW,
                                          long long M, int N, int K, int ntaps, const int* shifts_host,
                                          int transpose_out, myolo_stream stream);
extern "C" int myolo_gemm_taps_wgrad_tc(const float* A, long long lda, const float* D, long long ldd, float* dW,
                                        long long M, int N, int K, int ntaps, const int* shifts_host,
                                        int transpose_out, myolo_stream stream);
extern "C" int myolo_gemm_taps_wgrad_tc_supported(long long lda, long long ldd, long long M, int N, int K, int ntaps);

extern "C" int myolo_gemm_taps_win(const float* A, long long lda, const float* Bt, float* C, long long ldc, long long M,
                                   int N, int K, int ntaps, const int* shifts_host, const float* bias, const float* scale,
                                   const float* shift_c, int act, int pf_w1, int pf_blk, int accumulate, myolo_stream stream);
extern "C" int myolo_gemm_taps_win_supported(long long lda, long long ldc, long long M, int N, int K, int ntaps,
                                             const int* shifts_host, int accumulate);

namespace myolo {
static int g_precision = MYOLO_PREC_FP32;
static int use_win() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MYOLO_NO_WIN");
    v = (e && atoi(e)) ? 0 : 1;
  }
  return v;
}
}

extern "C" int myolo_set_precision(int mode) {
  MYOLO_CHECK_ARG(mode == MYOLO_PREC_FP32 || mode == MYOLO_PREC_TF32);
  myolo::g_precision = mode;
  return MYOLO_OK;
}
extern "C" int myolo_get_precision(void) { return myolo::g_precision; }

extern "C" int myolo_gemm_taps(const float* A, long long lda, const float* Bt, float* C, long long ldc, long long M,
                               int N, int K, int ntaps, const int* shifts_host, const float* bias, const float* scale,
                               const float* shift_c, int act, int pf_w1, int pf_blk, int accumulate,
                               myolo_stream stream) {
  if (myolo::g_precision == MYOLO_PREC_TF32 && myolo::use_win() &&
      myolo_gemm_taps_win_supported(lda, ldc, M, N, K, ntaps, shifts_host, accumulate))
    return myolo_gemm_taps_win(A, lda, Bt, C, ldc, M, N, K, ntaps, shifts_host, bias, scale, shift_c, act, pf_w1, pf_blk,
                               accumulate, stream);
  if (myolo::g_precision == MYOLO_PREC_TF32 && myolo_gemm_taps_tc_supported(lda, ldc, M, N, K, ntaps, accumulate))
    return myolo_gemm_taps_tc(A, lda, Bt, C, ldc, M, N, K, ntaps, shifts_host, bias, scale, shift_c, act, pf_w1, pf_blk,
                              accumulate, stream);
  return myolo_gemm_taps_ffma(A, lda, Bt, C, ldc, M, N, K, ntaps, shifts_host, bias, scale, shift_c, act, pf_w1, pf_blk,
                              accumulate, stream);
}

extern "C" int myolo_gemm_taps_wgrad(const float* A, long long lda, const float* D, long long ldd, float* dW,
                                     long long M, int N, int K, int ntaps, const int* shifts_host, int transpose_out,
                                     myolo_stream stream) {
  if (myolo::g_precision == MYOLO_PREC_TF32 && myolo_gemm_taps_wgrad_tc_supported(lda, ldd, M, N, K, ntaps))
    return myolo_gemm_taps_wgrad_tc(A, lda, D, ldd, dW, M, N, K, ntaps, shifts_host, transpose_out, stream);
  return myolo_gemm_taps_wgrad_ffma(A, lda, D, ldd, dW, M, N, K, ntaps, shifts_host, transpose_out, stream);
}

static void conv3x3_shifts(int W, int* s, bool negate) {
  for (int dy = 0; dy < 3; ++dy)
    for (int dx = 0; dx < 3; ++dx) {
      const int v = (dy - 1) * (W + 1) + (dx - 1);
      s[dy * 3 + dx] = negate ? -v : v;
    }
}

extern "C" int myolo_pwconv_fwd(const float* x, const float* wt, float* y, long long M, int Cin, int Cout,
                                const float* bias, myolo_stream stream) {
  return myolo_gemm_taps(x, Cin, wt, y, Cout, M, Cout, Cin, 1, nullptr, bias, nullptr, nullptr, MYOLO_ACT_NONE, 0, 0, 0, stream);
}
extern "C" int myolo_pwconv_dgrad(const float* dy, const float* w, float* dx, long long M, int Cin, int Cout,
                                  myolo_stream stream) {
  // dx[m,ci] = sum_co dy[m,co] * w[ci][co]  -> Bt = w ([N=Cin][K=Cout])
  return myolo_gemm_taps(dy, Cout, w, dx, Cin, M, Cin, Cout, 1, nullptr, nullptr, nullptr, nullptr, MYOLO_ACT_NONE, 0, 0, 0, stream);
}
extern "C" int myolo_pwconv_wgrad(const float* x, const float* dy, float* dw, long long M, int Cin, int Cout,
                                  myolo_stream stream) {
  MYOLO_CUDA(cudaMemsetAsync(dw, 0, (size_t)Cin * Cout * sizeof(float), myolo::as_stream(stream)));
  return myolo_gemm_taps_wgrad(x, Cin, dy, Cout, dw, M, Cout, Cin, 1, nullptr, 0, stream);
}

extern "C" int myolo_conv3x3_fwd(const float* x, const float* wt, float* y, int n_img, int H, int W, int Cin, int Cout,
                                 const float* bias, const float* scale, const float* shift_c, int act,
                                 myolo_stream stream) {
  MYOLO_CHECK_ARG(n_img > 0 && H > 0 && W > 0);
  int s[9];
  conv3x3_shifts(W, s, false);
  const long long M = (long long)n_img * (H + 1) * (W + 1);
  return myolo_gemm_taps(x, Cin, wt, y, Cout, M, Cout, Cin, 9, s, bias, scale, shift_c, act, W + 1, (H + 1) * (W + 1), 0, stream);
}
extern "C" int myolo_conv3x3_dgrad(const float* dy, const float* w, float* dx, int n_img, int H, int W, int Cin,
                                   int Cout, myolo_stream stream) {
  // dx[p,ci] = sum_t sum_co dy[p - shift_t, co] * w[t][ci][co]
  MYOLO_CHECK_ARG(n_img > 0 && H > 0 && W > 0);
  int s[9];
  conv3x3_shifts(W, s, true);
  const long long M = (long long)n_img * (H + 1) * (W + 1);
  return myolo_gemm_taps(dy, Cout, w, dx, Cin, M, Cin, Cout, 9, s, nullptr, nullptr, nullptr, MYOLO_ACT_NONE, W + 1, (H + 1) * (W + 1), 0, stream);
}
extern "C" int myolo_conv3x3_wgrad(const float* x, const float* dy, float* dw, int n_img, int H, int W, int Cin,
                                   int Cout, myolo_stream stream) {
  MYOLO_CHECK_ARG(n_img > 0 && H > 0 && W > 0);
  int s[9];
  conv3x3_shifts(W, s, false);
  const long long M = (long long)n_img * (H + 1) * (W + 1);
  MYOLO_CUDA(cudaMemsetAsync(dw, 0, (size_t)9 * Cin * Cout * sizeof(float), myolo::as_stream(stream)));
  return myolo_gemm_taps_wgrad(x, Cin, dy, Cout, dw, M, Cout, Cin, 9, s, 0, stream);
}
