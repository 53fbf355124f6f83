/* myolo_b200.h -- C ABI of libmyolo_sm100.so, the B200 (sm_100a) kernels behind the Mask-YOLO
 * hot path (jianing-sun/Mask-YOLO @ 402dbd9).
 *
 * The reference has no FFI: its hot path is a Keras/TensorFlow graph (myolo/model.py) whose
 * arithmetic runs inside TF ops.  Each entry point below replaces one TF-op call site of that
 * graph (cited per function as myolo/model.py:LINE and the SURVEY.md section 2.3 kernel id).
 * INTEGRATION.md shows the ctypes binding the Python host side uses.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in _host; fp32 unless stated;
 *  - activations are NHWC, described by a myolo_view (strided: innermost C contiguous);
 *  - conv kernels are HWIO ([tap][Cin][Cout]), depthwise [3][3][C], deconv [2][2][Cout][Cin];
 *  - nothing allocates, nothing synchronises; work is enqueued on `stream` (a cudaStream_t);
 *  - return value 0 on success, negative on error; myolo_last_error() describes the failure;
 *  - the library refuses to run on anything but compute capability 10.x (no fallback).
 */
#ifndef MYOLO_B200_H_
#define MYOLO_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef void* myolo_stream;            /* cudaStream_t */

/* Strided NHWC view.  pixel (n,h,w) channel c lives at p[n*sn + h*sh + w*c_stride... ] with the
 * pixel stride fixed to C:  addr = p + n*sn + h*sh + w*C + c.   Dense: sh=W*C, sn=H*W*C.
 * "Padded-flat" (PF) tensors are the same struct with sh=(W+1)*C,
 * sn=(H+1)*(W+1)*C and p pointing at pixel (0,0). */
typedef struct {
  float* p;
  long long sn, sh;
  int n, h, w, c;
} myolo_view;

enum { MYOLO_ACT_NONE = 0, MYOLO_ACT_RELU = 1, MYOLO_ACT_RELU6 = 2,
       /* OR-able flag on any `act` argument: round the stored result to tf32 (operand of a tcgen05 GEMM) */
       MYOLO_ROUND_TF32 = 0x100 };
enum { MYOLO_OK = 0, MYOLO_ERR_ARG = -1, MYOLO_ERR_CUDA = -2, MYOLO_ERR_DEVICE = -3 };

/* ---- library --------------------------------------------------------------------------- */
int myolo_version(void);
const char* myolo_last_error(void);
/* 0 iff device `dev` is sm_100-class; otherwise MYOLO_ERR_DEVICE (no other target is supported). */
int myolo_device_check(int dev);

/* ---- K1: first conv, myolo/model.py:45-50 (ZeroPad(1,1) + Conv 3x3 s2 VALID, 3->Cout) ---- */
int myolo_conv1_fwd(const float* x, const float* w, float* y, int B, int S, int Cout, myolo_stream stream);
int myolo_conv1_wgrad(const float* x, const float* dy, float* dw, int B, int S, int Cout, myolo_stream stream);

/* ---- K2: depthwise 3x3, myolo/model.py:68-77,256-268 (ZeroPad(1,1) + DepthwiseConv VALID) ---- */
/* x is a strided view (dense or padded-flat); y, dy, dx are dense NHWC. */
int myolo_dwconv3x3_fwd(const myolo_view* x, const float* w, float* y, int stride, myolo_stream stream);
/* fused variants.  in_*: BatchNormalization statistics / affine terms of the layer that PRODUCED x; when
 * given, x is that layer's pre-BN output and act(BN(x)) (in_act: MYOLO_ACT_*) is applied while x is read, zero padding
 * after it -- the post-BN activation never exists in HBM (replaces the FusedBatchNorm + Relu6 nodes between a pointwise
 * conv and the next depthwise conv, keras_applications _depthwise_conv_block).  out_mean / out_var: when given, receive
 * the batch mean / biased variance of y over (n, h, w), reduced in the epilogue (ws: BN workspace, zero before and after). */
int myolo_dwconv3x3_fwd_bn(const myolo_view* x, const float* w, float* y, int stride, const float* in_mean,
                           const float* in_var, const float* in_gamma, const float* in_beta, float eps, int in_act,
                           float* out_mean, float* out_var, double* ws, myolo_stream stream);
int myolo_dwconv3x3_bwd_filter_bn(const myolo_view* x, const float* dy, float* dw, int stride, const float* in_mean,
                                  const float* in_var, const float* in_gamma, const float* in_beta, float eps, int in_act,
                                  myolo_stream stream);
int myolo_dwconv3x3_bwd_data(const float* dy, const float* w, float* dx, int B, int H, int W, int C, int stride, myolo_stream stream);
int myolo_dwconv3x3_bwd_filter(const myolo_view* x, const float* dy, float* dw, int stride, myolo_stream stream);

/* ---- K3/K6/K8/K10: tap-GEMM family (pointwise 1x1, 3x3 SAME on padded-flat tiles, deconv) ----
 * C[m,n] = epi( sum_t sum_k A[m + shift[t], k] * Bt[t][n][k] ),  m in [0,M)
 *   A row-major [rows][K], leading dim lda;  Bt = per-tap [N][K] (K contiguous: the tensor-core
 *   "K-major" operand form);  C row-major, leading dim ldc.
 *   epi: (+ bias[n]) -> (* scale[n] + shift_c[n]) -> act;   bias/scale may be NULL.
 *   pf_w1>0: rows are a padded-flat tiling with (W+1)=pf_w1 and block pf_blk; pad rows are NOT
 *   written (they stay zero).  accumulate!=0: C += result (no epilogue affine/act allowed).
 *   Rows m+shift < 0 read as zero on the tcgen05 path (TMA fill); rows past M are read from memory
 *   on both paths (and below 0 on the fp32 path), so padded-flat buffers carry >= W+2 zero guard
 *   rows on both sides.  ntaps <= 32.
 * Precision mode (process-wide): MYOLO_PREC_FP32 = exact fp32 FFMA kernel; MYOLO_PREC_TF32 =
 * tcgen05.mma kind::tf32 with fp32 accumulation in TMEM when the shape qualifies
 * (K % 32 == 0, N % 32 == 0), else the fp32 kernel.
 * Replaces tf Conv2D call sites myolo/model.py:271, 688-713, 848 and keras_applications pw convs. */
enum { MYOLO_PREC_FP32 = 0, MYOLO_PREC_TF32 = 1 };
int myolo_set_precision(int mode);
int myolo_get_precision(void);
int myolo_gemm_taps(const float* A, long long lda, const float* Bt, float* C, long long ldc,
                    long long M, int N, int K, int ntaps, const int* shifts_host,
                    const float* bias, const float* scale, const float* shift_c, int act,
                    int pf_w1, int pf_blk, int accumulate, myolo_stream stream);
/* wgrad: dW[t][k][n] += sum_m A[m + shift[t], k] * D[m, n]   (fp32 atomics; dW pre-zeroed by caller)
 * transpose_out!=0 stores dW as [t][n][k] instead.  TF32 mode uses tcgen05 with both operands
 * MN-major when K % 128 == 0 and N % 32 == 0. */
int myolo_gemm_taps_wgrad(const float* A, long long lda, const float* D, long long ldd, float* dW,
                          long long M, int N, int K, int ntaps, const int* shifts_host,
                          int transpose_out, myolo_stream stream);
/* the two implementations, callable directly (tests, benchmarks) */
int myolo_gemm_taps_ffma(const float* A, long long lda, const float* Bt, float* C, long long ldc,
                         long long M, int N, int K, int ntaps, const int* shifts_host,
                         const float* bias, const float* scale, const float* shift_c, int act,
                         int pf_w1, int pf_blk, int accumulate, myolo_stream stream);
int myolo_gemm_taps_tc(const float* A, long long lda, const float* Bt, float* C, long long ldc,
                       long long M, int N, int K, int ntaps, const int* shifts_host,
                       const float* bias, const float* scale, const float* shift_c, int act,
                       int pf_w1, int pf_blk, int accumulate, myolo_stream stream);
int myolo_gemm_taps_tc_supported(long long lda, long long ldc, long long M, int N, int K, int ntaps, int accumulate);
/* myolo_gemm_taps_tc without bias / scale / activation, plus the batch statistics of its result in the epilogue: mean[n] /
 * var[n] (biased) of C[:, n] over the valid rows (n_valid of them; pad rows of a padded-flat result do not count).  This is
 * the FusedBatchNorm statistics pass of the BatchNormalization that follows a pointwise convolution (keras_applications
 * _depthwise_conv_block; call sites myolo/model.py:68-77, 256-268), folded into the convolution.  ws: the BN workspace
 * (zero before, zero after).  Supported whenever myolo_gemm_taps_tc_supported and N <= 1024. */
int myolo_gemm_taps_tc_stats(const float* A, long long lda, const float* Bt, float* C, long long ldc, long long M,
                             int N, int K, int ntaps, const int* shifts_host, int pf_w1, int pf_blk,
                             float* mean, float* var, double* ws, long long n_valid, myolo_stream stream);
/* persistent multi-tap variant with shared-memory reuse of the A rows across taps (conv_win_tcgen05.cu):
 * N == 256, 2 <= ntaps, every |shift| <= 16 (3x3 conv on padded-flat tiles up to 14 wide); A must carry
 * >= 16 zero guard rows after row M.  myolo_gemm_taps picks it automatically in TF32 mode. */
int myolo_gemm_taps_win(const float* A, long long lda, const float* Bt, float* C, long long ldc,
                        long long M, int N, int K, int ntaps, const int* shifts_host,
                        const float* bias, const float* scale, const float* shift_c, int act,
                        int pf_w1, int pf_blk, int accumulate, myolo_stream stream);
/* plain GEMM over k-segments on the same persistent kernel: C[m][n] = sum_g sum_k A[m + seg_off[g]][k] * Bt[g][n][k]
 * (seg_off >= 0, host array of nseg <= 8 row offsets into ONE matrix A), and -- when mean / var are given -- the batch
 * statistics of C over its M rows in the epilogue (ws: BN workspace, zero before and after).  This is how the pointwise
 * convolutions of the backbone (keras_applications _depthwise_conv_block, call sites myolo/model.py:68-77, 256-268) run in
 * the 3xTF32 modes: A = [tf32 high part | tf32 low part] of the activation, lo_off rows apart, Bt = the [hi | hi | lo]
 * weight triple of myolo_prep_weights (round 2), seg_off = {0, lo_off, 0}.  M >= 4096, N % 128 == 0, K % 32 == 0. */
int myolo_gemm_segs_win(const float* A, long long lda, const float* Bt, float* C, long long ldc, long long M, int N,
                        int K, int nseg, const int* seg_off_host, float* mean, float* var, double* ws, myolo_stream stream);
int myolo_gemm_segs_win_supported(long long lda, long long ldc, long long M, int N, int K, int nseg);
int myolo_gemm_taps_win_supported(long long lda, long long ldc, long long M, int N, int K, int ntaps,
                                  const int* shifts_host, int accumulate);
int myolo_gemm_taps_wgrad_ffma(const float* A, long long lda, const float* D, long long ldd, float* dW,
                               long long M, int N, int K, int ntaps, const int* shifts_host,
                               int transpose_out, myolo_stream stream);
int myolo_gemm_taps_wgrad_tc(const float* A, long long lda, const float* D, long long ldd, float* dW,
                             long long M, int N, int K, int ntaps, const int* shifts_host,
                             int transpose_out, myolo_stream stream);
int myolo_gemm_taps_wgrad_tc_supported(long long lda, long long ldd, long long M, int N, int K, int ntaps);
/* weight staging: out[t][c][r] = f(in[t][r][c]) when transpose!=0, else out = f(in);
 * f rounds to tf32 (round-to-nearest) when round_tf32==1.  round_tf32==2 writes THREE stacked copies
 * [hi | hi | lo] (hi = rna_tf32(v), lo = rna_tf32(v-hi)), each ntaps*rows*cols floats: the B operands of
 * the 3xTF32 tap triple (A_hi,B_hi), (A_lo,B_hi), (A_hi,B_lo).  in != out. */
int myolo_prep_weights(const float* in, float* out, int ntaps, int rows, int cols, int transpose,
                       int round_tf32, myolo_stream stream);
/* every staging job of a step in one launch.  jobs_dev: device array of n_jobs records
 * { const float* in; void* out; int ntaps, rows, cols, transpose, mode, tile_begin, out_ld, out_total; } (48 bytes each)
 * with mode 0 copy / 1 tf32 rounding / 2 3xTF32 split / 3 IEEE half (= myolo_prep_weights with round_tf32 0/1/2 and
 * myolo_prep_weights_h) and tile_begin = running sum of ntaps*ceil(rows/32)*ceil(cols/32) over the jobs before it;
 * total_tiles = that sum over all jobs.  out_ld (0 = dense) is the leading dimension of the staged matrix and out_total
 * (0 = ntaps*rows*cols) the distance between the three copies of mode 2: larger values leave zero padding the job never
 * writes, which is how conv_23's N_BOX*(5+NC) output channels (27, 35, 430) become a tensor-core shape. */
int myolo_prep_weights_batch(const void* jobs_dev, int n_jobs, int total_tiles, myolo_stream stream);
/* dst[r][0..cols) = src[r][0..cols) between matrices of different row pitch (elements): un-pads the tensor-core result of
 * conv_23 (myolo/model.py:271) into the dense [B,G,G,N_BOX*(5+NC)] tensor the decode / loss kernels read, and pads its
 * gradient for the data-gradient GEMM. */
int myolo_copy_cols(const float* src, long long src_ld, float* dst, long long dst_ld, long long rows, int cols,
                    myolo_stream stream);
/* gather (scatter == 0): dst tile i = src tile list[i]; scatter: dst tile list[i] = src tile i, for i < n_list; a tile is
 * tile_bytes contiguous bytes (multiple of 16).  Moves the padded-flat tiles of the POSITIVE rois into / out of compact
 * tensors for the exact sparse backward of the mask head: only rois with a target class carry gradient above
 * myolo_mask_bn1 (myolo_mask_loss_graph, myolo/model.py:718-754, gathers the positive rois before the loss). */
int myolo_copy_tiles(const void* src, void* dst, const int* list, int n_list, long long tile_bytes, int scatter,
                     myolo_stream stream);
/* pointwise / 3x3 named wrappers (SURVEY 8b names).  w = HWIO kernel [t][Cin][Cout]; wt = its
 * per-tap transpose [t][Cout][Cin] (myolo_prep_weights). */
int myolo_pwconv_fwd(const float* x, const float* wt, float* y, long long M, int Cin, int Cout,
                     const float* bias, myolo_stream stream);
int myolo_pwconv_dgrad(const float* dy, const float* w, float* dx, long long M, int Cin, int Cout, myolo_stream stream);
int myolo_pwconv_wgrad(const float* x, const float* dy, float* dw, long long M, int Cin, int Cout, myolo_stream stream);
/* 3x3 SAME conv on a padded-flat tensor: n_img tiles of H x W, rows = n_img*(H+1)*(W+1). x and y point at
 * row 0 of the tiling (guards of >= W+2 zero rows must exist on both sides of x). */
int myolo_conv3x3_fwd(const float* x, const float* wt, float* y, int n_img, int H, int W, int Cin, int Cout,
                      const float* bias, const float* scale, const float* shift_c, int act, myolo_stream stream);
int myolo_conv3x3_dgrad(const float* dy, const float* w, float* dx, int n_img, int H, int W, int Cin, int Cout, myolo_stream stream);
int myolo_conv3x3_wgrad(const float* x, const float* dy, float* dw, int n_img, int H, int W, int Cin, int Cout, myolo_stream stream);

/* ---- K4/K5/K9: batch norm + activation (Keras BatchNormalization eps 1e-3) ----
 * Workspace `ws` of every entry point in this block: MYOLO_BN_WS_DOUBLES doubles (C <= 1024), ZERO before the
 * first call; the reductions finalize and reset their sums and tickets themselves (one launch each). */
#define MYOLO_BN_WS_DOUBLES 4112
/* mean[c], var[c] (biased) over all pixels of x, single pass (sum and sum of squares in fp64). */
int myolo_bn_stats(const myolo_view* x, float* mean, float* var, double* ws, myolo_stream stream);
/* y = act(gamma*(x-mean)*rsqrt(var+eps)+beta) */
int myolo_bn_apply(const myolo_view* x, const myolo_view* y, const float* mean, const float* var,
                   const float* gamma, const float* beta, float eps, int act, myolo_stream stream);
/* same, but the result v is stored as the operand pair of a 3xTF32 GEMM: y_hi = rna_tf32(v),
 * y_lo = rna_tf32(v - y_hi) (A*B ~= Ah*Bh + Al*Bh + Ah*Bl recovers fp32-grade accuracy on tcgen05). */
int myolo_bn_apply_split(const myolo_view* x, const myolo_view* y_hi, const myolo_view* y_lo, const float* mean,
                         const float* var, const float* gamma, const float* beta, float eps, int act,
                         myolo_stream stream);
/* hi = rna_tf32(src), lo = rna_tf32(src - hi), elementwise over equal-shape views. */
int myolo_split_tf32(const myolo_view* src, const myolo_view* hi, const myolo_view* lo, myolo_stream stream);
/* backward of act(BN(x)). train!=0: batch-statistics BN (mean/var are this batch's); else moving stats.
 * dgamma/dbeta are OVERWRITTEN. dx may alias dy. */
int myolo_bn_bwd(const myolo_view* x, const myolo_view* dy, const myolo_view* dx, const float* mean, const float* var,
                 const float* gamma, const float* beta, float eps, int act, int train,
                 float* dgamma, float* dbeta, double* ws, myolo_stream stream);
/* scale = gamma*rsqrt(var+eps), shift = beta - mean*scale: fixed-statistics BN (myolo_mask_bn2..4, model.py:695-708)
 * folded into the producing conv's epilogue through the scale / shift_c arguments of myolo_gemm_taps. */
int myolo_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                  float* scale, float* shift, int C, myolo_stream stream);
/* backward of a = act(BN_fixed_stats(conv + bias)) from the OUTPUT a only (the pre-BN tensor is never
 * stored): dx = dy*act'(a)*gamma*rs (may alias dy), dgamma/dbeta OVERWRITTEN, dbias (nullable) =
 * gamma*rs*dbeta = gradient of the conv bias. */
int myolo_bn_act_bwd_from_output(const myolo_view* a, const myolo_view* dy, const myolo_view* dx, const float* gamma,
                                 const float* beta, const float* var, float eps, int act, float* dgamma, float* dbeta,
                                 float* dbias, double* ws, myolo_stream stream);
/* dgrad GEMM with the backward of the PREVIOUS layer's fixed-statistics BN + activation fused into its epilogue
 * (tcgen05 persistent kernel, N == 256): the GEMM result is d(a); the kernel reads a_out ([M][N], same rows as C),
 * stores C = d(a)*act'(a)*gamma*rs and reduces dbeta / dgamma / dbias -- what myolo_gemm_taps followed by
 * myolo_bn_act_bwd_from_output computes, without the three extra passes over the gradient tensor.
 * ws: the BN workspace (MYOLO_BN_WS_DOUBLES, zero before and after). */
int myolo_gemm_taps_bnbwd(const float* A, long long lda, const float* Bt, float* C, long long ldc, long long M,
                          int N, int K, int ntaps, const int* shifts_host, int pf_w1, int pf_blk,
                          const float* a_out, const float* gamma, const float* beta, const float* var, float eps,
                          int act, float* dgamma, float* dbeta, float* dbias, double* ws, myolo_stream stream);
int myolo_gemm_taps_bnbwd_supported(long long lda, long long ldc, long long M, int N, int K, int ntaps,
                                    const int* shifts_host);
int myolo_bn_epi_finalize(double* sums, const float* gamma, const float* var, float eps, float* dgamma,
                          float* dbeta, float* dbias, int C, myolo_stream stream);
/* Keras moving-average update with TF zero-debias: biased -= (biased-value)*(1-momentum);
 * moving = biased/(1-momentum^step). value = mean, or var*bessel*n/(n-(1+eps)) when is_var. */
int myolo_bn_moving_update(const float* value, float* biased, float* moving, int C, float momentum, int step,
                           int is_var, double n, float eps, myolo_stream stream);
/* the same update for many (layer, statistic) pairs in ONE launch.  items_dev: device array of n_items records
 * { const float* value; float* biased; float* moving; int C; float corr; } (32 bytes each, corr = 1 for means,
 * bessel * n/(n-(1+eps)) for variances); every item shares the step count. */
int myolo_bn_moving_update_batch(const void* items_dev, int n_items, float momentum, int step, myolo_stream stream);
/* out[c] = sum over pixels of x[.,c]  (bias gradients). */
int myolo_colsum(const myolo_view* x, float* out, double* ws, myolo_stream stream);
/* dst (=|+=) src over valid pixels (dense <-> padded-flat moves, gradient joins).
 * accumulate bit 0: dst += src; bit 1: round the stored value to tf32. */
int myolo_view_copy(const myolo_view* src, const myolo_view* dst, int accumulate, myolo_stream stream);

/* ---- K7: PyramidROIAlign, myolo/model.py:385-387 = tf.image.crop_and_resize(feat, boxes, idx, (P,P)) ----
 * boxes[n_roi][4] are consumed as (y1,x1,y2,x2) exactly like the TF op; the reference passes
 * (x1,y1,x2,y2) and so does the host side here.  roi n samples image n / rois_per_img.
 * round_tf32!=0 rounds the pooled values to tf32 (they are the A operand of mask conv1). */
int myolo_roialign_fwd(const myolo_view* feat, const float* boxes, int n_roi, int rois_per_img, int pool,
                       const myolo_view* out, int round_tf32, myolo_stream stream);
int myolo_roialign_bwd(const myolo_view* dout, const float* boxes, int n_roi, int rois_per_img, int pool,
                       const myolo_view* dfeat, myolo_stream stream);

/* ---- half-operand (tcgen05 kind::f16) variants of the mask-head path, myolo/model.py:688-713 ----
 * The mask head is 98.8 % of the FLOPs.  tcgen05 has no fp32 MMA; kind::tf32 keeps 10 explicit mantissa bits, IEEE half
 * keeps the same 10 at TWICE the tensor rate and half the operand bytes.  In the "h16" precision mode the mask-head
 * activations, staged weights and (loss-scaled) gradients are stored as IEEE half, every accumulation stays fp32 in
 * TMEM, and every epilogue (bias, BN, ReLU, sigmoid, column sums) stays fp32.  `const void*` / `void*` arguments
 * below are device pointers to IEEE-half data; half myolo_views use the same struct with p pointing at half data
 * and sn / sh counted in ELEMENTS.  Conversions are round-to-nearest-even, saturating at +-65504. */
/* out[t][c][r] = half(in[t][r][c]) when transpose != 0, else out = half(in) (weight staging, see myolo_prep_weights) */
int myolo_prep_weights_h(const float* in, void* out_half, int ntaps, int rows, int cols, int transpose, myolo_stream stream);
/* myolo_gemm_taps on the persistent CTA-pair kernel with half A / Bt.  Output: EITHER C (fp32) OR Ch (half), the other
 * NULL (eight epilogue warps own one staging tile each).  acc_scale (nullable): DEVICE scalar
 * multiplied into the accumulator before the epilogue (un-scaling of loss-scaled gradients).  N % 256 == 0, K % 64 == 0,
 * ntaps >= 2 with |shift| <= 16, or a plain GEMM with M >= 4096. */
int myolo_gemm_taps_h(const void* A, long long lda, const void* Bt, float* C, long long ldc, void* Ch, long long ldch,
                      long long M, int N, int K, int ntaps, const int* shifts_host, const float* bias,
                      const float* scale, const float* shift_c, int act, int pf_w1, int pf_blk,
                      const float* acc_scale, myolo_stream stream);
int myolo_gemm_taps_h_supported(long long lda, long long M, int N, int K, int ntaps, const int* shifts_host);
/* myolo_deconv_mask_fwd with half a4 [rows][Cmid] and half kd [4*Cmid][Cmid] (masks, y4 stay fp32). */
/* myolo_gemm_taps_h with an fp32 result (bias added, no activation) and the batch statistics of that result in the
 * epilogue: mean[n] / var[n] (biased) over the valid rows (n_valid of them; pad rows of the padded-flat result do not
 * count).  myolo_mask_conv1 + the statistics pass of myolo_mask_bn1, the one mask-head BatchNormalization that follows
 * the learning phase (myolo/model.py:688-690).  pivot (nullable, [N]): subtracted from every value before the sums are
 * taken, e.g. the layer's moving mean, so that the variance does not lose mean^2 / variance digits; the results do not
 * depend on it beyond rounding.  N == 256; ws: the BN workspace (zero before, zero after). */
int myolo_gemm_taps_h_stats(const void* A, long long lda, const void* Bt, float* C, long long ldc, long long M, int N,
                            int K, int ntaps, const int* shifts_host, const float* bias, int pf_w1, int pf_blk,
                            const float* pivot, float* mean, float* var, double* ws, long long n_valid, myolo_stream stream);
/* The same with the result stored as IEEE half (Ch, pitch ldch) and the statistics taken from the half-rounded values,
 * i.e. of exactly the tensor the BatchNormalization that follows reads (default h16 path of myolo_mask_bn1). */
int myolo_gemm_taps_hh_stats(const void* A, long long lda, const void* Bt, void* Ch, long long ldch, long long M, int N,
                             int K, int ntaps, const int* shifts_host, const float* bias, int pf_w1, int pf_blk,
                             const float* pivot, float* mean, float* var, double* ws, long long n_valid, myolo_stream stream);
int myolo_deconv_mask_fwd_h(const void* a4, const void* kd, const float* bd, const float* w1, const float* b1,
                            float* masks, const int* target_ids, float* y4, int n_roi, int H, int W, int Cmid, int NC,
                            myolo_stream stream);
/* myolo_gemm_taps_bnbwd with half A / Bt / a_out; the result goes to EITHER C (fp32) OR Ch (half), [M][N] with pitch
 * ldc, the other NULL.  grad_unscale (nullable): DEVICE scalar applied to dgamma / dbeta / dbias (the incoming
 * gradient carries a loss scale; the stored d(pre-BN) keeps it). */
int myolo_gemm_taps_bnbwd_h(const void* A, long long lda, const void* Bt, float* C, void* Ch, long long ldc,
                            long long M, int N, int K, int ntaps, const int* shifts_host, int pf_w1, int pf_blk,
                            const void* a_out, const float* gamma, const float* beta, const float* var, float eps,
                            int act, float* dgamma, float* dbeta, float* dbias, double* ws,
                            const float* grad_unscale, myolo_stream stream);
int myolo_bn_epi_finalize_s(double* sums, const float* gamma, const float* var, float eps, float* dgamma,
                            float* dbeta, float* dbias, int C, const float* unscale, myolo_stream stream);
/* Backward of a BATCH-statistics BatchNormalization + activation in two launches instead of three passes
 * (myolo_mask_bn1 in the learning phase, myolo/model.py:688-690; keras BatchNormalization training=True):
 * 1. myolo_gemm_taps_bnbwd_sums_h = myolo_gemm_taps_bnbwd_h on the data-gradient GEMM that PRODUCES d(a): its epilogue
 *    stores g1 = gamma * rsqrt(var + eps) * d(a) * act'(a) as half (loss scale kept) and leaves sum(g), sum(g * xhat) in ws;
 *    var is the BATCH variance of the forward pass.  Nothing is finalised: ws is NOT zero on return.
 * 2. myolo_bn_bwd_batch_fix_hh: dgamma / dbeta from those sums (times *grad_unscale), ws zeroed again, and in place over
 *    the view's pixels  g_half <- g_half - gamma*rs*(S0/n + xhat*S1/n),  xhat = (z_half - mean) * rs,  n = pixels of the
 *    view (z_half = the half pre-BN tensor of the forward pass).  Between the two calls g_half may be rearranged (the
 *    sparse backward scatters a compact result into a zeroed full tensor); rows that stayed zero get the mean terms only. */
int myolo_gemm_taps_bnbwd_sums_h(const void* A, long long lda, const void* Bt, void* Ch, long long ldc, long long M,
                                 int N, int K, int ntaps, const int* shifts_host, int pf_w1, int pf_blk,
                                 const void* a_out, const float* gamma, const float* beta, const float* var, float eps,
                                 int act, double* ws, myolo_stream stream);
int myolo_bn_bwd_batch_fix_hh(const myolo_view* z_half, const myolo_view* g_half, const float* mean, const float* var,
                              const float* gamma, float eps, float* dgamma, float* dbeta, double* ws,
                              const float* grad_unscale, myolo_stream stream);
/* myolo_bn_apply from a half pre-BN tensor to a half result. */
int myolo_bn_apply_hh(const myolo_view* x_half, const myolo_view* y_half, const float* mean, const float* var,
                      const float* gamma, const float* beta, float eps, int act, myolo_stream stream);
/* myolo_roialign_fwd with the pooled values stored as half (out_half) and, when out != NULL, also as fp32. */
int myolo_roialign_fwd_h(const myolo_view* feat, const float* boxes, int n_roi, int rois_per_img, int pool,
                         const myolo_view* out, const myolo_view* out_half, myolo_stream stream);
/* myolo_roialign_bwd (CropAndResizeGradImage) on a half, loss-scaled gradient: every value read is multiplied by
 * *in_scale (device scalar, nullable = 1); dfeat stays fp32.  C = 128 or 256. */
int myolo_roialign_bwd_h(const myolo_view* dout_half, const float* boxes, int n_roi, int rois_per_img, int pool,
                         const myolo_view* dfeat, const float* in_scale, myolo_stream stream);
/* myolo_bn_apply with the result stored as half (y_half) and, when y != NULL, as fp32 holding the same rounded values. */
int myolo_bn_apply_h(const myolo_view* x, const myolo_view* y, const myolo_view* y_half, const float* mean,
                     const float* var, const float* gamma, const float* beta, float eps, int act, myolo_stream stream);

/* -- backward of the mask head in the "h16" mode.  Gradients are stored as half multiplied by a power-of-two loss
 * scale S that lives on the DEVICE: gs = {S, 1/S, scratch}, computed per step by myolo_grad_scale from max|dlogit| so
 * that the scaled gradient peaks in [8, 16) (no host synchronisation; S = 1 for an all-zero gradient).  Every
 * parameter gradient is un-scaled where it is reduced (out_scale / grad_unscale = gs + 1), so the flat gradient
 * buffer, Adam and the all-reduce never see S. */
int myolo_grad_scale(const float* g, long long n, float* gs, myolo_stream stream);
/* myolo_mask_out_bwd with dy4 stored as half * (*gscale); dw1 / db1 / dbd are unscaled.  target_ids (nullable):
 * the ids myolo_mask_loss was called with -- rois with id <= 0 have an identically zero dlogit, so their rows are
 * zero-filled without reading it.  prev_ids (nullable, int[n_roi], zero before the first call, owned by the dy4
 * buffer): the ids of the call that last wrote dy4_half; rows that were not positive then are still zero and are not
 * written again (the 2 GB zero fill per step was the whole cost of this kernel).  Updated to target_ids on return. */
int myolo_mask_out_bwd_h(const float* y4, const float* bd, const float* w1, const float* dlogit, void* dy4_half,
                         float* dw1, float* db1, float* dbd, int n_roi, int H, int W, int Cmid, int NC,
                         const float* gscale, const int* target_ids, int* prev_ids, myolo_stream stream);
/* myolo_gemm_taps_wgrad with half A [rows][K] and half D [rows][N] (both MN-major operands of tcgen05 kind::f16):
 * dW[t][k][n] += (*out_scale) * sum_m A[m + shift[t], k] * D[m, n]  (fp32 atomics; out_scale nullable = 1).
 * K % 64 == 0, N % 64 == 0, lda % 8 == 0, ldd % 8 == 0; dW 16-byte aligned unless transpose_out (128-bit vector reductions). */
int myolo_gemm_taps_wgrad_h(const void* A, long long lda, const void* D, long long ldd, float* dW, long long M,
                            int N, int K, int ntaps, const int* shifts_host, int transpose_out,
                            const float* out_scale, myolo_stream stream);
/* Width of the following myolo_gemm_taps_wgrad_h launches: their split-M CTAs are sized for n SMs (32..148, default 148 = one
 * CTA per SM).  A filter-gradient launch issued on its own stream next to the backbone's backward leaves 148 - n SMs to
 * that chain's kernels (the persistent CTAs hold 197 KB of shared memory each, so nothing tensor-core-sized fits beside
 * them).  Process-global, takes effect in issue order like every other entry point; it changes the launch geometry only
 * (how the reduction over rows is split), never which sums are formed; no reference counterpart. */
int myolo_set_wgrad_sms(int n);

int myolo_gemm_taps_wgrad_h_supported(long long lda, long long ldd, long long M, int N, int K, int ntaps);
/* myolo_bn_bwd (fp32 x, fp32 UNSCALED dy) whose dx is stored as half * (*out_scale) in the half view dx_half. */
int myolo_bn_bwd_h(const myolo_view* x, const myolo_view* dy, const myolo_view* dx_half, const float* mean,
                   const float* var, const float* gamma, const float* beta, float eps, int act, int train,
                   float* dgamma, float* dbeta, double* ws, const float* out_scale, myolo_stream stream);

/* myolo_bn_bwd with fp32 x and HALF, loss-scaled dy and dx (dx may alias dy): dx keeps dy's scale, dgamma / dbeta are
 * multiplied by *grad_unscale. */
int myolo_bn_bwd_hh(const myolo_view* x, const myolo_view* dy_half, const myolo_view* dx_half, const float* mean,
                    const float* var, const float* gamma, const float* beta, float eps, int act, int train,
                    float* dgamma, float* dbeta, double* ws, const float* grad_unscale, myolo_stream stream);

/* ---- K12: DecodeYOLOLayer / DetectionsLayer, myolo/model.py:1442-1473, 1493-1538 ---- */
int myolo_yolo_decode(const float* y_pred, const float* anchors, float* boxes, float* detections /*nullable*/,
                      int B, int GH, int GW, int NB, int NC, myolo_stream stream);

/* ---- K13: DetectMaskTargetLayer, myolo/model.py:457-602 (+ norm_boxes_graph 1394-1408) ----
 * gt_boxes are PIXEL (x1,y1,x2,y2) as fed to the model; normalisation happens inside.
 * gt_class_ids [B,M], gt_boxes [B,M,4] (M = TRUE_BOX_BUFFER columns as BatchGenerator pads them),
 * gt_masks [B,S,S,MM] bytes (MM = MAX_GT_INSTANCES channels).
 * Outputs: rois [B,R,4], target_ids [B,R] int32, target_masks [B,R,MH,MW], n_pos [B] int32,
 * roi_src [B,R] int32 (source proposal index, -1 for padding), roi_gt [B,R] int32 (matched GT, -1). */
int myolo_detect_mask_targets(const float* proposals, const int* gt_class_ids, const float* gt_boxes,
                              const unsigned char* gt_masks, int B, int R, int M, int MM, int S, int MH, int MW,
                              float* rois, int* target_ids, float* target_masks, int* n_pos,
                              int* roi_src, int* roi_gt, myolo_stream stream);

/* ---- K10+K11: mask output, myolo/model.py:711-713.  y4 = deconv GEMM output rows [p][(a,b,co)]
 * (padded-flat H x W tiles, pre-bias).  masks[n][2h+a][2w+b][k] = sigmoid(b1[k] + sum_co relu(y4+bd[co]) * w1[co][k]) */
int myolo_mask_out_fwd(const float* y4, const float* bd, const float* w1, const float* b1, float* masks,
                       int n_roi, int H, int W, int Cmid, int NC, myolo_stream stream);
/* K10+K11 fused (tcgen05): deconv GEMM a4 [rows,Cmid] x kd [4*Cmid][Cmid] with the whole mask tail in the
 * epilogue -> masks.  y4 rows are written ONLY for rois with target_ids > 0 (the rows myolo_mask_out_bwd
 * reads); target_ids may be NULL (inference: nothing is written to y4).  Cmid == 256, NC <= 7. */
int myolo_deconv_mask_fwd(const float* a4, const float* kd, const float* bd, const float* w1, const float* b1,
                          float* masks, const int* target_ids, float* y4, int n_roi, int H, int W, int Cmid, int NC,
                          myolo_stream stream);
int myolo_deconv_mask_fwd_supported(int Cmid, int NC);
/* backward: dlogit [n][2H][2W][NC] -> dy4 (same layout as y4), dw1 [Cmid][NC] +=, db1 [NC] +=, dbd [Cmid] += */
int myolo_mask_out_bwd(const float* y4, const float* bd, const float* w1, const float* dlogit,
                       float* dy4, float* dw1, float* db1, float* dbd,
                       int n_roi, int H, int W, int Cmid, int NC, myolo_stream stream);

/* ---- K16: myolo_mask_loss_graph, myolo/model.py:718-754 (Keras binary_crossentropy) ----
 * loss_out[0] = mean BCE over positive rois' class masks (0 if none). dlogit (pre-sigmoid gradient,
 * scaled by loss_weight) is written for every element (zeros off the positive/class entries).
 * ws: 2 doubles. */
int myolo_mask_loss(const float* masks, const float* target_masks, const int* target_ids, int n_roi,
                    int MH, int MW, int NC, float loss_weight, float* loss_out, float* dlogit /*nullable*/,
                    double* ws, myolo_stream stream);

/* ---- K15: yolo_custom_loss, myolo/model.py:86-242 ----
 * scales = {OBJECT, NO_OBJECT, COORD, CLASS}; warmup!=0 selects the warm-up branch (196-207).
 * loss_out[0..4] = total, xy, wh, conf, class. dy_pred (nullable) = d(loss_weight*total)/dy_pred. ws: 8 doubles. */
int myolo_yolo_loss(const float* y_true, const float* y_pred, const float* true_boxes, const float* anchors,
                    const float* class_weights, int B, int GH, int GW, int NB, int NC, int TB,
                    const float* scales_host, int warmup, float loss_weight,
                    float* loss_out, float* dy_pred, double* ws, myolo_stream stream);

/* ---- inference post-processing, myolo/model.py:1290-1304 + 1330-1391 (myolo_utils.py:88-113 NMB, 883-912 unmold_mask) ----
 * per image: detections without area are dropped first (decode_masks, model.py:1367-1375), then the top_k detections by
 * confidence that reach cs_threshold (ties: the higher index first, as np.argsort(scores)[::-1] orders them), NMB
 * suppression exactly as the reference does it (a candidate is dropped when any EARLIER candidate of the SAME class --
 * dropped or not -- has IoU >= nms_threshold with it), then each survivor's class mask pasted by unmold_mask's rule: the
 * box corners truncated with int(), x1/y1 clamped to [0,S], x2/y2 to [1,S], the mask resized (bilinear, pixel centres
 * aligned) into that CLIPPED box, thresholded at 0.5 and written into an S x S byte image.
 * detections [B,R,6] = DetectionsLayer output; masks [B,R,MH,MW,NC] (nullable together with out_masks).
 * out_index [B,top_k] (detection index or -1), out_boxes [B,top_k,4] int32 pixels (x1,y1,x2,y2) = the box the mask was
 * pasted into, out_class / out_score [B,top_k], out_count [B], out_masks [B,top_k,S,S] bytes.  top_k <= 32. */
int myolo_detect_postprocess(const float* detections, const float* masks, int B, int R, int NC, int S, int MH, int MW,
                             int top_k, float cs_threshold, float nms_threshold, int* out_index, int* out_boxes,
                             int* out_class, float* out_score, int* out_count, unsigned char* out_masks,
                             myolo_stream stream);

/* ---- target encoding on the device (host loops of myolo_utils.py:247-271 extract_bboxes and 769-820 BatchGenerator) ----
 * boxes[b][m] = (x1,y1,x2,y2) of mask column m of gt_masks [B,S,S,M] bytes, x2/y2 exclusive, zeros for an empty mask. */
int myolo_extract_bboxes(const unsigned char* gt_masks, int B, int S, int M, int* boxes, myolo_stream stream);
/* yolo_target [B,G,G,NB,5+NC] and true_boxes [B,TB,4] (both zero-filled first) from padded int32 gt arrays
 * [B,M] / [B,M,4] (rows with class id 0 are padding): centre and size in grid units, the cell holding the centre,
 * the anchor with the best IoU against (0,0,w,h) (first maximum wins), later instances overwrite earlier ones. */
int myolo_encode_yolo_targets(const int* gt_class_ids, const int* gt_boxes, int B, int M, int S, int G, int NB, int NC,
                              int TB, const float* anchors, float* yolo_target, float* true_boxes, myolo_stream stream);

/* ---- the Shapes workload on the device: example/shapes/dataset_shapes.py:80-135 (load_image, load_mask,
 * draw_shape = cv2.rectangle / cv2.circle / cv2.fillPoly, later shapes occlude earlier ones) + myolo_utils.py:274-366
 * load_image_gt (instances with an empty visible mask are dropped, order kept) + 247-271 extract_bboxes + the padding and
 * `image / 255.` of BatchGenerator.__getitem__ (821-851), for a whole batch.
 * specs [B, 4 + 8*MS] int32 as the host generator draws them: bg r, g, b, n_shapes, then per shape
 *   type (1 square, 2 circle, 3 triangle = class id), r, g, b, x, y, s, 0.
 * ws: B*MS*(2*S+1) ints (row extents per shape + compaction slots), 8-byte aligned.
 * Outputs: image_f32 [B,S,S,3] = float32(uint8 / 255.) (nullable), image_u8 [B,S,S,3] (nullable), gt_masks [B,S,S,M] bytes,
 * gt_class_ids [B,TB], gt_boxes [B,TB,4] int32 (x1,y1,x2,y2; x2/y2 exclusive), gt_boxes_f the same as float (nullable);
 * all zero padded.  S % 16 == 0, MS <= 8, MS <= M <= 128, MS <= TB.  Pixel-exact with OpenCV 4.13 (the row-extent
 * functions are compared with cv2 on the CPU for arbitrary centres / sizes / image shapes, tests/test_shapes_raster.py). */
int myolo_shapes_raster(const int* specs, int B, int S, int MS, int M, int TB, int* ws, float* image_f32,
                        unsigned char* image_u8, unsigned char* gt_masks, int* gt_class_ids, int* gt_boxes,
                        float* gt_boxes_f, myolo_stream stream);

/* ---- VIA polygon annotations on the device: RiceDataset.load_mask, example/rice/rice_dataset.py:135-159
 * (`rr, cc = skimage.draw.polygon(all_points_y, all_points_x); mask[rr, cc, i] = 1` per instance) for all instances of
 * one image in one launch.  verts_y / verts_x: the concatenated float64 vertex rows / columns of the instances,
 * offsets [n_inst + 1] int32 (instance i owns vertices offsets[i] .. offsets[i+1]-1), all device pointers.
 * masks [H, W, M] bytes (the reference's layout), channels >= n_inst zero; n_inst <= M <= 128, masks 4-byte aligned.
 * ws: 4*M ints, 16-byte aligned (the instances' candidate boxes, written by the first of the two launches).
 * Inclusion rule: skimage's float64 crossing-number test, restated in csrc/polygon_pip.h (scikit-image is absent here:
 * parity pinned to the oracle's restatement, not to the library).  Pixels outside the image are clipped; the reference
 * raises IndexError there, which the Python wrapper (myolo.rice) reproduces before the launch. */
int myolo_polygon_masks(const double* verts_y, const double* verts_x, const int* offsets, int n_inst, int H, int W, int M,
                        int* ws, unsigned char* masks, myolo_stream stream);

/* ---- the data-parallel exchange step: ONE sum all-reduce of the flat fp32 gradient buffer per step, issued
 * as two slices (mask-head tail first, overlapping the backbone backward).  The reference has no distributed code; this
 * is the NCCL communicator behind the C ABI (bound at run time to the libnccl.so.2 PyTorch ships and has loaded).
 * unique_id: rank 0 fills 128 bytes and hands them to every rank (any side channel, e.g. a torch.distributed broadcast);
 * init: collective over all ranks, uses the calling thread's current CUDA device, returns the handle in *comm_out;
 * run: in-place sum over ranks of buf[0..n) on `stream` (asynchronous, like every other entry point);
 * destroy: releases the communicator. */
int myolo_allreduce_unique_id(char* id128);
int myolo_allreduce_init(const char* id128, int rank, int world, void** comm_out);
int myolo_allreduce_run(void* comm, float* buf, long long n, myolo_stream stream);
int myolo_allreduce_destroy(void* comm);

/* ---- K17: Keras Adam, myolo/model.py:1071-1075 ----  lr_t = lr*sqrt(1-b2^t)/(1-b1^t) computed by caller */
int myolo_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr_t,
                    float b1, float b2, float eps, float grad_scale, myolo_stream stream);
/* the same update restricted to the variables Keras would hand the optimizer: `trainable[i]` is 1 for elements of
 * trainable variables and 0 for frozen ones (MaskYOLO.set_trainable, model.py:1120-1155; yolo_trainable=False,
 * model.py:854-868); a frozen element keeps p, m and v bit-for-bit. */
int myolo_adam_step_masked(float* p, const float* g, float* m, float* v, const float* trainable, long long n,
                           float lr_t, float b1, float b2, float eps, float grad_scale, myolo_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* MYOLO_B200_H_ */
