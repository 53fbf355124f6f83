// mask_out.cu -- tail of the mask head (K10 epilogue + K11) and the Keras Adam update (K17).
//
//   myolo/model.py:711  Conv2DTranspose(256, 2x2, stride 2) + bias + ReLU
//   myolo/model.py:713  Conv2D(NC, 1x1) + bias + sigmoid
//
// The transposed convolution itself is a GEMM  y4[p][(a,b,co)] = sum_ci a4[p][ci] * Kd[a][b][co][ci]
// (tap-GEMM family, gemm_*.cu); this file applies bias+ReLU, contracts the 256 deconv channels
// against the 1x1 kernel and writes the pixel-shuffled sigmoid masks, so the 28x28x256 activation
// is never materialised.  HBM-bound: one warp per output pixel reads its 1 KB channel vector with
// two 128-bit loads per lane and warp-reduces NC partial sums.
#include "common.cuh"

namespace myolo {

constexpr int kMaxCmid = 256;  // channels per (a,b) group: 8 per lane

__device__ __forceinline__ int pf_row(int n, int h, int w, int H, int W) {
  return (n * (H + 1) + h + 1) * (W + 1) + w + 1;
}

// grid-stride over items = n_roi*H*W*4 (pixel p, sub-position ab); one warp per item.
__global__ void __launch_bounds__(256)
mask_out_fwd_kernel(const float* __restrict__ y4, const float* __restrict__ bd, const float* __restrict__ w1,
                    const float* __restrict__ b1, float* __restrict__ masks, int n_roi, int H, int W, int Cmid,
                    int NC) {
  extern __shared__ __align__(16) float sm[];  // w1 TRANSPOSED [NC][Cmid] (conflict-free float4 reads), b1 [NC]
  for (int i = threadIdx.x; i < Cmid * NC; i += blockDim.x) sm[(i % NC) * Cmid + i / NC] = w1[i];
  for (int i = threadIdx.x; i < NC; i += blockDim.x) sm[Cmid * NC + i] = b1[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long items = (long long)n_roi * H * W * 4;
  const int nq = Cmid >> 7;  // float4 loads per lane (Cmid multiple of 128)
  float4 bq[2];
#pragma unroll
  for (int j = 0; j < 2; ++j)
    bq[j] = (j < nq) ? __ldg(reinterpret_cast<const float4*>(bd) + j * 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long it = warp; it < items; it += nwarps) {
    const int ab = (int)(it & 3);
    const unsigned t = (unsigned)(it >> 2);      // pixel index < 2^31 (checked by the host side): 32-bit div/mod
    const unsigned t2 = t / (unsigned)W;
    const int w = (int)(t - t2 * (unsigned)W);
    const int n = (int)(t2 / (unsigned)H);
    const int h = (int)(t2 - (unsigned)n * (unsigned)H);
    const size_t row = (size_t)pf_row(n, h, w, H, W);
    const float4* src = reinterpret_cast<const float4*>(y4 + row * (size_t)(4 * Cmid) + (size_t)ab * Cmid);
    float v[8];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j < nq) q = ld_stream(src + j * 32 + lane);
      v[4 * j + 0] = fmaxf(q.x + bq[j].x, 0.f);
      v[4 * j + 1] = fmaxf(q.y + bq[j].y, 0.f);
      v[4 * j + 2] = fmaxf(q.z + bq[j].z, 0.f);
      v[4 * j + 3] = fmaxf(q.w + bq[j].w, 0.f);
    }
    const int a = ab >> 1, b = ab & 1;
    float* out = masks + ((((size_t)n * 2 * H + 2 * h + a) * 2 * W) + 2 * w + b) * NC;
    for (int k = 0; k < NC; ++k) {
      float s = 0.f;
      const float4* wk = reinterpret_cast<const float4*>(sm + (size_t)k * Cmid);
#pragma unroll
      for (int j = 0; j < 2; ++j)
        if (j < nq) {
          const float4 ww = wk[j * 32 + lane];
          s = fmaf(v[4 * j + 0], ww.x, s);
          s = fmaf(v[4 * j + 1], ww.y, s);
          s = fmaf(v[4 * j + 2], ww.z, s);
          s = fmaf(v[4 * j + 3], ww.w, s);
        }
      s = warp_sum(s);
      if (lane == (k & 31)) out[k] = 1.f / (1.f + expf(-(s + sm[Cmid * NC + k])));
    }
  }
}

// backward of the same tail.  dlogit is zero except on the class channel of positive ROIs, so the
// per-pixel work is a ballot over the NC gradients and a loop over the (usually one) non-zero class.
__device__ __forceinline__ uint32_t pack_half2_sat(float lo, float hi) {
  uint16_t a, b;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(a) : "f"(lo));
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(b) : "f"(hi));
  return (uint32_t)a | ((uint32_t)b << 16);
}

// HALF: dy4 is stored as IEEE half, multiplied by the loss scale *gscale (the parameter gradients dw1 / db1 / dbd
// are reduced from the unscaled values).
template <bool HALF>
__global__ void __launch_bounds__(256)
mask_out_bwd_kernel(const float* __restrict__ y4, const float* __restrict__ bd, const float* __restrict__ w1,
                    const float* __restrict__ dlogit, float* __restrict__ dy4, float* __restrict__ dw1,
                    float* __restrict__ db1, float* __restrict__ dbd, int n_roi, int H, int W, int Cmid, int NC,
                    const float* __restrict__ gscale, const int* __restrict__ ids, const int* __restrict__ prev_ids) {
  extern __shared__ float sm[];  // w1 [Cmid][NC] | acc_w1 [Cmid][NC] | acc_b1 [NC]
  float* s_w1 = sm;
  float* a_w1 = sm + Cmid * NC;
  float* a_b1 = a_w1 + Cmid * NC;
  for (int i = threadIdx.x; i < Cmid * NC; i += blockDim.x) {  // both TRANSPOSED: [NC][Cmid]
    s_w1[(i % NC) * Cmid + i / NC] = w1[i];
    a_w1[i] = 0.f;
  }
  for (int i = threadIdx.x; i < NC; i += blockDim.x) a_b1[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long items = (long long)n_roi * H * W * 4;
  const int nq = Cmid >> 7;
  float4 bq[2];
#pragma unroll
  for (int j = 0; j < 2; ++j)
    bq[j] = (j < nq) ? __ldg(reinterpret_cast<const float4*>(bd) + j * 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
  float dbd_acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) dbd_acc[e] = 0.f;
  bool any_local = false;
  const float gs = (HALF && gscale) ? __ldg(gscale) : 1.f;
  // a warp takes 32 consecutive items (pixel, sub-position) of ONE roi at a time, so the roi-level decisions below
  // cost one lookup per 32 items (as a per-item test their dependent loads were the whole run time of the kernel
  // once the zero fill was gone)
  const int per_roi = H * W * 4;
  const int chunks = (per_roi + 31) / 32;
  const long long tasks = (long long)n_roi * chunks;
  (void)items;
  for (long long task = warp; task < tasks; task += nwarps) {
   const int n = (int)(task / chunks);
   const int i0 = (int)(task - (long long)n * chunks) * 32;
   // not a positive roi: its dlogit is identically zero (myolo_mask_loss writes it so) -> a pure zero fill, without
   // the dependent load of the gradient.  With prev_ids (the ids of the call that last wrote this dy4 buffer) rows
   // that were not positive then are zero already: nothing to do.
   const bool zero_only = HALF && ids && __ldg(ids + n) <= 0;
   if (zero_only && prev_ids && __ldg(prev_ids + n) <= 0) continue;
   const int i1 = min(per_roi, i0 + 32);
   for (int it = i0; it < i1; ++it) {
    const int ab = it & 3;
    const int pix = it >> 2;
    const int h = pix / W, w = pix - h * W;
    const size_t row = (size_t)pf_row(n, h, w, H, W);
    const size_t off = row * (size_t)(4 * Cmid) + (size_t)ab * Cmid;
    if (zero_only) {
#pragma unroll
      for (int j = 0; j < 2; ++j)
        if (j < nq) reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(dy4) + off)[j * 32 + lane] = make_uint2(0u, 0u);
      continue;
    }
    const int a = ab >> 1, b = ab & 1;
    const float* g = dlogit + ((((size_t)n * 2 * H + 2 * h + a) * 2 * W) + 2 * w + b) * NC;
    float d[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) d[e] = 0.f;
    float hv[8];
    bool loaded = false;
    for (int kb = 0; kb < NC; kb += 32) {
      const float gk = (kb + lane < NC) ? __ldg(g + kb + lane) : 0.f;
      unsigned nz = __ballot_sync(0xffffffffu, gk != 0.f);
      if (nz && !loaded) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
          if (j < nq) q = ld_stream(reinterpret_cast<const float4*>(y4 + off) + j * 32 + lane);
          hv[4 * j + 0] = fmaxf(q.x + bq[j].x, 0.f);
          hv[4 * j + 1] = fmaxf(q.y + bq[j].y, 0.f);
          hv[4 * j + 2] = fmaxf(q.z + bq[j].z, 0.f);
          hv[4 * j + 3] = fmaxf(q.w + bq[j].w, 0.f);
        }
        loaded = true;
      }
      while (nz) {
        const int src = __ffs(nz) - 1;
        nz &= nz - 1;
        const int k = kb + src;
        const float gv = __shfl_sync(0xffffffffu, gk, src);
        if (lane == 0) atomicAdd(a_b1 + k, gv);
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (j < nq) {
              const int co = (j * 32 + lane) * 4 + e;
              d[4 * j + e] = fmaf(gv, s_w1[k * Cmid + co], d[4 * j + e]);
              atomicAdd(a_w1 + k * Cmid + co, hv[4 * j + e] * gv);
            }
          }
      }
    }
    if (loaded) {
      any_local = true;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        d[e] = hv[e] > 0.f ? d[e] : 0.f;
        dbd_acc[e] += d[e];
      }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j)
      if (j < nq) {
        if (HALF)
          reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(dy4) + off)[j * 32 + lane] =
              make_uint2(pack_half2_sat(d[4 * j] * gs, d[4 * j + 1] * gs), pack_half2_sat(d[4 * j + 2] * gs, d[4 * j + 3] * gs));
        else
          reinterpret_cast<float4*>(dy4 + off)[j * 32 + lane] = make_float4(d[4 * j], d[4 * j + 1], d[4 * j + 2], d[4 * j + 3]);
      }
   }
  }
  if (any_local) {
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (j < nq && dbd_acc[4 * j + e] != 0.f) atomicAdd(dbd + (j * 32 + lane) * 4 + e, dbd_acc[4 * j + e]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Cmid * NC; i += blockDim.x) {
    const float v = a_w1[(i % NC) * Cmid + i / NC];
    if (v != 0.f) atomicAdd(dw1 + i, v);
  }
  for (int i = threadIdx.x; i < NC; i += blockDim.x)
    if (a_b1[i] != 0.f) atomicAdd(db1 + i, a_b1[i]);
}

// Keras Adam (optimizers.py, Keras 2.x): m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ;
// p -= lr_t * m / (sqrt(v) + eps), lr_t = lr sqrt(1-b2^t)/(1-b1^t) folded by the caller.
// MASKED: `mask` (1 = trainable, 0 = frozen) is read per element; a frozen element keeps p, m and v untouched (Keras
// leaves non-trainable weights out of the optimizer's update list altogether, model.py:1120-1155).
template <bool MASKED>
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            const float* __restrict__ mask, long long n, float lr_t, float b1, float b2, float eps, float gs) {
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* P = &pp.x;
    const float* G = &gg.x;
    float* M = &mm.x;
    float* V = &vv.x;
    float4 kk = make_float4(1.f, 1.f, 1.f, 1.f);
    if (MASKED) {
      kk = reinterpret_cast<const float4*>(mask)[i];
      if (kk.x == 0.f && kk.y == 0.f && kk.z == 0.f && kk.w == 0.f) continue;
    }
    const float* Kp = &kk.x;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (MASKED && Kp[e] == 0.f) continue;
      const float ge = G[e] * gs;
      M[e] = b1 * M[e] + (1.f - b1) * ge;
      V[e] = b2 * V[e] + (1.f - b2) * ge * ge;
      P[e] = P[e] - lr_t * M[e] / (sqrtf(V[e]) + eps);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = (n4 << 2) + threadIdx.x;
    if (MASKED && mask[i] == 0.f) return;
    const float ge = g[i] * gs;
    const float me = b1 * m[i] + (1.f - b1) * ge;
    const float ve = b2 * v[i] + (1.f - b2) * ge * ge;
    m[i] = me;
    v[i] = ve;
    p[i] = p[i] - lr_t * me / (sqrtf(ve) + eps);
  }
}

}  // namespace myolo

using namespace myolo;

extern "C" int myolo_mask_out_fwd(const float* y4, const float* bd, const float* w1, const float* b1, float* masks,
                                  int n_roi, int H, int W, int Cmid, int NC, myolo_stream stream) {
  MYOLO_CHECK_ARG(y4 && bd && w1 && b1 && masks && n_roi > 0 && H > 0 && W > 0 && NC > 0);
  MYOLO_CHECK_ARG(Cmid > 0 && (Cmid % 128) == 0 && Cmid <= kMaxCmid && (long long)n_roi * H * W < (1LL << 31));
  const size_t smem = (size_t)(Cmid * NC + NC) * sizeof(float);
  MYOLO_CHECK_ARG(smem <= 200 * 1024);
  if (smem > 48 * 1024) MYOLO_CUDA(cudaFuncSetAttribute(mask_out_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long items = (long long)n_roi * H * W * 4;
  const int blocks = (int)max(1LL, min(ceil_div(items, 8), (long long)kNumSMs * 8));
  mask_out_fwd_kernel<<<blocks, 256, smem, as_stream(stream)>>>(y4, bd, w1, b1, masks, n_roi, H, W, Cmid, NC);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

extern "C" int myolo_mask_out_bwd(const float* y4, const float* bd, const float* w1, const float* dlogit, float* dy4,
                                  float* dw1, float* db1, float* dbd, int n_roi, int H, int W, int Cmid, int NC,
                                  myolo_stream stream) {
  MYOLO_CHECK_ARG(y4 && bd && w1 && dlogit && dy4 && dw1 && db1 && dbd && n_roi > 0 && H > 0 && W > 0 && NC > 0);
  MYOLO_CHECK_ARG(Cmid > 0 && (Cmid % 128) == 0 && Cmid <= kMaxCmid && (long long)n_roi * H * W < (1LL << 31));
  const size_t smem = (size_t)(2 * Cmid * NC + NC) * sizeof(float);
  MYOLO_CHECK_ARG(smem <= 200 * 1024);
  if (smem > 48 * 1024) MYOLO_CUDA(cudaFuncSetAttribute(mask_out_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long items = (long long)n_roi * H * W * 4;
  const int per_sm = smem > 64 * 1024 ? 1 : 4;
  const int blocks = (int)max(1LL, min(ceil_div(items, 8), (long long)kNumSMs * per_sm));
  mask_out_bwd_kernel<false><<<blocks, 256, smem, as_stream(stream)>>>(y4, bd, w1, dlogit, dy4, dw1, db1, dbd, n_roi, H, W, Cmid, NC, nullptr, nullptr, nullptr);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

extern "C" int myolo_mask_out_bwd_h(const float* y4, const float* bd, const float* w1, const float* dlogit, void* dy4_half,
                                    float* dw1, float* db1, float* dbd, int n_roi, int H, int W, int Cmid, int NC,
                                    const float* gscale, const int* target_ids, int* prev_ids, myolo_stream stream) {
  MYOLO_CHECK_ARG(y4 && bd && w1 && dlogit && dy4_half && dw1 && db1 && dbd && n_roi > 0 && H > 0 && W > 0 && NC > 0);
  MYOLO_CHECK_ARG(Cmid > 0 && (Cmid % 128) == 0 && Cmid <= kMaxCmid && ((uintptr_t)dy4_half & 7) == 0 &&
                  (long long)n_roi * H * W < (1LL << 31));
  const size_t smem = (size_t)(2 * Cmid * NC + NC) * sizeof(float);
  MYOLO_CHECK_ARG(smem <= 200 * 1024);
  if (smem > 48 * 1024) MYOLO_CUDA(cudaFuncSetAttribute(mask_out_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long items = (long long)n_roi * H * W * 4;
  const int per_sm = smem > 64 * 1024 ? 1 : 4;
  const int blocks = (int)max(1LL, min(ceil_div(items, 8), (long long)kNumSMs * per_sm));
  mask_out_bwd_kernel<true><<<blocks, 256, smem, as_stream(stream)>>>(y4, bd, w1, dlogit, reinterpret_cast<float*>(dy4_half), dw1,
                                                                      db1, dbd, n_roi, H, W, Cmid, NC, gscale, target_ids,
                                                                      target_ids ? prev_ids : nullptr);
  MYOLO_CHECK_LAUNCH();
  if (target_ids && prev_ids)   // remember which rows of this buffer are non-zero now
    MYOLO_CUDA(cudaMemcpyAsync(prev_ids, target_ids, (size_t)n_roi * sizeof(int), cudaMemcpyDeviceToDevice, as_stream(stream)));
  return MYOLO_OK;
}

// ---- loss scale of the half-precision backward pass -------------------------------------------------------------
// gs[0] = S = 2^(4 - e) with max|g| in [2^(e-1), 2^e)  (the scaled gradient peaks in [8, 16): 12 binades of headroom
// below the half maximum for growth through the five backward GEMMs, 18 binades of full precision below the peak),
// gs[1] = 1/S, gs[2] = scratch (max bits, zero between calls).  S = 1 when the gradient is identically zero.
namespace myolo {
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ g, long long n, unsigned* __restrict__ out) {
  float m = 0.f;
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(g)[i];
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) m = fmaxf(m, fabsf(g[(n4 << 2) + threadIdx.x]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));   // non-negative floats order like their bits
}
__global__ void grad_scale_finalize_kernel(float* __restrict__ gs) {
  const float m = __uint_as_float(reinterpret_cast<unsigned*>(gs)[2]);
  float S = 1.f;
  if (m > 0.f && m < 3.0e38f) {
    int e;
    frexpf(m, &e);                    // m = f * 2^e, f in [0.5, 1)
    e = 4 - e;
    e = e < -60 ? -60 : (e > 60 ? 60 : e);
    S = ldexpf(1.f, e);
  }
  gs[0] = S;
  gs[1] = 1.f / S;
  reinterpret_cast<unsigned*>(gs)[2] = 0u;
}
}  // namespace myolo

extern "C" int myolo_grad_scale(const float* g, long long n, float* gs, myolo_stream stream) {
  MYOLO_CHECK_ARG(g && gs && n > 0 && ((uintptr_t)g & 15) == 0);
  const int blocks = (int)max(1LL, min(ceil_div(n / 4 + 1, 256), (long long)kNumSMs * 8));
  absmax_kernel<<<blocks, 256, 0, as_stream(stream)>>>(g, n, reinterpret_cast<unsigned*>(gs) + 2);
  grad_scale_finalize_kernel<<<1, 1, 0, as_stream(stream)>>>(gs);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

extern "C" int myolo_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr_t, float b1,
                               float b2, float eps, float grad_scale, myolo_stream stream) {
  MYOLO_CHECK_ARG(p && g && m && v && n > 0);
  MYOLO_CHECK_ARG((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0);
  const int blocks = (int)max(1LL, min(ceil_div(n / 4 + 1, 256), (long long)kNumSMs * 8));
  adam_kernel<false><<<blocks, 256, 0, as_stream(stream)>>>(p, g, m, v, nullptr, n, lr_t, b1, b2, eps, grad_scale);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

extern "C" int myolo_adam_step_masked(float* p, const float* g, float* m, float* v, const float* trainable, long long n,
                                      float lr_t, float b1, float b2, float eps, float grad_scale, myolo_stream stream) {
  MYOLO_CHECK_ARG(p && g && m && v && trainable && n > 0);
  MYOLO_CHECK_ARG((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v | (uintptr_t)trainable) & 15) == 0);
  const int blocks = (int)max(1LL, min(ceil_div(n / 4 + 1, 256), (long long)kNumSMs * 8));
  adam_kernel<true><<<blocks, 256, 0, as_stream(stream)>>>(p, g, m, v, trainable, n, lr_t, b1, b2, eps, grad_scale);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}
