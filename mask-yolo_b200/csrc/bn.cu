// bn.cu -- BatchNormalization (+ fused activation) forward/backward, per-channel reductions and
// strided-view moves.  All HBM-bound: float4 channel vectors, 128 B per pixel per 8-thread group,
// per-block smem reduction then ONE fp64 atomic per (block, channel) so the batch statistics are
// reproducible to fp32 rounding.  Replaces TF FusedBatchNorm / FusedBatchNormGrad (K4, K9) and the
// Relu / Relu6 nodes (K5); Keras semantics restated in SURVEY.md Q7.
#include "common.cuh"

namespace myolo {

// Division by a runtime constant as multiply-high + add + shift (Granlund-Montgomery round-up form;
// exact for dividends below 2^31).  The 64-bit div/mod chains these replaced made every elementwise
// BN kernel ALU-bound at ~3 TB/s.
struct FastDiv {
  uint32_t mul, shr, d;
};
static inline FastDiv make_fd(uint32_t d) {
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;
  const uint64_t m = ((1ull << 32) * ((1ull << l) - d)) / d + 1;
  return FastDiv{(uint32_t)m, l, d};
}
__device__ __forceinline__ uint32_t fd_div(uint32_t n, const FastDiv& f) { return (__umulhi(f.mul, n) + n) >> f.shr; }

struct V {  // device copy of myolo_view + precomputed divisors
  float* p;
  long long sn, sh;
  int n, h, w, c;
  FastDiv fw, fh, fc4;
  int dense;  // pixels are contiguous: offset = pixel * c
};
static inline V to_v(const myolo_view* v) {
  const int dense = (v->sh == (long long)v->w * v->c) && (v->sn == (long long)v->h * v->sh);
  return V{v->p, v->sn, v->sh, v->n, v->h, v->w, v->c, make_fd((uint32_t)v->w), make_fd((uint32_t)v->h),
           make_fd((uint32_t)(v->c / 4 > 0 ? v->c / 4 : 1)), dense};
}

__device__ __forceinline__ size_t pix_off(const V& v, long long pix) {
  if (v.dense) return (size_t)pix * (size_t)v.c;
  const uint32_t p = (uint32_t)pix;
  const uint32_t t = fd_div(p, v.fw);
  const uint32_t w = p - t * (uint32_t)v.w;
  const uint32_t n = fd_div(t, v.fh);
  const uint32_t h = t - n * (uint32_t)v.h;
  return (size_t)((long long)n * v.sn + (long long)h * v.sh + (long long)w * v.c);
}

__device__ __forceinline__ float act_grad_mask(float y, int act) {
  const int a = act & 0xff;
  if (a == MYOLO_ACT_RELU) return y > 0.f ? 1.f : 0.f;
  if (a == MYOLO_ACT_RELU6) return (y > 0.f && y < 6.f) ? 1.f : 0.f;
  return 1.f;
}

// MODE 0: s0 += x                 MODE 1: s0 += (x-mean)^2
// MODE 2: g = dy*act'(y); s0 += g; s1 += g*xhat
// MODE 3: s0 += x; s1 += x^2   (single-pass batch statistics)
//
// Every reduction ends with the "last block finalizes" pattern: blocks add their partial sums to the fp64
// workspace with atomics, take a ticket, and the block that draws the last ticket of its channel group
// turns the sums into the result (mean/var, dgamma/dbeta + the dx coefficient table, plain column sums)
// and ZEROES sums and ticket again.  One launch instead of memset + reduce + finalize (x2 for the
// statistics); the workspace must be zero before the first call and is zero after every call.
struct Fin {
  float* o0;          // MODE 0: column sums | MODE 2: dbeta | MODE 3: mean
  float* o1;          //                       MODE 2: dgamma | MODE 3: var
  float* coef;        // MODE 2: [rs | gamma*rs | mean(g) | mean(g*xhat)] x C   (nullable)
  int* ticket;        // one counter per 32-channel group
  double inv_count;   // 1 / pixels
  int train;          // MODE 2: batch-statistics BN (mean terms) or fixed statistics
  int C;
  const float* unscale;   // MODE 2 with a loss-scaled half dy: device scalar applied to dbeta / dgamma only (the
                          // mean terms of the coefficient table stay in scaled units, like the dx they produce)
};

__device__ __forceinline__ float4 load_half4(const float* base, size_t off) {   // 4 IEEE halves at element offset `off`
  const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(base) + off);
  float4 r;
  asm("cvt.f32.f16 %0, %1;" : "=f"(r.x) : "h"((uint16_t)(u.x & 0xffffu)));
  asm("cvt.f32.f16 %0, %1;" : "=f"(r.y) : "h"((uint16_t)(u.x >> 16)));
  asm("cvt.f32.f16 %0, %1;" : "=f"(r.z) : "h"((uint16_t)(u.y & 0xffffu)));
  asm("cvt.f32.f16 %0, %1;" : "=f"(r.w) : "h"((uint16_t)(u.y >> 16)));
  return r;
}

// DYH: dy is an IEEE-half view (MODE 2 only)
template <int MODE, bool DYH = false>
__global__ void __launch_bounds__(256)
colreduce_kernel(V x, V dy, const float* __restrict__ mean, const float* __restrict__ var,
                 const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int act,
                 double* __restrict__ out0, double* __restrict__ out1, long long chunk, Fin fin) {
  pdl_entry();
  __shared__ float red[2][32][33];
  __shared__ int s_last;
  const int tid = threadIdx.x;
  const int cq = tid & 7, pg = tid >> 3;
  const int c0 = blockIdx.y * 32 + cq * 4;
  const long long total = (long long)x.n * x.h * x.w;
  const long long p0 = blockIdx.x * chunk, p1 = min(total, p0 + chunk);
  float4 mu = make_float4(0.f, 0.f, 0.f, 0.f), rs = mu, ga = mu, be = mu;
  if (MODE >= 1) mu = *reinterpret_cast<const float4*>(mean + c0);
  if (MODE == 2) {
    const float4 vv = *reinterpret_cast<const float4*>(var + c0);
    rs = make_float4(1.f / sqrtf(vv.x + eps), 1.f / sqrtf(vv.y + eps), 1.f / sqrtf(vv.z + eps), 1.f / sqrtf(vv.w + eps));
    ga = *reinterpret_cast<const float4*>(gamma + c0);
    be = *reinterpret_cast<const float4*>(beta + c0);
  }
  float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
  for (long long p = p0 + pg; p < p1; p += 32) {
    const float4 v = *reinterpret_cast<const float4*>(x.p + pix_off(x, p) + c0);
    if (MODE == 0) {
      s0.x += v.x; s0.y += v.y; s0.z += v.z; s0.w += v.w;
    } else if (MODE == 1) {
      const float a = v.x - mu.x, b = v.y - mu.y, c = v.z - mu.z, d = v.w - mu.w;
      s0.x = fmaf(a, a, s0.x); s0.y = fmaf(b, b, s0.y); s0.z = fmaf(c, c, s0.z); s0.w = fmaf(d, d, s0.w);
    } else if (MODE == 3) {
      s0.x += v.x; s0.y += v.y; s0.z += v.z; s0.w += v.w;
      s1.x = fmaf(v.x, v.x, s1.x); s1.y = fmaf(v.y, v.y, s1.y); s1.z = fmaf(v.z, v.z, s1.z); s1.w = fmaf(v.w, v.w, s1.w);
    } else {
      const float4 g = DYH ? load_half4(dy.p, pix_off(dy, p) + c0) : *reinterpret_cast<const float4*>(dy.p + pix_off(dy, p) + c0);
      const float xh0 = (v.x - mu.x) * rs.x, xh1 = (v.y - mu.y) * rs.y, xh2 = (v.z - mu.z) * rs.z, xh3 = (v.w - mu.w) * rs.w;
      const float g0 = g.x * act_grad_mask(fmaf(xh0, ga.x, be.x), act);
      const float g1 = g.y * act_grad_mask(fmaf(xh1, ga.y, be.y), act);
      const float g2 = g.z * act_grad_mask(fmaf(xh2, ga.z, be.z), act);
      const float g3 = g.w * act_grad_mask(fmaf(xh3, ga.w, be.w), act);
      s0.x += g0; s0.y += g1; s0.z += g2; s0.w += g3;
      s1.x = fmaf(g0, xh0, s1.x); s1.y = fmaf(g1, xh1, s1.y); s1.z = fmaf(g2, xh2, s1.z); s1.w = fmaf(g3, xh3, s1.w);
    }
  }
  red[0][cq * 4 + 0][pg] = s0.x; red[0][cq * 4 + 1][pg] = s0.y; red[0][cq * 4 + 2][pg] = s0.z; red[0][cq * 4 + 3][pg] = s0.w;
  if (MODE >= 2) {
    red[1][cq * 4 + 0][pg] = s1.x; red[1][cq * 4 + 1][pg] = s1.y; red[1][cq * 4 + 2][pg] = s1.z; red[1][cq * 4 + 3][pg] = s1.w;
  }
  __syncthreads();
  const int nred = (MODE >= 2) ? 64 : 32;
  if (tid < nred) {
    const int which = tid >> 5, c = tid & 31;
    double s = 0.0;
#pragma unroll 8
    for (int j = 0; j < 32; ++j) s += (double)red[which][c][j];
    atomicAdd((which ? out1 : out0) + blockIdx.y * 32 + c, s);
  }
  if (fin.ticket == nullptr) return;
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(fin.ticket + blockIdx.y, 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (tid < 32) {
    const int c = blockIdx.y * 32 + tid;
    const double S0 = *(volatile double*)(out0 + c);
    const double S1 = (MODE >= 2) ? *(volatile double*)(out1 + c) : 0.0;
    if (MODE == 0) {
      fin.o0[c] = (float)S0;
    } else if (MODE == 3) {
      const double m = S0 * fin.inv_count;
      const double vv = S1 * fin.inv_count - m * m;
      fin.o0[c] = (float)m;
      fin.o1[c] = (float)(vv > 0.0 ? vv : 0.0);
    } else if (MODE == 2) {
      const double us = fin.unscale ? (double)__ldg(fin.unscale) : 1.0;
      fin.o0[c] = (float)(S0 * us);
      fin.o1[c] = (float)(S1 * us);
      if (fin.coef) {
        const float r = 1.f / sqrtf(var[c] + eps);
        fin.coef[c] = r;
        fin.coef[fin.C + c] = gamma[c] * r;
        fin.coef[2 * fin.C + c] = fin.train ? (float)(S0 * fin.inv_count) : 0.f;
        fin.coef[3 * fin.C + c] = fin.train ? (float)(S1 * fin.inv_count) : 0.f;
      }
    }
    out0[c] = 0.0;
    if (MODE >= 2) out1[c] = 0.0;
  }
  if (tid == 0) fin.ticket[blockIdx.y] = 0;
}

// arbitrary (small) C: one thread per channel, serial over pixels.  Only used for tiny tensors.
__global__ void colsum_generic_kernel(V x, double* __restrict__ out) {
  pdl_entry();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= x.c) return;
  const long long total = (long long)x.n * x.h * x.w;
  double s = 0.0;
  for (long long p = blockIdx.y; p < total; p += gridDim.y) s += (double)x.p[pix_off(x, p) + c];
  atomicAdd(out + c, s);
}

__global__ void finalize_div_kernel(double* __restrict__ s, float* __restrict__ out, int C, double inv, int accumulate) {
  pdl_entry();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    out[c] = (accumulate ? out[c] : 0.f) + (float)(s[c] * inv);
    s[c] = 0.0;   // workspace invariant: zero after every call
  }
}

// Per-channel-quad BN constants.  The element-wise kernels below walk a grid-stride loop over float4 channel quads; with
// 256-thread blocks and a power-of-two channel count the stride is a multiple of C/4, so a thread meets the SAME quad in
// every iteration and the constants (with their sqrt + reciprocal per channel) are computed once, not per element --
// the 8 MUFU operations per float4 were what kept these kernels at ~3.5 TB/s.
struct BnQuad {
  float4 mu, rs, ga, be;
};
__device__ __forceinline__ BnQuad bn_quad(const float* __restrict__ mean, const float* __restrict__ var,
                                          const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int q) {
  BnQuad c;
  c.mu = __ldg(reinterpret_cast<const float4*>(mean + q));
  const float4 vv = __ldg(reinterpret_cast<const float4*>(var + q));
  c.rs = make_float4(1.f / sqrtf(vv.x + eps), 1.f / sqrtf(vv.y + eps), 1.f / sqrtf(vv.z + eps), 1.f / sqrtf(vv.w + eps));
  c.ga = __ldg(reinterpret_cast<const float4*>(gamma + q));
  c.be = __ldg(reinterpret_cast<const float4*>(beta + q));
  return c;
}

// SPLIT: y receives hi = rna_tf32(v) and ylo receives rna_tf32(v - hi) (operand pair of a 3xTF32 GEMM)
template <bool SPLIT>
__global__ void __launch_bounds__(256)
bn_apply_kernel(V x, V y, V ylo, const float* __restrict__ mean, const float* __restrict__ var,
                const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int act) {
  pdl_entry();
  const int C4 = x.c >> 2;
  const FastDiv x_fc4 = x.fc4;
  const long long total = (long long)x.n * x.h * x.w * C4;
  const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  const bool fixed_q = (stride % C4) == 0;
  BnQuad co = bn_quad(mean, var, gamma, beta, eps, (int)(i0 % C4) * 4);
  for (long long i = i0; i < total; i += stride) {
    const uint32_t pp = fd_div((uint32_t)i, x_fc4);
    const int q = (int)((uint32_t)i - pp * (uint32_t)C4) * 4;
    const long long p = pp;
    const float4 v = *reinterpret_cast<const float4*>(x.p + pix_off(x, p) + q);
    if (!fixed_q) co = bn_quad(mean, var, gamma, beta, eps, q);
    float4 o;
    o.x = apply_act(fmaf((v.x - co.mu.x) * co.rs.x, co.ga.x, co.be.x), act);
    o.y = apply_act(fmaf((v.y - co.mu.y) * co.rs.y, co.ga.y, co.be.y), act);
    o.z = apply_act(fmaf((v.z - co.mu.z) * co.rs.z, co.ga.z, co.be.z), act);
    o.w = apply_act(fmaf((v.w - co.mu.w) * co.rs.w, co.ga.w, co.be.w), act);
    if (SPLIT) {
      const float4 hi = make_float4(round_tf32(o.x), round_tf32(o.y), round_tf32(o.z), round_tf32(o.w));
      *reinterpret_cast<float4*>(y.p + pix_off(y, p) + q) = hi;
      *reinterpret_cast<float4*>(ylo.p + pix_off(ylo, p) + q) =
          make_float4(round_tf32(o.x - hi.x), round_tf32(o.y - hi.y), round_tf32(o.z - hi.z), round_tf32(o.w - hi.w));
    } else {
      *reinterpret_cast<float4*>(y.p + pix_off(y, p) + q) = o;
    }
  }
}

__device__ __forceinline__ uint32_t pack_half2_sat(float lo, float hi) {
  uint16_t a, b;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(a) : "f"(lo));
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(b) : "f"(hi));
  return (uint32_t)a | ((uint32_t)b << 16);
}
__device__ __forceinline__ float half_bits_to_float(uint32_t h) {
  float r;
  asm("cvt.f32.f16 %0, %1;" : "=f"(r) : "h"((uint16_t)h));
  return r;
}

// BN + activation with the result stored as IEEE half (yh, operand of a kind::f16 GEMM) and, when y.p is not
// null, as fp32 holding the SAME half-rounded values (what the backward pass reads).
template <bool XH>
__global__ void __launch_bounds__(256)
bn_apply_h_kernel(V x, V y, V yh, const float* __restrict__ mean, const float* __restrict__ var,
                  const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int act, int same_geo) {
  pdl_entry();
  const int C4 = x.c >> 2;
  const FastDiv x_fc4 = x.fc4;
  const long long total = (long long)x.n * x.h * x.w * C4;
  const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  const bool fixed_q = (stride % C4) == 0;
  BnQuad co = bn_quad(mean, var, gamma, beta, eps, (int)(i0 % C4) * 4);
  for (long long i = i0; i < total; i += stride) {
    const uint32_t pp = fd_div((uint32_t)i, x_fc4);
    const int q = (int)((uint32_t)i - pp * (uint32_t)C4) * 4;
    const long long p = pp;
    const size_t xo = pix_off(x, p);
    const float4 v = XH ? load_half4(x.p, xo + q) : *reinterpret_cast<const float4*>(x.p + xo + q);
    if (!fixed_q) co = bn_quad(mean, var, gamma, beta, eps, q);
    float4 o;
    o.x = apply_act(fmaf((v.x - co.mu.x) * co.rs.x, co.ga.x, co.be.x), act & 0xff);
    o.y = apply_act(fmaf((v.y - co.mu.y) * co.rs.y, co.ga.y, co.be.y), act & 0xff);
    o.z = apply_act(fmaf((v.z - co.mu.z) * co.rs.z, co.ga.z, co.be.z), act & 0xff);
    o.w = apply_act(fmaf((v.w - co.mu.w) * co.rs.w, co.ga.w, co.be.w), act & 0xff);
    const uint32_t p0 = pack_half2_sat(o.x, o.y), p1 = pack_half2_sat(o.z, o.w);
    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(yh.p) + (same_geo ? xo : pix_off(yh, p)) + q) = make_uint2(p0, p1);
    if (y.p)
      *reinterpret_cast<float4*>(y.p + (same_geo ? xo : pix_off(y, p)) + q) =
          make_float4(half_bits_to_float(p0 & 0xffffu), half_bits_to_float(p0 >> 16), half_bits_to_float(p1 & 0xffffu),
                      half_bits_to_float(p1 >> 16));
  }
}

__global__ void __launch_bounds__(256) split_tf32_kernel(V src, V hi, V lo) {
  pdl_entry();
  const int C4 = src.c >> 2;
  const FastDiv x_fc4 = src.fc4;
  const long long total = (long long)src.n * src.h * src.w * C4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const uint32_t pp = fd_div((uint32_t)i, x_fc4);
    const int q = (int)((uint32_t)i - pp * (uint32_t)C4) * 4;
    const long long p = pp;
    const float4 v = *reinterpret_cast<const float4*>(src.p + pix_off(src, p) + q);
    const float4 h = make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
    *reinterpret_cast<float4*>(hi.p + pix_off(hi, p) + q) = h;
    *reinterpret_cast<float4*>(lo.p + pix_off(lo, p) + q) =
        make_float4(round_tf32(v.x - h.x), round_tf32(v.y - h.y), round_tf32(v.z - h.z), round_tf32(v.w - h.w));
  }
}

// HALF: dx is an IEEE-half view and receives dx * (*oscale) (loss-scaled operand of the kind::f16 GEMMs)
// DYH: dy is an IEEE-half view (already loss-scaled: dx inherits the scale, oscale must be null)
template <bool HALF, bool DYH = false>
__global__ void __launch_bounds__(256)
bn_bwd_dx_kernel(V x, V dy, V dx, const float* __restrict__ mean, const float* __restrict__ gamma,
                 const float* __restrict__ beta, int act, const float* __restrict__ coef, const float* __restrict__ oscale,
                 int same_geo) {
  pdl_entry();
  const float os = (HALF && oscale) ? __ldg(oscale) : 1.f;
  const int C = x.c, C4 = C >> 2;
  const FastDiv x_fc4 = x.fc4;
  const long long total = (long long)x.n * x.h * x.w * C4;
  const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  const bool fixed_q = (stride % C4) == 0;     // see BnQuad: the same channel quad in every iteration
  float4 mu, ga, be, rs, k1, m0, m1;
  auto load_q = [&](int q) {
    mu = __ldg(reinterpret_cast<const float4*>(mean + q));
    ga = __ldg(reinterpret_cast<const float4*>(gamma + q));
    be = __ldg(reinterpret_cast<const float4*>(beta + q));
    rs = __ldg(reinterpret_cast<const float4*>(coef + q));
    k1 = __ldg(reinterpret_cast<const float4*>(coef + C + q));
    m0 = __ldg(reinterpret_cast<const float4*>(coef + 2 * C + q));
    m1 = __ldg(reinterpret_cast<const float4*>(coef + 3 * C + q));
  };
  load_q((int)(i0 % C4) * 4);
  for (long long i = i0; i < total; i += stride) {
    const uint32_t pp = fd_div((uint32_t)i, x_fc4);
    const int q = (int)((uint32_t)i - pp * (uint32_t)C4) * 4;
    const long long p = pp;
    const size_t xo = pix_off(x, p);
    const size_t go = same_geo ? xo : pix_off(dy, p);
    const float4 v = *reinterpret_cast<const float4*>(x.p + xo + q);
    const float4 g = DYH ? load_half4(dy.p, go + q) : *reinterpret_cast<const float4*>(dy.p + go + q);
    if (!fixed_q) load_q(q);
    const float vin[4] = {v.x, v.y, v.z, v.w}, gin[4] = {g.x, g.y, g.z, g.w};
    const float mua[4] = {mu.x, mu.y, mu.z, mu.w}, gaa[4] = {ga.x, ga.y, ga.z, ga.w}, bea[4] = {be.x, be.y, be.z, be.w};
    const float rsa[4] = {rs.x, rs.y, rs.z, rs.w}, k1a[4] = {k1.x, k1.y, k1.z, k1.w};
    const float m0a[4] = {m0.x, m0.y, m0.z, m0.w}, m1a[4] = {m1.x, m1.y, m1.z, m1.w};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float xh = (vin[j] - mua[j]) * rsa[j];
      const float gg = gin[j] * act_grad_mask(fmaf(xh, gaa[j], bea[j]), act);
      o[j] = k1a[j] * (gg - m0a[j] - xh * m1a[j]);     // inference-mode BN: m0 = m1 = 0
      if (act & MYOLO_ROUND_TF32) o[j] = round_tf32(o[j]);
    }
    if (HALF)
      *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(dx.p) + (same_geo ? xo : pix_off(dx, p)) + q) =
          make_uint2(pack_half2_sat(o[0] * os, o[1] * os), pack_half2_sat(o[2] * os, o[3] * os));
    else
      *reinterpret_cast<float4*>(dx.p + (same_geo ? xo : pix_off(dx, p)) + q) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// Second half of the backward of a BATCH-statistics BN whose first half ran in a GEMM epilogue (myolo_gemm_taps_bnbwd_sums_h):
// the epilogue left g1 = gamma*rs * g (g = dy * act', half, loss-scaled) and the column sums S0 = sum g, S1 = sum g*xhat.
// dx = gamma*rs * (g - S0/n - xhat * S1/n) = g1 - A - B * (z - mean),  A = gamma*rs*S0/n,  B = gamma*rs^2*S1/n.
__global__ void bn_batch_coef_kernel(double* __restrict__ sums, const float* __restrict__ gamma, const float* __restrict__ var,
                                     float eps, double inv_count, const float* __restrict__ unscale, float* __restrict__ dgamma,
                                     float* __restrict__ dbeta, float* __restrict__ coef, int C) {
  pdl_entry();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double S0 = sums[c], S1 = sums[C + c];
  const double us = unscale ? (double)__ldg(unscale) : 1.0;
  dbeta[c] = (float)(S0 * us);
  dgamma[c] = (float)(S1 * us);
  const double r = 1.0 / sqrt((double)var[c] + (double)eps), gr = (double)gamma[c] * r;
  coef[c] = (float)(gr * S0 * inv_count);
  coef[C + c] = (float)(gr * r * S1 * inv_count);
  sums[c] = 0.0;
  sums[C + c] = 0.0;
}

__global__ void __launch_bounds__(256)
bn_batch_fix_hh_kernel(V z, V g, const float* __restrict__ mean, const float* __restrict__ coef, int same_geo) {
  pdl_entry();
  const int C = z.c, C4 = C >> 2;
  const FastDiv x_fc4 = z.fc4;
  const long long total = (long long)z.n * z.h * z.w * C4;
  const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  const bool fixed_q = (stride % C4) == 0;
  float4 mu, ka, kb;
  auto load_q = [&](int q) {
    mu = __ldg(reinterpret_cast<const float4*>(mean + q));
    ka = __ldg(reinterpret_cast<const float4*>(coef + q));
    kb = __ldg(reinterpret_cast<const float4*>(coef + C + q));
  };
  load_q((int)(i0 % C4) * 4);
  for (long long i = i0; i < total; i += stride) {
    const uint32_t pp = fd_div((uint32_t)i, x_fc4);
    const int q = (int)((uint32_t)i - pp * (uint32_t)C4) * 4;
    const long long p = pp;
    const size_t zo = pix_off(z, p);
    const size_t go = same_geo ? zo : pix_off(g, p);
    const float4 v = load_half4(z.p, zo + q);
    const float4 gv = load_half4(g.p, go + q);
    if (!fixed_q) load_q(q);
    const float o0 = gv.x - ka.x - kb.x * (v.x - mu.x), o1 = gv.y - ka.y - kb.y * (v.y - mu.y);
    const float o2 = gv.z - ka.z - kb.z * (v.z - mu.z), o3 = gv.w - ka.w - kb.w * (v.w - mu.w);
    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(g.p) + go + q) = make_uint2(pack_half2_sat(o0, o1), pack_half2_sat(o2, o3));
  }
}

// ---- 8 channels (16 bytes of half) per thread and two pixels in flight: the one-tensor passes of myolo_mask_bn1 over the
// 0.5 GB half tensors of the mask head.  (The 4-channel kernels above move 8 bytes per load with one load in flight per
// thread: 3.0 TB/s on the apply pass against 4.8 TB/s for the correction pass, which has two.)
__device__ __forceinline__ void unpack_half8(const uint4& u, float (&f)[8]) {
  f[0] = half_bits_to_float(u.x & 0xffffu); f[1] = half_bits_to_float(u.x >> 16);
  f[2] = half_bits_to_float(u.y & 0xffffu); f[3] = half_bits_to_float(u.y >> 16);
  f[4] = half_bits_to_float(u.z & 0xffffu); f[5] = half_bits_to_float(u.z >> 16);
  f[6] = half_bits_to_float(u.w & 0xffffu); f[7] = half_bits_to_float(u.w >> 16);
}
__device__ __forceinline__ uint4 pack_half8(const float (&f)[8]) {
  return make_uint4(pack_half2_sat(f[0], f[1]), pack_half2_sat(f[2], f[3]), pack_half2_sat(f[4], f[5]), pack_half2_sat(f[6], f[7]));
}

__global__ void __launch_bounds__(256)
bn_apply_hh8_kernel(V x, V yh, const float* __restrict__ mean, const float* __restrict__ var, const float* __restrict__ gamma,
                    const float* __restrict__ beta, float eps, int act, int same_geo, FastDiv fc8) {
  pdl_entry();
  const int C8 = x.c >> 3;
  const long long total = (long long)x.n * x.h * x.w * C8;
  const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  const bool fixed_q = (stride % C8) == 0;
  float mu[8], rs[8], ga[8], be[8];      // the same arithmetic as bn_apply_kernel: act(((x - mu) * rs) * ga + be)
  auto load_q = [&](int q) {
    const BnQuad a = bn_quad(mean, var, gamma, beta, eps, q), b = bn_quad(mean, var, gamma, beta, eps, q + 4);
    mu[0] = a.mu.x; mu[1] = a.mu.y; mu[2] = a.mu.z; mu[3] = a.mu.w; mu[4] = b.mu.x; mu[5] = b.mu.y; mu[6] = b.mu.z; mu[7] = b.mu.w;
    rs[0] = a.rs.x; rs[1] = a.rs.y; rs[2] = a.rs.z; rs[3] = a.rs.w; rs[4] = b.rs.x; rs[5] = b.rs.y; rs[6] = b.rs.z; rs[7] = b.rs.w;
    ga[0] = a.ga.x; ga[1] = a.ga.y; ga[2] = a.ga.z; ga[3] = a.ga.w; ga[4] = b.ga.x; ga[5] = b.ga.y; ga[6] = b.ga.z; ga[7] = b.ga.w;
    be[0] = a.be.x; be[1] = a.be.y; be[2] = a.be.z; be[3] = a.be.w; be[4] = b.be.x; be[5] = b.be.y; be[6] = b.be.z; be[7] = b.be.w;
  };
  load_q((int)(i0 % C8) * 8);
  const uint16_t* xp = reinterpret_cast<const uint16_t*>(x.p);
  uint16_t* yp = reinterpret_cast<uint16_t*>(yh.p);
  for (long long i = i0; i < total; i += 2 * stride) {
    const long long i2 = i + stride;
    const bool two = i2 < total;
    const uint32_t pa = fd_div((uint32_t)i, fc8), pb = fd_div((uint32_t)(two ? i2 : i), fc8);
    const int qa = (int)((uint32_t)i - pa * (uint32_t)C8) * 8, qb = (int)((uint32_t)(two ? i2 : i) - pb * (uint32_t)C8) * 8;
    const size_t xa = pix_off(x, pa), xb = pix_off(x, pb);
    const uint4 ua = *reinterpret_cast<const uint4*>(xp + xa + qa);
    const uint4 ub = *reinterpret_cast<const uint4*>(xp + xb + qb);
    float f[8];
    if (!fixed_q) load_q(qa);
    unpack_half8(ua, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = apply_act(fmaf((f[j] - mu[j]) * rs[j], ga[j], be[j]), act & 0xff);
    *reinterpret_cast<uint4*>(yp + (same_geo ? xa : pix_off(yh, pa)) + qa) = pack_half8(f);
    if (two) {
      if (!fixed_q) load_q(qb);
      unpack_half8(ub, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = apply_act(fmaf((f[j] - mu[j]) * rs[j], ga[j], be[j]), act & 0xff);
      *reinterpret_cast<uint4*>(yp + (same_geo ? xb : pix_off(yh, pb)) + qb) = pack_half8(f);
    }
  }
}

__global__ void __launch_bounds__(256)
bn_batch_fix_hh8_kernel(V z, V g, const float* __restrict__ mean, const float* __restrict__ coef, int same_geo, FastDiv fc8) {
  pdl_entry();
  const int C = z.c, C8 = C >> 3;
  const long long total = (long long)z.n * z.h * z.w * C8;
  const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  const bool fixed_q = (stride % C8) == 0;
  float mu[8], ka[8], kb[8];
  auto load_q = [&](int q) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 m = __ldg(reinterpret_cast<const float4*>(mean + q + 4 * h));
      const float4 a = __ldg(reinterpret_cast<const float4*>(coef + q + 4 * h));
      const float4 b = __ldg(reinterpret_cast<const float4*>(coef + C + q + 4 * h));
      mu[4 * h] = m.x; mu[4 * h + 1] = m.y; mu[4 * h + 2] = m.z; mu[4 * h + 3] = m.w;
      ka[4 * h] = a.x; ka[4 * h + 1] = a.y; ka[4 * h + 2] = a.z; ka[4 * h + 3] = a.w;
      kb[4 * h] = b.x; kb[4 * h + 1] = b.y; kb[4 * h + 2] = b.z; kb[4 * h + 3] = b.w;
    }
  };
  load_q((int)(i0 % C8) * 8);
  const uint16_t* zp = reinterpret_cast<const uint16_t*>(z.p);
  uint16_t* gp = reinterpret_cast<uint16_t*>(g.p);
  for (long long i = i0; i < total; i += stride) {
    const uint32_t pp = fd_div((uint32_t)i, fc8);
    const int q = (int)((uint32_t)i - pp * (uint32_t)C8) * 8;
    const size_t zo = pix_off(z, pp);
    const size_t go = same_geo ? zo : pix_off(g, pp);
    const uint4 uz = *reinterpret_cast<const uint4*>(zp + zo + q);
    const uint4 ug = *reinterpret_cast<const uint4*>(gp + go + q);
    if (!fixed_q) load_q(q);
    float fz[8], fg[8];
    unpack_half8(uz, fz);
    unpack_half8(ug, fg);
#pragma unroll
    for (int j = 0; j < 8; ++j) fg[j] = fg[j] - ka[j] - kb[j] * (fz[j] - mu[j]);
    *reinterpret_cast<uint4*>(gp + go + q) = pack_half8(fg);
  }
}

__global__ void bn_moving_update_kernel(const float* __restrict__ value, float* __restrict__ biased,
                                        float* __restrict__ moving, int C, float momentum, float corr, float debias) {
  pdl_entry();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    const float v = value[c] * corr;
    const float b = biased[c] - (biased[c] - v) * (1.f - momentum);
    biased[c] = b;
    moving[c] = b / debias;
  }
}

// all BN layers of the model in one launch: block = one (layer, statistic) item
struct MovingItem {
  const float* value;
  float* biased;
  float* moving;
  int C;
  float corr;
};
__global__ void bn_moving_update_batch_kernel(const MovingItem* __restrict__ items, float momentum, float debias) {
  pdl_entry();
  const MovingItem it = items[blockIdx.x];
  for (int c = threadIdx.x; c < it.C; c += blockDim.x) {
    const float v = it.value[c] * it.corr;
    const float b = it.biased[c] - (it.biased[c] - v) * (1.f - momentum);
    it.biased[c] = b;
    it.moving[c] = b / debias;
  }
}

__global__ void __launch_bounds__(256) view_copy_kernel(V src, V dst, int accumulate) {
  pdl_entry();
  const int C4 = src.c >> 2;
  const FastDiv x_fc4 = src.fc4;
  const long long total = (long long)src.n * src.h * src.w * C4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const uint32_t pp = fd_div((uint32_t)i, x_fc4);
    const int q = (int)((uint32_t)i - pp * (uint32_t)C4) * 4;
    const long long p = pp;
    float4 v = *reinterpret_cast<const float4*>(src.p + pix_off(src, p) + q);
    float4* d = reinterpret_cast<float4*>(dst.p + pix_off(dst, p) + q);
    if (accumulate & 1) {
      const float4 o = *d;
      v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
    }
    if (accumulate & 2) v = make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
    *d = v;
  }
}

// scale = gamma * rsqrt(var + eps), shift = beta - mean * scale: inference-mode BN folded into the
// producing GEMM's epilogue (myolo_gemm_taps scale/shift_c arguments).
__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ mean, const float* __restrict__ var, float eps,
                               float* __restrict__ scale, float* __restrict__ shift, int C) {
  pdl_entry();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    const float s = gamma[c] * (1.f / sqrtf(var[c] + eps));
    scale[c] = s;
    shift[c] = fmaf(-mean[c], s, beta[c]);
  }
}

// Backward of a = act(gamma * xhat + beta) with FIXED (moving) statistics, computed from the
// OUTPUT a alone: where the activation passes, gamma * xhat = a - beta.  One pass: reads a and dy,
// writes dx = dy * act'(a) * gamma * rs, accumulates dbeta = sum dy*act' and dgamma = sum dy*act'*xhat.
__global__ void __launch_bounds__(256)
bn_act_bwd_from_output_kernel(V a, V dy, V dx, const float* __restrict__ gamma, const float* __restrict__ beta,
                              const float* __restrict__ var, float eps, int act, double* __restrict__ out0,
                              double* __restrict__ out1, long long chunk, float* __restrict__ dgamma,
                              float* __restrict__ dbeta, float* __restrict__ dbias, int* __restrict__ ticket) {
  pdl_entry();
  __shared__ float red[2][32][33];
  __shared__ int s_last;
  const int tid = threadIdx.x;
  const int cq = tid & 7, pg = tid >> 3;
  const int c0 = blockIdx.y * 32 + cq * 4;
  const long long total = (long long)a.n * a.h * a.w;
  const long long p0 = blockIdx.x * chunk, p1 = min(total, p0 + chunk);
  const float4 ga = *reinterpret_cast<const float4*>(gamma + c0);
  const float4 be = *reinterpret_cast<const float4*>(beta + c0);
  const float4 vv = *reinterpret_cast<const float4*>(var + c0);
  const float gav[4] = {ga.x, ga.y, ga.z, ga.w}, bev[4] = {be.x, be.y, be.z, be.w};
  const float vvv[4] = {vv.x, vv.y, vv.z, vv.w};
  float sc[4], ig[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sc[j] = gav[j] * (1.f / sqrtf(vvv[j] + eps));
    ig[j] = 1.f / (fabsf(gav[j]) < 1e-20f ? copysignf(1e-20f, gav[j]) : gav[j]);
  }
  float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
  const int aa = act & 0xff;
  for (long long p = p0 + pg; p < p1; p += 32) {
    const float4 av = *reinterpret_cast<const float4*>(a.p + pix_off(a, p) + c0);
    const float4 gv = *reinterpret_cast<const float4*>(dy.p + pix_off(dy, p) + c0);
    const float ain[4] = {av.x, av.y, av.z, av.w}, gin[4] = {gv.x, gv.y, gv.z, gv.w};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      bool pass = true;
      if (aa == MYOLO_ACT_RELU) pass = ain[j] > 0.f;
      else if (aa == MYOLO_ACT_RELU6) pass = ain[j] > 0.f && ain[j] < 6.f;
      const float g = pass ? gin[j] : 0.f;
      s0[j] += g;
      s1[j] = fmaf(g, (ain[j] - bev[j]) * ig[j], s1[j]);
      o[j] = g * sc[j];
      if (act & MYOLO_ROUND_TF32) o[j] = round_tf32(o[j]);
    }
    *reinterpret_cast<float4*>(dx.p + pix_off(dx, p) + c0) = make_float4(o[0], o[1], o[2], o[3]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    red[0][cq * 4 + j][pg] = s0[j];
    red[1][cq * 4 + j][pg] = s1[j];
  }
  __syncthreads();
  if (tid < 64) {
    const int which = tid >> 5, c = tid & 31;
    double s = 0.0;
#pragma unroll 8
    for (int j = 0; j < 32; ++j) s += (double)red[which][c][j];
    atomicAdd((which ? out1 : out0) + blockIdx.y * 32 + c, s);
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(ticket + blockIdx.y, 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (tid < 32) {
    const int c = blockIdx.y * 32 + tid;
    const double S0 = *(volatile double*)(out0 + c), S1 = *(volatile double*)(out1 + c);
    dbeta[c] = (float)S0;
    dgamma[c] = (float)S1;
    // the conv bias sits before the (fixed-statistics) BN, so dbias = gamma*rs*dbeta
    if (dbias) dbias[c] = (float)(S0 * (double)(gamma[c] * (1.f / sqrtf(var[c] + eps))));
    out0[c] = 0.0;
    out1[c] = 0.0;
  }
  if (tid == 0) ticket[blockIdx.y] = 0;
}

static bool view_ok(const myolo_view* v) {
  return v && v->p && v->n > 0 && v->h > 0 && v->w > 0 && v->c > 0 && (v->c % 4) == 0 && (v->sn % 4) == 0 && (v->sh % 4) == 0;
}
// a half view whose pixels can be moved 8 channels (16 bytes) at a time
static bool half8_ok(const myolo_view* v) {
  return (v->c % 8) == 0 && (v->sn % 8) == 0 && (v->sh % 8) == 0 && ((uintptr_t)v->p & 15) == 0 &&
         (long long)v->n * v->h * v->w * (v->c / 8) < (1LL << 31);
}
static bool same_shape(const myolo_view* a, const myolo_view* b) {
  return a->n == b->n && a->h == b->h && a->w == b->w && a->c == b->c;
}
static void reduce_grid(long long total, int C, dim3* grid, long long* chunk) {
  const int cg = C / 32;
  long long nch = max(1LL, min(ceil_div(total, 64), (long long)(kNumSMs * 8) / cg + 1));
  *chunk = ceil_div(total, nch);
  nch = ceil_div(total, *chunk);
  *grid = dim3((unsigned)nch, (unsigned)cg);
}
// workspace layout (doubles), FIXED so that calls with different C never see each other's leftovers:
//   [0,16) int tickets | [16, 16+2048) sums (2*C used, zero between calls) | [2064, 4112) float coefficient table
constexpr int kWsMaxC = 1024, kWsSums = 16, kWsCoef = 16 + 2 * kWsMaxC;
static int* ws_ticket(double* ws, int) { return reinterpret_cast<int*>(ws); }
static int ew_blocks(long long total) { return (int)max(1LL, min(ceil_div(total, 256), (long long)kNumSMs * 16)); }

}  // namespace myolo

using namespace myolo;

extern "C" int myolo_bn_stats(const myolo_view* x, float* mean, float* var, double* ws, myolo_stream stream) {
  MYOLO_CHECK_ARG(view_ok(x) && mean && var && ws && (x->c % 32) == 0 && x->c <= kWsMaxC);
  cudaStream_t st = as_stream(stream);
  const int C = x->c;
  const long long total = (long long)x->n * x->h * x->w;
  dim3 grid;
  long long chunk;
  reduce_grid(total, C, &grid, &chunk);
  V vx = to_v(x);
  // single pass: sum and sum of squares in fp64 (per-thread fp32 partials over <= chunk/32 pixels)
  Fin fin{mean, var, nullptr, ws_ticket(ws, C), 1.0 / (double)total, 1, C};
  MYOLO_LAUNCH(colreduce_kernel<3>, grid, 256, 0, st, vx, vx, nullptr, nullptr, nullptr, nullptr, 0.f, 0, ws + kWsSums, ws + kWsSums + C, chunk, fin);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

extern "C" int myolo_bn_apply(const myolo_view* x, const myolo_view* y, const float* mean, const float* var,
                              const float* gamma, const float* beta, float eps, int act, myolo_stream stream) {
  MYOLO_CHECK_ARG(view_ok(x) && view_ok(y) && same_shape(x, y) && mean && var && gamma && beta);
  const long long total = (long long)x->n * x->h * x->w * (x->c / 4);
  MYOLO_LAUNCH(bn_apply_kernel<false>, ew_blocks(total), 256, 0, as_stream(stream), to_v(x), to_v(y), to_v(y), mean, var, gamma, beta, eps, act);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

extern "C" int myolo_bn_apply_split(const myolo_view* x, const myolo_view* y_hi, const myolo_view* y_lo, const float* mean,
                                    const float* var, const float* gamma, const float* beta, float eps, int act,
                                    myolo_stream stream) {
  MYOLO_CHECK_ARG(view_ok(x) && view_ok(y_hi) && view_ok(y_lo) && same_shape(x, y_hi) && same_shape(x, y_lo));
  MYOLO_CHECK_ARG(mean && var && gamma && beta);
  const long long total = (long long)x->n * x->h * x->w * (x->c / 4);
  MYOLO_LAUNCH(bn_apply_kernel<true>, ew_blocks(total), 256, 0, as_stream(stream), to_v(x), to_v(y_hi), to_v(y_lo), mean, var, gamma, beta, eps, act);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

extern "C" int myolo_bn_apply_h(const myolo_view* x, const myolo_view* y, const myolo_view* y_half, const float* mean,
                                const float* var, const float* gamma, const float* beta, float eps, int act,
                                myolo_stream stream) {
  MYOLO_CHECK_ARG(view_ok(x) && view_ok(y_half) && same_shape(x, y_half) && (!y || (view_ok(y) && same_shape(x, y))));
  MYOLO_CHECK_ARG(mean && var && gamma && beta && ((uintptr_t)y_half->p & 7) == 0);
  const long long total = (long long)x->n * x->h * x->w * (x->c / 4);
  V vy = y ? to_v(y) : to_v(x);
  if (!y) vy.p = nullptr;
  const int same_geo = x->sn == y_half->sn && x->sh == y_half->sh && (!y || (y->sn == x->sn && y->sh == x->sh));
  MYOLO_LAUNCH(bn_apply_h_kernel<false>, ew_blocks(total), 256, 0, as_stream(stream), to_v(x), vy, to_v(y_half), mean, var, gamma, beta, eps, act, same_geo);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

// BN + activation from an IEEE-half pre-BN tensor to an IEEE-half result (myolo_mask_bn1 in h16 mode)
extern "C" int myolo_bn_apply_hh(const myolo_view* x_half, const myolo_view* y_half, const float* mean, const float* var,
                                 const float* gamma, const float* beta, float eps, int act, myolo_stream stream) {
  MYOLO_CHECK_ARG(view_ok(x_half) && view_ok(y_half) && same_shape(x_half, y_half));
  MYOLO_CHECK_ARG(mean && var && gamma && beta && ((uintptr_t)y_half->p & 7) == 0 && ((uintptr_t)x_half->p & 7) == 0);
  const long long total = (long long)x_half->n * x_half->h * x_half->w * (x_half->c / 4);
  V vy = to_v(x_half);
  vy.p = nullptr;
  const int same_geo = x_half->sn == y_half->sn && x_half->sh == y_half->sh;
  if (half8_ok(x_half) && half8_ok(y_half)) {
    const long long t8 = total / 2;
    MYOLO_LAUNCH(bn_apply_hh8_kernel, ew_blocks(t8 / 2 + 1), 256, 0, as_stream(stream), to_v(x_half), to_v(y_half), mean, var, gamma, beta, eps, act,
                                                                           same_geo, make_fd((uint32_t)(x_half->c / 8)));
    MYOLO_CHECK_LAUNCH();
    return MYOLO_OK;
  }
  MYOLO_LAUNCH(bn_apply_h_kernel<true>, ew_blocks(total), 256, 0, as_stream(stream), to_v(x_half), vy, to_v(y_half), mean, var, gamma, beta, eps, act, same_geo);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

// Finishes the backward of a batch-statistics BN + activation whose masked, gamma*rs-scaled gradient and column sums were
// produced by myolo_gemm_taps_bnbwd_sums_h on the SAME workspace: dgamma / dbeta (un-scaled by *grad_unscale), the sums
// zeroed again, and g_half <- g_half - A - B * (z_half - mean) in place over the view's pixels.
extern "C" int myolo_bn_bwd_batch_fix_hh(const myolo_view* z_half, const myolo_view* g_half, const float* mean, const float* var,
                                         const float* gamma, float eps, float* dgamma, float* dbeta, double* ws,
                                         const float* grad_unscale, myolo_stream stream) {
  MYOLO_CHECK_ARG(view_ok(z_half) && view_ok(g_half) && same_shape(z_half, g_half));
  MYOLO_CHECK_ARG(mean && var && gamma && dgamma && dbeta && ws && z_half->c <= kWsMaxC);
  MYOLO_CHECK_ARG(((uintptr_t)z_half->p & 7) == 0 && ((uintptr_t)g_half->p & 7) == 0);
  cudaStream_t st = as_stream(stream);
  const int C = z_half->c;
  const long long total = (long long)z_half->n * z_half->h * z_half->w;
  float* coef = reinterpret_cast<float*>(ws + kWsCoef);
  MYOLO_LAUNCH(bn_batch_coef_kernel, (C + 127) / 128, 128, 0, st, ws + kWsSums, gamma, var, eps, 1.0 / (double)total, grad_unscale, dgamma, dbeta, coef, C);
  const int same_geo = z_half->sn == g_half->sn && z_half->sh == g_half->sh;
  if (half8_ok(z_half) && half8_ok(g_half)) {
    MYOLO_LAUNCH(bn_batch_fix_hh8_kernel, ew_blocks(total * (C / 8)), 256, 0, st, to_v(z_half), to_v(g_half), mean, coef, same_geo,
                                                                        make_fd((uint32_t)(C / 8)));
    MYOLO_CHECK_LAUNCH();
    return MYOLO_OK;
  }
  MYOLO_LAUNCH(bn_batch_fix_hh_kernel, ew_blocks(total * (C / 4)), 256, 0, st, to_v(z_half), to_v(g_half), mean, coef, same_geo);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

extern "C" int myolo_split_tf32(const myolo_view* src, const myolo_view* hi, const myolo_view* lo, myolo_stream stream) {
  MYOLO_CHECK_ARG(view_ok(src) && view_ok(hi) && view_ok(lo) && same_shape(src, hi) && same_shape(src, lo));
  const long long total = (long long)src->n * src->h * src->w * (src->c / 4);
  MYOLO_LAUNCH(split_tf32_kernel, ew_blocks(total), 256, 0, as_stream(stream), to_v(src), to_v(hi), to_v(lo));
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

extern "C" int myolo_bn_bwd(const myolo_view* x, const myolo_view* dy, const myolo_view* dx, const float* mean,
                            const float* var, const float* gamma, const float* beta, float eps, int act, int train,
                            float* dgamma, float* dbeta, double* ws, myolo_stream stream) {
  MYOLO_CHECK_ARG(view_ok(x) && view_ok(dy) && view_ok(dx) && same_shape(x, dy) && same_shape(x, dx));
  MYOLO_CHECK_ARG(mean && var && gamma && beta && dgamma && dbeta && ws && (x->c % 32) == 0 && x->c <= kWsMaxC);
  cudaStream_t st = as_stream(stream);
  const int C = x->c;
  const long long total = (long long)x->n * x->h * x->w;
  dim3 grid;
  long long chunk;
  reduce_grid(total, C, &grid, &chunk);
  float* coef = reinterpret_cast<float*>(ws + kWsCoef);   // 4*C floats
  Fin fin{dbeta, dgamma, coef, ws_ticket(ws, C), 1.0 / (double)total, train, C};
  MYOLO_LAUNCH(colreduce_kernel<2>, grid, 256, 0, st, to_v(x), to_v(dy), mean, var, gamma, beta, eps, act, ws + kWsSums, ws + kWsSums + C, chunk, fin);
  const int same_geo = x->sn == dy->sn && x->sh == dy->sh && x->sn == dx->sn && x->sh == dx->sh;
  MYOLO_LAUNCH(bn_bwd_dx_kernel<false>, ew_blocks(total * (C / 4)), 256, 0, st, to_v(x), to_v(dy), to_v(dx), mean, gamma, beta, act, coef, nullptr, same_geo);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

extern "C" int myolo_bn_bwd_h(const myolo_view* x, const myolo_view* dy, const myolo_view* dx_half, const float* mean,
                              const float* var, const float* gamma, const float* beta, float eps, int act, int train,
                              float* dgamma, float* dbeta, double* ws, const float* out_scale, myolo_stream stream) {
  MYOLO_CHECK_ARG(view_ok(x) && view_ok(dy) && view_ok(dx_half) && same_shape(x, dy) && same_shape(x, dx_half));
  MYOLO_CHECK_ARG(mean && var && gamma && beta && dgamma && dbeta && ws && (x->c % 32) == 0 && x->c <= kWsMaxC);
  MYOLO_CHECK_ARG(((uintptr_t)dx_half->p & 7) == 0 && !(act & MYOLO_ROUND_TF32));
  cudaStream_t st = as_stream(stream);
  const int C = x->c;
  const long long total = (long long)x->n * x->h * x->w;
  dim3 grid;
  long long chunk;
  reduce_grid(total, C, &grid, &chunk);
  float* coef = reinterpret_cast<float*>(ws + kWsCoef);
  Fin fin{dbeta, dgamma, coef, ws_ticket(ws, C), 1.0 / (double)total, train, C};
  MYOLO_LAUNCH(colreduce_kernel<2>, grid, 256, 0, st, to_v(x), to_v(dy), mean, var, gamma, beta, eps, act, ws + kWsSums, ws + kWsSums + C, chunk, fin);
  const int same_geo = x->sn == dy->sn && x->sh == dy->sh && x->sn == dx_half->sn && x->sh == dx_half->sh;
  MYOLO_LAUNCH(bn_bwd_dx_kernel<true>, ew_blocks(total * (C / 4)), 256, 0, st, to_v(x), to_v(dy), to_v(dx_half), mean, gamma, beta, act, coef, out_scale, same_geo);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

extern "C" int myolo_bn_bwd_hh(const myolo_view* x, const myolo_view* dy_half, const myolo_view* dx_half, const float* mean,
                               const float* var, const float* gamma, const float* beta, float eps, int act, int train,
                               float* dgamma, float* dbeta, double* ws, const float* grad_unscale, myolo_stream stream) {
  MYOLO_CHECK_ARG(view_ok(x) && view_ok(dy_half) && view_ok(dx_half) && same_shape(x, dy_half) && same_shape(x, dx_half));
  MYOLO_CHECK_ARG(mean && var && gamma && beta && dgamma && dbeta && ws && (x->c % 32) == 0 && x->c <= kWsMaxC);
  MYOLO_CHECK_ARG(((uintptr_t)dx_half->p & 7) == 0 && ((uintptr_t)dy_half->p & 7) == 0 && !(act & MYOLO_ROUND_TF32));
  cudaStream_t st = as_stream(stream);
  const int C = x->c;
  const long long total = (long long)x->n * x->h * x->w;
  dim3 grid;
  long long chunk;
  reduce_grid(total, C, &grid, &chunk);
  float* coef = reinterpret_cast<float*>(ws + kWsCoef);
  Fin fin{dbeta, dgamma, coef, ws_ticket(ws, C), 1.0 / (double)total, train, C, grad_unscale};
  MYOLO_LAUNCH((colreduce_kernel<2, true>), grid, 256, 0, st, to_v(x), to_v(dy_half), mean, var, gamma, beta, eps, act, ws + kWsSums, ws + kWsSums + C, chunk, fin);
  const int same_geo = x->sn == dy_half->sn && x->sh == dy_half->sh && x->sn == dx_half->sn && x->sh == dx_half->sh;
  MYOLO_LAUNCH((bn_bwd_dx_kernel<true, true>), ew_blocks(total * (C / 4)), 256, 0, st, to_v(x), to_v(dy_half), to_v(dx_half), mean, gamma, beta, act, coef,
                                                                         nullptr, same_geo);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

extern "C" int myolo_bn_moving_update(const float* value, float* biased, float* moving, int C, float momentum, int step,
                                      int is_var, double n, float eps, myolo_stream stream) {
  MYOLO_CHECK_ARG(value && biased && moving && C > 0 && step >= 1 && n > 0);
  double corr = 1.0;
  if (is_var) corr = (n / (n > 1 ? n - 1 : 1.0)) * (n / (n - (1.0 + (double)eps)));
  double debias = 1.0;
  {
    double pw = 1.0;
    for (int i = 0; i < step && pw > 1e-300; ++i) pw *= (double)momentum;
    debias = 1.0 - pw;
  }
  MYOLO_LAUNCH(bn_moving_update_kernel, (C + 127) / 128, 128, 0, as_stream(stream), value, biased, moving, C, momentum, (float)corr, (float)debias);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

extern "C" int myolo_colsum(const myolo_view* x, float* out, double* ws, myolo_stream stream) {
  MYOLO_CHECK_ARG(x && x->p && x->n > 0 && x->h > 0 && x->w > 0 && x->c > 0 && x->c <= 2 * kWsMaxC && out && ws);
  cudaStream_t st = as_stream(stream);
  const int C = x->c;
  const long long total = (long long)x->n * x->h * x->w;
  V vx = to_v(x);
  if ((C % 32) == 0 && (x->sn % 4) == 0 && (x->sh % 4) == 0) {
    dim3 grid;
    long long chunk;
    reduce_grid(total, C, &grid, &chunk);
    Fin fin{out, nullptr, nullptr, ws_ticket(ws, C), 1.0, 0, C};
    MYOLO_LAUNCH(colreduce_kernel<0>, grid, 256, 0, st, vx, vx, nullptr, nullptr, nullptr, nullptr, 0.f, 0, ws + kWsSums, nullptr, chunk, fin);
  } else {
    dim3 grid((C + 63) / 64, (unsigned)min(total, 256LL));
    MYOLO_LAUNCH(colsum_generic_kernel, grid, 64, 0, st, vx, ws + kWsSums);
    MYOLO_LAUNCH(finalize_div_kernel, (C + 127) / 128, 128, 0, st, ws + kWsSums, out, C, 1.0, 0);
  }
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

extern "C" int myolo_view_copy(const myolo_view* src, const myolo_view* dst, int accumulate, myolo_stream stream) {
  MYOLO_CHECK_ARG(view_ok(src) && view_ok(dst) && same_shape(src, dst));
  const long long total = (long long)src->n * src->h * src->w * (src->c / 4);
  MYOLO_LAUNCH(view_copy_kernel, ew_blocks(total), 256, 0, as_stream(stream), to_v(src), to_v(dst), accumulate);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

extern "C" int myolo_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                             float* scale, float* shift, int C, myolo_stream stream) {
  MYOLO_CHECK_ARG(gamma && beta && mean && var && scale && shift && C > 0);
  MYOLO_LAUNCH(bn_fold_kernel, (C + 127) / 128, 128, 0, as_stream(stream), gamma, beta, mean, var, eps, scale, shift, C);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

extern "C" int myolo_bn_act_bwd_from_output(const myolo_view* a, const myolo_view* dy, const myolo_view* dx,
                                            const float* gamma, const float* beta, const float* var, float eps, int act,
                                            float* dgamma, float* dbeta, float* dbias, double* ws, myolo_stream stream) {
  MYOLO_CHECK_ARG(view_ok(a) && view_ok(dy) && view_ok(dx) && same_shape(a, dy) && same_shape(a, dx));
  MYOLO_CHECK_ARG(gamma && beta && var && dgamma && dbeta && ws && (a->c % 32) == 0 && a->c <= kWsMaxC);
  cudaStream_t st = as_stream(stream);
  const int C = a->c;
  const long long total = (long long)a->n * a->h * a->w;
  dim3 grid;
  long long chunk;
  reduce_grid(total, C, &grid, &chunk);
  MYOLO_LAUNCH(bn_act_bwd_from_output_kernel, grid, 256, 0, st, to_v(a), to_v(dy), to_v(dx), gamma, beta, var, eps, act, ws + kWsSums, ws + kWsSums + C, chunk,
                                                      dgamma, dbeta, dbias, ws_ticket(ws, C));
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

extern "C" int myolo_bn_moving_update_batch(const void* items_dev, int n_items, float momentum, int step, myolo_stream stream) {
  MYOLO_CHECK_ARG(items_dev && n_items > 0 && step >= 1);
  double pw = 1.0;
  for (int i = 0; i < step && pw > 1e-300; ++i) pw *= (double)momentum;
  MYOLO_LAUNCH(bn_moving_update_batch_kernel, n_items, 256, 0, as_stream(stream), reinterpret_cast<const MovingItem*>(items_dev), momentum,
                                                                         (float)(1.0 - pw));
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

namespace myolo {
__global__ void bn_epi_finalize_kernel(double* __restrict__ sums, const float* __restrict__ gamma, const float* __restrict__ var,
                                       float eps, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias,
                                       int C, const float* __restrict__ unscale) {
  pdl_entry();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    const double us = unscale ? (double)__ldg(unscale) : 1.0;   // the gradient tensor carried a loss scale
    const double S0 = sums[c] * us, S1 = sums[C + c] * us;
    dbeta[c] = (float)S0;
    dgamma[c] = (float)S1;
    if (dbias) dbias[c] = (float)(S0 * (double)(gamma[c] * (1.f / sqrtf(var[c] + eps))));
    sums[c] = 0.0;
    sums[C + c] = 0.0;
  }
}
}  // namespace myolo

// finishes the column sums a GEMM epilogue accumulated (myolo_gemm_taps_bnbwd) and zeroes them again
extern "C" int myolo_bn_epi_finalize_s(double* sums, const float* gamma, const float* var, float eps, float* dgamma,
                                       float* dbeta, float* dbias, int C, const float* unscale, myolo_stream stream) {
  MYOLO_CHECK_ARG(sums && gamma && var && dgamma && dbeta && C > 0);
  MYOLO_LAUNCH(myolo::bn_epi_finalize_kernel, (C + 127) / 128, 128, 0, myolo::as_stream(stream), sums, gamma, var, eps, dgamma, dbeta, dbias, C, unscale);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}
extern "C" int myolo_bn_epi_finalize(double* sums, const float* gamma, const float* var, float eps, float* dgamma,
                                     float* dbeta, float* dbias, int C, myolo_stream stream) {
  return myolo_bn_epi_finalize_s(sums, gamma, var, eps, dgamma, dbeta, dbias, C, nullptr, stream);
}
