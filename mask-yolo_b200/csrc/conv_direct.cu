// conv_direct.cu -- HBM-bound direct convolutions: the 3->32 stem conv (K1) and the 14 depthwise
// 3x3 layers (K2) of the truncated MobileNet backbone.  No tensor cores: 9 (or 27) MACs per
// loaded element, so these are pure streaming kernels -- 128-bit channel-vector accesses, the
// input halo tile staged once in shared memory, grid sized in (tile x channel-group x image).
// Reference call sites: myolo/model.py:42-52 (conv_block) and 68-77 / 256-268
// (keras_applications _depthwise_conv_block: ZeroPad(1,1) + DepthwiseConv2D 3x3 VALID stride s).
#include <stdlib.h>
#include "common.cuh"

namespace myolo {

// ------------------------------------------------------------------------------------------
// depthwise 3x3 forward.  block = 256 threads = 32 pixel-slots x 8 channel-quads (32 channels).
// tile = 8x8 outputs; input halo tile (7*S+3)^2 x 32ch staged in smem with float4 loads.
// ------------------------------------------------------------------------------------------
// Fusions (BN statistics in the producer's epilogue, BN apply in the consumer's load):
//   INBN : the input is the PRE-BN output of the producing layer; a = act(gamma*(x-mean)*rs + beta) is applied while the
//          halo tile is staged (zero padding stays zero: ZeroPad comes after the activation), so the producer's post-BN
//          activation is never written to HBM;
//   STATS: per-channel sum / sum of squares of THIS layer's output, reduced per block and added to the fp64 workspace of
//          the BN family (bn.cu); the block that draws the last ticket of its 32-channel group turns them into the batch
//          mean / biased variance and zeroes the workspace again (same contract as colreduce_kernel<3>).
struct DwBn {
  const float* mean;     // INBN: statistics / affine terms of the producing layer's BN
  const float* var;
  const float* gamma;
  const float* beta;
  float eps;
  int act;
};
struct DwStats {
  double* sums;          // [2][C]: zero before the launch, zero after it
  int* ticket;           // one counter per 32-channel group
  float* mean;           // out: batch mean / biased variance of y
  float* var;
  double inv_count;      // 1 / (B*Ho*Wo)
  int C;
};

// a = act(x * sc + sh) with sc = gamma * rs, sh = beta - mean * sc (one FMA + the clamp per element: these kernels are
// instruction-bound, and the filter-gradient kernel meets every input nine times)
__device__ __forceinline__ float4 dw_in_bn(float4 v, const float4& sc, const float4& sh, int act) {
  v.x = apply_act(fmaf(v.x, sc.x, sh.x), act);
  v.y = apply_act(fmaf(v.y, sc.y, sh.y), act);
  v.z = apply_act(fmaf(v.z, sc.z, sh.z), act);
  v.w = apply_act(fmaf(v.w, sc.w, sh.w), act);
  return v;
}
__device__ __forceinline__ void dw_bn_consts(const DwBn& bn, int c, float4& sc, float4& sh) {
  const float4 mu = __ldg(reinterpret_cast<const float4*>(bn.mean + c));
  const float4 vv = __ldg(reinterpret_cast<const float4*>(bn.var + c));
  const float4 ga = __ldg(reinterpret_cast<const float4*>(bn.gamma + c));
  const float4 be = __ldg(reinterpret_cast<const float4*>(bn.beta + c));
  sc = make_float4(ga.x / sqrtf(vv.x + bn.eps), ga.y / sqrtf(vv.y + bn.eps), ga.z / sqrtf(vv.z + bn.eps), ga.w / sqrtf(vv.w + bn.eps));
  sh = make_float4(fmaf(-mu.x, sc.x, be.x), fmaf(-mu.y, sc.y, be.y), fmaf(-mu.z, sc.z, be.z), fmaf(-mu.w, sc.w, be.w));
}

template <int S, bool INBN, bool STATS>
__global__ void __launch_bounds__(256) dw_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                     float* __restrict__ y, int H, int W, int C, int Ho, int Wo,
                                                     int ntx, long long xsn, long long xsh, DwBn bn, DwStats st) {
  pdl_entry();
  constexpr int IT = 7 * S + 3;
  constexpr int kTile = IT * IT * 32, kRed = 2 * 32 * 33;
  __shared__ __align__(16) float tile[kTile > kRed ? kTile : kRed];
  __shared__ int s_last;
  const int tid = threadIdx.x;
  const int cq = tid & 7, pg = tid >> 3;
  const int c0 = blockIdx.y * 32;
  const int b = blockIdx.z;
  const int oy0 = (blockIdx.x / ntx) * 8, ox0 = (blockIdx.x % ntx) * 8;
  const int iy0 = oy0 * S - 1, ix0 = ox0 * S - 1;
  const float* xb = x + (size_t)b * xsn + c0;
  float4 bsc = make_float4(0.f, 0.f, 0.f, 0.f), bsh = bsc;
  if (INBN) dw_bn_consts(bn, c0 + cq * 4, bsc, bsh);     // the staging loop below keeps (i & 7) == cq: one channel quad per thread
  for (int i = tid; i < IT * IT * 8; i += 256) {
    const int pix = i >> 3, q = i & 7;
    const int gy = iy0 + pix / IT, gx = ix0 + pix % IT;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
      v = __ldg(reinterpret_cast<const float4*>(xb + (size_t)gy * xsh + (size_t)gx * C + q * 4));
      if (INBN) v = dw_in_bn(v, bsc, bsh, bn.act);
    }
    *reinterpret_cast<float4*>(&tile[pix * 32 + q * 4]) = v;
  }
  float4 wr[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) wr[k] = __ldg(reinterpret_cast<const float4*>(w + (size_t)k * C + c0 + cq * 4));
  __syncthreads();
  float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int o = pg + 32 * j;
    const int oy = o >> 3, ox = o & 7;
    if (oy0 + oy >= Ho || ox0 + ox >= Wo) continue;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float4 v = *reinterpret_cast<const float4*>(&tile[((oy * S + ky) * IT + ox * S + kx) * 32 + cq * 4]);
        const float4 ww = wr[ky * 3 + kx];
        acc.x = fmaf(v.x, ww.x, acc.x);
        acc.y = fmaf(v.y, ww.y, acc.y);
        acc.z = fmaf(v.z, ww.z, acc.z);
        acc.w = fmaf(v.w, ww.w, acc.w);
      }
    *reinterpret_cast<float4*>(y + (((size_t)b * Ho + oy0 + oy) * Wo + ox0 + ox) * C + c0 + cq * 4) = acc;
    if (STATS) {
      s0.x += acc.x; s0.y += acc.y; s0.z += acc.z; s0.w += acc.w;
      s1.x = fmaf(acc.x, acc.x, s1.x); s1.y = fmaf(acc.y, acc.y, s1.y); s1.z = fmaf(acc.z, acc.z, s1.z); s1.w = fmaf(acc.w, acc.w, s1.w);
    }
  }
  if (!STATS) return;
  __syncthreads();                      // every thread is done reading the halo tile: reuse it for the reduction
  float (*red)[32][33] = reinterpret_cast<float (*)[32][33]>(tile);
  red[0][cq * 4 + 0][pg] = s0.x; red[0][cq * 4 + 1][pg] = s0.y; red[0][cq * 4 + 2][pg] = s0.z; red[0][cq * 4 + 3][pg] = s0.w;
  red[1][cq * 4 + 0][pg] = s1.x; red[1][cq * 4 + 1][pg] = s1.y; red[1][cq * 4 + 2][pg] = s1.z; red[1][cq * 4 + 3][pg] = s1.w;
  __syncthreads();
  if (tid < 64) {
    const int which = tid >> 5, c = tid & 31;
    double s = 0.0;
#pragma unroll 8
    for (int j = 0; j < 32; ++j) s += (double)red[which][c][j];
    atomicAdd(st.sums + (size_t)which * st.C + c0 + c, s);
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(st.ticket + blockIdx.y, 1) == (int)(gridDim.x * gridDim.z) - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (tid < 32) {
    const int c = c0 + tid;
    const double S0 = *(volatile double*)(st.sums + c), S1 = *(volatile double*)(st.sums + st.C + c);
    const double m = S0 * st.inv_count;
    const double vv = S1 * st.inv_count - m * m;
    st.mean[c] = (float)m;
    st.var[c] = (float)(vv > 0.0 ? vv : 0.0);
    st.sums[c] = 0.0;
    st.sums[st.C + c] = 0.0;
  }
  if (tid == 0) st.ticket[blockIdx.y] = 0;
}

// ------------------------------------------------------------------------------------------
// Strip kernels (round 2).  The tile kernels above stage an 8x8 halo tile per 32 channels in shared memory: load,
// barrier, compute, store, with ~25 KB in flight per SM -- measured 0.2-0.36 of the HBM peak even on the 100 MB layers,
// bound by latency and instruction issue.  Here a thread owns one channel quad of one output COLUMN (b, ox) and walks
// a strip of TY output rows with the three input columns it needs in registers: every input row costs three 128-bit
// loads (two of them shared with the neighbouring threads through L1), no shared memory, no barrier, ~30 independent
// loads in flight per thread.  A warp covers 4 adjacent columns x 32 channels = 512 contiguous bytes of an NHWC row.
// ------------------------------------------------------------------------------------------
// a += v * w on a channel quad as two packed fp32 pairs (Blackwell FFMA2: each half rounded like a scalar fmaf)
__device__ __forceinline__ void fma4(float4& a, const float4& v, const float4& w) {
#ifdef MYOLO_DW_SCALAR_FMA
  a.x = fmaf(v.x, w.x, a.x);
  a.y = fmaf(v.y, w.y, a.y);
  a.z = fmaf(v.z, w.z, a.z);
  a.w = fmaf(v.w, w.w, a.w);
#else
  asm("{\n\t.reg .b64 a0, a1, v0, v1, w0, w1;\n\t"
      "mov.b64 a0, {%0, %1};\n\tmov.b64 a1, {%2, %3};\n\t"
      "mov.b64 v0, {%4, %5};\n\tmov.b64 v1, {%6, %7};\n\t"
      "mov.b64 w0, {%8, %9};\n\tmov.b64 w1, {%10, %11};\n\t"
      "fma.rn.f32x2 a0, v0, w0, a0;\n\tfma.rn.f32x2 a1, v1, w1, a1;\n\t"
      "mov.b64 {%0, %1}, a0;\n\tmov.b64 {%2, %3}, a1;\n\t}"
      : "+f"(a.x), "+f"(a.y), "+f"(a.z), "+f"(a.w)
      : "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "f"(w.x), "f"(w.y), "f"(w.z), "f"(w.w));
#endif
}

// forward (and, with FLIP, the stride-1 data gradient: the same correlation with the kernel rotated by 180 degrees)
#ifndef MYOLO_DW_MINBLK
#define MYOLO_DW_MINBLK 2
#endif
#ifndef MYOLO_DW_TY1
#define MYOLO_DW_TY1 8
#endif
template <int S, int TY, bool INBN, bool STATS, bool FLIP>
__global__ void __launch_bounds__(256, MYOLO_DW_MINBLK) dw_strip_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       float* __restrict__ y, int B, int H, int W, int C, int Ho, int Wo,
                                                       long long xsn, long long xsh, DwBn bn, DwStats st) {
  pdl_entry();
  __shared__ float red[STATS ? 2 * 32 * 33 : 1];
  __shared__ int s_last;
  const int tid = threadIdx.x;
  const int cq = tid & 7, pg = tid >> 3;
  const int c = blockIdx.y * 32 + cq * 4;
  const int col = blockIdx.x * 32 + pg;
  const bool active = col < B * Wo;
  const int b = active ? col / Wo : 0, ox = active ? col - (col / Wo) * Wo : 0;
  const int oy0 = blockIdx.z * TY;
  float4 wr[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) wr[k] = __ldg(reinterpret_cast<const float4*>(w + (size_t)(FLIP ? 8 - k : k) * C + c));
  float4 bsc = make_float4(0.f, 0.f, 0.f, 0.f), bsh = bsc;
  if (INBN) dw_bn_consts(bn, c, bsc, bsh);
  float4 acc[TY];
#pragma unroll
  for (int t = 0; t < TY; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* xb = x + (size_t)b * xsn + c;
  const int ix0 = ox * S - 1;
  constexpr int ROWS = S * (TY - 1) + 3;       // input rows a strip touches
  // every load of the strip is issued before the first FMA (ROWS x 3 quads in registers): the kernel lives on
  // memory-level parallelism, not on occupancy (measured on B200, dw1: 38.9 us with the loads interleaved row by row at
  // 102 registers, 30.7 us once the compiler was given 128 registers and hoisted them)
  float4 vin[ROWS][3];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const int iy = oy0 * S - 1 + r;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int ix = ix0 + j;
      vin[r][j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (active && iy >= 0 && iy < H && ix >= 0 && ix < W)
        vin[r][j] = __ldg(reinterpret_cast<const float4*>(xb + (size_t)iy * xsh + (size_t)ix * C));
    }
  }
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const int iy = oy0 * S - 1 + r;
    float4 v[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int ix = ix0 + j;
      v[j] = vin[r][j];
      if (INBN && active && iy >= 0 && iy < H && ix >= 0 && ix < W) v[j] = dw_in_bn(v[j], bsc, bsh, bn.act);
    }
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      // input row r is tap row ky of output row t with S*t + ky == r
      if ((r - ky) >= 0 && (r - ky) % S == 0 && (r - ky) / S < TY) {
        const int t = (r - ky) / S;
        fma4(acc[t], v[0], wr[ky * 3 + 0]);
        fma4(acc[t], v[1], wr[ky * 3 + 1]);
        fma4(acc[t], v[2], wr[ky * 3 + 2]);
      }
    }
  }
  float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
#pragma unroll
  for (int t = 0; t < TY; ++t) {
    const int oy = oy0 + t;
    if (active && oy < Ho) {
      *reinterpret_cast<float4*>(y + (((size_t)b * Ho + oy) * Wo + ox) * C + c) = acc[t];
      if (STATS) {
        s0.x += acc[t].x; s0.y += acc[t].y; s0.z += acc[t].z; s0.w += acc[t].w;
        s1.x = fmaf(acc[t].x, acc[t].x, s1.x); s1.y = fmaf(acc[t].y, acc[t].y, s1.y);
        s1.z = fmaf(acc[t].z, acc[t].z, s1.z); s1.w = fmaf(acc[t].w, acc[t].w, s1.w);
      }
    }
  }
  if (!STATS) return;
  float (*rd)[32][33] = reinterpret_cast<float (*)[32][33]>(red);
  rd[0][cq * 4 + 0][pg] = s0.x; rd[0][cq * 4 + 1][pg] = s0.y; rd[0][cq * 4 + 2][pg] = s0.z; rd[0][cq * 4 + 3][pg] = s0.w;
  rd[1][cq * 4 + 0][pg] = s1.x; rd[1][cq * 4 + 1][pg] = s1.y; rd[1][cq * 4 + 2][pg] = s1.z; rd[1][cq * 4 + 3][pg] = s1.w;
  __syncthreads();
  if (tid < 64) {
    const int which = tid >> 5, cc = tid & 31;
    double sum = 0.0;
#pragma unroll 8
    for (int j = 0; j < 32; ++j) sum += (double)rd[which][cc][j];
    atomicAdd(st.sums + (size_t)which * st.C + blockIdx.y * 32 + cc, sum);
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(st.ticket + blockIdx.y, 1) == (int)(gridDim.x * gridDim.z) - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (tid < 32) {
    const int cc = blockIdx.y * 32 + tid;
    const double S0 = *(volatile double*)(st.sums + cc), S1 = *(volatile double*)(st.sums + st.C + cc);
    const double m = S0 * st.inv_count;
    const double vv = S1 * st.inv_count - m * m;
    st.mean[cc] = (float)m;
    st.var[cc] = (float)(vv > 0.0 ? vv : 0.0);
    st.sums[cc] = 0.0;
    st.sums[st.C + cc] = 0.0;
  }
  if (tid == 0) st.ticket[blockIdx.y] = 0;
}

// filter gradient: thread = channel quad x output column (b, ox) x a strip of `rows` output rows.  Stride 1 keeps the
// 3x3 input window in registers and loads one new input row (3 quads) + one dy quad per output row; stride 2 loads two.
template <int S, bool INBN>
__global__ void __launch_bounds__(256) dw_bwd_filter_strip_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                  float* __restrict__ dw, int B, int H, int W, int C, int Ho,
                                                                  int Wo, int rows, long long xsn, long long xsh, DwBn bn) {
  pdl_entry();
  __shared__ float red[9 * 32 * 33];
  const int tid = threadIdx.x;
  const int cq = tid & 7, pg = tid >> 3;
  const int c = blockIdx.y * 32 + cq * 4;
  const int col = blockIdx.x * 32 + pg;
  const bool active = col < B * Wo;
  const int b = active ? col / Wo : 0, ox = active ? col - (col / Wo) * Wo : 0;
  const int oy0 = blockIdx.z * rows, oy1 = min(Ho, oy0 + rows);
  float4 bsc = make_float4(0.f, 0.f, 0.f, 0.f), bsh = bsc;
  if (INBN) dw_bn_consts(bn, c, bsc, bsh);
  float4 acc[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* xb = x + (size_t)b * xsn + c;
  const int ix0 = ox * S - 1;
  auto load_row = [&](int iy, float4 (&v)[3]) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int ix = ix0 + j;
      v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (active && iy >= 0 && iy < H && ix >= 0 && ix < W) {
        v[j] = __ldg(reinterpret_cast<const float4*>(xb + (size_t)iy * xsh + (size_t)ix * C));
        if (INBN) v[j] = dw_in_bn(v[j], bsc, bsh, bn.act);
      }
    }
  };
  // RB output rows per trip: the S*RB new input rows and the RB dy quads of a trip are loaded before its first FMA
  // (memory-level parallelism, see dw_strip_kernel); the 3-S input rows the next trip shares are carried in registers.
  constexpr int RB = S == 1 ? 4 : 2;
  constexpr int XR = S * RB + (3 - S);          // input rows a trip reads: S*RB new ones + the carried ones
  float4 xr[XR][3];
  if (oy0 < oy1) {
#pragma unroll
    for (int r = 0; r < 3 - S; ++r) load_row(oy0 * S - 1 + r, xr[r]);
  }
  for (int oy = oy0; oy < oy1; oy += RB) {
#pragma unroll
    for (int r = 3 - S; r < XR; ++r) load_row(oy * S - 1 + r, xr[r]);
    float4 g[RB];
#pragma unroll
    for (int t = 0; t < RB; ++t) {
      g[t] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (active && oy + t < oy1) g[t] = __ldg(reinterpret_cast<const float4*>(dy + (((size_t)b * Ho + oy + t) * Wo + ox) * C + c));
    }
#pragma unroll
    for (int t = 0; t < RB; ++t)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int j = 0; j < 3; ++j) fma4(acc[ky * 3 + j], xr[S * t + ky][j], g[t]);
#pragma unroll
    for (int r = 0; r < 3 - S; ++r)
#pragma unroll
      for (int j = 0; j < 3; ++j) xr[r][j] = xr[S * RB + r][j];
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    float* r = &red[(k * 32 + cq * 4) * 33 + pg];
    r[0] = acc[k].x;
    r[33] = acc[k].y;
    r[66] = acc[k].z;
    r[99] = acc[k].w;
  }
  __syncthreads();
  for (int i = tid; i < 9 * 32; i += 256) {
    const float* r = &red[i * 33];
    float sum = 0.f;
#pragma unroll 8
    for (int j = 0; j < 32; ++j) sum += r[j];
    const int k = i / 32, cc = i % 32;
    if (sum != 0.f) atomicAdd(dw + (size_t)k * C + blockIdx.y * 32 + cc, sum);
  }
}

// depthwise backward w.r.t. input: gather form, one thread = one input pixel x 4 channels.
template <int S>
__global__ void __launch_bounds__(256) dw_bwd_data_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                          float* __restrict__ dx, int B, int H, int W, int C, int Ho,
                                                          int Wo) {
  pdl_entry();
  const unsigned C4 = C >> 2;
  const unsigned total = (unsigned)B * H * W * C4;   // < 2^31 (checked by the caller): 32-bit index math
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int q = (int)(i % C4);
    unsigned p = i / C4;
    const int ix = (int)(p % (unsigned)W);
    p /= (unsigned)W;
    const int iy = (int)(p % (unsigned)H);
    const int b = (int)(p / (unsigned)H);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int ty = iy + 1 - ky;
      if (ty < 0 || (ty % S) != 0) continue;
      const int oy = ty / S;
      if (oy >= Ho) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int tx = ix + 1 - kx;
        if (tx < 0 || (tx % S) != 0) continue;
        const int ox = tx / S;
        if (ox >= Wo) continue;
        const float4 g = __ldg(reinterpret_cast<const float4*>(dy + (((size_t)b * Ho + oy) * Wo + ox) * C + q * 4));
        const float4 ww = __ldg(reinterpret_cast<const float4*>(w + (size_t)(ky * 3 + kx) * C + q * 4));
        acc.x = fmaf(g.x, ww.x, acc.x);
        acc.y = fmaf(g.y, ww.y, acc.y);
        acc.z = fmaf(g.z, ww.z, acc.z);
        acc.w = fmaf(g.w, ww.w, acc.w);
      }
    }
    *reinterpret_cast<float4*>(dx + (((size_t)b * H + iy) * W + ix) * C + q * 4) = acc;
  }
}

// depthwise backward w.r.t. filter: per-block register accumulation over a pixel chunk, smem
// reduction across the 32 pixel slots, one atomicAdd per (tap, channel) per block.
// INBN: x is the producing layer's PRE-BN output; its BN + activation is applied on load (see dw_fwd_kernel)
template <int S, bool INBN>
__global__ void __launch_bounds__(256) dw_bwd_filter_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                            float* __restrict__ dw, int B, int H, int W, int C, int Ho,
                                                            int Wo, long long chunk, long long xsn, long long xsh, DwBn bn) {
  pdl_entry();
  __shared__ float red[9 * 32 * 33];
  const int tid = threadIdx.x;
  const int cq = tid & 7, pg = tid >> 3;
  const int c0 = blockIdx.y * 32 + cq * 4;
  const long long total = (long long)B * Ho * Wo;
  const long long p0 = blockIdx.x * chunk;
  const long long p1 = min(total, p0 + chunk);
  float4 acc[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 bsc = make_float4(0.f, 0.f, 0.f, 0.f), bsh = bsc;
  if (INBN) dw_bn_consts(bn, c0, bsc, bsh);
  for (long long p = p0 + pg; p < p1; p += 32) {
    const unsigned pu = (unsigned)p;
    const int ox = (int)(pu % (unsigned)Wo);
    const unsigned t = pu / (unsigned)Wo;
    const int oy = (int)(t % (unsigned)Ho);
    const int b = (int)(t / (unsigned)Ho);
    const float4 g = __ldg(reinterpret_cast<const float4*>(dy + (size_t)p * C + c0));
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * S + ky - 1;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = ox * S + kx - 1;
        if (ix < 0 || ix >= W) continue;
        float4 v = __ldg(reinterpret_cast<const float4*>(x + (size_t)b * xsn + (size_t)iy * xsh + (size_t)ix * C + c0));
        if (INBN) v = dw_in_bn(v, bsc, bsh, bn.act);
        float4& a = acc[ky * 3 + kx];
        a.x = fmaf(v.x, g.x, a.x);
        a.y = fmaf(v.y, g.y, a.y);
        a.z = fmaf(v.z, g.z, a.z);
        a.w = fmaf(v.w, g.w, a.w);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    float* r = &red[(k * 32 + cq * 4) * 33 + pg];
    r[0] = acc[k].x;
    r[33] = acc[k].y;
    r[66] = acc[k].z;
    r[99] = acc[k].w;
  }
  __syncthreads();
  for (int i = tid; i < 9 * 32; i += 256) {
    const float* r = &red[i * 33];
    float s = 0.f;
#pragma unroll 8
    for (int j = 0; j < 32; ++j) s += r[j];
    const int k = i / 32, c = i % 32;
    atomicAdd(dw + (size_t)k * C + blockIdx.y * 32 + c, s);
  }
}

// ------------------------------------------------------------------------------------------
// stem conv 3x3 s2 pad 1, Cin=3 -> Cout (K=27: too thin for MMA, HBM/L1 bound direct conv).
// thread = one output pixel x 8 output channels; 864-float filter in smem.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        float* __restrict__ y, int B, int S, int Cout) {
  pdl_entry();
  extern __shared__ __align__(16) float ws[];  // [27][Cout]
  for (int i = threadIdx.x; i < 27 * Cout; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const int So = S / 2;
  const int octs = Cout >> 3;
  const unsigned total = (unsigned)B * So * So * octs;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int o8 = (int)(i % (unsigned)octs);
    unsigned p = i / (unsigned)octs;
    const int ox = (int)(p % (unsigned)So);
    p /= (unsigned)So;
    const int oy = (int)(p % (unsigned)So);
    const int b = (int)(p / (unsigned)So);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * 2 + ky - 1;
      if (iy < 0 || iy >= S) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = ox * 2 + kx - 1;
        if (ix < 0 || ix >= S) continue;
        const float* px = x + (((size_t)b * S + iy) * S + ix) * 3;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
          const float v = __ldg(px + ci);
          const float* wp = ws + ((ky * 3 + kx) * 3 + ci) * Cout + o8 * 8;
          const float4 w0 = *reinterpret_cast<const float4*>(wp);
          const float4 w1 = *reinterpret_cast<const float4*>(wp + 4);
          acc[0] = fmaf(v, w0.x, acc[0]);
          acc[1] = fmaf(v, w0.y, acc[1]);
          acc[2] = fmaf(v, w0.z, acc[2]);
          acc[3] = fmaf(v, w0.w, acc[3]);
          acc[4] = fmaf(v, w1.x, acc[4]);
          acc[5] = fmaf(v, w1.y, acc[5]);
          acc[6] = fmaf(v, w1.z, acc[6]);
          acc[7] = fmaf(v, w1.w, acc[7]);
        }
      }
    }
    float* py = y + (((size_t)b * So + oy) * So + ox) * Cout + o8 * 8;
    *reinterpret_cast<float4*>(py) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *reinterpret_cast<float4*>(py + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}

// stem conv wgrad (Cout == 32).  lane = output pixel, warp w of the block = output channels 4w .. 4w+3, 27 x 4 accumulators
// per thread: one pixel costs a thread 27 scalar loads of x (neighbouring lanes 24 bytes apart) and one 16-byte load of
// dy for 108 FMAs (54 packed FFMA2).  The first version (lane = output channel, one pixel per warp trip) issued 28 loads
// and three 64-bit div / mod per 27 FMAs and took 155 us at the very end of the backward pass, after the last kernel it
// could overlap with; the sums are reduced across the 32 pixels of a warp by shuffles at the end, one atomic per
// (block, tap, channel).
__global__ void __launch_bounds__(256) conv1_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                          float* __restrict__ dw, int B, int S, unsigned chunks_per_block) {
  pdl_entry();
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const unsigned So = (unsigned)S / 2u;
  const unsigned total = (unsigned)B * So * So;          // < 2^31 (checked on the host)
  const unsigned c_beg = blockIdx.x * chunks_per_block, c_end = c_beg + chunks_per_block;
  float acc[27][4];
#pragma unroll
  for (int t = 0; t < 27; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;
  // every load of a trip (TWO chunks of 32 pixels: 54 x values and two dy quads per thread) is issued before the first
  // FMA: with one block of eight warps per SM the kernel lives on loads in flight (one chunk per trip: 129 us)
  auto load_chunk = [&](unsigned ch, float (&xv)[27], float4& g) {
    const unsigned p = ch * 32u + (unsigned)lane;
    const bool live = ch < c_end && p < total;
    const unsigned pp = live ? p : 0u;
    const unsigned ox = pp % So, t2 = pp / So;
    const unsigned oy = t2 % So, b = t2 / So;
    g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) g = __ldg(reinterpret_cast<const float4*>(dy + (size_t)p * 32) + wp);
    const float* xb = x + (size_t)b * S * S * 3;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = (int)oy * 2 + ky - 1;
      const bool vy = live && iy >= 0 && iy < S;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = (int)ox * 2 + kx - 1;
        const bool v = vy && ix >= 0 && ix < S;
        const float* px = xb + ((size_t)(v ? iy : 0) * S + (v ? ix : 0)) * 3;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) xv[(ky * 3 + kx) * 3 + ci] = v ? __ldg(px + ci) : 0.f;
      }
    }
  };
  auto fma_chunk = [&](const float (&xv)[27], const float4& g) {
#pragma unroll
    for (int t = 0; t < 27; ++t) {
      acc[t][0] = fmaf(xv[t], g.x, acc[t][0]);
      acc[t][1] = fmaf(xv[t], g.y, acc[t][1]);
      acc[t][2] = fmaf(xv[t], g.z, acc[t][2]);
      acc[t][3] = fmaf(xv[t], g.w, acc[t][3]);
    }
  };
  for (unsigned ch = c_beg; ch < c_end && ch * 32u < total; ch += 2) {
    float xa[27], xb2[27];
    float4 ga, gb;
    load_chunk(ch, xa, ga);
    load_chunk(ch + 1, xb2, gb);
    fma_chunk(xa, ga);
    fma_chunk(xb2, gb);
  }
  // sum over the 32 pixels of the warp; lane 0 adds the block's share into dw[t][4*wp + c]
#pragma unroll
  for (int t = 0; t < 27; ++t) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float v = acc[t][c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && v != 0.f) atomicAdd(dw + t * 32 + wp * 4 + c, v);
    }
  }
}

}  // namespace myolo

using namespace myolo;

extern "C" int myolo_conv1_fwd(const float* x, const float* w, float* y, int B, int S, int Cout, myolo_stream stream) {
  MYOLO_CHECK_ARG(x && w && y && B > 0 && S > 0 && (S % 2) == 0 && Cout > 0 && (Cout % 8) == 0 && Cout <= 128);
  MYOLO_CHECK_ARG((long long)B * (S / 2) * (S / 2) * (Cout / 8) < (1LL << 31));
  const long long total = (long long)B * (S / 2) * (S / 2) * (Cout / 8);
  const int blocks = (int)min(ceil_div(total, 256), (long long)kNumSMs * 16);
  MYOLO_LAUNCH(conv1_fwd_kernel, blocks, 256, 27 * Cout * sizeof(float), as_stream(stream), x, w, y, B, S, Cout);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

extern "C" int myolo_conv1_wgrad(const float* x, const float* dy, float* dw, int B, int S, int Cout, myolo_stream stream) {
  MYOLO_CHECK_ARG(x && dy && dw && B > 0 && S > 0 && (S % 2) == 0);
  MYOLO_CHECK_ARG(Cout == 32);
  const long long total = (long long)B * (S / 2) * (S / 2);
  MYOLO_CHECK_ARG(total < (1LL << 31) - 64);
  const long long chunks = ceil_div(total, 32);                       // 32 output pixels per warp trip
  const int blocks = (int)max(1LL, min(chunks, (long long)kNumSMs));     // 173 registers: one block per SM
  const unsigned per_block = (unsigned)ceil_div(chunks, blocks);
  MYOLO_CUDA(cudaMemsetAsync(dw, 0, 27 * 32 * sizeof(float), as_stream(stream)));
  MYOLO_LAUNCH(conv1_wgrad_kernel, blocks, 256, 0, as_stream(stream), x, dy, dw, B, S, per_block);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

static inline int dw_out(int n, int s) { return (n + 2 - 3) / s + 1; }

static bool dw_view_ok(const myolo_view* v) {
  return v && v->p && v->n > 0 && v->h > 0 && v->w > 0 && v->c > 0 && (v->c % 32) == 0 && (v->sn % 4) == 0 && (v->sh % 4) == 0;
}

// workspace layout of the BN family (bn.cu): [0,16) doubles of int tickets | sums from double 16 on
static constexpr int kDwWsSums = 16, kDwWsMaxC = 1024;

extern "C" int myolo_dwconv3x3_fwd_bn(const myolo_view* xv, const float* w, float* y, int stride, const float* in_mean,
                                      const float* in_var, const float* in_gamma, const float* in_beta, float eps, int in_act,
                                      float* out_mean, float* out_var, double* ws, myolo_stream stream) {
  MYOLO_CHECK_ARG(dw_view_ok(xv) && w && y && (stride == 1 || stride == 2));
  const bool inbn = in_mean != nullptr, stats = out_mean != nullptr;
  MYOLO_CHECK_ARG(!inbn || (in_var && in_gamma && in_beta));
  MYOLO_CHECK_ARG(!stats || (out_var && ws && xv->c <= kDwWsMaxC));
  const float* x = xv->p;
  const int B = xv->n, H = xv->h, W = xv->w, C = xv->c;
  const int Ho = dw_out(H, stride), Wo = dw_out(W, stride);
  const int ntx = (Wo + 7) / 8, nty = (Ho + 7) / 8;
  dim3 grid(ntx * nty, C / 32, B);
  DwBn bn{in_mean, in_var, in_gamma, in_beta, eps, in_act};
  DwStats st{ws ? ws + kDwWsSums : nullptr, reinterpret_cast<int*>(ws), out_mean, out_var, 1.0 / ((double)B * Ho * Wo), C};
  cudaStream_t cs = as_stream(stream);
  static int use_tile = -1;
  if (use_tile < 0) {
    const char* e = getenv("MYOLO_DW_TILE");      // 1: the shared-memory tile kernels of round 1 (A/B switch)
    use_tile = (e && atoi(e)) ? 1 : 0;
  }
  if (use_tile) {
#define MYOLO_DW_FWD(S_, I_, T_) MYOLO_LAUNCH((dw_fwd_kernel<S_, I_, T_>), grid, 256, 0, cs, x, w, y, H, W, C, Ho, Wo, ntx, xv->sn, xv->sh, bn, st)
    if (stride == 1) {
      if (inbn && stats) MYOLO_DW_FWD(1, true, true);
      else if (inbn) MYOLO_DW_FWD(1, true, false);
      else if (stats) MYOLO_DW_FWD(1, false, true);
      else MYOLO_DW_FWD(1, false, false);
    } else {
      if (inbn && stats) MYOLO_DW_FWD(2, true, true);
      else if (inbn) MYOLO_DW_FWD(2, true, false);
      else if (stats) MYOLO_DW_FWD(2, false, true);
      else MYOLO_DW_FWD(2, false, false);
    }
#undef MYOLO_DW_FWD
  } else {
    constexpr int TY1 = MYOLO_DW_TY1, TY2 = 4;
    dim3 g1((unsigned)ceil_div((long long)B * Wo, 32), C / 32, (unsigned)ceil_div(Ho, stride == 1 ? TY1 : TY2));
#define MYOLO_DW_STRIP(S_, T_, I_, ST_) \
  MYOLO_LAUNCH((dw_strip_kernel<S_, T_, I_, ST_, false>), g1, 256, 0, cs, x, w, y, B, H, W, C, Ho, Wo, xv->sn, xv->sh, bn, st)
    if (stride == 1) {
      if (inbn && stats) MYOLO_DW_STRIP(1, TY1, true, true);
      else if (inbn) MYOLO_DW_STRIP(1, TY1, true, false);
      else if (stats) MYOLO_DW_STRIP(1, TY1, false, true);
      else MYOLO_DW_STRIP(1, TY1, false, false);
    } else {
      if (inbn && stats) MYOLO_DW_STRIP(2, TY2, true, true);
      else if (inbn) MYOLO_DW_STRIP(2, TY2, true, false);
      else if (stats) MYOLO_DW_STRIP(2, TY2, false, true);
      else MYOLO_DW_STRIP(2, TY2, false, false);
    }
#undef MYOLO_DW_STRIP
  }
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

extern "C" int myolo_dwconv3x3_fwd(const myolo_view* xv, const float* w, float* y, int stride, myolo_stream stream) {
  return myolo_dwconv3x3_fwd_bn(xv, w, y, stride, nullptr, nullptr, nullptr, nullptr, 0.f, 0, nullptr, nullptr, nullptr, stream);
}

extern "C" int myolo_dwconv3x3_bwd_data(const float* dy, const float* w, float* dx, int B, int H, int W, int C,
                                        int stride, myolo_stream stream) {
  MYOLO_CHECK_ARG(dy && w && dx && B > 0 && H > 0 && W > 0 && C > 0 && (C % 4) == 0 && (stride == 1 || stride == 2));
  MYOLO_CHECK_ARG((long long)B * H * W * (C / 4) < (1LL << 31));
  const int Ho = dw_out(H, stride), Wo = dw_out(W, stride);
  const long long total = (long long)B * H * W * (C / 4);
  const int blocks = (int)min(ceil_div(total, 256), (long long)kNumSMs * 32);
  static int use_old = -1;
  if (use_old < 0) {
    const char* e = getenv("MYOLO_DW_TILE");
    use_old = (e && atoi(e)) ? 1 : 0;
  }
  if (stride == 1 && !use_old && (C % 32) == 0) {
    // stride 1: dx = dy correlated with the kernel rotated by 180 degrees, same geometry -> the forward strip kernel
    constexpr int TY = MYOLO_DW_TY1;
    dim3 grid((unsigned)ceil_div((long long)B * W, 32), C / 32, (unsigned)ceil_div(H, TY));
    DwBn bn{};
    DwStats st{};
    MYOLO_LAUNCH((dw_strip_kernel<1, TY, false, false, true>), grid, 256, 0, as_stream(stream), dy, w, dx, B, H, W, C, H, W,
                                                                                     (long long)H * W * C, (long long)W * C, bn, st);
  } else if (stride == 1)
    MYOLO_LAUNCH(dw_bwd_data_kernel<1>, blocks, 256, 0, as_stream(stream), dy, w, dx, B, H, W, C, Ho, Wo);
  else
    MYOLO_LAUNCH(dw_bwd_data_kernel<2>, blocks, 256, 0, as_stream(stream), dy, w, dx, B, H, W, C, Ho, Wo);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

extern "C" int myolo_dwconv3x3_bwd_filter_bn(const myolo_view* xv, const float* dy, float* dw, int stride, const float* in_mean,
                                             const float* in_var, const float* in_gamma, const float* in_beta, float eps,
                                             int in_act, myolo_stream stream) {
  MYOLO_CHECK_ARG(dw_view_ok(xv) && dy && dw && (stride == 1 || stride == 2));
  const bool inbn = in_mean != nullptr;
  MYOLO_CHECK_ARG(!inbn || (in_var && in_gamma && in_beta));
  DwBn bn{in_mean, in_var, in_gamma, in_beta, eps, in_act};
  const float* x = xv->p;
  const int B = xv->n, H = xv->h, W = xv->w, C = xv->c;
  const int Ho = dw_out(H, stride), Wo = dw_out(W, stride);
  const long long total = (long long)B * Ho * Wo;
  const int cgroups = C / 32;
  long long nchunks = max(1LL, min(ceil_div(total, 64), (long long)(kNumSMs * 4) / cgroups + 1));
  const long long chunk = ceil_div(total, nchunks);
  nchunks = ceil_div(total, chunk);
  MYOLO_CUDA(cudaMemsetAsync(dw, 0, (size_t)9 * C * sizeof(float), as_stream(stream)));
  dim3 grid((unsigned)nchunks, cgroups);
  cudaStream_t cs = as_stream(stream);
  static int use_old = -1;
  if (use_old < 0) {
    const char* e = getenv("MYOLO_DW_TILE");
    use_old = (e && atoi(e)) ? 1 : 0;
  }
  if (!use_old) {
    // strips of output rows per thread: about four blocks per SM in total
    const long long coltiles = ceil_div((long long)B * Wo, 32);
    long long nstrips = max(1LL, min((long long)Ho, ceil_div((long long)kNumSMs * 4, coltiles * cgroups)));
    const int rows = (int)ceil_div(Ho, nstrips);
    nstrips = ceil_div(Ho, rows);
    dim3 g2((unsigned)coltiles, cgroups, (unsigned)nstrips);
    if (stride == 1) {
      if (inbn) MYOLO_LAUNCH((dw_bwd_filter_strip_kernel<1, true>), g2, 256, 0, cs, x, dy, dw, B, H, W, C, Ho, Wo, rows, xv->sn, xv->sh, bn);
      else MYOLO_LAUNCH((dw_bwd_filter_strip_kernel<1, false>), g2, 256, 0, cs, x, dy, dw, B, H, W, C, Ho, Wo, rows, xv->sn, xv->sh, bn);
    } else {
      if (inbn) MYOLO_LAUNCH((dw_bwd_filter_strip_kernel<2, true>), g2, 256, 0, cs, x, dy, dw, B, H, W, C, Ho, Wo, rows, xv->sn, xv->sh, bn);
      else MYOLO_LAUNCH((dw_bwd_filter_strip_kernel<2, false>), g2, 256, 0, cs, x, dy, dw, B, H, W, C, Ho, Wo, rows, xv->sn, xv->sh, bn);
    }
    MYOLO_CHECK_LAUNCH();
    return MYOLO_OK;
  }
  if (stride == 1) {
    if (inbn) MYOLO_LAUNCH((dw_bwd_filter_kernel<1, true>), grid, 256, 0, cs, x, dy, dw, B, H, W, C, Ho, Wo, chunk, xv->sn, xv->sh, bn);
    else MYOLO_LAUNCH((dw_bwd_filter_kernel<1, false>), grid, 256, 0, cs, x, dy, dw, B, H, W, C, Ho, Wo, chunk, xv->sn, xv->sh, bn);
  } else {
    if (inbn) MYOLO_LAUNCH((dw_bwd_filter_kernel<2, true>), grid, 256, 0, cs, x, dy, dw, B, H, W, C, Ho, Wo, chunk, xv->sn, xv->sh, bn);
    else MYOLO_LAUNCH((dw_bwd_filter_kernel<2, false>), grid, 256, 0, cs, x, dy, dw, B, H, W, C, Ho, Wo, chunk, xv->sn, xv->sh, bn);
  }
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

extern "C" int myolo_dwconv3x3_bwd_filter(const myolo_view* xv, const float* dy, float* dw, int stride,
                                          myolo_stream stream) {
  return myolo_dwconv3x3_bwd_filter_bn(xv, dy, dw, stride, nullptr, nullptr, nullptr, nullptr, 0.f, 0, stream);
}
