// roialign.cu -- PyramidROIAlign (K7): tf.image.crop_and_resize(feature, boxes, box_idx, (P,P),
// bilinear, extrapolation 0) forward and its image gradient (CropAndResizeGradImage).
// Reference: myolo/model.py:327-410 (crop at 385-387; every ROI maps to pyramid level 0, the
// final top-k re-sort is the identity).  TF 1.x kernel semantics restated in SURVEY.md Q3.
//
// HBM-bound gather: one warp per (roi, output row); per sample 4 coalesced channel-vector reads
// (C*4 bytes each, served from L2 after first touch: the whole feature map is B*F*F*C*4 bytes)
// and one C*4-byte streaming write.  Coordinates are computed with explicitly rounded fp32 ops in
// the reference's order so the sample positions are bit-identical to the CPU oracle.
#include <cuda_fp16.h>
#include "common.cuh"
#include "crop.cuh"

namespace myolo {

struct V {
  float* p;
  long long sn, sh;
  int n, h, w, c;
};
static inline V to_v(const myolo_view* v) { return V{v->p, v->sn, v->sh, v->n, v->h, v->w, v->c}; }

__device__ __forceinline__ uint32_t pack_half2_sat(float lo, float hi) {
  uint16_t a, b;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(a) : "f"(lo));
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(b) : "f"(hi));
  return (uint32_t)a | ((uint32_t)b << 16);
}

// HALF: the pooled values are (also) stored as IEEE half into `outh` (operand of the kind::f16 mask conv1);
// out.p may then be null.
// NQ > 0 (C == 128 * NQ): the feature columns a sample row touches are kept in registers and re-used by the next sample
// when it falls on the same columns -- 14 samples over a 3..10 pixel wide box read 28 column quads per row otherwise,
// most of them twice or more (the kernel was bound by these L1 / L2 gathers, not by its 0.4 GB of output).  The values
// and their arithmetic are unchanged, so the result stays bit-exact.  NQ == 0: any C, one load per sample and corner.
template <bool HALF, int NQ>
__global__ void __launch_bounds__(256)
roialign_fwd_kernel(V feat, const float* __restrict__ boxes, int n_roi, int rois_per_img, int pool, V out, int rnd, V outh) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int C4 = feat.c >> 2;
  const long long items = (long long)n_roi * pool;
  for (long long it = warp; it < items; it += nwarps) {
    const int r = (int)(it / pool), y = (int)(it % pool);
    const float4 bx = __ldg(reinterpret_cast<const float4*>(boxes) + r);  // consumed as (y1,x1,y2,x2)
    const int b = r / rois_per_img;
    const Sample sy = crop_coord(bx.x, bx.z, y, pool, feat.h);
    const float* fb = feat.p + (size_t)b * feat.sn;
    float* orow = out.p ? out.p + (size_t)r * out.sn + (size_t)y * out.sh : nullptr;
    uint16_t* hrow = HALF ? reinterpret_cast<uint16_t*>(outh.p) + (size_t)r * outh.sn + (size_t)y * outh.sh : nullptr;
    constexpr int NQR = NQ > 0 ? NQ : 1;
    float4 ct[2][NQR], cb[2][NQR];     // cached column pair: [0] = column c_lo, [1] = column c_hi; top / bottom feature row
    int c_lo = -1, c_hi = -1;
    for (int x = 0; x < pool; ++x) {
      const Sample sx = crop_coord(bx.y, bx.w, x, pool, feat.w);
      float4* op = reinterpret_cast<float4*>(orow + (size_t)x * out.c);
      uint2* hp = reinterpret_cast<uint2*>(hrow + (size_t)x * feat.c);
      if (!(sy.valid && sx.valid)) {
        for (int q = lane; q < C4; q += 32) {
          if (orow) op[q] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (HALF) hp[q] = make_uint2(0u, 0u);
        }
        continue;
      }
      const float4* rowt = reinterpret_cast<const float4*>(fb + (size_t)sy.lo * feat.sh);
      const float4* rowb = reinterpret_cast<const float4*>(fb + (size_t)sy.hi * feat.sh);
      if (NQ > 0) {
        // bring (sx.lo, sx.hi) into the two cache slots, loading only the columns that are not there yet
        auto load_col = [&](int slot, int col) {
#pragma unroll
          for (int qq = 0; qq < NQR; ++qq) {
            ct[slot][qq] = __ldg(rowt + (size_t)col * C4 + qq * 32 + lane);
            cb[slot][qq] = __ldg(rowb + (size_t)col * C4 + qq * 32 + lane);
          }
        };
        auto move_col = [&](int dst, int src) {
#pragma unroll
          for (int qq = 0; qq < NQR; ++qq) {
            ct[dst][qq] = ct[src][qq];
            cb[dst][qq] = cb[src][qq];
          }
        };
        if (sx.lo != c_lo) {
          if (sx.lo == c_hi) {
            if (sx.hi == c_lo) {          // swapped pair (a reversed box stepping back by one column)
#pragma unroll
              for (int qq = 0; qq < NQR; ++qq) {
                const float4 t0 = ct[0][qq], t1 = cb[0][qq];
                ct[0][qq] = ct[1][qq]; cb[0][qq] = cb[1][qq];
                ct[1][qq] = t0; cb[1][qq] = t1;
              }
              c_hi = c_lo;
            } else {
              move_col(0, 1);
              c_hi = -1;
            }
          } else if (sx.hi == c_lo && sx.hi != sx.lo) {
            move_col(1, 0);
            c_hi = c_lo;
            load_col(0, sx.lo);
          } else {
            load_col(0, sx.lo);
          }
          c_lo = sx.lo;
        }
        if (sx.hi != c_hi) {
          if (sx.hi == c_lo) move_col(1, 0);
          else load_col(1, sx.hi);
          c_hi = sx.hi;
        }
#pragma unroll
        for (int qq = 0; qq < NQR; ++qq) {
          const float4 a = ct[0][qq], bq = ct[1][qq], c = cb[0][qq], d = cb[1][qq];
          float4 o;
          o.x = lerp_rn(lerp_rn(a.x, bq.x, sx.lerp), lerp_rn(c.x, d.x, sx.lerp), sy.lerp);
          o.y = lerp_rn(lerp_rn(a.y, bq.y, sx.lerp), lerp_rn(c.y, d.y, sx.lerp), sy.lerp);
          o.z = lerp_rn(lerp_rn(a.z, bq.z, sx.lerp), lerp_rn(c.z, d.z, sx.lerp), sy.lerp);
          o.w = lerp_rn(lerp_rn(a.w, bq.w, sx.lerp), lerp_rn(c.w, d.w, sx.lerp), sy.lerp);
          if (rnd) o = make_float4(round_tf32(o.x), round_tf32(o.y), round_tf32(o.z), round_tf32(o.w));
          const int q = qq * 32 + lane;
          if (orow) op[q] = o;
          if (HALF) hp[q] = make_uint2(pack_half2_sat(o.x, o.y), pack_half2_sat(o.z, o.w));
        }
        continue;
      }
      const float4* tl = rowt + (size_t)sx.lo * C4;
      const float4* tr = rowt + (size_t)sx.hi * C4;
      const float4* bl = rowb + (size_t)sx.lo * C4;
      const float4* br = rowb + (size_t)sx.hi * C4;
      for (int q = lane; q < C4; q += 32) {
        const float4 a = __ldg(tl + q), bq = __ldg(tr + q), c = __ldg(bl + q), d = __ldg(br + q);
        float4 o;
        o.x = lerp_rn(lerp_rn(a.x, bq.x, sx.lerp), lerp_rn(c.x, d.x, sx.lerp), sy.lerp);
        o.y = lerp_rn(lerp_rn(a.y, bq.y, sx.lerp), lerp_rn(c.y, d.y, sx.lerp), sy.lerp);
        o.z = lerp_rn(lerp_rn(a.z, bq.z, sx.lerp), lerp_rn(c.z, d.z, sx.lerp), sy.lerp);
        o.w = lerp_rn(lerp_rn(a.w, bq.w, sx.lerp), lerp_rn(c.w, d.w, sx.lerp), sy.lerp);
        if (rnd) o = make_float4(round_tf32(o.x), round_tf32(o.y), round_tf32(o.z), round_tf32(o.w));
        if (orow) op[q] = o;
        if (HALF) hp[q] = make_uint2(pack_half2_sat(o.x, o.y), pack_half2_sat(o.z, o.w));
      }
    }
  }
}

__device__ __forceinline__ void red_add4(float* p, float4 v) {
  atomicAdd(reinterpret_cast<float4*>(p), v);  // sm_90+: 128-bit vector reduction
}

__global__ void __launch_bounds__(256)
roialign_bwd_kernel(V dout, const float* __restrict__ boxes, int n_roi, int rois_per_img, int pool, V dfeat) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int C4 = dfeat.c >> 2;
  const long long items = (long long)n_roi * pool;
  for (long long it = warp; it < items; it += nwarps) {
    const int r = (int)(it / pool), y = (int)(it % pool);
    const float4 bx = __ldg(reinterpret_cast<const float4*>(boxes) + r);
    const int b = r / rois_per_img;
    const Sample sy = crop_coord(bx.x, bx.z, y, pool, dfeat.h);
    if (!sy.valid) continue;
    float* fb = dfeat.p + (size_t)b * dfeat.sn;
    const float* grow = dout.p + (size_t)r * dout.sn + (size_t)y * dout.sh;
    for (int x = 0; x < pool; ++x) {
      const Sample sx = crop_coord(bx.y, bx.w, x, pool, dfeat.w);
      if (!sx.valid) continue;
      const float4* gp = reinterpret_cast<const float4*>(grow + (size_t)x * dout.c);
      float* tl = fb + (size_t)sy.lo * dfeat.sh + (size_t)sx.lo * dfeat.c;
      float* tr = fb + (size_t)sy.lo * dfeat.sh + (size_t)sx.hi * dfeat.c;
      float* bl = fb + (size_t)sy.hi * dfeat.sh + (size_t)sx.lo * dfeat.c;
      float* br = fb + (size_t)sy.hi * dfeat.sh + (size_t)sx.hi * dfeat.c;
      const float wt = 1.f - sy.lerp, wb = sy.lerp, wl = 1.f - sx.lerp, wr = sx.lerp;
      for (int q = lane; q < C4; q += 32) {
        const float4 g = gp[q];
        const float4 dt = make_float4(wt * g.x, wt * g.y, wt * g.z, wt * g.w);
        const float4 db = make_float4(wb * g.x, wb * g.y, wb * g.z, wb * g.w);
        red_add4(tl + q * 4, make_float4(wl * dt.x, wl * dt.y, wl * dt.z, wl * dt.w));
        red_add4(tr + q * 4, make_float4(wr * dt.x, wr * dt.y, wr * dt.z, wr * dt.w));
        red_add4(bl + q * 4, make_float4(wl * db.x, wl * db.y, wl * db.z, wl * db.w));
        red_add4(br + q * 4, make_float4(wr * db.x, wr * db.y, wr * db.z, wr * db.w));
      }
    }
  }
}

// Run-length form of the backward scatter: a warp walks the P samples of one (roi, output row) left to right and
// keeps the contributions to the current pair of feature columns (lo, lo+1) x (top, bottom row) in registers; they
// go to memory as vector reductions only when the sample position moves on to another column.  A 14-sample row over
// a 3..10 pixel wide box issues 1.4..4.7x fewer L2 atomics than one reduction per sample and corner (the atomics,
// not the 0.94 GB gradient read, bound the per-sample kernel: 0.58 ms).  NQ = C / 128 float4 per lane.
// GH: dout is an IEEE-half view holding the gradient times a loss scale; *in_scale (device scalar, nullable = 1) removes it
template <int NQ, bool GH = false>
__global__ void __launch_bounds__(256)
roialign_bwd_rl_kernel(V dout, const float* __restrict__ boxes, int n_roi, int rois_per_img, int pool, V dfeat,
                       const float* __restrict__ in_scale = nullptr) {
  const float gsc = (GH && in_scale) ? __ldg(in_scale) : 1.f;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const long long items = (long long)n_roi * pool;
  for (long long it = warp; it < items; it += nwarps) {
    const int r = (int)(it / pool), y = (int)(it % pool);
    const float4 bx = __ldg(reinterpret_cast<const float4*>(boxes) + r);
    const int b = r / rois_per_img;
    const Sample sy = crop_coord(bx.x, bx.z, y, pool, dfeat.h);
    if (!sy.valid) continue;
    float* fb = dfeat.p + (size_t)b * dfeat.sn;
    float* rowt = fb + (size_t)sy.lo * dfeat.sh;
    float* rowb = fb + (size_t)sy.hi * dfeat.sh;
    const float* grow = dout.p + (size_t)r * dout.sn + (size_t)y * dout.sh;
    const float wt = 1.f - sy.lerp, wb = sy.lerp;
    const bool two_rows = sy.hi != sy.lo;      // hi == lo: the bottom weight is exactly zero
    float4 at[2][NQ], ab[2][NQ];               // [column lo / lo+1][channel chunk], top and bottom row
    int col = -1;                               // feature column of at[0] / ab[0]; -1 = nothing pending
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int q = 0; q < NQ; ++q) at[c][q] = ab[c][q] = make_float4(0.f, 0.f, 0.f, 0.f);
    auto flush = [&](int c, int column) {
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const size_t off = (size_t)column * dfeat.c + (size_t)(q * 32 + lane) * 4;
        red_add4(rowt + off, at[c][q]);
        if (two_rows) red_add4(rowb + off, ab[c][q]);
        at[c][q] = ab[c][q] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    for (int x = 0; x < pool; ++x) {
      const Sample sx = crop_coord(bx.y, bx.w, x, pool, dfeat.w);
      if (!sx.valid) continue;
      if (col >= 0 && sx.lo != col) {
        if (sx.lo == col + 1) {                // moved on by one column: the old "lo+1" becomes "lo"
          flush(0, col);
#pragma unroll
          for (int q = 0; q < NQ; ++q) {
            at[0][q] = at[1][q]; ab[0][q] = ab[1][q];
            at[1][q] = ab[1][q] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        } else {
          flush(0, col);
          if (col + 1 < dfeat.w) flush(1, col + 1);
        }
      }
      col = sx.lo;
      const float wl = 1.f - sx.lerp, wr = sx.lerp;
      const float4* gp = reinterpret_cast<const float4*>(grow + (size_t)x * dout.c);
      const uint2* gph = reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(dout.p) + (size_t)r * dout.sn +
                                                        (size_t)y * dout.sh + (size_t)x * dout.c);
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        float4 g;
        if (GH) {
          const uint2 u = __ldg(gph + q * 32 + lane);
          g = make_float4(gsc * __half2float(__ushort_as_half((unsigned short)(u.x & 0xffffu))),
                          gsc * __half2float(__ushort_as_half((unsigned short)(u.x >> 16))),
                          gsc * __half2float(__ushort_as_half((unsigned short)(u.y & 0xffffu))),
                          gsc * __half2float(__ushort_as_half((unsigned short)(u.y >> 16))));
        } else {
          g = gp[q * 32 + lane];
        }
        const float4 dt = make_float4(wt * g.x, wt * g.y, wt * g.z, wt * g.w);
        const float4 db = make_float4(wb * g.x, wb * g.y, wb * g.z, wb * g.w);
        at[0][q].x += wl * dt.x; at[0][q].y += wl * dt.y; at[0][q].z += wl * dt.z; at[0][q].w += wl * dt.w;
        ab[0][q].x += wl * db.x; ab[0][q].y += wl * db.y; ab[0][q].z += wl * db.z; ab[0][q].w += wl * db.w;
        if (sx.hi != sx.lo) {                  // hi == lo: the right weight is exactly zero
          at[1][q].x += wr * dt.x; at[1][q].y += wr * dt.y; at[1][q].z += wr * dt.z; at[1][q].w += wr * dt.w;
          ab[1][q].x += wr * db.x; ab[1][q].y += wr * db.y; ab[1][q].z += wr * db.z; ab[1][q].w += wr * db.w;
        }
      }
    }
    if (col >= 0) {
      flush(0, col);
      if (col + 1 < dfeat.w) flush(1, col + 1);
    }
  }
}

static bool view_ok(const myolo_view* v) {
  return v && v->p && v->n > 0 && v->h > 0 && v->w > 0 && v->c > 0 && (v->c % 4) == 0 && (v->sn % 4) == 0 && (v->sh % 4) == 0;
}

}  // namespace myolo

using namespace myolo;

extern "C" int myolo_roialign_fwd(const myolo_view* feat, const float* boxes, int n_roi, int rois_per_img, int pool,
                                  const myolo_view* out, int round_tf32, myolo_stream stream) {
  MYOLO_CHECK_ARG(view_ok(feat) && view_ok(out) && boxes && n_roi > 0 && rois_per_img > 0 && pool > 0);
  MYOLO_CHECK_ARG(out->n == n_roi && out->h == pool && out->w == pool && out->c == feat->c);
  MYOLO_CHECK_ARG((n_roi + rois_per_img - 1) / rois_per_img <= feat->n);
  const long long items = (long long)n_roi * pool;
  const int blocks = (int)max(1LL, min(ceil_div(items, 8), (long long)kNumSMs * 8));
  if (feat->c == 256)
    roialign_fwd_kernel<false, 2><<<blocks, 256, 0, as_stream(stream)>>>(to_v(feat), boxes, n_roi, rois_per_img, pool, to_v(out), round_tf32, to_v(out));
  else if (feat->c == 128)
    roialign_fwd_kernel<false, 1><<<blocks, 256, 0, as_stream(stream)>>>(to_v(feat), boxes, n_roi, rois_per_img, pool, to_v(out), round_tf32, to_v(out));
  else
    roialign_fwd_kernel<false, 0><<<blocks, 256, 0, as_stream(stream)>>>(to_v(feat), boxes, n_roi, rois_per_img, pool, to_v(out), round_tf32, to_v(out));
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

extern "C" int myolo_roialign_fwd_h(const myolo_view* feat, const float* boxes, int n_roi, int rois_per_img, int pool,
                                    const myolo_view* out, const myolo_view* out_half, myolo_stream stream) {
  MYOLO_CHECK_ARG(view_ok(feat) && view_ok(out_half) && (!out || view_ok(out)) && boxes && n_roi > 0 && rois_per_img > 0 && pool > 0);
  MYOLO_CHECK_ARG(out_half->n == n_roi && out_half->h == pool && out_half->w == pool && out_half->c == feat->c);
  MYOLO_CHECK_ARG(!out || (out->n == n_roi && out->h == pool && out->w == pool && out->c == feat->c));
  MYOLO_CHECK_ARG((n_roi + rois_per_img - 1) / rois_per_img <= feat->n && ((uintptr_t)out_half->p & 7) == 0);
  const long long items = (long long)n_roi * pool;
  const int blocks = (int)max(1LL, min(ceil_div(items, 8), (long long)kNumSMs * 8));
  V vo = out ? to_v(out) : V{nullptr, 0, 0, n_roi, pool, pool, feat->c};
  if (feat->c == 256)
    roialign_fwd_kernel<true, 2><<<blocks, 256, 0, as_stream(stream)>>>(to_v(feat), boxes, n_roi, rois_per_img, pool, vo, 0, to_v(out_half));
  else if (feat->c == 128)
    roialign_fwd_kernel<true, 1><<<blocks, 256, 0, as_stream(stream)>>>(to_v(feat), boxes, n_roi, rois_per_img, pool, vo, 0, to_v(out_half));
  else
    roialign_fwd_kernel<true, 0><<<blocks, 256, 0, as_stream(stream)>>>(to_v(feat), boxes, n_roi, rois_per_img, pool, vo, 0, to_v(out_half));
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

extern "C" int myolo_roialign_bwd(const myolo_view* dout, const float* boxes, int n_roi, int rois_per_img, int pool,
                                  const myolo_view* dfeat, myolo_stream stream) {
  MYOLO_CHECK_ARG(view_ok(dfeat) && view_ok(dout) && boxes && n_roi > 0 && rois_per_img > 0 && pool > 0);
  MYOLO_CHECK_ARG(dout->n == n_roi && dout->h == pool && dout->w == pool && dout->c == dfeat->c);
  MYOLO_CHECK_ARG((n_roi + rois_per_img - 1) / rois_per_img <= dfeat->n);
  const long long items = (long long)n_roi * pool;
  const int blocks = (int)max(1LL, min(ceil_div(items, 8), (long long)kNumSMs * 8));
  if (dfeat->c == 128)
    roialign_bwd_rl_kernel<1><<<blocks, 256, 0, as_stream(stream)>>>(to_v(dout), boxes, n_roi, rois_per_img, pool, to_v(dfeat));
  else if (dfeat->c == 256)
    roialign_bwd_rl_kernel<2><<<blocks, 256, 0, as_stream(stream)>>>(to_v(dout), boxes, n_roi, rois_per_img, pool, to_v(dfeat));
  else
    roialign_bwd_kernel<<<blocks, 256, 0, as_stream(stream)>>>(to_v(dout), boxes, n_roi, rois_per_img, pool, to_v(dfeat));
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

// myolo_roialign_bwd on a HALF, loss-scaled gradient (the data gradient of myolo_mask_conv1 in h16 mode); *in_scale
// (device scalar, nullable) is multiplied into every value read, dfeat stays fp32 and unscaled.  C = 128 or 256.
extern "C" int myolo_roialign_bwd_h(const myolo_view* dout_half, const float* boxes, int n_roi, int rois_per_img, int pool,
                                    const myolo_view* dfeat, const float* in_scale, myolo_stream stream) {
  MYOLO_CHECK_ARG(view_ok(dfeat) && view_ok(dout_half) && boxes && n_roi > 0 && rois_per_img > 0 && pool > 0);
  MYOLO_CHECK_ARG(dout_half->n == n_roi && dout_half->h == pool && dout_half->w == pool && dout_half->c == dfeat->c);
  MYOLO_CHECK_ARG((n_roi + rois_per_img - 1) / rois_per_img <= dfeat->n && ((uintptr_t)dout_half->p & 7) == 0);
  MYOLO_CHECK_ARG(dfeat->c == 128 || dfeat->c == 256);
  const long long items = (long long)n_roi * pool;
  const int blocks = (int)max(1LL, min(ceil_div(items, 8), (long long)kNumSMs * 8));
  if (dfeat->c == 128)
    roialign_bwd_rl_kernel<1, true><<<blocks, 256, 0, as_stream(stream)>>>(to_v(dout_half), boxes, n_roi, rois_per_img, pool, to_v(dfeat), in_scale);
  else
    roialign_bwd_rl_kernel<2, true><<<blocks, 256, 0, as_stream(stream)>>>(to_v(dout_half), boxes, n_roi, rois_per_img, pool, to_v(dfeat), in_scale);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}
