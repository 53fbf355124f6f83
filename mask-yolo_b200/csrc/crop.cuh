// crop.cuh -- tf.image.crop_and_resize sampling arithmetic (TF 1.x CPU kernel),
// shared by ROIAlign (K7) and the mask-target crop (K13).  Every operation is an explicitly
// rounded fp32 op in the reference's evaluation order (no FMA contraction), so sample positions,
// interpolation weights and the half-to-even rounding of mask targets are bit-identical to the
// CPU oracle.
#pragma once
#include <cuda_runtime.h>

namespace myolo {

struct Sample {
  bool valid;
  int lo, hi;
  float lerp;
};

// coordinate of output index i (of `crop`) along an axis of `size` pixels for the box edge pair (c1,c2)
__device__ __forceinline__ Sample crop_coord(float c1, float c2, int i, int crop, int size) {
  const float sm1 = (float)(size - 1);
  float in;
  if (crop > 1) {
    const float scale = __fdiv_rn(__fmul_rn(__fsub_rn(c2, c1), sm1), (float)(crop - 1));
    in = __fadd_rn(__fmul_rn(c1, sm1), __fmul_rn((float)i, scale));
  } else {
    in = __fmul_rn(__fmul_rn(0.5f, __fadd_rn(c1, c2)), sm1);
  }
  Sample s;
  s.valid = (in >= 0.f) && (in <= sm1);  // NaN -> invalid -> extrapolation value 0
  const float f = floorf(in);
  s.lo = s.valid ? (int)f : 0;
  s.hi = s.valid ? (int)ceilf(in) : 0;
  s.lerp = s.valid ? __fsub_rn(in, f) : 0.f;
  return s;
}

__device__ __forceinline__ float lerp_rn(float a, float b, float t) {
  return __fadd_rn(a, __fmul_rn(__fsub_rn(b, a), t));
}

}  // namespace myolo
