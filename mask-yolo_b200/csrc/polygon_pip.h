// polygon_pip.h -- the point-in-polygon rule of skimage.draw.polygon, shared by the device kernel (polygon.cu) and a host
// build for the CPU tests (tests/polygon_pip_harness.cpp).
//
// The reference rasterises VIA polygon annotations with `skimage.draw.polygon(all_points_y, all_points_x)`
// (example/rice/rice_dataset.py:151-153, example/food/rice_dataset.py likewise).  scikit-image is a third-party,
// un-pinned dependency that is absent from this image, so its published algorithm is restated here
// (scikit-image 0.13/0.14, the releases contemporary with the reference: skimage/draw/_draw.pyx `_polygon` and
// skimage/_shared/geometry.pxd `point_in_polygon`):
//
//   minr = int(max(0, r.min()));  maxr = int(ceil(r.max()));  the same for c      (no upper clamp without `shape`)
//   for r_i in minr..maxr, c_i in minc..maxc:  keep (r_i, c_i) iff point_in_polygon(c, r, c_i, r_i)
//   point_in_polygon: crossing number over the edges (j -> i), j = i - 1 cyclic, in float64:
//       if ((yp[i] <= y < yp[j]) or (yp[j] <= y < yp[i])) and x < (xp[j]-xp[i]) * (y-yp[i]) / (yp[j]-yp[i]) + xp[i]:  c = !c
//
// Every operation is an individually rounded IEEE double operation in that order (no FMA contraction): the device build
// uses the explicit round-to-nearest intrinsics, the host build is compiled with -ffp-contract=off.
// PARITY UNPINNED against scikit-image itself (not installable here); pinned against the oracle's plain-Python
// restatement and hand-computed cases (tests/test_via_polygons.py).
#pragma once

#if defined(__CUDACC__)
#define MYOLO_PIP_HD __host__ __device__ __forceinline__
#else
#define MYOLO_PIP_HD inline
#endif

namespace myolo_polygon {

MYOLO_PIP_HD double pip_mul(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
MYOLO_PIP_HD double pip_add(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}
MYOLO_PIP_HD double pip_div(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __ddiv_rn(a, b);
#else
  return a / b;
#endif
}

// xp / yp: the polygon's n vertices (x = column, y = row); (x, y): the pixel.  Returns the crossing parity.
MYOLO_PIP_HD bool point_in_polygon(int n, const double* xp, const double* yp, double x, double y) {
  bool c = false;
  int j = n - 1;
  for (int i = 0; i < n; ++i) {
    const double yi = yp[i], yj = yp[j];
    if (((yi <= y) && (y < yj)) || ((yj <= y) && (y < yi))) {
      const double xi = xp[i];
      const double t = pip_add(pip_div(pip_mul(pip_add(xp[j], -xi), pip_add(y, -yi)), pip_add(yj, -yi)), xi);
      if (x < t) c = !c;
    }
    j = i;
  }
  return c;
}

}  // namespace myolo_polygon
