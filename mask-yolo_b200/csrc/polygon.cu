// polygon.cu -- VIA polygon annotations rasterised on the device (second half).
//
// Replaces the host loop of RiceDataset.load_mask, example/rice/rice_dataset.py:135-159 (and its twin under example/food):
//     mask = zeros([height, width, len(polygons)], uint8)
//     for i, p in enumerate(polygons):
//         rr, cc = skimage.draw.polygon(p['all_points_y'], p['all_points_x']);  mask[rr, cc, i] = 1
// for all instances of one image in one launch.  The inclusion rule is skimage's crossing-number test in float64
// (polygon_pip.h restates it; the same header is compiled for the host and compared with the oracle on the CPU).
// Layout and roof: the output is the reference's [height, width, instances] byte tensor, H*W*M bytes written once --
// an HBM-write-bound kernel.  A CTA owns 256 consecutive pixels = 256*M consecutive bytes, composes them in shared memory
// and stores them as 32-bit words; vertex reads are warp-uniform (every lane tests the same edge) and stay in L1.
#include "common.cuh"
#include "polygon_pip.h"

namespace myolo {

constexpr int kPolyMaxInst = 128;

__device__ __forceinline__ int poly_clamp_int(double v) {      // int() of the reference on a value made safe to convert
  return (int)fmin(fmax(v, -1.0e9), 1.0e9);
}

__global__ void __launch_bounds__(256)
polygon_masks_kernel(const double* __restrict__ vy, const double* __restrict__ vx, const int* __restrict__ off, int n_inst,
                     int H, int W, int M, unsigned char* __restrict__ masks) {
  extern __shared__ __align__(16) unsigned char s_mask[];      // [256][M]
  __shared__ int s_box[kPolyMaxInst][4];                       // minr, maxr, minc, maxc as skimage's _polygon computes them
  const int tid = threadIdx.x;
  for (int i = tid; i < n_inst; i += 256) {
    const int b = off[i], e = off[i + 1];
    double rmin = 1.0e300, rmax = -1.0e300, cmin = 1.0e300, cmax = -1.0e300;
    for (int v = b; v < e; ++v) {
      const double r = vy[v], c = vx[v];
      rmin = fmin(rmin, r); rmax = fmax(rmax, r);
      cmin = fmin(cmin, c); cmax = fmax(cmax, c);
    }
    const bool any = e > b;
    s_box[i][0] = any ? poly_clamp_int(fmax(0.0, rmin)) : 1;
    s_box[i][1] = any ? poly_clamp_int(ceil(rmax)) : 0;
    s_box[i][2] = any ? poly_clamp_int(fmax(0.0, cmin)) : 1;
    s_box[i][3] = any ? poly_clamp_int(ceil(cmax)) : 0;
  }
  const int words = 256 * M / 4;                               // 256*M is a multiple of 4
  for (int w = tid; w < words; w += 256) reinterpret_cast<uint32_t*>(s_mask)[w] = 0u;
  __syncthreads();
  const long long npix = (long long)H * W;
  const long long pix0 = (long long)blockIdx.x * 256;
  const long long pix = pix0 + tid;
  if (pix < npix) {
    const int r = (int)(pix / W), c = (int)(pix - (long long)r * W);
    const double y = (double)r, x = (double)c;
    for (int i = 0; i < n_inst; ++i) {
      if (r < s_box[i][0] || r > s_box[i][1] || c < s_box[i][2] || c > s_box[i][3]) continue;
      const int b = off[i];
      if (myolo_polygon::point_in_polygon(off[i + 1] - b, vx + b, vy + b, x, y)) s_mask[tid * M + i] = 1;
    }
  }
  __syncthreads();
  unsigned char* dst = masks + pix0 * M;
  if (pix0 + 256 <= npix) {
    for (int w = tid; w < words; w += 256) reinterpret_cast<uint32_t*>(dst)[w] = reinterpret_cast<const uint32_t*>(s_mask)[w];
  } else {                                                     // the image's last, partial group of pixels
    const long long bytes = (npix - pix0) * M;
    for (long long k = tid; k < bytes; k += 256) dst[k] = s_mask[k];
  }
}

}  // namespace myolo

extern "C" int myolo_polygon_masks(const double* verts_y, const double* verts_x, const int* offsets, int n_inst, int H, int W,
                                   int M, unsigned char* masks, myolo_stream stream) {
  MYOLO_CHECK_ARG(verts_y && verts_x && offsets && masks);
  MYOLO_CHECK_ARG(H > 0 && W > 0 && H <= 65536 && W <= 65536);
  MYOLO_CHECK_ARG(n_inst >= 0 && n_inst <= M && M >= 1 && M <= myolo::kPolyMaxInst);
  MYOLO_CHECK_ARG((reinterpret_cast<uintptr_t>(masks) & 3) == 0);
  MYOLO_CHECK_ARG(((reinterpret_cast<uintptr_t>(verts_y) | reinterpret_cast<uintptr_t>(verts_x)) & 7) == 0);
  const long long blocks = ((long long)H * W + 255) / 256;
  myolo::polygon_masks_kernel<<<(unsigned)blocks, 256, (size_t)256 * M, myolo::as_stream(stream)>>>(
      verts_y, verts_x, offsets, n_inst, H, W, M, masks);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}
