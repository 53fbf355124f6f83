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
// Two launches: the candidate boxes of all instances once (a warp each), then the pixels; a CTA first lists the instances
// whose box meets its pixels, so that most pixels test nothing.
#include "common.cuh"
#include "polygon_pip.h"

namespace myolo {

constexpr int kPolyMaxInst = 128;

__device__ __forceinline__ int poly_clamp_int(double v) {      // int() of the reference on a value made safe to convert
  return (int)fmin(fmax(v, -1.0e9), 1.0e9);
}

// kernel 1 (one warp per instance): the candidate box skimage's _polygon scans, minr = int(max(0, r.min())),
// maxr = int(ceil(r.max())), the same for the columns, as (minr, maxr, minc, maxc); an instance without vertices gets an
// empty box.
__global__ void __launch_bounds__(32)
polygon_boxes_kernel(const double* __restrict__ vy, const double* __restrict__ vx, const int* __restrict__ off,
                     int4* __restrict__ boxes) {
  const int i = blockIdx.x, lane = threadIdx.x;
  const int b = off[i], e = off[i + 1];
  double rmin = 1.0e300, rmax = -1.0e300, cmin = 1.0e300, cmax = -1.0e300;
  for (int v = b + lane; v < e; v += 32) {
    const double r = vy[v], c = vx[v];
    rmin = fmin(rmin, r); rmax = fmax(rmax, r);
    cmin = fmin(cmin, c); cmax = fmax(cmax, c);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    rmin = fmin(rmin, __shfl_xor_sync(0xffffffffu, rmin, o)); rmax = fmax(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
    cmin = fmin(cmin, __shfl_xor_sync(0xffffffffu, cmin, o)); cmax = fmax(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
  }
  if (lane == 0)
    boxes[i] = e > b ? make_int4(poly_clamp_int(fmax(0.0, rmin)), poly_clamp_int(ceil(rmax)), poly_clamp_int(fmax(0.0, cmin)),
                                 poly_clamp_int(ceil(cmax)))
                     : make_int4(1, 0, 1, 0);
}

// kernel 2 (256 consecutive pixels per CTA): the instances whose box meets the CTA's pixels are listed once per CTA, every
// pixel then runs the crossing test against those only.
__global__ void __launch_bounds__(256)
polygon_masks_kernel(const double* __restrict__ vy, const double* __restrict__ vx, const int* __restrict__ off,
                     const int4* __restrict__ boxes, int n_inst, int H, int W, int M, unsigned char* __restrict__ masks) {
  extern __shared__ __align__(16) unsigned char s_mask[];      // [256][M]
  __shared__ int4 s_box[kPolyMaxInst];
  __shared__ int s_inst[kPolyMaxInst], s_beg[kPolyMaxInst], s_cnt[kPolyMaxInst];
  __shared__ int s_n;
  const int tid = threadIdx.x;
  const long long npix = (long long)H * W;
  const long long pix0 = (long long)blockIdx.x * 256;
  const long long pix1 = min(pix0 + 255, npix - 1);            // last pixel of this CTA
  const int row0 = (int)(pix0 / W), row1 = (int)(pix1 / W);
  const int col0 = (int)(pix0 - (long long)row0 * W), col1 = (int)(pix1 - (long long)row1 * W);
  if (tid == 0) s_n = 0;
  const int words = 256 * M / 4;                               // 256*M is a multiple of 4
  for (int w = tid; w < words; w += 256) reinterpret_cast<uint32_t*>(s_mask)[w] = 0u;
  __syncthreads();
  if (tid < n_inst) {
    const int4 bx = boxes[tid];
    bool hit = bx.x <= row1 && bx.y >= row0 && bx.z <= bx.w;
    if (hit && row0 == row1) hit = bx.z <= col1 && bx.w >= col0;     // a CTA inside one image row: its column span counts too
    if (hit) {
      const int k = atomicAdd(&s_n, 1);
      const int b = off[tid];
      s_inst[k] = tid; s_box[k] = bx; s_beg[k] = b; s_cnt[k] = off[tid + 1] - b;
    }
  }
  __syncthreads();
  const int n = s_n;
  const long long pix = pix0 + tid;
  if (pix < npix && n > 0) {
    const int r = (int)(pix / W), c = (int)(pix - (long long)r * W);
    const double y = (double)r, x = (double)c;
    for (int k = 0; k < n; ++k) {
      const int4 bx = s_box[k];
      if (r < bx.x || r > bx.y || c < bx.z || c > bx.w) continue;
      const int b = s_beg[k];
      if (myolo_polygon::point_in_polygon(s_cnt[k], vx + b, vy + b, x, y)) s_mask[tid * M + s_inst[k]] = 1;
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
                                   int M, int* ws, unsigned char* masks, myolo_stream stream) {
  MYOLO_CHECK_ARG(verts_y && verts_x && offsets && masks && ws);
  MYOLO_CHECK_ARG(H > 0 && W > 0 && H <= 65536 && W <= 65536);
  MYOLO_CHECK_ARG(n_inst >= 0 && n_inst <= M && M >= 1 && M <= myolo::kPolyMaxInst);
  MYOLO_CHECK_ARG((reinterpret_cast<uintptr_t>(masks) & 3) == 0 && (reinterpret_cast<uintptr_t>(ws) & 15) == 0);
  MYOLO_CHECK_ARG(((reinterpret_cast<uintptr_t>(verts_y) | reinterpret_cast<uintptr_t>(verts_x)) & 7) == 0);
  cudaStream_t st = myolo::as_stream(stream);
  int4* boxes = reinterpret_cast<int4*>(ws);
  if (n_inst > 0) {
    myolo::polygon_boxes_kernel<<<n_inst, 32, 0, st>>>(verts_y, verts_x, offsets, boxes);
    MYOLO_CHECK_LAUNCH();
  }
  const long long blocks = ((long long)H * W + 255) / 256;
  myolo::polygon_masks_kernel<<<(unsigned)blocks, 256, (size_t)256 * M, st>>>(verts_y, verts_x, offsets, boxes, n_inst, H, W, M,
                                                                            masks);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}
