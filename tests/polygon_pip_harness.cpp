// Host build of mask-yolo_b200/csrc/polygon_pip.h for tests/test_via_polygons.py (g++ -ffp-contract=off, no CUDA): the same
// inclusion test the device kernel of polygon.cu calls, over a C ABI, so that it can be compared with the oracle on the CPU.
#include "polygon_pip.h"

extern "C" void polygon_mask_host(int n, const double* ys, const double* xs, int H, int W, unsigned char* mask) {
  for (int r = 0; r < H; ++r)
    for (int c = 0; c < W; ++c)
      mask[(long long)r * W + c] = myolo_polygon::point_in_polygon(n, xs, ys, (double)c, (double)r) ? 1 : 0;
}
