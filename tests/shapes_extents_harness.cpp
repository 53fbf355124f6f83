// Host build of mask-yolo_b200/csrc/shapes_extents.h for tests/test_shapes_raster.py (g++, no CUDA): the same functions
// the device kernels of shapes.cu call, exposed over a C ABI so that they can be compared with cv2 on the CPU.
#include "shapes_extents.h"

extern "C" void shape_rows_host(int type, int x, int y, int s, int W, int H, int* lo, int* hi) {
  for (int r = 0; r < H; ++r) { lo[r] = W; hi[r] = -1; }
  myolo_shapes::shape_rows(lo, hi, 1, W, H, type, x, y, s);
}
