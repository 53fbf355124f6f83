// shapes_extents.h -- row extents of the three Shapes primitives exactly as OpenCV fills them.
//
// The Shapes workload (example/shapes/dataset_shapes.py:120-135, draw_shape) rasterises with
//   square   cv2.rectangle(img, (x-s, y-s), (x+s, y+s), color, -1)
//   circle   cv2.circle(img, (x, y), s, color, -1)
//   triangle cv2.fillPoly(img, int32[(x, y-s), (x - s/sin60, y+s), (x + s/sin60, y+s)], color)
// OpenCV is a third-party dependency of the reference (unpinned; 4.13.0 in this image).  Its algorithms for
// these three calls (imgproc/drawing.cpp: Circle() midpoint spans; fillPoly = Bresenham boundary lines through
// LineIterator/clipLine + the 16.16 fixed-point scanline of FillEdgeCollection) are restated here as per-row
// column intervals [lo[r], hi[r]] (every row of every such shape is one interval), clipped to a W x H image.
// The file is plain C++ without CUDA types so that tests/test_shapes_raster.py can compile it with g++ and
// compare it with cv2 pixel for pixel on the CPU; the device kernels (shapes.cu) include the same functions.
#pragma once

#if defined(__CUDACC__)
#define MYOLO_HD __host__ __device__
#else
#define MYOLO_HD
#endif

namespace myolo_shapes {

enum { SHAPE_SQUARE = 1, SHAPE_CIRCLE = 2, SHAPE_TRIANGLE = 3 };   // = class ids of dataset_shapes.py:61-63

// rows are addressed as lo[r * stride], hi[r * stride]; an empty row has lo > hi
MYOLO_HD static inline void ext_put(int* lo, int* hi, int stride, int W, int H, int r, int a, int b) {
  if ((unsigned)r >= (unsigned)H) return;
  if (a < 0) a = 0;
  if (b > W - 1) b = W - 1;
  if (a > b) return;
  if (a < lo[r * stride]) lo[r * stride] = a;
  if (b > hi[r * stride]) hi[r * stride] = b;
}

// cv::clipLine(Size, Point&, Point&): Cohen-Sutherland with OpenCV's double rounding
MYOLO_HD static inline bool clip_line(int W, int H, long long& x1, long long& y1, long long& x2, long long& y2) {
  const long long right = W - 1, bottom = H - 1;
  int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
  int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
  if ((c1 & c2) == 0 && (c1 | c2) != 0) {
    long long a;
    if (c1 & 12) {
      a = c1 < 8 ? 0 : bottom;
      x1 += (long long)((double)(a - y1) * (double)(x2 - x1) / (double)(y2 - y1));
      y1 = a;
      c1 = (x1 < 0) + (x1 > right) * 2;
    }
    if (c2 & 12) {
      a = c2 < 8 ? 0 : bottom;
      x2 += (long long)((double)(a - y2) * (double)(x2 - x1) / (double)(y2 - y1));
      y2 = a;
      c2 = (x2 < 0) + (x2 > right) * 2;
    }
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
      if (c1) {
        a = c1 == 1 ? 0 : right;
        y1 += (long long)((double)(a - x1) * (double)(y2 - y1) / (double)(x2 - x1));
        x1 = a;
        c1 = 0;
      }
      if (c2) {
        a = c2 == 1 ? 0 : right;
        y2 += (long long)((double)(a - x2) * (double)(y2 - y1) / (double)(x2 - x1));
        x2 = a;
        c2 = 0;
      }
    }
  }
  return (c1 | c2) == 0;
}

// cv::LineIterator(img, p1, p2, connectivity 8, left_to_right) as used by Line(): every visited pixel widens its row
MYOLO_HD static inline void line_rows(int* lo, int* hi, int stride, int W, int H, long long x1, long long y1, long long x2,
                                      long long y2) {
  if ((unsigned long long)x1 >= (unsigned long long)W || (unsigned long long)x2 >= (unsigned long long)W ||
      (unsigned long long)y1 >= (unsigned long long)H || (unsigned long long)y2 >= (unsigned long long)H) {
    if (!clip_line(W, H, x1, y1, x2, y2)) return;
  }
  long long dx = x2 - x1, dy = y2 - y1;
  int sx = 1, sy = 1;
  if (dx < 0) { dx = -dx; dy = -dy; x1 = x2; y1 = y2; }
  if (dy < 0) { dy = -dy; sy = -1; }
  const bool vert = dy > dx;
  if (vert) { long long t = dx; dx = dy; dy = t; }
  long long err = dx - (dy + dy);
  const long long plus = dx + dx, minus = -(dy + dy);
  long long x = x1, y = y1;
  for (long long k = 0; k <= dx; ++k) {
    ext_put(lo, hi, stride, W, H, (int)y, (int)x, (int)x);
    const bool m = err < 0;
    err += minus + (m ? plus : 0);
    if (vert) { y += sy; if (m) x += sx; }
    else      { x += sx; if (m) y += sy; }
  }
}

// cv2.circle(..., thickness -1): Circle(img, center, radius, color, fill = 1); all spans are centred on cx
MYOLO_HD static inline void circle_rows(int* lo, int* hi, int stride, int W, int H, int cx, int cy, int radius) {
  int err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1;
  while (dx >= dy) {
    ext_put(lo, hi, stride, W, H, cy - dy, cx - dx, cx + dx);
    ext_put(lo, hi, stride, W, H, cy + dy, cx - dx, cx + dx);
    ext_put(lo, hi, stride, W, H, cy - dx, cx - dy, cx + dy);
    ext_put(lo, hi, stride, W, H, cy + dx, cx - dy, cx + dy);
    dy++;
    err += plus;
    plus += 2;
    const int mask = (err <= 0) - 1;
    err -= minus & mask;
    dx += mask;
    minus -= mask & 2;
  }
}

// cv2.rectangle(..., thickness -1): the closed rectangle [x1, x2] x [y1, y2]
MYOLO_HD static inline void rect_rows(int* lo, int* hi, int stride, int W, int H, int x1, int y1, int x2, int y2) {
  if (x1 > x2) { int t = x1; x1 = x2; x2 = t; }
  if (y1 > y2) { int t = y1; y1 = y2; y2 = t; }
  for (int r = (y1 < 0 ? 0 : y1); r <= y2 && r < H; ++r) ext_put(lo, hi, stride, W, H, r, x1, x2);
}

// cv2.fillPoly of the triangle apex (ax, ay), base corners (bl, by) and (br, by), by > ay: CollectPolyEdges draws the three
// boundary lines and builds one PolyEdge per slanted side (16.16 fixed point, start at +1/2 pixel; when a side leaves the
// image its x comes from the CLIPPED end points -- even when clipping leaves a single pixel or rejects the side -- and its
// y from them only if they still span rows); FillEdgeCollection then fills rows ay .. by-1 from x_left >> 16 to
// (x_right - 1) >> 16.
MYOLO_HD static inline void triangle_rows(int* lo, int* hi, int stride, int W, int H, int ax, int ay, int bl, int br, int by) {
  const int XY_SHIFT = 16;
  const long long HALF = 1LL << (XY_SHIFT - 1);
  const long long px[3] = {ax, bl, br}, py[3] = {ay, by, by};
  long long ex[2] = {0, 0}, edx[2] = {0, 0};
  int ne = 0;
  long long p0x = px[2], p0y = py[2];
  for (int i = 0; i < 3; ++i) {
    const long long p1x = px[i], p1y = py[i];
    line_rows(lo, hi, stride, W, H, p0x, p0y, p1x, p1y);
    if (p0y != p1y && ne < 2) {
      long long c0x = (p0x << XY_SHIFT) + HALF, c0y = p0y, c1x = (p1x << XY_SHIFT) + HALF, c1y = p1y;
      if ((unsigned long long)p0x >= (unsigned long long)W || (unsigned long long)p1x >= (unsigned long long)W ||
          (unsigned long long)p0y >= (unsigned long long)H || (unsigned long long)p1y >= (unsigned long long)H) {
        long long t0x = p0x, t0y = p0y, t1x = p1x, t1y = p1y;
        clip_line(W, H, t0x, t0y, t1x, t1y);
        c0x = (t0x << XY_SHIFT) + HALF;                           // x always from the clipped points ...
        c1x = (t1x << XY_SHIFT) + HALF;
        if (t0y != t1y) { c0y = t0y; c1y = t1y; }                 // ... y only when the clipped side still spans rows
      }
      const long long d = (c1x - c0x) / (c1y - c0y);            // truncating division, as in C++
      const long long y0 = p0y < p1y ? p0y : p1y;
      ex[ne] = p0y < p1y ? c0x + (y0 - c0y) * d : c1x + (y0 - c1y) * d;
      edx[ne] = d;
      ++ne;
    }
    p0x = p1x; p0y = p1y;
  }
  if (ne != 2) return;
  for (long long y = ay; y < by && y < H; ++y) {
    if (y >= 0) {
      const long long xa = ex[0] < ex[1] ? ex[0] : ex[1], xb = ex[0] < ex[1] ? ex[1] : ex[0];
      const long long x1 = xa >> XY_SHIFT, x2 = (xb - 1) >> XY_SHIFT;
      if (x1 < W && x2 >= 0) ext_put(lo, hi, stride, W, H, (int)y, (int)x1, (int)x2);
    }
    ex[0] += edx[0];
    ex[1] += edx[1];
  }
}

// draw_shape (dataset_shapes.py:120-135): rows lo/hi must be initialised empty (lo = W, hi = -1) by the caller
MYOLO_HD static inline void shape_rows(int* lo, int* hi, int stride, int W, int H, int type, int x, int y, int s) {
  if (type == SHAPE_SQUARE) {
    rect_rows(lo, hi, stride, W, H, x - s, y - s, x + s, y + s);
  } else if (type == SHAPE_CIRCLE) {
    circle_rows(lo, hi, stride, W, H, x, y, s);
  } else if (type == SHAPE_TRIANGLE) {
    const double sin60 = 0x1.bb67ae8584caap-1;                   // math.sin(math.radians(60))
    const double half = (double)s / sin60;
    // np.array([...], dtype=np.int32) truncates the float64 corners toward zero
    triangle_rows(lo, hi, stride, W, H, x, y - s, (int)((double)x - half), (int)((double)x + half), y + s);
  }
}

}  // namespace myolo_shapes
