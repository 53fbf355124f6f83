// postprocess.cu -- inference post-processing on the device: what MaskYOLO.detect does
// in numpy after keras_model.predict (myolo/model.py:1290-1304, 1330-1391; myolo_utils.py:88-113 NMB,
// 883-912 unmold_mask): top-k detections by confidence, confidence threshold, NMB box suppression (the reference's rule),
// and pasting each survivor's class-specific 28x28 soft mask into its pixel box (bilinear resize,
// threshold 0.5).  One block per image for the selection, one block per (detection, image) for the paste.
#include <math_constants.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

namespace myolo {

constexpr int kMaxTopK = 32;

// myolo_utils._interval_overlap (231-244), kept branch for branch
__device__ __forceinline__ float interval_overlap(float a1, float a2, float b1, float b2) {
  if (b1 < a1) return b2 < a1 ? 0.f : fminf(a2, b2) - a1;
  return a2 < b1 ? 0.f : fminf(a2, b2) - b1;
}
__device__ __forceinline__ float box_iou(const float* p, const float* q) {   // bbox_iou_2 (201-228)
  const float iw = interval_overlap(p[0], p[2], q[0], q[2]);
  const float ih = interval_overlap(p[1], p[3], q[1], q[3]);
  const float inter = iw * ih;
  const float uni = (p[2] - p[0]) * (p[3] - p[1]) + (q[2] - q[0]) * (q[3] - q[1]) - inter;
  return inter / uni;
}

__global__ void __launch_bounds__(256)
detect_select_kernel(const float* __restrict__ det, int R, int S, int top_k, float cs_thr, float nms_thr,
                     int* __restrict__ out_idx, int* __restrict__ out_boxes, int* __restrict__ out_class,
                     float* __restrict__ out_score, int* __restrict__ out_count) {
  extern __shared__ float sc[];                     // [R] scores, -inf once taken
  __shared__ float rv[8];
  __shared__ int ri[8];
  __shared__ int cand[kMaxTopK];
  __shared__ int ncand;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
  const float* D = det + (size_t)b * R * 6;
  // decode_masks drops boxes without area before anything is ranked (model.py:1367-1375)
  for (int r = tid; r < R; r += blockDim.x) {
    const float* p = D + (size_t)r * 6;
    sc[r] = ((p[2] - p[0]) * (p[3] - p[1]) <= 0.f) ? -CUDART_INF_F : p[4];
  }
  if (tid == 0) ncand = 0;
  __syncthreads();
  for (int k = 0; k < top_k; ++k) {                 // k-th largest score; HIGHEST index on ties (np.argsort(scores)[::-1])
    float bv = -CUDART_INF_F;
    int bi = -1;
    for (int r = tid; r < R; r += blockDim.x) {
      const float v = sc[r];
      if (v >= bv) { bv = v; bi = r; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi > bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { rv[wp] = bv; ri[wp] = bi; }
    __syncthreads();
    if (tid == 0) {
      float v = rv[0];
      int i = ri[0];
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
        if (rv[w] > v || (rv[w] == v && ri[w] > i)) { v = rv[w]; i = ri[w]; }
      if (i >= 0 && v == -CUDART_INF_F) i = -1;      // nothing left to rank (taken already, or no area)
      if (i >= 0 && v >= cs_thr) cand[ncand++] = i;  // detections below the confidence threshold are dropped
      if (i >= 0) sc[i] = -CUDART_INF_F;
    }
    __syncthreads();
  }
  if (tid == 0) {
    // NMB exactly as myolo_utils.py:88-113 (pinned by the reference's own outputs, tests/golden): candidate c is
    // dropped when ANY earlier candidate j of the same class overlaps it by >= thr -- also a j that was dropped
    // itself (this is not greedy NMS).
    int kept = 0;
    int keep[kMaxTopK];
    for (int c = 0; c < ncand; ++c) {
      const float* p = D + (size_t)cand[c] * 6;
      bool ok = true;
      for (int j = 0; j < c && ok; ++j) {
        const float* q = D + (size_t)cand[j] * 6;
        ok = !(q[5] == p[5] && box_iou(q, p) >= nms_thr);
      }
      if (ok) keep[kept++] = cand[c];
    }
    for (int j = 0; j < top_k; ++j) {
      const int o = b * top_k + j;
      if (j < kept) {
        const float* p = D + (size_t)keep[j] * 6;
        out_idx[o] = keep[j];
        // the pixel box unmold_mask pastes into (myolo_utils.py:895-901): int() truncation, x1/y1 in [0,S], x2/y2 in [1,S]
        for (int e = 0; e < 4; ++e) out_boxes[o * 4 + e] = min(max(e < 2 ? 0 : 1, (int)(p[e] * (float)S)), S);
        out_class[o] = (int)p[5];
        out_score[o] = p[4];
      } else {
        out_idx[o] = -1;
        for (int e = 0; e < 4; ++e) out_boxes[o * 4 + e] = 0;
        out_class[o] = 0;
        out_score[o] = 0.f;
      }
    }
    out_count[b] = kept;
  }
}

// grid (top_k, B).  Full-size boolean mask of detection j of image b (all zero when j >= count).
__global__ void __launch_bounds__(256)
mask_paste_kernel(const float* __restrict__ det, const float* __restrict__ masks, const int* __restrict__ idx,
                  int R, int NC, int S, int MH, int MW, int top_k, unsigned char* __restrict__ out) {
  const int j = blockIdx.x, b = blockIdx.y;
  unsigned char* O = out + ((size_t)b * top_k + j) * S * S;
  const int r = idx[b * top_k + j];
  int x1 = 0, y1 = 0, x2 = 0, y2 = 0, cls = 0;
  if (r >= 0) {
    const float* p = det + ((size_t)b * R + r) * 6;
    // the reference's unmold_mask (myolo_utils.py:895-901): int() truncation, x1/y1 clamped to [0,S], x2/y2 to [1,S];
    // the mask is then resized into this CLIPPED box
    x1 = min(max(0, (int)(p[0] * (float)S)), S); y1 = min(max(0, (int)(p[1] * (float)S)), S);
    x2 = min(max(1, (int)(p[2] * (float)S)), S); y2 = min(max(1, (int)(p[3] * (float)S)), S);
    cls = (int)p[5];
  }
  const int bw = x2 - x1, bh = y2 - y1;
  const bool live = r >= 0 && bw > 0 && bh > 0 && cls >= 0 && cls < NC;
  const float* Mk = masks + (((size_t)b * R + (r >= 0 ? r : 0)) * MH * MW) * NC + cls;
  const float fx = (float)MW / (float)max(bw, 1), fy = (float)MH / (float)max(bh, 1);
  for (int i = threadIdx.x; i < S * S; i += blockDim.x) {
    const int y = i / S, x = i - y * S;
    unsigned char v = 0;
    if (live && x >= x1 && x < x2 && y >= y1 && y < y2) {
      // bilinear resize of the MHxMW mask to (bh, bw), pixel centres aligned, edges replicated
      float sx = ((float)(x - x1) + 0.5f) * fx - 0.5f, sy = ((float)(y - y1) + 0.5f) * fy - 0.5f;
      int ix = (int)floorf(sx), iy = (int)floorf(sy);
      float ax = sx - (float)ix, ay = sy - (float)iy;
      if (ix < 0) { ix = 0; ax = 0.f; }
      if (iy < 0) { iy = 0; ay = 0.f; }
      if (ix >= MW - 1) { ix = MW - 2 < 0 ? 0 : MW - 2; ax = MW > 1 ? 1.f : 0.f; }
      if (iy >= MH - 1) { iy = MH - 2 < 0 ? 0 : MH - 2; ay = MH > 1 ? 1.f : 0.f; }
      const int ix1 = min(ix + 1, MW - 1), iy1 = min(iy + 1, MH - 1);
      const float m00 = Mk[((size_t)iy * MW + ix) * NC], m01 = Mk[((size_t)iy * MW + ix1) * NC];
      const float m10 = Mk[((size_t)iy1 * MW + ix) * NC], m11 = Mk[((size_t)iy1 * MW + ix1) * NC];
      const float top = m00 + (m01 - m00) * ax, bot = m10 + (m11 - m10) * ax;
      v = (top + (bot - top) * ay) >= 0.5f ? 1 : 0;
    }
    O[i] = v;
  }
}

}  // namespace myolo

using namespace myolo;

extern "C" int myolo_detect_postprocess(const float* detections, const float* masks, int B, int R, int NC, int S, int MH,
                                        int MW, int top_k, float cs_threshold, float nms_threshold, int* out_index,
                                        int* out_boxes, int* out_class, float* out_score, int* out_count,
                                        unsigned char* out_masks, myolo_stream stream) {
  MYOLO_CHECK_ARG(detections && out_index && out_boxes && out_class && out_score && out_count);
  MYOLO_CHECK_ARG(B > 0 && R > 0 && NC > 0 && S > 0 && MH > 0 && MW > 0 && top_k > 0 && top_k <= kMaxTopK);
  MYOLO_CHECK_ARG((size_t)R * sizeof(float) <= 160 * 1024);
  MYOLO_CHECK_ARG((masks == nullptr) == (out_masks == nullptr));
  cudaStream_t st = as_stream(stream);
  const size_t smem = (size_t)R * sizeof(float);
  if (smem > 48 * 1024) MYOLO_CUDA(cudaFuncSetAttribute(detect_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  detect_select_kernel<<<B, 256, smem, st>>>(detections, R, S, top_k, cs_threshold, nms_threshold, out_index, out_boxes,
                                              out_class, out_score, out_count);
  if (masks) {
    dim3 grid(top_k, B);
    mask_paste_kernel<<<grid, 256, 0, st>>>(detections, masks, out_index, R, NC, S, MH, MW, top_k, out_masks);
  }
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

// ------------------------------------------------------------------------------------------------
// the host loops of BatchGenerator.__getitem__ (myolo_utils.py:769-820) and
// extract_bboxes (247-271) on the device.
// ------------------------------------------------------------------------------------------------
namespace myolo {

// one block per (instance, image): bounding box of a [S,S,M] byte mask column, x2/y2 exclusive
__global__ void __launch_bounds__(256)
extract_bboxes_kernel(const unsigned char* __restrict__ masks, int S, int M, int* __restrict__ boxes) {
  const int m = blockIdx.x, b = blockIdx.y;
  const unsigned char* P = masks + (size_t)b * S * S * M + m;
  int x1 = S, y1 = S, x2 = -1, y2 = -1;
  for (int i = threadIdx.x; i < S * S; i += blockDim.x) {
    if (P[(size_t)i * M]) {
      const int y = i / S, x = i - y * S;
      x1 = min(x1, x); x2 = max(x2, x);
      y1 = min(y1, y); y2 = max(y2, y);
    }
  }
  __shared__ int r[4][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    x1 = min(x1, __shfl_xor_sync(0xffffffffu, x1, o));
    y1 = min(y1, __shfl_xor_sync(0xffffffffu, y1, o));
    x2 = max(x2, __shfl_xor_sync(0xffffffffu, x2, o));
    y2 = max(y2, __shfl_xor_sync(0xffffffffu, y2, o));
  }
  const int wp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { r[0][wp] = x1; r[1][wp] = y1; r[2][wp] = x2; r[3][wp] = y2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) {
      x1 = min(x1, r[0][w]); y1 = min(y1, r[1][w]);
      x2 = max(x2, r[2][w]); y2 = max(y2, r[3][w]);
    }
    int* o = boxes + ((size_t)b * M + m) * 4;
    if (x2 < 0) { o[0] = o[1] = o[2] = o[3] = 0; }         // empty mask -> zero box
    else { o[0] = x1; o[1] = y1; o[2] = x2 + 1; o[3] = y2 + 1; }
  }
}

// one thread per image, instances in order (a later instance overwrites an earlier one in the same
// cell/anchor; the true-box buffer is a ring), exactly like the host loop.
__global__ void encode_yolo_targets_kernel(const int* __restrict__ ids, const int* __restrict__ boxes, int B, int M, int S,
                                           int G, int NB, int NC, int TB, const float* __restrict__ anchors,
                                           float* __restrict__ yolo_target, float* __restrict__ true_boxes) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int D = 5 + NC;
  float* Y = yolo_target + (size_t)b * G * G * NB * D;
  float* T = true_boxes + (size_t)b * TB * 4;
  const double cell = (double)S / (double)G;
  int slot = 0;
  for (int i = 0; i < M; ++i) {
    const int cls = ids[(size_t)b * M + i];
    if (cls <= 0) continue;                                  // zero padding of the [B, TRUE_BOX_BUFFER] arrays
    const int* q = boxes + ((size_t)b * M + i) * 4;
    const double cx = 0.5 * (q[0] + q[2]) / cell, cy = 0.5 * (q[1] + q[3]) / cell;
    const int gx = (int)floor(cx), gy = (int)floor(cy);
    if (gx < G && gy < G) {
      const double w = (q[2] - q[0]) / cell, h = (q[3] - q[1]) / cell;
      int best = -1;
      double best_iou = -1.0;
      for (int a = 0; a < NB; ++a) {                         // bbox_iou((0,0,w,h), (0,0,aw,ah)); first maximum wins
        const double aw = (double)anchors[2 * a], ah = (double)anchors[2 * a + 1];
        const double iw = (aw < 0.0) ? 0.0 : fmin(w, aw), ih = (ah < 0.0) ? 0.0 : fmin(h, ah);
        const double inter = iw * ih;
        const double iou = inter / (w * h + aw * ah - inter);
        if (best_iou < iou) { best = a; best_iou = iou; }
      }
      if (best >= 0 && cls < NC) {
        float* y = Y + (((size_t)gy * G + gx) * NB + best) * D;
        y[0] = (float)cx; y[1] = (float)cy; y[2] = (float)w; y[3] = (float)h; y[4] = 1.f;
        y[5 + cls] = 1.f;
        float* t = T + (size_t)slot * 4;
        t[0] = (float)cx; t[1] = (float)cy; t[2] = (float)w; t[3] = (float)h;
        slot = (slot + 1) % TB;
      }
    }
  }
}

}  // namespace myolo

extern "C" int myolo_extract_bboxes(const unsigned char* gt_masks, int B, int S, int M, int* boxes, myolo_stream stream) {
  MYOLO_CHECK_ARG(gt_masks && boxes && B > 0 && S > 0 && M > 0);
  dim3 grid(M, B);
  myolo::extract_bboxes_kernel<<<grid, 256, 0, myolo::as_stream(stream)>>>(gt_masks, S, M, boxes);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

extern "C" int myolo_encode_yolo_targets(const int* gt_class_ids, const int* gt_boxes, int B, int M, int S, int G, int NB,
                                         int NC, int TB, const float* anchors, float* yolo_target, float* true_boxes,
                                         myolo_stream stream) {
  MYOLO_CHECK_ARG(gt_class_ids && gt_boxes && anchors && yolo_target && true_boxes);
  MYOLO_CHECK_ARG(B > 0 && M > 0 && S > 0 && G > 0 && NB > 0 && NC > 0 && TB > 0);
  cudaStream_t st = myolo::as_stream(stream);
  MYOLO_CUDA(cudaMemsetAsync(yolo_target, 0, (size_t)B * G * G * NB * (5 + NC) * sizeof(float), st));
  MYOLO_CUDA(cudaMemsetAsync(true_boxes, 0, (size_t)B * TB * 4 * sizeof(float), st));
  myolo::encode_yolo_targets_kernel<<<(B + 63) / 64, 64, 0, st>>>(gt_class_ids, gt_boxes, B, M, S, G, NB, NC, TB, anchors,
                                                                  yolo_target, true_boxes);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}
