// postprocess.cu -- inference post-processing on the device: what MaskYOLO.detect does
// in numpy after keras_model.predict (myolo/model.py:1290-1304, 1330-1391; myolo_utils.py:88-113 NMB,
// 883-912 unmold_mask): top-k detections by confidence, confidence threshold, greedy box suppression,
// and pasting each survivor's class-specific 28x28 soft mask into its pixel box (bilinear resize,
// threshold 0.5).  One block per image for the selection, one block per (detection, image) for the paste.
#include <math_constants.h>
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
  for (int r = tid; r < R; r += blockDim.x) sc[r] = D[(size_t)r * 6 + 4];
  if (tid == 0) ncand = 0;
  __syncthreads();
  for (int k = 0; k < top_k; ++k) {                 // k-th largest score, lowest index on ties
    float bv = -CUDART_INF_F;
    int bi = -1;
    for (int r = tid; r < R; r += blockDim.x) {
      const float v = sc[r];
      if (v > bv) { bv = v; bi = r; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi >= 0 && (bi < 0 || oi < bi))) { bv = ov; bi = oi; }
    }
    if (lane == 0) { rv[wp] = bv; ri[wp] = bi; }
    __syncthreads();
    if (tid == 0) {
      float v = rv[0];
      int i = ri[0];
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
        if (rv[w] > v || (rv[w] == v && ri[w] >= 0 && (i < 0 || ri[w] < i))) { v = rv[w]; i = ri[w]; }
      if (i >= 0 && v >= cs_thr) cand[ncand++] = i;  // detections below the confidence threshold are dropped
      if (i >= 0) sc[i] = -CUDART_INF_F;
    }
    __syncthreads();
  }
  if (tid == 0) {                                   // NMB: greedy, a box survives if IoU < thr with every kept one
    int kept = 0;
    int keep[kMaxTopK];
    for (int c = 0; c < ncand; ++c) {
      const float* p = D + (size_t)cand[c] * 6;
      bool ok = true;
      for (int j = 0; j < kept && ok; ++j) ok = box_iou(D + (size_t)keep[j] * 6, p) < nms_thr;
      if (ok) keep[kept++] = cand[c];
    }
    for (int j = 0; j < top_k; ++j) {
      const int o = b * top_k + j;
      if (j < kept) {
        const float* p = D + (size_t)keep[j] * 6;
        out_idx[o] = keep[j];
        for (int e = 0; e < 4; ++e) out_boxes[o * 4 + e] = min(max((int)rintf(p[e] * (float)S), 0), S);
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
    x1 = (int)rintf(p[0] * (float)S); y1 = (int)rintf(p[1] * (float)S);     // decode_masks: np.round(d[:4]*S), unclipped
    x2 = (int)rintf(p[2] * (float)S); y2 = (int)rintf(p[3] * (float)S);
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
