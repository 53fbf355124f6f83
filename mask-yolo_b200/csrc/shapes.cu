// shapes.cu -- the Shapes workload generated on the device.
//
// Replaces, for a whole batch in two launches, the host chain
//   ShapesDataset.load_image / load_mask / draw_shape   example/shapes/dataset_shapes.py:80-135 (cv2 rasterisation,
//                                                        occlusion of earlier shapes by later ones 112-116)
//   load_image_gt                                        myolo/myolo_utils.py:274-366 (instances whose visible mask is
//                                                        empty are dropped, the rest keep their order)
//   extract_bboxes                                       myolo/myolo_utils.py:247-271
//   BatchGenerator.__getitem__ padding / `image / 255.`  myolo/myolo_utils.py:821-851
// Input is the spec table the host generator draws (a few integers per image); outputs are the model's ground-truth
// inputs, byte for byte what the host chain produces.  The YOLO target / true-box tensors follow with
// myolo_encode_yolo_targets on the boxes written here.
//
// Every shape's pixel set is one column interval per row (shapes_extents.h, OpenCV's algorithms restated and pinned
// against cv2 on the CPU).  Because a later shape overwrites an earlier one in the image AND removes it from the
// earlier one's mask, each pixel has exactly one "owner" (the last shape covering it): image colour and mask channel
// both follow from the owner.
//   kernel 1 (one CTA per image): row extents per shape -> workspace; visible-pixel count and bounding box per shape;
//            compaction slots; class ids and boxes, zero padded.
//   kernel 2 (256 pixels per CTA): owner per pixel; image (fp32 /255 and/or bytes) and the M mask bytes per pixel staged
//            in shared memory and written out as full 16-byte / 4-byte words (HBM-write bound: 12 + M bytes per pixel).
#include "common.cuh"
#include "shapes_extents.h"

namespace myolo {

constexpr int kMaxShapes = 8;
constexpr int kSpecHead = 4, kSpecShape = 8;      // [bg r, g, b, n] + n x [type, r, g, b, x, y, s, 0]

__global__ void __launch_bounds__(128)
shapes_extents_kernel(const int* __restrict__ specs, int S, int MS, int M, int TB, int* __restrict__ ws,
                      int* __restrict__ gt_class_ids, int* __restrict__ gt_boxes, float* __restrict__ gt_boxes_f) {
  const int b = blockIdx.x, tid = threadIdx.x;
  const int* spec = specs + (size_t)b * (kSpecHead + kSpecShape * MS);
  int* ext = ws + (size_t)b * MS * 2 * S;                    // ws = [B][MS][S][2] extents, then [B][MS] slots
  int* slots = ws + (size_t)gridDim.x * MS * 2 * S + (size_t)b * MS;
  const int n = min(max(spec[3], 0), MS);
  __shared__ int cnt[kMaxShapes], bx[kMaxShapes][4];
  if (tid < kMaxShapes) { cnt[tid] = 0; bx[tid][0] = S; bx[tid][1] = S; bx[tid][2] = -1; bx[tid][3] = -1; }
  for (int k = tid; k < n * S; k += blockDim.x) { ext[2 * k] = S; ext[2 * k + 1] = -1; }
  __syncthreads();
  if (tid < n) {
    const int* sh = spec + kSpecHead + kSpecShape * tid;
    int* e = ext + (size_t)tid * 2 * S;
    myolo_shapes::shape_rows(e, e + 1, 2, S, S, sh[0], sh[4], sh[5], sh[6]);
  }
  __syncthreads();
  // visible pixels of shape i in row r: its interval minus the intervals of every later shape
  for (int k = tid; k < n * S; k += blockDim.x) {
    const int i = k / S, r = k - i * S;
    const int lo = ext[2 * k], hi = ext[2 * k + 1];
    int c_n = 0, c_first = S, c_last = -1;
    for (int c = lo; c <= hi; ++c) {
      bool vis = true;
      for (int j = i + 1; j < n; ++j) {
        const int* ej = ext + ((size_t)j * S + r) * 2;
        if (c >= ej[0] && c <= ej[1]) { vis = false; c = ej[1]; break; }   // skip the rest of the occluder's interval
      }
      if (vis) { ++c_n; c_first = min(c_first, c); c_last = c; }
    }
    if (c_n) {
      atomicAdd(&cnt[i], c_n);
      atomicMin(&bx[i][0], c_first); atomicMax(&bx[i][2], c_last);
      atomicMin(&bx[i][1], r);       atomicMax(&bx[i][3], r);
    }
  }
  __syncthreads();
  if (tid == 0) {
    int* ids = gt_class_ids + (size_t)b * TB;
    int* bo = gt_boxes + (size_t)b * TB * 4;
    int k = 0;
    for (int i = 0; i < MS; ++i) {
      int s = -1;
      if (i < n && cnt[i] > 0 && k < TB && k < M) {
        s = k++;
        ids[s] = spec[kSpecHead + kSpecShape * i];               // class id = shape type (dataset_shapes.py:61-63, 117)
        bo[4 * s + 0] = bx[i][0]; bo[4 * s + 1] = bx[i][1];
        bo[4 * s + 2] = bx[i][2] + 1; bo[4 * s + 3] = bx[i][3] + 1;
      }
      slots[i] = s;
    }
    for (; k < TB; ++k) { ids[k] = 0; bo[4 * k] = bo[4 * k + 1] = bo[4 * k + 2] = bo[4 * k + 3] = 0; }
    if (gt_boxes_f)
      for (int q = 0; q < 4 * TB; ++q) gt_boxes_f[(size_t)b * TB * 4 + q] = (float)bo[q];
  }
}

// dynamic shared memory: 256*M mask bytes (padded to 16) | 768 floats | 768 bytes
__global__ void __launch_bounds__(256)
shapes_paint_kernel(const int* __restrict__ specs, int B, int S, int MS, int M, const int* __restrict__ ws,
                    float* __restrict__ image_f32, unsigned char* __restrict__ image_u8,
                    unsigned char* __restrict__ gt_masks) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int tid = threadIdx.x;
  const size_t pix0 = (size_t)blockIdx.x * 256;                  // S*S is a multiple of 256: a CTA never straddles images
  const int b = (int)(pix0 / ((size_t)S * S));
  const int p = (int)(pix0 - (size_t)b * S * S) + tid;
  const int r = p / S, c = p - r * S;
  const int mask_bytes = 256 * M, mask_pad = (mask_bytes + 15) & ~15;
  unsigned char* s_mask = smem;
  float* s_img = reinterpret_cast<float*>(smem + mask_pad);
  unsigned char* s_u8 = smem + mask_pad + 768 * 4;
  __shared__ int s_col[kMaxShapes + 1][3], s_slot[kMaxShapes];
  __shared__ float s_norm[kMaxShapes + 1][3];
  const int* spec = specs + (size_t)b * (kSpecHead + kSpecShape * MS);
  const int* ext = ws + (size_t)b * MS * 2 * S;
  const int n = min(max(spec[3], 0), MS);
  if (tid < 3 * (n + 1)) {
    const int i = tid / 3, ch = tid - 3 * i;                      // i = 0: background, i >= 1: shape i-1
    const int v = (i == 0 ? spec[ch] : spec[kSpecHead + kSpecShape * (i - 1) + 1 + ch]) & 255;
    s_col[i][ch] = v;
    s_norm[i][ch] = __double2float_rn(__ddiv_rn((double)v, 255.0));   // numpy: uint8 / 255. in float64, stored as float32
  }
  if (tid < n) s_slot[tid] = ws[(size_t)B * MS * 2 * S + (size_t)b * MS + tid];
  for (int w = tid; w < mask_pad / 4; w += 256) reinterpret_cast<uint32_t*>(s_mask)[w] = 0u;
  __syncthreads();
  int owner = -1;
  for (int i = n - 1; i >= 0; --i) {
    const int2 e = *reinterpret_cast<const int2*>(ext + ((size_t)i * S + r) * 2);
    if (c >= e.x && c <= e.y) { owner = i; break; }
  }
  if (owner >= 0 && s_slot[owner] >= 0) s_mask[tid * M + s_slot[owner]] = 1;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    s_img[tid * 3 + ch] = s_norm[owner + 1][ch];
    s_u8[tid * 3 + ch] = (unsigned char)s_col[owner + 1][ch];
  }
  __syncthreads();
  {
    uint32_t* dst = reinterpret_cast<uint32_t*>(gt_masks + pix0 * M);          // 256*M bytes: a multiple of 4, 4 B aligned
    for (int w = tid; w < mask_bytes / 4; w += 256) dst[w] = reinterpret_cast<const uint32_t*>(s_mask)[w];
  }
  if (image_f32) {
    float4* dst = reinterpret_cast<float4*>(image_f32 + pix0 * 3);
    if (tid < 192) dst[tid] = reinterpret_cast<const float4*>(s_img)[tid];
  }
  if (image_u8) {
    uint32_t* dst = reinterpret_cast<uint32_t*>(image_u8 + pix0 * 3);
    if (tid < 192) dst[tid] = reinterpret_cast<const uint32_t*>(s_u8)[tid];
  }
}

}  // namespace myolo

extern "C" int myolo_shapes_raster(const int* specs, int B, int S, int MS, int M, int TB, int* ws, float* image_f32,
                                   unsigned char* image_u8, unsigned char* gt_masks, int* gt_class_ids, int* gt_boxes,
                                   float* gt_boxes_f, myolo_stream stream) {
  MYOLO_CHECK_ARG(specs && ws && gt_masks && gt_class_ids && gt_boxes);
  MYOLO_CHECK_ARG(B > 0 && S >= 16 && S % 16 == 0 && S <= 16384);
  MYOLO_CHECK_ARG(MS >= 1 && MS <= myolo::kMaxShapes && M >= MS && M <= 128 && TB >= MS);
  MYOLO_CHECK_ARG((reinterpret_cast<uintptr_t>(ws) & 7) == 0 && (reinterpret_cast<uintptr_t>(gt_masks) & 3) == 0);
  MYOLO_CHECK_ARG((reinterpret_cast<uintptr_t>(image_f32) & 15) == 0 && (reinterpret_cast<uintptr_t>(image_u8) & 3) == 0);
  cudaStream_t st = myolo::as_stream(stream);
  myolo::shapes_extents_kernel<<<B, 128, 0, st>>>(specs, S, MS, M, TB, ws, gt_class_ids, gt_boxes, gt_boxes_f);
  MYOLO_CHECK_LAUNCH();
  const size_t smem = (size_t)((256 * M + 15) & ~15) + 768 * 4 + 768;
  const long long blocks = (long long)B * S * S / 256;
  myolo::shapes_paint_kernel<<<(unsigned)blocks, 256, smem, st>>>(specs, B, S, MS, M, ws, image_f32, image_u8, gt_masks);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}
