// yolo_ops.cu -- the small / index-heavy ops of the hot path:
//   K12 DecodeYOLOLayer + DetectionsLayer   myolo/model.py:1442-1473, 1493-1538
//   K13 DetectMaskTargetLayer               myolo/model.py:457-602 (+ norm_boxes_graph 1394-1408,
//                                            overlaps_graph 420-454, trim_zeros_graph 1411-1420)
//   K15 yolo_custom_loss fwd+bwd            myolo/model.py:86-242
//   K16 myolo_mask_loss_graph fwd+bwd       myolo/model.py:718-754
// Compiled with -fmad=false: the IoU / partition / mask-target arithmetic must reproduce the
// reference's fp32 operation order bit for bit ("ROI index selection bit-exact").
#include <math_constants.h>
#include "common.cuh"
#include "crop.cuh"

namespace myolo {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// ------------------------------------------------------------------------------------------
// K12 decode
// ------------------------------------------------------------------------------------------
__global__ void yolo_decode_kernel(const float* __restrict__ y_pred, const float* __restrict__ anchors,
                                   float* __restrict__ boxes, float* __restrict__ det, int B, int GH, int GW, int NB,
                                   int NC) {
  const int total = B * GH * GW * NB;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int nb = i % NB;
  const int gx = (i / NB) % GW;
  const int gy = (i / (NB * GW)) % GH;
  const float* t = y_pred + (size_t)i * (5 + NC);
  const float gw = (float)GW;
  // cell[...,0] = column; cell[...,1] = transpose(cell_x) = row (reference requires GH == GW)
  const float x = (sigmoidf_(t[0]) + (float)gx) / gw;
  const float y = (sigmoidf_(t[1]) + (float)gy) / gw;
  const float w = expf(t[2]) * anchors[2 * nb + 0] / gw;
  const float h = expf(t[3]) * anchors[2 * nb + 1] / gw;
  const float hw = w / 2.f, hh = h / 2.f;
  const float x1 = x - hw, y1 = y - hh, x2 = x + hw, y2 = y + hh;
  if (boxes) *reinterpret_cast<float4*>(boxes + (size_t)i * 4) = make_float4(x1, y1, x2, y2);
  if (det) {
    float best = t[5];
    int bi = 0;
    for (int k = 1; k < NC; ++k)
      if (t[5 + k] > best) { best = t[5 + k]; bi = k; }
    float* d = det + (size_t)i * 6;
    d[0] = x1; d[1] = y1; d[2] = x2; d[3] = y2;
    d[4] = sigmoidf_(t[4]);
    d[5] = (float)bi;
  }
}

// ------------------------------------------------------------------------------------------
// K13 targets: one block per image.  IoU -> pos/neg flags -> stable partition (positives first,
// ascending proposal index inside each group) -> class ids + matched GT index.
// ------------------------------------------------------------------------------------------
constexpr int kMaxGT = 64;

__global__ void __launch_bounds__(256)
detect_targets_kernel(const float* __restrict__ proposals, const int* __restrict__ gt_class_ids,
                      const float* __restrict__ gt_boxes, int R, int M, int S, float* __restrict__ rois,
                      int* __restrict__ target_ids, int* __restrict__ n_pos, int* __restrict__ roi_src,
                      int* __restrict__ roi_gt) {
  extern __shared__ int smem_i[];
  __shared__ float gtb[kMaxGT][4];
  __shared__ int kept[kMaxGT];
  __shared__ int nkept, s_npos, s_nneg;
  int* flag = smem_i;          // [R] 1 = pos, 2 = neg, 0 = neither (NaN)
  int* assign = smem_i + R;    // [R] matched (original) GT index
  int* rank = smem_i + 2 * R;  // [R] rank inside its group
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* P = proposals + (size_t)b * R * 4;

  if (tid < M) {
    // norm_boxes_graph: (box - [0,0,1,1]) / ([S,S,S,S] - 1)
    const float sc = (float)S - 1.0f;
    const float* g = gt_boxes + ((size_t)b * M + tid) * 4;
    gtb[tid][0] = (g[0] - 0.f) / sc;
    gtb[tid][1] = (g[1] - 0.f) / sc;
    gtb[tid][2] = (g[2] - 1.f) / sc;
    gtb[tid][3] = (g[3] - 1.f) / sc;
  }
  __syncthreads();
  if (tid == 0) {
    int k = 0;
    for (int m = 0; m < M; ++m) {
      const float s = fabsf(gtb[m][0]) + fabsf(gtb[m][1]) + fabsf(gtb[m][2]) + fabsf(gtb[m][3]);
      if (s != 0.f) kept[k++] = m;  // trim_zeros_graph (NaN sum casts to True as well)
    }
    nkept = k;
  }
  __syncthreads();
  for (int r = tid; r < R; r += blockDim.x) {
    const float4 a = *reinterpret_cast<const float4*>(P + (size_t)r * 4);
    float best = -CUDART_INF_F;
    int bi = 0;
    bool nan = false;
    const float a_area = (a.w - a.y) * (a.z - a.x);
    for (int j = 0; j < nkept; ++j) {
      const int m = kept[j];
      const float bx1 = gtb[m][0], by1 = gtb[m][1], bx2 = gtb[m][2], by2 = gtb[m][3];
      const float x1 = fmaxf(a.x, bx1), y1 = fmaxf(a.y, by1);
      const float x2 = fminf(a.z, bx2), y2 = fminf(a.w, by2);
      const float inter = fmaxf(x2 - x1, 0.f) * fmaxf(y2 - y1, 0.f);
      const float b_area = (by2 - by1) * (bx2 - bx1);
      const float uni = (a_area + b_area) - inter;
      float iou = inter / uni;
      if (isnan(a.x) || isnan(a.y) || isnan(a.z) || isnan(a.w)) iou = CUDART_NAN_F;
      if (isnan(iou)) nan = true;
      if (iou > best) { best = iou; bi = m; }
    }
    int f = 0;
    if (!nan) f = (best >= 0.5f) ? 1 : 2;  // best = -inf when no GT -> negative
    flag[r] = f;
    assign[r] = bi;
  }
  __syncthreads();
  if (tid < 32) {
    int cp = 0, cn = 0;
    for (int base = 0; base < R; base += 32) {
      const int r = base + tid;
      const int f = (r < R) ? flag[r] : 0;
      const unsigned bp = __ballot_sync(0xffffffffu, f == 1);
      const unsigned bn = __ballot_sync(0xffffffffu, f == 2);
      const unsigned lt = (1u << tid) - 1u;
      if (f == 1) rank[r] = cp + __popc(bp & lt);
      if (f == 2) rank[r] = cn + __popc(bn & lt);
      cp += __popc(bp);
      cn += __popc(bn);
    }
    if (tid == 0) { s_npos = cp; s_nneg = cn; n_pos[b] = cp; }
  }
  __syncthreads();
  const int np = s_npos, nn = s_nneg;
  float* Ro = rois + (size_t)b * R * 4;
  int* ids = target_ids + (size_t)b * R;
  int* src = roi_src + (size_t)b * R;
  int* rgt = roi_gt + (size_t)b * R;
  for (int r = tid; r < R; r += blockDim.x) {
    const int f = flag[r];
    if (f == 0) continue;
    const int d = (f == 1) ? rank[r] : np + rank[r];
    *reinterpret_cast<float4*>(Ro + (size_t)d * 4) = *reinterpret_cast<const float4*>(P + (size_t)r * 4);
    ids[d] = (f == 1) ? gt_class_ids[(size_t)b * M + assign[r]] : 0;
    src[d] = r;
    rgt[d] = (f == 1) ? assign[r] : -1;
  }
  for (int d = np + nn + tid; d < R; d += blockDim.x) {
    *reinterpret_cast<float4*>(Ro + (size_t)d * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    ids[d] = 0;
    src[d] = -1;
    rgt[d] = -1;
  }
}

// mask targets: round(crop_and_resize(float(gt_mask[assign]), (y1,x1,y2,x2), MHxMW)) for positive
// rois, zeros elsewhere.  grid (R, B), block MH*MW threads.
__global__ void mask_targets_kernel(const float* __restrict__ rois, const int* __restrict__ roi_gt,
                                    const int* __restrict__ n_pos, const unsigned char* __restrict__ gt_masks, int R,
                                    int M, int S, int MH, int MW, float* __restrict__ out) {
  // M here = number of mask channels (MAX_GT_INSTANCES); a matched GT index beyond it has no mask
  const int j = blockIdx.x, b = blockIdx.y;
  const int t = threadIdx.x;
  if (t >= MH * MW) return;
  float* o = out + (((size_t)b * R + j) * MH) * MW + t;
  if (j >= n_pos[b]) { *o = 0.f; return; }
  const float4 bx = *reinterpret_cast<const float4*>(rois + ((size_t)b * R + j) * 4);  // x1,y1,x2,y2
  const int g = roi_gt[(size_t)b * R + j];
  const int y = t / MW, x = t % MW;
  const Sample sy = crop_coord(bx.y, bx.w, y, MH, S);
  const Sample sx = crop_coord(bx.x, bx.z, x, MW, S);
  float v = 0.f;
  if (sy.valid && sx.valid && g >= 0 && g < M) {
    const unsigned char* mb = gt_masks + (size_t)b * S * S * M + g;
    const float tl = mb[((size_t)sy.lo * S + sx.lo) * M] ? 1.f : 0.f;
    const float tr = mb[((size_t)sy.lo * S + sx.hi) * M] ? 1.f : 0.f;
    const float bl = mb[((size_t)sy.hi * S + sx.lo) * M] ? 1.f : 0.f;
    const float br = mb[((size_t)sy.hi * S + sx.hi) * M] ? 1.f : 0.f;
    v = lerp_rn(lerp_rn(tl, tr, sx.lerp), lerp_rn(bl, br, sx.lerp), sy.lerp);
  }
  *o = rintf(v);  // tf.round: half to even
}

// ------------------------------------------------------------------------------------------
// K16 mask loss
// ------------------------------------------------------------------------------------------
__global__ void count_pos_kernel(const int* __restrict__ ids, int n, double* __restrict__ ws) {
  int c = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) c += (ids[i] > 0);
  float s = warp_sum((float)c);
  if ((threadIdx.x & 31) == 0 && s != 0.f) atomicAdd(ws, (double)s);
}

// grid = n_roi blocks, block 256.  ws[0] = #positive rois, ws[1] = sum of BCE
__global__ void __launch_bounds__(256)
mask_loss_kernel(const float* __restrict__ masks, const float* __restrict__ tm, const int* __restrict__ ids, int HW,
                 int NC, float loss_weight, float* __restrict__ dlogit, double* __restrict__ ws) {
  const int r = blockIdx.x;
  const int cls = ids[r];
  const bool pos = cls > 0 && cls < NC;
  const double cnt = ws[0] * (double)HW;
  const float gscale = (cnt > 0.0) ? (float)((double)loss_weight / cnt) : 0.f;
  const float eps = 1e-7f;
  float lsum = 0.f;
  const float* mp = masks + (size_t)r * HW * NC;
  float* dp = dlogit ? dlogit + (size_t)r * HW * NC : nullptr;
  if (!pos) {
    if (dp)
      for (int i = threadIdx.x; i < HW * NC; i += blockDim.x) dp[i] = 0.f;
    return;
  }
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    const float p = mp[(size_t)i * NC + cls];
    const float y = tm[(size_t)r * HW + i];
    const float pc = fminf(fmaxf(p, eps), 1.f - eps);
    const float z = logf(pc / (1.f - pc));
    lsum += fmaxf(z, 0.f) - z * y + log1pf(expf(-fabsf(z)));
    if (dp) {
      for (int k = 0; k < NC; ++k) dp[(size_t)i * NC + k] = 0.f;
      const float inclip = (p >= eps && p <= 1.f - eps) ? 1.f : 0.f;
      // dl/dz = pc - y ; dz/dpc = 1/(pc(1-pc)) ; dp/dx = p(1-p)
      dp[(size_t)i * NC + cls] = gscale * inclip * (pc - y) / (pc * (1.f - pc)) * (p * (1.f - p));
    }
  }
  __shared__ float red[8];
  lsum = warp_sum(lsum);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = lsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += (double)red[i];
    atomicAdd(ws + 1, s);
  }
}

__global__ void mask_loss_finalize_kernel(const double* __restrict__ ws, int HW, float* __restrict__ loss_out) {
  const double cnt = ws[0] * (double)HW;
  loss_out[0] = cnt > 0.0 ? (float)(ws[1] / cnt) : 0.f;
}

// ------------------------------------------------------------------------------------------
// K15 yolo loss.  One thread per predictor (b, gy, gx, nb).
// ws: 0 nb_coord, 1 nb_conf, 2 nb_class, 3 sum_xy, 4 sum_wh, 5 sum_conf, 6 sum_class
// ------------------------------------------------------------------------------------------
struct YoloScales {
  float object, no_object, coord, cls;
};

struct YoloTerms {
  float pxy[2], pwh[2], pconf, txy[2], twh[2], obj, coord_mask, conf_mask, class_mask, tconf, iou, ce;
  int tcls;
  // intermediates for the IoU gradient
  float inter, uni;
};

__device__ __forceinline__ float iou_xywh(const float* pxy, const float* pwh, const float* bxy, const float* bwh,
                                          float* inter_o, float* uni_o) {
  float iw[2];
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    const float pmin = pxy[d] - pwh[d] / 2.f, pmax = pxy[d] + pwh[d] / 2.f;
    const float tmin = bxy[d] - bwh[d] / 2.f, tmax = bxy[d] + bwh[d] / 2.f;
    iw[d] = fmaxf(fminf(pmax, tmax) - fmaxf(pmin, tmin), 0.f);
  }
  const float inter = iw[0] * iw[1];
  const float uni = pwh[0] * pwh[1] + bwh[0] * bwh[1] - inter;
  if (inter_o) { *inter_o = inter; *uni_o = uni; }
  return inter / uni;
}

__device__ void yolo_terms(const float* __restrict__ yt, const float* __restrict__ yp, const float* __restrict__ tb,
                           const float* __restrict__ anchors, const float* __restrict__ cw, int gx, int gy, int nb,
                           int NC, int TB, YoloScales sc, int warmup, YoloTerms& o) {
  const float cell[2] = {(float)gx, (float)gy};
  o.pxy[0] = sigmoidf_(yp[0]) + cell[0];
  o.pxy[1] = sigmoidf_(yp[1]) + cell[1];
  o.pwh[0] = expf(yp[2]) * anchors[2 * nb + 0];
  o.pwh[1] = expf(yp[3]) * anchors[2 * nb + 1];
  o.pconf = sigmoidf_(yp[4]);
  o.txy[0] = yt[0]; o.txy[1] = yt[1]; o.twh[0] = yt[2]; o.twh[1] = yt[3];
  o.obj = yt[4];
  o.iou = iou_xywh(o.pxy, o.pwh, o.txy, o.twh, &o.inter, &o.uni);
  o.tconf = o.iou * o.obj;
  int tc = 0;
  float bv = yt[5];
  for (int k = 1; k < NC; ++k)
    if (yt[5 + k] > bv) { bv = yt[5 + k]; tc = k; }
  o.tcls = tc;
  o.coord_mask = o.obj * sc.coord;
  float best = -CUDART_INF_F;
  bool nan = false;
  for (int j = 0; j < TB; ++j) {
    const float v = iou_xywh(o.pxy, o.pwh, tb + 4 * j, tb + 4 * j + 2, nullptr, nullptr);
    if (isnan(v)) nan = true;
    best = fmaxf(best, v);
  }
  const float lt = (!nan && best < 0.6f) ? 1.f : 0.f;
  o.conf_mask = lt * (1.f - o.obj) * sc.no_object + o.obj * sc.object;
  o.class_mask = o.obj * cw[tc] * sc.cls;
  if (warmup) {
    const float nobox = (o.coord_mask < sc.coord / 2.f) ? 1.f : 0.f;
    o.txy[0] += (0.5f + cell[0]) * nobox;
    o.txy[1] += (0.5f + cell[1]) * nobox;
    o.twh[0] += anchors[2 * nb + 0] * nobox;
    o.twh[1] += anchors[2 * nb + 1] * nobox;
    o.coord_mask = 1.f;
  }
  // sparse softmax CE
  float mx = yp[5];
  for (int k = 1; k < NC; ++k) mx = fmaxf(mx, yp[5 + k]);
  float se = 0.f;
  for (int k = 0; k < NC; ++k) se += expf(yp[5 + k] - mx);
  o.ce = (logf(se) + mx) - yp[5 + tc];
}

__global__ void __launch_bounds__(128)
yolo_loss_reduce_kernel(const float* __restrict__ y_true, const float* __restrict__ y_pred,
                        const float* __restrict__ true_boxes, const float* __restrict__ anchors,
                        const float* __restrict__ cw, int B, int GH, int GW, int NB, int NC, int TB, YoloScales sc,
                        int warmup, double* __restrict__ ws) {
  const int total = B * GH * GW * NB;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float v[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (i < total) {
    const int nb = i % NB, gx = (i / NB) % GW, gy = (i / (NB * GW)) % GH, b = i / (NB * GW * GH);
    YoloTerms t;
    yolo_terms(y_true + (size_t)i * (5 + NC), y_pred + (size_t)i * (5 + NC), true_boxes + (size_t)b * TB * 4, anchors,
               cw, gx, gy, nb, NC, TB, sc, warmup, t);
    v[0] = t.coord_mask > 0.f ? 1.f : 0.f;
    v[1] = t.conf_mask > 0.f ? 1.f : 0.f;
    v[2] = t.class_mask > 0.f ? 1.f : 0.f;
    const float dx = t.txy[0] - t.pxy[0], dy = t.txy[1] - t.pxy[1];
    const float dw = t.twh[0] - t.pwh[0], dh = t.twh[1] - t.pwh[1];
    v[3] = (dx * dx + dy * dy) * t.coord_mask;
    v[4] = (dw * dw + dh * dh) * t.coord_mask;
    const float dc = t.tconf - t.pconf;
    v[5] = dc * dc * t.conf_mask;
    v[6] = t.ce * t.class_mask;
  }
  __shared__ double red[4][7];
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    const double s = warp_sum_d((double)v[k]);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = s;
  }
  __syncthreads();
  if (threadIdx.x < 7) {
    const double s = red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
    if (s != 0.0) atomicAdd(ws + threadIdx.x, s);
  }
}

__global__ void yolo_loss_finalize_kernel(const double* __restrict__ ws, float* __restrict__ loss_out) {
  const float nb_coord = (float)ws[0], nb_conf = (float)ws[1], nb_class = (float)ws[2];
  const float lxy = (float)ws[3] / (nb_coord + 1e-6f) / 2.f;
  const float lwh = (float)ws[4] / (nb_coord + 1e-6f) / 2.f;
  const float lcf = (float)ws[5] / (nb_conf + 1e-6f) / 2.f;
  const float lcl = (float)ws[6] / (nb_class + 1e-6f);
  loss_out[0] = lxy + lwh + lcf + lcl;
  loss_out[1] = lxy; loss_out[2] = lwh; loss_out[3] = lcf; loss_out[4] = lcl;
}

__global__ void __launch_bounds__(128)
yolo_loss_grad_kernel(const float* __restrict__ y_true, const float* __restrict__ y_pred,
                      const float* __restrict__ true_boxes, const float* __restrict__ anchors,
                      const float* __restrict__ cw, int B, int GH, int GW, int NB, int NC, int TB, YoloScales sc,
                      int warmup, float loss_weight, const double* __restrict__ ws, float* __restrict__ dyp) {
  const int total = B * GH * GW * NB;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int nb = i % NB, gx = (i / NB) % GW, gy = (i / (NB * GW)) % GH, b = i / (NB * GW * GH);
  const float* yt = y_true + (size_t)i * (5 + NC);
  const float* yp = y_pred + (size_t)i * (5 + NC);
  YoloTerms t;
  yolo_terms(yt, yp, true_boxes + (size_t)b * TB * 4, anchors, cw, gx, gy, nb, NC, TB, sc, warmup, t);
  const float inv_coord = loss_weight / ((float)ws[0] + 1e-6f);
  const float inv_conf = loss_weight / ((float)ws[1] + 1e-6f);
  const float inv_class = loss_weight / ((float)ws[2] + 1e-6f);
  float* g = dyp + (size_t)i * (5 + NC);
  // direct coordinate terms: d/dp [ (t-p)^2 * m / n / 2 ] = -(t-p) m / n
  float gpxy[2], gpwh[2];
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    gpxy[d] = -(t.txy[d] - t.pxy[d]) * t.coord_mask * inv_coord;
    gpwh[d] = -(t.twh[d] - t.pwh[d]) * t.coord_mask * inv_coord;
  }
  const float dconf = (t.tconf - t.pconf) * t.conf_mask * inv_conf;  // dL/d tconf ; dL/d pconf = -dconf
  // tconf = iou * obj depends on the prediction (TF does not stop the gradient)
  const float giou = dconf * t.obj;
  if (giou != 0.f) {
    // iou gradient against the ORIGINAL y_true box (tconf is computed before the warm-up shift)
    const float oxy[2] = {yt[0], yt[1]}, owh[2] = {yt[2], yt[3]};
    float iw[2], dmin[2], dmax[2];
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      const float pmin = t.pxy[d] - t.pwh[d] / 2.f, pmax = t.pxy[d] + t.pwh[d] / 2.f;
      const float tmin = oxy[d] - owh[d] / 2.f, tmax = oxy[d] + owh[d] / 2.f;
      const float hi = fminf(pmax, tmax), lo = fmaxf(pmin, tmin);
      const float raw = hi - lo;
      iw[d] = fmaxf(raw, 0.f);
      const float open = raw >= 0.f ? 1.f : 0.f;
      dmax[d] = (pmax <= tmax) ? open : 0.f;   // d iw / d pmax
      dmin[d] = (pmin >= tmin) ? -open : 0.f;  // d iw / d pmin
    }
    const float I = t.inter, U = t.uni;
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      const float other = iw[1 - d];
      const float dI_dxy = (dmax[d] + dmin[d]) * other;
      const float dI_dwh = (dmax[d] - dmin[d]) * 0.5f * other;
      const float dU_dxy = -dI_dxy;
      const float dU_dwh = t.pwh[1 - d] - dI_dwh;
      gpxy[d] += giou * (dI_dxy / U - I * dU_dxy / (U * U));
      gpwh[d] += giou * (dI_dwh / U - I * dU_dwh / (U * U));
    }
  }
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    const float s = t.pxy[d] - (d == 0 ? (float)gx : (float)gy);  // sigmoid(t)
    g[d] = gpxy[d] * s * (1.f - s);
    g[2 + d] = gpwh[d] * t.pwh[d];
  }
  g[4] = -dconf * t.pconf * (1.f - t.pconf);
  // class: (softmax - onehot) * class_mask / nb_class
  float mx = yp[5];
  for (int k = 1; k < NC; ++k) mx = fmaxf(mx, yp[5 + k]);
  float se = 0.f;
  for (int k = 0; k < NC; ++k) se += expf(yp[5 + k] - mx);
  const float cm = t.class_mask * inv_class;
  for (int k = 0; k < NC; ++k) g[5 + k] = (expf(yp[5 + k] - mx) / se - (k == t.tcls ? 1.f : 0.f)) * cm;
}

}  // namespace myolo

using namespace myolo;

extern "C" int myolo_yolo_decode(const float* y_pred, const float* anchors, float* boxes, float* detections, int B,
                                 int GH, int GW, int NB, int NC, myolo_stream stream) {
  MYOLO_CHECK_ARG(y_pred && anchors && (boxes || detections) && B > 0 && GH > 0 && GW > 0 && NB > 0 && NC > 0);
  MYOLO_CHECK_ARG(GH == GW);  // the reference builds cell_y as the transpose of cell_x (model.py:1447)
  const int total = B * GH * GW * NB;
  yolo_decode_kernel<<<(total + 127) / 128, 128, 0, as_stream(stream)>>>(y_pred, anchors, boxes, detections, B, GH, GW, NB, NC);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

extern "C" int myolo_detect_mask_targets(const float* proposals, const int* gt_class_ids, const float* gt_boxes,
                                         const unsigned char* gt_masks, int B, int R, int M, int MM, int S, int MH, int MW,
                                         float* rois, int* target_ids, float* target_masks, int* n_pos, int* roi_src,
                                         int* roi_gt, myolo_stream stream) {
  MYOLO_CHECK_ARG(proposals && gt_class_ids && gt_boxes && gt_masks && rois && target_ids && target_masks && n_pos && roi_src && roi_gt);
  MYOLO_CHECK_ARG(B > 0 && R > 0 && M > 0 && M <= kMaxGT && MM > 0 && S > 1 && MH > 0 && MW > 0 && MH * MW <= 1024);
  const size_t smem = (size_t)3 * R * sizeof(int);
  MYOLO_CHECK_ARG(smem <= 40 * 1024);
  cudaStream_t st = as_stream(stream);
  detect_targets_kernel<<<B, 256, smem, st>>>(proposals, gt_class_ids, gt_boxes, R, M, S, rois, target_ids, n_pos, roi_src, roi_gt);
  dim3 grid(R, B);
  const int threads = ((MH * MW + 31) / 32) * 32;
  mask_targets_kernel<<<grid, threads, 0, st>>>(rois, roi_gt, n_pos, gt_masks, R, MM, S, MH, MW, target_masks);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

extern "C" int myolo_mask_loss(const float* masks, const float* target_masks, const int* target_ids, int n_roi, int MH,
                               int MW, int NC, float loss_weight, float* loss_out, float* dlogit, double* ws,
                               myolo_stream stream) {
  MYOLO_CHECK_ARG(masks && target_masks && target_ids && loss_out && ws && n_roi > 0 && MH > 0 && MW > 0 && NC > 0);
  cudaStream_t st = as_stream(stream);
  MYOLO_CUDA(cudaMemsetAsync(ws, 0, 2 * sizeof(double), st));
  count_pos_kernel<<<min((n_roi + 255) / 256, 64), 256, 0, st>>>(target_ids, n_roi, ws);
  mask_loss_kernel<<<n_roi, 256, 0, st>>>(masks, target_masks, target_ids, MH * MW, NC, loss_weight, dlogit, ws);
  mask_loss_finalize_kernel<<<1, 1, 0, st>>>(ws, MH * MW, loss_out);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}

extern "C" int myolo_yolo_loss(const float* y_true, const float* y_pred, const float* true_boxes, const float* anchors,
                               const float* class_weights, int B, int GH, int GW, int NB, int NC, int TB,
                               const float* scales_host, int warmup, float loss_weight, float* loss_out,
                               float* dy_pred, double* ws, myolo_stream stream) {
  MYOLO_CHECK_ARG(y_true && y_pred && true_boxes && anchors && class_weights && scales_host && loss_out && ws);
  MYOLO_CHECK_ARG(B > 0 && GH > 0 && GW > 0 && NB > 0 && NC > 0 && TB > 0 && GH == GW);
  cudaStream_t st = as_stream(stream);
  YoloScales sc{scales_host[0], scales_host[1], scales_host[2], scales_host[3]};
  const int total = B * GH * GW * NB;
  MYOLO_CUDA(cudaMemsetAsync(ws, 0, 8 * sizeof(double), st));
  yolo_loss_reduce_kernel<<<(total + 127) / 128, 128, 0, st>>>(y_true, y_pred, true_boxes, anchors, class_weights, B, GH, GW, NB, NC, TB, sc, warmup, ws);
  yolo_loss_finalize_kernel<<<1, 1, 0, st>>>(ws, loss_out);
  if (dy_pred)
    yolo_loss_grad_kernel<<<(total + 127) / 128, 128, 0, st>>>(y_true, y_pred, true_boxes, anchors, class_weights, B, GH, GW, NB, NC, TB, sc, warmup, loss_weight, ws, dy_pred);
  MYOLO_CHECK_LAUNCH();
  return MYOLO_OK;
}
