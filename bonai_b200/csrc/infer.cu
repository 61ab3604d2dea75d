// Test-time post-processing kernels of the LOFT path.
//
// loft_paste_masks -- FCNMaskHead.get_seg_masks + _do_paste_mask
// (mmdet/models/roi_heads/mask_heads/fcn_mask_head.py:151-308): every detection's 28x28 mask
// probability is resampled over its box and thresholded into a full-image bitmap.  The reference
// evaluates F.grid_sample over the WHOLE image for every detection in 1 GB chunks (N x H x W fp32
// grids + outputs); outside the box (+1 pixel) every bilinear tap falls outside the mask and the
// result is exactly 0, so only the window [floor(x0) - 1, ceil(x1) + 1) needs arithmetic: the
// caller zero-fills the uint8 output (one memset) and this kernel writes each detection's window,
// one block per (detection, band of rows), the sigmoid of the 28x28 logits held in shared memory.
//
// loft_offset_fusion_decode -- OffsetHeadExpandFeature.offset_fusion('max') + DeltaXYOffsetCoder
// .decode (attribute_heads/offset_head_expand_feature.py:346-448, core/bbox/coder/
// delta_xy_offset_coder.py:67-88): the four rotated branches' predictions -> one offset per
// detection, in one pass.
#include "common.cuh"
#include "loft_b200.h"

namespace {

constexpr int kMaskMax = 32;   // mask side (28 in the LOFT config)

// grid_sample(bilinear, padding zeros, align_corners=False) of prob[M][M] at normalised (gx, gy)
__device__ __forceinline__ float sample_mask(const float* __restrict__ prob, int M, float gx,
                                             float gy) {
  const float px = ((gx + 1.f) * (float)M - 1.f) * 0.5f;
  const float py = ((gy + 1.f) * (float)M - 1.f) * 0.5f;
  const float fx = floorf(px), fy = floorf(py);
  const int x0 = (int)fx, y0 = (int)fy;
  const float ax = px - fx, ay = py - fy;
  float v = 0.f;
  const bool x0ok = x0 >= 0 && x0 < M, x1ok = x0 + 1 >= 0 && x0 + 1 < M;
  if (y0 >= 0 && y0 < M) {
    if (x0ok) v += prob[y0 * M + x0] * (1.f - ax) * (1.f - ay);
    if (x1ok) v += prob[y0 * M + x0 + 1] * ax * (1.f - ay);
  }
  if (y0 + 1 >= 0 && y0 + 1 < M) {
    if (x0ok) v += prob[(y0 + 1) * M + x0] * (1.f - ax) * ay;
    if (x1ok) v += prob[(y0 + 1) * M + x0 + 1] * ax * ay;
  }
  return v;
}

__global__ void paste_masks_kernel(const float* __restrict__ logits, long long ld_n, int ld_px,
                                   const float* __restrict__ boxes, int ld_box, int N, int M,
                                   int img_h, int img_w, float thr, int rows_per_block,
                                   uint8_t* __restrict__ out) {
  __shared__ float prob[kMaskMax * kMaskMax];
  const int n = blockIdx.x;
  const float* bx = boxes + (long long)n * ld_box;
  const float x0 = bx[0], y0 = bx[1], x1 = bx[2], y1 = bx[3];
  // window of the CPU path of the reference (skip_empty=True, one detection per chunk)
  const int wx0 = max((int)floorf(x0) - 1, 0), wy0 = max((int)floorf(y0) - 1, 0);
  const int wx1 = min((int)ceilf(x1) + 1, img_w), wy1 = min((int)ceilf(y1) + 1, img_h);
  const int r0 = wy0 + blockIdx.y * rows_per_block;
  if (r0 >= wy1 || wx1 <= wx0) return;
  const int r1 = min(r0 + rows_per_block, wy1);
  for (int i = threadIdx.x; i < M * M; i += blockDim.x) {
    const float z = logits[(long long)n * ld_n + (long long)i * ld_px];
    prob[i] = 1.f / (1.f + expf(-z));
  }
  __syncthreads();
  const float sx = 2.f / (x1 - x0), sy = 2.f / (y1 - y0);
  const int ww = wx1 - wx0;
  const int total = (r1 - r0) * ww;
  uint8_t* o = out + (long long)n * img_h * img_w;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int iy = r0 + i / ww, ix = wx0 + i % ww;
    float gx = ((float)ix + 0.5f - x0) * sx - 1.f;    // (img - x0) / (x1 - x0) * 2 - 1
    float gy = ((float)iy + 0.5f - y0) * sy - 1.f;
    if (isinf(gx)) gx = 0.f;                          // zero-width boxes (fcn_mask_head.py:297-302)
    if (isinf(gy)) gy = 0.f;
    const float v = sample_mask(prob, M, gx, gy);
    o[(long long)iy * img_w + ix] = thr >= 0.f ? (v >= thr ? 1 : 0) : (uint8_t)(v * 255.f);
  }
}

__global__ void offset_fusion_decode_kernel(const float* __restrict__ pred, int ld, long long n,
                                            const float* __restrict__ boxes, int ld_box, float std_x,
                                            float std_y, float max_x, float max_y,
                                            float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // branch b of detection i is row b * n + i; branches 1 and 3 (90 / 270 degrees) swap x and y
  const float* p0 = pred + i * ld;
  const float* p1 = pred + (n + i) * ld;
  const float* p2 = pred + (2 * n + i) * ld;
  const float* p3 = pred + (3 * n + i) * ld;
  const float ax = fmaxf(fmaxf(fabsf(p0[0]), fabsf(p1[1])), fmaxf(fabsf(p2[0]), fabsf(p3[1])));
  const float ay = fmaxf(fmaxf(fabsf(p0[1]), fabsf(p1[0])), fmaxf(fabsf(p2[1]), fabsf(p3[0])));
  const float dx = (p0[0] > 0.f ? ax : -ax) * std_x;      // polarity of the un-rotated branch
  const float dy = (p0[1] > 0.f ? ay : -ay) * std_y;
  const float* b = boxes + i * ld_box;
  const float pw = b[2] - b[0], ph = b[3] - b[1];
  float gx = pw * dx, gy = ph * dy;
  if (max_x > 0.f) {
    gx = fminf(fmaxf(gx, -max_x), max_x);
    gy = fminf(fmaxf(gy, -max_y), max_y);
  }
  out[i * 2] = gx;
  out[i * 2 + 1] = gy;
}

}  // namespace

extern "C" {

int loft_paste_masks(const float* logits, long long ld_n, int ld_px, const float* boxes, int ld_box,
                     int N, int M, int img_h, int img_w, float thr, unsigned char* out,
                     cudaStream_t stream) {
  LOFT_CHECK_ARG(logits && boxes && out, "paste_masks: null pointer");
  LOFT_CHECK_SHAPE(M >= 1 && M <= kMaskMax && img_h > 0 && img_w > 0,
                   "paste_masks: mask side %d (<= %d), image %d x %d", M, kMaskMax, img_h, img_w);
  if (N == 0) return LOFT_OK;
  const int rows = 16;
  dim3 grid((unsigned)N, (unsigned)((img_h + 2 + rows - 1) / rows));
  paste_masks_kernel<<<grid, 256, 0, stream>>>(logits, ld_n, ld_px, boxes, ld_box, N, M, img_h,
                                               img_w, thr, rows, out);
  LOFT_CUDA_LAUNCH_CHECK("paste_masks");
  return LOFT_OK;
}

int loft_offset_fusion_decode(const float* pred, int ld, long long n, const float* boxes, int ld_box,
                              float std_x, float std_y, float max_x, float max_y, float* out,
                              cudaStream_t stream) {
  LOFT_CHECK_ARG(pred && boxes && out, "offset_fusion_decode: null pointer");
  if (n == 0) return LOFT_OK;
  offset_fusion_decode_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(
      pred, ld, n, boxes, ld_box, std_x, std_y, max_x, max_y, out);
  LOFT_CUDA_LAUNCH_CHECK("offset_fusion_decode");
  return LOFT_OK;
}

}  // extern "C"
