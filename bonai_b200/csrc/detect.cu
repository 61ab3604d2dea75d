// Box-side kernels of the LOFT path: fused IoU + max-IoU assignment, RPN top-k decode
// (anchor generation + delta2bbox), batched NMS (bitmask + chunked scan), box / offset target
// encoders.  All index outputs are designed to be bit-exact against the reference arithmetic
// (fp32, no FMA contraction -- explicit _rn intrinsics).
// Reference: core/bbox/iou_calculators/iou2d_calculator.py:39-130,
// core/bbox/assigners/max_iou_assigner.py:127-212, core/anchor/anchor_generator.py:142-271,
// core/bbox/coder/delta_xywh_bbox_coder.py:74-197, dense_heads/rpn_head.py:79-168,
// mmcv.ops.batched_nms / nms [mmcv-full 1.0.5],
// core/bbox/coder/delta_xy_offset_coder.py:46-65,
// roi_heads/attribute_heads/offset_head_expand_feature.py:271-344.
#include "common.cuh"
#include "loft_b200.h"
#include <stdlib.h>

namespace {

__device__ __forceinline__ float box_iou_ref(float ax1, float ay1, float ax2, float ay2, float aarea,
                                             float bx1, float by1, float bx2, float by2,
                                             float barea, float eps) {
  // bbox_overlaps: wh = clamp(min(rb) - max(lt), 0); overlap = w*h; union = a1 + a2 - overlap;
  // union = max(union, eps); iou = overlap / union
  const float w = fmaxf(__fsub_rn(fminf(ax2, bx2), fmaxf(ax1, bx1)), 0.f);
  const float h = fmaxf(__fsub_rn(fminf(ay2, by2), fmaxf(ay1, by1)), 0.f);
  const float ov = __fmul_rn(w, h);
  const float un = fmaxf(__fsub_rn(__fadd_rn(aarea, barea), ov), eps);
  return __fdiv_rn(ov, un);
}

constexpr int kMaxGt = 1024;

// pass 1: per box max / argmax over gts, per gt max over boxes (atomicMax on the float bits)
__global__ void iou_max_kernel(const float* __restrict__ boxes, long long n,
                               const float* __restrict__ gts, int G, float* __restrict__ max_ov,
                               int* __restrict__ argmax, unsigned int* __restrict__ gt_max_bits) {
  extern __shared__ float sg[];  // G*5: x1,y1,x2,y2,area
  unsigned int* sgmax = reinterpret_cast<unsigned int*>(sg + G * 5);
  for (int i = threadIdx.x; i < G; i += blockDim.x) {
    const float x1 = gts[i * 4], y1 = gts[i * 4 + 1], x2 = gts[i * 4 + 2], y2 = gts[i * 4 + 3];
    sg[i * 5] = x1;
    sg[i * 5 + 1] = y1;
    sg[i * 5 + 2] = x2;
    sg[i * 5 + 3] = y2;
    sg[i * 5 + 4] = __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
    sgmax[i] = 0u;
  }
  __syncthreads();
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) {
    const float4 b = reinterpret_cast<const float4*>(boxes)[j];
    const float barea = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    float best = -1.f;
    int besti = 0;
    for (int i = 0; i < G; ++i) {
      const float iou = box_iou_ref(sg[i * 5], sg[i * 5 + 1], sg[i * 5 + 2], sg[i * 5 + 3],
                                    sg[i * 5 + 4], b.x, b.y, b.z, b.w, barea, 1e-6f);
      if (iou > best) {
        best = iou;
        besti = i;
      }
      if (iou > 0.f) atomicMax(&sgmax[i], __float_as_uint(iou));
    }
    max_ov[j] = G > 0 ? best : 0.f;
    argmax[j] = besti;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < G; i += blockDim.x)
    if (sgmax[i]) atomicMax(&gt_max_bits[i], sgmax[i]);
}

// pass 2: assignment incl. low-quality matches (later gts overwrite earlier ones)
__global__ void iou_assign_kernel(const float* __restrict__ boxes, long long n,
                                  const float* __restrict__ gts, int G,
                                  const float* __restrict__ max_ov, const int* __restrict__ argmax,
                                  const unsigned int* __restrict__ gt_max_bits, float pos_thr,
                                  float neg_thr, float min_pos, int match_low_quality,
                                  long long* __restrict__ gt_inds) {
  extern __shared__ float sg[];
  float* sgmax = sg + G * 5;
  for (int i = threadIdx.x; i < G; i += blockDim.x) {
    const float x1 = gts[i * 4], y1 = gts[i * 4 + 1], x2 = gts[i * 4 + 2], y2 = gts[i * 4 + 3];
    sg[i * 5] = x1;
    sg[i * 5 + 1] = y1;
    sg[i * 5 + 2] = x2;
    sg[i * 5 + 3] = y2;
    sg[i * 5 + 4] = __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
    sgmax[i] = __uint_as_float(gt_max_bits[i]);
  }
  __syncthreads();
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  if (G == 0) {
    gt_inds[j] = 0;
    return;
  }
  const float mo = max_ov[j];
  long long a = -1;
  if (mo >= 0.f && mo < neg_thr) a = 0;
  if (mo >= pos_thr) a = argmax[j] + 1;
  if (match_low_quality) {
    const float4 b = reinterpret_cast<const float4*>(boxes)[j];
    const float barea = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    for (int i = 0; i < G; ++i) {
      const float gm = sgmax[i];
      if (gm >= min_pos) {
        const float iou = box_iou_ref(sg[i * 5], sg[i * 5 + 1], sg[i * 5 + 2], sg[i * 5 + 3],
                                      sg[i * 5 + 4], b.x, b.y, b.z, b.w, barea, 1e-6f);
        if (iou == gm) a = i + 1;
      }
    }
  }
  gt_inds[j] = a;
}

// ---------------------------------------------------------------- RPN decode
// For rank r of level-local top-k index idx (into the (h,w,a) flattening): anchor + delta2bbox.
__global__ void rpn_decode_kernel(const float* __restrict__ head_out, int ld, int reg_off,
                                  const long long* __restrict__ topk_idx, int k, int fw, int A,
                                  const float* __restrict__ base_anchors, float stride,
                                  float max_ratio, float img_h, float img_w,
                                  float* __restrict__ boxes_out, long long head_stride,
                                  long long idx_stride, long long out_stride) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= k) return;
  head_out += (long long)blockIdx.y * head_stride;   // image blockIdx.y
  topk_idx += (long long)blockIdx.y * idx_stride;
  boxes_out += (long long)blockIdx.y * out_stride;
  const long long idx = topk_idx[r];
  const int a = (int)(idx % A);
  const long long pos = idx / A;
  const int w = (int)(pos % fw), h = (int)(pos / fw);
  const float sx = (float)w * stride, sy = (float)h * stride;
  const float ax1 = __fadd_rn(base_anchors[a * 4 + 0], sx), ay1 = __fadd_rn(base_anchors[a * 4 + 1], sy);
  const float ax2 = __fadd_rn(base_anchors[a * 4 + 2], sx), ay2 = __fadd_rn(base_anchors[a * 4 + 3], sy);
  const float* d = head_out + pos * ld + reg_off + a * 4;
  const float dx = d[0], dy = d[1];
  const float dw = fminf(fmaxf(d[2], -max_ratio), max_ratio);
  const float dh = fminf(fmaxf(d[3], -max_ratio), max_ratio);
  const float px = __fmul_rn(__fadd_rn(ax1, ax2), 0.5f), py = __fmul_rn(__fadd_rn(ay1, ay2), 0.5f);
  const float pw = __fsub_rn(ax2, ax1), ph = __fsub_rn(ay2, ay1);
  const float gw = __fmul_rn(pw, expf(dw)), gh = __fmul_rn(ph, expf(dh));
  const float gx = __fadd_rn(px, __fmul_rn(pw, dx)), gy = __fadd_rn(py, __fmul_rn(ph, dy));
  float x1 = __fsub_rn(gx, __fmul_rn(gw, 0.5f)), y1 = __fsub_rn(gy, __fmul_rn(gh, 0.5f));
  float x2 = __fadd_rn(gx, __fmul_rn(gw, 0.5f)), y2 = __fadd_rn(gy, __fmul_rn(gh, 0.5f));
  x1 = fminf(fmaxf(x1, 0.f), img_w);
  y1 = fminf(fmaxf(y1, 0.f), img_h);
  x2 = fminf(fmaxf(x2, 0.f), img_w);
  y2 = fminf(fmaxf(y2, 0.f), img_h);
  reinterpret_cast<float4*>(boxes_out)[r] = make_float4(x1, y1, x2, y2);
}

// ---------------------------------------------------------------- NMS
__global__ void fill_f32_kernel(float* p, float v) { *p = v; }

__global__ void max_coord_kernel(const float* __restrict__ boxes, long long n4,
                                 float* __restrict__ out) {
  // out must be initialised to -inf bits; boxes may be negative => use ordered-int trick
  float m = -INFINITY;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x)
    m = fmaxf(m, boxes[i]);
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) {
    int* addr = reinterpret_cast<int*>(out);
    int old = *addr, assumed;
    do {
      assumed = old;
      if (__int_as_float(assumed) >= m) break;
      old = atomicCAS(addr, assumed, __float_as_int(m));
    } while (assumed != old);
  }
}

// mask[i][w] bit b set  <=>  box j = w*64+b (j > i) is suppressed by box i.
// Boxes are pre-sorted by score (desc, stable).  Coordinates are offset by idx*(max_coord+1) in
// fp32 exactly like mmcv batched_nms, so boxes of different idx never overlap.
__global__ void nms_mask_kernel(const float* __restrict__ boxes, const long long* __restrict__ idxs,
                                const float* __restrict__ max_coord, int n, float thr,
                                unsigned long long* __restrict__ mask, int nwords) {
  const int row = blockIdx.y, colb = blockIdx.x;
  if (colb < row) return;
  const float off1 = max_coord ? __fadd_rn(*max_coord, 1.f) : 0.f;
  __shared__ float sb[64][5];
  const int t = threadIdx.x;
  const int cj = colb * 64 + t;
  if (cj < n) {
    const float o = idxs ? __fmul_rn((float)idxs[cj], off1) : 0.f;
    const float x1 = __fadd_rn(boxes[cj * 4 + 0], o), y1 = __fadd_rn(boxes[cj * 4 + 1], o);
    const float x2 = __fadd_rn(boxes[cj * 4 + 2], o), y2 = __fadd_rn(boxes[cj * 4 + 3], o);
    sb[t][0] = x1;
    sb[t][1] = y1;
    sb[t][2] = x2;
    sb[t][3] = y2;
    sb[t][4] = __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
  }
  __syncthreads();
  const int i = row * 64 + t;
  if (i >= n) return;
  const float o = idxs ? __fmul_rn((float)idxs[i], off1) : 0.f;
  const float x1 = __fadd_rn(boxes[i * 4 + 0], o), y1 = __fadd_rn(boxes[i * 4 + 1], o);
  const float x2 = __fadd_rn(boxes[i * 4 + 2], o), y2 = __fadd_rn(boxes[i * 4 + 3], o);
  const float area = __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
  unsigned long long bits = 0ull;
  const int jstart = (row == colb) ? t + 1 : 0;
  const int jend = min(64, n - colb * 64);
  for (int j = jstart; j < jend; ++j) {
    const float w = fmaxf(0.f, __fsub_rn(fminf(x2, sb[j][2]), fmaxf(x1, sb[j][0])));
    const float h = fmaxf(0.f, __fsub_rn(fminf(y2, sb[j][3]), fmaxf(y1, sb[j][1])));
    const float inter = __fmul_rn(w, h);
    const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area, sb[j][4]), inter));
    if (iou > thr) bits |= 1ull << j;
  }
  mask[(long long)i * nwords + colb] = bits;
}

// One block per image: greedy scan in 64-box chunks.  For chunk c the suppression word of the
// chunk is gathered lazily: OR over every box kept so far of mask[kept][c] (independent loads
// spread over the block, reduced with warp shuffles), then one thread resolves the 64 boxes of
// the chunk against the diagonal block.  No per-chunk update of a global "removed" bitmap.
constexpr int kScanThreads = 512;
constexpr int kMaxWords = 2048;  // n <= 131072
__global__ void __launch_bounds__(kScanThreads)
nms_scan_kernel(const unsigned long long* __restrict__ mask, int n, int nwords, int max_keep,
                long long* __restrict__ keep, int* __restrict__ num_keep, long long mask_stride,
                long long keep_stride) {
  __shared__ unsigned long long diag[64];
  __shared__ unsigned long long warp_or[kScanThreads / 32];
  __shared__ int s_count;
  const int img = blockIdx.x;
  mask += (long long)img * mask_stride;
  keep += (long long)img * keep_stride;
  const int t = threadIdx.x;
  if (t == 0) s_count = 0;
  __syncthreads();
  for (int c = 0; c < nwords; ++c) {
    const int cnt0 = s_count;
    if (cnt0 >= max_keep) break;
    unsigned long long acc = 0ull;
    for (int k = t; k < cnt0; k += kScanThreads) acc |= mask[keep[k] * (long long)nwords + c];
    if (t < 64) {
      const int i = c * 64 + t;
      diag[t] = (i < n) ? mask[(long long)i * nwords + c] : 0ull;
    }
    unsigned int lo = (unsigned int)acc, hi = (unsigned int)(acc >> 32);
    lo = __reduce_or_sync(0xffffffffu, lo);
    hi = __reduce_or_sync(0xffffffffu, hi);
    if ((t & 31) == 0) warp_or[t >> 5] = ((unsigned long long)hi << 32) | lo;
    __syncthreads();
    if (t == 0) {
      unsigned long long rem = 0ull;
      for (int w = 0; w < kScanThreads / 32; ++w) rem |= warp_or[w];
      int cnt = cnt0;
      const int lim = min(64, n - c * 64);
      for (int b = 0; b < lim && cnt < max_keep; ++b) {
        if (!((rem >> b) & 1ull)) {
          keep[cnt++] = c * 64 + b;
          rem |= diag[b];
        }
      }
      s_count = cnt;
    }
    __syncthreads();
  }
  if (t == 0) num_keep[img] = s_count;
}


// ---------------------------------------------------------------- segmented (per-level) NMS
// batched_nms never lets boxes of different idx suppress each other (their coordinates are
// pushed (max+1) apart), so the 12 768 x 12 768 pair matrix of the RPN is block diagonal: one
// block per FPN level.  Boxes arrive SEGMENT-MAJOR (all of level 0, then level 1, ...), each
// segment sorted by score; `order` lists them in global score order.  Per segment: pair mask and
// greedy scan (same kernels' arithmetic, same fp32 coordinate offset => identical decisions),
// then one compaction pass walks the global order and emits the first max_keep kept boxes.
constexpr int kMaxSeg = 16;
struct SegTable {
  int L;
  int off[kMaxSeg + 1];        // box offsets of the segments
  long long moff[kMaxSeg + 1]; // mask word offsets of the segments
};

__global__ void nms_mask_seg_kernel(const float* __restrict__ boxes, const float* __restrict__ maxc,
                                    int n, float thr, unsigned long long* __restrict__ ws,
                                    long long ws_stride_words, long long maxc_off_words,
                                    const SegTable tab) {
  const int s = blockIdx.z % tab.L, b = blockIdx.z / tab.L;
  const int ks = tab.off[s + 1] - tab.off[s];
  const int nw = (ks + 63) >> 6;
  const int row = blockIdx.y, colb = blockIdx.x;
  if (row >= nw || colb >= nw || colb < row) return;
  unsigned long long* wsb = ws + (long long)b * ws_stride_words;
  const float mx = *reinterpret_cast<const float*>(wsb + maxc_off_words);
  (void)maxc;
  const float off1 = __fadd_rn(mx, 1.f);
  const float o = __fmul_rn((float)s, off1);
  const float* bx = boxes + ((long long)b * n + tab.off[s]) * 4;
  unsigned long long* mask = wsb + tab.moff[s];
  __shared__ float sb[64][5];
  const int t = threadIdx.x;
  const int cj = colb * 64 + t;
  if (cj < ks) {
    const float x1 = __fadd_rn(bx[cj * 4 + 0], o), y1 = __fadd_rn(bx[cj * 4 + 1], o);
    const float x2 = __fadd_rn(bx[cj * 4 + 2], o), y2 = __fadd_rn(bx[cj * 4 + 3], o);
    sb[t][0] = x1;
    sb[t][1] = y1;
    sb[t][2] = x2;
    sb[t][3] = y2;
    sb[t][4] = __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
  }
  __syncthreads();
  const int i = row * 64 + t;
  if (i >= ks) return;
  const float x1 = __fadd_rn(bx[i * 4 + 0], o), y1 = __fadd_rn(bx[i * 4 + 1], o);
  const float x2 = __fadd_rn(bx[i * 4 + 2], o), y2 = __fadd_rn(bx[i * 4 + 3], o);
  const float area = __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
  unsigned long long bits = 0ull;
  const int jstart = (row == colb) ? t + 1 : 0;
  const int jend = min(64, ks - colb * 64);
  // The decision is the reference's fl(inter / union) > thr.  The IEEE division is ~40
  // instructions; a pair is only divided when inter is within 1e-4 (relative) of thr * union --
  // outside that band the rounded quotient cannot land on the other side of thr.
  const float thr_hi = thr * 1.0001f, thr_lo = thr * 0.9999f;
  for (int j = jstart; j < jend; ++j) {
    const float w = fmaxf(0.f, __fsub_rn(fminf(x2, sb[j][2]), fmaxf(x1, sb[j][0])));
    const float h = fmaxf(0.f, __fsub_rn(fminf(y2, sb[j][3]), fmaxf(y1, sb[j][1])));
    const float inter = __fmul_rn(w, h);
    const float uni = __fsub_rn(__fadd_rn(area, sb[j][4]), inter);
    bool over;
    if (uni > 0.f && inter > thr_hi * uni) over = true;
    else if (uni > 0.f && inter < thr_lo * uni) over = false;
    else over = __fdiv_rn(inter, uni) > thr;
    if (over) bits |= 1ull << j;
  }
  mask[(long long)i * nw + colb] = bits;
}

// One block per (image, segment): the greedy scan over the segment's own pair mask; kept boxes are
// flagged at their segment-major position.  Same decisions as nms_scan_kernel, different data
// flow: that kernel PULLS, per 64-box chunk, the mask words of every box kept so far (a gather
// over up to 3000 rows behind two dependent loads: 7 us per chunk, 330 us for the 47 chunks of an
// RPN level at random init, with 138 SMs idle behind it inside the forward graph); this one keeps
// the "suppressed" bitmap of the whole segment in shared memory and PUSHES into it the mask rows
// of the <= 64 boxes a chunk has just kept (independent loads, shared-memory atomicOr), while the
// chunk itself is resolved by one thread over the 64 diagonal words held in registers.
constexpr int kSegMaxWords = 2048;   // boxes per segment <= 131072
__global__ void __launch_bounds__(kScanThreads)
nms_scan_seg_kernel(unsigned long long* __restrict__ ws, long long ws_stride_words,
                    long long keep_off_words, long long flag_off_words, int n, int max_keep,
                    const SegTable tab) {
  __shared__ unsigned long long removed[kSegMaxWords];
  __shared__ unsigned long long diag[64];
  __shared__ int newkept[64];
  __shared__ int s_count, s_new;
  const int s = blockIdx.x % tab.L, b = blockIdx.x / tab.L;
  const int ks = tab.off[s + 1] - tab.off[s];
  const int nw = (ks + 63) >> 6;
  unsigned long long* wsb = ws + (long long)b * ws_stride_words;
  const unsigned long long* mask = wsb + tab.moff[s];
  int* keep = reinterpret_cast<int*>(wsb + keep_off_words) + tab.off[s];
  unsigned char* flags = reinterpret_cast<unsigned char*>(wsb + flag_off_words) + tab.off[s];
  const int lim_keep = min(max_keep, ks);
  const int t = threadIdx.x;
  for (int w = t; w < nw; w += kScanThreads) removed[w] = 0ull;
  if (t == 0) s_count = 0;
  unsigned long long dnext = 0ull;                 // diagonal word of box c*64 + t, one chunk ahead
  if (t < 64 && t < ks) dnext = mask[(long long)t * nw];
  __syncthreads();
  for (int c = 0; c < nw; ++c) {
    if (s_count >= lim_keep) break;
    if (t < 64) diag[t] = dnext;
    __syncthreads();
    if (t < 64) {                                  // prefetch the next chunk's diagonal
      const int i = (c + 1) * 64 + t;
      dnext = (c + 1 < nw && i < ks) ? mask[(long long)i * nw + c + 1] : 0ull;
    }
    if (t == 0) {
      unsigned long long d[64];
#pragma unroll
      for (int bb = 0; bb < 64; ++bb) d[bb] = diag[bb];
      unsigned long long rem = removed[c];
      const int lim = min(64, ks - c * 64);
      if (lim < 64) rem |= ~0ull << lim;           // positions past the end of the segment
      int cnt = s_count, nn = 0;
#pragma unroll
      for (int bb = 0; bb < 64; ++bb) {
        if (!((rem >> bb) & 1ull) && cnt < lim_keep) {
          keep[cnt++] = c * 64 + bb;
          newkept[nn++] = c * 64 + bb;
          rem |= d[bb];
        }
      }
      s_count = cnt;
      s_new = nn;
    }
    __syncthreads();
    const int nn = s_new, nrem = nw - (c + 1);
    for (int idx = t; idx < nn * nrem; idx += kScanThreads) {
      const int q = idx / nrem;
      const int w = c + 1 + (idx - q * nrem);
      const unsigned long long v = mask[(long long)newkept[q] * nw + w];
      if (v) atomicOr(&removed[w], v);
    }
    __syncthreads();
  }
  const int cnt = s_count;
  for (int k = t; k < cnt; k += kScanThreads) flags[keep[k]] = 1;
}

// One block per image: walk the global score order, keep[pos] = global rank of the pos-th kept box.
__global__ void __launch_bounds__(1024)
nms_compact_kernel(const unsigned long long* __restrict__ ws, long long ws_stride_words,
                   long long flag_off_words, const long long* __restrict__ order, int n,
                   int max_keep, long long* __restrict__ keep, int* __restrict__ num_keep) {
  __shared__ int wsum[32];
  __shared__ int s_base;
  const int b = blockIdx.x, t = threadIdx.x;
  const unsigned char* flags =
      reinterpret_cast<const unsigned char*>(ws + (long long)b * ws_stride_words + flag_off_words);
  const long long* ord = order + (long long)b * n;
  long long* kp = keep + (long long)b * n;
  if (t == 0) s_base = 0;
  __syncthreads();
  for (int g0 = 0; g0 < n; g0 += 1024) {
    const int g = g0 + t;
    const int f = (g < n) ? (int)flags[ord[g]] : 0;
    int incl = f;
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if ((t & 31) >= o) incl += v;
    }
    if ((t & 31) == 31) wsum[t >> 5] = incl;
    __syncthreads();
    if (t < 32) {
      int v = wsum[t];
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v, o);
        if (t >= o) v += u;
      }
      wsum[t] = v;
    }
    __syncthreads();
    const int base = s_base + ((t >> 5) ? wsum[(t >> 5) - 1] : 0);
    const int pos = base + incl - f;
    if (f && pos < max_keep) kp[pos] = g;
    __syncthreads();
    if (t == 1023) s_base = base + incl;
    __syncthreads();
    if (s_base >= max_keep) break;
  }
  if (t == 0) num_keep[b] = min(s_base, max_keep);
}

// ---------------------------------------------------------------- RCNN RoI sampling
// BaseSampler.sample + RandomSampler (core/bbox/samplers/base_sampler.py:34-101,
// random_sampler.py:31-75) for one image per block, after MaxIoUAssigner has labelled the
// proposals: candidates are [gt boxes (add_gt_as_proposals), proposals]; up to num_pos positives
// (assigned gt > 0) and then num - (#positives taken) negatives (assigned gt == 0) are drawn
// uniformly WITHOUT replacement and emitted in ascending candidate order (the reference's
// `.unique()` sorts them), positives first.  Replaces ~120 ATen launches per step
// (cat / where / nonzero / randperm / sort / index chains issued one by one from Python).
// A uniform subset of size m = the m smallest of independent random 64-bit keys (hash of a seed
// and the candidate index, made unique by the index in the low bits): radix select of the m-th
// smallest key, then an ordered compaction of everything not above it.
constexpr int kSampThreads = 1024;
constexpr int kSampPer = 4;                       // candidates per thread: n <= 4096
constexpr int kSampMaxOut = 1024;

__device__ __forceinline__ unsigned samp_hash(unsigned long long seed, unsigned j) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(j + 1u);   // splitmix64
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (unsigned)(z >> 32);
}

// m-th smallest (m >= 1) of the keys of class `cls` among this block's candidates; keys are
// unique.  Threads hold their candidates' (key, class) in registers.
__device__ unsigned long long samp_select(const unsigned long long (&key)[kSampPer],
                                          const int (&kls)[kSampPer], int cls, int m,
                                          unsigned* s_hist, unsigned long long* s_prefix,
                                          int* s_want) {
  const int t = threadIdx.x;
  if (t == 0) {
    *s_prefix = 0ull;
    *s_want = m;
  }
  __syncthreads();
  for (int pass = 0; pass < 8; ++pass) {           // 8 bits per pass, most significant first
    for (int i = t; i < 256; i += kSampThreads) s_hist[i] = 0u;
    __syncthreads();
    const int sh = 56 - 8 * pass;
    const unsigned long long prefix = *s_prefix;
    const unsigned long long pmask = pass == 0 ? 0ull : (~0ull << (sh + 8));
#pragma unroll
    for (int k = 0; k < kSampPer; ++k)
      if (kls[k] == cls && (key[k] & pmask) == prefix)
        atomicAdd(&s_hist[(unsigned)(key[k] >> sh) & 255u], 1u);
    __syncthreads();
    if (t == 0) {
      int want = *s_want, d = 0;
      unsigned c = s_hist[0];
      while ((int)c < want) {                      // walk up from the smallest digit
        want -= (int)c;
        ++d;
        c = s_hist[d];
      }
      *s_prefix = prefix | ((unsigned long long)d << sh);
      *s_want = want;
    }
    __syncthreads();
  }
  return *s_prefix;                                // all 64 bits fixed: the m-th smallest key
}

__global__ void __launch_bounds__(kSampThreads)
rcnn_sample_kernel(const float* __restrict__ props, long long prop_stride, int K, int prop_ld,
                   const int* __restrict__ num_valid, const long long* __restrict__ prop_gt_inds,
                   const float* __restrict__ gts, const int* __restrict__ gt_off, int num,
                   int num_pos_max, unsigned long long seed, long long* __restrict__ sel,
                   float* __restrict__ out_boxes, long long* __restrict__ out_gt,
                   unsigned char* __restrict__ out_isgt, int* __restrict__ cnt) {
  __shared__ unsigned s_hist[256];
  __shared__ unsigned long long s_prefix;
  __shared__ int s_want;
  __shared__ int s_warp[kSampThreads / 32];
  __shared__ int s_tot[2];
  const int img = blockIdx.x, t = threadIdx.x;
  const int g0 = gt_off[img], G = gt_off[img + 1] - g0;
  const int nv = num_valid ? min(num_valid[img], K) : K;
  const int n = G + K;
  const float* pr = props + (long long)img * prop_stride;
  const long long* pg = prop_gt_inds + (long long)img * K;
  // candidate j = t * kSampPer + k (consecutive per thread: the compaction keeps index order)
  unsigned long long key[kSampPer];
  int kls[kSampPer];                               // 1 positive, 0 negative, -1 neither
  long long gti[kSampPer];
  int npos = 0, nneg = 0;
#pragma unroll
  for (int k = 0; k < kSampPer; ++k) {
    const int j = t * kSampPer + k;
    kls[k] = -1;
    gti[k] = -1;
    key[k] = ~0ull;
    if (j < n) {
      long long gi;
      if (j < G) gi = j + 1;                       // a gt box is assigned to itself
      else gi = (j - G < nv) ? (G > 0 ? pg[j - G] : 0) : -1;     // padding rows are ignored
      gti[k] = gi;
      kls[k] = gi > 0 ? 1 : (gi == 0 ? 0 : -1);
      key[k] = ((unsigned long long)samp_hash(seed + (unsigned long long)img * 0x100000001B3ull,
                                              (unsigned)j) << 32) | (unsigned)j;
      npos += kls[k] == 1;
      nneg += kls[k] == 0;
    }
  }
  // block totals of both classes
  int v = (npos << 16) | nneg;                     // n <= 4096: both fit in 16 bits
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((t & 31) == 0) s_warp[t >> 5] = v;
  __syncthreads();
  if (t == 0) {
    int a = 0;
    for (int w = 0; w < kSampThreads / 32; ++w) a += s_warp[w];
    s_tot[0] = a >> 16;
    s_tot[1] = a & 0xffff;
  }
  __syncthreads();
  const int tot_pos = s_tot[0], tot_neg = s_tot[1];
  const int take_pos = min(tot_pos, num_pos_max);
  const int take_neg = min(tot_neg, num - take_pos);
  unsigned long long thr_pos = ~0ull, thr_neg = ~0ull;          // take everything of the class
  if (take_pos < tot_pos) thr_pos = samp_select(key, kls, 1, take_pos, s_hist, &s_prefix, &s_want);
  __syncthreads();
  if (take_neg < tot_neg && take_neg > 0)
    thr_neg = samp_select(key, kls, 0, take_neg, s_hist, &s_prefix, &s_want);
  __syncthreads();
  // ordered compaction: positives to [0, take_pos), negatives to [take_pos, take_pos + take_neg)
  int fp[kSampPer], fn[kSampPer], cp = 0, cn = 0;
#pragma unroll
  for (int k = 0; k < kSampPer; ++k) {
    fp[k] = kls[k] == 1 && take_pos > 0 && key[k] <= thr_pos;
    fn[k] = kls[k] == 0 && take_neg > 0 && key[k] <= thr_neg;
    cp += fp[k];
    cn += fn[k];
  }
  int pv = (cp << 16) | cn, incl = pv;             // inclusive warp scan of the packed counts
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, incl, o);
    if ((t & 31) >= o) incl += u;
  }
  if ((t & 31) == 31) s_warp[t >> 5] = incl;
  __syncthreads();
  if (t < 32) {
    int w = s_warp[t], wi = w;
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, wi, o);
      if (t >= o) wi += u;
    }
    s_warp[t] = wi - w;                            // exclusive prefix of the warp totals
  }
  __syncthreads();
  const int excl = incl - pv + s_warp[t >> 5];
  int op = excl >> 16, on = take_pos + (excl & 0xffff);
  long long* sel_i = sel + (long long)img * num;
  float* ob = out_boxes + (long long)img * num * 4;
  long long* og = out_gt + (long long)img * num;
  unsigned char* oi = out_isgt + (long long)img * num;
#pragma unroll
  for (int k = 0; k < kSampPer; ++k) {
    if (!(fp[k] || fn[k])) continue;
    const int j = t * kSampPer + k;
    const int o = fp[k] ? op++ : on++;
    float4 b;
    if (j < G) b = reinterpret_cast<const float4*>(gts)[g0 + j];
    else {
      const float* q = pr + (long long)(j - G) * prop_ld;
      b = make_float4(q[0], q[1], q[2], q[3]);
    }
    sel_i[o] = j;
    reinterpret_cast<float4*>(ob)[o] = b;
    og[o] = fp[k] ? gti[k] - 1 : -1;
    oi[o] = j < G ? 1 : 0;
  }
  if (t == 0) {
    cnt[img * 2 + 0] = take_pos;
    cnt[img * 2 + 1] = take_neg;
  }
}

// ---------------------------------------------------------------- RPN sampling + targets
// AnchorHead._get_targets_single after the assigner (anchor_head.py:206-278, allowed_border -1,
// sampling on) + RandomSampler (random_sampler.py:31-75, add_gt_as_proposals off) + images_to_levels
// (anchor_head.py:363-380) for the whole batch in three wide launches and NO host read-back: up to
// num_pos_max positives (assigned gt > 0), then num - (#positives taken) negatives (assigned gt ==
// 0) are drawn uniformly without replacement over the ~262 k anchors of each image, and labels /
// label weights / box targets / box weights are written directly in the per-level (n, h, w, a)
// order the loss kernel reads.  The reference path is ~80 small ATen launches and three host
// syncs per step (nonzero / randperm / index_put per image, the sample count for the loss scale).
// A uniform subset of size m = the m smallest of unique pseudo-random 64-bit keys (hash(seed, img,
// j) << 32 | j):  (1) histogram of the keys' top 12 bits per (image, class);  (2) every block
// derives the bin holding the m-th smallest key and appends that bin's keys (~64 of them) to a
// list;  (3) every block ranks the list to find the exact m-th key and writes the targets of its
// anchors.  The total sample count (the loss' avg_factor) is left on the device.
constexpr int kMaxRpnLevels = 8;                 // as in loss.cu (rpn_loss_fused)
constexpr int kRpnBins = 4096;
constexpr int kRpnListCap = 2048;
constexpr int kRpnThreads = 1024;

struct RpnMeta {            // per image
  int tot[2];               // candidates of class 0 (negative) / 1 (positive)
  int take[2];              // how many of them are sampled
  int bin[2];               // top-12-bit bin of the take-th smallest key (-1: take everything)
  int want[2];              // rank of that key within its bin (1-based)
};

struct RpnTargetArgs {
  const float* anchors;               // [A, 4], level-major
  const long long* gt_inds;           // [B, A]
  const float* gts;                   // [sum G, 4]
  const int* gt_off;                  // [B + 1]
  long long lvl_off[kMaxRpnLevels + 1];   // anchor offsets of the levels
  int n_levels, B;
  long long A;
  int num, num_pos_max;
  unsigned long long seed;
  float s0, s1, s2, s3, pos_weight;
  // workspace
  unsigned* hist;                     // [B][2][kRpnBins]
  unsigned long long* list;           // [B][2][kRpnListCap]
  int* list_n;                        // [B][2]
  RpnMeta* meta;                      // [B]
  int* err;                           // != 0: a list overflowed
  // outputs: level l occupies [B * lvl_off[l], B * lvl_off[l + 1]) in (image, anchor) order
  float *labels, *label_w, *bbox_t, *bbox_w;
  float* total;                       // [1]: sum over images of max(#pos, 1) + max(#neg, 1)
};

__device__ __forceinline__ unsigned long long rpn_key(unsigned long long seed, int img, long long j) {
  return ((unsigned long long)samp_hash(seed + (unsigned long long)img * 0x100000001B3ull,
                                        (unsigned)j) << 32) | (unsigned)j;
}

__global__ void __launch_bounds__(kRpnThreads) rpn_hist_kernel(const RpnTargetArgs a) {
  __shared__ unsigned s_hist[2 * kRpnBins];
  const int img = blockIdx.y, t = threadIdx.x;
  for (int i = t; i < 2 * kRpnBins; i += kRpnThreads) s_hist[i] = 0u;
  __syncthreads();
  const long long* gi = a.gt_inds + (long long)img * a.A;
  for (long long j = (long long)blockIdx.x * kRpnThreads + t; j < a.A;
       j += (long long)gridDim.x * kRpnThreads) {
    const long long g = gi[j];
    if (g < 0) continue;
    const int cls = g > 0 ? 1 : 0;
    atomicAdd(&s_hist[cls * kRpnBins + (unsigned)(rpn_key(a.seed, img, j) >> 52)], 1u);
  }
  __syncthreads();
  unsigned* gh = a.hist + (size_t)img * 2 * kRpnBins;
  for (int i = t; i < 2 * kRpnBins; i += kRpnThreads)
    if (s_hist[i]) atomicAdd(&gh[i], s_hist[i]);
}

// per (image, class): totals, how many to take, and the (bin, rank in bin) of the take-th key
__device__ void rpn_plan(const RpnTargetArgs& a, int img, unsigned* s_scan, int* s_warp,
                         RpnMeta* s_meta) {
  const int t = threadIdx.x;
  const unsigned* gh = a.hist + (size_t)img * 2 * kRpnBins;
  constexpr int kPer = kRpnBins / kRpnThreads;   // 4 bins per thread
  for (int cls = 1; cls >= 0; --cls) {           // positives first: they fix the negatives' quota
    unsigned v[kPer], sum = 0;
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      v[k] = gh[cls * kRpnBins + t * kPer + k];
      sum += v[k];
    }
    unsigned incl = sum;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned u = __shfl_up_sync(0xffffffffu, incl, o);
      if ((t & 31) >= o) incl += u;
    }
    if ((t & 31) == 31) s_warp[t >> 5] = (int)incl;
    __syncthreads();
    if (t < 32) {
      int w = s_warp[t], wi = w;
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, wi, o);
        if (t >= o) wi += u;
      }
      s_warp[t] = wi - w;
      if (t == 31) s_meta->tot[cls] = wi;
    }
    __syncthreads();
    unsigned excl = incl - sum + (unsigned)s_warp[t >> 5];   // candidates in bins before mine
    if (t == 0) {
      const int tot = s_meta->tot[cls];
      const int quota = cls == 1 ? a.num_pos_max : a.num - s_meta->take[1];
      s_meta->take[cls] = tot < quota ? tot : (quota > 0 ? quota : 0);
      s_meta->bin[cls] = -1;
      s_meta->want[cls] = 0;
    }
    __syncthreads();
    const int take = s_meta->take[cls];
    if (take > 0 && take < s_meta->tot[cls]) {
#pragma unroll
      for (int k = 0; k < kPer; ++k) {
        if ((int)excl < take && (int)(excl + v[k]) >= take) {   // exactly one bin satisfies this
          s_meta->bin[cls] = t * kPer + k;
          s_meta->want[cls] = take - (int)excl;
        }
        excl += v[k];
      }
    }
    __syncthreads();
  }
  (void)s_scan;
}

__global__ void __launch_bounds__(kRpnThreads) rpn_gather_kernel(const RpnTargetArgs a) {
  __shared__ int s_warp[32];
  __shared__ RpnMeta s_meta;
  const int img = blockIdx.y, t = threadIdx.x;
  rpn_plan(a, img, nullptr, s_warp, &s_meta);
  if (blockIdx.x == 0 && t == 0) {
    a.meta[img] = s_meta;
    atomicAdd(a.total, (float)((s_meta.take[1] > 1 ? s_meta.take[1] : 1) +
                               (s_meta.take[0] > 1 ? s_meta.take[0] : 1)));
  }
  const int bin0 = s_meta.bin[0], bin1 = s_meta.bin[1];
  if (bin0 < 0 && bin1 < 0) return;
  const long long* gi = a.gt_inds + (long long)img * a.A;
  for (long long j = (long long)blockIdx.x * kRpnThreads + t; j < a.A;
       j += (long long)gridDim.x * kRpnThreads) {
    const long long g = gi[j];
    if (g < 0) continue;
    const int cls = g > 0 ? 1 : 0;
    const int want_bin = cls ? bin1 : bin0;
    if (want_bin < 0) continue;
    const unsigned long long key = rpn_key(a.seed, img, j);
    if ((int)(key >> 52) != want_bin) continue;
    const int slot = atomicAdd(&a.list_n[img * 2 + cls], 1);
    if (slot < kRpnListCap) a.list[((size_t)img * 2 + cls) * kRpnListCap + slot] = key;
    else *a.err = 1;
  }
}

__global__ void __launch_bounds__(kRpnThreads) rpn_write_kernel(const RpnTargetArgs a) {
  __shared__ unsigned long long s_list[kRpnListCap];
  __shared__ unsigned long long s_thr[2];
  const int img = blockIdx.y, t = threadIdx.x;
  const RpnMeta m = a.meta[img];
  if (t < 2) s_thr[t] = ~0ull;                       // take every candidate of the class
  __syncthreads();
  for (int cls = 0; cls < 2; ++cls) {
    if (m.bin[cls] < 0) continue;
    const int n = min(a.list_n[img * 2 + cls], kRpnListCap);
    const unsigned long long* gl = a.list + ((size_t)img * 2 + cls) * kRpnListCap;
    for (int i = t; i < n; i += kRpnThreads) s_list[i] = gl[i];
    __syncthreads();
    for (int i = t; i < n; i += kRpnThreads) {       // the key of rank want-1 is the threshold
      const unsigned long long k = s_list[i];
      int r = 0;
      for (int q = 0; q < n; ++q) r += s_list[q] < k;
      if (r == m.want[cls] - 1) s_thr[cls] = k;
    }
    __syncthreads();
  }
  const unsigned long long thr0 = s_thr[0], thr1 = s_thr[1];
  const long long* gi = a.gt_inds + (long long)img * a.A;
  const float4* anc = reinterpret_cast<const float4*>(a.anchors);
  const float4* gtb = reinterpret_cast<const float4*>(a.gts) + a.gt_off[img];
  int lvl = 0;
  for (long long j = (long long)blockIdx.x * kRpnThreads + t; j < a.A;
       j += (long long)gridDim.x * kRpnThreads) {
    while (j >= a.lvl_off[lvl + 1]) ++lvl;           // j only grows
    const long long nl = a.lvl_off[lvl + 1] - a.lvl_off[lvl];
    const long long o = (long long)a.B * a.lvl_off[lvl] + (long long)img * nl + (j - a.lvl_off[lvl]);
    const long long g = gi[j];
    bool pos = false, neg = false;
    if (g >= 0) {
      const unsigned long long key = rpn_key(a.seed, img, j);
      if (g > 0) pos = m.take[1] > 0 && key <= thr1;
      else neg = m.take[0] > 0 && key <= thr0;
    }
    float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pos) {
      const float4 p = anc[j];
      const float4 q = gtb[g - 1];
      const float px = __fmul_rn(__fadd_rn(p.x, p.z), 0.5f), py = __fmul_rn(__fadd_rn(p.y, p.w), 0.5f);
      const float pw = __fsub_rn(p.z, p.x), ph = __fsub_rn(p.w, p.y);
      const float gx = __fmul_rn(__fadd_rn(q.x, q.z), 0.5f), gy = __fmul_rn(__fadd_rn(q.y, q.w), 0.5f);
      const float gw = __fsub_rn(q.z, q.x), gh = __fsub_rn(q.w, q.y);
      d.x = __fdiv_rn(__fdiv_rn(__fsub_rn(gx, px), pw), a.s0);
      d.y = __fdiv_rn(__fdiv_rn(__fsub_rn(gy, py), ph), a.s1);
      d.z = __fdiv_rn(logf(__fdiv_rn(gw, pw)), a.s2);
      d.w = __fdiv_rn(logf(__fdiv_rn(gh, ph)), a.s3);
    }
    const float bw = pos ? 1.f : 0.f;
    a.labels[o] = pos ? 1.f : 0.f;
    a.label_w[o] = pos ? a.pos_weight : (neg ? 1.f : 0.f);
    reinterpret_cast<float4*>(a.bbox_t)[o] = d;
    reinterpret_cast<float4*>(a.bbox_w)[o] = make_float4(bw, bw, bw, bw);
  }
}

// ---------------------------------------------------------------- soft-NMS (linear), test time
// mmcv.ops.soft_nms(method='linear') [mmcv-full 1.0.5, CPU-only there]: repeatedly select the
// highest-scoring live box, decay every other live box j by (1 - iou) if iou > thr, drop boxes whose
// score falls below min_score.  One block; scores live in shared memory.
constexpr int kSoftThreads = 1024;
constexpr int kSoftMax = 8192;
__global__ void __launch_bounds__(kSoftThreads)
soft_nms_kernel(const float* __restrict__ boxes, const float* __restrict__ scores_in,
                const long long* __restrict__ idxs, int n, float thr, float min_score, int max_keep,
                float* __restrict__ dets, long long* __restrict__ keep, int* __restrict__ num_keep) {
  __shared__ float sc[kSoftMax];
  __shared__ float red_v[kSoftThreads / 32];
  __shared__ int red_i[kSoftThreads / 32];
  __shared__ float s_max;
  __shared__ int s_arg;
  const int t = threadIdx.x;
  // class-aware offset like batched_nms: boxes + idx * (max_coord + 1)
  float off1 = 0.f;
  if (idxs != nullptr) {
    float m = -INFINITY;
    for (int i = t; i < n * 4; i += kSoftThreads) m = fmaxf(m, boxes[i]);
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((t & 31) == 0) red_v[t >> 5] = m;
    __syncthreads();
    if (t == 0) {
      float mm = -INFINITY;
      for (int w = 0; w < kSoftThreads / 32; ++w) mm = fmaxf(mm, red_v[w]);
      s_max = mm;
    }
    __syncthreads();
    off1 = __fadd_rn(s_max, 1.f);
    __syncthreads();
  }
  for (int i = t; i < n; i += kSoftThreads) sc[i] = scores_in[i];   // dead boxes get -inf
  __syncthreads();
  int count = 0;
  while (count < max_keep) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = t; i < n; i += kSoftThreads) {
      const float v = sc[i];
      if (v > bv || (v == bv && i < bi && v > -INFINITY)) {
        bv = v;
        bi = i;
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) {
        bv = ov;
        bi = oi;
      }
    }
    if ((t & 31) == 0) {
      red_v[t >> 5] = bv;
      red_i[t >> 5] = bi;
    }
    __syncthreads();
    if (t == 0) {
      float v = red_v[0];
      int a = red_i[0];
      for (int w = 1; w < kSoftThreads / 32; ++w)
        if (red_v[w] > v || (red_v[w] == v && red_i[w] < a)) {
          v = red_v[w];
          a = red_i[w];
        }
      s_max = v;
      s_arg = a;
    }
    __syncthreads();
    const float mv = s_max;
    const int ma = s_arg;
    if (!(mv > -INFINITY)) break;  // nothing alive
    const float o_m = idxs ? __fmul_rn((float)idxs[ma], off1) : 0.f;
    const float mx1 = __fadd_rn(boxes[ma * 4 + 0], o_m), my1 = __fadd_rn(boxes[ma * 4 + 1], o_m);
    const float mx2 = __fadd_rn(boxes[ma * 4 + 2], o_m), my2 = __fadd_rn(boxes[ma * 4 + 3], o_m);
    const float marea = __fmul_rn(__fsub_rn(mx2, mx1), __fsub_rn(my2, my1));
    if (t == 0) {
      dets[count * 5 + 0] = boxes[ma * 4 + 0];
      dets[count * 5 + 1] = boxes[ma * 4 + 1];
      dets[count * 5 + 2] = boxes[ma * 4 + 2];
      dets[count * 5 + 3] = boxes[ma * 4 + 3];
      dets[count * 5 + 4] = mv;
      keep[count] = ma;
    }
    for (int i = t; i < n; i += kSoftThreads) {
      if (i == ma) {
        sc[i] = -INFINITY;
        continue;
      }
      const float v = sc[i];
      if (!(v > -INFINITY)) continue;
      const float o_i = idxs ? __fmul_rn((float)idxs[i], off1) : 0.f;
      const float x1 = __fadd_rn(boxes[i * 4 + 0], o_i), y1 = __fadd_rn(boxes[i * 4 + 1], o_i);
      const float x2 = __fadd_rn(boxes[i * 4 + 2], o_i), y2 = __fadd_rn(boxes[i * 4 + 3], o_i);
      const float w = fmaxf(__fsub_rn(fminf(mx2, x2), fmaxf(mx1, x1)), 0.f);
      const float h = fmaxf(__fsub_rn(fminf(my2, y2), fmaxf(my1, y1)), 0.f);
      const float inter = __fmul_rn(w, h);
      const float area = __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
      const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(marea, area), inter));
      const float wgt = ovr > thr ? __fsub_rn(1.f, ovr) : 1.f;
      const float nv = __fmul_rn(v, wgt);
      sc[i] = nv < min_score ? -INFINITY : nv;
    }
    ++count;
    __syncthreads();
  }
  if (t == 0) *num_keep = count;
}

// Same arithmetic and tie-breaking for n <= 4096 (every inference call: <= 3000 candidates per
// image, test_cfg.rpn.max_num), restructured for latency: the selection loop is serial in the number of kept boxes, so
// what matters is the time of ONE iteration.  256 threads hold their <= 8 candidates (class-offset
// box, area, live score) in registers, the offset boxes also sit in shared memory for the
// broadcast read of the selected one; the decay of iteration k and the thread-local argmax for
// iteration k+1 are one pass; the block-wide argmax is a warp shuffle + 8 double-buffered shared
// entries that every thread scans itself: ONE __syncthreads per iteration, no global-memory reads
// on the critical path (the general kernel above: three barriers, a serial 32-entry scan by
// thread 0 and ~6 dependent global loads per iteration; 6.0 ms per tile at 1000 candidates).
constexpr int kSoftFastThreads = 256;
constexpr int kSoftFastMax = kSoftFastThreads * 16;
template <int kSoftFastPer>      // candidates per thread: 2, 4, 8 or 16 (n <= 512 ... 4096)
__global__ void __launch_bounds__(kSoftFastThreads)
soft_nms_fast_kernel(const float* __restrict__ boxes, const float* __restrict__ scores_in,
                     const long long* __restrict__ idxs, int n, float thr, float min_score,
                     int max_keep, float* __restrict__ dets, long long* __restrict__ keep,
                     int* __restrict__ num_keep) {
  extern __shared__ float4 s_box[];   // kSoftFastThreads * kSoftFastPer offset boxes
  __shared__ float red_v[2][kSoftFastThreads / 32];
  __shared__ int red_i[2][kSoftFastThreads / 32];
  __shared__ float s_max;
  const int t = threadIdx.x;
  float off1 = 0.f;
  if (idxs != nullptr) {
    float m = -INFINITY;
    for (int i = t; i < n * 4; i += kSoftFastThreads) m = fmaxf(m, boxes[i]);
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((t & 31) == 0) red_v[0][t >> 5] = m;
    __syncthreads();
    if (t == 0) {
      float mm = -INFINITY;
      for (int w = 0; w < kSoftFastThreads / 32; ++w) mm = fmaxf(mm, red_v[0][w]);
      s_max = mm;
    }
    __syncthreads();
    off1 = __fadd_rn(s_max, 1.f);
  }
  float v[kSoftFastPer], area[kSoftFastPer];
  float4 b[kSoftFastPer], orig[kSoftFastPer];
  float bv = -INFINITY;
  int bi = 0x7fffffff;
#pragma unroll
  for (int k = 0; k < kSoftFastPer; ++k) {
    const int i = t + kSoftFastThreads * k;
    v[k] = -INFINITY;
    b[k] = orig[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    area[k] = 0.f;
    if (i < n) {
      const float o_i = idxs ? __fmul_rn((float)idxs[i], off1) : 0.f;
      const float4 r = reinterpret_cast<const float4*>(boxes)[i];
      orig[k] = r;
      b[k] = make_float4(__fadd_rn(r.x, o_i), __fadd_rn(r.y, o_i), __fadd_rn(r.z, o_i),
                         __fadd_rn(r.w, o_i));
      area[k] = __fmul_rn(__fsub_rn(b[k].z, b[k].x), __fsub_rn(b[k].w, b[k].y));
      v[k] = scores_in[i];
      s_box[i] = b[k];
      if (v[k] > bv) {         // ascending i within a thread: the first maximum has the lowest index
        bv = v[k];
        bi = i;
      }
    }
  }
  int count = 0, par = 0;
  while (count < max_keep) {
    float rv = bv;
    int ri = bi;
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, rv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, ri, o);
      if (ov > rv || (ov == rv && oi < ri)) {
        rv = ov;
        ri = oi;
      }
    }
    if ((t & 31) == 0) {
      red_v[par][t >> 5] = rv;
      red_i[par][t >> 5] = ri;
    }
    __syncthreads();
    float mv = red_v[par][0];
    int ma = red_i[par][0];
#pragma unroll
    for (int w = 1; w < kSoftFastThreads / 32; ++w) {
      const float wv = red_v[par][w];
      const int wi = red_i[par][w];
      if (wv > mv || (wv == mv && wi < ma)) {
        mv = wv;
        ma = wi;
      }
    }
    if (!(mv > -INFINITY)) break;  // nothing alive
    const float4 mb = s_box[ma];
    const float marea = __fmul_rn(__fsub_rn(mb.z, mb.x), __fsub_rn(mb.w, mb.y));
    bv = -INFINITY;
    bi = 0x7fffffff;
#pragma unroll
    for (int k = 0; k < kSoftFastPer; ++k) {
      const int i = t + kSoftFastThreads * k;
      // branch-free over the 8 slots (the decays of a thread's candidates overlap; a dead slot
      // carries -inf through, a zero-area corner case only ever produces wgt = 1)
      const float x0 = v[k];
      const float w = fmaxf(__fsub_rn(fminf(mb.z, b[k].z), fmaxf(mb.x, b[k].x)), 0.f);
      const float h = fmaxf(__fsub_rn(fminf(mb.w, b[k].w), fmaxf(mb.y, b[k].y)), 0.f);
      const float inter = __fmul_rn(w, h);
      const float uni = __fsub_rn(__fadd_rn(marea, area[k]), inter);
      // The correctly rounded quotient is only needed where the overlap exceeds the threshold
      // (then 1 - ovr enters the score).  IEEE division is a ~50-instruction subroutine and the
      // selection loop is serial, so a 2-ulp reciprocal screens first: below thr * (1 - 2^-10)
      // the exact quotient cannot exceed thr and the weight is exactly 1.
      float wgt = 1.f;
      if (__fdividef(inter, uni) >= thr * 0.999f) {
        const float ovr = __fdiv_rn(inter, uni);
        wgt = ovr > thr ? __fsub_rn(1.f, ovr) : 1.f;
      }
      const float nv = __fmul_rn(x0, wgt);
      float x = (x0 > -INFINITY && !(nv < min_score)) ? nv : -INFINITY;
      if (i == ma) {             // the owner of the selected box emits it (from registers: a global
        x = -INFINITY;           // read here would be latency on every iteration)
        dets[count * 5 + 0] = orig[k].x;
        dets[count * 5 + 1] = orig[k].y;
        dets[count * 5 + 2] = orig[k].z;
        dets[count * 5 + 3] = orig[k].w;
        dets[count * 5 + 4] = mv;
        keep[count] = ma;
      }
      v[k] = x;
      if (x > bv) {
        bv = x;
        bi = i;
      }
    }
    ++count;
    par ^= 1;
  }
  if (t == 0) *num_keep = count;
}

// ---------------------------------------------------------------- target encoders
// bbox2delta with means 0: deltas[i] = ((gx-px)/pw, (gy-py)/ph, log(gw/pw), log(gh/ph)) / stds
__global__ void bbox_encode_kernel(const float* __restrict__ props, const float* __restrict__ gts,
                                   long long n, float s0, float s1, float s2, float s3,
                                   float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = reinterpret_cast<const float4*>(props)[i];
  const float4 g = reinterpret_cast<const float4*>(gts)[i];
  const float px = __fmul_rn(__fadd_rn(p.x, p.z), 0.5f), py = __fmul_rn(__fadd_rn(p.y, p.w), 0.5f);
  const float pw = __fsub_rn(p.z, p.x), ph = __fsub_rn(p.w, p.y);
  const float gx = __fmul_rn(__fadd_rn(g.x, g.z), 0.5f), gy = __fmul_rn(__fadd_rn(g.y, g.w), 0.5f);
  const float gw = __fsub_rn(g.z, g.x), gh = __fsub_rn(g.w, g.y);
  float4 d;
  d.x = __fdiv_rn(__fdiv_rn(__fsub_rn(gx, px), pw), s0);
  d.y = __fdiv_rn(__fdiv_rn(__fsub_rn(gy, py), ph), s1);
  d.z = __fdiv_rn(logf(__fdiv_rn(gw, pw)), s2);
  d.w = __fdiv_rn(logf(__fdiv_rn(gh, ph)), s3);
  reinterpret_cast<float4*>(out)[i] = d;
}

// FOA offset targets, closed form of the reference's polar rotation (verified equal):
//   branch 0: (x/pw, y/ph)/std   90: (y/ph, -x/pw)/std   180: (-x/pw, -y/ph)/std   270: (-y/ph, x/pw)/std
// out is [4P, 2] branch-major.
__global__ void offset_target_kernel(const float* __restrict__ props,
                                     const float* __restrict__ gt_offsets,
                                     const long long* __restrict__ gt_inds, long long P, float std_x,
                                     float std_y, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const float4 p = reinterpret_cast<const float4*>(props)[i];
  const float pw = __fsub_rn(p.z, p.x), ph = __fsub_rn(p.w, p.y);
  const long long g = gt_inds[i];
  const float x = __fdiv_rn(__fdiv_rn(gt_offsets[g * 2], pw), std_x);
  const float y = __fdiv_rn(__fdiv_rn(gt_offsets[g * 2 + 1], ph), std_y);
  out[(0 * P + i) * 2 + 0] = x;
  out[(0 * P + i) * 2 + 1] = y;
  out[(1 * P + i) * 2 + 0] = y;
  out[(1 * P + i) * 2 + 1] = -x;
  out[(2 * P + i) * 2 + 0] = -x;
  out[(2 * P + i) * 2 + 1] = -y;
  out[(3 * P + i) * 2 + 0] = -y;
  out[(3 * P + i) * 2 + 1] = x;
}

}  // namespace

extern "C" {

size_t loft_iou_assign_workspace(long long n, int G) {
  // max_ov[n] f32, argmax[n] i32, gt_max[G] u32
  return (size_t)n * 8 + (size_t)(G > 0 ? G : 1) * 4 + 64;
}

int loft_iou_assign(const float* boxes, long long n, const float* gts, int G, float pos_thr,
                    float neg_thr, float min_pos, int match_low_quality, long long* gt_inds,
                    float* max_overlaps, void* workspace, size_t ws_bytes, cudaStream_t stream) {
  LOFT_CHECK_ARG(gt_inds && max_overlaps && workspace, "iou_assign: null pointer");
  LOFT_CHECK_SHAPE(G <= kMaxGt, "iou_assign: at most %d gt boxes supported, got %d", kMaxGt, G);
  LOFT_CHECK_ARG(ws_bytes >= loft_iou_assign_workspace(n, G), "iou_assign: workspace too small");
  if (n == 0) return LOFT_OK;
  int* argmax = reinterpret_cast<int*>(workspace);
  unsigned int* gtmax = reinterpret_cast<unsigned int*>(argmax + n);
  cudaMemsetAsync(gtmax, 0, (size_t)(G > 0 ? G : 1) * 4, stream);
  const int threads = 256;
  const unsigned blocks = (unsigned)((n + threads - 1) / threads);
  const size_t smem = (size_t)(G > 0 ? G : 1) * 6 * sizeof(float);
  iou_max_kernel<<<blocks, threads, smem, stream>>>(boxes, n, gts, G, max_overlaps, argmax, gtmax);
  LOFT_CUDA_LAUNCH_CHECK("iou_max");
  iou_assign_kernel<<<blocks, threads, smem, stream>>>(boxes, n, gts, G, max_overlaps, argmax, gtmax,
                                                      pos_thr, neg_thr, min_pos, match_low_quality,
                                                      gt_inds);
  LOFT_CUDA_LAUNCH_CHECK("iou_assign");
  return LOFT_OK;
}

size_t loft_rpn_targets_workspace(int B) {
  size_t b = 64;                                             // total (float), err (int)
  b += (size_t)B * 2 * kRpnBins * sizeof(unsigned);          // hist
  b += (size_t)B * 2 * sizeof(int) + 64;                     // list_n
  b += (size_t)B * sizeof(RpnMeta) + 64;                     // meta
  b += (size_t)B * 2 * kRpnListCap * sizeof(unsigned long long);
  return b;
}

// RPN sampling + target construction for a batch (see rpn_hist_kernel).  gt_inds [B, A] is the
// assigner's output per image; lvl_off[n_levels + 1] are the anchor offsets of the pyramid levels;
// outputs are flat buffers of B*A (labels, label_w) and B*A*4 (bbox_t, bbox_w) floats in which
// level l occupies [B*lvl_off[l], B*lvl_off[l+1]) in (image, anchor) order; total[0] receives the
// loss' avg_factor (anchor_head.py:441-449).  Asynchronous; nothing is read back.
int loft_rpn_targets(const float* anchors, const long long* gt_inds, const float* gts,
                     const int* gt_off, const long long* lvl_off, int n_levels, int B, long long A,
                     int num, int num_pos_max, unsigned long long seed, float s0, float s1,
                     float s2, float s3, float pos_weight, float* labels, float* label_w,
                     float* bbox_t, float* bbox_w, float* total, void* workspace, size_t ws_bytes,
                     cudaStream_t stream) {
  LOFT_CHECK_ARG(anchors && gt_inds && gts && gt_off && lvl_off && labels && label_w && bbox_t &&
                     bbox_w && total && workspace,
                 "rpn_targets: null pointer");
  LOFT_CHECK_SHAPE(n_levels >= 1 && n_levels <= kMaxRpnLevels && B >= 1 && A >= 1 &&
                       A < (1ll << 32) && num >= 1 && num_pos_max >= 0 && lvl_off[n_levels] == A,
                   "rpn_targets: bad sizes levels=%d B=%d A=%lld num=%d", n_levels, B, A, num);
  LOFT_CHECK_ARG(ws_bytes >= loft_rpn_targets_workspace(B), "rpn_targets: workspace too small");
  LOFT_CHECK_ARG((reinterpret_cast<uintptr_t>(anchors) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(gts) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(bbox_t) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(bbox_w) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
                 "rpn_targets: buffers must be 16-byte aligned");
  RpnTargetArgs a{};
  a.anchors = anchors;
  a.gt_inds = gt_inds;
  a.gts = gts;
  a.gt_off = gt_off;
  for (int l = 0; l <= n_levels; ++l) a.lvl_off[l] = lvl_off[l];
  a.n_levels = n_levels;
  a.B = B;
  a.A = A;
  a.num = num;
  a.num_pos_max = num_pos_max;
  a.seed = seed;
  a.s0 = s0; a.s1 = s1; a.s2 = s2; a.s3 = s3;
  a.pos_weight = pos_weight;
  unsigned char* w = reinterpret_cast<unsigned char*>(workspace);
  a.err = reinterpret_cast<int*>(w + 16);
  size_t o = 64;
  a.hist = reinterpret_cast<unsigned*>(w + o);
  o += (size_t)B * 2 * kRpnBins * sizeof(unsigned);
  a.list_n = reinterpret_cast<int*>(w + o);
  o += ((size_t)B * 2 * sizeof(int) + 63) / 64 * 64;
  const size_t zero_bytes = o;                               // err, hist, list_n start at zero
  a.meta = reinterpret_cast<RpnMeta*>(w + o);
  o += ((size_t)B * sizeof(RpnMeta) + 63) / 64 * 64;
  a.list = reinterpret_cast<unsigned long long*>(w + o);
  a.labels = labels;
  a.label_w = label_w;
  a.bbox_t = bbox_t;
  a.bbox_w = bbox_w;
  a.total = total;
  cudaMemsetAsync(w, 0, zero_bytes, stream);
  cudaMemsetAsync(total, 0, sizeof(float), stream);
  const int per_img = (int)((A + 8 * kRpnThreads - 1) / (8 * kRpnThreads));   // ~8 anchors / thread
  dim3 grid(per_img < 1 ? 1 : (per_img > 64 ? 64 : per_img), B);
  rpn_hist_kernel<<<grid, kRpnThreads, 0, stream>>>(a);
  LOFT_CUDA_LAUNCH_CHECK("rpn_hist");
  rpn_gather_kernel<<<grid, kRpnThreads, 0, stream>>>(a);
  LOFT_CUDA_LAUNCH_CHECK("rpn_gather");
  rpn_write_kernel<<<grid, kRpnThreads, 0, stream>>>(a);
  LOFT_CUDA_LAUNCH_CHECK("rpn_write");
  return LOFT_OK;
}

// 1 if a candidate list of the last loft_rpn_targets call on this workspace overflowed (the
// sample is then short); a device -> host read, for tests.
int loft_rpn_targets_overflowed(const void* workspace, cudaStream_t stream) {
  int e = 0;
  cudaMemcpyAsync(&e, reinterpret_cast<const unsigned char*>(workspace) + 16, sizeof(int),
                  cudaMemcpyDeviceToHost, stream);
  cudaStreamSynchronize(stream);
  return e;
}

int loft_rpn_decode(const float* head_out, int ld, int reg_off, const long long* topk_idx, int k,
                    int fw, int A, const float* base_anchors, float stride, float max_ratio,
                    float img_h, float img_w, float* boxes_out, int batch, long long head_stride,
                    long long idx_stride, long long out_stride, cudaStream_t stream) {
  LOFT_CHECK_ARG(head_out && topk_idx && base_anchors && boxes_out, "rpn_decode: null pointer");
  if (k == 0 || batch == 0) return LOFT_OK;
  dim3 grid(loft_cdiv(k, 128), batch);
  rpn_decode_kernel<<<grid, 128, 0, stream>>>(head_out, ld, reg_off, topk_idx, k, fw, A,
                                              base_anchors, stride, max_ratio, img_h, img_w,
                                              boxes_out, head_stride, idx_stride, out_stride);
  LOFT_CUDA_LAUNCH_CHECK("rpn_decode");
  return LOFT_OK;
}

size_t loft_nms_workspace(int n) {
  const size_t nwords = (size_t)(n + 63) / 64;
  return (size_t)n * nwords * 8 + 64;
}

// boxes [B,n,4] sorted by score desc within each image; idxs [B,n] (may be NULL); per-image
// max_coord computed here.  keep [B,n] positions into the sorted order, num_keep [B].
int loft_nms_sorted(const float* boxes, const long long* idxs, int B, int n, float iou_thr,
                    int max_keep, long long* keep, int* num_keep, void* workspace, size_t ws_bytes,
                    cudaStream_t stream) {
  LOFT_CHECK_ARG(boxes && keep && num_keep && workspace, "nms_sorted: null pointer");
  const int nwords = (n + 63) / 64;
  LOFT_CHECK_SHAPE(nwords <= kMaxWords, "nms_sorted: n=%d too large", n);
  LOFT_CHECK_ARG(ws_bytes >= (size_t)B * loft_nms_workspace(n), "nms_sorted: workspace too small");
  if (n == 0 || B == 0) {
    if (B) cudaMemsetAsync(num_keep, 0, sizeof(int) * B, stream);
    return LOFT_OK;
  }
  const size_t per_img = loft_nms_workspace(n);
  for (int b = 0; b < B; ++b) {
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace) + (size_t)b * per_img;
    unsigned long long* mask = reinterpret_cast<unsigned long long*>(ws);
    float* maxc = reinterpret_cast<float*>(ws + (size_t)n * nwords * 8);
    const float* bx = boxes + (size_t)b * n * 4;
    const long long* ix = idxs ? idxs + (size_t)b * n : nullptr;
    if (ix) {
      fill_f32_kernel<<<1, 1, 0, stream>>>(maxc, -INFINITY);
      max_coord_kernel<<<32, 256, 0, stream>>>(bx, (long long)n * 4, maxc);
      LOFT_CUDA_LAUNCH_CHECK("max_coord");
    }
    dim3 grid(nwords, nwords);
    nms_mask_kernel<<<grid, 64, 0, stream>>>(bx, ix, ix ? maxc : nullptr, n, iou_thr, mask, nwords);
    LOFT_CUDA_LAUNCH_CHECK("nms_mask");
  }
  if (max_keep <= 0 || max_keep > n) max_keep = n;
  nms_scan_kernel<<<B, kScanThreads, 0, stream>>>(
      reinterpret_cast<const unsigned long long*>(workspace), n, nwords, max_keep, keep, num_keep,
      (long long)(per_img / 8), (long long)n);
  LOFT_CUDA_LAUNCH_CHECK("nms_scan");
  return LOFT_OK;
}

namespace {
int build_seg_table(const int* seg_off, int L, SegTable& tab, int n) {
  if (L < 1 || L > kMaxSeg || seg_off[0] != 0 || seg_off[L] != n) return -1;
  tab.L = L;
  long long m = 0;
  for (int s = 0; s <= L; ++s) tab.off[s] = seg_off[s];
  for (int s = 0; s < L; ++s) {
    const long long ks = seg_off[s + 1] - seg_off[s];
    if (ks < 0) return -1;
    tab.moff[s] = m;
    m += ks * ((ks + 63) / 64);
  }
  tab.moff[L] = m;
  return 0;
}
}  // namespace

// per-image workspace: [pair masks][keep lists n x i32][flags n x u8][max coord]
size_t loft_nms_segmented_workspace(const int* seg_off, int L) {
  SegTable tab;
  if (L < 1 || L > kMaxSeg) return 0;
  const int n = seg_off[L];
  if (build_seg_table(seg_off, L, tab, n)) return 0;
  const size_t keep_words = ((size_t)n * 4 + 7) / 8, flag_words = ((size_t)n + 7) / 8;
  return ((size_t)tab.moff[L] + keep_words + flag_words + 8) * 8;
}

int loft_nms_segmented(const float* boxes, const int* seg_off, int L, const long long* order, int B,
                       int n, float iou_thr, int max_keep, long long* keep, int* num_keep,
                       void* workspace, size_t ws_bytes, cudaStream_t stream) {
  LOFT_CHECK_ARG(boxes && seg_off && order && keep && num_keep && workspace,
                 "nms_segmented: null pointer");
  SegTable tab;
  LOFT_CHECK_SHAPE(build_seg_table(seg_off, L, tab, n) == 0,
                   "nms_segmented: bad segment table (L=%d, n=%d)", L, n);
  const size_t per_img = loft_nms_segmented_workspace(seg_off, L);
  LOFT_CHECK_ARG(ws_bytes >= (size_t)B * per_img, "nms_segmented: workspace too small");
  if (n == 0 || B == 0) {
    if (B) cudaMemsetAsync(num_keep, 0, sizeof(int) * B, stream);
    return LOFT_OK;
  }
  const long long stride_w = (long long)(per_img / 8);
  const long long keep_off = tab.moff[L];
  const long long flag_off = keep_off + (long long)(((size_t)n * 4 + 7) / 8);
  const long long maxc_off = flag_off + (long long)(((size_t)n + 7) / 8);
  unsigned long long* ws = reinterpret_cast<unsigned long long*>(workspace);
  int max_ks = 0;
  for (int s = 0; s < L; ++s) max_ks = max(max_ks, seg_off[s + 1] - seg_off[s]);
  const int max_nw = (max_ks + 63) / 64;
  LOFT_CHECK_SHAPE(max_nw <= kSegMaxWords, "nms_segmented: at most %d boxes per segment, got %d",
                   kSegMaxWords * 64, max_ks);
  for (int b = 0; b < B; ++b) {
    unsigned long long* wsb = ws + (long long)b * stride_w;
    float* maxc = reinterpret_cast<float*>(wsb + maxc_off);
    cudaMemsetAsync(wsb + flag_off, 0, (size_t)n, stream);
    fill_f32_kernel<<<1, 1, 0, stream>>>(maxc, -INFINITY);
    max_coord_kernel<<<32, 256, 0, stream>>>(boxes + (size_t)b * n * 4, (long long)n * 4, maxc);
    LOFT_CUDA_LAUNCH_CHECK("max_coord");
  }
  if (max_keep <= 0 || max_keep > n) max_keep = n;
  dim3 grid(max_nw, max_nw, B * L);
  nms_mask_seg_kernel<<<grid, 64, 0, stream>>>(boxes, nullptr, n, iou_thr, ws, stride_w, maxc_off,
                                               tab);
  LOFT_CUDA_LAUNCH_CHECK("nms_mask_seg");
  nms_scan_seg_kernel<<<B * L, kScanThreads, 0, stream>>>(ws, stride_w, keep_off, flag_off, n,
                                                         max_keep, tab);
  LOFT_CUDA_LAUNCH_CHECK("nms_scan_seg");
  nms_compact_kernel<<<B, 1024, 0, stream>>>(ws, stride_w, flag_off, order, n, max_keep, keep,
                                             num_keep);
  LOFT_CUDA_LAUNCH_CHECK("nms_compact");
  return LOFT_OK;
}

// RoI sampling of the RCNN stage for a batch of images, one block each (see rcnn_sample_kernel).
// props [B, K, prop_ld] (box in the first 4 columns), num_valid [B] (rows >= it are padding; NULL =
// all K), prop_gt_inds [B, K] from loft_iou_assign on the proposals, gts = all images' gt boxes
// [sum G, 4] with gt_off [B+1].  Outputs per image, `num` slots: sel (candidate index: gt boxes
// first, then proposals), out_boxes, out_gt (assigned gt index or -1), out_isgt, cnt [B,2] =
// (#positives, #negatives) -- positives first, each group in ascending candidate order.
int loft_rcnn_sample(const float* props, long long prop_stride, int K, int prop_ld,
                     const int* num_valid, const long long* prop_gt_inds, const float* gts,
                     const int* gt_off, int B, int max_gt, int num, int num_pos_max,
                     unsigned long long seed, long long* sel, float* out_boxes, long long* out_gt,
                     unsigned char* out_isgt, int* cnt, cudaStream_t stream) {
  LOFT_CHECK_ARG(props && prop_gt_inds && gt_off && sel && out_boxes && out_gt && out_isgt && cnt,
                 "rcnn_sample: null pointer");
  LOFT_CHECK_SHAPE(B >= 0 && K >= 0 && max_gt >= 0 && K + max_gt <= kSampThreads * kSampPer &&
                       num > 0 && num <= kSampMaxOut && num_pos_max >= 0 && num_pos_max <= num,
                   "rcnn_sample: bad sizes B=%d K=%d max_gt=%d num=%d (K + max_gt <= %d)", B, K,
                   max_gt, num, kSampThreads * kSampPer);
  LOFT_CHECK_ARG(max_gt == 0 || (gts && (reinterpret_cast<uintptr_t>(gts) & 15) == 0),
                 "rcnn_sample: gt boxes must be 16-byte aligned");
  if (B == 0) return LOFT_OK;
  rcnn_sample_kernel<<<B, kSampThreads, 0, stream>>>(props, prop_stride, K, prop_ld, num_valid,
                                                     prop_gt_inds, gts, gt_off, num, num_pos_max,
                                                     seed, sel, out_boxes, out_gt, out_isgt, cnt);
  LOFT_CUDA_LAUNCH_CHECK("rcnn_sample");
  return LOFT_OK;
}

// Linear soft-NMS of one image.  dets [n,5], keep [n] (selection order), num_keep [1].
int loft_soft_nms_linear(const float* boxes, const float* scores, const long long* idxs, int n,
                         float iou_thr, float min_score, int max_keep, float* dets, long long* keep,
                         int* num_keep, cudaStream_t stream) {
  LOFT_CHECK_ARG(boxes && scores && dets && keep && num_keep, "soft_nms_linear: null pointer");
  LOFT_CHECK_SHAPE(n <= kSoftMax, "soft_nms_linear: at most %d boxes, got %d", kSoftMax, n);
  if (max_keep <= 0 || max_keep > n) max_keep = n;
  if (n == 0) {
    cudaMemsetAsync(num_keep, 0, sizeof(int), stream);
    return LOFT_OK;
  }
  static int fast = -1;                       // LOFT_SOFT_NMS_FAST=0: the general kernel only (A/B)
  if (fast < 0) {
    const char* e = getenv("LOFT_SOFT_NMS_FAST");
    fast = (e == nullptr || e[0] != '0') ? 1 : 0;
  }
  if (fast && n <= kSoftFastMax && (reinterpret_cast<uintptr_t>(boxes) & 15) == 0) {
    static bool attr_set = false;
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute(soft_nms_fast_kernel<16>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           kSoftFastThreads * 16 * 16);
      if (e != cudaSuccess) {
        loft_set_error("soft_nms_linear: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
        return LOFT_ERR_CUDA;
      }
      attr_set = true;
    }
#define LOFT_SOFT_FAST(PER)                                                                    \
  soft_nms_fast_kernel<PER><<<1, kSoftFastThreads, kSoftFastThreads * PER * 16, stream>>>(     \
      boxes, scores, idxs, n, iou_thr, min_score, max_keep, dets, keep, num_keep)
    if (n <= 2 * kSoftFastThreads) LOFT_SOFT_FAST(2);
    else if (n <= 4 * kSoftFastThreads) LOFT_SOFT_FAST(4);
    else if (n <= 8 * kSoftFastThreads) LOFT_SOFT_FAST(8);
    else LOFT_SOFT_FAST(16);
#undef LOFT_SOFT_FAST
  } else
    soft_nms_kernel<<<1, kSoftThreads, 0, stream>>>(boxes, scores, idxs, n, iou_thr, min_score,
                                                     max_keep, dets, keep, num_keep);
  LOFT_CUDA_LAUNCH_CHECK("soft_nms_linear");
  return LOFT_OK;
}

int loft_bbox_encode(const float* props, const float* gts, long long n, float s0, float s1, float s2,
                     float s3, float* out, cudaStream_t stream) {
  LOFT_CHECK_ARG(props && gts && out, "bbox_encode: null pointer");
  if (n == 0) return LOFT_OK;
  bbox_encode_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(props, gts, n, s0, s1, s2, s3,
                                                                      out);
  LOFT_CUDA_LAUNCH_CHECK("bbox_encode");
  return LOFT_OK;
}

int loft_offset_target(const float* props, const float* gt_offsets, const long long* gt_inds,
                       long long P, float std_x, float std_y, float* out, cudaStream_t stream) {
  LOFT_CHECK_ARG(props && gt_offsets && gt_inds && out, "offset_target: null pointer");
  if (P == 0) return LOFT_OK;
  offset_target_kernel<<<(unsigned)((P + 127) / 128), 128, 0, stream>>>(props, gt_offsets, gt_inds, P,
                                                                        std_x, std_y, out);
  LOFT_CUDA_LAUNCH_CHECK("offset_target");
  return LOFT_OK;
}

}  // extern "C"
