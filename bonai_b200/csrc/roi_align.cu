// RoIAlign (avg, aligned=True, adaptive sampling grid) over an NHWC FPN pyramid, forward and
// backward, with the FPN level mapping fused in; plus the mask-target sampler that runs the same
// op over uint8 GT bitmaps.
// Replaces mmcv.ops.RoIAlign / roi_align [mmcv-full 1.0.5] at the reference call sites
//   roi_heads/roi_extractors/single_level_roi_extractor.py:32-80 (map_roi_levels + per-level op)
//   core/mask/structures.py:261-291 (BitmapMasks.crop_and_resize), core/mask/mask_target.py:31-62.
// Algorithm: SURVEY.md Appendix A (Detectron "aligned" RoIAlign).
#include "common.cuh"
#include "loft_b200.h"

namespace {

struct Pyramid {
  const float* feat[4];
  float* grad[4];
  int H[4], W[4];
  float scale[4];
  int num_levels;
};

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// floor(log2(sqrt(w*h)/finest + 1e-6)) clamped to [0, L-1]  (single_level_roi_extractor.py:47-50)
__device__ __forceinline__ int roi_level(float x1, float y1, float x2, float y2, float finest,
                                         int L) {
  const float s = sqrtf((x2 - x1) * (y2 - y1));
  const float l = floorf(log2f(s / finest + 1e-6f));
  int lv = (int)fminf(fmaxf(l, 0.f), (float)(L - 1));
  return lv;
}

// start + p*bin + (i+.5)*bin/g with the reference's operation order and no FMA contraction
__device__ __forceinline__ float sample_coord(float start, int p, float bin, int i, int g) {
  return __fadd_rn(__fadd_rn(start, __fmul_rn((float)p, bin)),
                   __fdiv_rn(__fmul_rn((float)i + 0.5f, bin), (float)g));
}

struct Tap {
  int lo, hi;
  float wl, wh;  // weight of lo / hi
  bool valid;
};

__device__ __forceinline__ Tap axis_tap(float y, int size) {
  Tap t;
  t.valid = !(y < -1.0f || y > (float)size);
  if (y <= 0.f) y = 0.f;
  int lo = (int)y;
  int hi;
  if (lo >= size - 1) {
    hi = lo = size - 1;
    y = (float)lo;
  } else {
    hi = lo + 1;
  }
  const float l = y - (float)lo;
  t.lo = lo;
  t.hi = hi;
  t.wh = l;
  t.wl = 1.f - l;
  return t;
}

// One block per RoI; blockDim = (C/4 float4 lanes, bins in flight), every thread row walks the
// RoI's S x S bins.
//
// The average over a bin's gh x gw bilinear samples is SEPARABLE: every sample's weight on pixel
// (Y, X) is wy(iy, Y) * wx(ix, X) and a sample is dropped iff its row OR its column is out of
// range, so
//     sum_{iy,ix} sum_{taps} w * f  =  sum_Y sum_X  Wy[Y] * Wx[X] * f[Y, X],
//     Wy[Y] = sum_iy wy(iy, Y),  Wx[X] = sum_ix wx(ix, X).
// Sample spacing is bin/g <= 1 pixel, so the pixels a bin touches are a contiguous window of at
// most (gh + 1) x (gw + 1): that many loads (or atomics, backward) per bin instead of 4 * gh * gw
// (16 instead of 36 for the common 3 x 3 grid), all independent of each other.  Wy depends only on
// the bin ROW and Wx only on the bin COLUMN: the 2 * S weight vectors of the RoI are built ONCE
// per block into shared memory (all the divisions, floors and range tests of the reference's
// sample loop), and the S x S bins then run a division-free loop of loads and FMAs.  Grids wider
// than kMaxWin - 1 samples (extreme aspect ratios) take the sample-by-sample path.
constexpr int kMaxWin = 16;
constexpr int kMaxS = 16;

struct RoiGeom {
  float wy[kMaxS][kMaxWin], wx[kMaxS][kMaxWin];
  int y0[kMaxS], ny[kMaxS], x0[kMaxS], nx[kMaxS];
};

template <bool kBackward>
__global__ void __launch_bounds__(256)
roi_align_kernel(Pyramid pyr, const float* __restrict__ rois, long long K, int S, int C,
                 float finest, float* __restrict__ out, const float* __restrict__ dout,
                 int* __restrict__ levels_out) {
  __shared__ RoiGeom G;
  const long long k = blockIdx.x;
  const float* r = rois + k * 5;
  const int b = (int)r[0];
  const float rx1 = r[1], ry1 = r[2], rx2 = r[3], ry2 = r[4];
  const int lv = roi_level(rx1, ry1, rx2, ry2, finest, pyr.num_levels);
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
  if (levels_out && tid == 0 && blockIdx.y == 0) levels_out[k] = lv;
  const float sc = pyr.scale[lv];
  const int H = pyr.H[lv], W = pyr.W[lv];
  const float x1 = __fsub_rn(__fmul_rn(rx1, sc), 0.5f), y1 = __fsub_rn(__fmul_rn(ry1, sc), 0.5f);
  const float x2 = __fsub_rn(__fmul_rn(rx2, sc), 0.5f), y2 = __fsub_rn(__fmul_rn(ry2, sc), 0.5f);
  const float rw = x2 - x1, rh = y2 - y1;
  const float bin_h = rh / (float)S, bin_w = rw / (float)S;
  const int gh = (int)ceilf(rh / (float)S), gw = (int)ceilf(rw / (float)S);
  const float count = fmaxf((float)(gh * gw), 1.f);
  const float inv_count = 1.f / count;
  const int c4 = threadIdx.x;  // float4 index along C
  const int C4 = C / 4;
  const bool windowed = gh < kMaxWin && gw < kMaxWin && gh > 0 && gw > 0 && S <= kMaxS;

  // ---- the RoI's 2 * S weight vectors: thread t -> (axis, bin index p, window position j)
  if (windowed) {
    for (int t = tid; t < 2 * S * kMaxWin; t += nthr) {
      const int j = t % kMaxWin;
      const int p = (t / kMaxWin) % S;
      const bool is_x = t >= S * kMaxWin;
      const int g = is_x ? gw : gh, size = is_x ? W : H;
      const float start = is_x ? x1 : y1, bsz = is_x ? bin_w : bin_h;
      const int base = axis_tap(sample_coord(start, p, bsz, 0, g), size).lo;
      const int last = axis_tap(sample_coord(start, p, bsz, g - 1, g), size).hi;
      float w = 0.f;
      if (j <= last - base) {
        for (int i = 0; i < g; ++i) {
          const Tap tp = axis_tap(sample_coord(start, p, bsz, i, g), size);
          if (!tp.valid) continue;
          if (tp.lo == base + j) w += tp.wl;
          if (tp.hi == base + j) w += tp.wh;
        }
      }
      if (is_x) {
        G.wx[p][j] = w;
        if (j == 0) {
          G.x0[p] = base;
          G.nx[p] = last - base + 1;
        }
      } else {
        G.wy[p][j] = w;
        if (j == 0) {
          G.y0[p] = base;
          G.ny[p] = last - base + 1;
        }
      }
    }
  }
  __syncthreads();

  const long long fbase = (long long)b * H * W * C4;
  // gridDim.y blocks share one RoI when there are too few RoIs to fill the machine
  for (int bin = blockIdx.y * blockDim.y + threadIdx.y; bin < S * S; bin += blockDim.y * gridDim.y) {
    const int ph = bin / S, pw = bin - ph * S;
    const long long obase = ((k * S + ph) * S + pw) * C;
    if (!kBackward) {
      const float4* f = reinterpret_cast<const float4*>(pyr.feat[lv]) + fbase;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (windowed) {
        const int ny = G.ny[ph], nx = G.nx[pw], y0 = G.y0[ph], x0 = G.x0[pw];
        for (int jy = 0; jy < ny; ++jy) {
          const float wy = G.wy[ph][jy];
          if (wy == 0.f) continue;
          const float4* row = f + ((long long)(y0 + jy) * W + x0) * C4 + c4;
#pragma unroll 4
          for (int jx = 0; jx < nx; ++jx) {
            const float w = wy * G.wx[pw][jx];
            const float4 v = row[(long long)jx * C4];
            acc.x += w * v.x;
            acc.y += w * v.y;
            acc.z += w * v.z;
            acc.w += w * v.w;
          }
        }
      } else {
        for (int iy = 0; iy < gh; ++iy) {
          const Tap ty = axis_tap(sample_coord(y1, ph, bin_h, iy, gh), H);
          for (int ix = 0; ix < gw; ++ix) {
            const Tap tx = axis_tap(sample_coord(x1, pw, bin_w, ix, gw), W);
            if (!(ty.valid && tx.valid)) continue;
            const float4 v1 = f[((long long)ty.lo * W + tx.lo) * C4 + c4];
            const float4 v2 = f[((long long)ty.lo * W + tx.hi) * C4 + c4];
            const float4 v3 = f[((long long)ty.hi * W + tx.lo) * C4 + c4];
            const float4 v4 = f[((long long)ty.hi * W + tx.hi) * C4 + c4];
            const float w1 = ty.wl * tx.wl, w2 = ty.wl * tx.wh, w3 = ty.wh * tx.wl,
                        w4 = ty.wh * tx.wh;
            acc.x += w1 * v1.x + w2 * v2.x + w3 * v3.x + w4 * v4.x;
            acc.y += w1 * v1.y + w2 * v2.y + w3 * v3.y + w4 * v4.y;
            acc.z += w1 * v1.z + w2 * v2.z + w3 * v3.z + w4 * v4.z;
            acc.w += w1 * v1.w + w2 * v2.w + w3 * v3.w + w4 * v4.w;
          }
        }
      }
      float4 o;
      o.x = tf32_rna(acc.x / count);
      o.y = tf32_rna(acc.y / count);
      o.z = tf32_rna(acc.z / count);
      o.w = tf32_rna(acc.w / count);
      reinterpret_cast<float4*>(out + obase)[c4] = o;
    } else {
      float4* g = reinterpret_cast<float4*>(pyr.grad[lv]) + fbase;
      float4 d = reinterpret_cast<const float4*>(dout + obase)[c4];
      d.x *= inv_count;
      d.y *= inv_count;
      d.z *= inv_count;
      d.w *= inv_count;
      if (windowed) {
        const int ny = G.ny[ph], nx = G.nx[pw], y0 = G.y0[ph], x0 = G.x0[pw];
        for (int jy = 0; jy < ny; ++jy) {
          const float wy = G.wy[ph][jy];
          if (wy == 0.f) continue;
          float4* row = g + ((long long)(y0 + jy) * W + x0) * C4 + c4;
          for (int jx = 0; jx < nx; ++jx) {
            const float w = wy * G.wx[pw][jx];
            if (w == 0.f) continue;
            atomicAdd(&row[(long long)jx * C4], make_float4(d.x * w, d.y * w, d.z * w, d.w * w));
          }
        }
      } else {
        for (int iy = 0; iy < gh; ++iy) {
          const Tap ty = axis_tap(sample_coord(y1, ph, bin_h, iy, gh), H);
          for (int ix = 0; ix < gw; ++ix) {
            const Tap tx = axis_tap(sample_coord(x1, pw, bin_w, ix, gw), W);
            if (!(ty.valid && tx.valid)) continue;
            const float w1 = ty.wl * tx.wl, w2 = ty.wl * tx.wh, w3 = ty.wh * tx.wl,
                        w4 = ty.wh * tx.wh;
            atomicAdd(&g[((long long)ty.lo * W + tx.lo) * C4 + c4],
                      make_float4(d.x * w1, d.y * w1, d.z * w1, d.w * w1));
            atomicAdd(&g[((long long)ty.lo * W + tx.hi) * C4 + c4],
                      make_float4(d.x * w2, d.y * w2, d.z * w2, d.w * w2));
            atomicAdd(&g[((long long)ty.hi * W + tx.lo) * C4 + c4],
                      make_float4(d.x * w3, d.y * w3, d.z * w3, d.w * w3));
            atomicAdd(&g[((long long)ty.hi * W + tx.hi) * C4 + c4],
                      make_float4(d.x * w4, d.y * w4, d.z * w4, d.w * w4));
          }
        }
      }
    }
  }
}

// mask targets: one thread per output bin; roi p samples bitmap masks[gt_inds[p]] (uint8, H x W),
// spatial_scale 1, aligned, adaptive grid, threshold >= 0.5 -> {0,1}.
__global__ void mask_target_kernel(const uint8_t* __restrict__ masks, const float* __restrict__ boxes,
                                   const long long* __restrict__ gt_inds, long long P, int S, int H,
                                   int W, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * S * S) return;
  const int pw = (int)(i % S);
  const int ph = (int)((i / S) % S);
  const long long p = i / ((long long)S * S);
  const float* bx = boxes + p * 4;
  // mask_target_single clips proposals to the image first (mask_target.py:48-51)
  const float cx1 = fminf(fmaxf(bx[0], 0.f), (float)W), cy1 = fminf(fmaxf(bx[1], 0.f), (float)H);
  const float cx2 = fminf(fmaxf(bx[2], 0.f), (float)W), cy2 = fminf(fmaxf(bx[3], 0.f), (float)H);
  const float x1 = cx1 - 0.5f, y1 = cy1 - 0.5f, x2 = cx2 - 0.5f, y2 = cy2 - 0.5f;
  const float rw = x2 - x1, rh = y2 - y1;
  const float bin_h = rh / (float)S, bin_w = rw / (float)S;
  const int gh = (int)ceilf(rh / (float)S), gw = (int)ceilf(rw / (float)S);
  const float count = fmaxf((float)(gh * gw), 1.f);
  const uint8_t* m = masks + gt_inds[p] * (long long)H * W;
  float acc = 0.f;
  for (int iy = 0; iy < gh; ++iy) {
    const float y = sample_coord(y1, ph, bin_h, iy, gh);
    const Tap ty = axis_tap(y, H);
    for (int ix = 0; ix < gw; ++ix) {
      const float x = sample_coord(x1, pw, bin_w, ix, gw);
      const Tap tx = axis_tap(x, W);
      if (!(ty.valid && tx.valid)) continue;
      const float v1 = (float)m[(long long)ty.lo * W + tx.lo];
      const float v2 = (float)m[(long long)ty.lo * W + tx.hi];
      const float v3 = (float)m[(long long)ty.hi * W + tx.lo];
      const float v4 = (float)m[(long long)ty.hi * W + tx.hi];
      const float w1 = __fmul_rn(ty.wl, tx.wl), w2 = __fmul_rn(ty.wl, tx.wh);
      const float w3 = __fmul_rn(ty.wh, tx.wl), w4 = __fmul_rn(ty.wh, tx.wh);
      const float sv = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w1, v1), __fmul_rn(w2, v2)),
                                           __fmul_rn(w3, v3)),
                                 __fmul_rn(w4, v4));
      acc = __fadd_rn(acc, sv);
    }
  }
  out[i] = (__fdiv_rn(acc, count) >= 0.5f) ? 1.f : 0.f;
}

int fill_pyramid(Pyramid& pyr, const float* const* feats, float* const* grads, const int* Hs,
                 const int* Ws, const float* scales, int L) {
  LOFT_CHECK_SHAPE(L >= 1 && L <= 4, "roi_align: 1..4 pyramid levels supported, got %d", L);
  pyr.num_levels = L;
  for (int i = 0; i < 4; ++i) {
    const int j = i < L ? i : L - 1;
    pyr.feat[i] = feats ? feats[j] : nullptr;
    pyr.grad[i] = grads ? grads[j] : nullptr;
    pyr.H[i] = Hs[j];
    pyr.W[i] = Ws[j];
    pyr.scale[i] = scales[j];
  }
  return LOFT_OK;
}

}  // namespace

extern "C" {

int loft_roi_align_fwd(const float* const* feats, const int* Hs, const int* Ws, const float* scales,
                       int num_levels, const float* rois, long long K, int S, int C,
                       float finest_scale, float* out, int* levels_out, cudaStream_t stream) {
  LOFT_CHECK_ARG(feats && rois && out, "roi_align_fwd: null pointer");
  LOFT_CHECK_SHAPE(C % 4 == 0 && C / 4 <= 256, "roi_align_fwd: C=%d must be a multiple of 4, <=1024", C);
  if (K == 0) return LOFT_OK;
  Pyramid pyr;
  int r = fill_pyramid(pyr, feats, nullptr, Hs, Ws, scales, num_levels);
  if (r) return r;
  const int tx = C / 4;
  int by = 256 / tx < 1 ? 1 : 256 / tx;
  if (by > S * S) by = S * S;
  dim3 block(tx, by);
  // ~8 blocks per SM in flight: split the bins of a RoI over several blocks if K alone is short
  int splits = (int)((8LL * loft_num_sms() + K - 1) / K);
  const int max_splits = (S * S + by - 1) / by;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  roi_align_kernel<false><<<dim3((unsigned)K, (unsigned)splits), block, 0, stream>>>(
      pyr, rois, K, S, C, finest_scale, out, nullptr, levels_out);
  LOFT_CUDA_LAUNCH_CHECK("roi_align_fwd");
  return LOFT_OK;
}

int loft_roi_align_bwd(float* const* grads, const int* Hs, const int* Ws, const float* scales,
                       int num_levels, const float* rois, long long K, int S, int C,
                       float finest_scale, const float* dout, cudaStream_t stream) {
  LOFT_CHECK_ARG(grads && rois && dout, "roi_align_bwd: null pointer");
  LOFT_CHECK_SHAPE(C % 4 == 0 && C / 4 <= 256, "roi_align_bwd: C=%d must be a multiple of 4, <=1024", C);
  if (K == 0) return LOFT_OK;
  Pyramid pyr;
  int r = fill_pyramid(pyr, nullptr, grads, Hs, Ws, scales, num_levels);
  if (r) return r;
  const int tx = C / 4;
  int by = 256 / tx < 1 ? 1 : 256 / tx;
  if (by > S * S) by = S * S;
  dim3 block(tx, by);
  // ~8 blocks per SM in flight: split the bins of a RoI over several blocks if K alone is short
  int splits = (int)((8LL * loft_num_sms() + K - 1) / K);
  const int max_splits = (S * S + by - 1) / by;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  roi_align_kernel<true><<<dim3((unsigned)K, (unsigned)splits), block, 0, stream>>>(
      pyr, rois, K, S, C, finest_scale, nullptr, dout, nullptr);
  LOFT_CUDA_LAUNCH_CHECK("roi_align_bwd");
  return LOFT_OK;
}

int loft_mask_target(const uint8_t* masks, const float* boxes, const long long* gt_inds, long long P,
                     int S, int H, int W, float* out, cudaStream_t stream) {
  LOFT_CHECK_ARG(masks && boxes && gt_inds && out, "mask_target: null pointer");
  if (P == 0) return LOFT_OK;
  const long long n = P * S * S;
  mask_target_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(masks, boxes, gt_inds, P, S, H,
                                                                      W, out);
  LOFT_CUDA_LAUNCH_CHECK("mask_target");
  return LOFT_OK;
}

}  // extern "C"
