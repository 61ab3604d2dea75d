// Device side of the reference's post-load training pipeline for BONAI tiles
// (configs/_base_/datasets/bonai_instance.py:5-17): RandomFlip -> Normalize -> Pad ->
// DefaultFormatBundle on the image, and RandomFlip -> Pad on the uint8 GT bitmaps, each as ONE
// pass over the bytes:
//   mmdet/datasets/pipelines/transforms.py:484-488 (mmcv.imflip), :655-676 (mmcv.imnormalize:
//   BGR->RGB, subtract mean, multiply by 1/std), :571-600 (pad to a multiple of 32 with 0),
//   formating.py:191-230 (HWC -> CHW), core/mask/structures.py:218-240 (BitmapMasks.flip / pad).
// Both are pure HBM streams: 3 B read + 12 B written per pixel, 1 B + 1 B per mask pixel.
#include "common.cuh"
#include "loft_b200.h"

namespace {

struct NormParams {
  float mean[3];
  double stdinv[3];
};

// one thread per 4 consecutive output pixels of one row (all 3 channels): 16-byte stores
__global__ void image_prep_kernel(const uint8_t* __restrict__ img, float* __restrict__ out, int H,
                                  int W, int Hp, int Wp, NormParams np, int to_rgb, int flip) {
  const int Wq = Wp >> 2;
  const long long total = (long long)Hp * Wq;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x0 = (int)(i % Wq) << 2;
    const int y = (int)(i / Wq);
    float v[3][4];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int k = 0; k < 4; ++k) v[c][k] = 0.f;                 // Pad: pad_val 0 after Normalize
    if (y < H) {
      const int ys = (flip == 2) ? H - 1 - y : y;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int x = x0 + k;
        if (x >= W) continue;
        const int xs = (flip == 1) ? W - 1 - x : x;
        const uint8_t* px = img + ((long long)ys * W + xs) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float raw = (float)px[to_rgb ? 2 - c : c];        // cv2.cvtColor(BGR2RGB)
          const float d = __fsub_rn(raw, np.mean[c]);             // cv2.subtract, float32
          v[c][k] = (float)((double)d * np.stdinv[c]);            // cv2.multiply, double scalar
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
      *reinterpret_cast<float4*>(out + ((long long)c * Hp + y) * Wp + x0) =
          make_float4(v[c][0], v[c][1], v[c][2], v[c][3]);
  }
}

// one thread per 16 consecutive output bytes of one mask row
__global__ void mask_flip_pad_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                     long long G, int H, int W, int Hp, int Wp, int flip,
                                     int aligned16) {
  const int Wq = Wp >> 4;
  const long long total = G * Hp * Wq;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x0 = (int)(i % Wq) << 4;
    const long long t = i / Wq;
    const int y = (int)(t % Hp);
    const long long g = t / Hp;
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    if (y < H) {
      const int ys = (flip == 2) ? H - 1 - y : y;
      const uint8_t* row = in + (g * H + ys) * (long long)W;
      if (aligned16) {
        // rows are 16-byte aligned and W % 16 == 0: one 16-byte load per thread; a horizontal
        // flip reads the mirrored chunk and reverses its bytes in registers
        if (x0 < W) {
          if (flip == 1) {
            const uint4 v = *reinterpret_cast<const uint4*>(row + (W - 16 - x0));
            w[0] = __byte_perm(v.w, 0, 0x0123);
            w[1] = __byte_perm(v.z, 0, 0x0123);
            w[2] = __byte_perm(v.y, 0, 0x0123);
            w[3] = __byte_perm(v.x, 0, 0x0123);
          } else {
            const uint4 v = *reinterpret_cast<const uint4*>(row + x0);
            w[0] = v.x;
            w[1] = v.y;
            w[2] = v.z;
            w[3] = v.w;
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const int x = x0 + k;
          if (x < W) {
            const int xs = (flip == 1) ? W - 1 - x : x;
            w[k >> 2] |= (uint32_t)row[xs] << (8 * (k & 3));
          }
        }
      }
    }
    *reinterpret_cast<uint4*>(out + (g * Hp + y) * (long long)Wp + x0) =
        make_uint4(w[0], w[1], w[2], w[3]);
  }
}


// ---------------------------------------------------------------------------------------------
// Polygon -> bitmap (LoadAnnotations(poly2mask=True), pipelines/loading.py:301-326,345-368: every
// BONAI building mask is `maskUtils.decode(maskUtils.merge(maskUtils.frPyObjects(polys, h, w)))`).
// The rasteriser is the one of pycocotools 2.0.x (common/maskApi.c: rleFrPoly / rleMerge /
// rleDecode), restated so the bitmaps are the reference's bit for bit:
//   1. vertices are scaled by 5 and rounded ((int)(5*x + .5), C truncation);
//   2. every edge is walked one unit step along its longer axis (endpoints swapped so the walk
//      ascends, the slope a double, positions rounded with +.5 and truncation);
//   3. wherever the upsampled column index changes between consecutive boundary points, the point
//      is mapped back to pixel units; it is kept only if it lies exactly on a pixel column inside
//      the image, and becomes the run boundary  a = x*h + ceil(clamp(y, 0, h))  of the
//      COLUMN-major run-length code;
//   4. pixel i (column-major) is set iff an odd number of boundaries are <= i; parts are OR-ed.
// Kernel 1 finds and sorts the run boundaries of one polygon part per block, kernel 2 writes the
// row-major uint8 bitmap, one thread per image column (a warp stores 32 consecutive bytes of a
// row; every output byte is written exactly once, zeros included).
constexpr int kPolyMaxCross = 8192;     // run boundaries per part held (and sorted) in shared memory

__device__ __forceinline__ int poly_up(double c) {      // (int)(scale*c + .5), no FMA contraction
  return (int)__dadd_rn(__dmul_rn(5.0, c), 0.5);
}

// boundary point `d` (0..max(dx,dy)) of edge (xs,ys)->(xe,ye) in walk order
__device__ __forceinline__ void poly_edge_point(int xs, int ys, int xe, int ye, int d, int& u,
                                                int& v) {
  const int dx = abs(xe - xs), dy = abs(ys - ye);
  const bool flip = (dx >= dy && xs > xe) || (dx < dy && ys > ye);
  if (flip) {
    int t = xs; xs = xe; xe = t;
    t = ys; ys = ye; ye = t;
  }
  if (dx >= dy) {
    const double s = dx == 0 ? 0.0 : __ddiv_rn((double)(ye - ys), (double)dx);
    const int t = flip ? dx - d : d;
    u = t + xs;
    v = (int)__dadd_rn(__dadd_rn((double)ys, __dmul_rn(s, (double)t)), 0.5);
  } else {
    const double s = __ddiv_rn((double)(xe - xs), (double)dy);
    const int t = flip ? dy - d : d;
    v = t + ys;
    u = (int)__dadd_rn(__dadd_rn((double)xs, __dmul_rn(s, (double)t)), 0.5);
  }
}

__global__ void __launch_bounds__(256)
poly_cross_kernel(const double* __restrict__ xy, const long long* __restrict__ part_off, int h, int w,
                  long long* __restrict__ edge_off,      // [total vertices + parts] scratch
                  unsigned* __restrict__ cross,          // [parts][kPolyMaxCross] sorted boundaries
                  int* __restrict__ cross_n,             // [parts]
                  int* __restrict__ err) {
  __shared__ unsigned s_a[kPolyMaxCross];
  __shared__ int s_n;
  __shared__ long long s_m;
  const int part = blockIdx.x;
  const long long v0 = part_off[part];
  const int k = (int)(part_off[part + 1] - v0);
  const double* p = xy + 2 * v0;
  long long* eoff = edge_off + v0 + part;                 // k + 1 entries
  if (threadIdx.x == 0) {
    s_n = 0;
    long long m = 0;                                       // dense boundary points before edge j
    for (int j = 0; j < k; ++j) {
      const int jn = j + 1 == k ? 0 : j + 1;
      const int dx = abs(poly_up(p[2 * j]) - poly_up(p[2 * jn]));
      const int dy = abs(poly_up(p[2 * j + 1]) - poly_up(p[2 * jn + 1]));
      eoff[j] = m;
      m += (dx > dy ? dx : dy) + 1;
    }
    eoff[k] = m;
    s_m = m;
  }
  __syncthreads();
  const long long m = s_m;
  for (long long j = 1 + threadIdx.x; j < m; j += blockDim.x) {
    // edge of point j: last e with eoff[e] <= j
    int lo = 0, hi = k - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (eoff[mid] <= j) lo = mid; else hi = mid - 1;
    }
    const int e = lo, en = e + 1 == k ? 0 : e + 1;
    const int d = (int)(j - eoff[e]);
    int u, v, up, vp;
    poly_edge_point(poly_up(p[2 * e]), poly_up(p[2 * e + 1]), poly_up(p[2 * en]),
                    poly_up(p[2 * en + 1]), d, u, v);
    if (d > 0) {
      poly_edge_point(poly_up(p[2 * e]), poly_up(p[2 * e + 1]), poly_up(p[2 * en]),
                      poly_up(p[2 * en + 1]), d - 1, up, vp);
    } else {                                               // last point of the previous edge
      const int ep = e - 1;
      poly_edge_point(poly_up(p[2 * ep]), poly_up(p[2 * ep + 1]), poly_up(p[2 * e]),
                      poly_up(p[2 * e + 1]), (int)(eoff[e] - eoff[ep]) - 1, up, vp);
    }
    if (u == up) continue;
    double xd = (double)(u < up ? u : u - 1);
    xd = __dadd_rn(__ddiv_rn(__dadd_rn(xd, 0.5), 5.0), -0.5);
    if (floor(xd) != xd || xd < 0 || xd > (double)(w - 1)) continue;
    double yd = (double)(v < vp ? v : vp);
    yd = __dadd_rn(__ddiv_rn(__dadd_rn(yd, 0.5), 5.0), -0.5);
    if (yd < 0) yd = 0; else if (yd > (double)h) yd = (double)h;
    yd = ceil(yd);
    const int slot = atomicAdd(&s_n, 1);
    if (slot < kPolyMaxCross) s_a[slot] = (unsigned)((int)xd * h + (int)yd);
  }
  __syncthreads();
  int n = s_n;
  if (n > kPolyMaxCross) {
    if (threadIdx.x == 0) atomicExch(err, 1 + part);
    n = kPolyMaxCross;
  }
  // bitonic sort of the first pow2 >= n entries (padding = UINT_MAX)
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  for (int i = n + threadIdx.x; i < np2; i += blockDim.x) s_a[i] = 0xffffffffu;
  __syncthreads();
  for (int size = 2; size <= np2; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < np2; i += blockDim.x) {
        const int jx = i ^ stride;
        if (jx > i) {
          const unsigned a = s_a[i], b = s_a[jx];
          const bool asc = (i & size) == 0;
          if ((a > b) == asc) {
            s_a[i] = b;
            s_a[jx] = a;
          }
        }
      }
      __syncthreads();
    }
  unsigned* dst = cross + (long long)part * kPolyMaxCross;
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = s_a[i];
  if (threadIdx.x == 0) cross_n[part] = n;
}

__global__ void __launch_bounds__(128)
poly_fill_kernel(const unsigned* __restrict__ cross, const int* __restrict__ cross_n,
                 const int* __restrict__ inst_part_off, int h, int w, uint8_t* __restrict__ out) {
  const int inst = blockIdx.y;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= w) return;
  uint8_t* col = out + (long long)inst * h * w + x;
  const int p0 = inst_part_off[inst], p1 = inst_part_off[inst + 1];
  if (p0 == p1) {
    for (int y = 0; y < h; ++y) col[(long long)y * w] = 0;
    return;
  }
  for (int part = p0; part < p1; ++part) {
    const unsigned* a = cross + (long long)part * kPolyMaxCross;
    const int n = cross_n[part];
    const unsigned base = (unsigned)x * (unsigned)h;
    int lo = 0, hi = n;                                   // first boundary >= base
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (a[mid] < base) lo = mid + 1; else hi = mid;
    }
    int ptr = lo;
    unsigned parity = (unsigned)lo & 1u;                  // boundaries in earlier columns carry over
    unsigned next = ptr < n ? a[ptr] : 0xffffffffu;
    for (int y = 0; y < h; ++y) {
      const unsigned i = base + (unsigned)y;
      while (next <= i) {
        parity ^= 1u;
        ++ptr;
        next = ptr < n ? a[ptr] : 0xffffffffu;
      }
      if (part == p0) col[(long long)y * w] = (uint8_t)parity;
      else if (parity) col[(long long)y * w] = 1;         // rleMerge(union): OR of the parts
    }
  }
}


// ---------------------------------------------------------------------------------------------
// Resize(keep_ratio) for tiles that are not already at img_scale (transforms.py:186-215,244-253):
// mmcv.imrescale -> cv2.resize.  Image: INTER_LINEAR on uint8, restated from OpenCV's
// imgproc/resize.cpp (resizeGeneric_ / HResizeLinear / VResizeLinear<uchar>): source coordinate
// (float)((d + .5) * scale - .5), 11-bit fixed-point weights rounded half-to-even, horizontal pass
// in int, vertical pass ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2.
// Bitmaps: INTER_NEAREST, source index min(floor(d * scale), size - 1).  cv2 is not in this image:
// the oracle (oracle/pipeline_cpu.py) is the same restatement -- PARITY UNPINNED for this step.
// The BONAI configuration never takes this path (1024^2 tiles at img_scale 1024: identity).
struct LinTap {
  int s0, s1;      // source indices (clamped)
  int a0, a1;      // fixed-point weights, sum 2048
};

__device__ __forceinline__ LinTap lin_tap_x(int d, double scale, int ssize) {
  float f = (float)__dadd_rn(__dmul_rn((double)d + 0.5, scale), -0.5);
  int s = (int)floorf(f);
  f -= (float)s;
  if (s < 0) {
    f = 0.f;
    s = 0;
  }
  if (s >= ssize - 1) {
    f = 0.f;
    s = ssize - 1;
  }
  LinTap t;
  t.s0 = s;
  t.s1 = s + 1 < ssize ? s + 1 : ssize - 1;
  t.a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
  t.a1 = __float2int_rn(__fmul_rn(f, 2048.f));
  return t;
}

__device__ __forceinline__ LinTap lin_tap_y(int d, double scale, int ssize) {
  float f = (float)__dadd_rn(__dmul_rn((double)d + 0.5, scale), -0.5);
  const int s = (int)floorf(f);
  f -= (float)s;                       // the weights are NOT reset at the border, the rows are clipped
  LinTap t;
  t.s0 = min(max(s, 0), ssize - 1);
  t.s1 = min(max(s + 1, 0), ssize - 1);
  t.a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
  t.a1 = __float2int_rn(__fmul_rn(f, 2048.f));
  return t;
}

__global__ void resize_bilinear_u8c3_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                            int sh, int sw, int dh, int dw, double scale_x,
                                            double scale_y) {
  const long long total = (long long)dh * dw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int dx = (int)(i % dw), dy = (int)(i / dw);
    const LinTap tx = lin_tap_x(dx, scale_x, sw), ty = lin_tap_y(dy, scale_y, sh);
    const uint8_t* r0 = src + (long long)ty.s0 * sw * 3;
    const uint8_t* r1 = src + (long long)ty.s1 * sw * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int h0 = (int)r0[tx.s0 * 3 + c] * tx.a0 + (int)r0[tx.s1 * 3 + c] * tx.a1;
      const int h1 = (int)r1[tx.s0 * 3 + c] * tx.a0 + (int)r1[tx.s1 * 3 + c] * tx.a1;
      const int v = (((ty.a0 * (h0 >> 4)) >> 16) + ((ty.a1 * (h1 >> 4)) >> 16) + 2) >> 2;
      dst[i * 3 + c] = (uint8_t)min(max(v, 0), 255);
    }
  }
}

__global__ void resize_nearest_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                         long long G, int sh, int sw, int dh, int dw, double scale_x,
                                         double scale_y) {
  const long long total = G * dh * dw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int dx = (int)(i % dw);
    const long long t = i / dw;
    const int dy = (int)(t % dh);
    const long long g = t / dh;
    const int sx = min((int)floor(__dmul_rn((double)dx, scale_x)), sw - 1);
    const int sy = min((int)floor(__dmul_rn((double)dy, scale_y)), sh - 1);
    dst[i] = src[(g * sh + sy) * (long long)sw + sx];
  }
}

inline int grid_for(long long n, int per_block = 256, int max_blocks = 148 * 16) {
  long long b = (n + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return (int)b;
}

}  // namespace

extern "C" {

int loft_image_prep(const uint8_t* img_hwc, float* out_chw, int H, int W, int Hp, int Wp,
                    const float* mean3, const float* std3, int to_rgb, int flip,
                    cudaStream_t stream) {
  LOFT_CHECK_ARG(img_hwc && out_chw && mean3 && std3, "image_prep: null pointer");
  LOFT_CHECK_SHAPE(H > 0 && W > 0 && Hp >= H && Wp >= W && Wp % 4 == 0,
                   "image_prep: bad sizes H=%d W=%d Hp=%d Wp=%d (Wp must be a multiple of 4)", H, W,
                   Hp, Wp);
  LOFT_CHECK_ARG(flip >= 0 && flip <= 2, "image_prep: flip must be 0 (none), 1 (horizontal), 2 (vertical)");
  NormParams np;
  for (int c = 0; c < 3; ++c) {
    np.mean[c] = mean3[c];                       // host pointers: three floats each
    np.stdinv[c] = 1.0 / (double)std3[c];
  }
  const long long total = (long long)Hp * (Wp / 4);
  image_prep_kernel<<<grid_for(total), 256, 0, stream>>>(img_hwc, out_chw, H, W, Hp, Wp, np, to_rgb,
                                                         flip);
  LOFT_CUDA_LAUNCH_CHECK("image_prep");
  return LOFT_OK;
}

int loft_mask_flip_pad(const uint8_t* in, uint8_t* out, long long G, int H, int W, int Hp, int Wp,
                       int flip, cudaStream_t stream) {
  LOFT_CHECK_ARG(in && out, "mask_flip_pad: null pointer");
  LOFT_CHECK_SHAPE(H > 0 && W > 0 && Hp >= H && Wp >= W && Wp % 16 == 0,
                   "mask_flip_pad: bad sizes H=%d W=%d Hp=%d Wp=%d (Wp must be a multiple of 16)",
                   H, W, Hp, Wp);
  LOFT_CHECK_ARG(flip >= 0 && flip <= 2, "mask_flip_pad: flip must be 0, 1 or 2");
  if (G == 0) return LOFT_OK;
  const long long total = G * Hp * (Wp / 16);
  const int aligned16 = (W % 16 == 0) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0);
  mask_flip_pad_kernel<<<grid_for(total), 256, 0, stream>>>(in, out, G, H, W, Hp, Wp, flip,
                                                            aligned16);
  LOFT_CUDA_LAUNCH_CHECK("mask_flip_pad");
  return LOFT_OK;
}


long long loft_poly_scratch_bytes(long long total_vertices, int n_parts) {
  // edge offsets (8 B each) + sorted boundaries + counts + error flag, each 16-byte aligned
  long long b = ((total_vertices + n_parts) * 8 + 15) / 16 * 16;
  b += (long long)n_parts * kPolyMaxCross * 4;
  b += ((long long)n_parts * 4 + 15) / 16 * 16;
  return b + 16;
}

int loft_poly_rasterize(const double* xy, const long long* part_off, const int* inst_part_off,
                        int n_inst, int n_parts, long long total_vertices, int H, int W,
                        uint8_t* out, void* scratch, int* err, cudaStream_t stream) {
  LOFT_CHECK_ARG(out && inst_part_off && err, "poly_rasterize: null pointer");
  LOFT_CHECK_SHAPE(n_inst >= 0 && n_parts >= 0 && H > 0 && W > 0 &&
                       (long long)H * W < (1ll << 31) - H,
                   "poly_rasterize: bad sizes n_inst=%d n_parts=%d H=%d W=%d", n_inst, n_parts, H, W);
  if (n_inst == 0) return LOFT_OK;
  LOFT_CHECK_ARG(n_parts == 0 || (xy && part_off && scratch), "poly_rasterize: null pointer");
  char* sp = static_cast<char*>(scratch);
  long long* edge_off = reinterpret_cast<long long*>(sp);
  sp += ((total_vertices + n_parts) * 8 + 15) / 16 * 16;
  unsigned* cross = reinterpret_cast<unsigned*>(sp);
  sp += (long long)n_parts * kPolyMaxCross * 4;
  int* cross_n = reinterpret_cast<int*>(sp);
  if (n_parts > 0) {
    poly_cross_kernel<<<n_parts, 256, 0, stream>>>(xy, part_off, H, W, edge_off, cross, cross_n, err);
    LOFT_CUDA_LAUNCH_CHECK("poly_cross");
  }
  poly_fill_kernel<<<dim3(loft_cdiv(W, 128), n_inst), 128, 0, stream>>>(cross, cross_n,
                                                                        inst_part_off, H, W, out);
  LOFT_CUDA_LAUNCH_CHECK("poly_fill");
  return LOFT_OK;
}


int loft_resize_bilinear_u8(const uint8_t* src_hwc, uint8_t* dst_hwc, int sh, int sw, int dh, int dw,
                            cudaStream_t stream) {
  LOFT_CHECK_ARG(src_hwc && dst_hwc, "resize_bilinear_u8: null pointer");
  LOFT_CHECK_SHAPE(sh > 0 && sw > 0 && dh > 0 && dw > 0, "resize_bilinear_u8: bad sizes %dx%d -> %dx%d",
                   sh, sw, dh, dw);
  // cv2.resize with dsize: inv_scale = dsize / ssize, scale = 1 / inv_scale
  const double scale_x = 1.0 / ((double)dw / (double)sw), scale_y = 1.0 / ((double)dh / (double)sh);
  resize_bilinear_u8c3_kernel<<<grid_for((long long)dh * dw), 256, 0, stream>>>(
      src_hwc, dst_hwc, sh, sw, dh, dw, scale_x, scale_y);
  LOFT_CUDA_LAUNCH_CHECK("resize_bilinear_u8");
  return LOFT_OK;
}

int loft_resize_nearest_u8(const uint8_t* src, uint8_t* dst, long long G, int sh, int sw, int dh,
                           int dw, cudaStream_t stream) {
  LOFT_CHECK_ARG(src && dst, "resize_nearest_u8: null pointer");
  LOFT_CHECK_SHAPE(G >= 0 && sh > 0 && sw > 0 && dh > 0 && dw > 0,
                   "resize_nearest_u8: bad sizes G=%lld %dx%d -> %dx%d", G, sh, sw, dh, dw);
  if (G == 0) return LOFT_OK;
  const double scale_x = 1.0 / ((double)dw / (double)sw), scale_y = 1.0 / ((double)dh / (double)sh);
  resize_nearest_u8_kernel<<<grid_for(G * dh * dw), 256, 0, stream>>>(src, dst, G, sh, sw, dh, dw,
                                                                      scale_x, scale_y);
  LOFT_CUDA_LAUNCH_CHECK("resize_nearest_u8");
  return LOFT_OK;
}

}  // extern "C"
