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

}  // extern "C"
