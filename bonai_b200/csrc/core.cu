// Error string, device query and ABI version of libloft_b200.so.
#include "common.cuh"
#include "loft_b200.h"
#include <stdarg.h>
#include <string.h>

static thread_local char g_err[512] = "";

void loft_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int g_sm_reserve = 0;

// SMs the persistent grids may fill: all of them, minus what loft_reserve_sms() set aside for a
// concurrently running collective (its CTAs cannot share an SM with a 197 KB-smem GEMM CTA; a
// statically striped persistent grid that does not fit runs a ragged second wave instead).
int loft_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  const int avail = n - g_sm_reserve;
  return avail > 2 ? (avail & ~1) : 2;
}

__global__ void mask_windows_kernel(const unsigned char* __restrict__ src,
                                    const float* __restrict__ boxes, unsigned char* __restrict__ dst,
                                    int H, int W, int pad) {
  const int g = blockIdx.x;
  const float4 b = reinterpret_cast<const float4*>(boxes)[g];
  const int x0 = max(0, min(W, (int)floorf(b.x) - pad)), y0 = max(0, min(H, (int)floorf(b.y) - pad));
  const int x1 = max(0, min(W, (int)ceilf(b.z) + pad)), y1 = max(0, min(H, (int)ceilf(b.w) + pad));
  const int w = x1 - x0, h = y1 - y0;
  if (w <= 0 || h <= 0) return;
  const size_t base = (size_t)g * H * W;
  for (int i = threadIdx.x; i < w * h; i += blockDim.x) {
    const int y = i / w, x = i - y * w;
    const size_t o = base + (size_t)(y0 + y) * W + x0 + x;
    dst[o] = src[o];
  }
}

extern "C" {
const char* loft_last_error(void) { return g_err; }
int loft_abi_version(void) { return 3; }
// Host -> device transfer of the gt-box WINDOWS of a stack of bitmaps [G, H, W] (uint8, pinned
// host memory, read by the kernel itself over PCIe -- pinned allocations are device-addressable
// under unified addressing) into the same positions of a dense device stack the caller has zeroed.
// A building's bitmap is zero outside its box, so this moves ~0.5 MB per tile instead of
// G * H * W = 80 MB (BitmapMasks of mmdet/core/mask/structures.py:20-60 as fed by
// pipelines/formating.py:191-230).  One block per GT; the window is the box (device copy of
// gt_bboxes) grown by `pad` pixels and clipped to the tile, so the host issues ONE launch per
// image (a strided cudaMemcpy2DAsync per box cost the launch thread ~1 ms per step).
int loft_h2d_mask_windows(const unsigned char* src_host, const float* boxes_dev,
                          unsigned char* dst_dev, int G, int H, int W, int pad,
                          cudaStream_t stream) {
  LOFT_CHECK_ARG((src_host && boxes_dev && dst_dev) || G == 0, "h2d_mask_windows: null pointer");
  LOFT_CHECK_SHAPE(G >= 0 && H > 0 && W > 0 && pad >= 0, "h2d_mask_windows: bad shape");
  if (G == 0) return LOFT_OK;
  mask_windows_kernel<<<G, 256, 0, stream>>>(src_host, boxes_dev, dst_dev, H, W, pad);
  LOFT_CUDA_LAUNCH_CHECK("h2d_mask_windows");
  return LOFT_OK;
}
int loft_reserve_sms(int n) {
  const int prev = g_sm_reserve;
  g_sm_reserve = n > 0 ? n : 0;
  return prev;
}
}
