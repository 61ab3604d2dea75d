// Probe: does cuTensorMapEncodeTiled accept OVERLAPPING strides (stride of dim 1 smaller than the
// extent of dim 0), and does the TMA then deliver the sliding-window rows?  Used to decide the
// operand form of the direct 7x7/2 stem conv (loft_stem_conv): a padded NHWC4 image, where the
// 8 taps x 4 channels of one kernel row are 128 contiguous bytes starting every 32 bytes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o build/tma_overlap_probe tools/probes/tma_overlap_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

__global__ void probe_kernel(const __grid_constant__ CUtensorMap tm, float* out, int rank, int c1,
                             int c2, int c3, int c4, int bytes) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
  const uint32_t sm_a = (uint32_t)__cvta_generic_to_shared(smem);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes));
    if (rank == 4)
      asm volatile(
          "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
          "{%3, %4, %5, %6}], [%2];" ::"r"(sm_a),
          "l"(&tm), "r"(bar_a), "r"(0), "r"(c1), "r"(c2), "r"(c3)
          : "memory");
    else
      asm volatile(
          "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
          "{%3, %4, %5, %6, %7}], [%2];" ::"r"(sm_a),
          "l"(&tm), "r"(bar_a), "r"(0), "r"(0), "r"(c2), "r"(c3), "r"(c4)
          : "memory");
  }
  uint32_t done = 0;
  while (!done)
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
        : "=r"(done)
        : "r"(bar_a), "r"(0)
        : "memory");
  for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) out[i] = reinterpret_cast<float*>(smem)[i];
}

int main() {
  cuInit(0);
  const int N = 2, Hp = 38, Wp = 72;   // padded image, NHWC4, rows split into even / odd planes
  const int Hh = Hp / 2;
  const int Wo = 32;
  std::vector<float> h((size_t)N * 2 * Hh * Wp * 4);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
  float *d, *o;
  cudaMalloc(&d, h.size() * 4);
  cudaMalloc(&o, 64 * 1024);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  const int tw = 16, th = 4;
  int fails = 0;
  for (int form = 0; form < 2; ++form) {
    for (int swz = 0; swz < 2; ++swz) {
      CUtensorMap tm;
      cuuint64_t gd[5], gs[4];
      cuuint32_t bx[5], es[5] = {1, 1, 1, 1, 1};
      int rank;
      if (form == 0) {   // A: (32 floats, ox @32 B, half-row, plane)
        rank = 4;
        gd[0] = 32; gd[1] = Wo; gd[2] = Hh; gd[3] = (cuuint64_t)N * 2;
        gs[0] = 32; gs[1] = (cuuint64_t)Wp * 16; gs[2] = (cuuint64_t)Hh * Wp * 16;
        bx[0] = 32; bx[1] = tw; bx[2] = th; bx[3] = 1;
      } else {           // B: (8 floats, 4 sub-runs @32 B, ox @32 B, half-row, plane)
        rank = 5;
        gd[0] = 8; gd[1] = 4; gd[2] = Wo; gd[3] = Hh; gd[4] = (cuuint64_t)N * 2;
        gs[0] = 32; gs[1] = 32; gs[2] = (cuuint64_t)Wp * 16; gs[3] = (cuuint64_t)Hh * Wp * 16;
        bx[0] = 8; bx[1] = 4; bx[2] = tw; bx[3] = th; bx[4] = 1;
      }
      CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, d, gd, gs, bx, es,
                                          CU_TENSOR_MAP_INTERLEAVE_NONE,
                                          swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      printf("form %c swizzle %d: encode CUresult %d\n", form ? 'B' : 'A', swz, (int)r);
      if (r != CUDA_SUCCESS) { ++fails; continue; }
      const int w0 = 16, h0 = 3, pl = 3, bytes = tw * th * 128;
      cudaMemset(o, 0xff, 64 * 1024);
      cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
      probe_kernel<<<1, 128, 32 * 1024>>>(tm, o, rank, w0, form ? w0 : h0, form ? h0 : pl, pl, bytes);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("  kernel failed: %s\n", cudaGetErrorString(e)); return 2; }
      std::vector<float> got(bytes / 4);
      cudaMemcpy(got.data(), o, bytes, cudaMemcpyDeviceToHost);
      long bad = 0;
      for (int y = 0; y < th; ++y)
        for (int x = 0; x < tw; ++x)
          for (int e4 = 0; e4 < 32; ++e4) {
            const size_t src = (((size_t)pl * Hh + (h0 + y)) * Wp) * 4 + (size_t)(w0 + x) * 8 + e4;
            int row = y * tw + x, chunk = e4 / 4, within = e4 % 4;
            int pc = swz ? (chunk ^ (row & 7)) : chunk;
            float v = got[(size_t)row * 32 + pc * 4 + within];
            if (v != h[src]) { if (bad < 4) printf("  mismatch y%d x%d e%d got %.0f want %.0f\n", y, x, e4, v, h[src]); ++bad; }
          }
      printf("  data: %ld mismatches of %d\n", bad, bytes / 4);
      if (bad) ++fails;
    }
  }
  printf("PROBE %s\n", fails ? "SOME FORMS FAILED" : "ALL FORMS OK");
  return 0;
}
