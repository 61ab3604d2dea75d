/* loft_b200.h -- C ABI of the B200-native LOFT/FOA training hot path.
 *
 * Every entry point is `extern "C"`, takes plain device pointers + sizes + a cudaStream_t, never
 * allocates, never synchronises the device, and returns 0 or a negative error code
 * (LOFT_ERR_ARG -1 bad argument, LOFT_ERR_SHAPE -2 unsupported shape, LOFT_ERR_CUDA -3 CUDA error;
 * loft_last_error() holds the message).  All tensors are fp32, activations are NHWC
 * (= torch channels_last), conv weights are [Cout][kh][kw][Cin] (= torch OIHW in channels_last).
 *
 * Each group cites the reference call site it replaces (paths under jwwangchn/BONAI @ aeafa46;
 * "mmcv" = the un-vendored mmcv-full==1.0.5 dependency pinned at mmdet/__init__.py:18-26).
 */
#ifndef LOFT_B200_H_
#define LOFT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

const char* loft_last_error(void);
int loft_abi_version(void);

/* Fused epilogue of the dense kernels.  For pixel p, channel c:
 *   acc = sum_k ...;  if (raw_out) raw_out[p,c] = acc;           (pre-BN conv output)
 *   v = acc*scale[c] + shift[c]                                   (BN-eval affine / bias)
 *   v += residual[p,c]  (or residual[up2x(p),c] if res_upsample2x: FPN top-down, fpn.py:185-186)
 *   if (relu) v = max(v,0);  if (mask) v = mask[p,c] > 0 ? v : 0  (ReLU backward)
 *   out[p,c] = v   (deconv_shuffle: c=(i*2+j)*Co+co scatters to pixel (2h+i,2w+j), channel co)
 */
typedef struct loft_epilogue_t {
  float* raw_out;
  const float* scale;
  const float* shift;
  const float* residual;
  const float* mask;
  long long ldr;
  int res_upsample2x;
  int relu;
  int deconv_shuffle;
} loft_epilogue_t;

/* ---- dense contractions on tcgen05 (TF32 operands, fp32 accumulate in TMEM) -------------------
 * replace cuDNN/cuBLAS behind torch.nn.Conv2d / nn.Linear / ConvTranspose2d at
 * resnet.py:163-203,260-300; fpn.py:116-132,164-216; rpn_head.py:26-44;
 * convfc_bbox_head.py:118-173; fcn_mask_head.py:64-126; offset_head_expand_feature.py:72-161. */
int loft_gemm_fprop(const float* x, const float* w, float* y, long long P, int K, int Cout,
                    long long ldx, long long ldw, long long ldy, int H, int W,
                    const loft_epilogue_t* epi, cudaStream_t stream);
int loft_gemm_dgrad(const float* dy, const float* w, float* dx, long long P, int Cin, int Cout,
                    long long lddy, long long ldw, long long lddx, const loft_epilogue_t* epi,
                    cudaStream_t stream);
int loft_gemm_wgrad(const float* dy, const float* x, float* dw, long long P, int Cin, int Cout,
                    long long lddy, long long ldx, long long lddw, cudaStream_t stream);
int loft_conv3x3_fprop(const float* x, const float* w, float* y, int N, int H, int W, int Cin,
                       int Cout, const loft_epilogue_t* epi, cudaStream_t stream);
int loft_conv3x3_dgrad(const float* dy, const float* w, float* dx, int N, int H, int W, int Cin,
                       int Cout, const loft_epilogue_t* epi, cudaStream_t stream);
int loft_conv3x3_wgrad(const float* dy, const float* x, float* dw, int N, int H, int W, int Cin,
                       int Cout, cudaStream_t stream);
/* bring-up only: override UMMA descriptor fields (-1 = keep default) */
void loft_debug_set_desc(long long a_desc, long long b_desc, long long a_kstep, long long b_kstep,
                         long long idesc);

#ifdef __cplusplus
}
#endif
#endif /* LOFT_B200_H_ */
