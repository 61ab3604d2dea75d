// Fused loss kernels (forward sums with warp-shuffle reductions; backward with a device-side
// upstream-gradient scalar so nothing synchronises the host).
// Reference: models/losses/cross_entropy_loss.py:9-125 (CE / sigmoid BCE / mask BCE),
// models/losses/smooth_l1_loss.py:8-42 (SmoothL1 / L1), models/losses/utils.py:26-52
// (weight_reduce_loss), models/losses/accuracy.py:4-48, models/losses/focal_loss.py:10-86 and
// mmcv.ops.sigmoid_focal_loss [mmcv-full 1.0.5].
#include "common.cuh"
#include "loft_b200.h"
#include <float.h>

namespace {

constexpr int kT = 256;

enum { kBceLogits = 0, kL1 = 1, kSmoothL1 = 2 };

__device__ __forceinline__ float block_sum(float v) {
  __shared__ float sm[kT / 32];
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < kT / 32) t = sm[threadIdx.x];
  if (threadIdx.x < 32) t = warp_sum(t);
  __syncthreads();
  return t;
}

__device__ __forceinline__ float bce_logits(float x, float t) {
  // max(x,0) - x*t + log(1 + exp(-|x|))  (F.binary_cross_entropy_with_logits)
  return fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
}

// element e -> (row, col) = (e / ncols, e % ncols); pred at pred[row*ld + col_off + col];
// target / weight are dense [rows*ncols] (weight NULL => 1).
template <bool kBackward>
__global__ void elem_loss_kernel(int mode, const float* __restrict__ pred, long long ld, int col_off,
                                 int ncols, long long rows, const float* __restrict__ target,
                                 const float* __restrict__ weight, float beta, float scale,
                                 const float* __restrict__ gscale, float* __restrict__ out_sum,
                                 float* __restrict__ dpred) {
  const long long n = rows * ncols;
  float acc = 0.f;
  const float gs = kBackward ? scale * (gscale ? *gscale : 1.f) : 0.f;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
       e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / ncols;
    const int c = (int)(e - r * ncols);
    const long long pi = r * ld + col_off + c;
    const float x = pred[pi], t = target[e];
    const float w = weight ? weight[e] : 1.f;
    if (!kBackward) {
      float l;
      if (mode == kBceLogits) {
        l = bce_logits(x, t);
      } else {
        const float d = fabsf(x - t);
        l = (mode == kL1) ? d : (d < beta ? 0.5f * d * d / beta : d - 0.5f * beta);
      }
      acc += l * w;
    } else {
      float g;
      if (mode == kBceLogits) {
        g = 1.f / (1.f + expf(-x)) - t;
      } else {
        const float d = x - t;
        if (mode == kL1)
          g = (d > 0.f) ? 1.f : (d < 0.f ? -1.f : 0.f);
        else
          g = (fabsf(d) < beta) ? d / beta : (d > 0.f ? 1.f : -1.f);
      }
      dpred[pi] = g * w * gs;
    }
  }
  if (!kBackward) {
    const float t = block_sum(acc);
    if (threadIdx.x == 0) atomicAdd(out_sum, t * scale);
  }
}

// softmax CE over rows of [n, C] logits (row pitch ld) with per-row weights; out[0] += sum(ce*w)*
// scale, out[1] += #(argmax == label) (top-1 accuracy numerator).
template <bool kBackward>
__global__ void softmax_ce_kernel(const float* __restrict__ logits, long long ld, int C, long long n,
                                  const long long* __restrict__ labels,
                                  const float* __restrict__ weight, float scale,
                                  const float* __restrict__ gscale, float* __restrict__ out,
                                  float* __restrict__ dlogits) {
  float acc = 0.f, correct = 0.f;
  const float gs = kBackward ? scale * (gscale ? *gscale : 1.f) : 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float* x = logits + i * ld;
    float m = x[0];
    int am = 0;
    for (int c = 1; c < C; ++c)
      if (x[c] > m) {
        m = x[c];
        am = c;
      }
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += expf(x[c] - m);
    const int lab = (int)labels[i];
    const float w = weight ? weight[i] : 1.f;
    if (!kBackward) {
      acc += (logf(s) + m - x[lab]) * w;
      correct += (am == lab) ? 1.f : 0.f;
    } else {
      for (int c = 0; c < C; ++c)
        dlogits[i * ld + c] = (expf(x[c] - m) / s - (c == lab ? 1.f : 0.f)) * w * gs;
    }
  }
  if (!kBackward) {
    const float t = block_sum(acc);
    const float k = block_sum(correct);
    if (threadIdx.x == 0) {
      atomicAdd(out, t * scale);
      atomicAdd(out + 1, k);
    }
  }
}

// mmcv sigmoid_focal_loss: target[n] in [0, C], C = background.
template <bool kBackward>
__global__ void focal_kernel(const float* __restrict__ x, const long long* __restrict__ target,
                             const float* __restrict__ weight, long long n, int C, float gamma,
                             float alpha, float* __restrict__ loss, float* __restrict__ dx,
                             const float* __restrict__ dloss) {
  const long long total = n * C;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long i = e / C;
    const int c = (int)(e - i * C);
    const long long t = target[i];
    const float v = x[e];
    const float p = 1.f / (1.f + expf(-v));
    const float w = weight ? weight[i] : 1.f;
    const bool pos = (t == c);
    const bool neg = (t != c) && (t >= 0);
    if (!kBackward) {
      // -[t==c] a (1-p)^g log(max(p,FLT_MIN)) - [t!=c] (1-a) p^g log(1-p)
      const float t1 = powf(1.f - p, gamma) * logf(fmaxf(p, FLT_MIN));
      const float t2 = powf(p, gamma) * (-v * (v >= 0.f) - log1pf(expf(v - 2.f * v * (v >= 0.f))));
      loss[e] = (-(pos ? alpha * t1 : 0.f) - (neg ? (1.f - alpha) * t2 : 0.f)) * w;
    } else {
      const float t1 = powf(1.f - p, gamma) * (1.f - p - gamma * p * logf(fmaxf(p, FLT_MIN)));
      const float t2 = powf(p, gamma) * ((-v * (v >= 0.f) - log1pf(expf(v - 2.f * v * (v >= 0.f)))) *
                                             (1.f - p) * gamma -
                                         p);
      const float g = -(pos ? alpha * t1 : 0.f) - (neg ? (1.f - alpha) * t2 : 0.f);
      dx[e] = g * w * dloss[e];
    }
  }
}

inline int blocks_for(long long n) {
  long long b = (n + kT - 1) / kT;
  if (b < 1) b = 1;
  if (b > 148 * 8) b = 148 * 8;
  return (int)b;
}



// ---------------------------------------------------------------- fused RPN loss, all levels
// One launch for AnchorHead.loss over every pyramid level (anchor_head.py:382-497): per level the
// sigmoid-BCE classification sum and the L1 / SmoothL1 regression sum (weight_reduce_loss with
// avg_factor = num_total_samples), and -- the total loss being the plain sum of the terms
// (detectors/base.py:175-208) -- the gradient w.r.t. the fused head output [rows, ld] written in
// full (padding columns zero) into the buffer the RPN backward program reads.  Replaces 10 forward
// and 10 backward elem_loss launches, their zero-fills, and 5 gradient copies per step.
constexpr int kMaxRpnLevels = 8;
struct RpnLossArgs {
  loft_rpn_level_t lv[kMaxRpnLevels];
  long long row_end[kMaxRpnLevels];   // prefix sums of rows
  int n_levels, A, ld, mode_bbox;
  float beta, cls_scale, bbox_scale;
  const float* inv_denom;   // optional device scalar: both scales are divided by it (avg_factor)
};

__global__ void rpn_loss_fused_kernel(const RpnLossArgs a, float* __restrict__ sums) {
  __shared__ float s_acc[2 * kMaxRpnLevels];
  if (threadIdx.x < 2 * kMaxRpnLevels) s_acc[threadIdx.x] = 0.f;
  __syncthreads();
  const long long total = a.row_end[a.n_levels - 1] * a.ld;
  const int A = a.A, ld = a.ld;
  const float inv = a.inv_denom ? 1.f / fmaxf(*a.inv_denom, 1.f) : 1.f;
  const float cls_scale = a.cls_scale * inv, bbox_scale = a.bbox_scale * inv;
  int cur = -1;
  float acc_c = 0.f, acc_b = 0.f;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long grow = e / ld;
    const int col = (int)(e - grow * ld);
    int l = 0;
    while (grow >= a.row_end[l]) ++l;
    if (l != cur) {
      if (cur >= 0) {
        atomicAdd(&s_acc[2 * cur], acc_c);
        atomicAdd(&s_acc[2 * cur + 1], acc_b);
      }
      cur = l;
      acc_c = acc_b = 0.f;
    }
    const loft_rpn_level_t& L = a.lv[l];
    const long long row = grow - (l ? a.row_end[l - 1] : 0);
    const float x = L.out[row * ld + col];
    float g = 0.f;
    if (col < A) {
      const long long ti = row * A + col;
      const float t = L.labels[ti], w = L.label_w[ti];
      acc_c += bce_logits(x, t) * w;
      g = (1.f / (1.f + expf(-x)) - t) * w * cls_scale;
    } else if (col < 5 * A) {
      const long long ti = row * 4 * A + (col - A);
      const float t = L.bbox_t[ti], w = L.bbox_w[ti];
      const float d = x - t, ad = fabsf(d);
      float gd;
      if (a.mode_bbox == kL1) {
        acc_b += ad * w;
        gd = (d > 0.f) ? 1.f : (d < 0.f ? -1.f : 0.f);
      } else {
        acc_b += (ad < a.beta ? 0.5f * d * d / a.beta : ad - 0.5f * a.beta) * w;
        gd = (ad < a.beta) ? d / a.beta : (d > 0.f ? 1.f : -1.f);
      }
      g = gd * w * bbox_scale;
    }
    if (L.grad != nullptr) L.grad[row * ld + col] = g;
  }
  if (cur >= 0) {
    atomicAdd(&s_acc[2 * cur], acc_c);
    atomicAdd(&s_acc[2 * cur + 1], acc_b);
  }
  __syncthreads();
  if (threadIdx.x < 2 * a.n_levels) {
    const int l = threadIdx.x >> 1, which = threadIdx.x & 1;
    const float v = s_acc[threadIdx.x] * (which ? bbox_scale : cls_scale);
    if (v != 0.f) atomicAdd(&sums[which * a.n_levels + l], v);
  }
}

}  // namespace

extern "C" {

int loft_rpn_loss_fused(const loft_rpn_level_t* levels, int n_levels, int A, int ld, int mode_bbox,
                        float beta, float cls_scale, float bbox_scale, const float* denom,
                        float* sums, cudaStream_t stream) {
  LOFT_CHECK_ARG(levels && sums, "rpn_loss_fused: null pointer");
  LOFT_CHECK_SHAPE(n_levels >= 1 && n_levels <= kMaxRpnLevels && 5 * A <= ld,
                   "rpn_loss_fused: n_levels=%d A=%d ld=%d", n_levels, A, ld);
  LOFT_CHECK_ARG(mode_bbox == kL1 || mode_bbox == kSmoothL1, "rpn_loss_fused: bad bbox mode %d",
                 mode_bbox);
  RpnLossArgs a{};
  long long rows = 0;
  for (int l = 0; l < n_levels; ++l) {
    LOFT_CHECK_ARG(levels[l].out && levels[l].labels && levels[l].label_w && levels[l].bbox_t &&
                       levels[l].bbox_w,
                   "rpn_loss_fused: null pointer in level %d", l);
    a.lv[l] = levels[l];
    rows += levels[l].rows;
    a.row_end[l] = rows;
  }
  a.n_levels = n_levels;
  a.A = A;
  a.ld = ld;
  a.mode_bbox = mode_bbox;
  a.beta = beta;
  a.cls_scale = cls_scale;
  a.bbox_scale = bbox_scale;
  a.inv_denom = denom;
  cudaMemsetAsync(sums, 0, sizeof(float) * 2 * n_levels, stream);
  if (rows == 0) return LOFT_OK;
  rpn_loss_fused_kernel<<<blocks_for(rows * ld), kT, 0, stream>>>(a, sums);
  LOFT_CUDA_LAUNCH_CHECK("rpn_loss_fused");
  return LOFT_OK;
}

int loft_elem_loss_fwd(int mode, const float* pred, long long ld, int col_off, int ncols,
                       long long rows, const float* target, const float* weight, float beta,
                       float scale, float* out_sum, cudaStream_t stream) {
  LOFT_CHECK_ARG(pred && target && out_sum, "elem_loss_fwd: null pointer");
  LOFT_CHECK_ARG(mode >= 0 && mode <= 2, "elem_loss_fwd: bad mode %d", mode);
  if (rows * ncols == 0) return LOFT_OK;
  elem_loss_kernel<false><<<blocks_for(rows * ncols), kT, 0, stream>>>(
      mode, pred, ld, col_off, ncols, rows, target, weight, beta, scale, nullptr, out_sum, nullptr);
  LOFT_CUDA_LAUNCH_CHECK("elem_loss_fwd");
  return LOFT_OK;
}

int loft_elem_loss_bwd(int mode, const float* pred, long long ld, int col_off, int ncols,
                       long long rows, const float* target, const float* weight, float beta,
                       float scale, const float* gscale, float* dpred, cudaStream_t stream) {
  LOFT_CHECK_ARG(pred && target && dpred, "elem_loss_bwd: null pointer");
  LOFT_CHECK_ARG(mode >= 0 && mode <= 2, "elem_loss_bwd: bad mode %d", mode);
  if (rows * ncols == 0) return LOFT_OK;
  elem_loss_kernel<true><<<blocks_for(rows * ncols), kT, 0, stream>>>(
      mode, pred, ld, col_off, ncols, rows, target, weight, beta, scale, gscale, nullptr, dpred);
  LOFT_CUDA_LAUNCH_CHECK("elem_loss_bwd");
  return LOFT_OK;
}

int loft_softmax_ce_fwd(const float* logits, long long ld, int C, long long n,
                        const long long* labels, const float* weight, float scale, float* out2,
                        cudaStream_t stream) {
  LOFT_CHECK_ARG(logits && labels && out2, "softmax_ce_fwd: null pointer");
  if (n == 0) return LOFT_OK;
  softmax_ce_kernel<false><<<blocks_for(n), kT, 0, stream>>>(logits, ld, C, n, labels, weight, scale,
                                                            nullptr, out2, nullptr);
  LOFT_CUDA_LAUNCH_CHECK("softmax_ce_fwd");
  return LOFT_OK;
}

int loft_softmax_ce_bwd(const float* logits, long long ld, int C, long long n,
                        const long long* labels, const float* weight, float scale,
                        const float* gscale, float* dlogits, cudaStream_t stream) {
  LOFT_CHECK_ARG(logits && labels && dlogits, "softmax_ce_bwd: null pointer");
  if (n == 0) return LOFT_OK;
  softmax_ce_kernel<true><<<blocks_for(n), kT, 0, stream>>>(logits, ld, C, n, labels, weight, scale,
                                                           gscale, nullptr, dlogits);
  LOFT_CUDA_LAUNCH_CHECK("softmax_ce_bwd");
  return LOFT_OK;
}

int loft_sigmoid_focal_loss_fwd(const float* x, const long long* target, const float* weight,
                                long long n, int C, float gamma, float alpha, float* loss,
                                cudaStream_t stream) {
  LOFT_CHECK_ARG(x && target && loss, "sigmoid_focal_loss_fwd: null pointer");
  if (n * C == 0) return LOFT_OK;
  focal_kernel<false><<<blocks_for(n * C), kT, 0, stream>>>(x, target, weight, n, C, gamma, alpha,
                                                           loss, nullptr, nullptr);
  LOFT_CUDA_LAUNCH_CHECK("sigmoid_focal_loss_fwd");
  return LOFT_OK;
}

int loft_sigmoid_focal_loss_bwd(const float* x, const long long* target, const float* weight,
                                long long n, int C, float gamma, float alpha, const float* dloss,
                                float* dx, cudaStream_t stream) {
  LOFT_CHECK_ARG(x && target && dloss && dx, "sigmoid_focal_loss_bwd: null pointer");
  if (n * C == 0) return LOFT_OK;
  focal_kernel<true><<<blocks_for(n * C), kT, 0, stream>>>(x, target, weight, n, C, gamma, alpha,
                                                          nullptr, dx, dloss);
  LOFT_CUDA_LAUNCH_CHECK("sigmoid_focal_loss_bwd");
  return LOFT_OK;
}

}  // extern "C"
