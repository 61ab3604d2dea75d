// HBM-bound helper kernels of the LOFT path: TF32 rounding / weight repacks, BN folding and
// activation backward (+ per-channel reductions), stem im2col, max-pool, 2x subsampling, FPN
// top-down backward, FOA rot90, fused grad-norm + clip + SGD.
// Reference semantics: backbones/resnet.py:260-300,525-571,623-649 (BN eval, ReLU, maxpool),
// necks/fpn.py:164-216, attribute_heads/offset_head_expand_feature.py:163-196 (rotation),
// mmcv OptimizerHook + torch.optim.SGD (configs/_base_/schedules/schedule_2x_bonai.py:2-3).
#include "common.cuh"
#include "loft_b200.h"

namespace {

constexpr int kT = 256;

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

inline int grid_for(long long n, int per_block = kT, int max_blocks = 148 * 16) {
  long long b = (n + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return (int)b;
}

// ---------------------------------------------------------------- copy / round / repack
// dst[r*ldd + c] (=|+=) round?(src[r*lds + c])
__global__ void copy2d_kernel(const float* __restrict__ src, long long lds, float* __restrict__ dst,
                              long long ldd, long long rows, int cols, int accumulate, int round) {
  const long long n = rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = (int)(i - r * cols);
    float v = src[r * lds + c];
    if (round) v = tf32_rna(v);
    float* d = dst + r * ldd + c;
    *d = accumulate ? (*d + v) : v;
  }
}

// Up to kMaxCopies independent copy2d jobs in ONE launch (blockIdx.y = job): the per-step refresh of
// the fused / padded head weights and the scatter of their gradients were ~40 launches of a few
// hundred elements each, back to back on the critical path between the optimizer and the forward.
constexpr int kMaxCopies = 48;
struct CopyJobs {
  loft_copy2d_t job[kMaxCopies];
};

__global__ void copy2d_multi_kernel(const CopyJobs jobs) {
  const loft_copy2d_t& j = jobs.job[blockIdx.y];
  const long long n = j.rows * j.cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / j.cols;
    const int c = (int)(i - r * j.cols);
    float v = j.src[r * j.lds + c];
    if (j.round_tf32) v = tf32_rna(v);
    float* d = j.dst + r * j.ldd + c;
    *d = j.accumulate ? (*d + v) : v;
  }
}

// Generic 3-axis permute of a [A][B][C] tensor into [A][C][B] (dst) or back, with optional TF32
// rounding / accumulation.  Covers fc (C,HW)->(HW,C) weight repacks and their gradient un-packs.
__global__ void permute_acb_kernel(const float* __restrict__ src, float* __restrict__ dst, int A,
                                   int B, int C, int accumulate, int round) {
  __shared__ float tile[32][33];
  const int a = blockIdx.z;
  const int b0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const float* s = src + (long long)a * B * C;
  float* d = dst + (long long)a * B * C;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int b = b0 + i, c = c0 + threadIdx.x;
    if (b < B && c < C) tile[i][threadIdx.x] = s[(long long)b * C + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, b = b0 + threadIdx.x;
    if (b < B && c < C) {
      float v = tile[threadIdx.x][i];
      if (round) v = tf32_rna(v);
      float* p = d + (long long)c * B + b;
      *p = accumulate ? (*p + v) : v;
    }
  }
}

__global__ void fill_kernel(float4* __restrict__ p4, float* __restrict__ p, long long n4,
                            long long n, float v) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long k = i; k < n4; k += stride) p4[k] = make_float4(v, v, v, v);
  for (long long k = 4 * n4 + i; k < n; k += stride) p[k] = v;
}

// ---------------------------------------------------------------- BN fold / activation backward
__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ mean, const float* __restrict__ var,
                               float eps, float* __restrict__ scale, float* __restrict__ shift,
                               float* __restrict__ rstd, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float r = 1.0f / sqrtf(var[c] + eps);
  const float s = gamma[c] * r;
  scale[c] = s;
  shift[c] = beta[c] - mean[c] * s;
  if (rstd) rstd[c] = r;
}

// ---------------------------------------------------------------- BN folded into conv weights
// Eval-mode BN after a conv is y = s_c * (W_c . x) + shift_c with s_c = gamma_c * rstd_c.  The
// tensor cores read W'_c = tf32(s_c * W_c), so the forward epilogue is shift-only, the data
// gradient needs no per-channel scaling pass, and the weight-gradient kernels produce dW'.
// One block per weight row (= output channel); rows are described by a table built once.
__global__ void __launch_bounds__(128)
bn_fold_weights_kernel(const float* __restrict__ P, float* __restrict__ T,
                       const long long* __restrict__ row_off, const int* __restrict__ row_k,
                       const int* __restrict__ row_ch, const float* __restrict__ scale) {
  const long long r = blockIdx.x;
  const long long off = row_off[r];
  const int K = row_k[r];
  const float s = scale[row_ch[r]];
  if (((off | K) & 3) == 0) {
    const float4* src = reinterpret_cast<const float4*>(P + off);
    float4* dst = reinterpret_cast<float4*>(T + off);
    for (int i = threadIdx.x; i < (K >> 2); i += blockDim.x) {
      float4 v = src[i];
      dst[i] = make_float4(tf32_rna(v.x * s), tf32_rna(v.y * s), tf32_rna(v.z * s),
                           tf32_rna(v.w * s));
    }
  } else {
    for (int i = threadIdx.x; i < K; i += blockDim.x) T[off + i] = tf32_rna(P[off + i] * s);
  }
}

// After backward: G rows hold dW'.  dgamma_c += rstd_c * (W_c . dW'_c - mean_c * dbeta_c)
// (dL/dgamma through W' = gamma*rstd*W and shift = beta - mean*gamma*rstd), then
// dW_c = s_c * dW'_c in place.  Replaces the sum over pixels of dy * x_hat that
// torch.nn.BatchNorm2d's backward computes from the saved conv output (resnet.py:260-300).
__global__ void __launch_bounds__(128)
bn_finalize_kernel(const float* __restrict__ P, float* __restrict__ G,
                   const long long* __restrict__ row_off, const int* __restrict__ row_k,
                   const int* __restrict__ row_ch, const float* __restrict__ scale,
                   const float* __restrict__ rstd, const float* __restrict__ mean,
                   const float* __restrict__ dbeta, float* __restrict__ dgamma) {
  __shared__ float red[4];
  const long long r = blockIdx.x;
  const long long off = row_off[r];
  const int K = row_k[r];
  const int ch = row_ch[r];
  const float s = scale[ch];
  float dot = 0.f;
  if (((off | K) & 3) == 0) {
    const float4* w = reinterpret_cast<const float4*>(P + off);
    float4* g = reinterpret_cast<float4*>(G + off);
    for (int i = threadIdx.x; i < (K >> 2); i += blockDim.x) {
      const float4 a = w[i];
      float4 b = g[i];
      dot += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
      g[i] = make_float4(b.x * s, b.y * s, b.z * s, b.w * s);
    }
  } else {
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
      const float b = G[off + i];
      dot += P[off + i] * b;
      G[off + i] = b * s;
    }
  }
  for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
  __syncthreads();
  if (threadIdx.x == 0) {
    dot = red[0] + red[1] + red[2] + red[3];
    dgamma[ch] += rstd[ch] * (dot - mean[ch] * dbeta[ch]);
  }
}

// g = dy * (y>0 if relu);  dz = round(g*scale) ; dres = g ; dbeta += sum g ; dgamma += sum g*xhat
// float4 along C; a block covers `rpb` rows per iteration over its slab of rows, partial channel
// sums are combined in shared memory before one atomicAdd per channel per block.
__global__ void __launch_bounds__(256)
act_bwd_kernel(const float4* __restrict__ dy, const float4* __restrict__ y,
               const float4* __restrict__ z, const float4* __restrict__ scale,
               const float4* __restrict__ mean, const float4* __restrict__ rstd,
               float4* __restrict__ dz, float4* __restrict__ dres, float* __restrict__ dgamma,
               float* __restrict__ dbeta, long long P, int cvec, int relu, long long rows_per_block) {
  __shared__ float4 sm_b[256], sm_g[256];
  const int tpr = cvec < 256 ? cvec : 256;  // threads per row
  const int rpb = 256 / tpr;                // rows per block iteration
  const int tr = threadIdx.x / tpr, tc = threadIdx.x % tpr;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > P) r1 = P;
  const bool active = tr < rpb;
  for (int cg = tc; cg < cvec; cg += tpr) {
    const float4 sc = scale ? scale[cg] : make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 mu = mean ? mean[cg] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 rs = rstd ? rstd[cg] : make_float4(1.f, 1.f, 1.f, 1.f);
    float4 sb = make_float4(0.f, 0.f, 0.f, 0.f), sg = sb;
    if (active) {
      for (long long r = r0 + tr; r < r1; r += rpb) {
        const long long i = r * cvec + cg;
        float4 g = dy[i];
        if (relu) {
          const float4 yy = y[i];
          g.x = yy.x > 0.f ? g.x : 0.f;
          g.y = yy.y > 0.f ? g.y : 0.f;
          g.z = yy.z > 0.f ? g.z : 0.f;
          g.w = yy.w > 0.f ? g.w : 0.f;
        }
        sb.x += g.x; sb.y += g.y; sb.z += g.z; sb.w += g.w;
        if (z) {
          const float4 zz = z[i];
          sg.x += g.x * (zz.x - mu.x) * rs.x;
          sg.y += g.y * (zz.y - mu.y) * rs.y;
          sg.z += g.z * (zz.z - mu.z) * rs.z;
          sg.w += g.w * (zz.w - mu.w) * rs.w;
        }
        if (dres) dres[i] = g;
        if (dz)
          dz[i] = make_float4(tf32_rna(g.x * sc.x), tf32_rna(g.y * sc.y), tf32_rna(g.z * sc.z),
                              tf32_rna(g.w * sc.w));
      }
    }
    if (dbeta || dgamma) {
      sm_b[threadIdx.x] = sb;
      sm_g[threadIdx.x] = sg;
      __syncthreads();
      if (tr == 0) {
        for (int k = 1; k < rpb; ++k) {
          const float4 b2 = sm_b[k * tpr + tc], g2 = sm_g[k * tpr + tc];
          sb.x += b2.x; sb.y += b2.y; sb.z += b2.z; sb.w += b2.w;
          sg.x += g2.x; sg.y += g2.y; sg.z += g2.z; sg.w += g2.w;
        }
        if (dbeta) {
          atomicAdd(dbeta + 4 * cg + 0, sb.x);
          atomicAdd(dbeta + 4 * cg + 1, sb.y);
          atomicAdd(dbeta + 4 * cg + 2, sb.z);
          atomicAdd(dbeta + 4 * cg + 3, sb.w);
        }
        if (dgamma) {
          atomicAdd(dgamma + 4 * cg + 0, sg.x);
          atomicAdd(dgamma + 4 * cg + 1, sg.y);
          atomicAdd(dgamma + 4 * cg + 2, sg.z);
          atomicAdd(dgamma + 4 * cg + 3, sg.w);
        }
      }
      __syncthreads();
    }
  }
}

// ---------------------------------------------------------------- im2col / col2im (NHWC)
// col[(n,ho,wo), (r,s,c)] ; K order matches the [Cout][kh][kw][Cin] weight layout. Kpad >= kh*kw*C.
// NHWC input with C % 4 == 0: one thread per (row, tap, 4 channels), 16-byte loads and stores.
__global__ void im2col_v4_kernel(const float4* __restrict__ x, float4* __restrict__ col, int N,
                                 int H, int W, int C4, int kh, int kw, int stride, int pad, int Ho,
                                 int Wo, int K4, int Kpad4) {
  const long long total = (long long)N * Ho * Wo * Kpad4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int k4 = (int)(i % Kpad4);
    const long long m = i / Kpad4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k4 < K4) {
      const int c4 = k4 % C4;
      const int rs = k4 / C4;
      const int s = rs % kw, r = rs / kw;
      const int wo = (int)(m % Wo);
      const long long t = m / Wo;
      const int ho = (int)(t % Ho);
      const int n = (int)(t / Ho);
      const int h = ho * stride - pad + r, w = wo * stride - pad + s;
      if (h >= 0 && h < H && w >= 0 && w < W) {
        v = x[(((long long)n * H + h) * W + w) * C4 + c4];
        v.x = tf32_rna(v.x);
        v.y = tf32_rna(v.y);
        v.z = tf32_rna(v.z);
        v.w = tf32_rna(v.w);
      }
    }
    col[i] = v;
  }
}

// NCHW fp32 image (the 7x7/2 stem, C = 3).  One block per run of kTP consecutive output pixels of
// one output row: the image patch they read (C x kh rows x ((kTP-1)*stride + kw) columns, 5.8 KB
// for the stem) is staged in shared memory with coalesced loads, and the kTP im2col rows -- which
// are CONTIGUOUS in the output (kTP * Kpad floats) -- are then written with fully coalesced
// stores, every thread decoding (pixel, kernel row, kernel column, channel) from its output index.
// (The first version wrote 21-float runs per thread with scalar stores: 0.39 ms for the 310 MB of
// the stem's im2col matrix = 0.8 TB/s, profiles/r02_launch_summary.txt.)
constexpr int kTP = 32;

__global__ void im2col_nchw_kernel(const float* __restrict__ x, float* __restrict__ col, int N, int H,
                                   int W, int C, int kh, int kw, int stride, int pad, int Ho, int Wo,
                                   int Kpad) {
  extern __shared__ float patch[];                 // [C][kh][pw]
  const int chunks = (Wo + kTP - 1) / kTP;
  const int chunk = blockIdx.x % chunks;
  const int ho = (blockIdx.x / chunks) % Ho;
  const int n = blockIdx.x / (chunks * Ho);
  const int wo0 = chunk * kTP;
  const int npix = min(kTP, Wo - wo0);
  const int pw = (kTP - 1) * stride + kw;
  const int h0 = ho * stride - pad, w0 = wo0 * stride - pad;
  for (int i = threadIdx.x; i < C * kh * pw; i += blockDim.x) {
    const int j = i % pw;
    const int r = (i / pw) % kh;
    const int c = i / (pw * kh);
    const int h = h0 + r, w = w0 + j;
    float v = 0.f;
    if (h >= 0 && h < H && w >= 0 && w < W) v = x[(((long long)n * C + c) * H + h) * W + w];
    patch[i] = tf32_rna(v);
  }
  __syncthreads();
  const int K = kh * kw * C;
  float* dst = col + (((long long)n * Ho + ho) * Wo + wo0) * Kpad;
  for (int i = threadIdx.x; i < npix * Kpad; i += blockDim.x) {
    const int p = i / Kpad, k = i - p * Kpad;
    float v = 0.f;
    if (k < K) {
      const int r = k / (kw * C), rem = k - r * (kw * C);
      const int sx = rem / C, c = rem - sx * C;
      v = patch[(c * kh + r) * pw + p * stride + sx];
    }
    dst[i] = v;
  }
}

// Operand of the direct 7x7/2 stem conv (loft_stem_conv7x7): the NCHW fp32 image as a zero-padded
// NHWC4 tensor, TF32-rounded, with its rows split into an even and an odd plane:
//   xp[n][par][hh][wp][4],  padded row 2*hh + par = iy + 3,  padded column wp = ix + 3,
//   Hh = ceil((H + 6) / 2),  Wp = W + 8.
// For one kernel row the 8 taps x 4 channels an output pixel reads (tap 7 and channel 3 carry zero
// weights) are then 128 contiguous bytes that start every 32 bytes along the row: a TMA view with
// overlapping strides delivers the im2col rows without materialising them (the im2col matrix of
// the 1024^2 stem was 310 MB written and read back per step, 0.36 ms for the write alone).
__global__ void stem_pack_kernel(const float* __restrict__ x, float4* __restrict__ xp, int N, int H,
                                 int W, int C, int Hh, int Wp) {
  const long long total = (long long)N * 2 * Hh * Wp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int wp = (int)(i % Wp);
    long long t = i / Wp;
    const int hh = (int)(t % Hh);
    t /= Hh;
    const int par = (int)(t & 1);
    const int n = (int)(t >> 1);
    const int iy = 2 * hh + par - 3, ix = wp - 3;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
      const float* src = x + ((long long)n * C * H + iy) * W + ix;
      const long long plane = (long long)H * W;
      v.x = tf32_rna(src[0]);
      if (C > 1) v.y = tf32_rna(src[plane]);
      if (C > 2) v.z = tf32_rna(src[2 * plane]);
      if (C > 3) v.w = tf32_rna(src[3 * plane]);
    }
    xp[i] = v;
  }
}

// dst[r, :] = src[idx[r], :] (gather) or dst[idx[r], :] += src[r, :] (scatter-add; idx unique),
// rows of `cols4` float4.  Lets the FOA head read the positives' rows of the bbox head's RoI
// features instead of a third RoIAlign over the same RoIs, and add its gradient back into the
// same rows of the bbox features' gradient.
template <bool kScatterAdd>
__global__ void rows_kernel(const float4* __restrict__ src, const long long* __restrict__ idx,
                            float4* __restrict__ dst, long long rows, int cols4) {
  const long long total = rows * cols4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols4;
    const int c = (int)(i - r * cols4);
    const long long j = idx[r];
    if (kScatterAdd) {
      float4 a = dst[j * cols4 + c];
      const float4 b = src[i];
      a.x += b.x;
      a.y += b.y;
      a.z += b.z;
      a.w += b.w;
      dst[j * cols4 + c] = a;
    } else {
      dst[i] = src[j * cols4 + c];
    }
  }
}

// FOA input assembly (offset_head_expand_feature.py:163-196,198-214) fused with the row gather:
//   y[b*P + p][i][j][:] = x[idx[p]][rot_kb(i, j)][:]   for the nb branches (rotation kb * 90 deg),
// i.e. gather_rows + one rot90 per branch + torch.cat in one launch; and its backward
//   gx[idx[p]][s][:] += sum_b gy[b*P + p][rot_kb^-1(s)][:]
// (the branches' gradients summed and added into the rows of the bbox features' gradient).
struct Rot4 {
  int k[4];
  int nb;
};

__device__ __forceinline__ void rot_src(int k, int S, int i, int j, int& si, int& sj) {
  switch (k & 3) {           // as rot90_kernel: y[i][j] = x[si][sj]
    case 0: si = i; sj = j; break;
    case 1: si = j; sj = S - 1 - i; break;
    case 2: si = S - 1 - i; sj = S - 1 - j; break;
    default: si = S - 1 - j; sj = i; break;
  }
}

__global__ void gather_rot_kernel(const float4* __restrict__ x, const long long* __restrict__ idx,
                                  float4* __restrict__ y, long long P, int S, int C4, Rot4 r) {
  const long long per = (long long)S * S * C4;
  const long long total = (long long)r.nb * P * per;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % C4);
    long long t = e / C4;
    const int j = (int)(t % S);
    t /= S;
    const int i = (int)(t % S);
    t /= S;
    const long long p = t % P;
    const int b = (int)(t / P);
    int si, sj;
    rot_src(r.k[b], S, i, j, si, sj);
    y[e] = x[(idx[p] * S + si) * (long long)S * C4 + (long long)sj * C4 + c];
  }
}

__global__ void scatter_rot_add_kernel(const float4* __restrict__ gy,
                                       const long long* __restrict__ idx, float4* __restrict__ gx,
                                       long long P, int S, int C4, Rot4 r) {
  const long long per = (long long)S * S * C4;
  const long long total = P * per;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % C4);
    long long t = e / C4;
    const int j = (int)(t % S);
    t /= S;
    const int i = (int)(t % S);
    const long long p = t / S;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = 0; b < r.nb; ++b) {
      int si, sj;
      rot_src(-r.k[b], S, i, j, si, sj);      // gx[i][j] collects gy_b at the inverse rotation
      const float4 v = gy[((b * P + p) * S + si) * (long long)S * C4 + (long long)sj * C4 + c];
      acc.x += v.x;
      acc.y += v.y;
      acc.z += v.z;
      acc.w += v.w;
    }
    float4* d = gx + (idx[p] * S + i) * (long long)S * C4 + (long long)j * C4 + c;
    float4 a = *d;
    a.x += acc.x;
    a.y += acc.y;
    a.z += acc.z;
    a.w += acc.w;
    *d = a;
  }
}

// ---------------------------------------------------------------- narrow 1x1 heads (Cout <= 4)
// y[p, 0..3] = x[p, :] . w[0..3, :] + b   over P pixel rows of C channels (the mask-logits conv:
// 256 -> 1 channel over P*28*28 pixels, fcn_mask_head.py:118-126).  Through the 128-row tensor-core
// tile this layer was epilogue-bound: fprop 48 us, dgrad 101 us (a rank-1 update written through
// the full epilogue), wgrad 43 us + its pad / un-pad copies, bias pass 13 us, against 25 us of HBM
// time per pass over x.  Here one warp owns a pixel row (16-byte loads, shuffle reduction) and the
// backward is ONE pass: dx = mask(x > 0) * (dy . w) rounded to TF32, dW += dy^T x, db += sum dy and
// the per-channel sum of dx (the bias gradient of the producing layer) from the same read of x.
// Row r of a space-to-depth packed tensor [(n, h, w), (i, j)] (the un-shuffled output of a 2x2 /
// stride-2 deconv GEMM) -> its pixel index in the [n, 2h + i, 2w + j] image; identity if sH == 0.
__device__ __forceinline__ long long s2d_pixel(long long r, int sH, int sW) {
  if (sH == 0) return r;
  const int ij = (int)(r & 3);
  long long t = r >> 2;
  const int w = (int)(t % sW);
  t /= sW;
  const int h = (int)(t % sH);
  const long long n = t / sH;
  return (n * (2 * sH) + 2 * h + (ij >> 1)) * (2 * sW) + 2 * w + (ij & 1);
}

template <int NOUT>    // real output channels (1..4); the padded rows of w are not touched
__global__ void __launch_bounds__(256)
narrow_fwd_kernel(const float4* __restrict__ x, const float4* __restrict__ w,
                  const float* __restrict__ b, float4* __restrict__ y, long long P, int C4, int sH,
                  int sW) {
  extern __shared__ float4 sw[];                    // [NOUT][C4]
  for (int i = threadIdx.x; i < NOUT * C4; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarp = (long long)gridDim.x * (blockDim.x >> 5);
  float bias[4] = {0.f, 0.f, 0.f, 0.f};
  if (b)
    for (int j = 0; j < 4; ++j) bias[j] = b[j];
  // two pixel rows per iteration: twice the loads in flight per warp
  for (long long p = 2 * warp; p < P; p += 2 * nwarp) {
    const bool two = p + 1 < P;
    float a[2][NOUT];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int j = 0; j < NOUT; ++j) a[r][j] = 0.f;
    const float4* x0 = x + p * C4;
    const float4* x1 = x0 + (two ? C4 : 0);
    for (int i = lane; i < C4; i += 32) {
      const float4 v0 = x0[i], v1 = x1[i];
#pragma unroll
      for (int j = 0; j < NOUT; ++j) {
        const float4 wj = sw[j * C4 + i];
        a[0][j] += v0.x * wj.x + v0.y * wj.y + v0.z * wj.z + v0.w * wj.w;
        a[1][j] += v1.x * wj.x + v1.y * wj.y + v1.z * wj.z + v1.w * wj.w;
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int j = 0; j < NOUT; ++j)
        for (int o = 16; o > 0; o >>= 1) a[r][j] += __shfl_xor_sync(0xffffffffu, a[r][j], o);
    if (lane < 2 && (lane == 0 || two)) {
      float o4[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) o4[j] = (j < NOUT ? a[lane][j] : 0.f) + bias[j];
      y[s2d_pixel(p + lane, sH, sW)] = make_float4(o4[0], o4[1], o4[2], o4[3]);
    }
  }
}

template <int T, int NOUT>       // float4 columns per lane (C = 128 * T), real output channels
__global__ void __launch_bounds__(256)
narrow_bwd_kernel(const float4* __restrict__ dy, const float4* __restrict__ x,
                  const float4* __restrict__ w, float4* __restrict__ dx, float* __restrict__ dw,
                  float* __restrict__ db, float* __restrict__ colsum, long long P, int premask,
                  int sH, int sW) {
  constexpr int C4 = 32 * T;
  constexpr int C = 4 * C4;
  __shared__ float s_red[(NOUT + 1) * C + 4];       // dW [NOUT][C], colsum [C], db [4]
  for (int i = threadIdx.x; i < (NOUT + 1) * C + 4; i += blockDim.x) s_red[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarp = (long long)gridDim.x * (blockDim.x >> 5);
  float4 wr[NOUT][T], gw[NOUT][T], cs[T];
  float gb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int t = 0; t < T; ++t) {
    cs[t] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < NOUT; ++j) {
      wr[j][t] = w[j * C4 + lane + 32 * t];
      gw[j][t] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  for (long long p = warp; p < P; p += nwarp) {
    const float4 d = dy[s2d_pixel(p, sH, sW)];
    const float dj[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) gb[j] += dj[j];
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const long long o = p * C4 + lane + 32 * t;
      const float4 xv = x[o];
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < NOUT; ++j) {
        g.x += dj[j] * wr[j][t].x; g.y += dj[j] * wr[j][t].y;
        g.z += dj[j] * wr[j][t].z; g.w += dj[j] * wr[j][t].w;
        gw[j][t].x += dj[j] * xv.x; gw[j][t].y += dj[j] * xv.y;
        gw[j][t].z += dj[j] * xv.z; gw[j][t].w += dj[j] * xv.w;
      }
      if (premask) {
        g.x = xv.x > 0.f ? g.x : 0.f; g.y = xv.y > 0.f ? g.y : 0.f;
        g.z = xv.z > 0.f ? g.z : 0.f; g.w = xv.w > 0.f ? g.w : 0.f;
      }
      g.x = tf32_rna(g.x); g.y = tf32_rna(g.y); g.z = tf32_rna(g.z); g.w = tf32_rna(g.w);
      if (dx) dx[o] = g;
      cs[t].x += g.x; cs[t].y += g.y; cs[t].z += g.z; cs[t].w += g.w;
    }
  }
  // block reduction in shared memory, then one global atomic per value and block
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const int c = (lane + 32 * t) * 4;
#pragma unroll
    for (int j = 0; j < NOUT; ++j) {
      atomicAdd(&s_red[j * C + c + 0], gw[j][t].x);
      atomicAdd(&s_red[j * C + c + 1], gw[j][t].y);
      atomicAdd(&s_red[j * C + c + 2], gw[j][t].z);
      atomicAdd(&s_red[j * C + c + 3], gw[j][t].w);
    }
    atomicAdd(&s_red[NOUT * C + c + 0], cs[t].x);
    atomicAdd(&s_red[NOUT * C + c + 1], cs[t].y);
    atomicAdd(&s_red[NOUT * C + c + 2], cs[t].z);
    atomicAdd(&s_red[NOUT * C + c + 3], cs[t].w);
  }
  if (lane == 0)                                    // every lane saw the same dy rows
    for (int j = 0; j < 4; ++j) atomicAdd(&s_red[(NOUT + 1) * C + j], gb[j]);
  __syncthreads();
  for (int i = threadIdx.x; i < (NOUT + 1) * C + 4; i += blockDim.x) {
    const float v = s_red[i];
    if (v == 0.f) continue;
    if (i < NOUT * C) {
      if (dw) atomicAdd(&dw[i], v);
    } else if (i < (NOUT + 1) * C) {
      if (colsum) atomicAdd(&colsum[i - NOUT * C], v);
    } else if (db) {
      atomicAdd(&db[i - (NOUT + 1) * C], v);
    }
  }
}

template <int NOUT>
static void narrow_bwd_launch(int grid, cudaStream_t stream, int C, const float4* dy,
                              const float4* x, const float4* w, float4* dx, float* dw, float* db,
                              float* colsum, long long P, int premask, int sH, int sW) {
  if (C == 128)
    narrow_bwd_kernel<1, NOUT><<<grid, 256, 0, stream>>>(dy, x, w, dx, dw, db, colsum, P, premask,
                                                         sH, sW);
  else if (C == 256)
    narrow_bwd_kernel<2, NOUT><<<grid, 256, 0, stream>>>(dy, x, w, dx, dw, db, colsum, P, premask,
                                                         sH, sW);
  else
    narrow_bwd_kernel<4, NOUT><<<grid, 256, 0, stream>>>(dy, x, w, dx, dw, db, colsum, P, premask,
                                                         sH, sW);
}

// dx[n,h,w,c] = sum over taps of dcol[(n,ho,wo),(r,s,c)] (gather form, no atomics), float4 over c
__global__ void col2im_v4_kernel(const float4* __restrict__ dcol, float4* __restrict__ dx,
                                 const float4* __restrict__ mask, int N, int H, int W, int C4, int kh,
                                 int kw, int stride, int pad, int Ho, int Wo, int Kpad4) {
  const long long total = (long long)N * H * W * C4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    long long t = i / C4;
    const int w = (int)(t % W);
    t /= W;
    const int h = (int)(t % H);
    const int n = (int)(t / H);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < kh; ++r) {
      const int hh = h + pad - r;
      if (hh < 0 || hh % stride) continue;
      const int ho = hh / stride;
      if (ho >= Ho) continue;
      for (int s = 0; s < kw; ++s) {
        const int ww = w + pad - s;
        if (ww < 0 || ww % stride) continue;
        const int wo = ww / stride;
        if (wo >= Wo) continue;
        const float4 v =
            dcol[(((long long)n * Ho + ho) * Wo + wo) * Kpad4 + (r * kw + s) * C4 + c4];
        acc.x += v.x;
        acc.y += v.y;
        acc.z += v.z;
        acc.w += v.w;
      }
    }
    if (mask) {
      const float4 mk = mask[i];
      acc.x = mk.x > 0.f ? acc.x : 0.f;
      acc.y = mk.y > 0.f ? acc.y : 0.f;
      acc.z = mk.z > 0.f ? acc.z : 0.f;
      acc.w = mk.w > 0.f ? acc.w : 0.f;
    }
    dx[i] = make_float4(tf32_rna(acc.x), tf32_rna(acc.y), tf32_rna(acc.z), tf32_rna(acc.w));
  }
}

// ---------------------------------------------------------------- pooling / sampling
__global__ void maxpool3x3s2_v4_kernel(const float4* __restrict__ x, float4* __restrict__ y, int N,
                                       int H, int W, int C4, int Ho, int Wo) {
  const long long total = (long long)N * Ho * Wo * C4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    long long t = i / C4;
    const int wo = (int)(t % Wo);
    t /= Wo;
    const int ho = (int)(t % Ho);
    const int n = (int)(t / Ho);
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int h = ho * 2 - 1 + r;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int w = wo * 2 - 1 + s;
        if (w < 0 || w >= W) continue;
        const float4 v = x[(((long long)n * H + h) * W + w) * C4 + c];
        m.x = fmaxf(m.x, v.x);
        m.y = fmaxf(m.y, v.y);
        m.z = fmaxf(m.z, v.z);
        m.w = fmaxf(m.w, v.w);
      }
    }
    y[i] = m;
  }
}

__global__ void maxpool3x3s2_kernel(const float* __restrict__ x, float* __restrict__ y, int N,
                                    int H, int W, int C, int Ho, int Wo) {
  const long long total = (long long)N * Ho * Wo * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long t = i / C;
    const int wo = (int)(t % Wo);
    t /= Wo;
    const int ho = (int)(t % Ho);
    const int n = (int)(t / Ho);
    float m = -INFINITY;
    for (int r = 0; r < 3; ++r) {
      const int h = ho * 2 - 1 + r;
      if (h < 0 || h >= H) continue;
      for (int s = 0; s < 3; ++s) {
        const int w = wo * 2 - 1 + s;
        if (w < 0 || w >= W) continue;
        m = fmaxf(m, x[(((long long)n * H + h) * W + w) * C + c]);
      }
    }
    y[i] = m;
  }
}

// y[n,h,w,:] = x[n,2h,2w,:]  (stride-2 1x1 conv input; FPN P6 = max_pool2d(P5,1,stride=2))
__global__ void subsample2_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int H,
                                  int W, int C, int Ho, int Wo) {
  const long long total = (long long)N * Ho * Wo * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long t = i / C;
    const int wo = (int)(t % Wo);
    t /= Wo;
    const int ho = (int)(t % Ho);
    const int n = (int)(t / Ho);
    y[i] = x[(((long long)n * H + 2 * ho) * W + 2 * wo) * C + c];
  }
}

// dx[n,h,w,:] = (h,w even) ? dy[n,h/2,w/2,:] : 0   (+ optional ReLU mask of the forward input)
__global__ void subsample2_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx,
                                      const float* __restrict__ mask, int N, int H, int W, int C,
                                      int Ho, int Wo) {
  const long long total = (long long)N * H * W * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long t = i / C;
    const int w = (int)(t % W);
    t /= W;
    const int h = (int)(t % H);
    const int n = (int)(t / H);
    float v = 0.f;
    if (!(h & 1) && !(w & 1) && (h >> 1) < Ho && (w >> 1) < Wo)
      v = dy[(((long long)n * Ho + (h >> 1)) * Wo + (w >> 1)) * C + c];
    if (mask && !(mask[i] > 0.f)) v = 0.f;
    dx[i] = v;
  }
}

// out[n,h,w,:] = base[n,h,w,:] + sum_{i,j<2} fine[n,2h+i,2w+j,:]   (FPN top-down backward)
__global__ void sum2x2_add_kernel(const float* __restrict__ fine, const float* __restrict__ base,
                                  float* __restrict__ out, int N, int H, int W, int C) {
  const long long total = (long long)N * H * W * C;
  const int Hf = 2 * H, Wf = 2 * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long t = i / C;
    const int w = (int)(t % W);
    t /= W;
    const int h = (int)(t % H);
    const int n = (int)(t / H);
    const float* f = fine + (((long long)n * Hf + 2 * h) * Wf + 2 * w) * C + c;
    float v = f[0] + f[C] + f[(long long)Wf * C] + f[(long long)Wf * C + C];
    if (base) v += base[i];
    out[i] = tf32_rna(v);
  }
}

// torch.rot90(x, k, dims=(H,W)) on [K,S,S,C] NHWC:  k=1: y[i][j] = x[j][S-1-i]
__global__ void rot90_kernel(const float* __restrict__ x, float* __restrict__ y, long long K, int S,
                             int C, int k) {
  const long long total = K * S * S * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long t = i / C;
    const int j = (int)(t % S);
    t /= S;
    const int ii = (int)(t % S);
    const long long n = t / S;
    int si, sj;
    switch (k & 3) {
      case 0: si = ii; sj = j; break;
      case 1: si = j; sj = S - 1 - ii; break;
      case 2: si = S - 1 - ii; sj = S - 1 - j; break;
      default: si = S - 1 - j; sj = ii; break;
    }
    y[i] = x[((n * S + si) * S + sj) * C + c];
  }
}

// out = round?(a + b) over [rows, C] and colsum[c] += sum over rows of out[., c] in the same pass
// (the FPN output convs' bias gradients are the per-channel sums of the gradient maps the trunk
// assembles with this add: a separate read-only pass over 178 MB otherwise).  256 % (C/4) == 0, so
// a thread's float4 column is fixed across its grid-stride loop.
__global__ void __launch_bounds__(256)
add_colsum_kernel(const float4* __restrict__ a, const float4* __restrict__ b,
                  float4* __restrict__ out, long long n4, int C4, float* __restrict__ colsum,
                  int round) {
  __shared__ float s_sum[256 * 4];
  for (int i = threadIdx.x; i < C4 * 4; i += 256) s_sum[i] = 0.f;
  __syncthreads();
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long stride = (long long)gridDim.x * 256;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n4; i += stride) {
    const float4 x = a[i], y = b[i];
    float4 v = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
    if (round) {
      v.x = tf32_rna(v.x); v.y = tf32_rna(v.y); v.z = tf32_rna(v.z); v.w = tf32_rna(v.w);
    }
    out[i] = v;
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  const int c = (threadIdx.x % C4) * 4;
  atomicAdd(&s_sum[c + 0], acc.x);
  atomicAdd(&s_sum[c + 1], acc.y);
  atomicAdd(&s_sum[c + 2], acc.z);
  atomicAdd(&s_sum[c + 3], acc.w);
  __syncthreads();
  for (int i = threadIdx.x; i < C4 * 4; i += 256)
    if (s_sum[i] != 0.f) atomicAdd(&colsum[i], s_sum[i]);
}

__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b,
                           float* __restrict__ out, long long n, int round) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float v = a[i] + b[i];
    out[i] = round ? tf32_rna(v) : v;
  }
}

// ---------------------------------------------------------------- optimizer
__global__ void sqnorm_kernel(const float* __restrict__ g, long long n, double* __restrict__ out) {
  double acc = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float v = g[i];
    acc += (double)v * v;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double sm[kT / 32];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < kT / 32; ++i) t += sm[i];
    atomicAdd(out, t);
  }
}

// clip_grad_norm_(max_norm) + SGD(momentum, weight_decay); grad_scale pre-multiplies the gradient
// (1/world_size after the all-reduce).  Also refreshes the TF32-rounded copy the GEMMs read.
__global__ void sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                           float* __restrict__ p_tf32, long long n, float lr, float momentum,
                           float wd, float max_norm, float grad_scale,
                           const double* __restrict__ sqnorm) {
  float coef = grad_scale;
  if (max_norm > 0.f && sqnorm) {
    const float total = (float)sqrt(*sqnorm) * grad_scale;
    const float c = max_norm / (total + 1e-6f);
    if (c < 1.f) coef *= c;
  }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float pv = p[i];
    const float gv = g[i] * coef + wd * pv;
    const float mv = m[i] * momentum + gv;
    m[i] = mv;
    const float nv = pv - lr * mv;
    p[i] = nv;
    if (p_tf32) p_tf32[i] = tf32_rna(nv);
  }
}

}  // namespace

extern "C" {

int loft_copy2d(const float* src, long long lds, float* dst, long long ldd, long long rows, int cols,
                int accumulate, int round_tf32, cudaStream_t stream) {
  LOFT_CHECK_ARG(src && dst, "copy2d: null pointer");
  if (rows * cols == 0) return LOFT_OK;
  copy2d_kernel<<<grid_for(rows * cols), kT, 0, stream>>>(src, lds, dst, ldd, rows, cols,
                                                           accumulate, round_tf32);
  LOFT_CUDA_LAUNCH_CHECK("copy2d");
  return LOFT_OK;
}

int loft_copy2d_multi(const loft_copy2d_t* items, int n, cudaStream_t stream) {
  LOFT_CHECK_ARG(items || n == 0, "copy2d_multi: null pointer");
  for (int o = 0; o < n; o += kMaxCopies) {
    const int m = n - o < kMaxCopies ? n - o : kMaxCopies;
    CopyJobs jobs{};
    long long big = 0;
    for (int i = 0; i < m; ++i) {
      jobs.job[i] = items[o + i];
      LOFT_CHECK_ARG((jobs.job[i].src && jobs.job[i].dst) || jobs.job[i].rows * jobs.job[i].cols == 0,
                     "copy2d_multi: null pointer in job %d", o + i);
      const long long e = jobs.job[i].rows * jobs.job[i].cols;
      big = e > big ? e : big;
    }
    if (big == 0) continue;
    dim3 grid(grid_for(big, kT, 64), m);
    copy2d_multi_kernel<<<grid, kT, 0, stream>>>(jobs);
    LOFT_CUDA_LAUNCH_CHECK("copy2d_multi");
  }
  return LOFT_OK;
}

int loft_permute_acb(const float* src, float* dst, int A, int B, int C, int accumulate,
                     int round_tf32, cudaStream_t stream) {
  LOFT_CHECK_ARG(src && dst, "permute_acb: null pointer");
  if ((long long)A * B * C == 0) return LOFT_OK;
  LOFT_CHECK_SHAPE(A <= 65535 && loft_cdiv(B, 32) <= 65535, "permute_acb: grid too large");
  dim3 grid(loft_cdiv(C, 32), loft_cdiv(B, 32), A), block(32, 8);
  permute_acb_kernel<<<grid, block, 0, stream>>>(src, dst, A, B, C, accumulate, round_tf32);
  LOFT_CUDA_LAUNCH_CHECK("permute_acb");
  return LOFT_OK;
}

int loft_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var,
                 float eps, float* scale, float* shift, float* rstd, int C, cudaStream_t stream) {
  LOFT_CHECK_ARG(gamma && beta && mean && var && scale && shift, "bn_fold: null pointer");
  bn_fold_kernel<<<loft_cdiv(C, 128), 128, 0, stream>>>(gamma, beta, mean, var, eps, scale, shift,
                                                        rstd, C);
  LOFT_CUDA_LAUNCH_CHECK("bn_fold");
  return LOFT_OK;
}

int loft_fill(float* p, long long n, float v, cudaStream_t stream) {
  LOFT_CHECK_ARG(p, "fill: null pointer");
  if (n == 0) return LOFT_OK;
  const bool al = (reinterpret_cast<uintptr_t>(p) & 15) == 0;
  const long long n4 = al ? n / 4 : 0;
  fill_kernel<<<grid_for(n4 > 0 ? n4 : n, kT, 148 * 8), kT, 0, stream>>>(
      reinterpret_cast<float4*>(p), p, n4, n, v);
  LOFT_CUDA_LAUNCH_CHECK("fill");
  return LOFT_OK;
}

int loft_bn_fold_weights(const float* P, float* T, const long long* row_off, const int* row_k,
                         const int* row_ch, const float* scale, long long rows,
                         cudaStream_t stream) {
  LOFT_CHECK_ARG(P && T && row_off && row_k && row_ch && scale, "bn_fold_weights: null pointer");
  if (rows == 0) return LOFT_OK;
  bn_fold_weights_kernel<<<(unsigned)rows, 128, 0, stream>>>(P, T, row_off, row_k, row_ch, scale);
  LOFT_CUDA_LAUNCH_CHECK("bn_fold_weights");
  return LOFT_OK;
}

int loft_bn_finalize(const float* P, float* G, const long long* row_off, const int* row_k,
                     const int* row_ch, const float* scale, const float* rstd, const float* mean,
                     const float* dbeta, float* dgamma, long long rows, cudaStream_t stream) {
  LOFT_CHECK_ARG(P && G && row_off && row_k && row_ch && scale && rstd && mean && dbeta && dgamma,
                 "bn_finalize: null pointer");
  if (rows == 0) return LOFT_OK;
  bn_finalize_kernel<<<(unsigned)rows, 128, 0, stream>>>(P, G, row_off, row_k, row_ch, scale, rstd,
                                                         mean, dbeta, dgamma);
  LOFT_CUDA_LAUNCH_CHECK("bn_finalize");
  return LOFT_OK;
}

int loft_act_bwd(const float* dy, const float* y, const float* z, const float* scale,
                 const float* mean, const float* rstd, float* dz, float* dres, float* dgamma,
                 float* dbeta, long long P, int C, int relu, cudaStream_t stream) {
  LOFT_CHECK_ARG(dy, "act_bwd: null dy");
  LOFT_CHECK_ARG(!relu || y, "act_bwd: relu needs y");
  LOFT_CHECK_SHAPE(C % 4 == 0, "act_bwd: C=%d must be a multiple of 4", C);
  if (P == 0) return LOFT_OK;
  const int cvec = C / 4;
  const int tpr = cvec < 256 ? cvec : 256;
  const int rpb = 256 / tpr;
  long long blocks = (P + rpb * 8 - 1) / (rpb * 8);  // >= 8 rows per thread-row
  const long long maxb = 148 * 6;
  if (blocks > maxb) blocks = maxb;
  if (blocks < 1) blocks = 1;
  const long long rows = (P + blocks - 1) / blocks;
  blocks = (P + rows - 1) / rows;
  act_bwd_kernel<<<(unsigned)blocks, 256, 0, stream>>>(
      reinterpret_cast<const float4*>(dy), reinterpret_cast<const float4*>(y),
      reinterpret_cast<const float4*>(z), reinterpret_cast<const float4*>(scale),
      reinterpret_cast<const float4*>(mean), reinterpret_cast<const float4*>(rstd),
      reinterpret_cast<float4*>(dz), reinterpret_cast<float4*>(dres), dgamma, dbeta, P, cvec, relu,
      rows);
  LOFT_CUDA_LAUNCH_CHECK("act_bwd");
  return LOFT_OK;
}

int loft_im2col(const float* x, float* col, int N, int H, int W, int C, int kh, int kw, int stride,
                int pad, int Kpad, int nchw_input, cudaStream_t stream) {
  LOFT_CHECK_ARG(x && col, "im2col: null pointer");
  const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
  LOFT_CHECK_SHAPE(Kpad >= kh * kw * C && Kpad % 4 == 0, "im2col: bad Kpad %d", Kpad);
  if ((long long)N * Ho * Wo == 0) return LOFT_OK;
  if (nchw_input) {
    const int chunks = (Wo + kTP - 1) / kTP;
    const size_t smem = (size_t)C * kh * ((kTP - 1) * stride + kw) * sizeof(float);
    LOFT_CHECK_SHAPE(smem <= 48 * 1024, "im2col: NCHW patch of %zu B exceeds shared memory", smem);
    im2col_nchw_kernel<<<(unsigned)((long long)N * Ho * chunks), 256, smem, stream>>>(
        x, col, N, H, W, C, kh, kw, stride, pad, Ho, Wo, Kpad);
  } else {
    LOFT_CHECK_SHAPE(C % 4 == 0, "im2col: NHWC input needs C %% 4 == 0, got %d", C);
    const long long total = (long long)N * Ho * Wo * (Kpad / 4);
    im2col_v4_kernel<<<grid_for(total, kT, 148 * 32), kT, 0, stream>>>(
        reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(col), N, H, W, C / 4, kh, kw,
        stride, pad, Ho, Wo, kh * kw * C / 4, Kpad / 4);
  }
  LOFT_CUDA_LAUNCH_CHECK("im2col");
  return LOFT_OK;
}

int loft_stem_pack(const float* x, float* xp, int N, int H, int W, int C, cudaStream_t stream) {
  LOFT_CHECK_ARG(x && xp, "stem_pack: null pointer");
  LOFT_CHECK_SHAPE(C >= 1 && C <= 4, "stem_pack: C=%d must be 1..4", C);
  const int Hh = (H + 7) / 2, Wp = W + 8;
  const long long total = (long long)N * 2 * Hh * Wp;
  if (total == 0) return LOFT_OK;
  stem_pack_kernel<<<grid_for(total, kT, 148 * 32), kT, 0, stream>>>(
      x, reinterpret_cast<float4*>(xp), N, H, W, C, Hh, Wp);
  LOFT_CUDA_LAUNCH_CHECK("stem_pack");
  return LOFT_OK;
}

int loft_gather_rows(const float* src, const long long* idx, float* dst, long long rows,
                     long long cols, cudaStream_t stream) {
  LOFT_CHECK_ARG(src && idx && dst, "gather_rows: null pointer");
  LOFT_CHECK_SHAPE(cols % 4 == 0 && cols / 4 < (1ll << 31), "gather_rows: cols=%lld must be a multiple of 4", cols);
  if (rows * cols == 0) return LOFT_OK;
  rows_kernel<false><<<grid_for(rows * (cols / 4), kT, 148 * 16), kT, 0, stream>>>(
      reinterpret_cast<const float4*>(src), idx, reinterpret_cast<float4*>(dst), rows, (int)(cols / 4));
  LOFT_CUDA_LAUNCH_CHECK("gather_rows");
  return LOFT_OK;
}

int loft_scatter_add_rows(const float* src, const long long* idx, float* dst, long long rows,
                          long long cols, cudaStream_t stream) {
  LOFT_CHECK_ARG(src && idx && dst, "scatter_add_rows: null pointer");
  LOFT_CHECK_SHAPE(cols % 4 == 0 && cols / 4 < (1ll << 31), "scatter_add_rows: cols=%lld must be a multiple of 4", cols);
  if (rows * cols == 0) return LOFT_OK;
  rows_kernel<true><<<grid_for(rows * (cols / 4), kT, 148 * 16), kT, 0, stream>>>(
      reinterpret_cast<const float4*>(src), idx, reinterpret_cast<float4*>(dst), rows, (int)(cols / 4));
  LOFT_CUDA_LAUNCH_CHECK("scatter_add_rows");
  return LOFT_OK;
}

int loft_gather_rot(const float* x, const long long* idx, float* y, long long P, int S, int C,
                    const int* k, int nb, cudaStream_t stream) {
  LOFT_CHECK_ARG(x && idx && y && k, "gather_rot: null pointer");
  LOFT_CHECK_SHAPE(C % 4 == 0 && nb >= 1 && nb <= 4 && S >= 1, "gather_rot: C=%d nb=%d S=%d", C, nb, S);
  if (P == 0) return LOFT_OK;
  Rot4 r{};
  r.nb = nb;
  for (int b = 0; b < nb; ++b) r.k[b] = ((k[b] % 4) + 4) % 4;
  gather_rot_kernel<<<grid_for((long long)nb * P * S * S * (C / 4), kT, 148 * 16), kT, 0, stream>>>(
      reinterpret_cast<const float4*>(x), idx, reinterpret_cast<float4*>(y), P, S, C / 4, r);
  LOFT_CUDA_LAUNCH_CHECK("gather_rot");
  return LOFT_OK;
}

int loft_scatter_rot_add(const float* gy, const long long* idx, float* gx, long long P, int S, int C,
                         const int* k, int nb, cudaStream_t stream) {
  LOFT_CHECK_ARG(gy && idx && gx && k, "scatter_rot_add: null pointer");
  LOFT_CHECK_SHAPE(C % 4 == 0 && nb >= 1 && nb <= 4 && S >= 1, "scatter_rot_add: C=%d nb=%d S=%d", C, nb, S);
  if (P == 0) return LOFT_OK;
  Rot4 r{};
  r.nb = nb;
  for (int b = 0; b < nb; ++b) r.k[b] = ((k[b] % 4) + 4) % 4;
  scatter_rot_add_kernel<<<grid_for(P * S * S * (C / 4), kT, 148 * 16), kT, 0, stream>>>(
      reinterpret_cast<const float4*>(gy), idx, reinterpret_cast<float4*>(gx), P, S, C / 4, r);
  LOFT_CUDA_LAUNCH_CHECK("scatter_rot_add");
  return LOFT_OK;
}

int loft_narrow_head_fwd(const float* x, const float* w, const float* b, float* y, long long P,
                         int C, int n_out, int s2d_h, int s2d_w, cudaStream_t stream) {
  LOFT_CHECK_ARG(x && w && y, "narrow_head_fwd: null pointer");
  LOFT_CHECK_SHAPE(C % 4 == 0 && C >= 4 && C <= 2048 && n_out >= 1 && n_out <= 4,
                   "narrow_head_fwd: C=%d n_out=%d", C, n_out);
  LOFT_CHECK_SHAPE(s2d_h == 0 || (s2d_h > 0 && s2d_w > 0 && P % (4ll * s2d_h * s2d_w) == 0),
                   "narrow_head_fwd: rows do not tile %d x %d x 4 sub-pixels", s2d_h, s2d_w);
  if (P == 0) return LOFT_OK;
  const long long blocks = (P + 15) / 16;
  const int grid = (int)(blocks < 148 * 8 ? blocks : 148 * 8);
  const float4* x4 = reinterpret_cast<const float4*>(x);
  const float4* w4 = reinterpret_cast<const float4*>(w);
  float4* y4 = reinterpret_cast<float4*>(y);
  const size_t smem = (size_t)n_out * C * sizeof(float);
  switch (n_out) {
    case 1: narrow_fwd_kernel<1><<<grid, 256, smem, stream>>>(x4, w4, b, y4, P, C / 4, s2d_h, s2d_w); break;
    case 2: narrow_fwd_kernel<2><<<grid, 256, smem, stream>>>(x4, w4, b, y4, P, C / 4, s2d_h, s2d_w); break;
    case 3: narrow_fwd_kernel<3><<<grid, 256, smem, stream>>>(x4, w4, b, y4, P, C / 4, s2d_h, s2d_w); break;
    default: narrow_fwd_kernel<4><<<grid, 256, smem, stream>>>(x4, w4, b, y4, P, C / 4, s2d_h, s2d_w); break;
  }
  LOFT_CUDA_LAUNCH_CHECK("narrow_head_fwd");
  return LOFT_OK;
}

int loft_narrow_head_bwd(const float* dy, const float* x, const float* w, float* dx, float* dw,
                         float* db, float* colsum, long long P, int C, int n_out, int premask,
                         int s2d_h, int s2d_w, cudaStream_t stream) {
  LOFT_CHECK_ARG(dy && x && w, "narrow_head_bwd: null pointer");
  LOFT_CHECK_SHAPE((C == 128 || C == 256 || C == 512) && n_out >= 1 && n_out <= 4,
                   "narrow_head_bwd: C=%d must be 128, 256 or 512, n_out=%d in 1..4", C, n_out);
  LOFT_CHECK_SHAPE(s2d_h == 0 || (s2d_h > 0 && s2d_w > 0 && P % (4ll * s2d_h * s2d_w) == 0),
                   "narrow_head_bwd: rows do not tile %d x %d x 4 sub-pixels", s2d_h, s2d_w);
  if (P == 0) return LOFT_OK;
  const long long blocks = (P + 7) / 8;
  const int grid = (int)(blocks < 148 * 4 ? blocks : 148 * 4);
  const float4* dy4 = reinterpret_cast<const float4*>(dy);
  const float4* x4 = reinterpret_cast<const float4*>(x);
  const float4* w4 = reinterpret_cast<const float4*>(w);
  float4* dx4 = reinterpret_cast<float4*>(dx);
  if (n_out == 1)
    narrow_bwd_launch<1>(grid, stream, C, dy4, x4, w4, dx4, dw, db, colsum, P, premask, s2d_h,
                         s2d_w);
  else if (n_out == 2)
    narrow_bwd_launch<2>(grid, stream, C, dy4, x4, w4, dx4, dw, db, colsum, P, premask, s2d_h,
                         s2d_w);
  else
    narrow_bwd_launch<4>(grid, stream, C, dy4, x4, w4, dx4, dw, db, colsum, P, premask, s2d_h,
                         s2d_w);
  LOFT_CUDA_LAUNCH_CHECK("narrow_head_bwd");
  return LOFT_OK;
}

int loft_col2im(const float* dcol, float* dx, const float* mask, int N, int H, int W, int C, int kh,
                int kw, int stride, int pad, int Kpad, cudaStream_t stream) {
  LOFT_CHECK_ARG(dcol && dx, "col2im: null pointer");
  LOFT_CHECK_SHAPE(C % 4 == 0 && Kpad % 4 == 0, "col2im: C and Kpad must be multiples of 4");
  const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
  const long long total = (long long)N * H * W * (C / 4);
  if (total == 0) return LOFT_OK;
  col2im_v4_kernel<<<grid_for(total, kT, 148 * 32), kT, 0, stream>>>(
      reinterpret_cast<const float4*>(dcol), reinterpret_cast<float4*>(dx),
      reinterpret_cast<const float4*>(mask), N, H, W, C / 4, kh, kw, stride, pad, Ho, Wo, Kpad / 4);
  LOFT_CUDA_LAUNCH_CHECK("col2im");
  return LOFT_OK;
}

int loft_maxpool3x3s2(const float* x, float* y, int N, int H, int W, int C, cudaStream_t stream) {
  LOFT_CHECK_ARG(x && y, "maxpool: null pointer");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const long long total = (long long)N * Ho * Wo * C;
  if (total == 0) return LOFT_OK;
  if (C % 4 == 0)
    maxpool3x3s2_v4_kernel<<<grid_for(total / 4, kT, 148 * 32), kT, 0, stream>>>(
        reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), N, H, W, C / 4, Ho, Wo);
  else
    maxpool3x3s2_kernel<<<grid_for(total, kT, 148 * 32), kT, 0, stream>>>(x, y, N, H, W, C, Ho, Wo);
  LOFT_CUDA_LAUNCH_CHECK("maxpool3x3s2");
  return LOFT_OK;
}

int loft_subsample2(const float* x, float* y, int N, int H, int W, int C, cudaStream_t stream) {
  LOFT_CHECK_ARG(x && y, "subsample2: null pointer");
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const long long total = (long long)N * Ho * Wo * C;
  if (total == 0) return LOFT_OK;
  subsample2_kernel<<<grid_for(total), kT, 0, stream>>>(x, y, N, H, W, C, Ho, Wo);
  LOFT_CUDA_LAUNCH_CHECK("subsample2");
  return LOFT_OK;
}

int loft_subsample2_bwd(const float* dy, float* dx, const float* mask, int N, int H, int W, int C,
                        cudaStream_t stream) {
  LOFT_CHECK_ARG(dy && dx, "subsample2_bwd: null pointer");
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const long long total = (long long)N * H * W * C;
  if (total == 0) return LOFT_OK;
  subsample2_bwd_kernel<<<grid_for(total, kT, 148 * 32), kT, 0, stream>>>(dy, dx, mask, N, H, W, C,
                                                                         Ho, Wo);
  LOFT_CUDA_LAUNCH_CHECK("subsample2_bwd");
  return LOFT_OK;
}

int loft_sum2x2_add(const float* fine, const float* base, float* out, int N, int H, int W, int C,
                    cudaStream_t stream) {
  LOFT_CHECK_ARG(fine && out, "sum2x2_add: null pointer");
  const long long total = (long long)N * H * W * C;
  if (total == 0) return LOFT_OK;
  sum2x2_add_kernel<<<grid_for(total), kT, 0, stream>>>(fine, base, out, N, H, W, C);
  LOFT_CUDA_LAUNCH_CHECK("sum2x2_add");
  return LOFT_OK;
}

int loft_rot90(const float* x, float* y, long long K, int S, int C, int k, cudaStream_t stream) {
  LOFT_CHECK_ARG(x && y, "rot90: null pointer");
  if (K * S * S * C == 0) return LOFT_OK;
  rot90_kernel<<<grid_for(K * S * S * C), kT, 0, stream>>>(x, y, K, S, C, ((k % 4) + 4) % 4);
  LOFT_CUDA_LAUNCH_CHECK("rot90");
  return LOFT_OK;
}

int loft_add(const float* a, const float* b, float* out, long long n, int round_tf32,
             cudaStream_t stream) {
  LOFT_CHECK_ARG(a && b && out, "add: null pointer");
  if (n == 0) return LOFT_OK;
  add_kernel<<<grid_for(n), kT, 0, stream>>>(a, b, out, n, round_tf32);
  LOFT_CUDA_LAUNCH_CHECK("add");
  return LOFT_OK;
}

int loft_add_colsum(const float* a, const float* b, float* out, long long rows, int C,
                    float* colsum, int round_tf32, cudaStream_t stream) {
  LOFT_CHECK_ARG(a && b && out && colsum, "add_colsum: null pointer");
  LOFT_CHECK_SHAPE(C % 4 == 0 && C >= 4 && C <= 1024 && 256 % (C / 4) == 0,
                   "add_colsum: C=%d (C/4 must divide 256)", C);
  if (rows == 0) return LOFT_OK;
  const long long n4 = rows * (C / 4);
  add_colsum_kernel<<<grid_for(n4, kT, 148 * 8), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b),
      reinterpret_cast<float4*>(out), n4, C / 4, colsum, round_tf32);
  LOFT_CUDA_LAUNCH_CHECK("add_colsum");
  return LOFT_OK;
}

int loft_grad_sqnorm(const float* g, long long n, double* out, cudaStream_t stream) {
  LOFT_CHECK_ARG(g && out, "grad_sqnorm: null pointer");
  if (n == 0) return LOFT_OK;
  sqnorm_kernel<<<grid_for(n, kT, 148 * 8), kT, 0, stream>>>(g, n, out);
  LOFT_CUDA_LAUNCH_CHECK("grad_sqnorm");
  return LOFT_OK;
}

int loft_sgd_clip_step(float* p, const float* g, float* m, float* p_tf32, long long n, float lr,
                       float momentum, float weight_decay, float max_norm, float grad_scale,
                       const double* sqnorm, cudaStream_t stream) {
  LOFT_CHECK_ARG(p && g && m, "sgd_clip_step: null pointer");
  if (n == 0) return LOFT_OK;
  sgd_kernel<<<grid_for(n, kT, 148 * 8), kT, 0, stream>>>(p, g, m, p_tf32, n, lr, momentum,
                                                          weight_decay, max_norm, grad_scale,
                                                          sqnorm);
  LOFT_CUDA_LAUNCH_CHECK("sgd_clip_step");
  return LOFT_OK;
}

}  // extern "C"
