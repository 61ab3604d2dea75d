// tcgen05 / TMA implicit-GEMM core for every dense contraction on the LOFT path
// (conv fprop / dgrad / wgrad, Linear, deconv) -- TF32 operands (fp32 in HBM), fp32 accumulate
// in TMEM.  One persistent warp-specialised kernel, six operand-fetch modes.
//
// Orientation: the UMMA "M" side (128 TMEM lanes) is the OUTPUT-CHANNEL axis and the UMMA "N"
// side (TMEM columns, <=256) is the PIXEL axis, i.e. we compute D[c, p] = sum_k A[c,k] B[p,k].
// That way an epilogue thread owns one channel: bias / BN scale+shift are one register each and
// a warp's store for a given pixel is 128 contiguous bytes of the NHWC tensor.
//
// Replaces (reference call sites): cuDNN conv fwd/bwd behind torch.nn.Conv2d
// (mmdet/models/backbones/resnet.py:163-203, necks/fpn.py:116-132, dense_heads/rpn_head.py:26-30,
// roi_heads/mask_heads/fcn_mask_head.py:64-104, attribute_heads/offset_head_expand_feature.py:72-77)
// and cuBLAS behind nn.Linear (bbox_heads/convfc_bbox_head.py:118-123,
// offset_head_expand_feature.py:97-104).
#include "common.cuh"
#include "loft_b200.h"
#include <stdlib.h>
#include <type_traits>

namespace {

constexpr int kStages = 4;                    // stages of the largest (256-pixel) tile
constexpr int kMaxStages = 8;                 // narrower tiles split the same arena into more
constexpr int kBlockC = 128;                  // UMMA M: channels per tile
constexpr int kMaxN = 256;                    // UMMA N max: pixels per tile
constexpr int kKB = 32;                       // k elements per stage (128 B of fp32)
constexpr int kABytes = kBlockC * kKB * 4;    // 16 KB
constexpr int kBBytes = kMaxN * kKB * 4;      // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kRowTabBytes = 2 * 2 * kMaxN * 4 + 64;  // 2 accumulator stages x (out row, residual
                                                      // row) + per-32-column validity masks
constexpr int kSmemBytes =
    kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + kRowTabBytes;
constexpr int kThreads = 320;                 // warp0 TMA, warp1 MMA, warps2-9 epilogue
constexpr int kTmemCols = 512;                // 2 accumulator stages x 256 columns

enum Mode {
  FPROP_2D = 0,    // A: W[Cout,K] K-major      B: X[P,K] K-major
  FPROP_CONV = 1,  // A: W[Cout,9*Cin] K-major  B: X[N,H,W,Cin] 4-D boxes shifted per tap
  DGRAD_2D = 2,    // A: W[Cout,Cin] MN-major (M=cin,k=cout)  B: dY[P,Cout] K-major
  DGRAD_CONV = 3,  // A: W[Cout,9*Cin] MN-major per tap       B: dY 4-D boxes shifted by -tap
  WGRAD_2D = 4,    // A: dY[P,Cout] MN-major (M=cout,k=p)     B: X[P,Cin] MN-major
  WGRAD_CONV = 5,  // same with 5-D boxes; one tap per work item
};

struct GemmParams {
  int mode;
  int nct;        // channel tiles (M side, 128 each)
  int npt;        // pixel tiles (FPROP/DGRAD) or N-side cin tiles (WGRAD)
  int num_tiles;  // total work items
  int num_kb;     // k-blocks per tile (FPROP/DGRAD); total k-blocks (WGRAD)
  int kb_per_split;
  int n_mma;  // UMMA N
  int n_half; // B rows (pixels / cin) one CTA stages per k-block: n_mma, or n_mma/2 in pair mode
  int stages;            // operand pipeline depth (barrier slots) and bytes per k-block, set by launch()
  int kgroup;            // k-blocks per barrier slot (1 or 2): one full / empty round trip per slot
  uint32_t stage_bytes;
  uint32_t tx_bytes;
  uint64_t a_desc, b_desc;  // smem descriptor templates (address field zero)
  uint32_t a_kstep, b_kstep;
  uint32_t idesc;
  // geometry (pixel side)
  int N, H, W;      // image batch / height / width of the pixel grid
  int tn, th, tw;   // pixel tile (conv modes)
  int tiles_h, tiles_w;
  int cchunks;      // Cin/32 (FPROP_CONV) or Cout/32 (DGRAD_CONV): k-chunks per tap
  int stem;         // FPROP_CONV over the packed stem image (loft_stem_conv7x7): k-block = kernel
                    // row kh, read from plane 2n + (kh & 1) at half-row h + (kh >> 1)
  int ntaps;        // 9 or 1
  int P;            // total pixels (2D modes)
  int Cm;           // number of valid channels on the M side (Cout for fprop, Cin for dgrad)
  int Cn;           // WGRAD: valid columns on the N side (Cin)
  // epilogue
  float* out;
  float* raw_out;
  const float* scale;
  const float* shift;
  const float* residual;
  const float* mask;
  long long ldo, ldr;
  int res_mode;   // 0 none, 1 same pixel, 2 nearest-upsample x2 (residual has H/2 x W/2 pixels),
                  // 3 zero-stuffed x2 (residual only at even (h,w): backward of a stride-2 1x1)
  int relu;
  int out_map;    // 0 plain rows, 1 deconv 2x2/s2 pixel shuffle (out is [N,2H,2W,Cm/4])
  int round_out;  // round `out` to TF32 (RNA) so the next MMA's operand truncation is exact
  int fast_epi;   // all element offsets of out / residual / mask fit in 31 bits
  float* colsum;  // per-channel sum over pixels of the value written to `out` (atomic accumulate):
  float* colsum2; //   the bias / BN-beta gradient of the layer that produced this dgrad's input
  long long colsum_gstride;  // stride of colsum between groups (floats)
  long long ldw;  // WGRAD: row pitch of dW
  int tap_stride; // WGRAD_CONV: column offset per tap in dW (= Cin)
  // grouped conv modes (FOA: 4 branches with their own weights in one launch)
  int groups;               // >= 1
  int group_n;              // images (RoIs) per group along the N axis of the activations
  int splits;               // WGRAD: split-K factor per (group, tap, tile)
  int pair;                 // host only: launch the cta_group::2 form (clusters of 2 CTAs)
  unsigned long long* trace;  // debug (loft_debug_set_trace): 8 globaltimer stamps per CTA
  int dbg_skip;             // debug (LOFT_GEMM_SKIP, results invalid): 1 no A loads, 2 no B loads,
                            // 4 one MMA per k-block -- attributes the k-loop time
  long long vec_gstride;    // scale / shift stride between groups (floats)
  long long out_gstride;    // WGRAD: dW stride between groups (floats)
};

struct DebugOverrides {
  long long a_desc = -1, b_desc = -1;
  long long a_kstep = -1, b_kstep = -1;
  long long idesc = -1;
};
DebugOverrides g_dbg;
unsigned long long* g_trace = nullptr;   // device buffer, 8 x u64 per CTA (tools/gemm_timeline.py)

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void tile_pixel_origin(const GemmParams& p, int ptile, int& n0, int& h0,
                                                  int& w0) {
  int twi = ptile % p.tiles_w;
  int r = ptile / p.tiles_w;
  int thi = r % p.tiles_h;
  int tni = r / p.tiles_h;
  n0 = tni * p.tn;
  h0 = thi * p.th;
  w0 = twi * p.tw;
}

// kPair = true: the kernel runs as clusters of two CTAs (one TPC) that share each tile with
// tcgen05.mma.cta_group::2 -- the pair covers 256 channels x n_mma pixels, each CTA stages its own
// 128 channel rows of A and HALF of the pixel rows of B per k-block (32 KB instead of 48 KB for
// the same 128 x 256 x 32 block of MMA work per SM), the leader CTA issues the MMAs for both, and
// each CTA drains its own 128 TMEM lanes.  Operand delivery (L2 -> SM), not the tensor pipe, is
// what bounds the 1-CTA form (profiles/r01_ncu_layer4_conv.txt: 33x operand re-fetch).
template <bool kPair, int kGroup>
__global__ void __launch_bounds__(kThreads, 1)
loft_gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmap_a,
                      const __grid_constant__ CUtensorMap tmap_b, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* full_bar = bars;                       // [kMaxStages]
  uint64_t* empty_bar = bars + kMaxStages;         // [kMaxStages]
  uint64_t* tfull_bar = bars + 2 * kMaxStages;     // [2]
  uint64_t* tempty_bar = bars + 2 * kMaxStages + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);
  // The 192 KB operand arena holds p.stages stages of p.stage_bytes each (16 KB of A + 128 B per
  // pixel column of B): 4 for 256-column tiles, up to 8 for 64-column ones.  Small tiles are
  // bound by the TMA round trip per stage, not by bytes, so depth is what they need (measured:
  // layer4's 3x3 took 51 us whatever the tile width with 4 stages).
  // Barrier slots hold p.kgroup k-blocks each: the producer and issuer loops are single-warp
  // dependent chains (try_wait ~90 cycles + address arithmetic + issue: measured ~175 ns per round
  // trip even with no loads and one MMA, LOFT_GEMM_SKIP=7), so at <= 128 columns, where a
  // k-block's MMAs take <= 133 ns, one round trip per k-block bounds the k-loop.
  const int n_stages = p.stages;
  constexpr int kgroup = kGroup;
  const uint32_t stage_bytes = p.stage_bytes;
  int* row_tab = reinterpret_cast<int*>(smem + kStages * kStageBytes + 256);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  unsigned long long* trace = p.trace ? p.trace + 16ull * blockIdx.x : nullptr;
  if (trace && threadIdx.x == 0) trace[0] = gtime();
  // pair mode: rank 0 is the leader (owns the full / tmem-empty barriers, issues the MMAs)
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  const bool leader = rank == 0u;
  const int tile0 = kPair ? (int)cluster_id_x() : (int)blockIdx.x;
  const int tile_step = kPair ? (int)cluster_nctaid_x() : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < kMaxStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kPair ? 16 : 8);   // epilogue warps of both CTAs drain a pair tile
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    if constexpr (kPair) tmem_alloc_pair(tmem_slot, kTmemCols);
    else tmem_alloc(tmem_slot, kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  // the peer's barriers must exist before anything of ours signals them (TMA complete_tx on the
  // leader's full barriers, multicast commits, remote tmem-empty arrivals)
  if constexpr (kPair) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor
  // prefetch) overlaps the tail of the previous kernel in the stream; global memory is only
  // touched after the previous grid has completed and flushed.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (trace && threadIdx.x == 0) trace[1] = gtime();

  const bool is_wgrad = p.mode >= WGRAD_2D;

  // The producer and MMA roles are executed by WHOLE warps in convergent control flow; only the
  // asynchronous instructions themselves are issued by one elected lane.  Everything they consume
  // (stage addresses, coordinates, descriptors) is then warp-uniform and lives in uniform
  // registers.  With the loops inside `if (lane == 0)` the compiler could not prove uniformity and
  // wrapped every UTMALDG / UTCHMMA / UTCBAR in an ELECT + R2UR.BROADCAST waterfall: measured
  // (LOFT_GEMM_SKIP runs, gpurun_out/gemm_timeline_c/d.txt) ~125 ns + 45 ns per MMA on the issue
  // side and ~400 ns per k-block on the load side whatever the tile width -- the k-loop of every
  // tile narrower than 256 columns ran at that floor instead of at its MMA time.
  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    const bool elected = elect_one_sync();
    int s = 0;
    uint32_t ph = 0;
    const bool conv_b = p.mode == FPROP_CONV || p.mode == DGRAD_CONV;
    for (int tile = tile0; tile < p.num_tiles; tile += tile_step) {
      int t = tile;
      // pair mode: (ct, pt) of the tile index name a PAIR of 128-channel tiles and a pair of
      // B half-tiles; this CTA stages the members with its rank
      const int ct = kPair ? (t % p.nct) * 2 + (int)rank : t % p.nct;
      t /= p.nct;
      int pt = kPair ? (t % p.npt) * 2 + (int)rank : t % p.npt, tap = 0, split = 0, grp;
      t /= p.npt;
      if (is_wgrad) {
        tap = t % p.ntaps;
        t /= p.ntaps;
        split = t % p.splits;
        grp = t / p.splits;
      } else {
        grp = t;
      }
      const int gn0 = grp * p.group_n;  // first image of this group
      int kb_begin = 0, kb_count = p.num_kb;
      if (is_wgrad) {
        kb_begin = split * p.kb_per_split;
        kb_count = min(p.kb_per_split, p.num_kb - kb_begin);
      }
      int n0 = 0, h0 = 0, w0 = 0;
      if (conv_b) {
        tile_pixel_origin(p, pt, n0, h0, w0);
        n0 += gn0;
      }
      const int tdh = tap / 3 - 1, tdw = tap % 3 - 1;  // WGRAD_CONV tap shift
      // The k-loop is specialised per mode (one switch per tile, not per k-block): the loop is a
      // single-warp dependent chain, so an indirect branch and the other modes' counters inside it
      // are per-k-block latency.  Coordinates are running counters (no division in the loop):
      //   conv fprop / dgrad: k-block = (filter tap (dh, dw), 32-channel chunk ch)
      //   conv wgrad:         k-block = 32-pixel patch (kwi, khi, kni) of the group's images
      auto k_loop = [&](auto mode_c) {
        constexpr int kM = decltype(mode_c)::value;
        int ch = 0, tp = 0, dh = -1, dw = -1;
        int kwi = 0, khi = 0, kni = 0;
        if constexpr (kM == WGRAD_CONV) {
          kwi = kb_begin % p.tiles_w;
          const int r = kb_begin / p.tiles_w;
          khi = r % p.tiles_h;
          kni = r / p.tiles_h;
        }
        const uint32_t smem_a0 = smem_u32(smem);
        // pair mode: both CTAs' loads complete on the LEADER's full barrier, which expects the
        // bytes of both (a peer complete_tx that lands before the leader's expect_tx only drives
        // the tx-count negative; the phase cannot complete before the leader's arrival)
        const uint32_t fb0 = kPair ? mapa_shared(smem_u32(&full_bar[0]), 0u) : smem_u32(&full_bar[0]);
#ifdef LOFT_KTRACE
        long long kt_wait = 0;
        const long long kt_begin = clock64();
#endif
        for (int kbi = 0; kbi < kb_count; kbi += kgroup) {
          const int nsub = min(kgroup, kb_count - kbi);
#ifdef LOFT_KTRACE
          const long long kt0 = clock64();
#endif
          mbar_wait(&empty_bar[s], ph ^ 1u);
#ifdef LOFT_KTRACE
          kt_wait += clock64() - kt0;
#endif
          const uint32_t fb = fb0 + 8u * (uint32_t)s;
          if (elected && (!kPair || leader))
            mbar_expect_tx(&full_bar[s], p.tx_bytes * (uint32_t)nsub);
#pragma unroll
          for (int sub = 0; sub < kgroup; ++sub) {
            if (sub >= nsub) break;
            const int kb = kb_begin + kbi + sub;
            const uint32_t sa = smem_a0 + (uint32_t)(s * kgroup + sub) * stage_bytes;
            const uint32_t sb = sa + kABytes;
            if (elected) {
              if constexpr (kM == FPROP_2D) {
                // (dbg_skip: stage only one operand, or none; launch() reduced tx_bytes to match)
                if (!(p.dbg_skip & 1)) tma_load_2d<kPair>(sa, &tmap_a, fb, kb * kKB, ct * kBlockC);
                if (!(p.dbg_skip & 2)) tma_load_2d<kPair>(sb, &tmap_b, fb, kb * kKB, pt * p.n_half);
              } else if constexpr (kM == FPROP_CONV) {
                tma_load_3d<kPair>(sa, &tmap_a, fb, kb * kKB, ct * kBlockC, grp);
                if (p.stem)
                  tma_load_4d<kPair>(sb, &tmap_b, fb, 0, w0, h0 + (tp >> 1), 2 * n0 + (tp & 1));
                else
                  tma_load_4d<kPair>(sb, &tmap_b, fb, ch * kKB, w0 + dw, h0 + dh, n0);
              } else if constexpr (kM == DGRAD_2D) {
                // A view: (32 cin, Cout rows, Cin/32 chunks); k-block = 32 cout rows
                tma_load_3d<kPair>(sa, &tmap_a, fb, 0, kb * kKB, ct * (kBlockC / 32));
                tma_load_2d<kPair>(sb, &tmap_b, fb, kb * kKB, pt * p.n_half);
              } else if constexpr (kM == DGRAD_CONV) {
                // A view of W[Cout][ntaps*Cin]: chunk index = (tap*Cin + cin0)/32
                tma_load_4d<kPair>(sa, &tmap_a, fb, 0, ch * kKB,
                                   tp * (p.tap_stride / 32) + ct * (kBlockC / 32), grp);
                tma_load_4d<kPair>(sb, &tmap_b, fb, ch * kKB, w0 - dw, h0 - dh, n0);
              } else if constexpr (kM == WGRAD_2D) {
                tma_load_3d<kPair>(sa, &tmap_a, fb, 0, kb * kKB, ct * (kBlockC / 32));
                tma_load_3d<kPair>(sb, &tmap_b, fb, 0, kb * kKB, pt * (p.n_half / 32));
              } else {
                const int kw0 = kwi * p.tw, kh0 = khi * p.th, kn0 = kni * p.tn + gn0;
                tma_load_5d<kPair>(sa, &tmap_a, fb, 0, kw0, kh0, kn0, ct * (kBlockC / 32));
                tma_load_5d<kPair>(sb, &tmap_b, fb, 0, kw0 + tdw, kh0 + tdh, kn0,
                                   pt * (p.n_half / 32));
              }
            }
            if constexpr (kM == FPROP_CONV || kM == DGRAD_CONV) {
              if (++ch == p.cchunks) {          // next filter tap
                ch = 0;
                ++tp;
                if (++dw == 2) {
                  dw = -1;
                  ++dh;
                }
              }
            }
            if constexpr (kM == WGRAD_CONV) {
              if (++kwi == p.tiles_w) {         // next pixel patch
                kwi = 0;
                if (++khi == p.tiles_h) {
                  khi = 0;
                  ++kni;
                }
              }
            }
          }
          if (++s == n_stages) {
            s = 0;
            ph ^= 1u;
          }
        }
#ifdef LOFT_KTRACE
        if (trace && tile == tile0 && elected) {   // producer: cycles waiting for a free slot / total
          trace[10] = (unsigned long long)kt_wait;
          trace[11] = (unsigned long long)(clock64() - kt_begin);
        }
#endif
      };
      switch (p.mode) {
        case FPROP_2D: k_loop(std::integral_constant<int, FPROP_2D>{}); break;
        case FPROP_CONV: k_loop(std::integral_constant<int, FPROP_CONV>{}); break;
        case DGRAD_2D: k_loop(std::integral_constant<int, DGRAD_2D>{}); break;
        case DGRAD_CONV: k_loop(std::integral_constant<int, DGRAD_CONV>{}); break;
        case WGRAD_2D: k_loop(std::integral_constant<int, WGRAD_2D>{}); break;
        default: k_loop(std::integral_constant<int, WGRAD_CONV>{}); break;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (!kPair || leader) {
      const bool elected = elect_one_sync();
      uint32_t lt = 0;
      int s = 0;
      uint32_t ph = 0;
      const uint32_t smem_base = smem_u32(smem);
      const uint32_t idesc = p.idesc;
      const uint32_t a_hi = (uint32_t)(p.a_desc >> 32), b_hi = (uint32_t)(p.b_desc >> 32);
      const uint32_t a_lo0 = (uint32_t)p.a_desc + ((smem_base >> 4) & 0x3FFFu);
      const uint32_t b_lo0 = (uint32_t)p.b_desc + (((smem_base + kABytes) >> 4) & 0x3FFFu);
      const uint32_t kb_step = stage_bytes >> 4, slot_step = kb_step * kgroup;
      for (int tile = tile0; tile < p.num_tiles; tile += tile_step, ++lt) {
        int kb_count = p.num_kb;
        if (is_wgrad) {
          const int split = (tile / (p.nct * p.npt * p.ntaps)) % p.splits;
          kb_count = min(p.kb_per_split, p.num_kb - split * p.kb_per_split);
        }
        const uint32_t as = lt & 1u, aph = (lt >> 1) & 1u;
        mbar_wait(&tempty_bar[as], aph ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * kMaxN;
#ifdef LOFT_KTRACE
        long long kt_wait = 0;
        const long long kt_begin = clock64();
#endif
        // k-step of the descriptors inside a k-block as compile-time constants (K-major: 32 B,
        // MN-major: 1024 B, in 16-byte units): the four descriptors of a k-block are then
        // independent immediate adds instead of a chain fed by constant-bank loads
        auto k_loop = [&](auto ak_c, auto bk_c) {
          constexpr uint32_t kAK = decltype(ak_c)::value, kBK = decltype(bk_c)::value;
          for (int kbi = 0; kbi < kb_count; kbi += kgroup) {
            const int nsub = min(kgroup, kb_count - kbi);
#ifdef LOFT_KTRACE
            const long long kt0 = clock64();
            mbar_wait(&full_bar[s], ph);
            kt_wait += clock64() - kt0;
#else
            mbar_wait(&full_bar[s], ph);
#endif
            tc_fence_after();
            if (trace && lt == 0 && kbi == 0 && elected) trace[2] = gtime();
            // 32-bit descriptor arithmetic: the address field (bits 0-13 of the low word, 16-byte
            // units) never carries into the LBO field for offsets inside the 192 KB arena
            const uint32_t a_lo = a_lo0 + (uint32_t)s * slot_step;
            const uint32_t b_lo = b_lo0 + (uint32_t)s * slot_step;
            if (elected) {
#pragma unroll
              for (int sub = 0; sub < kgroup; ++sub) {
                if (sub < nsub) {
#pragma unroll
                  for (int ks = 0; ks < kKB / 8; ++ks) {
#ifdef LOFT_KTRACE
                    if ((p.dbg_skip & 4) && ks > 0) break;
#endif
                    umma_tf32_lohi<kPair>(tmem_d, a_lo + sub * kb_step + ks * kAK, a_hi,
                                          b_lo + sub * kb_step + ks * kBK, b_hi, idesc,
                                          (kbi | sub | ks) != 0 ? 1u : 0u);
                  }
                }
              }
              // frees the slot in BOTH CTAs of a pair (the MMA read both shared memories)
              if constexpr (kPair) umma_commit_pair(&empty_bar[s], 3);
              else umma_commit(&empty_bar[s]);
            }
            if (++s == n_stages) {
              s = 0;
              ph ^= 1u;
            }
          }
        };
        if (p.mode <= FPROP_CONV)
          k_loop(std::integral_constant<uint32_t, 2>{}, std::integral_constant<uint32_t, 2>{});
        else if (p.mode <= DGRAD_CONV)
          k_loop(std::integral_constant<uint32_t, 64>{}, std::integral_constant<uint32_t, 2>{});
        else
          k_loop(std::integral_constant<uint32_t, 64>{}, std::integral_constant<uint32_t, 64>{});
        if (elected) {
          if constexpr (kPair) umma_commit_pair(&tfull_bar[as], 3);
          else umma_commit(&tfull_bar[as]);
          if (trace && lt == 0) trace[3] = gtime();
#ifdef LOFT_KTRACE
          if (trace && lt == 0) {   // issuer: cycles waiting for operands / total, first tile
            trace[8] = (unsigned long long)kt_wait;
            trace[9] = (unsigned long long)(clock64() - kt_begin);
            trace[12] = (unsigned long long)kb_count;
          }
#endif
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9)
    // Lane quarter q = warp & 3 (TMEM access rule); the two warps of a quarter split the 16-column
    // chunks.  Output row indices are computed ONCE per tile (one column per epilogue thread)
    // into shared memory, so the per-element loop is: LDS row, IMAD.WIDE, (LDG), FFMA, STG.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int et = (int)threadIdx.x - 64;  // 0..255
    uint32_t lt = 0;
    const uint32_t tempty_remote =
        kPair ? mapa_shared(smem_u32(&tempty_bar[0]), 0u) : 0u;   // leader's tempty_bar[0]
    for (int tile = tile0; tile < p.num_tiles; tile += tile_step, ++lt) {
      int t = tile;
      const int ct = kPair ? (t % p.nct) * 2 + (int)rank : t % p.nct;
      t /= p.nct;
      // pt here is the index of the (pair) tile: its columns are the half-tiles 2*pt, 2*pt+1 side
      // by side (n_half columns each) in pair mode
      int pt = t % p.npt, tap = 0, grp;
      t /= p.npt;
      if (is_wgrad) {
        tap = t % p.ntaps;
        grp = t / (p.ntaps * p.splits);
      } else {
        grp = t;
      }
      const int n_lim = p.groups > 1 ? (grp + 1) * p.group_n : p.N;
      const uint32_t as = lt & 1u, aph = (lt >> 1) & 1u;
      int* s_row = row_tab + as * (2 * kMaxN);
      int* s_rrow = s_row + kMaxN;
      unsigned* s_vmask = reinterpret_cast<unsigned*>(row_tab + 4 * kMaxN) + as * 8;
      if (!is_wgrad) {
        const int col = et;
        int row = -1, rrow = 0;
        if (col < p.n_mma) {
          long long pix;
          int pn = 0, ph_ = 0, pw_ = 0;
          bool ok;
          if (p.mode == FPROP_CONV || p.mode == DGRAD_CONV) {
            int n0, h0, w0;
            const int hf = kPair ? (col >= p.n_half ? 1 : 0) : 0;
            const int lc = col - hf * p.n_half;             // column within the half-tile's box
            tile_pixel_origin(p, kPair ? 2 * pt + hf : pt, n0, h0, w0);
            n0 += grp * p.group_n;
            const int thw = p.th * p.tw;
            const int jn = lc / thw, r = lc - jn * thw;
            const int jh = r / p.tw, jw = r - jh * p.tw;
            pn = n0 + jn;
            ph_ = h0 + jh;
            pw_ = w0 + jw;
            ok = (jn < p.tn) && (pn < n_lim) && (ph_ < p.H) && (pw_ < p.W);
            pix = ((long long)pn * p.H + ph_) * p.W + pw_;
          } else {
            pix = (long long)pt * p.n_mma + col;
            ok = pix < p.P;
            if (ok && (p.res_mode >= 2 || p.out_map == 1)) {
              pw_ = (int)(pix % p.W);
              const long long r = pix / p.W;
              ph_ = (int)(r % p.H);
              pn = (int)(r / p.H);
            }
          }
          if (ok) {
            row = (p.out_map == 1) ? (pn * (2 * p.H) + 2 * ph_) * (2 * p.W) + 2 * pw_ : (int)pix;
            if (p.res_mode == 2) rrow = (pn * (p.H >> 1) + (ph_ >> 1)) * (p.W >> 1) + (pw_ >> 1);
            if (p.res_mode == 3)
              rrow = ((ph_ | pw_) & 1) ? -1
                                       : (pn * ((p.H + 1) >> 1) + (ph_ >> 1)) * ((p.W + 1) >> 1) +
                                             (pw_ >> 1);
          }
        }
        s_row[col] = row;
        s_rrow[col] = rrow;
        // validity of each 32-column chunk (= the columns one epilogue pass covers): a fully
        // valid chunk takes the branch-free fast path below
        const unsigned vm = __ballot_sync(0xffffffffu, row >= 0);
        if (lane == 0) s_vmask[et >> 5] = vm;
      }
      const int c = ct * kBlockC + q * 32 + lane;
      const bool c_ok = c < p.Cm;
      // per-channel scale / shift are fetched BEFORE waiting for the accumulator: a dependent
      // global load after the wait is ~1 us of exposed latency per tile (measured: the epilogue of
      // a 128-column tile took 2.3 us with neither tensor-memory loads nor stores)
      const long long vo = (long long)grp * p.vec_gstride;
      float sc = 1.f, sh = 0.f;
      if (!is_wgrad && c_ok) {
        if (p.scale != nullptr) sc = __ldg(p.scale + vo + c);
        if (p.shift != nullptr) sh = __ldg(p.shift + vo + c);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      if (trace && lt == 0 && et == 0) trace[4] = gtime();
#ifdef LOFT_KTRACE
      const long long ke0 = clock64();
#endif
      const uint32_t taddr = tmem_base + as * kMaxN + ((uint32_t)(q * 32) << 16);

      if (is_wgrad) {
        float* drow = p.out + (long long)grp * p.out_gstride + (long long)c * p.ldw +
                      (long long)tap * p.tap_stride + (long long)pt * p.n_mma;
        const int ncol = min(p.n_mma, p.Cn - pt * p.n_mma);
        // 16-byte vector reductions (red.global.add.v4.f32): a quarter of the L2 atomic traffic
        // of scalar adds; rows of dW are 16 B aligned (Cin % 32 == 0, flat-buffer offsets % 4 == 0)
        for (int cc = half * 32; cc < p.n_mma; cc += 64) {
          float v[32];
          tmem_ld32(taddr + cc, v);
          if (c_ok) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (cc + j + 3 < ncol) {
                atomicAdd(reinterpret_cast<float4*>(drow + cc + j),
                          make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
              } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  if (cc + j + k < ncol) atomicAdd(drow + cc + j + k, v[j + k]);
              }
          }
        }
      } else {
        int ocol = c, row_add = 0;
        if (p.out_map == 1) {
          // deconv 2x2 stride 2: channel index c = (i*2 + j2)*Co + co
          const int co_n = p.Cm >> 2;
          const int ij = c / co_n;
          ocol = c - ij * co_n;
          row_add = (ij >> 1) * (2 * p.W) + (ij & 1);
        }
        const float* __restrict__ res = p.residual;
        const float* __restrict__ msk = p.mask;
        constexpr int kCh = 32;  // columns per chunk: 32 independent loads in flight per thread
        const long long ldo = p.ldo, ldr = p.ldr;
        float* __restrict__ outb = p.out + (long long)row_add * ldo + ocol;
        float* __restrict__ rawb = p.raw_out ? p.raw_out + (long long)row_add * ldo + ocol : nullptr;
        const float* __restrict__ mskb = msk ? msk + (long long)row_add * ldo + ocol : nullptr;
        const float* __restrict__ resb =
            res ? res + (p.res_mode == 1 ? (long long)row_add * ldr : 0) + ocol : nullptr;
        const int res_mode = p.res_mode, relu = p.relu, round_out = p.round_out;
        const bool contig = (p.mode == FPROP_2D || p.mode == DGRAD_2D) && p.out_map == 0;
        const bool do_colsum = p.colsum != nullptr;
        float csum = 0.f;
        // ---- the generic chunk body: any flag combination, per-column validity
        auto generic_chunk = [&](const int cc, float (&v)[kCh]) {
          int rows[kCh];
          float rv[kCh];
          if (p.fast_epi && s_vmask[cc >> 5] == 0xffffffffu) {
            // ---- fast path: every column of the chunk is a valid pixel.  No per-element
            // predicates or branches; in the 2-D modes the rows are consecutive, so addresses
            // are base + j*pitch.  (The profiled epilogue spent ~30 instructions per stored
            // value, issue-bound with two warps per scheduler: profiles/r01_ncu_epilogue.txt.)
            // 32-bit element offsets (the host only enables the fast path when they fit): 2-D
            // modes (r0 + j) * pitch, conv modes row-table lookups
            const int ldo32 = (int)ldo, ldr32 = (int)ldr;
            int* off = rows;
            if (contig) {
              const int r0 = s_row[cc];
#pragma unroll
              for (int j = 0; j < kCh; ++j) off[j] = (r0 + j) * ldo32;
            } else {
#pragma unroll
              for (int j = 0; j < kCh; ++j) off[j] = s_row[cc + j] * ldo32;
            }
#define LOFT_OFF(j) off[j]
            if (rawb != nullptr) {
#pragma unroll
              for (int j = 0; j < kCh; ++j) rawb[LOFT_OFF(j)] = v[j];
            }
#pragma unroll
            for (int j = 0; j < kCh; ++j) v[j] = fmaf(v[j], sc, sh);
            if (res_mode != 0) {
              if (res_mode == 1) {
                if (ldr == ldo) {
#pragma unroll
                  for (int j = 0; j < kCh; ++j) rv[j] = resb[LOFT_OFF(j)];
                } else {
#pragma unroll
                  for (int j = 0; j < kCh; ++j) rv[j] = resb[s_row[cc + j] * ldr32];
                }
              } else {
#pragma unroll
                for (int j = 0; j < kCh; ++j) {
                  const int rr = s_rrow[cc + j];
                  rv[j] = rr >= 0 ? resb[rr * ldr32] : 0.f;
                }
              }
              asm volatile("" ::: "memory");
#pragma unroll
              for (int j = 0; j < kCh; ++j) v[j] += rv[j];
            }
            if (relu) {
#pragma unroll
              for (int j = 0; j < kCh; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (mskb != nullptr) {
#pragma unroll
              for (int j = 0; j < kCh; ++j) rv[j] = mskb[LOFT_OFF(j)];
              asm volatile("" ::: "memory");
#pragma unroll
              for (int j = 0; j < kCh; ++j) v[j] = (rv[j] > 0.f) ? v[j] : 0.f;
            }
            if (round_out) {
#pragma unroll
              for (int j = 0; j < kCh; ++j) {
                uint32_t rr;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(rr) : "f"(v[j]));
                v[j] = __uint_as_float(rr);
              }
            }
            if (do_colsum) {
#pragma unroll
              for (int j = 0; j < kCh; ++j) csum += v[j];
            }
#ifdef LOFT_KTRACE
            if (p.dbg_skip & 8) {   // attribution: no global stores (keep v alive with one store)
              float z = 0.f;
#pragma unroll
              for (int j = 0; j < kCh; ++j) z += v[j];
              if (z == 123.456f) outb[off[0]] = z;
              return;
            }
#endif
#pragma unroll
            for (int j = 0; j < kCh; ++j) outb[LOFT_OFF(j)] = v[j];
#undef LOFT_OFF
            return;
          }
#pragma unroll
          for (int j = 0; j < kCh; ++j) rows[j] = s_row[cc + j];
          if (rawb != nullptr) {
#pragma unroll
            for (int j = 0; j < kCh; ++j)
              if (rows[j] >= 0) rawb[(long long)rows[j] * ldo] = v[j];
          }
#pragma unroll
          for (int j = 0; j < kCh; ++j) v[j] = fmaf(v[j], sc, sh);
          if (res_mode != 0) {
            if (res_mode == 1) {
#pragma unroll
              for (int j = 0; j < kCh; ++j)
                rv[j] = rows[j] >= 0 ? __ldg(resb + (long long)rows[j] * ldr) : 0.f;
            } else {
#pragma unroll
              for (int j = 0; j < kCh; ++j) {
                const int rr = s_rrow[cc + j];
                rv[j] = (rows[j] >= 0 && rr >= 0) ? __ldg(resb + (long long)rr * ldr) : 0.f;
              }
            }
            asm volatile("" ::: "memory");
#pragma unroll
            for (int j = 0; j < kCh; ++j) v[j] += rv[j];
          }
          if (relu) {
#pragma unroll
            for (int j = 0; j < kCh; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (mskb != nullptr) {
#pragma unroll
            for (int j = 0; j < kCh; ++j)
              rv[j] = rows[j] >= 0 ? __ldg(mskb + (long long)rows[j] * ldo) : 0.f;
          }
          // the ReLU-mask select lives in the (branchy) store loop on purpose: uses in a later
          // basic block keep the 32 mask loads above issued back to back
#pragma unroll
          for (int j = 0; j < kCh; ++j) {
            if (rows[j] < 0) continue;
            float acc = v[j];
            if (mskb != nullptr) acc = (rv[j] > 0.f) ? acc : 0.f;
            if (round_out) {
              uint32_t rr;
              asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(rr) : "f"(acc));
              acc = __uint_as_float(rr);
            }
            csum += acc;
            outb[(long long)rows[j] * ldo] = acc;
          }
        };
        // ---- tile-level dispatch.  A chunk of 32 fully valid columns with the common flag
        // combinations (no raw copy, residual at the output's own pitch or none) runs a body
        // specialised at compile time on {consecutive rows, residual, mask}: the generic body
        // re-decides every flag per chunk with uniform branches fed by constant-bank loads, and
        // with two warps per scheduler each of those is exposed latency (measured: 2.3 us for a
        // 128 x 128 tile with neither stores nor tensor-memory loads).  ReLU is a max against
        // 0 / -inf and the column sum is always accumulated, so they cost no variants.
        const bool spec_ok = p.fast_epi && rawb == nullptr &&
                             (res_mode == 0 || (res_mode == 1 && ldr == ldo));
        const float relu_lo = relu ? 0.f : -INFINITY;
        auto spec_loop = [&](auto contig_c, auto res_c, auto mask_c) {
          constexpr bool kContig = decltype(contig_c)::value, kRes = decltype(res_c)::value,
                         kMask = decltype(mask_c)::value;
          const int ldo32 = (int)ldo;
          for (int cc = half * kCh; cc < p.n_mma; cc += 2 * kCh) {
            float v[kCh];
#ifdef LOFT_KTRACE
            if (p.dbg_skip & 16) {
#pragma unroll
              for (int j = 0; j < kCh; ++j) v[j] = (float)(cc + j);
            } else
#endif
            tmem_ld32(taddr + cc, v);
            if (!c_ok) continue;
            if (s_vmask[cc >> 5] != 0xffffffffu) {
              // ragged chunk (image border, last tile, columns past n_mma): same arithmetic with
              // per-column predicates
              int off[kCh];
              float rv[kCh];
#pragma unroll
              for (int j = 0; j < kCh; ++j) {
                const int r = s_row[cc + j];
                off[j] = r >= 0 ? r * ldo32 : -1;
              }
              if constexpr (kRes) {
#pragma unroll
                for (int j = 0; j < kCh; ++j) rv[j] = off[j] >= 0 ? resb[off[j]] : 0.f;
              }
#pragma unroll
              for (int j = 0; j < kCh; ++j) v[j] = fmaf(v[j], sc, sh);
              if constexpr (kRes) {
#pragma unroll
                for (int j = 0; j < kCh; ++j) v[j] += rv[j];
              }
              if constexpr (kMask) {
#pragma unroll
                for (int j = 0; j < kCh; ++j) rv[j] = off[j] >= 0 ? mskb[off[j]] : 0.f;
              }
#pragma unroll
              for (int j = 0; j < kCh; ++j) {
                float a = fmaxf(v[j], relu_lo);
                if constexpr (kMask) a = (rv[j] > 0.f) ? a : 0.f;
                if (round_out) {
                  uint32_t rr;
                  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(rr) : "f"(a));
                  a = __uint_as_float(rr);
                }
                if (off[j] >= 0) {
                  csum += a;
                  outb[off[j]] = a;
                }
              }
              continue;
            }
            int off[kCh];
            if constexpr (kContig) {
              const int r0 = s_row[cc] * ldo32;
#pragma unroll
              for (int j = 0; j < kCh; ++j) off[j] = r0 + j * ldo32;
            } else {
#pragma unroll
              for (int j = 0; j < kCh; ++j) off[j] = s_row[cc + j] * ldo32;
            }
            float rv[kCh];
            if constexpr (kRes) {
#pragma unroll
              for (int j = 0; j < kCh; ++j) rv[j] = resb[off[j]];
            }
#pragma unroll
            for (int j = 0; j < kCh; ++j) v[j] = fmaf(v[j], sc, sh);
            if constexpr (kRes) {
#pragma unroll
              for (int j = 0; j < kCh; ++j) v[j] += rv[j];
            }
            if constexpr (kMask) {
#pragma unroll
              for (int j = 0; j < kCh; ++j) rv[j] = mskb[off[j]];
            }
#pragma unroll
            for (int j = 0; j < kCh; ++j) v[j] = fmaxf(v[j], relu_lo);
            if constexpr (kMask) {
#pragma unroll
              for (int j = 0; j < kCh; ++j) v[j] = (rv[j] > 0.f) ? v[j] : 0.f;
            }
            if (round_out) {
#pragma unroll
              for (int j = 0; j < kCh; ++j) {
                uint32_t rr;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(rr) : "f"(v[j]));
                v[j] = __uint_as_float(rr);
              }
            }
#pragma unroll
            for (int j = 0; j < kCh; ++j) csum += v[j];
#ifdef LOFT_KTRACE
            if (p.dbg_skip & 8) {
              if (csum == 123.456f) outb[off[0]] = csum;
              continue;
            }
#endif
#pragma unroll
            for (int j = 0; j < kCh; ++j) outb[off[j]] = v[j];
          }
        };
        using T_ = std::true_type;
        using F_ = std::false_type;
        const int variant = !spec_ok ? -1
                                     : (contig ? 1 : 0) | (res_mode == 1 ? 2 : 0) | (mskb ? 4 : 0);
        switch (variant) {
          case 0: spec_loop(F_{}, F_{}, F_{}); break;
          case 1: spec_loop(T_{}, F_{}, F_{}); break;
          case 2: spec_loop(F_{}, T_{}, F_{}); break;
          case 3: spec_loop(T_{}, T_{}, F_{}); break;
          case 4: spec_loop(F_{}, F_{}, T_{}); break;
          case 5: spec_loop(T_{}, F_{}, T_{}); break;
          case 6: spec_loop(F_{}, T_{}, T_{}); break;
          case 7: spec_loop(T_{}, T_{}, T_{}); break;
          default:
            for (int cc = half * kCh; cc < p.n_mma; cc += 2 * kCh) {
              float v[kCh];
              tmem_ld32(taddr + cc, v);
              if (!c_ok) continue;
              generic_chunk(cc, v);
            }
        }
        if (p.colsum != nullptr && c_ok) {
          atomicAdd(p.colsum + (long long)grp * p.colsum_gstride + ocol, csum);
          if (p.colsum2 != nullptr) atomicAdd(p.colsum2 + ocol, csum);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (!kPair || leader) mbar_arrive(&tempty_bar[as]);
        else mbar_arrive_cluster(tempty_remote + as * 8u);
      }
      if (trace && et == 0) {
#ifdef LOFT_KTRACE
        if (lt == 0) trace[13] = (unsigned long long)(clock64() - ke0);   // epilogue cycles, warp 2
#endif
        if (lt == 0) trace[5] = gtime();
        trace[6] = gtime();
        trace[7] = lt + 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  // neither CTA of a pair may retire (or free tensor memory) while the other can still read its
  // shared memory through an in-flight MMA or signal one of its barriers
  if constexpr (kPair) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (kPair) tmem_dealloc_pair(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

// dims/box are innermost-first; strides (bytes) are for dims 1..rank-1.
// mn_major selects the 32B-atom 128B swizzle, the only smem layout tcgen05 accepts for MN-major
// TF32 operands (UMMA LayoutType SWIZZLE_128B_BASE32B); K-major operands use plain SWIZZLE_128B.
int make_tmap(CUtensorMap* m, int rank, const void* ptr, const uint64_t* dims,
              const uint64_t* strides, const uint32_t* box, bool mn_major = false) {
  PFN_encodeTiled enc = get_encode();
  if (enc == nullptr) {
    loft_set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return LOFT_ERR_CUDA;
  }
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gs[i - 1] = strides[i - 1];
    if (box[i] == 0 || box[i] > 256) {
      loft_set_error("tensor map box dim %d = %u out of range", i, box[i]);
      return LOFT_ERR_SHAPE;
    }
    if (i > 0 && (strides[i - 1] % 16) != 0) {
      loft_set_error("tensor map stride %d = %llu not a multiple of 16 B", i,
                     (unsigned long long)strides[i - 1]);
      return LOFT_ERR_SHAPE;
    }
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) {
    loft_set_error("tensor map base pointer not 16 B aligned");
    return LOFT_ERR_ARG;
  }
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(ptr), gd,
                   gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    loft_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d)", (int)r, rank);
    return LOFT_ERR_CUDA;
  }
  return LOFT_OK;
}

int pair_slots() { return loft_num_sms() / 2; }   // CTA pairs resident at once (one per TPC)

constexpr uint64_t desc_template(uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  // SmemDescriptor (sm_100): [0,14) addr>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1,
  // [61,64) layout type (2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B)
  return ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
         (1ull << 46) | ((uint64_t)layout_type << 61);
}

constexpr uint32_t make_idesc(int n_mma, bool a_mn, bool b_mn, bool pair) {
  // InstrDescriptor: c_format F32 (1) @4, a_format TF32 (2) @7, b_format TF32 (2) @10,
  // a_major @15, b_major @16, N>>3 @17, M>>4 @24 (M = 256 across the two CTAs of a pair)
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(n_mma >> 3) << 17) | ((uint32_t)((pair ? 2 * kBlockC : kBlockC) >> 4) << 24);
}

void fill_descs(GemmParams& p, bool a_mn, bool b_mn) {
  // K-major SW128: rows of 128 B, 8-row groups 1024 B apart (SBO); LBO unused.
  // MN-major TF32 (SW128 with 32 B atoms): smem is [chunk][k rows][32 floats]; the swizzle atom
  //   is 4 k-rows x 128 B = 512 B (SBO), 32-wide MN chunks are kKB*128 B apart (LBO); one
  //   K=8 MMA consumes two atoms = 1024 B.
  p.a_desc = a_mn ? desc_template(kKB * 128, 512, 1) : desc_template(16, 1024, 2);
  p.b_desc = b_mn ? desc_template(kKB * 128, 512, 1) : desc_template(16, 1024, 2);
  p.a_kstep = a_mn ? (1024 >> 4) : (32 >> 4);
  p.b_kstep = b_mn ? (1024 >> 4) : (32 >> 4);
  p.idesc = make_idesc(p.n_mma, a_mn, b_mn, p.pair != 0);
  p.n_half = p.pair ? p.n_mma / 2 : p.n_mma;
  if (g_dbg.a_desc >= 0) p.a_desc = (uint64_t)g_dbg.a_desc;
  if (g_dbg.b_desc >= 0) p.b_desc = (uint64_t)g_dbg.b_desc;
  if (g_dbg.a_kstep >= 0) p.a_kstep = (uint32_t)g_dbg.a_kstep;
  if (g_dbg.b_kstep >= 0) p.b_kstep = (uint32_t)g_dbg.b_kstep;
  if (g_dbg.idesc >= 0) p.idesc = (uint32_t)g_dbg.idesc;
}

int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p_in, cudaStream_t stream) {
  GemmParams p = p_in;
  p.trace = g_trace;
  {
    // largest pixel-row index the epilogue can form, times the widest pitch
    long long rows = (p.mode == FPROP_CONV || p.mode == DGRAD_CONV) ? (long long)p.N * p.H * p.W
                                                                    : (long long)p.P;
    if (p.out_map == 1) rows *= 4;
    const long long ld = p.ldo > p.ldr ? p.ldo : p.ldr;
    p.fast_epi = (rows + 1) * ld < (1ll << 31) ? 1 : 0;
    // operand arena: as many stages as fit (n_mma is a multiple of 16 -> stages stay 1 KB aligned)
    p.stage_bytes = (uint32_t)(kABytes + p.n_half * kKB * 4);
    int st = (kStages * kStageBytes) / (int)p.stage_bytes;
    static int max_st = -1;
    if (max_st < 0) {
      const char* e = getenv("LOFT_MAX_STAGES");
      max_st = e ? atoi(e) : kMaxStages;
      if (max_st < 2) max_st = 2;
      if (max_st > kMaxStages) max_st = kMaxStages;
    }
    st = st < max_st ? st : max_st;
    // two k-blocks per barrier slot for tiles of <= 64 columns per CTA (8 stages -> 4 slots), the
    // only ones whose MMAs (4 x 32 cycles per k-block) are shorter than a single-warp barrier round
    // trip + issue (~300 cycles): measured per shape (gpurun_out/gemm_shapes_kgroup{1,2}.txt) the
    // layer4 / P5 / P6 3x3 convs gain 10 %, wider tiles lose 3-8 % (fewer, larger refills).
    static int kg_env = -1;
    if (kg_env < 0) {
      const char* e = getenv("LOFT_KGROUP");
      kg_env = e ? atoi(e) : 0;
    }
    p.kgroup = kg_env > 0 ? (kg_env > 2 ? 2 : kg_env) : (st >= 8 && p.mode < WGRAD_2D ? 2 : 1);
    if (st / p.kgroup < 2) p.kgroup = 1;
    p.stages = st / p.kgroup;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaSuccess;
    for (const void* fn : {(const void*)loft_gemm_tf32_kernel<false, 1>,
                           (const void*)loft_gemm_tf32_kernel<false, 2>,
                           (const void*)loft_gemm_tf32_kernel<true, 1>,
                           (const void*)loft_gemm_tf32_kernel<true, 2>})
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) {
      loft_set_error("gemm: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return LOFT_ERR_CUDA;
    }
    attr_set = true;
  }
  if (p.num_tiles <= 0) return LOFT_OK;
  int grid = p.num_tiles < loft_num_sms() ? p.num_tiles : loft_num_sms();
  static int dbg_skip = -1, dbg_max_ctas = 0;
  if (dbg_skip < 0) {
    const char* e = getenv("LOFT_GEMM_SKIP");
    dbg_skip = e ? atoi(e) : 0;
    e = getenv("LOFT_GEMM_MAXCTAS");
    dbg_max_ctas = e ? atoi(e) : 0;
  }
  if (dbg_skip & 24) p.dbg_skip = dbg_skip & 24;   // epilogue attribution bits (KTRACE builds)
  if ((dbg_skip & 7) && !p.pair && p.mode == FPROP_2D) {
    p.dbg_skip = dbg_skip;
    if (dbg_skip & 1) p.tx_bytes -= kABytes;
    if (dbg_skip & 2) p.tx_bytes -= (uint32_t)(p.n_half * kKB * 4);
  }
  if (dbg_max_ctas > 0 && grid > dbg_max_ctas) grid = dbg_max_ctas;
  if (p.pair) {   // one cluster (CTA pair) per pair tile, at most one per TPC
    const int slots = pair_slots();
    grid = 2 * (p.num_tiles < slots ? p.num_tiles : slots);
    if (dbg_max_ctas > 0 && grid > dbg_max_ctas) grid = dbg_max_ctas & ~1;
  }
  static int use_pdl = -1;
  if (use_pdl < 0) {
    const char* e = getenv("LOFT_PDL");
    use_pdl = (e == nullptr || e[0] != '0') ? 1 : 0;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (use_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (p.pair) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaError_t le;
  if (p.pair)
    le = p.kgroup == 2 ? cudaLaunchKernelEx(&cfg, loft_gemm_tf32_kernel<true, 2>, ta, tb, p)
                       : cudaLaunchKernelEx(&cfg, loft_gemm_tf32_kernel<true, 1>, ta, tb, p);
  else
    le = p.kgroup == 2 ? cudaLaunchKernelEx(&cfg, loft_gemm_tf32_kernel<false, 2>, ta, tb, p)
                       : cudaLaunchKernelEx(&cfg, loft_gemm_tf32_kernel<false, 1>, ta, tb, p);
  if (le != cudaSuccess) {
    loft_set_error("loft_gemm_tf32_kernel: launch failed: %s", cudaGetErrorString(le));
    return LOFT_ERR_CUDA;
  }
  return LOFT_OK;
}

int round16(int x) { return (x + 15) & ~15; }
int round8(int x) { return (x + 7) & ~7; }

// ---- pair (cta_group::2) policy -----------------------------------------------------------
// LOFT_2CTA=0 never, 1 (default) by the rules below, 2 whenever eligible.
// Measured per shape (gpurun_out/gemm_shapes_v2_2cta{0,2}.txt): a pair launch costs ~1.5 us more
// to start and drain (cluster barrier, paired TMEM allocation) and, for the same tile width, has
// half as many independent CTAs; it wins 6-12 % where the launch runs for more than ~1.7 waves of
// 1-CTA tiles with a k-loop long enough to be operand-delivery bound (FPN/RPN P2-P3 convs, the
// mask head, the grouped FOA convs, the 12544-wide fc layers' dgrad / wgrad, wgrads over >= 32 K
// pixels), and loses 5-15 % on single-wave launches (layer3/4, P4-P6, the 1024-wide fc layers).
constexpr long long kPairMinTiles = 256;   // 1-CTA work items (148 SMs -> 1.7 waves)
constexpr int kPairMinKb = 16;             // k-blocks per tile (K >= 512)
int pair_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LOFT_2CTA");
    v = e ? atoi(e) : 1;
  }
  return v;
}
// Per-k-block time of one tile in units of "MMA columns" (a 128 x n x 32 TF32 block of MMAs takes
// 2n cycles): the MMA needs n, staging the operands needs bytes/128 at the ~64 B/cycle one SM
// ingests from L2 -- 128 + n for a 1-CTA tile (16 KB of A + 128 B per column), 128 + n/2 for a
// CTA of a pair.  +48: fixed per-tile overhead (pipeline fill, epilogue tail), as in wave_cost.
long long tile_time(int n, bool pair) {
  const int fill = pair ? 128 + n / 2 : 128 + n;
  return (n > fill ? n : fill) + 48;
}
long long fill_cost(long long tiles, int n, bool pair) {
  const long long slots = pair ? pair_slots() : loft_num_sms();
  return ((tiles + slots - 1) / slots) * tile_time(n, pair);
}

// Relative cost of covering `tiles` tiles of n columns on the machine: full waves x (n + fixed
// per-tile overhead).  Used to pick the UMMA N (pixels per tile) for small problems so that the
// grid fills the 148 SMs instead of leaving most of them idle.
long long wave_cost(long long tiles, int n) {
  const long long waves = (tiles + loft_num_sms() - 1) / loft_num_sms();
  return waves * (n + 48);
}

int forced_tile_n() {   // bring-up / tuning only: LOFT_TILE_N=256|128|64 overrides the pickers
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LOFT_TILE_N");
    v = e ? atoi(e) : 0;
  }
  return v;
}

int pick_n_2d(long long P, int nct) {
  if (P < 256) return round16((int)P);
  if (forced_tile_n() > 0) return forced_tile_n();
  int best = 256;
  long long bc = wave_cost((long long)nct * loft_cdiv(P, 256), 256);
  for (int n : {128, 64}) {
    const long long c = wave_cost((long long)nct * loft_cdiv(P, n), n);
    if (c < bc) {
      bc = c;
      best = n;
    }
  }
  return best;
}

// Pair form: UMMA N in {256,128,64} (halves of 128/64/32 rows per CTA); returns 0 if the 1-CTA
// choice n1 is predicted to be at least as fast.  nct = number of 128-channel tiles (even).
int pick_n_2d_pair(long long P, int nct, int n1, int num_kb) {
  if (pair_mode() == 0 || nct < 2 || (nct & 1)) return 0;
  if (P < 32) return 0;
  if (pair_mode() == 1 &&
      (num_kb < kPairMinKb || (long long)nct * loft_cdiv(P, n1) < kPairMinTiles))
    return 0;
  int best = 0;
  long long bc = -1;
  if (P < 256) {
    const int n = round16((int)P);
    const long long c = fill_cost(nct / 2, n, true);
    return (bc < 0 || c < bc) ? n : 0;
  }
  for (int n : {256, 128, 64}) {
    const long long c = fill_cost((long long)(nct / 2) * loft_cdiv(P, n), n, true);
    if (bc < 0 || c < bc) {
      bc = c;
      best = n;
    }
  }
  return best;
}

// Pixel tile for the N side of FPROP_CONV / DGRAD_CONV: tn*th*tw <= target.
void pixel_tile_for(int target, int N, int H, int W, int& tn, int& th, int& tw) {
  tw = W < 16 ? W : 16;
  th = H < (target / tw) ? H : (target / tw);
  if (th < 1) th = 1;
  tn = target / (tw * th);
  if (tn > N) tn = N;
  if (tn < 1) tn = 1;
}

void pick_pixel_tile(int N, int H, int W, int nct, int& tn, int& th, int& tw) {
  long long bc = -1;
  auto consider = [&](int a, int b, int c) {
    const long long tiles =
        (long long)nct * loft_cdiv(W, c) * loft_cdiv(H, b) * loft_cdiv(N, a);
    const long long cost = wave_cost(tiles, round16(a * b * c));
    if (bc < 0 || cost < bc) {
      bc = cost;
      tn = a;
      th = b;
      tw = c;
    }
  };
  if (forced_tile_n() > 0) {
    pixel_tile_for(forced_tile_n(), N, H, W, tn, th, tw);
    return;
  }
  for (int target : {256, 128, 64}) {
    int a, b, c;
    pixel_tile_for(target, N, H, W, a, b, c);
    consider(a, b, c);
  }
  // small maps (RoI heads: 7x7, 14x14): whole images per tile, any count that fits 256 columns --
  // picks the count that leaves the fewest idle SMs in the last wave
  if (W <= 16 && H * W <= kMaxN)
    for (int a = 1; a <= N && a * H * W <= kMaxN; ++a) consider(a, H, W);
}

// Pair form of the conv-mode tile choice: the box (tn, th, tw) is what ONE CTA stages (<= 128
// pixels, rounded up to 8 rows of B); a pair tile is two consecutive boxes.  Returns false if
// the 1-CTA tile (t1 columns, tiles1 tiles) is predicted to be at least as fast.
bool pick_pixel_tile_pair(int N, int H, int W, int nct, int G, long long tiles1, int n1, int& tn,
                          int& th, int& tw) {
  if (pair_mode() == 0 || nct < 2 || (nct & 1)) return false;
  if (pair_mode() == 1 && tiles1 < kPairMinTiles) return false;
  long long bc = -1;    // eligible and worthwhile: the cheapest pair tiling
  bool found = false;
  auto consider = [&](int a, int b, int c) {
    if (a * b * c > 128 || a * b * c < 8) return;
    const long long halves = (long long)loft_cdiv(W, c) * loft_cdiv(H, b) * loft_cdiv(N, a);
    const long long tiles = (long long)(nct / 2) * G * ((halves + 1) / 2);
    const long long cost = fill_cost(tiles, 2 * round8(a * b * c), true);
    if (bc < 0 || cost < bc) {
      bc = cost;
      tn = a;
      th = b;
      tw = c;
      found = true;
    }
  };
  for (int target : {128, 64, 32}) {
    int a, b, c;
    pixel_tile_for(target, N, H, W, a, b, c);
    consider(a, b, c);
  }
  if (W <= 16) {
    // small maps (RoI heads): whole maps per CTA, or an even split of the rows of one map
    for (int a = 1; a <= N && a * H * W <= 128; ++a) consider(a, H, W);
    for (int k = 2; k <= H; ++k) consider(1, loft_cdiv(H, k), W);
  }
  return found;
}

// Pixel patch of exactly kKB (=32) positions for the K side of WGRAD_CONV (OOB -> zero fill).
void pick_k_patch(int H, int W, int& tn, int& th, int& tw) {
  tw = W > 8 ? 16 : (W > 4 ? 8 : 4);
  int hmax = kKB / tw;
  th = 1;
  while (th < hmax && th < H) th <<= 1;
  tn = kKB / (tw * th);
}

// Pixel tiling of FPROP_CONV / DGRAD_CONV (p.nct = number of 128-channel tiles on entry): picks the
// 1-CTA tile, then the pair form if eligible and predicted faster.
void set_conv_tiles(GemmParams& p, int Ng, int H, int W, int G) {
  pick_pixel_tile(Ng, H, W, p.nct * G, p.tn, p.th, p.tw);
  const int n1 = round16(p.tn * p.th * p.tw);
  const long long tiles1 = (long long)p.nct * G * loft_cdiv(W, p.tw) * loft_cdiv(H, p.th) *
                           loft_cdiv(Ng, p.tn);
  int a, b, c;
  if (pick_pixel_tile_pair(Ng, H, W, p.nct, G, tiles1, n1, a, b, c)) {
    p.pair = 1;
    p.nct /= 2;
    p.tn = a;
    p.th = b;
    p.tw = c;
    p.n_mma = 2 * round8(a * b * c);
  } else {
    p.n_mma = n1;
  }
  p.tiles_w = loft_cdiv(W, p.tw);
  p.tiles_h = loft_cdiv(H, p.th);
  const int boxes = p.tiles_w * p.tiles_h * loft_cdiv(Ng, p.tn);   // per group
  p.npt = p.pair ? (boxes + 1) / 2 : boxes;
  p.num_tiles = p.nct * p.npt * G;
  const int box_bytes = p.tn * p.th * p.tw * kKB * 4;
  p.tx_bytes = p.pair ? 2 * (kABytes + box_bytes) : kABytes + box_bytes;
}

void set_epilogue(GemmParams& p, const loft_epilogue_t* e, float* out, long long ldo) {
  p.out = out;
  p.ldo = ldo;
  p.raw_out = e ? e->raw_out : nullptr;
  p.scale = e ? e->scale : nullptr;
  p.shift = e ? e->shift : nullptr;
  p.residual = e ? e->residual : nullptr;
  p.mask = e ? e->mask : nullptr;
  p.ldr = e ? (e->ldr ? e->ldr : ldo) : ldo;
  p.res_mode = (e && e->residual) ? (e->res_upsample2x == 1 ? 2 : (e->res_upsample2x == 2 ? 3 : 1)) : 0;
  p.colsum = e ? e->colsum : nullptr;
  p.colsum2 = e ? e->colsum2 : nullptr;
  p.colsum_gstride = e ? e->colsum_gstride : 0;
  p.relu = e ? e->relu : 0;
  p.out_map = e ? e->deconv_shuffle : 0;
  p.round_out = e ? e->round_out : 0;
}

}  // namespace

extern "C" {

// Debug: every later GEMM launch writes 8 globaltimer stamps per CTA (stride 16 words) into `buf`
// (device memory, >= 16 * 8 * num_SMs bytes; NULL turns it off): kernel entry, operands may be read (after the
// programmatic-dependency wait), first stage landed, last MMA of the first tile issued, first
// accumulator complete, first tile stored, last tile stored, tiles done by the CTA.
void loft_debug_set_trace(unsigned long long* buf) { g_trace = buf; }

void loft_debug_set_desc(long long a_desc, long long b_desc, long long a_kstep, long long b_kstep,
                         long long idesc) {
  g_dbg.a_desc = a_desc;
  g_dbg.b_desc = b_desc;
  g_dbg.a_kstep = a_kstep;
  g_dbg.b_kstep = b_kstep;
  g_dbg.idesc = idesc;
}

// y[P,Cout] = epi( x[P,K] . w[Cout,K]^T )
int loft_gemm_fprop(const float* x, const float* w, float* y, long long P, int K, int Cout,
                    long long ldx, long long ldw, long long ldy, int H, int W,
                    const loft_epilogue_t* epi, cudaStream_t stream) {
  LOFT_CHECK_ARG(x && w && y, "gemm_fprop: null pointer");
  LOFT_CHECK_SHAPE(P >= 0 && K > 0 && Cout > 0, "gemm_fprop: bad sizes P=%lld K=%d Cout=%d", P, K,
                   Cout);
  LOFT_CHECK_SHAPE(ldx % 4 == 0 && ldw % 4 == 0, "gemm_fprop: row pitches must be multiples of 4");
  if (P == 0) return LOFT_OK;
  GemmParams p{};
  p.mode = FPROP_2D;
  p.groups = 1;
  p.splits = 1;
  p.nct = loft_cdiv(Cout, kBlockC);
  p.n_mma = pick_n_2d(P, p.nct);
  if (const int n2 = pick_n_2d_pair(P, p.nct, p.n_mma, loft_cdiv(K, kKB))) {
    p.pair = 1;
    p.n_mma = n2;
    p.nct /= 2;
  }
  p.npt = loft_cdiv(P, p.n_mma);
  p.num_tiles = p.nct * p.npt;
  p.num_kb = loft_cdiv(K, kKB);
  p.kb_per_split = p.num_kb;
  p.ntaps = 1;
  p.tx_bytes = p.pair ? 2 * (kABytes + (p.n_mma / 2) * kKB * 4) : kABytes + p.n_mma * kKB * 4;
  p.P = (int)P;
  p.Cm = Cout;
  p.N = 1;
  p.H = H > 0 ? H : 1;
  p.W = W > 0 ? W : (int)P;
  fill_descs(p, false, false);
  set_epilogue(p, epi, y, ldy);
  CUtensorMap ta, tb;
  {
    uint64_t d[2] = {(uint64_t)K, (uint64_t)Cout};
    uint64_t s[1] = {(uint64_t)ldw * 4};
    uint32_t b[2] = {kKB, kBlockC};
    int r = make_tmap(&ta, 2, w, d, s, b);
    if (r) return r;
  }
  {
    uint64_t d[2] = {(uint64_t)K, (uint64_t)P};
    uint64_t s[1] = {(uint64_t)ldx * 4};
    uint32_t b[2] = {kKB, (uint32_t)p.n_half};
    int r = make_tmap(&tb, 2, x, d, s, b);
    if (r) return r;
  }
  return launch(ta, tb, p, stream);
}

// dx[P,Cin] = maskrelu( dy[P,Cout] . w[Cout,Cin] )
int loft_gemm_dgrad_hw(const float* dy, const float* w, float* dx, long long P, int Cin, int Cout,
                       long long lddy, long long ldw, long long lddx, int H, int W,
                       const loft_epilogue_t* epi, cudaStream_t stream) {
  LOFT_CHECK_ARG(dy && w && dx, "gemm_dgrad: null pointer");
  LOFT_CHECK_SHAPE(Cin % 32 == 0, "gemm_dgrad: Cin=%d must be a multiple of 32", Cin);
  LOFT_CHECK_SHAPE(lddy % 4 == 0 && ldw % 4 == 0, "gemm_dgrad: row pitches must be multiples of 4");
  if (P == 0) return LOFT_OK;
  GemmParams p{};
  p.mode = DGRAD_2D;
  p.groups = 1;
  p.splits = 1;
  p.nct = loft_cdiv(Cin, kBlockC);
  p.n_mma = pick_n_2d(P, p.nct);
  if (const int n2 = pick_n_2d_pair(P, p.nct, p.n_mma, loft_cdiv(Cout, kKB))) {
    p.pair = 1;
    p.n_mma = n2;
    p.nct /= 2;
  }
  p.npt = loft_cdiv(P, p.n_mma);
  p.num_tiles = p.nct * p.npt;
  p.num_kb = loft_cdiv(Cout, kKB);
  p.kb_per_split = p.num_kb;
  p.ntaps = 1;
  p.tx_bytes = p.pair ? 2 * (kABytes + (p.n_mma / 2) * kKB * 4) : kABytes + p.n_mma * kKB * 4;
  p.P = (int)P;
  p.Cm = Cin;
  p.N = 1;
  p.H = H > 0 ? H : 1;
  p.W = W > 0 ? W : (int)P;
  fill_descs(p, true, false);
  set_epilogue(p, epi, dx, lddx);
  CUtensorMap ta, tb;
  {
    uint64_t d[3] = {32, (uint64_t)Cout, (uint64_t)(Cin / 32)};
    uint64_t s[2] = {(uint64_t)ldw * 4, 128};
    uint32_t b[3] = {32, kKB, kBlockC / 32};
    int r = make_tmap(&ta, 3, w, d, s, b, true);
    if (r) return r;
  }
  {
    uint64_t d[2] = {(uint64_t)Cout, (uint64_t)P};
    uint64_t s[1] = {(uint64_t)lddy * 4};
    uint32_t b[2] = {kKB, (uint32_t)p.n_half};
    int r = make_tmap(&tb, 2, dy, d, s, b);
    if (r) return r;
  }
  return launch(ta, tb, p, stream);
}

int loft_gemm_dgrad(const float* dy, const float* w, float* dx, long long P, int Cin, int Cout,
                    long long lddy, long long ldw, long long lddx, const loft_epilogue_t* epi,
                    cudaStream_t stream) {
  return loft_gemm_dgrad_hw(dy, w, dx, P, Cin, Cout, lddy, ldw, lddx, 0, 0, epi, stream);
}

// dw[Cout,Cin] += dy[P,Cout]^T . x[P,Cin]   (atomic accumulate; caller zero-fills once per step)
int loft_gemm_wgrad(const float* dy, const float* x, float* dw, long long P, int Cin, int Cout,
                    long long lddy, long long ldx, long long lddw, cudaStream_t stream) {
  LOFT_CHECK_ARG(dy && x && dw, "gemm_wgrad: null pointer");
  LOFT_CHECK_SHAPE(Cin % 32 == 0 && Cout % 32 == 0,
                   "gemm_wgrad: Cin=%d and Cout=%d must be multiples of 32", Cin, Cout);
  LOFT_CHECK_SHAPE(lddy % 4 == 0 && ldx % 4 == 0, "gemm_wgrad: row pitches must be multiples of 4");
  if (P == 0) return LOFT_OK;
  GemmParams p{};
  p.mode = WGRAD_2D;
  p.n_mma = Cin >= 256 ? 256 : Cin;  // multiple of 32
  p.nct = loft_cdiv(Cout, kBlockC);
  // pair form: 256 output channels per CTA pair, each CTA stages half of the Cin columns
  if (pair_mode() != 0 && p.nct >= 2 && (p.nct & 1) == 0 && p.n_mma % 64 == 0 &&
      (pair_mode() >= 2 || P >= 32768 ||
       (long long)p.nct * loft_cdiv(Cin, p.n_mma) >= kPairMinTiles)) {
    p.pair = 1;
    p.nct /= 2;
  }
  p.npt = loft_cdiv(Cin, p.n_mma);
  p.ntaps = 1;
  p.num_kb = loft_cdiv(P, kKB);
  int base = p.nct * p.npt;
  int splits = (p.pair ? pair_slots() : loft_num_sms()) / base;
  if (splits < 1) splits = 1;
  if (splits > p.num_kb) splits = p.num_kb;
  p.kb_per_split = loft_cdiv(p.num_kb, splits);
  splits = loft_cdiv(p.num_kb, p.kb_per_split);
  p.num_tiles = base * splits;
  p.groups = 1;
  p.splits = splits;
  p.tx_bytes = p.pair ? 2 * (kABytes + (p.n_mma / 2) * kKB * 4) : kABytes + p.n_mma * kKB * 4;
  p.Cm = Cout;
  p.Cn = Cin;
  p.ldw = lddw;
  p.tap_stride = 0;
  p.out = dw;
  fill_descs(p, true, true);
  CUtensorMap ta, tb;
  {
    uint64_t d[3] = {32, (uint64_t)P, (uint64_t)(Cout / 32)};
    uint64_t s[2] = {(uint64_t)lddy * 4, 128};
    uint32_t b[3] = {32, kKB, kBlockC / 32};
    int r = make_tmap(&ta, 3, dy, d, s, b, true);
    if (r) return r;
  }
  {
    uint64_t d[3] = {32, (uint64_t)P, (uint64_t)(Cin / 32)};
    uint64_t s[2] = {(uint64_t)ldx * 4, 128};
    uint32_t b[3] = {32, kKB, (uint32_t)(p.n_half / 32)};
    int r = make_tmap(&tb, 3, x, d, s, b, true);
    if (r) return r;
  }
  return launch(ta, tb, p, stream);
}

// 3x3 / pad 1 / stride 1 convolution over NHWC, weights [Cout][3][3][Cin].
// Grouped form: the N images are G consecutive groups of N/G images; group g uses the weights at
// w + g*w_gstride (and scale/shift at + g*vec_gstride): the four FOA branches in one launch.
int loft_conv3x3_fprop_grouped(const float* x, const float* w, float* y, int N, int H, int W,
                               int Cin, int Cout, int G, long long w_gstride,
                               long long vec_gstride, const loft_epilogue_t* epi,
                               cudaStream_t stream) {
  LOFT_CHECK_ARG(x && w && y, "conv3x3_fprop: null pointer");
  LOFT_CHECK_SHAPE(Cin % 32 == 0, "conv3x3_fprop: Cin=%d must be a multiple of 32", Cin);
  LOFT_CHECK_SHAPE(G >= 1 && N % G == 0 && (G == 1 || w_gstride % 4 == 0),
                   "conv3x3_fprop: bad grouping N=%d G=%d", N, G);
  if (N == 0) return LOFT_OK;
  const int Ng = N / G;
  GemmParams p{};
  p.mode = FPROP_CONV;
  p.groups = G;
  p.group_n = Ng;
  p.splits = 1;
  p.vec_gstride = vec_gstride;
  p.nct = loft_cdiv(Cout, kBlockC);
  set_conv_tiles(p, Ng, H, W, G);
  p.cchunks = Cin / 32;
  p.ntaps = 9;
  p.num_kb = 9 * p.cchunks;
  p.kb_per_split = p.num_kb;
  p.N = N;
  p.H = H;
  p.W = W;
  p.Cm = Cout;
  fill_descs(p, false, false);
  set_epilogue(p, epi, y, Cout);
  CUtensorMap ta, tb;
  {
    uint64_t d[3] = {(uint64_t)9 * Cin, (uint64_t)Cout, (uint64_t)G};
    uint64_t s[2] = {(uint64_t)9 * Cin * 4,
                     (uint64_t)(G > 1 ? w_gstride : (long long)9 * Cin * Cout) * 4};
    uint32_t b[3] = {kKB, kBlockC, 1};
    int r = make_tmap(&ta, 3, w, d, s, b);
    if (r) return r;
  }
  {
    uint64_t d[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t s[3] = {(uint64_t)Cin * 4, (uint64_t)W * Cin * 4, (uint64_t)H * W * Cin * 4};
    uint32_t b[4] = {kKB, (uint32_t)p.tw, (uint32_t)p.th, (uint32_t)p.tn};
    int r = make_tmap(&tb, 4, x, d, s, b);
    if (r) return r;
  }
  return launch(ta, tb, p, stream);
}

// 7x7 / stride 2 / pad 3 conv of an image of <= 4 channels (the ResNet stem, resnet.py:525-571)
// WITHOUT an im2col matrix.  xp is the packed image of loft_stem_pack (zero-padded NHWC4, even and
// odd rows in separate planes); w is [Cout][7][8][4] (kernel row, tap padded 7 -> 8, channel
// padded to 4; the pads are zero).  For kernel row kh the 32 k-elements (8 taps x 4 channels) of
// output pixel (oy, ox) are the 128 contiguous bytes at padded row 2*oy + kh, column 2*ox: the B
// operand is a 4-D tensor map (32 floats, ox with a 32-byte stride, half-row, plane) whose second
// stride OVERLAPS the first extent -- a sliding window the TMA delivers straight into the swizzled
// operand stage (tools/probes/tma_overlap_probe.cu: accepted by cuTensorMapEncodeTiled and exact).
// 7 k-blocks per tile; fused epilogue as everywhere else.
int loft_stem_conv7x7(const float* xp, const float* w, float* y, int N, int H, int W, int Cout,
                      const loft_epilogue_t* epi, cudaStream_t stream) {
  LOFT_CHECK_ARG(xp && w && y, "stem_conv7x7: null pointer");
  LOFT_CHECK_SHAPE(N >= 0 && H >= 1 && W >= 1 && Cout >= 1, "stem_conv7x7: bad shape");
  if (N == 0) return LOFT_OK;
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const int Hh = (H + 7) / 2, Wp = W + 8;
  GemmParams p{};
  p.mode = FPROP_CONV;
  p.stem = 1;
  p.groups = 1;
  p.group_n = N;
  p.splits = 1;
  p.nct = loft_cdiv(Cout, kBlockC);
  // one image per tile (the planes of consecutive images are not consecutive rows of the view)
  p.tn = 1;
  p.tw = Wo < kMaxN ? Wo : kMaxN;
  p.th = kMaxN / p.tw < Ho ? kMaxN / p.tw : Ho;
  p.n_mma = round16(p.th * p.tw);
  p.tiles_w = loft_cdiv(Wo, p.tw);
  p.tiles_h = loft_cdiv(Ho, p.th);
  p.npt = p.tiles_w * p.tiles_h * N;
  p.num_tiles = p.nct * p.npt;
  p.tx_bytes = kABytes + p.th * p.tw * kKB * 4;
  p.cchunks = 1;
  p.ntaps = 7;
  p.num_kb = 7;
  p.kb_per_split = p.num_kb;
  p.N = N;
  p.H = Ho;
  p.W = Wo;
  p.Cm = Cout;
  fill_descs(p, false, false);
  set_epilogue(p, epi, y, Cout);
  CUtensorMap ta, tb;
  {
    uint64_t d[3] = {(uint64_t)7 * kKB, (uint64_t)Cout, 1};
    uint64_t s[2] = {(uint64_t)7 * kKB * 4, (uint64_t)7 * kKB * 4 * Cout};
    uint32_t b[3] = {kKB, kBlockC, 1};
    int r = make_tmap(&ta, 3, w, d, s, b);
    if (r) return r;
  }
  {
    uint64_t d[4] = {kKB, (uint64_t)Wo, (uint64_t)Hh, (uint64_t)2 * N};
    uint64_t s[3] = {32, (uint64_t)Wp * 16, (uint64_t)Hh * Wp * 16};
    uint32_t b[4] = {kKB, (uint32_t)p.tw, (uint32_t)p.th, 1};
    int r = make_tmap(&tb, 4, xp, d, s, b);
    if (r) return r;
  }
  return launch(ta, tb, p, stream);
}

int loft_conv3x3_fprop(const float* x, const float* w, float* y, int N, int H, int W, int Cin,
                       int Cout, const loft_epilogue_t* epi, cudaStream_t stream) {
  return loft_conv3x3_fprop_grouped(x, w, y, N, H, W, Cin, Cout, 1, 0, 0, epi, stream);
}

int loft_conv3x3_dgrad_grouped(const float* dy, const float* w, float* dx, int N, int H, int W,
                               int Cin, int Cout, int G, long long w_gstride,
                               const loft_epilogue_t* epi, cudaStream_t stream) {
  LOFT_CHECK_ARG(dy && w && dx, "conv3x3_dgrad: null pointer");
  LOFT_CHECK_SHAPE(Cin % 32 == 0 && Cout % 4 == 0,
                   "conv3x3_dgrad: Cin=%d must be a multiple of 32, Cout=%d of 4", Cin, Cout);
  LOFT_CHECK_SHAPE(G >= 1 && N % G == 0 && (G == 1 || w_gstride % 4 == 0),
                   "conv3x3_dgrad: bad grouping N=%d G=%d", N, G);
  if (N == 0) return LOFT_OK;
  const int Ng = N / G;
  GemmParams p{};
  p.mode = DGRAD_CONV;
  p.groups = G;
  p.group_n = Ng;
  p.splits = 1;
  p.nct = loft_cdiv(Cin, kBlockC);
  set_conv_tiles(p, Ng, H, W, G);
  p.cchunks = loft_cdiv(Cout, 32);
  p.ntaps = 9;
  p.num_kb = 9 * p.cchunks;
  p.kb_per_split = p.num_kb;
  p.N = N;
  p.H = H;
  p.W = W;
  p.Cm = Cin;
  p.tap_stride = Cin;
  fill_descs(p, true, false);
  set_epilogue(p, epi, dx, Cin);
  CUtensorMap ta, tb;
  {
    uint64_t d[4] = {32, (uint64_t)Cout, (uint64_t)(9 * Cin / 32), (uint64_t)G};
    uint64_t s[3] = {(uint64_t)9 * Cin * 4, 128,
                     (uint64_t)(G > 1 ? w_gstride : (long long)9 * Cin * Cout) * 4};
    uint32_t b[4] = {32, kKB, kBlockC / 32, 1};
    int r = make_tmap(&ta, 4, w, d, s, b, true);
    if (r) return r;
  }
  {
    uint64_t d[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t s[3] = {(uint64_t)Cout * 4, (uint64_t)W * Cout * 4, (uint64_t)H * W * Cout * 4};
    uint32_t b[4] = {kKB, (uint32_t)p.tw, (uint32_t)p.th, (uint32_t)p.tn};
    int r = make_tmap(&tb, 4, dy, d, s, b);
    if (r) return r;
  }
  return launch(ta, tb, p, stream);
}

int loft_conv3x3_dgrad(const float* dy, const float* w, float* dx, int N, int H, int W, int Cin,
                       int Cout, const loft_epilogue_t* epi, cudaStream_t stream) {
  return loft_conv3x3_dgrad_grouped(dy, w, dx, N, H, W, Cin, Cout, 1, 0, epi, stream);
}

int loft_conv3x3_wgrad_grouped(const float* dy, const float* x, float* dw, int N, int H, int W,
                               int Cin, int Cout, int G, long long dw_gstride,
                               cudaStream_t stream) {
  LOFT_CHECK_ARG(dy && x && dw, "conv3x3_wgrad: null pointer");
  LOFT_CHECK_SHAPE(Cin % 32 == 0 && Cout % 32 == 0,
                   "conv3x3_wgrad: Cin=%d and Cout=%d must be multiples of 32", Cin, Cout);
  LOFT_CHECK_SHAPE(G >= 1 && N % G == 0, "conv3x3_wgrad: bad grouping N=%d G=%d", N, G);
  if (N == 0) return LOFT_OK;
  const int Ng = N / G;
  GemmParams p{};
  p.mode = WGRAD_CONV;
  p.groups = G;
  p.group_n = Ng;
  p.out_gstride = dw_gstride;
  pick_k_patch(H, W, p.tn, p.th, p.tw);
  LOFT_CHECK_SHAPE(G == 1 || Ng % p.tn == 0,
                   "conv3x3_wgrad: group size %d not a multiple of the k-patch depth %d", Ng, p.tn);
  p.tiles_w = loft_cdiv(W, p.tw);
  p.tiles_h = loft_cdiv(H, p.th);
  p.n_mma = Cin >= 256 ? 256 : Cin;
  p.nct = loft_cdiv(Cout, kBlockC);
  if (pair_mode() != 0 && p.nct >= 2 && (p.nct & 1) == 0 && p.n_mma % 64 == 0 &&
      (pair_mode() >= 2 || (long long)N * H * W >= 32768)) {
    p.pair = 1;
    p.nct /= 2;
  }
  p.npt = loft_cdiv(Cin, p.n_mma);
  p.ntaps = 9;
  p.num_kb = p.tiles_w * p.tiles_h * loft_cdiv(Ng, p.tn);  // k-blocks per group
  int base = p.nct * p.npt * 9 * G;
  // floor: one full wave, never a ragged second one
  int splits = (p.pair ? pair_slots() : loft_num_sms()) / base;
  if (splits < 1) splits = 1;
  if (splits > p.num_kb) splits = p.num_kb;
  p.kb_per_split = loft_cdiv(p.num_kb, splits);
  splits = loft_cdiv(p.num_kb, p.kb_per_split);
  p.splits = splits;
  p.num_tiles = base * splits;
  p.tx_bytes = p.pair ? 2 * (kABytes + (p.n_mma / 2) * kKB * 4) : kABytes + p.n_mma * kKB * 4;
  p.N = N;
  p.H = H;
  p.W = W;
  p.Cm = Cout;
  p.Cn = Cin;
  p.ldw = (long long)9 * Cin;
  p.tap_stride = Cin;
  p.out = dw;
  fill_descs(p, true, true);
  CUtensorMap ta, tb;
  {
    uint64_t d[5] = {32, (uint64_t)W, (uint64_t)H, (uint64_t)N, (uint64_t)(Cout / 32)};
    uint64_t s[4] = {(uint64_t)Cout * 4, (uint64_t)W * Cout * 4, (uint64_t)H * W * Cout * 4, 128};
    uint32_t b[5] = {32, (uint32_t)p.tw, (uint32_t)p.th, (uint32_t)p.tn, kBlockC / 32};
    int r = make_tmap(&ta, 5, dy, d, s, b, true);
    if (r) return r;
  }
  {
    uint64_t d[5] = {32, (uint64_t)W, (uint64_t)H, (uint64_t)N, (uint64_t)(Cin / 32)};
    uint64_t s[4] = {(uint64_t)Cin * 4, (uint64_t)W * Cin * 4, (uint64_t)H * W * Cin * 4, 128};
    uint32_t b[5] = {32, (uint32_t)p.tw, (uint32_t)p.th, (uint32_t)p.tn, (uint32_t)(p.n_half / 32)};
    int r = make_tmap(&tb, 5, x, d, s, b, true);
    if (r) return r;
  }
  return launch(ta, tb, p, stream);
}

int loft_conv3x3_wgrad(const float* dy, const float* x, float* dw, int N, int H, int W, int Cin,
                       int Cout, cudaStream_t stream) {
  return loft_conv3x3_wgrad_grouped(dy, x, dw, N, H, W, Cin, Cout, 1, 0, stream);
}

}  // extern "C"
