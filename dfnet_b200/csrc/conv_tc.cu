// Implicit-GEMM 2-D convolution on tcgen05 for the DFNet feature extractor
// (reference feature/dfnet.py:74-172: VGG-16 3x3 convs, 1x1 / 5x5 adaptation convs).
//
//   out[b,y,x,n] = act( bias[n] + sum_{ky,kx,c} in[b, y+ky-pad, x+kx-pad, c] * w[n,c,ky,kx] )
//
// im2col-free: an M tile is a 16 x 8 pixel patch.  For every 64-channel slice the patch PLUS ITS
// HALO is loaded once into shared memory as [8-channel panel][patch row][patch col][16 B]
// (ONE 5-D TMA box over the NHWC tensor, zero fill outside the image by the TMA unit).  The A
// operand of filter tap (ky,kx) is then just a shifted window of that patch: 8 consecutive pixels
// of a row are one 8x16B core matrix, the next row is SBO = patch_width*16 B away and the next
// 8-channel panel LBO = patch_rows*patch_width*16 B away, so tcgen05.mma reads all KH*KW taps
// straight from the same bytes (9x / 25x fewer loads than gathering per tap).  Weights stream as
// pre-packed 8-panel chunks (one tap of one channel slice) through cp.async.bulk; accumulators
// are double-buffered in TMEM; the epilogue applies bias / ReLU and writes NHWC 16-bit (next
// layer), an optional pre-activation tap, or fp32 NCHW (the layout the reference returns).
//
// Warps: 0-3 epilogue (thread = pixel), 4 patch producer (TMA), 8 weight producer, 9 MMA issuer (warps 5-7 idle).
// Default variant k_conv_tc2 (cta_group::2): CTA pairs work on two neighbouring M tiles with one weight stream and one
// MMA instruction for both SMs; k_conv_tc (DFB_CONV_CTA_GROUP=1, single-tile launches) is the 1-CTA variant.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include <cuda.h>
#include <cudaTypedefs.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace dfb {
namespace conv {

using namespace dfb::tc;

constexpr int kTH = 16, kTW = 8;  // pixel patch of one M tile (128 pixels)
constexpr int kBStages = 4;       // weight ring (1-CTA kernel: 32 KB stages)
constexpr int kBStages2 = 8;      // weight ring of the cta_group::2 kernel (16 KB half-stages per CTA)
constexpr int kAStages = 3;       // patch ring
constexpr int kThreads = 320;

enum Bar { A_FULL = 0, A_EMPTY = 4, B_FULL2 = 8, B_EMPTY2 = 16, D_FULL2 = 24, D_EMPTY2 = 26, N_BARS2 = 28 };

struct ConvArgs {
  // cta_group::2 only: 3-D tensor map over the pair-layout weight image ([half-stages][64 rows][256 B] = 16 KB boxes)
  alignas(64) CUtensorMap tmap;
  // input patches: 5-D tensor map over the NHWC input seen as [B][C/8 panels][H][W][8 channels] (dims innermost first:
  // 8 ch, W, H, panel, B), box = [1][8 panels][PH][PW][8 ch] = exactly the shared-memory patch image
  // [panel][patch row][patch col][16 B]; the halo outside the image is zero-filled by the TMA unit
  alignas(64) CUtensorMap tmap_in;
  const __half* in;   // NHWC [B,H,W,Cin]
  const uint8_t* wimg;
  const float* bias;  // [Cout]
  __half* out;        // NHWC [B,H,W,Cout], after activation (nullable)
  __half* tap;        // NHWC [B,H,W,Cout], before activation (nullable)
  float* out_nchw;    // fp32 [B,nchw_C,H,W], before activation (nullable)
  uint16_t* out_bf16; // NHWC [B,H,W,Cout] bf16 copy of `out` (after activation; nullable): the weight-gradient kernel's
                      // operand type, written here instead of by a separate conversion kernel (NeRF-Hist training)
  const uint16_t* mask;    // NHWC 16-bit [B,H,W,Cout]: result zeroed where mask <= 0 (ReLU backward; nullable)
  const uint16_t* addend;  // NHWC 16-bit [B,H,W,Cout] added after masking (gradient of a second consumer; nullable)
  int B, H, W, Cin, Cout, KH, KW, pad, relu;
  int nchw_C;         // channels written to out_nchw (<= Cout: the data gradient of conv1_1 has 3)
  int nt;             // N tile (64, 128, 256)
  int n_ntiles;
  int tiles_x, tiles_y;
  int n_cc;           // channel slices of <= 64 channels
  int cpp;            // Cin / 8
  int PH, PW;         // patch rows / cols incl. halo
  uint32_t a_bytes;   // patch buffer size (8 panels)
  uint32_t b_bytes;   // weight stage size (tps taps x 8 panels x nt rows)
  int tps;            // filter taps per weight stage (256 / nt): narrow layers get more MMA work per stage
  int n_wst;          // weight stages per channel slice = ceil(KH*KW / tps)
  int* error_flag;
  unsigned long long* prof;  // optional [gridDim.x][8] cycle counters (dfb_debug_conv_prof): issuer waits D_EMPTY / A_FULL /
                             // B_FULL / total, epilogue wait D_FULL / total, loader wait A_EMPTY, producer wait B_EMPTY
};

#define CPROF(slot, stmt) { const long long _t = clock64(); stmt; if (a.prof) pacc[slot] += clock64() - _t; }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// bias / activation / stores of one 32-channel block of one pixel
template <typename T>
__device__ __forceinline__ float ld16f(uint16_t u);
template <>
__device__ __forceinline__ float ld16f<__half>(uint16_t u) { return __half2float(__ushort_as_half(u)); }
template <>
__device__ __forceinline__ float ld16f<__nv_bfloat16>(uint16_t u) { return __uint_as_float((uint32_t)u << 16); }

template <typename T>
__device__ __forceinline__ void epi_store(const ConvArgs& a, const uint32_t (&v)[32], int64_t m, bool valid, int n,
                                          int64_t nchw_base, int64_t plane) {
  float xv[32];
  const float4* b4 = reinterpret_cast<const float4*>(a.bias + n);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 bb = __ldg(b4 + q);
    xv[4 * q + 0] = __uint_as_float(v[4 * q + 0]) + bb.x;
    xv[4 * q + 1] = __uint_as_float(v[4 * q + 1]) + bb.y;
    xv[4 * q + 2] = __uint_as_float(v[4 * q + 2]) + bb.z;
    xv[4 * q + 3] = __uint_as_float(v[4 * q + 3]) + bb.w;
  }
  if (!valid) return;
  const int64_t o = m * a.Cout + n;
  if (a.mask) {  // ReLU backward: post-activation values are >= 0, so "active" == a non-zero, non-negative pattern
    const uint4* mk = reinterpret_cast<const uint4*>(a.mask + o);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 mm = __ldg(mk + q);
      const uint32_t w4[4] = {mm.x, mm.y, mm.z, mm.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const uint32_t lo = w4[e] & 0xffffu, hi = w4[e] >> 16;
        if (lo == 0u || (lo & 0x8000u)) xv[8 * q + 2 * e] = 0.f;
        if (hi == 0u || (hi & 0x8000u)) xv[8 * q + 2 * e + 1] = 0.f;
      }
    }
  }
  if (a.addend) {
    const uint4* ad = reinterpret_cast<const uint4*>(a.addend + o);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 mm = __ldg(ad + q);
      const uint32_t w4[4] = {mm.x, mm.y, mm.z, mm.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        xv[8 * q + 2 * e] += ld16f<T>((uint16_t)(w4[e] & 0xffffu));
        xv[8 * q + 2 * e + 1] += ld16f<T>((uint16_t)(w4[e] >> 16));
      }
    }
  }
  if (a.out_nchw) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (n + j < a.nchw_C) a.out_nchw[nchw_base + (int64_t)(n + j) * plane] = xv[j];
  }
  if (a.tap) {
    uint4* d = reinterpret_cast<uint4*>(a.tap + o);
#pragma unroll
    for (int q = 0; q < 4; ++q)
      d[q] = make_uint4(pack2<T>(xv[8 * q], xv[8 * q + 1]), pack2<T>(xv[8 * q + 2], xv[8 * q + 3]),
                        pack2<T>(xv[8 * q + 4], xv[8 * q + 5]), pack2<T>(xv[8 * q + 6], xv[8 * q + 7]));
  }
  if (a.out) {
    if (a.relu) {
#pragma unroll
      for (int j = 0; j < 32; ++j) xv[j] = fmaxf(xv[j], 0.f);
    }
    uint4* d = reinterpret_cast<uint4*>(a.out + o);
#pragma unroll
    for (int q = 0; q < 4; ++q)
      d[q] = make_uint4(pack2<T>(xv[8 * q], xv[8 * q + 1]), pack2<T>(xv[8 * q + 2], xv[8 * q + 3]),
                        pack2<T>(xv[8 * q + 4], xv[8 * q + 5]), pack2<T>(xv[8 * q + 6], xv[8 * q + 7]));
    if (a.out_bf16) {
      uint4* d2 = reinterpret_cast<uint4*>(a.out_bf16 + o);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        d2[q] = make_uint4(pack2<__nv_bfloat16>(xv[8 * q], xv[8 * q + 1]), pack2<__nv_bfloat16>(xv[8 * q + 2], xv[8 * q + 3]),
                           pack2<__nv_bfloat16>(xv[8 * q + 4], xv[8 * q + 5]), pack2<__nv_bfloat16>(xv[8 * q + 6], xv[8 * q + 7]));
    }
  }
}

// MMAs of up to `tpn` consecutive filter taps of one channel slice (KSN K=16 steps each), issued by ONE thread.  The tap's
// A operand is the patch window shifted by (ky, kx); the position advances incrementally.  (The first version recomputed
// ky = tap / KW per tap and predicated every K step inside one unrolled body: ~200 instructions per tap on the single
// issuing thread, more than the 128-512 tensor cycles a tap is worth - the issuer, not the tensor pipe, bounded every layer.)
template <int CG, int KSN>
__device__ __forceinline__ void issue_taps(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t a_step, uint32_t b_lo, uint32_t b_hi,
                                           uint32_t b_step, uint32_t b_tap, uint32_t idesc, int tpn, uint32_t kx, uint32_t kw,
                                           uint32_t wrap, uint32_t acc) {
#pragma unroll 1
  for (int tt = 0; tt < tpn; ++tt) {
    umma_f16<CG>(d_tmem, mk64(a_lo, a_hi), mk64(b_lo, b_hi), idesc, acc);
#pragma unroll
    for (int ks = 1; ks < KSN; ++ks) umma_f16<CG>(d_tmem, mk64(a_lo + ks * a_step, a_hi), mk64(b_lo + ks * b_step, b_hi), idesc, 1u);
    acc = 1u;
    b_lo += b_tap;
    ++a_lo;
    if (++kx == kw) kx = 0, a_lo += wrap;
  }
}

// CG = 2 (cta_group::2): the two CTAs of a cluster work on two neighbouring 128-pixel M tiles with ONE weight stream:
// each CTA holds half of every weight stage (its nt/2 rows of B, a 16 KB image loaded by a 2-SM TMA that credits the
// leader's barrier), the leader issues tcgen05.mma.cta_group::2 (M = 256) for the pair and tcgen05.commit multicasts
// the stage / accumulator hand-offs to both CTAs.  Per SM the L2 -> shared-memory weight traffic halves (the 1-CTA
// kernel streams 295 KB of filter per 4608 tensor cycles at 256 channels: 64 B/clk/SM, which is what bounds it) and
// the ring holds twice as many stages in the same shared memory.
template <typename T, int CG>
__device__ __forceinline__ void conv_tc_body(const ConvArgs& a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int NB = CG == 2 ? kBStages2 : kBStages;
  const int nt = a.nt;
  const uint32_t b_stage = a.b_bytes / CG;   // bytes of a weight stage held by this CTA
  const uint32_t sA = smem_u32(smem);
  const uint32_t sB = sA + kAStages * a.a_bytes;
  const uint32_t sBar = sB + NB * b_stage;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kAStages * a.a_bytes + NB * b_stage + N_BARS2 * 8);
  const int tid = threadIdx.x, warp = tid >> 5;
  auto bar = [&](int i) { return sBar + 8u * i; };
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;

  if (tid == 0) {
    // descriptor fetches overlap the set-up (and, under programmatic dependent launch, the previous kernel's tail)
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&a.tmap_in)) : "memory");
    if (CG == 2) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&a.tmap)) : "memory");
    for (int i = 0; i < kAStages; ++i) mbar_init(bar(A_FULL + i), 1), mbar_init(bar(A_EMPTY + i), 1);
    for (int i = 0; i < NB; ++i) mbar_init(bar(B_FULL2 + i), 1), mbar_init(bar(B_EMPTY2 + i), 1);
    for (int i = 0; i < 2; ++i) mbar_init(bar(D_FULL2 + i), 1), mbar_init(bar(D_EMPTY2 + i), 128 * CG);
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc<CG>(smem_u32(tmem_slot), 512);
  tc_fence_before();
  if (CG == 2) cluster_sync(); else __syncthreads();
  tc_fence_after();
  // everything above touched only this CTA's shared memory / TMEM and ran next to the previous kernel's tail; from here on
  // global memory is read and written.  The successor may start its own set-up now: every CTA of this grid is resident.
  pdl_wait();
  pdl_launch_dependents();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_img = a.tiles_x * a.tiles_y;
  const int n_mtiles = tiles_img * a.B;
  // a "tile" of the schedule = CG neighbouring M tiles x one N tile; this CTA's M tile is CG * (t / n_ntiles) + rank
  const int n_tiles = ((n_mtiles + CG - 1) / CG) * a.n_ntiles;
  const int n_taps = a.KH * a.KW;
  const uint32_t panel_stride = (uint32_t)a.PH * a.PW * 16u;
  const int unit0 = blockIdx.x / CG, n_units = gridDim.x / CG;

  if (warp == 4) {
    // ===== patch loader: ONE TMA box per 64-channel slice (patch + halo, zero-filled outside the image) ============
    // (the first version gathered the patch with 16-byte cp.async from 128 threads: ~600 address-arithmetic instructions
    // per thread and slice, and the MMA issuer waited 28-48 % of its time for patches, 70 % on the 1x1 layers)
    long long pacc[4] = {0, 0, 0, 0};
    uint32_t seq = 0;
    for (int t = unit0; t < n_tiles; t += n_units) {
      const int mt = CG * (t / a.n_ntiles) + (int)rank;
      const bool mt_ok = mt < n_mtiles;   // the last pair of an odd tile count: this CTA's half is empty
      const int b = mt_ok ? mt / tiles_img : a.B, ti = mt_ok ? mt % tiles_img : 0;   // b == B: wholly out of bounds -> zeros
      const int y0 = (ti / a.tiles_x) * kTH - a.pad, x0 = (ti % a.tiles_x) * kTW - a.pad;
      for (int cc = 0; cc < a.n_cc; ++cc, ++seq) {
        const uint32_t st = seq % kAStages, ph = (seq / kAStages) & 1;
        CPROF(0, mbar_wait(bar(A_EMPTY + st), ph ^ 1, a.error_flag));
        if (elect_one()) {
          if (CG == 1) {
            mbar_expect_tx(bar(A_FULL + st), a.a_bytes);
            tma_load_5d<1>(sA + st * a.a_bytes, &a.tmap_in, 0, x0, y0, cc * 8, b, bar(A_FULL + st));
          } else {
            const uint32_t bar_leader = bar(A_FULL + st) & 0xFEFFFFFFu;
            if (rank == 0) mbar_expect_tx(bar(A_FULL + st), 2 * a.a_bytes);
            tma_load_5d<2>(sA + st * a.a_bytes, &a.tmap_in, 0, x0, y0, cc * 8, b, bar_leader);
          }
        }
        __syncwarp();
      }
    }
    if (a.prof && (tid & 31) == 0) a.prof[blockIdx.x * 8 + 6] = pacc[0];
  } else if (warp == 8) {
    // ===== weight producer: one stage = tps taps of one channel slice ==============================
    uint32_t it = 0;
    long long pacc[4] = {0, 0, 0, 0};
    const int per_tile = a.n_cc * a.n_wst;
    for (int t = unit0; t < n_tiles; t += n_units) {
      const int ntile = t % a.n_ntiles;
      const uint8_t* src = a.wimg + (size_t)ntile * per_tile * a.b_bytes;
      for (int c = 0; c < per_tile; ++c, ++it) {
        const uint32_t stage = it % NB, ph = (it / NB) & 1;
        CPROF(0, mbar_wait(bar(B_EMPTY2 + stage), ph ^ 1, a.error_flag));
        if (elect_one()) {
          if (CG == 1) {
            mbar_expect_tx(bar(B_FULL2 + stage), a.b_bytes);
            bulk_g2s(sB + stage * b_stage, src + (size_t)c * a.b_bytes, a.b_bytes, bar(B_FULL2 + stage));
          } else {
            // the leader arms ITS barrier with both halves' bytes; every CTA loads its own 16 KB half with a 2-SM TMA
            // whose completion is credited to the leader's barrier
            const uint32_t bar_leader = bar(B_FULL2 + stage) & 0xFEFFFFFFu;
            if (rank == 0) mbar_expect_tx(bar(B_FULL2 + stage), a.b_bytes);
            tma_load_img_2sm(sB + stage * b_stage, &a.tmap, (ntile * per_tile + c) * 2 + (int)rank, bar_leader);
          }
        }
        __syncwarp();
      }
    }
    if (a.prof && (tid & 31) == 0) a.prof[blockIdx.x * 8 + 7] = pacc[0];
  } else if (warp == 9) {
    // ===== MMA issuer (the leader CTA issues for the pair) ===========================================
    if (rank == 0) {
    const uint32_t idesc = make_idesc(std::is_same<T, __nv_bfloat16>::value ? 1 : 0, nt, 128 * CG);
    // A: SBO = patch row pitch, LBO = panel stride.  B: SBO = 128 B, LBO = rows held by a CTA * 16 B.
    const uint32_t b_rows = (uint32_t)nt / CG;
    const uint32_t a_hi = ((uint32_t)(a.PW * 16) >> 4) | (1u << 14);
    const uint32_t b_hi = (128u >> 4) | (1u << 14);
    const uint32_t a_lbo = (panel_stride >> 4) << 16, b_lbo = b_rows << 16;
    const uint32_t b_step = 2u * b_rows;               // two weight panels per K=16 step
    const uint32_t a_step = 2u * (panel_stride >> 4);  // two patch panels per K=16 step
    const uint32_t b_tap = 8u * b_rows;                // 8 panels * b_rows * 16 B per tap, in 16-byte units
    const uint32_t kw = (uint32_t)a.KW, wrap = (uint32_t)(a.PW - a.KW);   // next filter row: + PW - KW patch columns
    uint32_t ita = 0, itb = 0, tl = 0;
    long long pacc[4] = {0, 0, 0, 0};
    const long long t_start = clock64();
    for (int t = unit0; t < n_tiles; t += n_units, ++tl) {
      const uint32_t buf = tl & 1;
      CPROF(0, mbar_wait_cluster<CG>(bar(D_EMPTY2 + buf), ((tl >> 1) & 1) ^ 1, a.error_flag));
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + buf * 256;
      uint32_t acc = 0;
      for (int cc = 0; cc < a.n_cc; ++cc, ++ita) {
        const uint32_t ast = ita % kAStages, aph = (ita / kAStages) & 1;
        CPROF(1, mbar_wait_cluster<CG>(bar(A_FULL + ast), aph, a.error_flag));
        const int np = min(8, a.cpp - cc * 8);
        const int ks_n = (np + 1) >> 1;
        const uint32_t patch = (sA + ast * a.a_bytes) >> 4;
        // running tap position inside the filter: a_off = ky * PW + kx in 16-byte units (no division per tap)
        uint32_t kx = 0, a_off = 0;
        int taps_left = n_taps;
        for (int ws = 0; ws < a.n_wst; ++ws, ++itb) {
          const uint32_t stage = itb % NB, ph = (itb / NB) & 1;
          CPROF(2, mbar_wait_cluster<CG>(bar(B_FULL2 + stage), ph, a.error_flag));
          tc_fence_after();
          const int tpn = min(a.tps, taps_left);
          const uint32_t b_lo0 = ((sB + stage * b_stage) >> 4) | b_lbo;
          if (elect_one()) {
            const uint32_t a0 = (patch + a_off) | a_lbo;
            if (ks_n == 4) issue_taps<CG, 4>(d_tmem, a0, a_hi, a_step, b_lo0, b_hi, b_step, b_tap, idesc, tpn, kx, kw, wrap, acc);
            else if (ks_n == 1) issue_taps<CG, 1>(d_tmem, a0, a_hi, a_step, b_lo0, b_hi, b_step, b_tap, idesc, tpn, kx, kw, wrap, acc);
            else if (ks_n == 2) issue_taps<CG, 2>(d_tmem, a0, a_hi, a_step, b_lo0, b_hi, b_step, b_tap, idesc, tpn, kx, kw, wrap, acc);
            else issue_taps<CG, 3>(d_tmem, a0, a_hi, a_step, b_lo0, b_hi, b_step, b_tap, idesc, tpn, kx, kw, wrap, acc);
            umma_commit<CG>(bar(B_EMPTY2 + stage));
            if (ws == a.n_wst - 1) {
              umma_commit<CG>(bar(A_EMPTY + ast));
              if (cc == a.n_cc - 1) umma_commit<CG>(bar(D_FULL2 + buf));
            }
          }
          __syncwarp();
          // every lane advances the tap position (the elected lane may change between stages)
          for (int tt = 0; tt < tpn; ++tt) {
            ++a_off;
            if (++kx == kw) kx = 0, a_off += wrap;
          }
          taps_left -= tpn;
          acc = 1;
        }
      }
    }
    if (a.prof && (tid & 31) == 0) {
      for (int i = 0; i < 3; ++i) a.prof[blockIdx.x * 8 + i] = pacc[i];
      a.prof[blockIdx.x * 8 + 3] = clock64() - t_start;
    }
    }
  } else if (warp < 4) {
    // ===== epilogue (thread = pixel of the 16x8 patch) ===============================================
    const int r = tid;
    uint32_t tl = 0;
    long long pacc[4] = {0, 0, 0, 0};
    const long long t_start = clock64();
    for (int t = unit0; t < n_tiles; t += n_units, ++tl) {
      const uint32_t buf = tl & 1;
      const int mt = CG * (t / a.n_ntiles) + (int)rank, n0 = (t % a.n_ntiles) * nt;
      const bool mt_ok = mt < n_mtiles;
      const int b = mt_ok ? mt / tiles_img : 0, ti = mt_ok ? mt % tiles_img : 0;
      const int y = (ti / a.tiles_x) * kTH + (r >> 3), x = (ti % a.tiles_x) * kTW + (r & 7);
      const bool valid = mt_ok && y < a.H && x < a.W;
      const int64_t m = ((int64_t)b * a.H + y) * a.W + x;
      CPROF(0, mbar_wait(bar(D_FULL2 + buf), (tl >> 1) & 1, a.error_flag));
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16) + buf * 256;
      const int64_t plane = (int64_t)a.H * a.W;
      const int64_t nchw_base = ((int64_t)b * a.nchw_C * a.H + y) * a.W + x;
      uint32_t v0[32], v1[32];
      tmem_ld32(t_row, v0);
#pragma unroll 1
      for (int cb = 0; cb < nt / 32; cb += 2) {  // nt is a multiple of 64: two blocks per trip
        tmem_ld_wait(v0);
        tmem_ld32(t_row + (cb + 1) * 32, v1);
        epi_store<T>(a, v0, m, valid, n0 + cb * 32, nchw_base, plane);
        tmem_ld_wait(v1);
        if (cb + 2 < nt / 32) tmem_ld32(t_row + (cb + 2) * 32, v0);
        epi_store<T>(a, v1, m, valid, n0 + (cb + 1) * 32, nchw_base, plane);
      }
      tc_fence_before();
      arrive_leader<CG>(bar(D_EMPTY2 + buf));
    }
    if (a.prof && tid == 0) a.prof[blockIdx.x * 8 + 4] = pacc[0], a.prof[blockIdx.x * 8 + 5] = clock64() - t_start;
  }
  tc_fence_before();
  if (CG == 2) cluster_sync(); else __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc<CG>(tmem_base, 512);
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 1) k_conv_tc(const __grid_constant__ ConvArgs a) {
  conv_tc_body<T, 1>(a);
}

template <typename T>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) k_conv_tc2(const __grid_constant__ ConvArgs a) {
  conv_tc_body<T, 2>(a);
}

}  // namespace conv
}  // namespace dfb

// ------------------------------------------------------------------------------------------
// host: handle, weight packing, launch
// ------------------------------------------------------------------------------------------
struct DfbConv {
  int Cin, Cin_pad, Cout, KH, KW, pad, nt, n_ntiles, n_cc, cpp, tps, n_wst;
  int fmt = 0;     // 0: fp16 operands, 1: bf16 operands
  int dgrad = 0;   // 1: data-gradient convolution of the layer (Cin0, Cout0)
  int Cin0 = 0, Cout0 = 0;
  int nchw_C = 0;  // channels of the fp32 NCHW output (real output channels)
  size_t wimg_bytes = 0;
  uint8_t* wimg = nullptr;
  uint8_t* wimg2 = nullptr;  // the same filter in the cta_group::2 layout: every stage split into the two CTAs' row halves
  CUtensorMap tmap2;         // tensor map over wimg2 (encoded on first use; the image is re-packed in place, never moved)
  bool tmap2_valid = false;
  float* bias = nullptr;
  int num_sms = 0;
};

using namespace dfb;

extern "C" void dfb_conv_destroy(DfbConv* c);

namespace dfb {
namespace conv {

struct PackArgs {
  const float* w;   // [Cout0, Cin0, KH, KW] fp32 (device)
  const float* b;   // [Cout0] or null
  const float* sc;  // [Cout0] or null (eval-mode BatchNorm scale)
  const float* sh;  // [Cout0] or null
  uint16_t* img;
  float* bias;      // [Cout] folded bias of this launch
  int Cin0, Cout0, KH, KW, Cin, Cout, nt, n_ntiles, n_cc, tps, n_wst, fmt, dgrad;
  int cg;           // 1: [..][tap in stage][8 panels][nt rows][8]; 2: [..][row half][tap in stage][8 panels][nt/2 rows][8]
  int64_t total;    // elements of the image
};

// weight image: [n tile][channel slice][weight stage][tap in stage][8 panels][nt rows][8 elements]; one thread = the 8
// elements (input channels) of one row for every tap = KH * KW 16-byte stores.  Batched: a training loop re-packs every convolution of the
// pose regressor after each optimizer step (51 images), which as one launch per image cost ~1 ms of the 20 ms step.
constexpr int kMaxPack = 64;
struct PackBatchArgs {
  int n;
  int blk0[kMaxPack + 1];   // first block of every entry
  PackArgs e[kMaxPack];
};

__global__ void __launch_bounds__(256) k_pack_conv_weights(const __grid_constant__ PackBatchArgs b) {
  int lo = 0, hi = b.n;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if ((int)blockIdx.x >= b.blk0[mid]) lo = mid; else hi = mid;
  }
  const PackArgs& a = b.e[lo];
  // one thread = the 8 input channels of one image row FOR EVERY TAP: its source is one contiguous run of 8 * KH * KW floats
  // (forward) or eight runs of KH * KW floats next to its neighbours' (data gradient), and neighbouring threads store
  // neighbouring 16-byte rows.  (With one thread per row and tap every 4-byte read of a warp hit its own 32-byte sector,
  // 18 KB apart: 0.25 ms for the four images of the 13 encoder layers, L2-bound.)
  const int64_t i = (int64_t)((int)blockIdx.x - b.blk0[lo]) * blockDim.x + threadIdx.x;
  if (i < a.Cout) {
    float bf = 0.f;
    if (!a.dgrad) {  // y = scale * (conv + bias) + shift
      const float scn = a.sc ? a.sc[i] : 1.f;
      bf = scn * (a.b ? a.b[i] : 0.f) + (a.sh ? a.sh[i] : 0.f);
    }
    a.bias[i] = bf;
  }
  const int per_row = a.n_wst * a.tps;                 // rows of the image this thread writes (taps incl. stage padding)
  if (i * 8 * per_row >= a.total) return;
  int64_t r = i;
  const int rows = a.nt / a.cg;
  const int rr0 = (int)(r % rows); r /= rows;
  const int pp = (int)(r % 8); r /= 8;
  int half = 0;
  if (a.cg == 2) { half = (int)(r % 2); r /= 2; }
  const int cc = (int)(r % a.n_cc); r /= a.n_cc;
  const int t = (int)r;
  const int rr = rr0 + half * rows;
  const int ci0 = (cc * 8 + pp) * 8, n = t * a.nt + rr;
  const int64_t kk = (int64_t)a.KH * a.KW;
  const float scn = (!a.dgrad && a.sc) ? a.sc[n] : 1.f;
  float scd[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) scd[e] = (a.dgrad && a.sc && ci0 + e < a.Cout0) ? a.sc[ci0 + e] : 1.f;
  for (int ws = 0; ws < a.n_wst; ++ws)
    for (int tis = 0; tis < a.tps; ++tis) {
      const int tp = ws * a.tps + tis;
      float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (tp < a.KH * a.KW) {
        const int ky = tp / a.KW, kx = tp % a.KW;
        if (!a.dgrad) {
          const float* src = a.w + ((int64_t)n * a.Cin0 + ci0) * kk + ky * a.KW + kx;
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (ci0 + e < a.Cin0) v[e] = scn * src[e * kk];
        } else if (n < a.Cin0) {
          // data gradient: n = input channel of the layer, ci = its output channel, filter flipped
          const float* src = a.w + ((int64_t)ci0 * a.Cin0 + n) * kk + (a.KH - 1 - ky) * a.KW + (a.KW - 1 - kx);
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (ci0 + e < a.Cout0) v[e] = scd[e] * src[(int64_t)e * a.Cin0 * kk];
        }
      }
      uint32_t pk[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t lo16 = a.fmt ? __bfloat16_as_ushort(__float2bfloat16_rn(v[2 * q])) : __half_as_ushort(__float2half_rn(v[2 * q]));
        const uint32_t hi16 = a.fmt ? __bfloat16_as_ushort(__float2bfloat16_rn(v[2 * q + 1])) : __half_as_ushort(__float2half_rn(v[2 * q + 1]));
        pk[q] = lo16 | (hi16 << 16);
      }
      // image order: [n tile][channel slice][weight stage][row half (cta_group::2)][tap in stage][8 panels][rows][8]
      int64_t o = ((int64_t)t * a.n_cc + cc) * a.n_wst + ws;
      if (a.cg == 2) o = o * 2 + half;
      o = ((o * a.tps + tis) * 8 + pp) * rows + rr0;
      reinterpret_cast<uint4*>(a.img)[o] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
}

// Packing requests collected between pack_batch_begin() and pack_batch_flush() (dfb_dfnet_load_ex) go out as ONE launch
// on the legacy default stream; outside such a bracket every request is launched at once.
struct PackBatch {
  PackBatchArgs args;
  bool open = false;
};
static thread_local PackBatch g_pack;

static int pack_launch(cudaStream_t st) {
  if (g_pack.args.n == 0) return DFB_OK;
  k_pack_conv_weights<<<(unsigned)g_pack.args.blk0[g_pack.args.n], 256, 0, st>>>(g_pack.args);
  g_pack.args.n = 0;
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}
static int pack_push(const PackArgs& a, cudaStream_t st) {
  if (g_pack.args.n == kMaxPack) {
    const int rc = pack_launch(st);
    if (rc) return rc;
  }
  const int64_t groups = std::max<int64_t>(a.total / 8 / ((int64_t)a.n_wst * a.tps), a.Cout);   // one thread per (row, 8 channels)
  const int k = g_pack.args.n++;
  if (k == 0) g_pack.args.blk0[0] = 0;
  g_pack.args.e[k] = a;
  g_pack.args.blk0[k + 1] = g_pack.args.blk0[k] + (int)((groups + 255) / 256);
  return g_pack.open ? DFB_OK : pack_launch(st);
}

}  // namespace conv
}  // namespace dfb

// (Re)pack the filter of an existing handle from fp32 tensors in host or device memory.
int dfb_conv_update_impl(DfbConv* c, const float* weight, const float* bias, const float* bn_scale, const float* bn_shift,
                         void* stream) {
  DFB_REQUIRE(c && weight, DFB_ERR_INVALID, "dfb_conv_update: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t wn = (size_t)c->Cout0 * c->Cin0 * c->KH * c->KW;
  // Device-resident sources (a training loop re-loads the parameters after every optimizer step) are packed in place;
  // only host tensors are staged.  (Staging everything through cudaMallocAsync cost ~60 pool allocations per step, and
  // the pool hands its memory back to the driver at synchronisation points: sporadic 30-500 ms stalls.)
  // (the answer is remembered per pointer: a training loop asks about the same ~100 parameter tensors after every optimizer
  // step, and cudaPointerGetAttributes is 1-2 us a call.  Only "device" answers are kept; CUDA's address range is never
  // handed to pageable host allocations.)
  auto on_device = [](const void* p) {
    if (!p) return true;
    static thread_local const void* known[256] = {};
    const size_t slot = ((uintptr_t)p >> 8) * 0x9E3779B97F4A7C15ull >> 56;
    if (known[slot] == p) return true;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    const bool dev = at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
    if (dev) known[slot] = p;
    return dev;
  };
  float* stage = nullptr;
  const float *dw = weight, *db = bias, *dsc = bn_scale, *dsh = bn_shift;
  if (!(on_device(weight) && on_device(bias) && on_device(bn_scale) && on_device(bn_shift))) {
    const size_t need = (wn + 3 * (size_t)c->Cout0) * 4;
    DFB_CHECK_CUDA(cudaMallocAsync((void**)&stage, need, st));
    float* sb = stage + wn, *ssc = sb + c->Cout0, *ssh = ssc + c->Cout0;
    DFB_CHECK_CUDA(cudaMemcpyAsync(stage, weight, wn * 4, cudaMemcpyDefault, st));
    if (bias) DFB_CHECK_CUDA(cudaMemcpyAsync(sb, bias, c->Cout0 * 4, cudaMemcpyDefault, st));
    if (bn_scale) DFB_CHECK_CUDA(cudaMemcpyAsync(ssc, bn_scale, c->Cout0 * 4, cudaMemcpyDefault, st));
    if (bn_shift) DFB_CHECK_CUDA(cudaMemcpyAsync(ssh, bn_shift, c->Cout0 * 4, cudaMemcpyDefault, st));
    dw = stage, db = sb, dsc = ssc, dsh = ssh;
  }
  conv::PackArgs a = {};
  a.w = dw, a.b = bias ? db : nullptr, a.sc = bn_scale ? dsc : nullptr, a.sh = bn_shift ? dsh : nullptr;
  a.img = (uint16_t*)c->wimg, a.bias = c->bias;
  a.Cin0 = c->Cin0, a.Cout0 = c->Cout0, a.KH = c->KH, a.KW = c->KW, a.Cin = c->Cin, a.Cout = c->Cout, a.nt = c->nt;
  a.n_ntiles = c->n_ntiles, a.n_cc = c->n_cc, a.tps = c->tps, a.n_wst = c->n_wst, a.fmt = c->fmt, a.dgrad = c->dgrad;
  a.total = (int64_t)c->wimg_bytes / 2;
  a.cg = 1;
  // staged (host) sources are released right below, so their request cannot wait for a batched launch
  const bool was_open = conv::g_pack.open;
  if (stage) {
    const int rc0 = conv::pack_launch(st);
    if (rc0) return rc0;
    conv::g_pack.open = false;
  }
  int rc = conv::pack_push(a, st);
  if (!rc && c->wimg2) {
    a.img = (uint16_t*)c->wimg2, a.cg = 2;
    rc = conv::pack_push(a, st);
  }
  conv::g_pack.open = was_open;
  if (stage) DFB_CHECK_CUDA(cudaFreeAsync(stage, st));
  return rc;
}

// Bracket for callers that (re)pack many convolutions in a row on the legacy default stream (dfb_dfnet_load_ex).
void dfb_conv_pack_batch_begin() { conv::g_pack.open = true; }
int dfb_conv_pack_batch_flush(bool discard, void* stream) {
  conv::g_pack.open = false;
  if (discard) conv::g_pack.args.n = 0;   // a failed load may have destroyed handles the pending requests point to
  return conv::pack_launch((cudaStream_t)stream);
}
// ABI form of the bracket (nerf_train.py re-loads ~60 one-by-one convolutions after every optimizer step): between
// dfb_conv_pack_begin and dfb_conv_pack_end every dfb_conv_update of device-resident tensors is queued and goes out as one
// launch on `stream`; the sources must stay unchanged until then.
extern "C" int dfb_conv_pack_begin() {
  dfb_conv_pack_batch_begin();
  return DFB_OK;
}
extern "C" int dfb_conv_pack_end(void* stream) { return dfb_conv_pack_batch_flush(false, stream); }
// cta_group used by the convolution kernel.  2 (default): CTA pairs work on two neighbouring M tiles with one weight stream
// and one MMA instruction for both SMs; DFB_CONV_CTA_GROUP=1 selects the 1-CTA kernel (also used for single-tile launches).
// Read per call so that tests can exercise both variants in one process; both weight images are always packed.
static int conv_cg_env() {
  const char* e = getenv("DFB_CONV_CTA_GROUP");
  return (e && e[0] == '1') ? 1 : 2;
}
// rounds of the persistent grid for a B x H x W launch: ceil(schedule units / resident units), a unit being one CTA
// (1-CTA kernel) or a CTA pair working on two M tiles (cta_group::2)
int64_t dfb_conv_rounds(const DfbConv* c, int B, int H, int W) {
  const int64_t n_mtiles = (int64_t)((W + conv::kTW - 1) / conv::kTW) * ((H + conv::kTH - 1) / conv::kTH) * B;
  const int cg = (c->wimg2 && n_mtiles >= 2) ? conv_cg_env() : 1;
  const int64_t n_tiles = ((n_mtiles + cg - 1) / cg) * c->n_ntiles, units = c->num_sms / cg;
  return (n_tiles + units - 1) / units;
}

// fmt: 0 fp16 / 1 bf16 operands.  dgrad != 0 builds the DATA-GRADIENT convolution of the layer described by
// (Cin0, Cout0, weight [Cout0,Cin0,KH,KW], bn_scale): gI[b,y,x,c] = sum_{n,ky,kx} gO[b,y+pad-ky,x+pad-kx,n] * scale[n] * w[n,c,ky,kx],
// i.e. the same kernel run on the transposed, spatially flipped filter (its "Cin" is Cout0, its "Cout" is Cin0
// rounded up to 64, no bias).
// Tiles handle c launches for B x H x W pixels.  The DFNet forward keeps a second packing of its 512-channel
// layers with 128-wide output-channel tiles and uses it when the 256-wide tiling would leave more than half of the
// SMs idle (conv5_x at 480x640: 40 tiles on 148 SMs -> 80 tiles; measured 52 -> 42 us per layer).  Wider layers with
// enough tiles stay on 256: the narrower tile reloads the input patch per output-channel tile and measured 15-25 %
// slower there despite the better wave quantisation.
int64_t dfb_conv_tiles(const DfbConv* c, int B, int H, int W) {
  return (int64_t)((W + conv::kTW - 1) / conv::kTW) * ((H + conv::kTH - 1) / conv::kTH) * B * c->n_ntiles;
}
int dfb_conv_num_sms(const DfbConv* c) { return c->num_sms; }

int dfb_conv_create_impl(int Cin0, int Cout0, int KH, int KW, const float* weight, const float* bias, const float* bn_scale,
                         const float* bn_shift, int fmt, int dgrad, DfbConv** out, int nt_force) {
  const int Cin = dgrad ? Cout0 : Cin0, Cout = dgrad ? round_up(Cin0, 64) : Cout0;
  DFB_REQUIRE(weight && out && Cin >= 1 && Cout >= 64 && Cout % 64 == 0, DFB_ERR_INVALID,
              "dfb_conv_create: Cout must be a positive multiple of 64");
  DFB_REQUIRE(KH == KW && (KH == 1 || KH == 3 || KH == 5), DFB_ERR_UNSUPPORTED, "kernel size must be 1, 3 or 5");
  int dev = 0, major = 0, sms = 0;
  DFB_CHECK_CUDA(cudaGetDevice(&dev));
  DFB_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  DFB_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  DFB_REQUIRE(major == 10, DFB_ERR_UNSUPPORTED, "sm_100a code only");
  DfbConv* c = new DfbConv();
  c->fmt = fmt, c->dgrad = dgrad, c->Cin0 = Cin0, c->Cout0 = Cout0, c->nchw_C = dgrad ? Cin0 : Cout0;
  c->Cin = Cin, c->Cin_pad = round_up(Cin, 8), c->Cout = Cout, c->KH = KH, c->KW = KW, c->pad = KH / 2;
  c->nt = Cout % 256 == 0 ? 256 : (Cout % 128 == 0 ? 128 : 64);
  if (nt_force && Cout % nt_force == 0 && nt_force <= c->nt) c->nt = nt_force;  // narrower output-channel tiles (see dfb_conv_tiles)
  c->n_ntiles = Cout / c->nt;
  c->cpp = c->Cin_pad / 8;
  c->n_cc = (c->cpp + 7) / 8;
  c->tps = 256 / c->nt;
  c->n_wst = (KH * KW + c->tps - 1) / c->tps;
  c->num_sms = sms;
  c->wimg_bytes = (size_t)c->n_ntiles * c->n_cc * c->n_wst * c->tps * c->nt * 16 * 8;
  if (cudaMalloc(&c->wimg, c->wimg_bytes) != cudaSuccess || cudaMalloc(&c->wimg2, c->wimg_bytes) != cudaSuccess ||
      cudaMalloc(&c->bias, Cout * 4) != cudaSuccess) {
    dfb_conv_destroy(c);
    DFB_REQUIRE(false, DFB_ERR_CUDA, "dfb_conv_create: out of device memory");
  }
  const int rc = dfb_conv_update_impl(c, weight, bias, bn_scale, bn_shift, nullptr);
  if (rc) { dfb_conv_destroy(c); return rc; }
  DFB_CHECK_CUDA(cudaStreamSynchronize(nullptr));  // the source tensors may be released by the caller on return
  *out = c;
  return DFB_OK;
}

extern "C" int dfb_conv_create(int Cin, int Cout, int KH, int KW, const float* weight, const float* bias,
                               const float* bn_scale, const float* bn_shift, DfbConv** out) {
  return dfb_conv_create_impl(Cin, Cout, KH, KW, weight, bias, bn_scale, bn_shift, 0, 0, out, 0);
}

extern "C" int dfb_conv_update(DfbConv* c, const float* weight, const float* bias, const float* bn_scale, const float* bn_shift,
                               void* stream) {
  return dfb_conv_update_impl(c, weight, bias, bn_scale, bn_shift, stream);
}

// n handles in one call, bracketed as one packing launch (bias[i] nullable; no BatchNorm folding): the per-call cost of the
// foreign-function interface is what a training loop with ~60 small layers pays after every optimizer step.
extern "C" int dfb_conv_update_many(DfbConv* const* convs, const float* const* weights, const float* const* biases, int n, void* stream) {
  DFB_REQUIRE(n >= 0 && (n == 0 || (convs && weights && biases)), DFB_ERR_INVALID, "dfb_conv_update_many: bad arguments");
  int rc = dfb_conv_pack_begin();
  if (rc) return rc;
  for (int i = 0; i < n && !rc; ++i) rc = dfb_conv_update_impl(convs[i], weights[i], biases[i], nullptr, nullptr, stream);
  const int rc2 = dfb_conv_pack_end(stream);
  return rc ? rc : rc2;
}

extern "C" void dfb_conv_destroy(DfbConv* c) {
  if (!c) return;
  if (c->wimg) cudaFree(c->wimg);
  if (c->wimg2) cudaFree(c->wimg2);
  if (c->bias) cudaFree(c->bias);
  delete c;
}

int dfb_conv_run(DfbConv* c, const void* in_nhwc16, int B, int H, int W, int relu, void* out_nhwc16, void* tap_nhwc16,
                 float* out_nchw32, const void* mask_nhwc16, const void* addend_nhwc16, void* stream, void* out_bf16);

extern "C" int dfb_conv_fwd(DfbConv* c, const void* in_nhwc16, int B, int H, int W, int relu, void* out_nhwc16,
                            void* tap_nhwc16, float* out_nchw32, void* stream) {
  return dfb_conv_run(c, in_nhwc16, B, H, W, relu, out_nhwc16, tap_nhwc16, out_nchw32, nullptr, nullptr, stream, nullptr);
}

// 5-D tensor map over the NHWC 16-bit input for the patch loads of k_conv_tc (see ConvArgs::tmap_in).
// Encoding a tensor map costs 2-3 us on the host - a third of an asynchronous launch through this ABI - and a training
// step launches the same convolutions on the same (persistent) buffers every time: the maps are kept in a small
// direct-mapped cache keyed by everything that determines them (a map depends on nothing else, so entries never go stale).
namespace {
struct TmapKey { const void* in; int B, H, W, Cpad, PH, PW, np; };
struct TmapEntry { TmapKey k; CUtensorMap m; bool valid; };
thread_local TmapEntry g_tmap_cache[64];
}  // namespace

static int make_patch_tmap_uncached(const void* in, int B, int H, int W, int Cpad, int PH, int PW, CUtensorMap* out, int npanels);

int dfb::make_patch_tmap(const void* in, int B, int H, int W, int Cpad, int PH, int PW, CUtensorMap* out, int npanels) {
  const TmapKey k = {in, B, H, W, Cpad, PH, PW, npanels};
  uint64_t hsh = (uint64_t)(uintptr_t)in * 0x9E3779B97F4A7C15ull;
  hsh ^= ((uint64_t)B << 40) ^ ((uint64_t)H << 28) ^ ((uint64_t)W << 16) ^ ((uint64_t)Cpad << 6) ^ ((uint64_t)PH << 3) ^ (uint64_t)PW ^ ((uint64_t)npanels << 50);
  TmapEntry& e = g_tmap_cache[(hsh >> 20) & 63];
  if (e.valid && memcmp(&e.k, &k, sizeof(k)) == 0) {
    *out = e.m;
    return DFB_OK;
  }
  const int rc = make_patch_tmap_uncached(in, B, H, W, Cpad, PH, PW, out, npanels);
  if (rc) return rc;
  e.k = k, e.m = *out, e.valid = true;
  return DFB_OK;
}

static int make_patch_tmap_uncached(const void* in, int B, int H, int W, int Cpad, int PH, int PW, CUtensorMap* out, int npanels) {
  static PFN_cuTensorMapEncodeTiled encode = nullptr;
  if (npanels == -64) {
    // pixel-major 128-byte-swizzled box: NHWC as [B][H][W][C], box = PH x PW pixels x 64 channels (one 128-byte row per pixel)
    if (!encode) {
      void* fn = nullptr;
      cudaDriverEntryPointQueryResult qres;
      DFB_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
      DFB_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, DFB_ERR_CUDA, "cuTensorMapEncodeTiled not available");
      encode = (PFN_cuTensorMapEncodeTiled)fn;
    }
    DFB_REQUIRE(((uintptr_t)in & 15) == 0 && Cpad % 8 == 0, DFB_ERR_INVALID, "conv input must be 16-byte aligned NHWC with C % 8 == 0");
    const cuuint64_t dims[4] = {(cuuint64_t)Cpad, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)Cpad * 2, (cuuint64_t)W * Cpad * 2, (cuuint64_t)H * W * Cpad * 2};
    const cuuint32_t box[4] = {64, (cuuint32_t)PW, (cuuint32_t)PH, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, const_cast<void*>(in), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DFB_REQUIRE(r == CUDA_SUCCESS, DFB_ERR_CUDA, "cuTensorMapEncodeTiled (pixel-major, %dx%dx%dx%d) failed (%d)", B, H, W, Cpad, (int)r);
    return DFB_OK;
  }
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    DFB_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    DFB_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, DFB_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    encode = (PFN_cuTensorMapEncodeTiled)fn;
  }
  DFB_REQUIRE(((uintptr_t)in & 15) == 0 && Cpad % 8 == 0, DFB_ERR_INVALID, "conv input must be 16-byte aligned NHWC with C % 8 == 0");
  const cuuint64_t dims[5] = {8, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(Cpad / 8), (cuuint64_t)B};
  const cuuint64_t strides[4] = {(cuuint64_t)Cpad * 2, (cuuint64_t)W * Cpad * 2, 16, (cuuint64_t)H * W * Cpad * 2};
  const cuuint32_t box[5] = {8, (cuuint32_t)PW, (cuuint32_t)PH, (cuuint32_t)npanels, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, 5, const_cast<void*>(in), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DFB_REQUIRE(r == CUDA_SUCCESS, DFB_ERR_CUDA, "cuTensorMapEncodeTiled (conv input, %dx%dx%dx%d) failed (%d)", B, H, W, Cpad, (int)r);
  return DFB_OK;
}

static unsigned long long* g_conv_prof = nullptr;  // dfb_debug_conv_prof: counters of the LAST conv launch
static int g_conv_prof_grid = 0;

// Debug seam (tools/conv_prof.py): on != 0 makes every following conv launch record per-CTA cycle counters
// ([cta][8]: issuer waits D_EMPTY, A_FULL, B_FULL, issuer total, epilogue wait D_FULL, epilogue total, loader wait
// A_EMPTY, producer wait B_EMPTY); out_host (nullable) receives the counters of the last launch, *grid its CTA count.
extern "C" int dfb_debug_conv_prof(int on, unsigned long long* out_host, int max_cta, int* grid) {
  if (on && !g_conv_prof) DFB_CHECK_CUDA(cudaMalloc(&g_conv_prof, 512 * 8 * sizeof(unsigned long long)));
  if (!on && g_conv_prof && !out_host) { cudaFree(g_conv_prof); g_conv_prof = nullptr; }
  if (out_host && g_conv_prof) {
    DFB_CHECK_CUDA(cudaDeviceSynchronize());
    const int n = std::min(max_cta, g_conv_prof_grid);
    DFB_CHECK_CUDA(cudaMemcpy(out_host, g_conv_prof, (size_t)n * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (grid) *grid = n;
  }
  return DFB_OK;
}

int dfb_conv_run(DfbConv* c, const void* in_nhwc16, int B, int H, int W, int relu, void* out_nhwc16, void* tap_nhwc16,
                 float* out_nchw32, const void* mask_nhwc16, const void* addend_nhwc16, void* stream, void* out_bf16) {
  DFB_REQUIRE(c && in_nhwc16 && (out_nhwc16 || tap_nhwc16 || out_nchw32), DFB_ERR_INVALID, "dfb_conv_fwd: null argument");
  DFB_REQUIRE(!out_bf16 || out_nhwc16, DFB_ERR_INVALID, "dfb_conv_fwd: the bf16 copy accompanies the 16-bit output");
  DFB_REQUIRE(B >= 1 && H >= 1 && W >= 1, DFB_ERR_INVALID, "bad image size");
  int* error_flag = nullptr;
  {
    const int rc = device_error_flag(&error_flag);
    if (rc) return rc;
  }
  conv::ConvArgs a = {};
  a.in = (const __half*)in_nhwc16, a.wimg = c->wimg, a.bias = c->bias;
  a.out = (__half*)out_nhwc16, a.tap = (__half*)tap_nhwc16, a.out_nchw = out_nchw32, a.out_bf16 = (uint16_t*)out_bf16;
  a.mask = (const uint16_t*)mask_nhwc16, a.addend = (const uint16_t*)addend_nhwc16, a.nchw_C = c->nchw_C;
  a.B = B, a.H = H, a.W = W, a.Cin = c->Cin_pad, a.Cout = c->Cout, a.KH = c->KH, a.KW = c->KW, a.pad = c->pad, a.relu = relu;
  a.nt = c->nt, a.n_ntiles = c->n_ntiles, a.n_cc = c->n_cc, a.cpp = c->cpp;
  a.tiles_x = (W + conv::kTW - 1) / conv::kTW, a.tiles_y = (H + conv::kTH - 1) / conv::kTH;
  a.PH = conv::kTH + 2 * c->pad, a.PW = conv::kTW + 2 * c->pad;
  a.a_bytes = (uint32_t)a.PH * a.PW * 16u * 8u;
  a.b_bytes = (uint32_t)c->nt * 16u * 8u * (uint32_t)c->tps;
  a.tps = c->tps, a.n_wst = c->n_wst;
  a.error_flag = error_flag;
  const int64_t n_mtiles = (int64_t)a.tiles_x * a.tiles_y * B;
  DFB_REQUIRE(n_mtiles * a.n_ntiles < (1ll << 30), DFB_ERR_INVALID, "image too large");
  // cta_group::2 (CTA pairs share every weight stage, one MMA instruction feeds both SMs) is the default: with the lean
  // issue loop and CTA-scope barrier arrivals it is faster on every layer class (r02: conv1_2 79 -> 70 us, 5x5 level-0
  // head 252 -> 210 us, conv3_2 56 -> 54 us at 480x640, batch 2); earlier in the round the two kernels had measured equal.
  const int cg = (c->wimg2 && n_mtiles >= 2) ? conv_cg_env() : 1;
  const int64_t n_tiles = ((n_mtiles + cg - 1) / cg) * a.n_ntiles;
  const int grid = cg * (int)std::min<int64_t>(n_tiles, c->num_sms / cg);
  const int nb = cg == 2 ? conv::kBStages2 : conv::kBStages;
  const size_t smem = (size_t)conv::kAStages * a.a_bytes + (size_t)nb * (a.b_bytes / cg) + 256;
  DFB_REQUIRE(smem <= 232448, DFB_ERR_UNSUPPORTED, "shared memory budget exceeded");
  if (cg == 2) {
    a.wimg = c->wimg2;
    if (!c->tmap2_valid) {
      const int rc = make_weight_tmap(c->wimg2, c->wimg_bytes, &c->tmap2);
      if (rc) return rc;
      c->tmap2_valid = true;
    }
    a.tmap = c->tmap2;
  }
  {
    const int rc = make_patch_tmap(in_nhwc16, B, H, W, c->Cin_pad, a.PH, a.PW, &a.tmap_in);
    if (rc) return rc;
  }
  if (g_conv_prof) {
    DFB_CHECK_CUDA(cudaMemsetAsync(g_conv_prof, 0, 512 * 8 * sizeof(unsigned long long), (cudaStream_t)stream));
    a.prof = g_conv_prof, g_conv_prof_grid = grid;
  }
  // the four variants share one function-pointer type, so the "attribute already set" flag is per variant, not per lambda
  auto launch = [&](auto kern, int variant) -> int {
    static thread_local uint64_t attr_set[4] = {0, 0, 0, 0};   // bit d: set on device d (the attribute is per device)
    int dev = 0;
    DFB_CHECK_CUDA(cudaGetDevice(&dev));
    if (!((attr_set[variant] >> (dev & 63)) & 1)) {
      DFB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
      attr_set[variant] |= 1ull << (dev & 63);
    }
    DFB_CHECK_CUDA(dfb_launch_pdl(kern, dim3(grid), dim3(conv::kThreads), smem, (cudaStream_t)stream, true, a));
    DFB_LAUNCH_CHECK();
    return DFB_OK;
  };
  if (cg == 2) return c->fmt ? launch(conv::k_conv_tc2<__nv_bfloat16>, 3) : launch(conv::k_conv_tc2<__half>, 2);
  return c->fmt ? launch(conv::k_conv_tc<__nv_bfloat16>, 1) : launch(conv::k_conv_tc<__half>, 0);
}

extern "C" int dfb_conv_create_ex(int Cin, int Cout, int KH, int KW, const float* weight, const float* bias, const float* bn_scale,
                                  const float* bn_shift, int fmt, int dgrad, DfbConv** out) {
  DFB_REQUIRE(fmt == 0 || fmt == 1, DFB_ERR_INVALID, "fmt must be 0 (fp16) or 1 (bf16)");
  return dfb_conv_create_impl(Cin, Cout, KH, KW, weight, bias, bn_scale, bn_shift, fmt, dgrad ? 1 : 0, out, 0);
}

extern "C" int dfb_conv_fwd_ex(DfbConv* c, const void* in_nhwc16, int B, int H, int W, int relu, void* out_nhwc16,
                               void* tap_nhwc16, float* out_nchw32, const void* mask_nhwc16, const void* addend_nhwc16,
                               void* stream) {
  return dfb_conv_run(c, in_nhwc16, B, H, W, relu, out_nhwc16, tap_nhwc16, out_nchw32, mask_nhwc16, addend_nhwc16, stream, nullptr);
}

// dfb_conv_fwd_ex with a bf16 copy of the 16-bit output (the operand type of dfb_conv_wgrad), written by the same epilogue
extern "C" int dfb_conv_fwd_ex2(DfbConv* c, const void* in_nhwc16, int B, int H, int W, int relu, void* out_nhwc16,
                                void* tap_nhwc16, float* out_nchw32, const void* mask_nhwc16, const void* addend_nhwc16,
                                void* out_bf16, void* stream) {
  return dfb_conv_run(c, in_nhwc16, B, H, W, relu, out_nhwc16, tap_nhwc16, out_nchw32, mask_nhwc16, addend_nhwc16, stream, out_bf16);
}
