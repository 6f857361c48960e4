// tcgen05 backward of the fine NeRF-W MLP w.r.t. its inputs (8x256 network): what train.py needs from the
// renderer (reference feature/direct_feature_matching.py:342-378 -> autograd through models/nerfw.py:297-354
// and models/rendering.py:287,305; the NeRF weights are frozen and the depths detached, rendering.py:302).
//
// Same machine as the forward kernel (mlp_tc.cu): persistent CTAs, two 128-sample tiles ("slots") that
// alternate between the tensor pipe and the epilogue warps, weights streamed as pre-packed 16 KB chunks
// through a 4-stage shared-memory ring, accumulators in TMEM.  One pass over a tile runs 29 program steps:
//
//   0..11   forward recompute (trunk 0..7, dir|transient.0 with xyz_encoding_final folded in, transient 2,4,6);
//           the epilogues keep ONLY the ReLU masks, one bit per activation, in a per-CTA scratch that the same
//           thread reads back later (L2-resident, 32 B per row and layer)
//   11      + head derivatives d heads (9 per sample) from the saved forward outputs `raw` and d raw, written as a
//           16-wide A operand ("ghA", in the free positional-encoding panels)
//   12      g_T3 = (d heads) heads^T as ONE K=16 MMA                                      (short step)
//   13,14   transient branch:  g_in = (g_out W) . mask, W^T streamed as the B operand
//   15+16   [g_dir | g_t0]: d rgb * rgb head^T (K=16, columns 0..127) and g_T1 W_t0 (columns 128..255) in one
//           accumulator, one 256-column mask epilogue
//   17      d dirPE = g_dir W_dir[:, 256:283]            -> g_samp[:, 3:30]
//   18+19   g_h7 = [g_dir | g_t0] (W_dt W_final)  +  d sigma * w_sigma (K=16, accumulated)
//   20..27  trunk (the skip layer splits into a 64-wide positional-encoding part, parked in the scratch, and
//           the 256-wide hidden part)
//   28      d PE = g_h0 W_0 + skip part, contracted with the encoding's Jacobian -> d pts -> g_samp[:, 0:3]
//
// (The head derivatives used to be fp32 dot products in the epilogue, 5 + 3 + 1 FMAs per hidden unit with
// register-indexed constant-bank operands: 10 000 + 2 000 + 2 800 of the 47 000 epilogue cycles of a pass.)
//
// Gradients are carried as 16-bit MMA operands (same kind as the forward: fp16 or bf16) with fp32
// accumulation.  A mean-reduced loss gives |g| ~ 1e-8, far below fp16's range, so every row (sample) is
// scaled by a power of two chosen from its head derivatives; the chain is linear in g, the scale is undone
// exactly on the two outputs.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <type_traits>

#include <cuda.h>
#include <cudaTypedefs.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace dfb {
namespace tcb {

using namespace dfb::tc;

constexpr int kTileM = 128;
constexpr int kMaxStages = 6;   // ring depth is a launch parameter: 4 (forward recompute) or 5 (saved masks)
constexpr int kChunkBytes = 16384;
constexpr int kPanelBytes = kTileM * 16;
constexpr int kHPanels = 32;
constexpr int kPePanels = 8;
constexpr int kGhPanels = 2;     // saved-mask program: only ghA lives behind the hidden panels
constexpr int kSmemBar = 256;
// shared memory: [slot 0: 32 hidden panels + pe_panels][slot 1: same][ring: n_stages x 16 KB][barriers]
__host__ __device__ constexpr int slot_bytes(int pe_panels) { return (kHPanels + pe_panels) * kPanelBytes; }
__host__ __device__ constexpr int smem_total(int pe_panels, int n_stages) {
  return 2 * slot_bytes(pe_panels) + n_stages * kChunkBytes + kSmemBar;
}
static_assert(smem_total(kPePanels, 4) <= 232448 && smem_total(kGhPanels, 5) <= 232448, "shared memory budget");
constexpr int kThreads = 448;
constexpr int kSteps = 29;
constexpr int kGhPanel = kHPanels;   // ghA: panels 32, 33 of the slot (the positional-encoding area)
constexpr int kMaskLayers = 12;
// per-CTA scratch, in 32-bit words
constexpr int kScrMask = 0;                                   // [slot][layer][8 words][128 rows]
constexpr int kScrSkip = kScrMask + 2 * kMaskLayers * 8 * 128;  // [slot][64][128]  fp32
constexpr int kScrJac = kScrSkip + 2 * 64 * 128;              // [parity][slot][64][128] fp32
constexpr int kScrWords = kScrJac + 2 * 2 * 64 * 128;

enum Bar { W_FULL = 0, W_EMPTY = 6, D_FULL = 12, A_READY = 14, PASS_DONE = 16, PE_READY = 18, PE_FREE = 20, N_BARS = 22 };

// HEADS: the head-derivative part of F_T3 alone (no MMA, no accumulator read): first step when the forward saved its masks.
// SUB: first MMA group of a compound step (accumulator not committed, no epilogue).
enum Kind { F_HID, F_DT, F_T3, B_MASK, B_DIRPE, B_PESKIP, B_PE0, HEADS, SUB };

struct Step {
  int n_chunks;    // weight chunks
  int cbytes;      // bytes per chunk (1-CTA kernel): 16 KB, or the two 8-column panels of a short (K = 16) step
  int img_base;    // cta_group::2 kernel: index of the step's first 16 KB image (chunk c = images img_base + 2c, + 2c + 1)
  int n;           // MMA N
  int a_panel0;    // first A panel
  uint32_t w_off;  // byte offset of the first chunk in the packed image
  int kind;
  int ml;          // mask layer written (forward) / applied (backward)
  int brow;        // row of the packed bias table (forward hidden layers)
  int ncb;         // 32-column blocks the epilogue reads
  int d_col;       // first accumulator column of the MMAs
  int kshort;      // 1: a single K = 16 MMA on a one-chunk B operand
  int acc;         // 1: accumulate onto what the previous (chained) step left in the accumulator
  int chain;       // 1: no commit, the next step continues the same accumulator
};

struct BtArgs {
  // cta_group::2 only: 3-D tensor map over the pair-layout weight image ([images][64 rows][256 B] = 16 KB boxes), see mlp_tc.cu
  alignas(64) CUtensorMap tmap;
  Step steps[kSteps];
  const void* wimg;
  const float* rayrec;   // [n_rays,12]
  const float* z;        // [n_rays,S]
  const float* raybias;  // [n_rays,256]
  const float* raw;      // [P,9] forward outputs
  const float* g_raw;    // [P,9]
  float* g_samp;         // [P,32]: d pts (3) | d dirPE (27) | pad
  uint32_t* scratch;     // [gridDim.x][kScrWords]
  // ReLU masks saved by the training forward (mlp_tc.cu, FULL == 2), [tile][12][8][128]; when set, the program is the
  // 15 backward steps only (no forward recompute) and the masks are read from here instead of the scratch
  const uint32_t* saved_masks;
  int n_steps;           // 29 (recompute) or 18
  int n_stages;          // depth of the weight ring
  int slot_bytes;        // shared-memory stride of a slot
  int pe_free_step;      // last step whose MMAs read the positional-encoding panels
  int S;
  int64_t P, n_pass;
  int* error_flag;
  // debug seam (dfb_debug_bwd_masks): ReLU masks per sample, [P][12 layers][8 words] in this kernel's bit layout;
  // mask_out receives the masks of the forward recompute, mask_in (if set) replaces them before they are used
  const uint32_t* mask_in;
  uint32_t* mask_out;
  unsigned long long* prof;  // optional [gridDim.x][64] cycle counters (DFB_TC_PROF builds, tools/tcb_prof.py)
  float dt_bias[256];    // constant part of the dir|transient.0 bias (W_dt b_final)
  float t3_bias[128];
  uint32_t btbl[10 * 128];  // packed 16-bit bias pairs: rows 0..7 trunk, 8/9 transient_encoding.2/.4
};

template <typename T> __device__ __forceinline__ uint32_t gt0_mask2(uint32_t pk);
template <> __device__ __forceinline__ uint32_t gt0_mask2<__half>(uint32_t pk) {
  return __hgt2_mask(*reinterpret_cast<__half2*>(&pk), __floats2half2_rn(0.f, 0.f));
}
template <> __device__ __forceinline__ uint32_t gt0_mask2<__nv_bfloat16>(uint32_t pk) {
  return __hgt2_mask(*reinterpret_cast<__nv_bfloat162*>(&pk), __floats2bfloat162_rn(0.f, 0.f));
}
// Mask word of a 32-column block: bit q (q = 0..15) = column 2q, bit 16+q = column 2q+1, i.e. the two 16-bit
// lanes of packed pair q.  expand2 turns the pair's two bits into 0xFFFF lanes: shift them to the sign bits of
// bytes 1 and 3, then PRMT in sign-replicate mode.
__device__ __forceinline__ uint32_t expand2(uint32_t w, int q) {
  uint32_t r;
  asm("prmt.b32 %0, %1, 0, 0xBB99;" : "=r"(r) : "r"(w << (15 - q)));
  return r;
}
// One shift serves two pairs: after w << (15 - q) (q = 8..15) pair q sits on the sign bits of bytes 1 / 3 and pair q - 8 on
// those of bytes 0 / 2.
__device__ __forceinline__ void expand4(uint32_t w, int q, uint32_t& m_q, uint32_t& m_q8) {
  const uint32_t sh = w << (15 - q);
  asm("prmt.b32 %0, %1, 0, 0xBB99;" : "=r"(m_q) : "r"(sh));
  asm("prmt.b32 %0, %1, 0, 0xAA88;" : "=r"(m_q8) : "r"(sh));
}
__device__ __forceinline__ void apply_mask16(uint32_t (&pk)[16], uint32_t w) {
#pragma unroll
  for (int q = 8; q < 16; ++q) {
    uint32_t m1, m0;
    expand4(w, q, m1, m0);
    pk[q] &= m1, pk[q - 8] &= m0;
  }
}
__device__ __forceinline__ int mask_pos(int j) { return (j & 1) * 16 + (j >> 1); }

// Two 32-column blocks per trip, the next block's TMEM read in flight while the current one is processed.
// f(v, cb, w): v = 32 fp32 accumulators of block cb, w = mask word of the block (0 if mrow == null).
template <typename F>
__device__ __forceinline__ void for_blocks(uint32_t t_row, int ncb, const uint32_t* mrow, F&& f) {
  uint32_t v0[32], v1[32];
  tmem_ld32(t_row, v0);
#pragma unroll 1
  for (int cb = 0; cb < ncb; cb += 2) {
    const uint32_t w0 = mrow ? mrow[cb * 128] : 0u, w1 = mrow ? mrow[(cb + 1) * 128] : 0u;
    tmem_ld_wait(v0);
    tmem_ld32(t_row + (cb + 1) * 32, v1);
    f(v0, cb, w0);
    tmem_ld_wait(v1);
    if (cb + 2 < ncb) tmem_ld32(t_row + (cb + 2) * 32, v0);
    f(v1, cb + 1, w1);
  }
}

// Same, with the layer's mask words already in registers (loaded BEFORE the wait for the accumulator, so their L2 / DRAM
// latency hides behind the MMAs): the words rotate through mw[0..1], no dynamic register indexing.
template <typename F>
__device__ __forceinline__ void for_blocks_m(uint32_t t_row, int ncb, uint32_t (&mw)[8], F&& f) {
  uint32_t v0[32], v1[32];
  tmem_ld32(t_row, v0);
#pragma unroll 1
  for (int cb = 0; cb < ncb; cb += 2) {
    const uint32_t w0 = mw[0], w1 = mw[1];
#pragma unroll
    for (int i = 0; i < 6; ++i) mw[i] = mw[i + 2];
    tmem_ld_wait(v0);
    tmem_ld32(t_row + (cb + 1) * 32, v1);
    f(v0, cb, w0);
    tmem_ld_wait(v1);
    if (cb + 2 < ncb) tmem_ld32(t_row + (cb + 2) * 32, v0);
    f(v1, cb + 1, w1);
  }
}

template <typename T>
__device__ __forceinline__ void store_block(uint32_t dst, const uint32_t (&pk)[16]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) st_shared_v4(dst + q * kPanelBytes, pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
}

#ifdef DFB_TC_PROF
#define TCB_PROF_DECL unsigned long long pacc[4] = {0, 0, 0, 0}; const long long prof_t0 = clock64();
#define TCB_PROF_WAIT(i, stmt) { const long long _t = clock64(); stmt; pacc[i] += clock64() - _t; }
#define TCB_PROF_FLUSH(base)                                                                        \
  if (a.prof && (threadIdx.x & 31) == 0) {                                                          \
    for (int _i = 0; _i < 3; ++_i) a.prof[(size_t)blockIdx.x * 64 + (base) + _i] = pacc[_i];        \
    a.prof[(size_t)blockIdx.x * 64 + (base) + 3] = clock64() - prof_t0;                             \
  }
#define TCB_PROF_PTR pacc
#else
#define TCB_PROF_DECL
#define TCB_PROF_WAIT(i, stmt) { stmt; }
#define TCB_PROF_FLUSH(base)
#define TCB_PROF_PTR nullptr
#endif

// One weight chunk = KS K-steps.  Every ring stage has its own full / empty barrier pair: the refill of a stage starts as
// soon as ITS MMAs have completed and the stage is waited for on its own, so n_stages - 1 chunks of MMA work cover the
// refill latency (with stages armed in pairs only one pair did, and the issuer waited 21 % of the kernel for weights).
template <int CG, int KS>
__device__ __forceinline__ void issue_step(uint32_t& stage, uint32_t& phase, uint32_t n_stages, int nch, uint32_t a_lo,
                                           uint32_t b_rows, uint32_t d_tmem, uint32_t idesc, uint32_t sW, uint32_t sBar, int* err,
                                           unsigned long long* pacc, uint32_t acc) {
  const uint32_t desc_hi = (128u >> 4) | (1u << 14);
  const uint32_t b_step = 2u * b_rows;
#pragma unroll 1
  for (int c = 0; c < nch; ++c) {
#ifdef DFB_TC_PROF
    const long long _t = clock64();
#endif
    mbar_wait(sBar + 8u * (W_FULL + stage), phase, err);
#ifdef DFB_TC_PROF
    pacc[0] += clock64() - _t;
#endif
    tc_fence_after();
    const uint32_t b_lo = ((sW + stage * kChunkBytes) >> 4) | (b_rows << 16);
    if (elect_one()) {
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
        umma_f16<CG>(d_tmem, mk64(a_lo + ks * 256, desc_hi), mk64(b_lo + ks * b_step, desc_hi), idesc, acc | ks);
      umma_commit<CG>(sBar + 8u * (W_EMPTY + stage));
    }
    __syncwarp();
    a_lo += 256u * KS;
    acc = 1;
    if (++stage == n_stages) stage = 0, phase ^= 1;
  }
}

// CG = 2 (cta_group::2): the two CTAs of a cluster work on one 256-row tile per slot with ONE weight stream and one MMA
// instruction for both SMs, exactly like the forward kernel (mlp_tc.cu, mlp_tc_body): each CTA keeps its own 128 rows
// (A operand, TMEM accumulators, masks, epilogue) and half of every weight chunk (2-SM TMA crediting the leader's barrier);
// tcgen05.commit multicasts the stage / accumulator hand-offs, one lane per warp arrives on the leader's barriers.
// The 1-CTA kernel reads the whole B operand per SM and writes the whole weight stream into its shared memory: 187 B/clk
// of shared-memory traffic against 128 B/clk available, which is what bounded it.
template <typename T, int CG>
__device__ __forceinline__ void mlp_tc_bwd_body(const BtArgs& a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sA = smem_u32(smem);
  const uint32_t sW = sA + 2u * (uint32_t)a.slot_bytes;
  const uint32_t sBar = sW + (uint32_t)a.n_stages * kChunkBytes;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 2 * a.slot_bytes + a.n_stages * kChunkBytes + N_BARS * 8);
  const int tid = threadIdx.x, warp = tid >> 5;
  auto bar = [&](int i) { return sBar + 8u * i; };
  const int fmt = std::is_same<T, __nv_bfloat16>::value ? 1 : 0;
  uint32_t* scr = a.scratch + (size_t)blockIdx.x * kScrWords;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  const int64_t unit0 = blockIdx.x / CG, n_units = gridDim.x / CG;   // a unit = CTA (CG 1) or CTA pair (CG 2)

  if (tid == 0) {
    for (int i = 0; i < kMaxStages; ++i) mbar_init(bar(W_FULL + i), 1), mbar_init(bar(W_EMPTY + i), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar(D_FULL + s), 1);
      mbar_init(bar(A_READY + s), 4 * CG);     // one arrival per warp (lane 0, after __syncwarp), see mlp_tc.cu
      mbar_init(bar(PASS_DONE + s), 4 * CG);
      mbar_init(bar(PE_READY + s), 4 * CG);
      mbar_init(bar(PE_FREE + s), 1);
    }
    fence_barrier_init();
  }
  if (warp == 13) tmem_alloc<CG>(smem_u32(tmem_slot), 512);
  tc_fence_before();
  if (CG == 2) cluster_sync(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 12) {
    // ===== weight producer =====================================================================
    uint32_t stage = 0, phase = 0;
    TCB_PROF_DECL
    const uint8_t* wimg = reinterpret_cast<const uint8_t*>(a.wimg);
    for (int64_t p = unit0; p < a.n_pass; p += n_units)
      for (int s = 0; s < a.n_steps; ++s) {
        const int nch = a.steps[s].n_chunks, cbytes = a.steps[s].cbytes;
        const uint8_t* src0 = wimg + a.steps[s].w_off;
        const int img0 = a.steps[s].img_base + (int)rank;
        for (int slot = 0; slot < 2; ++slot) {
          const uint8_t* src = src0;
          for (int c = 0; c < nch; ++c, src += cbytes) {
            TCB_PROF_WAIT(0, mbar_wait(bar(W_EMPTY + stage), phase ^ 1, a.error_flag));
            if (elect_one()) {
              if (CG == 1) {
                mbar_expect_tx(bar(W_FULL + stage), cbytes);
                bulk_g2s(sW + stage * kChunkBytes, src, cbytes, bar(W_FULL + stage));
              } else {
                // the leader arms ITS barrier with both halves' bytes; every CTA loads its own 16 KB image with a 2-SM
                // TMA whose completion is credited to the leader's barrier
                const uint32_t bar_leader = bar(W_FULL + stage) & 0xFEFFFFFFu;
                if (rank == 0) mbar_expect_tx(bar(W_FULL + stage), 2 * kChunkBytes);
                tma_load_img_2sm(sW + stage * kChunkBytes, &a.tmap, img0 + 2 * c, bar_leader);
              }
            }
            __syncwarp();
            if (++stage == (uint32_t)a.n_stages) stage = 0, phase ^= 1;
          }
        }
      }
    TCB_PROF_FLUSH(0)
  } else if (warp == 13 && rank != 0) {
    // peer CTA of a pair: the leader issues every MMA
  } else if (warp == 13) {
    // ===== MMA issuer ==========================================================================
    uint32_t stage = 0, phase = 0, na = 0;   // na: A_READY phases consumed (the same for both slots)
    int lp = 0;
    TCB_PROF_DECL
    for (int64_t p = unit0; p < a.n_pass; p += n_units, ++lp)
      for (int s = 0; s < a.n_steps; ++s) {
        const int nch = a.steps[s].n_chunks, nn = a.steps[s].n;
        const uint32_t idesc = make_idesc(fmt, nn, kTileM * CG);
        const uint32_t b_rows = (uint32_t)nn / CG;   // B rows held by each CTA
        const bool first = s == 0 || !a.steps[s - 1].chain;   // first MMA group of its accumulator
        const uint32_t acc0 = (uint32_t)a.steps[s].acc;
        for (int slot = 0; slot < 2; ++slot) {
          if (s == 0) {
            if (lp > 0) TCB_PROF_WAIT(1, mbar_wait(bar(PASS_DONE + slot), (lp - 1) & 1, a.error_flag));
            TCB_PROF_WAIT(2, mbar_wait(bar(PE_READY + slot), lp & 1, a.error_flag));
          } else if (first) {
            TCB_PROF_WAIT(1, mbar_wait(bar(A_READY + slot), na & 1, a.error_flag));
          }
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + slot * 256 + a.steps[s].d_col;
          const uint32_t a_lo = ((sA + slot * a.slot_bytes + a.steps[s].a_panel0 * kPanelBytes) >> 4) | ((kPanelBytes >> 4) << 16);
          const uint32_t nst = (uint32_t)a.n_stages;
          if (a.steps[s].kshort) issue_step<CG, 1>(stage, phase, nst, nch, a_lo, b_rows, d_tmem, idesc, sW, sBar, a.error_flag, TCB_PROF_PTR, acc0);
          else if (nn == 256) issue_step<CG, 2 * CG>(stage, phase, nst, nch, a_lo, b_rows, d_tmem, idesc, sW, sBar, a.error_flag, TCB_PROF_PTR, acc0);
          else if (nn == 128) issue_step<CG, 4 * CG>(stage, phase, nst, nch, a_lo, b_rows, d_tmem, idesc, sW, sBar, a.error_flag, TCB_PROF_PTR, acc0);
          else issue_step<CG, 8 * CG>(stage, phase, nst, nch, a_lo, b_rows, d_tmem, idesc, sW, sBar, a.error_flag, TCB_PROF_PTR, acc0);
          if (elect_one()) {
            if (!a.steps[s].chain) umma_commit<CG>(bar(D_FULL + slot));
            // last reader of the positional-encoding panels (the skip layer, then ghA in the same panels)
            if (s == a.pe_free_step) umma_commit<CG>(bar(PE_FREE + slot));
          }
          __syncwarp();
        }
        if (s > 0 && first) ++na;
      }
    TCB_PROF_FLUSH(4)
  } else if (warp >= 8 && warp < 12) {
    // ===== encoder: positional encoding of the next pass + its Jacobian for step 25 ================
    const int r = tid - 256;
    int lp = 0;
    for (int64_t p = unit0; p < a.n_pass; p += n_units, ++lp)
      for (int slot = 0; slot < 2; ++slot) {
        if (lp > 0) mbar_wait_relaxed(bar(PE_FREE + slot), (lp - 1) & 1, a.error_flag);
        int64_t g = ((2 * p + slot) * CG + rank) * kTileM + r;
        g = g < a.P ? g : a.P - 1;
        const int64_t ray = g / a.S;
        const float* rr = a.rayrec + ray * kRayRec;
        const float zz = __ldg(a.z + g);
        float pt[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) pt[c] = __fadd_rn(__ldg(rr + c), __fmul_rn(__ldg(rr + 3 + c), zz));
        const uint32_t dst = sA + slot * a.slot_bytes + kHPanels * kPanelBytes + r * 16;
        float* jac = reinterpret_cast<float*>(scr + kScrJac + ((lp & 1) * 2 + slot) * 64 * 128) + r;
        const bool want_pe = a.saved_masks == nullptr;   // the forward recompute alone reads the encoding itself
        auto put = [&](int col, float v, float j) {
          T h = (T)v;
          if (want_pe) st_shared_b16(dst + (uint32_t)(col >> 3) * kPanelBytes + (col & 7) * 2, *reinterpret_cast<uint16_t*>(&h));
          jac[col * 128] = j;
        };
        put(0, pt[0], 1.f), put(1, pt[1], 1.f), put(2, pt[2], 1.f), put(63, 0.f, 0.f);
#pragma unroll 1
        for (int l = 0; l < 10; ++l) {
          const float fr = (float)(1 << l);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float sn, cs;
            sincosf(__fmul_rn(pt[c], fr), &sn, &cs);
            put(3 + 6 * l + c, sn, fr * cs);
            put(3 + 6 * l + 3 + c, cs, -fr * sn);
          }
        }
        __threadfence_block();
        fence_proxy_async();
        __syncwarp();
        if ((tid & 31) == 0) arrive_leader<CG>(bar(PE_READY + slot));
      }
  } else if (warp < 8) {
    // ===== epilogue warpgroups (thread = accumulator row = sample) ================================
    const int slot = warp >> 2;
    const int r = tid & 127;
    const uint32_t t_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + slot * 256;
    const uint32_t h_row = sA + slot * a.slot_bytes + r * 16;
    uint32_t* mbase = scr + kScrMask + slot * kMaskLayers * 8 * 128 + r;
    float* skip = reinterpret_cast<float*>(scr + kScrSkip + slot * 64 * 128) + r;
    uint32_t nd = 0;
    int lp = 0;
    TCB_PROF_DECL
#ifdef DFB_TC_PROF
    unsigned long long pstep[kSteps] = {0};
#endif
    for (int64_t p = unit0; p < a.n_pass; p += n_units, ++lp) {
      const int64_t tile = (2 * p + slot) * CG + rank;   // 128-row tile of the flattened [ray][sample] array
      const int64_t g = tile * kTileM + r;
      const bool valid = g < a.P;
      const int64_t gc = valid ? g : a.P - 1;
      const float* rb = a.raybias + (gc / a.S) * 256;
      asm volatile("prefetch.global.L1 [%0];" ::"l"(rb + (tid & 7) * 32));
      float isc = 1.f;
      // saved masks cover ceil(P/128) tiles; a tile past the end (odd tile count) has no rows and reads the scratch
      const uint32_t* mpass = a.saved_masks && tile * kTileM < a.P
                                  ? a.saved_masks + (size_t)tile * (kMaskLayers * 8 * 128) + r : mbase;
      for (int s = 0; s < a.n_steps; ++s) {
        const int kd = a.steps[s].kind, ncb = a.steps[s].ncb;
        if (kd == SUB) continue;   // first MMA group of a compound step: nothing to read yet
        uint32_t* mrow = const_cast<uint32_t*>(mpass) + a.steps[s].ml * 8 * 128;  // written only by the recompute kinds
        // mask words of the layer this step applies, requested before the accumulator is waited for (saved masks stream
        // from DRAM once: loaded at the point of use they exposed one memory latency per pair of 32-column blocks)
        uint32_t mw[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        if (kd == B_MASK) {
#pragma unroll
          for (int c = 0; c < 8; ++c) mw[c] = mrow[c * 128];
        }
        // head step: d heads = d raw . activation', scaled per row into fp16 range, written as the 16-wide A operand of the
        // short steps.  Nothing here depends on the accumulator, and ghA's panels have no reader left (see kPeFreeStep), so
        // it all happens before the wait.
        if (kd == HEADS || kd == F_T3) {
          float gh[10], mx = 0.f;
          gh[9] = 0.f;
#pragma unroll
          for (int c = 0; c < 9; ++c) {
            const float o = __ldg(a.raw + gc * 9 + c), gr = valid ? __ldg(a.g_raw + gc * 9 + c) : 0.f;
            const bool sg = c < 3 || (c >= 4 && c < 7);  // sigmoid outputs; the others are softplus
            gh[c] = sg ? gr * o * (1.f - o) : gr * (1.f - expf(-o));
            mx = fmaxf(mx, fabsf(gh[c]));
          }
          float sc = 1.f;
          isc = 1.f;
          if (mx > 0.f && mx < 3.0e38f) {
            int e;
            frexpf(mx, &e);
            e = max(-100, min(100, e));
            sc = exp2f((float)(8 - e)), isc = exp2f((float)(e - 8));
          } else {
#pragma unroll
            for (int c = 0; c < 9; ++c) gh[c] = 0.f;
          }
#pragma unroll
          for (int c = 0; c < 9; ++c) gh[c] *= sc;
          st_shared_v4(h_row + (uint32_t)kGhPanel * kPanelBytes, pack2<T>(gh[0], gh[1]), pack2<T>(gh[2], gh[3]), pack2<T>(gh[4], gh[5]),
                       pack2<T>(gh[6], gh[7]));
          st_shared_v4(h_row + (uint32_t)(kGhPanel + 1) * kPanelBytes, pack2<T>(gh[8], gh[9]), 0u, 0u, 0u);
        }
        TCB_PROF_WAIT(0, mbar_wait(bar(D_FULL + slot), nd & 1, a.error_flag));
        ++nd;
        tc_fence_after();
#ifdef DFB_TC_PROF
        const long long _ts = clock64();
#endif
        const int64_t mdbg = (gc * kMaskLayers + a.steps[s].ml) * 8;
        auto put_mask = [&](int cb, uint32_t m) {
          if (a.mask_out && valid) a.mask_out[mdbg + cb] = m;
          if (a.mask_in) m = a.mask_in[mdbg + cb];
          mrow[cb * 128] = m;
        };
        if (kd == F_HID) {
          // relu(acc + b) in packed 16-bit math exactly like the forward kernel; mask bit = result > 0
          const int boff = a.steps[s].brow * 128;
          for_blocks(t_row, ncb, nullptr, [&](const uint32_t (&v)[32], int cb, uint32_t) {
            uint32_t pk[16], m = 0;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              pk[q] = add_relu2<T>(pack2<T>(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1])), a.btbl[boff + cb * 16 + q]);
              m |= gt0_mask2<T>(pk[q]) & (0x00010001u << q);
            }
            store_block<T>(h_row + (uint32_t)(cb * 4) * kPanelBytes, pk);
            put_mask(cb, m);
          });
        } else if (kd == F_DT) {
          // dir_encoding | transient_encoding.0: per-ray bias in fp32; only the transient half feeds a later layer
          for_blocks(t_row, 8, nullptr, [&](const uint32_t (&v)[32], int cb, uint32_t) {
            const float4* b4 = reinterpret_cast<const float4*>(rb + cb * 32);
            float x[32];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 bb = __ldg(b4 + q);
              x[4 * q + 0] = __uint_as_float(v[4 * q + 0]) + (bb.x + a.dt_bias[cb * 32 + 4 * q + 0]);
              x[4 * q + 1] = __uint_as_float(v[4 * q + 1]) + (bb.y + a.dt_bias[cb * 32 + 4 * q + 1]);
              x[4 * q + 2] = __uint_as_float(v[4 * q + 2]) + (bb.z + a.dt_bias[cb * 32 + 4 * q + 2]);
              x[4 * q + 3] = __uint_as_float(v[4 * q + 3]) + (bb.w + a.dt_bias[cb * 32 + 4 * q + 3]);
            }
            uint32_t m = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) m |= x[j] > 0.f ? (1u << mask_pos(j)) : 0u;
            put_mask(cb, m);
            if (cb >= 4) {
              uint32_t pk[16];
#pragma unroll
              for (int q = 0; q < 16; ++q) pk[q] = pack2<T>(fmaxf(x[2 * q], 0.f), fmaxf(x[2 * q + 1], 0.f));
              store_block<T>(h_row + (uint32_t)((cb - 4) * 4) * kPanelBytes, pk);
            }
          });
        } else if (kd == F_T3 || kd == HEADS) {
          // last transient layer: only its mask is needed (HEADS: saved by the forward, nothing to do here)
          if (kd == F_T3)
            for_blocks(t_row, 4, nullptr, [&](const uint32_t (&v)[32], int cb, uint32_t) {
              uint32_t m = 0;
#pragma unroll
              for (int j = 0; j < 32; ++j) m |= (__uint_as_float(v[j]) + a.t3_bias[cb * 32 + j]) > 0.f ? (1u << mask_pos(j)) : 0u;
              put_mask(cb, m);
            });
        } else if (kd == B_MASK) {
          // g_in = (g_out W) . relu'(producer)
          for_blocks_m(t_row, ncb, mw, [&](const uint32_t (&v)[32], int cb, uint32_t w) {
            uint32_t pk[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) pk[q] = pack2<T>(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1]));
            apply_mask16(pk, w);
            store_block<T>(h_row + (uint32_t)(cb * 4) * kPanelBytes, pk);
          });
        } else if (kd == B_DIRPE) {
          uint32_t v[32];
          tmem_ld32(t_row, v);
          tmem_ld_wait(v);
          if (valid) {
            // g_samp[g][3..29]: one scalar, six 16-byte stores, two scalars
            float* o = a.g_samp + g * 32;
            o[3] = __uint_as_float(v[0]) * isc;
#pragma unroll
            for (int q = 0; q < 6; ++q)
              *reinterpret_cast<float4*>(o + 4 + 4 * q) =
                  make_float4(__uint_as_float(v[1 + 4 * q]) * isc, __uint_as_float(v[2 + 4 * q]) * isc,
                              __uint_as_float(v[3 + 4 * q]) * isc, __uint_as_float(v[4 + 4 * q]) * isc);
            o[28] = __uint_as_float(v[25]) * isc, o[29] = __uint_as_float(v[26]) * isc;
          }
        } else if (kd == B_PESKIP) {
          for_blocks(t_row, 2, nullptr, [&](const uint32_t (&v)[32], int cb, uint32_t) {
#pragma unroll
            for (int j = 0; j < 32; ++j) skip[(cb * 32 + j) * 128] = __uint_as_float(v[j]);
          });
        } else {  // B_PE0
          const float* jac = reinterpret_cast<const float*>(scr + kScrJac + ((lp & 1) * 2 + slot) * 64 * 128) + r;
          float d[3] = {0.f, 0.f, 0.f};
          uint32_t v0[32], v1[32];
          tmem_ld32(t_row, v0);
          tmem_ld32(t_row + 32, v1);
          tmem_ld_wait(v0);
#pragma unroll
          for (int j = 0; j < 32; ++j)   // the component of a column is periodic in 3 from column 0
            d[j % 3] = fmaf(__uint_as_float(v0[j]) + skip[j * 128], jac[j * 128], d[j % 3]);
          tmem_ld_wait(v1);
#pragma unroll
          for (int j = 0; j < 32; ++j)
            d[(32 + j) % 3] = fmaf(__uint_as_float(v1[j]) + skip[(32 + j) * 128], jac[(32 + j) * 128], d[(32 + j) % 3]);
          if (valid) {
            a.g_samp[g * 32 + 0] = d[0] * isc, a.g_samp[g * 32 + 1] = d[1] * isc, a.g_samp[g * 32 + 2] = d[2] * isc;
          }
        }
        tc_fence_before();
        fence_proxy_async();
        __syncwarp();
        if ((tid & 31) == 0) arrive_leader<CG>(bar((s + 1 < a.n_steps ? A_READY : PASS_DONE) + slot));
#ifdef DFB_TC_PROF
        pstep[s] += clock64() - _ts;
#endif
      }
    }
    if ((warp & 3) == 0) {
      TCB_PROF_FLUSH(8 + 4 * slot)
#ifdef DFB_TC_PROF
      if (a.prof && slot == 0 && (threadIdx.x & 31) == 0)
        for (int _i = 0; _i < kSteps; ++_i) a.prof[(size_t)blockIdx.x * 64 + 16 + _i] = pstep[_i];
#endif
    }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync(); else __syncthreads();
  if (warp == 13) {
    tc_fence_after();
    tmem_dealloc<CG>(tmem_base, 512);
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 1) k_mlp_tc_bwd(const __grid_constant__ BtArgs a) {
  mlp_tc_bwd_body<T, 1>(a);
}

template <typename T>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) k_mlp_tc_bwd2(const __grid_constant__ BtArgs a) {
  mlp_tc_bwd_body<T, 2>(a);
}

}  // namespace tcb

// ---------------------------------------------------------------------------------------
// host side: packed weight image of the 26 steps, tables, launch
// ---------------------------------------------------------------------------------------
namespace {

uint16_t f2h(float f) { __half h = __float2half_rn(f); uint16_t u; memcpy(&u, &h, 2); return u; }
uint16_t f2b(float f) { __nv_bfloat16 h = __float2bfloat16_rn(f); uint16_t u; memcpy(&u, &h, 2); return u; }

// logical matrices: 0..7 trunk forward, 8 dir|transient.0 (folded), 9..11 transient 2,4,6 forward;
// 20+i: transpose of transient_encoding.{2,4,6}[i]; 30 dirPE columns of dir_encoding; 31 folded^T;
// 40+i trunk layer i transposed (hidden part), 50 skip layer's PE part, 51 layer 0 transposed;
// short (K = 16, the head derivatives' index in `raw` order): 60 transient heads^T, 61 rgb head^T, 62 sigma head^T
struct BStep { int logical, K, N, a_panel0, kind, ml, brow, ncb, d_col, kshort, acc, chain; };

std::vector<BStep> bwd_program() {
  using namespace tcb;
  const int G = kGhPanel;
  std::vector<BStep> pr;
  for (int i = 0; i < 8; ++i) pr.push_back({i, i == 0 ? 64 : (i == 4 ? 320 : 256), 256, i == 0 ? 32 : 0, F_HID, i, i, 8, 0, 0, 0, 0});
  pr.push_back({8, 256, 256, 0, F_DT, 8, 0, 8, 0, 0, 0, 0});
  pr.push_back({9, 128, 128, 0, F_HID, 9, 8, 4, 0, 0, 0, 0});
  pr.push_back({10, 128, 128, 0, F_HID, 10, 9, 4, 0, 0, 0, 0});
  pr.push_back({11, 128, 128, 0, F_T3, 11, 0, 4, 0, 0, 0, 0});     // + head derivatives -> ghA
  pr.push_back({60, 16, 128, G, B_MASK, 11, 0, 4, 0, 1, 0, 0});    // g_T3 = d heads . transient heads^T
  pr.push_back({22, 128, 128, 0, B_MASK, 10, 0, 4, 0, 0, 0, 0});   // g_T2
  pr.push_back({21, 128, 128, 0, B_MASK, 9, 0, 4, 0, 0, 0, 0});    // g_T1
  pr.push_back({61, 16, 128, G, SUB, 0, 0, 0, 0, 1, 0, 1});        // g_dir (columns 0..127) = d rgb . rgb head^T
  pr.push_back({20, 128, 128, 0, B_MASK, 8, 0, 8, 128, 0, 0, 0});  // g_t0 (columns 128..255); mask of dir|transient.0
  pr.push_back({30, 128, 128, 0, B_DIRPE, 0, 0, 1, 0, 0, 0, 0});   // d dirPE
  pr.push_back({31, 256, 256, 0, SUB, 0, 0, 0, 0, 0, 0, 1});       // g_h7 = [g_dir | g_t0] (W_dt W_final) ...
  pr.push_back({62, 16, 256, G, B_MASK, 7, 0, 8, 0, 1, 1, 0});     // ... + d sigma * w_sigma
  pr.push_back({47, 256, 256, 0, B_MASK, 6, 0, 8, 0, 0, 0, 0});    // g_h6 = g_h7 W_7
  pr.push_back({46, 256, 256, 0, B_MASK, 5, 0, 8, 0, 0, 0, 0});
  pr.push_back({45, 256, 256, 0, B_MASK, 4, 0, 8, 0, 0, 0, 0});    // g_h4 = g_h5 W_5
  pr.push_back({50, 256, 64, 0, B_PESKIP, 0, 0, 2, 0, 0, 0, 0});   // skip layer, PE columns
  pr.push_back({44, 256, 256, 0, B_MASK, 3, 0, 8, 0, 0, 0, 0});    // g_h3 = g_h4 W_4[:, h]
  pr.push_back({43, 256, 256, 0, B_MASK, 2, 0, 8, 0, 0, 0, 0});
  pr.push_back({42, 256, 256, 0, B_MASK, 1, 0, 8, 0, 0, 0, 0});
  pr.push_back({41, 256, 256, 0, B_MASK, 0, 0, 8, 0, 0, 0, 0});    // g_h0 = g_h1 W_1
  pr.push_back({51, 256, 64, 0, B_PE0, 0, 0, 2, 0, 0, 0, 0});
  return pr;
}
constexpr int kFirstBwdStep = 12;   // program index of the first backward step (the saved-mask program starts here)
constexpr int kPeFreeStep = 19;     // the sigma short step: last reader of ghA (and, before it, of the encoding)

}  // namespace

int pack_tc_bwd_weights(DfbNerf* n, int which, const std::vector<std::vector<float>>& P_in) {
  NetPack& np = n->net[which];
  for (int k = 0; k < 2; ++k) {
    if (np.blob16b[k]) { cudaFree(np.blob16b[k]); np.blob16b[k] = nullptr; }
    if (np.blob16b2[k]) { cudaFree(np.blob16b2[k]); np.blob16b2[k] = nullptr; }
  }
  np.tcb_tbl.clear();
  if (!np.fine || !tc_padded_shape(np)) return DFB_OK;
  const std::vector<std::vector<float>> P = tc_pad_params(np, P_in);  // narrower networks: embedded in 8x256 with zeros
  const int W = 256, H = 128, in_xyz = np.in_xyz;
  const int kdd = W + np.in_dir + np.a_dim, ktt = W + np.t_dim;
  auto wdt = [&](int nn, int k) -> double {
    if (nn < H) return P[18][(size_t)nn * kdd + k];
    return P[24][(size_t)(nn - H) * ktt + k];
  };
  std::vector<float> fw((size_t)W * W), fb(W);
  for (int nn = 0; nn < W; ++nn) {
    double bacc = 0.0;
    for (int j = 0; j < W; ++j) bacc += wdt(nn, j) * (double)P[17][j];
    fb[nn] = (float)bacc;
    for (int k = 0; k < W; ++k) {
      double acc = 0.0;
      for (int j = 0; j < W; ++j) acc += wdt(nn, j) * (double)P[16][(size_t)j * W + k];
      fw[(size_t)nn * W + k] = (float)acc;
    }
  }
  // B[n][k] of every logical matrix (D = A B^T: k runs over the A operand's columns)
  auto wval = [&](int lg, int nn, int k) -> float {
    if (lg == 0) return k < in_xyz ? P[0][(size_t)nn * in_xyz + k] : 0.f;
    if (lg == 4) {  // K order [h(256) | pe(64)]; torch order is cat([input_xyz, h])
      if (k < W) return P[8][(size_t)nn * (W + in_xyz) + in_xyz + k];
      return k - W < in_xyz ? P[8][(size_t)nn * (W + in_xyz) + (k - W)] : 0.f;
    }
    if (lg < 8) return P[2 * lg][(size_t)nn * W + k];
    if (lg == 8) return fw[(size_t)nn * W + k];
    if (lg <= 11) return P[26 + 2 * (lg - 9)][(size_t)nn * H + k];
    if (lg >= 20 && lg <= 22) return P[26 + 2 * (lg - 20)][(size_t)k * H + nn];
    if (lg == 30) return nn < np.in_dir ? P[18][(size_t)k * kdd + W + nn] : 0.f;
    if (lg == 31) return fw[(size_t)k * W + nn];
    if (lg == 44) return P[8][(size_t)k * (W + in_xyz) + in_xyz + nn];
    if (lg >= 41 && lg <= 47) return P[2 * (lg - 40)][(size_t)k * W + nn];
    if (lg == 50) return nn < in_xyz ? P[8][(size_t)k * (W + in_xyz) + nn] : 0.f;
    if (lg == 51) return nn < in_xyz ? P[0][(size_t)k * in_xyz + nn] : 0.f;
    // head derivatives in `raw` order: k = 0..2 rgb, 3 sigma, 4..6 transient rgb, 7 transient sigma, 8 transient beta
    if (lg == 60) return k >= 4 && k < 7 ? P[34][(size_t)(k - 4) * H + nn] : (k == 7 ? P[32][nn] : (k == 8 ? P[36][nn] : 0.f));
    if (lg == 61) return k < 3 ? P[22][(size_t)k * H + nn] : 0.f;
    if (lg == 62) return k == 3 ? P[20][nn] : 0.f;
    return 0.f;
  };
  const std::vector<BStep> prog = bwd_program();
  size_t total = 0;   // bytes; a short step is K = 16: one chunk of two 8-column panels (N x 16 B each)
  for (const BStep& st : prog) total += (size_t)st.K * st.N * 2;
  std::vector<uint16_t> img16[2];
  img16[0].assign(total / 2, 0);
  img16[1].assign(total / 2, 0);
  size_t base = 0;    // elements
  for (const BStep& st : prog) {
    const int kc = st.kshort ? 16 : tcb::kChunkBytes / (st.N * 2);
    for (int k0 = 0; k0 < st.K; k0 += kc, base += (size_t)kc * st.N) {
      for (int kk = 0; kk < kc; ++kk)
        for (int r = 0; r < st.N; ++r) {
          const float v = wval(st.logical, r, k0 + kk);
          const size_t idx = base + (size_t)(kk / 8) * st.N * 8 + (size_t)r * 8 + kk % 8;
          img16[0][idx] = f2h(v);
          img16[1][idx] = f2b(v);
        }
    }
  }
  np.blob16b_bytes = total;
  for (int k = 0; k < 2; ++k) {
    DFB_CHECK_CUDA(cudaMalloc(&np.blob16b[k], np.blob16b_bytes));
    DFB_CHECK_CUDA(cudaMemcpy(np.blob16b[k], img16[k].data(), np.blob16b_bytes, cudaMemcpyHostToDevice));
  }
  // cta_group::2 layout: every chunk is two 16 KB images, image h = rows [h N/2, (h+1) N/2) of B over the chunk's K
  // columns (64 / 128 / 256 for N = 256 / 128 / 64; a short step fills the first N/2 x 32 bytes of its images)
  {
    size_t n_img = 0;
    for (const BStep& st : prog) {
      const int rows = st.N / 2, kc = st.kshort ? 16 : tcb::kChunkBytes / (rows * 2);
      DFB_REQUIRE(st.K % kc == 0, DFB_ERR_INVALID, "backward program: K = %d is not a whole number of %d-column chunks", st.K, kc);
      n_img += 2 * (size_t)(st.K / kc);
    }
    std::vector<uint16_t> im2[2];
    im2[0].assign(n_img * (tcb::kChunkBytes / 2), 0);
    im2[1].assign(n_img * (tcb::kChunkBytes / 2), 0);
    size_t img = 0;
    for (const BStep& st : prog) {
      const int rows = st.N / 2, kc = st.kshort ? 16 : tcb::kChunkBytes / (rows * 2);
      for (int k0 = 0; k0 < st.K; k0 += kc)
        for (int h = 0; h < 2; ++h, ++img) {
          const size_t b0 = img * (tcb::kChunkBytes / 2);
          for (int kk = 0; kk < kc; ++kk)
            for (int r = 0; r < rows; ++r) {
              const float v = wval(st.logical, h * rows + r, k0 + kk);
              const size_t idx = b0 + (size_t)(kk / 8) * rows * 8 + (size_t)r * 8 + kk % 8;
              im2[0][idx] = f2h(v);
              im2[1][idx] = f2b(v);
            }
        }
    }
    np.blob16b2_bytes = n_img * tcb::kChunkBytes;
    for (int k = 0; k < 2; ++k) {
      DFB_CHECK_CUDA(cudaMalloc(&np.blob16b2[k], np.blob16b2_bytes));
      DFB_CHECK_CUDA(cudaMemcpy(np.blob16b2[k], im2[k].data(), np.blob16b2_bytes, cudaMemcpyHostToDevice));
    }
  }
  // fp32 table: [10][256] biases of the packed rows | dt_bias 256 | t3_bias 128
  np.tcb_tbl.assign(10 * 256 + 256 + 128, 0.f);
  float* tb = np.tcb_tbl.data();
  for (int i = 0; i < 8; ++i) memcpy(tb + i * 256, P[2 * i + 1].data(), 256 * sizeof(float));
  memcpy(tb + 8 * 256, P[27].data(), H * sizeof(float));
  memcpy(tb + 9 * 256, P[29].data(), H * sizeof(float));
  float* q = tb + 10 * 256;
  memcpy(q, fb.data(), 256 * sizeof(float)), q += 256;
  memcpy(q, P[31].data(), H * sizeof(float));
  return DFB_OK;
}

bool tc_bwd_supported(const DfbNerf* n) {
  const NetPack& np = n->net[1];
  return np.loaded && np.fine && np.blob16b[0] != nullptr && np.blob16b2[0] != nullptr && !np.tcb_tbl.empty();
}

static unsigned long long* g_tcb_prof = nullptr;
const uint32_t* g_dbg_tc_mask_in = nullptr;
uint32_t* g_dbg_tc_mask_out = nullptr;

// Fine-network backward for P = n_rays*S samples: g_samp[P,32] from raw / g_raw (see the header comment).
int launch_mlp_tc_bwd(const DfbNerf* nerf, int kind, const float* rayrec, const float* z, const float* raybias,
                      const float* raw, const float* g_raw, int64_t n_rays, int S, float* g_samp, cudaStream_t st,
                      const uint32_t* saved_masks) {
  const NetPack& np = nerf->net[1];
  DFB_REQUIRE(tc_bwd_supported(nerf), DFB_ERR_UNSUPPORTED, "network shape not supported by the tcgen05 backward kernel");
  DFB_REQUIRE(kind == DFB_MMA_F16 || kind == DFB_MMA_BF16, DFB_ERR_INVALID, "bad mma kind");
  int* error_flag = nullptr;
  {
    const int rc = device_error_flag(&error_flag);
    if (rc) return rc;
  }
  tcb::BtArgs a;
  memset(&a, 0, sizeof(a));
  const std::vector<BStep> prog = bwd_program();
  DFB_REQUIRE((int)prog.size() == tcb::kSteps, DFB_ERR_INVALID, "backward program / kSteps mismatch");
  // DFB_TC_CTA_GROUP=1 selects the 1-CTA kernel (as for the forward kernel); the default is the cta_group::2 pair kernel
  const int cg = tc_cta_group_env();
  int ns = 0, img = 0;
  size_t woff = 0;
  for (int s = 0; s < tcb::kSteps; ++s) {
    const BStep& ls = prog[s];
    int cbytes, nch;
    if (cg == 1) {
      cbytes = ls.kshort ? ls.N * 32 : tcb::kChunkBytes;
      nch = (int)((size_t)ls.K * ls.N * 2 / cbytes);
    } else {
      cbytes = tcb::kChunkBytes;
      nch = ls.kshort ? 1 : ls.K / (tcb::kChunkBytes / ((ls.N / 2) * 2));
    }
    const tcb::Step st_full = {nch, cbytes, img, ls.N, ls.a_panel0, (uint32_t)woff, ls.kind, ls.ml, ls.brow,
                               ls.ncb, ls.d_col, ls.kshort, ls.acc, ls.chain};
    woff += (size_t)ls.K * ls.N * 2;
    img += 2 * (ls.kshort ? 1 : ls.K / (tcb::kChunkBytes / ((ls.N / 2) * 2)));
    if (!saved_masks) a.steps[ns++] = st_full;
    else if (s == kFirstBwdStep - 1) a.steps[ns++] = {0, tcb::kChunkBytes, 0, 128, 0, 0u, tcb::HEADS, 11, 0, 0, 0, 0, 0, 0};  // head derivatives only, no MMA
    else if (s >= kFirstBwdStep) a.steps[ns++] = st_full;
  }
  a.n_steps = ns;
  // saved masks: the positional encoding is not needed in shared memory (only ghA), which buys a fifth ring stage
  const int pe_panels = saved_masks ? tcb::kGhPanels : tcb::kPePanels;
  a.n_stages = saved_masks ? 5 : 4;
  a.slot_bytes = tcb::slot_bytes(pe_panels);
  const int smem_bytes = tcb::smem_total(pe_panels, a.n_stages);
  a.pe_free_step = saved_masks ? kPeFreeStep - (kFirstBwdStep - 1) : kPeFreeStep;
  a.saved_masks = saved_masks;
  a.wimg = cg == 2 ? np.blob16b2[kind == DFB_MMA_F16 ? 0 : 1] : np.blob16b[kind == DFB_MMA_F16 ? 0 : 1];
  if (cg == 2) {
    const int rc = make_weight_tmap(const_cast<void*>(a.wimg), np.blob16b2_bytes, &a.tmap);
    if (rc) return rc;
  }
  const float* tb = np.tcb_tbl.data();
  for (int i = 0; i < 10 * 128; ++i) {
    const float lo = tb[2 * i], hi = tb[2 * i + 1];
    a.btbl[i] = kind == DFB_MMA_F16 ? ((uint32_t)f2h(hi) << 16 | f2h(lo)) : ((uint32_t)f2b(hi) << 16 | f2b(lo));
  }
  const float* q = tb + 10 * 256;
  memcpy(a.dt_bias, q, sizeof(a.dt_bias)), q += 256;
  memcpy(a.t3_bias, q, sizeof(a.t3_bias));
  a.rayrec = rayrec, a.z = z, a.raybias = raybias, a.raw = raw, a.g_raw = g_raw, a.g_samp = g_samp;
  a.S = S, a.P = n_rays * S, a.error_flag = error_flag;
  a.mask_in = g_dbg_tc_mask_in, a.mask_out = g_dbg_tc_mask_out;
  if (a.P == 0) return DFB_OK;
#ifdef DFB_TC_PROF
  if (!g_tcb_prof) DFB_CHECK_CUDA(cudaMalloc(&g_tcb_prof, 256 * 64 * sizeof(unsigned long long)));
  DFB_CHECK_CUDA(cudaMemsetAsync(g_tcb_prof, 0, 256 * 64 * sizeof(unsigned long long), st));
  a.prof = g_tcb_prof;
#endif
  const int64_t tiles = (a.P + tcb::kTileM - 1) / tcb::kTileM;
  a.n_pass = (tiles + 2 * cg - 1) / (2 * cg);
  const int grid = cg * (int)std::min<int64_t>(a.n_pass, nerf->num_sms / cg);
  if (nerf->bwd_scratch_ctas < grid) {  // per handle (= per device), see DfbNerf
    if (nerf->bwd_scratch) cudaFree(nerf->bwd_scratch);
    nerf->bwd_scratch = nullptr, nerf->bwd_scratch_ctas = 0;
    DFB_CHECK_CUDA(cudaMalloc(&nerf->bwd_scratch, (size_t)nerf->num_sms * tcb::kScrWords * 4));
    nerf->bwd_scratch_ctas = nerf->num_sms;
  }
  a.scratch = nerf->bwd_scratch;
  auto launch = [&](auto kern) -> int {
    DFB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tcb::smem_total(tcb::kPePanels, 4)));
    kern<<<grid, tcb::kThreads, smem_bytes, st>>>(a);
    DFB_LAUNCH_CHECK();
    return DFB_OK;
  };
  if (cg == 2) return kind == DFB_MMA_F16 ? launch(tcb::k_mlp_tc_bwd2<__half>) : launch(tcb::k_mlp_tc_bwd2<__nv_bfloat16>);
  if (kind == DFB_MMA_F16) return launch(tcb::k_mlp_tc_bwd<__half>);
  return launch(tcb::k_mlp_tc_bwd<__nv_bfloat16>);
}

}  // namespace dfb

// Debug seam: cycle counters of the last tcgen05 backward launch, [n_cta][64] (DFB_TC_PROF builds only; tools/tcb_prof.py).
extern "C" int dfb_debug_tcb_prof(unsigned long long* out_host, int n_cta) {
  if (!dfb::g_tcb_prof) return DFB_ERR_UNSUPPORTED;
  cudaDeviceSynchronize();
  cudaMemcpy(out_host, dfb::g_tcb_prof, (size_t)std::min(n_cta, 256) * 64 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  return DFB_OK;
}
