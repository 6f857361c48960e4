// tcgen05 NeRF-W MLP (8x256 of BASELINE config[1]/[4], 8x128 of the reference's shipped configs as its own program,
// other widths zero-padded into 8x256): a persistent, warp-specialised kernel.  Each CTA owns all 512 TMEM columns and
// works on two 128-sample tiles ("slots") that alternate between the tensor pipe and the epilogue warps.  Default
// variant k_mlp_tc2: CTA PAIRS (cta_group::2) share every weight chunk and one MMA instruction feeds both SMs;
// k_mlp_tc (DFB_TC_CTA_GROUP=1) is the 1-CTA variant.
//
//   warp 12      weight producer : pre-packed 16 KB weight chunks from L2 into a 4-stage shared-memory ring
//                                  (2-SM TMA of this CTA's half of a chunk, credited to the leader's mbarrier;
//                                  1-CTA variant: cp.async.bulk)
//   warp 13      MMA issuer      : one elected thread of the pair's leader issues tcgen05.mma.cta_group::2.kind::f16
//                                  (M = 256 over the pair, N = 256|128|64, K = 16) with A = activations in shared
//                                  memory, B = weight chunk, D = fp32 accumulators in TMEM; multicast tcgen05.commit
//                                  releases ring stages and publishes finished layers in both CTAs
//   warps 0-7    epilogue        : thread = accumulator row (sample) of BOTH slots, the two warpgroups split every
//                                  step's columns; tcgen05.ld 32x32b, bias + ReLU as packed HFMA2.RELU, 16-bit store of
//                                  the next layer's A operand in place; the output heads are MMA steps of their own
//                                  (N = 64) whose epilogue applies softplus / sigmoid and, on the render_path step,
//                                  composites the ray segments in registers (fused_composite); ONE barrier arrival per
//                                  warp with CTA-scope release (see mbar_arrive_cluster in tc_common.cuh)
//   warps 8-11   encoder         : pts = o + d*z and the 63-wide positional encoding of the NEXT
//                                  pass's tiles, written straight into the A-operand layout
//
// Operand layout (both A and B): K-major, no swizzle, "core-matrix panels": a panel holds 8
// consecutive K elements for every row, 16 bytes per row, rows contiguous (SBO = 128 B between
// 8-row groups, LBO = rows*16 B between the two K-halves of one K=16 MMA).  An epilogue thread
// therefore stores 16 B at panel*2048 + row*16: consecutive rows -> consecutive addresses, no
// bank conflicts, and the host packs weights into the identical image so a chunk is ONE linear
// bulk copy (no tensor map needed).
//
// Reference arithmetic: models/nerfw.py:297-354 (NeRFW.forward), :105-133 (embedding),
// models/rendering.py:287,305 (pts).  Per-ray constant inputs (view-direction encoding,
// appearance and transient codes) enter as a per-ray bias computed by k_raybias.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <type_traits>

#include <cuda.h>
#include <cudaTypedefs.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace dfb {
namespace tc {

constexpr int kTileM = 128;
constexpr int kStages = 4;
constexpr int kChunkBytes = 16384;
constexpr int kPanelBytes = kTileM * 16;  // 2048
constexpr int kHPanels = 32;              // 256 hidden columns
constexpr int kPePanels = 8;              // 64 positional-encoding columns (63 + zero pad)
constexpr int kSlotBytes = (kHPanels + kPePanels) * kPanelBytes;  // 81920
constexpr int kSmemA = 2 * kSlotBytes;
constexpr int kSmemW = kStages * kChunkBytes;
constexpr int kSmemBar = 256;
constexpr int kSmemTotal = kSmemA + kSmemW + kSmemBar;  // 229632 B
constexpr int kThreads = 448;
constexpr int kMaxSteps = 16;   // epilogue steps (layers) of a program
constexpr int kMaxSub = 32;     // MMA sub-steps: the split-precision coarse pass runs three per layer
constexpr int kTblScal = kMaxSteps * 256;
constexpr int kTblFloats = kTblScal + 16;

// EPI_SIGMA / EPI_HEADS: the output heads run on the tensor pipe as N=64 steps (sigma on the trunk output; the
// transient heads on transient_encoding.6 and static_rgb on dir_encoding as ONE block-structured K=256 step);
// their epilogues read a handful of accumulator columns.  As fp32 dot products in the epilogue they were 60 %
// of its instructions, and the epilogue warps - not the tensor pipe - bounded the kernel.
enum EpiKind { EPI_HIDDEN, EPI_FINAL, EPI_T, EPI_DT, EPI_SIGMA, EPI_HEADS };

// barrier slots (8 bytes each) inside the kSmemBar region
enum Bar { W_FULL = 0, W_EMPTY = 4, D_FULL = 8, A_READY = 10, PASS_DONE = 12, PE_READY = 14, PE_FREE = 16, W_FULLP = 18, N_BARS = 22 };

struct Step {
  int n_chunks;    // 16 KB weight chunks streamed for this step
  int ksteps;      // K=16 MMA steps issued per chunk (the native 128-wide program leaves the tail of some chunks unused)
  int n;           // MMA N (256 or 128)
  int a_panel0;    // first A-operand panel of the step
  int chunk_base;  // index of the step's first chunk in the packed weight image
};

struct TcArgs {
  // cta_group::2 only: 3-D tensor map over the weight image ([images][64 rows][256 B]); 2-SM TMA loads let
  // BOTH CTAs' halves of a stage complete on the LEADER's mbarrier (no relay hop)
  alignas(64) CUtensorMap tmap;
  // MMA program: n_steps sub-steps.  A layer is one sub-step (first = last = 1) or, in the split-precision coarse
  // pass (X3), three that accumulate into the same TMEM columns: A_hi W_hi (first), A_lo W_hi, A_hi W_lo (last).
  // `first`: wait for the layer's A operand, start a fresh accumulator; `last`: publish the accumulator (D_FULL).
  Step steps[kMaxSub];
  uint8_t first[kMaxSub], last[kMaxSub];
  int n_steps;
  int n_epi;              // epilogue steps (= layers)
  int kind[kMaxSteps];    // EpiKind of every epilogue step (EPI_DT = head + store halves)
  int last_pe_step;       // last SUB-step that reads the PE panels (skip layer)
  // network width in 64-column units: 4 = the 8x256 program (also every narrower network embedded in it), 2 = the native
  // 8x128 program (the reference's default netwidth, models/options.py:31): 128-column trunk steps, 64-column transient
  // steps, positional encoding in panels 16..23 so that [h | PE] stays one contiguous K range, panels 24..31 zero
  int hb;
  int pe_panel0;          // first A panel of the positional encoding (32, or 16 in the native 128-wide program)
  const void* wimg;       // packed 16-bit weight image, chunk i at wimg + i*16 KB
  const float* bias32;    // X3 only: fp32 biases in global memory, [epilogue step][256]
  const float* rayrec;    // [n_rays,12]
  const float* z;         // [n_rays,S]
  const float* raybias;   // [n_rays,256] (fine) or null
  int S;
  int64_t P;              // n_rays * S
  int64_t n_pass;         // ceil(tiles / 2)
  float* raw;             // [P,1] or [P,9]
  // FULL == 3 (test-time fine pass, fused compositing): instead of raw, every warp of the heads epilogue composites
  // the ray segments inside its 32 rows and writes one record per segment, part[(ray * part_k + k) * 8] =
  // {prod(1-alpha), prod(1-alpha_static), rgb[3], acc, static depth, 0}, k = index of the warp within the ray
  float* part;
  int part_k;
  // FULL == 4: FULL == 3 on a COMPACTED sample list (opt-in early ray termination, DfbRenderCfg::ert_eps): row g of the
  // tiles is sample rowmap[g] of the flattened [ray][sample] array, ray r owns rows [offsets[r], offsets[r+1]) (its first
  // n_live samples), *P_dev = offsets[n_rays] is the number of live rows (known on the device only: no host round trip)
  const int* rowmap;
  const int* offsets;
  const int* P_dev;
  int* error_flag;
  unsigned long long* prof;  // optional [gridDim.x][16] cycle counters (DFB_TC_PROF builds)
  // FULL == 2 (training forward): ReLU masks of the 12 hidden layers, one bit per activation, for the tcgen05
  // backward (mlp_tc_bwd.cu), which then skips its forward recompute.  [tile][12 layers][8 words][128 rows]
  // uint32, tile = 128 consecutive rows of the flattened [ray][sample] array; bit 16*(c&1) + (c>>1) of word c/32
  // = column c.  mlayer[s] = mask layer written by step s (-1: none).
  uint32_t* masks;
  int mlayer[kMaxSteps];
  // fp32 biases, read through the constant bank (warp-uniform addresses): bias[s][256] for every step, then
  // the head scalars {sigma_b, rgb_b[3], trgb_b[3], tsig_b, tbeta_b}
  float tbl[kTblFloats];
  // the same per-step biases as packed 16-bit pairs (fp16 or bf16, matching the MMA kind) for the
  // packed-math epilogue of the plain hidden layers: btbl[s*128 + j] = {bias[2j], bias[2j+1]}
  uint32_t btbl[kMaxSteps * 128];
};

#ifdef DFB_TC_PROF
#define PROF_DECL unsigned long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; unsigned long long prof_step[16] = {0}; const long long prof_t0 = clock64();
#define PROF_STEP(s, stmt) { const long long _t = clock64(); stmt; prof_step[s] += clock64() - _t; }
#define PROF_FLUSH_STEPS if (a.prof && (threadIdx.x & 31) == 0) { for (int _i = 0; _i < 16; ++_i) a.prof[(size_t)(256 + blockIdx.x) * 16 + _i] = prof_step[_i]; }
#define PROF_WAIT(slot, stmt) { const long long _t = clock64(); stmt; prof_acc[slot] += clock64() - _t; }
#define PROF_PTR prof_acc
#define PROF_FLUSH(base)                                                                   \
  if (a.prof && (threadIdx.x & 31) == 0) {                                                 \
    for (int _i = 0; _i < 3; ++_i) a.prof[(size_t)blockIdx.x * 16 + (base) + _i] = prof_acc[_i]; \
    a.prof[(size_t)blockIdx.x * 16 + (base) + 3] = clock64() - prof_t0;                    \
  }
#else
#define PROF_DECL
#define PROF_WAIT(slot, stmt) { stmt; }
#define PROF_STEP(s, stmt) { stmt; }
#define PROF_FLUSH_STEPS
#define PROF_PTR nullptr
#define PROF_FLUSH(base)
#endif


// ---------------------------------------------------------------------------------------
// epilogue building block: one 32-column block of the accumulator row of this thread
// ---------------------------------------------------------------------------------------

struct EpiCtx {
  float sig;
};

// One 32-column block (columns cb*32 .. cb*32+31) of this thread's accumulator row: bias + activation and the
// 16-bit store of the next layer's A operand.
template <typename T> __device__ __forceinline__ uint32_t gt0_mask2(uint32_t pk);
template <> __device__ __forceinline__ uint32_t gt0_mask2<__half>(uint32_t pk) {
  return __hgt2_mask(*reinterpret_cast<__half2*>(&pk), __floats2half2_rn(0.f, 0.f));
}
template <> __device__ __forceinline__ uint32_t gt0_mask2<__nv_bfloat16>(uint32_t pk) {
  return __hgt2_mask(*reinterpret_cast<__nv_bfloat162*>(&pk), __floats2bfloat162_rn(0.f, 0.f));
}

// mrow (MK only): this row's mask words of the step's layer, word cb at mrow[cb * 128]
template <typename T, int KIND, bool MK>
__device__ __forceinline__ void epi_block(const uint32_t (&v)[32], const TcArgs& a, int bias_off, int cb,
                                          const float* __restrict__ rb, uint32_t h_row, uint32_t* mrow) {
  if (KIND == EPI_HIDDEN || KIND == EPI_T || KIND == EPI_FINAL) {
    // Plain hidden layer: round the fp32 accumulators to the 16-bit operand type first, then
    // bias + ReLU as ONE packed HFMA2.RELU per column pair (the rounding this adds is of the same
    // size as the operand rounding the next MMA applies anyway).
    const int off = (bias_off >> 1) + cb * 16;
    uint32_t pk[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const uint32_t xx = pack2<T>(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1]));
      pk[q] = KIND == EPI_FINAL ? add2<T>(xx, a.btbl[off + q]) : add_relu2<T>(xx, a.btbl[off + q]);
    }
    if (MK && KIND != EPI_FINAL && mrow) {
      uint32_t m = 0;
#pragma unroll
      for (int q = 0; q < 16; ++q) m |= gt0_mask2<T>(pk[q]) & (0x00010001u << q);
      mrow[cb * 128] = m;
    }
    const uint32_t dst = h_row + (uint32_t)(cb * 4) * kPanelBytes;
#pragma unroll
    for (int q = 0; q < 4; ++q) st_shared_v4(dst + q * kPanelBytes, pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
    return;
  }
  // EPI_DT: dir_encoding | transient_encoding.0.  The per-ray bias (view direction, appearance / transient codes
  // and the step's constant bias, k_raybias) arrives as packed 16-bit pairs in global memory: same packed
  // HFMA2.RELU as above.  transient_encoding.0 (columns 128..255) goes to panels 0..15 (input of
  // transient_encoding.2), dir_encoding (columns 0..127) to panels 16..31 (input of static_rgb).
  const uint4* b4 = reinterpret_cast<const uint4*>(reinterpret_cast<const uint32_t*>(rb) + cb * 16);
  uint32_t pk[16];
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4) {
    const uint4 bb = __ldg(b4 + q4);
    const uint32_t bw[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int q = 4 * q4 + e;
      pk[q] = add_relu2<T>(pack2<T>(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1])), bw[e]);
    }
  }
  if (MK && mrow) {
    uint32_t m = 0;
#pragma unroll
    for (int q = 0; q < 16; ++q) m |= gt0_mask2<T>(pk[q]) & (0x00010001u << q);
    mrow[cb * 128] = m;
  }
  const uint32_t dst = h_row + (uint32_t)(((cb + a.hb) & (2 * a.hb - 1)) * 4) * kPanelBytes;
#pragma unroll
  for (int q = 0; q < 4; ++q) st_shared_v4(dst + q * kPanelBytes, pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
}

// Blocks [cb0, cb1) of a step (cb1 - cb0 even): software-pipelined TMEM reads, block cb+1 is in
// flight while block cb is processed.  A runtime loop (two blocks per trip) keeps the code small
// enough for the instruction cache, which the MMA issuer shares.
template <typename T, int KIND, bool MK>
__device__ __forceinline__ void epi_blocks(uint32_t t_row, uint32_t h_row, const TcArgs& a, int bias_off, int cb0, int cb1,
                                           const float* rb, uint32_t* mrow) {
  if (cb0 >= cb1) return;
  uint32_t v0[32], v1[32];
  tmem_ld32(t_row + cb0 * 32, v0);
#pragma unroll 1
  for (int cb = cb0; cb < cb1; cb += 2) {
    tmem_ld_wait(v0);
    tmem_ld32(t_row + (cb + 1) * 32, v1);
    epi_block<T, KIND, MK>(v0, a, bias_off, cb, rb, h_row, mrow);
    tmem_ld_wait(v1);
    if (cb + 2 < cb1) tmem_ld32(t_row + (cb + 2) * 32, v0);
    epi_block<T, KIND, MK>(v1, a, bias_off, cb + 1, rb, h_row, mrow);
  }
}

// Fused test-time compositing of the fine pass (models/rendering.py:132-243 with test_time and static_only, i.e.
// what k_composite_fine_tt computes from the raw tensor): called by the 32 lanes of a heads-epilogue warp, lane = row g
// of the flattened [ray][sample] array.  Ray segments inside the warp are composited with segmented warp scans
// (transmittance = running product of 1 - alpha inside the segment) and the last lane of every segment writes the
// segment's record; k_composite_partials chains the <= ceil(S/32) + 1 records of a ray.  fp32 throughout: nothing
// downstream decides a sample index, and the tensor-core path is gated at 1e-3.
template <bool ERT>
__device__ __forceinline__ void fused_composite(const TcArgs& a, int64_t g, int64_t gfull, int ray, int64_t P, float ss, float st,
                                                const float (&cs)[3], const float (&ct)[3]) {
  // g: row of the tile grid (compacted when ERT), gfull: index of the sample in [ray][sample] order
  const int lane = threadIdx.x & 31;
  const bool valid = g < P;
  const int i = (int)(gfull - (int64_t)ray * a.S);
  const float zi = valid ? __ldg(a.z + gfull) : 0.f;
  const int rnext = __shfl_down_sync(0xffffffffu, ray, 1);
  float zn = __shfl_down_sync(0xffffffffu, zi, 1);
  // the next lane holds the next sample of the same ray, except at the end of the warp and (ERT) at a ray's last live
  // sample, whose successor was not evaluated but still has a depth
  // (rows past the end of the list are clamped onto the last ray: g + 1 >= P marks the last real row)
  if ((lane == 31 || rnext != ray || g + 1 >= P) && valid && i + 1 < a.S) zn = __ldg(a.z + gfull + 1);
  const float delta = (i + 1 < a.S) ? __fsub_rn(zn, zi) : 1e2f;
  float al = 0.f, als = 0.f, alt = 0.f, oma = 1.f, omas = 1.f;
  if (valid) {
    al = __fsub_rn(1.f, expf(__fmul_rn(-delta, __fadd_rn(ss, st))));
    als = __fsub_rn(1.f, expf(__fmul_rn(-delta, ss)));
    alt = __fsub_rn(1.f, expf(__fmul_rn(-delta, st)));
    oma = __fsub_rn(1.f, al), omas = __fsub_rn(1.f, als);
  }
  // inclusive segmented products
  float pa = oma, ps = omas;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const float ta = __shfl_up_sync(0xffffffffu, pa, off), ts = __shfl_up_sync(0xffffffffu, ps, off);
    const int rr = __shfl_up_sync(0xffffffffu, ray, off);
    if (lane >= off && rr == ray) pa *= ta, ps *= ts;
  }
  const int rprev = __shfl_up_sync(0xffffffffu, ray, 1);
  const float ta1 = __shfl_up_sync(0xffffffffu, pa, 1), ts1 = __shfl_up_sync(0xffffffffu, ps, 1);
  const bool cont = lane >= 1 && rprev == ray;
  const float Te = cont ? ta1 : 1.f, Tse = cont ? ts1 : 1.f;   // exclusive: transmittance in front of this sample
  const float sw = __fmul_rn(als, Te), tw = __fmul_rn(alt, Te);
  float acc[5];
#pragma unroll
  for (int c = 0; c < 3; ++c) acc[c] = __fadd_rn(__fmul_rn(sw, cs[c]), __fmul_rn(tw, ct[c]));
  acc[3] = __fmul_rn(al, Te);
  acc[4] = __fmul_rn(__fmul_rn(als, Tse), zi);
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int rr = __shfl_up_sync(0xffffffffu, ray, off);
    const bool take = lane >= off && rr == ray;
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      const float t = __shfl_up_sync(0xffffffffu, acc[c], off);
      if (take) acc[c] += t;
    }
  }
  const int64_t ray0 = ERT ? (int64_t)__ldg(a.offsets + ray) : (int64_t)ray * a.S;   // first row of the ray
  const int64_t seg_first = ray0 > (g & ~(int64_t)31) ? ray0 : (g & ~(int64_t)31);
  if ((lane == 31 || rnext != ray) && seg_first < P) {
    const int k = (int)((g >> 5) - (ray0 >> 5));
    float4* o = reinterpret_cast<float4*>(a.part + ((size_t)ray * a.part_k + k) * 8);
    o[0] = make_float4(pa, ps, acc[0], acc[1]);
    o[1] = make_float4(acc[2], acc[3], acc[4], 0.f);
  }
}

// Split-precision (X3) hidden layer: bias + ReLU in fp32, then the activation is stored as TWO fp16 operands,
// hi = rn16(x) in the slot's panels and lo = rn16(x - hi) in the same panel of the "lo" region (kSlotBytes further):
// hi + lo carries 22 mantissa bits, and the next layer accumulates A_hi W_hi + A_lo W_hi + A_hi W_lo in fp32.
__device__ __forceinline__ void epi_block_x3(const uint32_t (&v)[32], const float* __restrict__ bias, int cb, uint32_t h_row) {
  uint32_t hi[16], lo[16];
  const float4* b4 = reinterpret_cast<const float4*>(bias + cb * 32);
#pragma unroll
  for (int q4 = 0; q4 < 8; ++q4) {
    const float4 bb = __ldg(b4 + q4);
    const float bw[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int q = 2 * q4 + e;
      const float x0 = fmaxf(__uint_as_float(v[2 * q]) + bw[2 * e], 0.f);
      const float x1 = fmaxf(__uint_as_float(v[2 * q + 1]) + bw[2 * e + 1], 0.f);
      const __half2 h = __floats2half2_rn(x0, x1);
      const float2 hf = __half22float2(h);
      const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
      hi[q] = *reinterpret_cast<const uint32_t*>(&h);
      lo[q] = *reinterpret_cast<const uint32_t*>(&l);
    }
  }
  const uint32_t dst = h_row + (uint32_t)(cb * 4) * kPanelBytes;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    st_shared_v4(dst + q * kPanelBytes, hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
    st_shared_v4(dst + kSlotBytes + q * kPanelBytes, lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
  }
}

__device__ __forceinline__ void epi_blocks_x3(uint32_t t_row, uint32_t h_row, const float* bias, int cb0, int cb1) {
  uint32_t v0[32], v1[32];
  tmem_ld32(t_row + cb0 * 32, v0);
#pragma unroll 1
  for (int cb = cb0; cb < cb1; cb += 2) {
    tmem_ld_wait(v0);
    tmem_ld32(t_row + (cb + 1) * 32, v1);
    epi_block_x3(v0, bias, cb, h_row);
    tmem_ld_wait(v1);
    if (cb + 2 < cb1) tmem_ld32(t_row + (cb + 2) * 32, v0);
    epi_block_x3(v1, bias, cb + 1, h_row);
  }
}

// MMA issue for one (step, slot): KS K=16 MMAs per weight stage.  Kept as lean as possible:
// this single warp paces the tensor pipe.
template <int CG, int KS>
__device__ __forceinline__ void issue_step(uint32_t& stage, uint32_t& phase, int nch, uint32_t a_lo, uint32_t b_rows,
                                           uint32_t d_tmem, uint32_t idesc, uint32_t sW, uint32_t sBar, int* err,
                                           unsigned long long* prof_acc, uint32_t acc = 0, int kcap = KS) {
  // acc = 1: the sub-step continues the accumulator of the previous one (split-precision coarse pass)
  // kcap (CG = 2 only): K=16 steps issued per chunk
  const uint32_t desc_hi = (128u >> 4) | (1u << 14);  // SBO 128 B, descriptor version 1, no swizzle
  const uint32_t b_step = 2u * b_rows;                // (2 panels * b_rows * 16 B) >> 4
  if (CG == 1) {
    // Stages are filled and waited for in PAIRS (one "full" barrier per two 16 KB stages, see the
    // producer) but released one by one: half the mbarrier waits / elect blocks per MMA.
#pragma unroll 1
    for (int c = 0; c < nch; c += 2) {
      PROF_WAIT(0, mbar_wait(sBar + 8u * (W_FULL + stage), phase, err));
      tc_fence_after();
      const uint32_t b_lo = ((sW + stage * kChunkBytes) >> 4) | (b_rows << 16);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
          umma_f16<CG>(d_tmem, mk64(a_lo + ks * 256, desc_hi), mk64(b_lo + ks * b_step, desc_hi), idesc, acc | ks);
        umma_commit<CG>(sBar + 8u * (W_EMPTY + stage));
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
          umma_f16<CG>(d_tmem, mk64(a_lo + (KS + ks) * 256, desc_hi),
                       mk64(b_lo + (kChunkBytes >> 4) + ks * b_step, desc_hi), idesc, 1u);
        umma_commit<CG>(sBar + 8u * (W_EMPTY + stage + 1));
      }
      __syncwarp();
      a_lo += 512u * KS;
      acc = 1;
      stage += 2;
      if (stage == kStages) stage = 0, phase ^= 1;
    }
    return;
  }
#pragma unroll 1
  for (int c = 0; c < nch; ++c) {
    PROF_WAIT(0, mbar_wait_cluster<CG>(sBar + 8u * (W_FULL + stage), phase, err));  // both CTAs' halves landed
    tc_fence_after();
    const uint32_t b_lo = ((sW + stage * kChunkBytes) >> 4) | (b_rows << 16);
    if (elect_one()) {
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
        if (ks < kcap) umma_f16<CG>(d_tmem, mk64(a_lo + ks * 256, desc_hi), mk64(b_lo + ks * b_step, desc_hi), idesc, acc | ks);
      umma_commit<CG>(sBar + 8u * (W_EMPTY + stage));
    }
    __syncwarp();
    a_lo += 256u * KS;
    acc = 1;
    if (++stage == kStages) stage = 0, phase ^= 1;
  }
}

// Kernel body, shared by the 1-CTA (CG = 1) and CTA-pair (CG = 2, cta_group::2) variants.
//
// CG = 2: two CTAs of a cluster work on one 256-row tile per slot.  Each CTA keeps its own 128
// rows (A operand, TMEM accumulators, epilogue) and HALF of every weight chunk (its N/2 rows of
// B); the leader CTA issues tcgen05.mma.cta_group::2 for the pair, so one instruction feeds both
// SMs' tensor cores, every weight byte is fetched from L2 once per 256 rows, and the MMA issue
// overhead per row halves.  Cross-CTA signalling: tcgen05.commit multicasts to the barriers of
// both CTAs; epilogue / encoder threads of the peer arrive remotely on the leader's barriers
// (mapa + mbarrier.arrive.release.cluster); weight stages are filled by 2-SM TMA loads
// (cp.async.bulk.tensor...cta_group::2) that credit both CTAs' bytes to the leader's "full" barrier.
//
// X3 (coarse sigma-only pass, fp16): split-precision operands.  ONE 128-row slot per CTA whose activations are kept
// as hi + lo fp16 pairs (the "lo" copy lives where slot 1 would be), three MMA sub-steps per layer (see TcArgs) and an
// fp32 bias/ReLU epilogue: the products carry ~22 mantissa bits, i.e. the pass that decides WHERE the fine samples go
// is as exact as the fp32 kernels at 3x the tensor work of the fp16 pass (mlp kind DFB_MMA_F16_SPLIT_COARSE).
template <typename T, int FULL, int CG, bool X3 = false>
__device__ __forceinline__ void mlp_tc_body(const TcArgs& a) {
  static_assert(!X3 || (FULL == 0 && std::is_same<T, __half>::value), "split precision: fp16 coarse pass only");
  constexpr int NSLOT = X3 ? 1 : 2;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sA = smem_u32(smem);
  const uint32_t sW = sA + kSmemA;
  const uint32_t sBar = sW + kSmemW;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kSmemA + kSmemW + N_BARS * 8);
  const int tid = threadIdx.x, warp = tid >> 5;
  auto bar = [&](int i) { return sBar + 8u * i; };
  const int fmt = std::is_same<T, __nv_bfloat16>::value ? 1 : 0;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  const int64_t unit0 = blockIdx.x / CG, n_units = gridDim.x / CG;  // a unit = CTA (CG 1) or CTA pair (CG 2)

  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(bar(W_FULL + i), 1), mbar_init(bar(W_EMPTY + i), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar(D_FULL + s), 1);
      // ONE arrival per warp (lane 0, after __syncwarp): with one arrival per thread the 512 remote release-arrivals per
      // (layer, slot) were 38-46 % of the epilogue warps' time (tools/tc_prof.py)
      mbar_init(bar(A_READY + s), 8 * CG);     // both epilogue warpgroups arrive for every slot
      mbar_init(bar(PASS_DONE + s), 8 * CG);
      mbar_init(bar(PE_READY + s), 4 * CG);
      mbar_init(bar(PE_FREE + s), 1);
    }
    fence_barrier_init();
  }
  if (warp == 13) tmem_alloc<CG>(smem_u32(tmem_slot), 512);
  if (a.hb == 2) {
    // native 128-wide program: panels 24..31 of both slots are read by the K = 128 / 256 ranges of layer 0, the skip layer
    // and the sigma step against zero weights; they must hold finite values (0 x NaN is NaN)
    for (int i = tid; i < 2 * 8 * (kPanelBytes / 16); i += kThreads) {
      const int sl = i / (8 * (kPanelBytes / 16)), o = i % (8 * (kPanelBytes / 16));
      st_shared_v4(sA + sl * kSlotBytes + 24 * kPanelBytes + o * 16, 0u, 0u, 0u, 0u);
    }
    fence_proxy_async();
  }
  tc_fence_before();
  if (CG == 2) cluster_sync(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_steps = a.n_steps;
  constexpr bool ERT = FULL == 4;
  const int64_t P = ERT ? (int64_t)__ldg(a.P_dev) : a.P;
  const int64_t n_pass = ERT ? ((P + kTileM - 1) / kTileM + NSLOT * CG - 1) / (NSLOT * CG) : a.n_pass;

  if (warp == 12) {
    // ===== weight producer (warp-uniform control flow, one elected lane issues) ============
    PROF_DECL
    uint32_t stage = 0, phase = 0;
    const uint8_t* wimg = reinterpret_cast<const uint8_t*>(a.wimg) + (size_t)rank * kChunkBytes;
    for (int64_t p = unit0; p < n_pass; p += n_units)
      for (int s = 0; s < n_steps; ++s) {
        const int nch = a.steps[s].n_chunks;
        const uint8_t* src0 = wimg + (size_t)a.steps[s].chunk_base * (kChunkBytes * CG);
        const int img0 = a.steps[s].chunk_base * CG + (int)rank;  // image index of chunk 0 for this CTA (CG 2)
        for (int slot = 0; slot < NSLOT; ++slot) {
          const uint8_t* src = src0;
          for (int c = 0; c < nch; ++c, src += kChunkBytes * CG) {
            PROF_WAIT(0, mbar_wait(bar(W_EMPTY + stage), phase ^ 1, a.error_flag));
            if (elect_one()) {
              if (CG == 1) {
                // one "full" barrier per stage pair: armed with both stages' bytes by the even stage
                if ((stage & 1) == 0) mbar_expect_tx(bar(W_FULL + stage), 2 * kChunkBytes);
                bulk_g2s(sW + stage * kChunkBytes, src, kChunkBytes, bar(W_FULL + (stage & ~1u)));
              } else {
                // CTA pair: the leader arms ITS barrier with both halves' bytes; each CTA loads its own half
                // with a 2-SM TMA whose completion is credited to the leader's barrier
                const uint32_t bar_leader = bar(W_FULL + stage) & 0xFEFFFFFFu;  // same offset in cluster rank 0
                if (rank == 0) mbar_expect_tx(bar(W_FULL + stage), 2 * kChunkBytes);
                tma_load_img_2sm(sW + stage * kChunkBytes, &a.tmap, img0 + c * CG, bar_leader);
              }
            }
            __syncwarp();
            if (++stage == kStages) stage = 0, phase ^= 1;
          }
        }
      }
    PROF_FLUSH(0)
  } else if (warp == 13 && rank != 0) {
    // peer CTA of a pair: the leader issues every MMA
  } else if (warp == 13) {
    // ===== MMA issuer (warp-uniform control flow, one elected lane issues) ===================
    PROF_DECL
    uint32_t stage = 0, phase = 0;
    int lp = 0;
    const int n_epi = a.n_epi;
    for (int64_t p = unit0; p < n_pass; p += n_units, ++lp) {
      int e = 0;  // epilogue step (layer) the sub-step belongs to
      for (int s = 0; s < n_steps; ++s) {
        const int nch = a.steps[s].n_chunks, nn = a.steps[s].n;
        const uint32_t idesc = make_idesc(fmt, nn, kTileM * CG);
        const uint32_t b_rows = nn / CG;  // B rows held by each CTA; LBO = b_rows * 16 B
        const bool first = a.first[s] != 0, last = a.last[s] != 0;
        for (int slot = 0; slot < NSLOT; ++slot) {
          if (first) {
            if (e == 0) {
              if (lp > 0) PROF_WAIT(2, mbar_wait_cluster<CG>(bar(PASS_DONE + slot), (lp - 1) & 1, a.error_flag));
              PROF_WAIT(2, mbar_wait_cluster<CG>(bar(PE_READY + slot), lp & 1, a.error_flag));
            } else {
              PROF_STEP(e, mbar_wait_cluster<CG>(bar(A_READY + slot), (lp * (n_epi - 1) + e - 1) & 1, a.error_flag));
            }
            tc_fence_after();
          }
          const uint32_t d_tmem = tmem_base + slot * 256;
          // low word: start address >> 4 | LBO (2048 B >> 4) << 16; one K=16 step advances by 2 panels
          const uint32_t a_lo = ((sA + slot * kSlotBytes + a.steps[s].a_panel0 * kPanelBytes) >> 4) | ((kPanelBytes >> 4) << 16);
          const uint32_t acc0 = first ? 0u : 1u;
          const int kcap = a.steps[s].ksteps;
          if (nn == 256) issue_step<CG, 2 * CG>(stage, phase, nch, a_lo, b_rows, d_tmem, idesc, sW, sBar, a.error_flag, PROF_PTR, acc0, kcap);
          else if (nn == 128) issue_step<CG, 4 * CG>(stage, phase, nch, a_lo, b_rows, d_tmem, idesc, sW, sBar, a.error_flag, PROF_PTR, acc0, kcap);
          else issue_step<CG, 8 * CG>(stage, phase, nch, a_lo, b_rows, d_tmem, idesc, sW, sBar, a.error_flag, PROF_PTR, acc0, kcap);
          if (elect_one()) {
            if (last) umma_commit<CG>(bar(D_FULL + slot));
            if (s == a.last_pe_step) umma_commit<CG>(bar(PE_FREE + slot));
          }
          __syncwarp();
        }
        if (last) ++e;
      }
    }
    PROF_FLUSH(4)
    PROF_FLUSH_STEPS
  } else if (warp >= 8 && warp < 12) {
    // ===== encoder: positional encoding of the next pass (nerfw.py:128-133) =================
    const int r = tid - 256;
    int lp = 0;
    for (int64_t p = unit0; p < n_pass; p += n_units, ++lp)
      for (int slot = 0; slot < NSLOT; ++slot) {
        if (lp > 0) mbar_wait_relaxed(bar(PE_FREE + slot), (lp - 1) & 1, a.error_flag);
        int64_t g = ((NSLOT * p + slot) * CG + rank) * kTileM + r;
        g = g < P ? g : P - 1;
        if (ERT) g = __ldg(a.rowmap + g);
        const int64_t ray = g / a.S;
        const float* rr = a.rayrec + ray * kRayRec;
        const float zz = __ldg(a.z + g);
        float pt[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) pt[c] = __fadd_rn(__ldg(rr + c), __fmul_rn(__ldg(rr + 3 + c), zz));
        // column c of the encoding lives at panel c/8, byte (c%8)*2 of this row's 16-byte slot
        const uint32_t dst = sA + slot * kSlotBytes + a.pe_panel0 * kPanelBytes + r * 16;
        auto put = [&](int col, float v) {
          T h = (T)v;
          const uint32_t at = dst + (uint32_t)(col >> 3) * kPanelBytes + (col & 7) * 2;
          st_shared_b16(at, *reinterpret_cast<uint16_t*>(&h));
          if (X3) {  // lo part of the split operand, in the "lo" region of the slot
            T l = (T)(v - (float)h);
            st_shared_b16(at + kSlotBytes, *reinterpret_cast<uint16_t*>(&l));
          }
        };
        put(0, pt[0]), put(1, pt[1]), put(2, pt[2]), put(63, 0.f);
#pragma unroll 1
        for (int l = 0; l < 10; ++l) {
          const float fr = (float)(1 << l);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float sn, cs;
            sincosf(__fmul_rn(pt[c], fr), &sn, &cs);
            put(3 + 6 * l + c, sn);
            put(3 + 6 * l + 3 + c, cs);
          }
        }
        fence_proxy_async();
        __syncwarp();
        if ((tid & 31) == 0) arrive_leader<CG>(bar(PE_READY + slot));
      }
  } else if (warp < 8) {
    // ===== epilogue: BOTH warpgroups work on every (slot, step) ===================================
    // thread = accumulator row (sample) r of BOTH slots; warpgroup `wg` owns one half of the step's
    // columns (warps w and w+4 reach the same 32 TMEM lanes).  The two slots' epilogues alternate in
    // time anyway (each overlaps the other slot's MMAs), so splitting every epilogue over all eight
    // warps halves the MMA -> epilogue -> MMA chain of a slot and gives every scheduler two working
    // warps instead of one.  The head steps (EPI_SIGMA, EPI_HEADS) only read a few accumulator columns:
    // warpgroup 0 finishes them, warpgroup 1 just arrives.
    const int wg = warp >> 2;
    const int r = tid & 127;
    EpiCtx cx[2];
    PROF_DECL
    uint32_t nd = 0;
    for (int64_t p = unit0; p < n_pass; p += n_units) {
      // row of slot `sl` in the flattened [ray][sample] array (recomputed where needed: registers are scarce here)
      auto row_of = [&](int sl) { return ((NSLOT * p + sl) * CG + rank) * kTileM + r; };
      int rayi[2], gfi[2];
#pragma unroll
      for (int slot = 0; slot < NSLOT; ++slot) {
        const int64_t gg = row_of(slot);
        const int64_t gc = gg < P ? gg : P - 1;
        gfi[slot] = ERT ? __ldg(a.rowmap + gc) : (int)gc;   // sample index in [ray][sample] order (< 2^31: the chunk bound)
        rayi[slot] = gfi[slot] / a.S;
        // the per-ray bias row (1 KB) is read by the dir/transient layer much later in the pass: pull it
        // into L1 now so that those loads do not pay eight serial L2 round trips
        if (FULL) asm volatile("prefetch.global.L1 [%0];" ::"l"(a.raybias + (size_t)rayi[slot] * 256 + (tid & 7) * 32));
        cx[slot].sig = 0.f;
      }
      const int n_epi = a.n_epi;
      for (int s = 0; s < n_epi; ++s, ++nd) {
        const int boff = s * 256;
        const int kd = a.kind[s];
#pragma unroll
        for (int slot = 0; slot < NSLOT; ++slot) {
          const uint32_t t_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + slot * 256;
          const uint32_t h_row = sA + slot * kSlotBytes + r * 16;
          const float* rbs = FULL ? a.raybias + (size_t)rayi[slot] * 256 : nullptr;
          PROF_WAIT(0, mbar_wait(bar(D_FULL + slot), nd & 1, a.error_flag));
          tc_fence_after();
          // first 32-column block of this warpgroup: full-width steps (hb blocks each) and half-width steps (hb / 2 each;
          // in the native 128-wide program a half-width step is two blocks, both done by warpgroup 0)
          const int hb = a.hb, w0 = hb * wg, n0 = hb == 4 ? 2 * wg : 0, n1 = hb == 4 ? n0 + 2 : (wg == 0 ? 2 : 0);
          constexpr bool MK = FULL == 2;
          uint32_t* mrow = nullptr;
          // (tiles past the end of the sample array exist when the tile count is not a multiple of the tiles per
          // pass: they are computed on clamped rows but must not write masks — the buffer has ceil(P/128) tiles)
          if (MK && ((2 * p + slot) * CG + rank) * kTileM < P)
            mrow = a.masks + ((((2 * p + slot) * CG + rank) * 12 + a.mlayer[s]) * 8) * 128 + r;
          if (X3 && kd == EPI_HIDDEN) epi_blocks_x3(t_row, h_row, a.bias32 + boff, w0, w0 + 4);
          else if (kd == EPI_HIDDEN) epi_blocks<T, EPI_HIDDEN, MK>(t_row, h_row, a, boff, w0, w0 + hb, rbs, mrow);
          else if (kd == EPI_T) epi_blocks<T, EPI_T, MK>(t_row, h_row, a, boff, n0, n1, rbs, mrow);
          else if (kd == EPI_DT) epi_blocks<T, EPI_DT, MK>(t_row, h_row, a, boff, w0, w0 + hb, rbs, mrow);
          else if (kd == EPI_FINAL) epi_blocks<T, EPI_FINAL, false>(t_row, h_row, a, boff, w0, w0 + hb, rbs, mrow);
          else if (wg == 0) {
            // head steps: column 0 = sigma (EPI_SIGMA); columns 0..4 = transient rgb(3), sigma, beta and
            // columns 8..10 = static rgb (EPI_HEADS)
            uint32_t v[32];
            tmem_ld32(t_row, v);
            tmem_ld_wait(v);
            const int64_t gg = row_of(slot);
            if (FULL >= 3 && kd == EPI_HEADS) {
              // fused compositing: the accumulator has been read, so the slot is handed back to the tensor pipe
              // first and the activation / compositing arithmetic overlaps the next pass' first layers
              tc_fence_before();
              __syncwarp();
              if ((tid & 31) == 0) arrive_leader<CG>(bar(PASS_DONE + slot));
              const float cs[3] = {sigmoid_f(__uint_as_float(v[8]) + a.tbl[kTblScal + 1]),
                                   sigmoid_f(__uint_as_float(v[9]) + a.tbl[kTblScal + 2]),
                                   sigmoid_f(__uint_as_float(v[10]) + a.tbl[kTblScal + 3])};
              const float ct[3] = {sigmoid_f(__uint_as_float(v[0]) + a.tbl[kTblScal + 4]),
                                   sigmoid_f(__uint_as_float(v[1]) + a.tbl[kTblScal + 5]),
                                   sigmoid_f(__uint_as_float(v[2]) + a.tbl[kTblScal + 6])};
              const float st = softplus_f(__uint_as_float(v[3]) + a.tbl[kTblScal + 7]);
              fused_composite<ERT>(a, gg, ERT ? (int64_t)gfi[slot] : gg, rayi[slot], P, cx[slot].sig, st, cs, ct);
              continue;
            }
            if (kd == EPI_SIGMA) {
              cx[slot].sig = softplus_f(__uint_as_float(v[0]) + a.tbl[kTblScal]);
              if (!FULL && gg < P) a.raw[gg] = cx[slot].sig;
            } else if (FULL < 3 && gg < P) {
              float* o = a.raw + gg * 9;
              o[0] = sigmoid_f(__uint_as_float(v[8]) + a.tbl[kTblScal + 1]);
              o[1] = sigmoid_f(__uint_as_float(v[9]) + a.tbl[kTblScal + 2]);
              o[2] = sigmoid_f(__uint_as_float(v[10]) + a.tbl[kTblScal + 3]);
              o[3] = cx[slot].sig;
              o[4] = sigmoid_f(__uint_as_float(v[0]) + a.tbl[kTblScal + 4]);
              o[5] = sigmoid_f(__uint_as_float(v[1]) + a.tbl[kTblScal + 5]);
              o[6] = sigmoid_f(__uint_as_float(v[2]) + a.tbl[kTblScal + 6]);
              o[7] = softplus_f(__uint_as_float(v[3]) + a.tbl[kTblScal + 7]);
              o[8] = softplus_f(__uint_as_float(v[4]) + a.tbl[kTblScal + 8]);
            }
          }
          // every lane orders its own TMEM reads and shared-memory writes, the warp synchronises, lane 0 releases
          PROF_WAIT(1, tc_fence_before(); fence_proxy_async(); __syncwarp());
          PROF_WAIT(2, if ((tid & 31) == 0) arrive_leader<CG>(bar((s + 1 < n_epi ? A_READY : PASS_DONE) + slot)));
        }
      }
    }
    if (warp == 0) { PROF_FLUSH(8) }
    if (warp == 4) { PROF_FLUSH(12) }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync(); else __syncthreads();
  if (warp == 13) {
    tc_fence_after();
    tmem_dealloc<CG>(tmem_base, 512);
  }
}

template <typename T, int FULL>
__global__ void __launch_bounds__(kThreads, 1) k_mlp_tc(const __grid_constant__ TcArgs a) {
  mlp_tc_body<T, FULL, 1>(a);
}

template <typename T, int FULL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) k_mlp_tc2(const __grid_constant__ TcArgs a) {
  mlp_tc_body<T, FULL, 2>(a);
}

// split-precision coarse pass (see mlp_tc_body, X3)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) k_mlp_tc2_x3(const __grid_constant__ TcArgs a) {
  mlp_tc_body<__half, 0, 2, true>(a);
}

// ---------------------------------------------------------------------------------------
// single-tile UMMA self test: D[128,N] = A[128,K] * B[N,K]^T through the same descriptors
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128, 1) k_umma_selftest(const float* A, const float* Bm, int N, int K, int variant,
                                                            float* D, int* error_flag) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  T* sAm = reinterpret_cast<T*>(smem);                       // [K/8 panels][128 rows][8]
  T* sBm = reinterpret_cast<T*>(smem + (size_t)K * 128 * 2);  // [K/8 panels][N rows][8]
  for (int i = tid; i < 128 * K; i += 128) {
    const int row = i / K, k = i % K;
    sAm[(size_t)(k / 8) * 128 * 8 + row * 8 + k % 8] = (T)A[i];
  }
  for (int i = tid; i < N * K; i += 128) {
    const int row = i / K, k = i % K;
    sBm[(size_t)(k / 8) * N * 8 + row * 8 + k % 8] = (T)Bm[i];
  }
  if (tid == 0) { mbar_init(smem_u32(&mbar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&tslot), 256);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tslot;
  const int fmt = std::is_same<T, __nv_bfloat16>::value ? 1 : 0;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(fmt, N, 128);
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint32_t a_addr = smem_u32(sAm) + ks * 2 * 2048, b_addr = smem_u32(sBm) + ks * 2 * (N * 16);
      uint64_t ad, bd;
      if (variant == 0) ad = make_desc(a_addr, 2048, 128), bd = make_desc(b_addr, N * 16, 128);
      else ad = make_desc(a_addr, 128, 2048), bd = make_desc(b_addr, 128, N * 16);
      umma_f16(tb, ad, bd, idesc, ks > 0);
    }
    umma_commit(smem_u32(&mbar));
  }
  mbar_wait(smem_u32(&mbar), 0, error_flag);
  tc_fence_after();
  for (int cb = 0; cb < N / 32; ++cb) {
    uint32_t v[32];
    tmem_ld32(tb + ((uint32_t)(warp * 32) << 16) + cb * 32, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D[(size_t)tid * N + cb * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 256); }
}

// Tensor-pipe rate probe: every CTA issues `iters` x 16 back-to-back MMAs (M=128, N=n, K=16) on resident
// shared-memory operands in the kernel's no-swizzle panel layout and reports cycles per MMA.
__global__ void __launch_bounds__(128, 1) k_umma_rate(int iters, int n, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (65536 + 32768) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) { mbar_init(smem_u32(&mbar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<1>(smem_u32(&tslot), 256);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tslot;
  if (warp == 0) {
    const uint32_t idesc = make_idesc(0, n, 128);
    const uint32_t desc_hi = (128u >> 4) | (1u << 14);
    const uint32_t a_lo = (smem_u32(smem) >> 4) | ((2048u >> 4) << 16);
    const uint32_t b_lo = ((smem_u32(smem) + 65536) >> 4) | ((uint32_t)n << 16);
    const long long t0 = clock64();
    if (elect_one()) {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int ks = 0; ks < 16; ++ks)
          umma_f16<1>(tb, mk64(a_lo + ks * 256, desc_hi), mk64(b_lo + (ks & 1) * 2 * n, desc_hi), idesc, 1u);
      }
      umma_commit<1>(smem_u32(&mbar));
    }
    __syncwarp();
    mbar_wait(smem_u32(&mbar), 0, nullptr);
    const long long t1 = clock64();
    if (tid == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<1>(tb, 256); }
}

// TMEM read-rate probe: `nwarps` warps of one CTA each issue `iters` x 8 tcgen05.ld.32x32b.x32 (4 KB per
// instruction) on their own 32 lanes, two loads in flight; reports cycles per load instruction per warp.
__global__ void __launch_bounds__(128, 1) k_tmem_rate(int iters, int nwarps, unsigned long long* out) {
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc<1>(smem_u32(&tslot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tslot + ((uint32_t)(warp * 32) << 16);
  uint32_t sink = 0;
  if (warp < nwarps) {
    uint32_t v0[32], v1[32];
    const long long t0 = clock64();
    tmem_ld32(tb, v0);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int cb = 0; cb < 8; cb += 2) {
        tmem_ld_wait(v0);
        tmem_ld32(tb + (cb + 1) * 32, v1);
        sink ^= v0[cb];
        tmem_ld_wait(v1);
        tmem_ld32(tb + ((cb + 2) & 7) * 32, v0);
        sink ^= v1[cb];
      }
    }
    tmem_ld_wait(v0);
    const long long t1 = clock64();
    if ((tid & 31) == 0) out[blockIdx.x * 4 + warp] = (unsigned long long)(t1 - t0) + (sink == 0x12345678u ? 1 : 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<1>(tslot, 512); }
}

// TMEM read rate UNDER tensor-pipe load: warp 4 issues back-to-back MMAs (M=128, N=256, K=16) into columns 256..511
// while warps 0..nwarps-1 run the tcgen05.ld loop of k_tmem_rate on columns 0..255.
__global__ void __launch_bounds__(160, 1) k_tmem_rate_mma(int iters, int nwarps, int mma_iters, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (65536 + 32768) / 16; i += 160) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) { mbar_init(smem_u32(&mbar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<1>(smem_u32(&tslot), 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t sink = 0;
  if (warp == 4) {
    const uint32_t idesc = make_idesc(0, 256, 128);
    const uint32_t desc_hi = (128u >> 4) | (1u << 14);
    const uint32_t a_lo = (smem_u32(smem) >> 4) | ((2048u >> 4) << 16);
    const uint32_t b_lo = ((smem_u32(smem) + 65536) >> 4) | (256u << 16);
    const long long t0 = clock64();
    if (elect_one()) {
      for (int it = 0; it < mma_iters; ++it) {
#pragma unroll
        for (int ks = 0; ks < 16; ++ks)
          umma_f16<1>(tslot + 256, mk64(a_lo + ks * 256, desc_hi), mk64(b_lo + (ks & 1) * 512, desc_hi), idesc, 1u);
      }
      umma_commit<1>(smem_u32(&mbar));
    }
    __syncwarp();
    mbar_wait(smem_u32(&mbar), 0, nullptr);
    if ((tid & 31) == 0) out[blockIdx.x * 8 + 4] = (unsigned long long)(clock64() - t0);
  } else if (warp < nwarps) {
    const uint32_t tb = tslot + ((uint32_t)(warp * 32) << 16);
    uint32_t v0[32], v1[32];
    const long long t0 = clock64();
    tmem_ld32(tb, v0);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int cb = 0; cb < 8; cb += 2) {
        tmem_ld_wait(v0);
        tmem_ld32(tb + (cb + 1) * 32, v1);
        sink ^= v0[cb];
        tmem_ld_wait(v1);
        tmem_ld32(tb + ((cb + 2) & 7) * 32, v0);
        sink ^= v1[cb];
      }
    }
    tmem_ld_wait(v0);
    const long long t1 = clock64();
    if ((tid & 31) == 0) out[blockIdx.x * 8 + warp] = (unsigned long long)(t1 - t0) + (sink == 0x12345678u ? 1 : 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<1>(tslot, 512); }
}

}  // namespace tc

// ---------------------------------------------------------------------------------------
// host side: weight image, step tables, launch
// ---------------------------------------------------------------------------------------
namespace {

// 16-bit conversions on the host (round to nearest even), independent of device intrinsics
uint16_t f2h(float f) { __half h = __float2half_rn(f); uint16_t u; memcpy(&u, &h, 2); return u; }
uint16_t f2b(float f) { __nv_bfloat16 h = __float2bfloat16_rn(f); uint16_t u; memcpy(&u, &h, 2); return u; }

}  // namespace

int tc_cta_group_env() {  // read per launch so that tests can exercise both variants in one process
  const char* e = getenv("DFB_TC_CTA_GROUP");
  return (e && e[0] == '1') ? 1 : 2;
}

bool tc_padded_shape(const NetPack& np) {
  return np.D == 8 && np.skip == 4 && np.pek == 64 && np.W >= 32 && np.W <= 256 && np.W % 16 == 0;
}

std::vector<std::vector<float>> tc_pad_params(const NetPack& np, const std::vector<std::vector<float>>& P) {
  const int W = np.W, H = W / 2, W2 = 256, H2 = 128, in_xyz = np.in_xyz;
  if (W == W2) return P;
  std::vector<std::vector<float>> Q(P.size());
  // [rows x cols] -> [rows2 x cols2]; column c of the source goes to column cmap(c)
  auto pad2 = [&](const std::vector<float>& src, int rows, int cols, int rows2, int cols2, auto cmap) {
    std::vector<float> dst((size_t)rows2 * cols2, 0.f);
    for (int r = 0; r < rows; ++r)
      for (int c = 0; c < cols; ++c) dst[(size_t)r * cols2 + cmap(c)] = src[(size_t)r * cols + c];
    return dst;
  };
  auto pad1 = [&](const std::vector<float>& src, int n2) {
    std::vector<float> dst(n2, 0.f);
    std::copy(src.begin(), src.end(), dst.begin());
    return dst;
  };
  auto ident = [](int c) { return c; };
  for (int i = 0; i < 8; ++i) {
    if (i == 0) Q[0] = pad2(P[0], W, in_xyz, W2, in_xyz, ident);
    else if (i == 4) Q[8] = pad2(P[8], W, in_xyz + W, W2, in_xyz + W2, ident);  // cat([input_xyz, h]): h columns stay adjacent
    else Q[2 * i] = pad2(P[2 * i], W, W, W2, W2, ident);
    Q[2 * i + 1] = pad1(P[2 * i + 1], W2);
  }
  Q[16] = pad2(P[16], W, W, W2, W2, ident), Q[17] = pad1(P[17], W2);  // xyz_encoding_final
  const int nd = np.in_dir + np.a_dim;
  // dir_encoding [H, W + in_dir (+ a)]: the ray-constant columns move behind the 256 hidden columns
  Q[18] = pad2(P[18], H, W + nd, H2, W2 + nd, [&](int c) { return c < W ? c : W2 + (c - W); });
  Q[19] = pad1(P[19], H2);
  Q[20] = pad2(P[20], 1, W, 1, W2, ident), Q[21] = P[21];             // static_sigma
  Q[22] = pad2(P[22], 3, H, 3, H2, ident), Q[23] = P[23];             // static_rgb
  if (np.fine) {
    Q[24] = pad2(P[24], H, W + np.t_dim, H2, W2 + np.t_dim, [&](int c) { return c < W ? c : W2 + (c - W); });
    Q[25] = pad1(P[25], H2);
    for (int i = 0; i < 3; ++i) Q[26 + 2 * i] = pad2(P[26 + 2 * i], H, H, H2, H2, ident), Q[27 + 2 * i] = pad1(P[27 + 2 * i], H2);
    Q[32] = pad2(P[32], 1, H, 1, H2, ident), Q[33] = P[33];           // transient_sigma
    Q[34] = pad2(P[34], 3, H, 3, H2, ident), Q[35] = P[35];           // transient_rgb
    Q[36] = pad2(P[36], 1, H, 1, H2, ident), Q[37] = P[37];           // transient_beta
  }
  for (size_t i = 0; i < P.size(); ++i)
    if (Q[i].empty()) Q[i] = P[i];
  return Q;
}

bool tc_supported(const DfbNerf* n, int which, int mode) {
  const NetPack& np = n->net[which];
  if (!np.loaded || !tc_padded_shape(np)) return false;
  if (np.blob16[0][0] == nullptr || np.tc_tbl.empty()) return false;
  if (mode == MLP_SIGMA) return true;
  if (mode == MLP_FULL) return np.fine;
  return false;  // MLP_STATIC (train-mode coarse pass) runs on the fp32 path
}

namespace {

// One MMA step of the streamed program.  `logical`: 0..7 trunk layer, 8 xyz_encoding_final,
// 9 dir_encoding | transient_encoding.0 (on xyz_encoding_final), 19 the same two layers with
// xyz_encoding_final folded in (on the trunk output), 10..12 transient_encoding.{2,4,6},
// 20 static_sigma as an N=64 step (row 0) on the trunk output, 21 the remaining heads as ONE block-structured
// K=256, N=64 step: rows 0..4 = transient_rgb(3), transient_sigma, transient_beta on k < 128
// (transient_encoding.6 output, panels 0..15), rows 8..10 = static_rgb on k >= 128 (dir_encoding, panels 16..31).
// kcap: K=16 MMA steps issued per chunk (0 = the whole chunk); K is the range the chunks cover.
struct LStep { int logical, K, N, a_panel0, kind, kcap; };

// xyz_encoding_final has no activation, so W_dir*(W_f h + b_f) = (W_dir W_f) h + W_dir b_f: folding it
// removes one 256x256 layer per fine sample (DFB_TC_FOLD_FINAL=0 keeps the literal layer sequence).
bool fold_final() {
  static int f = [] {
    const char* e = getenv("DFB_TC_FOLD_FINAL");
    return (e && e[0] == '0') ? 0 : 1;
  }();
  return f != 0;
}

// native128: the 8x128 network (the reference's default netwidth) as its own program instead of the zero-padded 8x256
// embedding: 128-column trunk steps, 64-column transient steps, a quarter of the MMA work and half of the epilogue work.
// cta_group::2 chunking only (a 16 KB image per CTA holds 128 K columns of an N = 128 step, 256 of an N = 64 step).  The
// positional encoding sits in panels 16..23 and panels 24..31 hold zeros, so that layer 0 (A = panels 16..31), the skip
// layer ([h | PE | 0] = panels 0..31) and the sigma step read whole chunks; steps with a shorter K stop early (kcap).
std::vector<LStep> build_program(bool fine, bool native128 = false) {
  std::vector<LStep> pr;
  if (native128) {
    for (int i = 0; i < 8; ++i) pr.push_back({i, i == 4 ? 256 : 128, 128, i == 0 ? 16 : 0, tc::EPI_HIDDEN, 0});
    pr.push_back({20, 256, 64, 0, tc::EPI_SIGMA, 8});
    if (!fine) return pr;
    pr.push_back({19, 128, 128, 0, tc::EPI_DT, 0});
    for (int i = 0; i < 3; ++i) pr.push_back({10 + i, 256, 64, 0, tc::EPI_T, 4});
    pr.push_back({21, 256, 64, 0, tc::EPI_HEADS, 8});
    return pr;
  }
  for (int i = 0; i < 8; ++i) pr.push_back({i, i == 0 ? 64 : (i == 4 ? 320 : 256), 256, i == 0 ? 32 : 0, tc::EPI_HIDDEN, 0});
  pr.push_back({20, 256, 64, 0, tc::EPI_SIGMA, 0});
  if (!fine) return pr;
  if (fold_final()) {
    pr.push_back({19, 256, 256, 0, tc::EPI_DT, 0});
  } else {
    pr.push_back({8, 256, 256, 0, tc::EPI_FINAL, 0});
    pr.push_back({9, 256, 256, 0, tc::EPI_DT, 0});
  }
  for (int i = 0; i < 3; ++i) pr.push_back({10 + i, 128, 128, 0, tc::EPI_T, 0});
  pr.push_back({21, 256, 64, 0, tc::EPI_HEADS, 0});
  return pr;
}

// DFB_TC_NATIVE128=0 keeps 128-wide networks on the zero-padded 8x256 embedding (A/B measurements and tests); read per
// call, the native image is always packed
bool native128_enabled() {
  const char* e = getenv("DFB_TC_NATIVE128");
  return !(e && e[0] == '0') && fold_final();
}
}  // namespace

// Pack the network into the streaming order of the kernel.  For every step, K is cut into
// chunks; a chunk is stored as `cg` consecutive 16 KB images, image h holding rows
// [h*N/cg, (h+1)*N/cg) of B as [K/8 panels][N/cg rows][8 elements] (see the layout note above).
// native = false: the 8x256 program (P already zero-padded to 256); native = true: the 8x128 program on the unpadded
// parameters (cta_group::2 images only, no split-precision image).
static int pack_tc_variant(NetPack& np, const std::vector<std::vector<float>>& P, int W, bool native) {
  const int H = W / 2, in_xyz = np.in_xyz;
  const bool fine = np.fine;
  const std::vector<LStep> prog = build_program(fine, native);
  // dir_encoding[:, :W] stacked on transient_encoding.0[:, :W] (row nn, column k)
  auto wdt = [&](int nn, int k) -> double {
    if (nn < H) return P[18][(size_t)nn * (W + np.in_dir + np.a_dim) + k];
    return P[24][(size_t)(nn - H) * (W + np.t_dim) + k];
  };
  std::vector<float> folded_w, folded_b;
  if (fine && fold_final()) {
    folded_w.resize((size_t)W * W), folded_b.resize(W);
    for (int nn = 0; nn < W; ++nn) {
      double bacc = 0.0;
      for (int j = 0; j < W; ++j) bacc += wdt(nn, j) * (double)P[17][j];
      folded_b[nn] = (float)bacc;
      for (int k = 0; k < W; ++k) {
        double acc = 0.0;
        for (int j = 0; j < W; ++j) acc += wdt(nn, j) * (double)P[16][(size_t)j * W + k];
        folded_w[(size_t)nn * W + k] = (float)acc;
      }
    }
  }
  // value of the logical weight matrix at (n, k); zero beyond the layer's true K (chunks cover whole K ranges)
  auto wval = [&](int lg, int nn, int k) -> float {
    if (lg == 0) return k < in_xyz ? P[0][(size_t)nn * in_xyz + k] : 0.f;
    if (lg < 8) {
      if (lg == 4) {  // K order [h(W) | pe(64)]; torch order is cat([input_xyz, h])
        if (k < W) return P[8][(size_t)nn * (W + in_xyz) + in_xyz + k];
        const int c = k - W;
        return c < in_xyz ? P[8][(size_t)nn * (W + in_xyz) + c] : 0.f;
      }
      return k < W ? P[2 * lg][(size_t)nn * W + k] : 0.f;
    }
    if (lg == 8) return k < W ? P[16][(size_t)nn * W + k] : 0.f;  // xyz_encoding_final
    if (lg == 9) return k < W ? (float)wdt(nn, k) : 0.f;
    if (lg == 19) return k < W ? folded_w[(size_t)nn * W + k] : 0.f;
    if (lg == 20) return nn == 0 && k < W ? P[20][k] : 0.f;  // static_sigma
    if (lg == 21) {
      if (k < H) {
        if (nn < 3) return P[34][(size_t)nn * H + k];  // transient_rgb
        if (nn == 3) return P[32][k];                  // transient_sigma
        if (nn == 4) return P[36][k];                  // transient_beta
        return 0.f;
      }
      return (nn >= 8 && nn < 11 && k < 2 * H) ? P[22][(size_t)(nn - 8) * H + (k - H)] : 0.f;  // static_rgb
    }
    return k < H ? P[26 + 2 * (lg - 10)][(size_t)nn * H + k] : 0.f;  // transient_encoding.{2,4,6}
  };
  for (int cg = native ? 2 : 1; cg <= 2; ++cg) {
    size_t total_imgs = 0;
    for (const LStep& st : prog) total_imgs += (size_t)st.K * st.N * 2 / tc::kChunkBytes;
    std::vector<uint16_t> img16[2];
    img16[0].assign(total_imgs * tc::kChunkBytes / 2, 0);
    img16[1].assign(total_imgs * tc::kChunkBytes / 2, 0);
    size_t img = 0;
    for (const LStep& st : prog) {
      const int rows = st.N / cg;
      const int kc = tc::kChunkBytes / (rows * 2);  // K columns per chunk
      DFB_REQUIRE(st.K % kc == 0, DFB_ERR_INVALID, "tcgen05 program: K = %d is not a whole number of %d-column chunks", st.K, kc);
      for (int k0 = 0; k0 < st.K; k0 += kc)
        for (int h = 0; h < cg; ++h, ++img) {
          const size_t base = img * (tc::kChunkBytes / 2);
          for (int kk = 0; kk < kc; ++kk)
            for (int r = 0; r < rows; ++r) {
              const float v = wval(st.logical, h * rows + r, k0 + kk);
              const size_t idx = base + (size_t)(kk / 8) * rows * 8 + (size_t)r * 8 + kk % 8;
              img16[0][idx] = f2h(v);
              img16[1][idx] = f2b(v);
            }
        }
    }
    DFB_REQUIRE(img == total_imgs, DFB_ERR_INVALID, "tcgen05 program: image count mismatch");
    (native ? np.blob16n_bytes : np.blob16_bytes) = total_imgs * tc::kChunkBytes;
    for (int k = 0; k < 2; ++k) {
      void** dst = native ? &np.blob16n[k] : &np.blob16[k][cg - 1];
      DFB_CHECK_CUDA(cudaMalloc(dst, total_imgs * tc::kChunkBytes));
      DFB_CHECK_CUDA(cudaMemcpy(*dst, img16[k].data(), total_imgs * tc::kChunkBytes, cudaMemcpyHostToDevice));
    }
  }
  // split-precision image of the sigma-only program (the first 9 steps: trunk + sigma), cta_group::2 chunking:
  // all hi chunks, then all lo chunks
  if (!native) {
    if (np.blob16x3) { cudaFree(np.blob16x3); np.blob16x3 = nullptr; }
    const std::vector<LStep> cprog = build_program(false);
    size_t imgs = 0;
    for (const LStep& st : cprog) imgs += (size_t)st.K * st.N * 2 / tc::kChunkBytes;
    std::vector<uint16_t> img(2 * imgs * tc::kChunkBytes / 2, 0);
    size_t ii = 0;
    for (const LStep& st : cprog) {
      const int rows = st.N / 2;
      const int kc = tc::kChunkBytes / (rows * 2);
      for (int k0 = 0; k0 < st.K; k0 += kc)
        for (int h = 0; h < 2; ++h, ++ii) {
          const size_t base = ii * (tc::kChunkBytes / 2);
          for (int kk = 0; kk < kc; ++kk)
            for (int r = 0; r < rows; ++r) {
              const float v = wval(st.logical, h * rows + r, k0 + kk);
              const size_t idx = base + (size_t)(kk / 8) * rows * 8 + (size_t)r * 8 + kk % 8;
              const __half hh = __float2half_rn(v);
              img[idx] = f2h(v);
              img[imgs * (tc::kChunkBytes / 2) + idx] = f2h(v - __half2float(hh));
            }
        }
    }
    np.blob16x3_bytes = 2 * imgs * tc::kChunkBytes;
    DFB_CHECK_CUDA(cudaMalloc(&np.blob16x3, np.blob16x3_bytes));
    DFB_CHECK_CUDA(cudaMemcpy(np.blob16x3, img.data(), np.blob16x3_bytes, cudaMemcpyHostToDevice));
  }
  // fp32 table read through the constant bank by the epilogue (see TcArgs::tbl); biases by program step
  std::vector<float>& tbl = native ? np.tc_tbl_n : np.tc_tbl;
  tbl.assign(tc::kTblFloats, 0.f);
  float* tb = tbl.data();
  for (size_t s = 0; s < prog.size(); ++s) {
    const int lg = prog[s].logical;
    float* dst = tb + s * 256;
    if (lg < 8) memcpy(dst, P[2 * lg + 1].data(), W * sizeof(float));
    else if (lg == 8) memcpy(dst, P[17].data(), W * sizeof(float));
    else if (lg == 19) memcpy(dst, folded_b.data(), W * sizeof(float));
    else if (lg >= 10 && lg <= 12) memcpy(dst, P[27 + 2 * (lg - 10)].data(), H * sizeof(float));
    // lg == 9: the bias of dir_encoding / transient_encoding.0 is part of the per-ray bias
  }
  tb[tc::kTblScal] = P[21][0];
  if (!native) {
    if (np.tc_bias32_dev) cudaFree(np.tc_bias32_dev);
    np.tc_bias32_dev = nullptr;
    DFB_CHECK_CUDA(cudaMalloc(&np.tc_bias32_dev, tc::kTblScal * sizeof(float)));
    DFB_CHECK_CUDA(cudaMemcpy(np.tc_bias32_dev, tb, tc::kTblScal * sizeof(float), cudaMemcpyHostToDevice));
  }
  if (fine) {
    std::vector<float> dtb(W, 0.f);
    if (fold_final()) dtb = folded_b;
    float** dd = native ? &np.tc_dtbias_n_dev : &np.tc_dtbias_dev;
    if (*dd) cudaFree(*dd);
    *dd = nullptr;
    DFB_CHECK_CUDA(cudaMalloc(dd, W * sizeof(float)));
    DFB_CHECK_CUDA(cudaMemcpy(*dd, dtb.data(), W * sizeof(float), cudaMemcpyHostToDevice));
  }
  if (fine) {
    for (int c = 0; c < 3; ++c) tb[tc::kTblScal + 1 + c] = P[23][c], tb[tc::kTblScal + 4 + c] = P[35][c];
    tb[tc::kTblScal + 7] = P[33][0], tb[tc::kTblScal + 8] = P[37][0];
  }
  return DFB_OK;
}

int pack_tc_weights(DfbNerf* n, int which, const std::vector<std::vector<float>>& P_in) {
  NetPack& np = n->net[which];
  for (int k = 0; k < 2; ++k) {
    for (int g = 0; g < 2; ++g)
      if (np.blob16[k][g]) { cudaFree(np.blob16[k][g]); np.blob16[k][g] = nullptr; }
    if (np.blob16n[k]) { cudaFree(np.blob16n[k]); np.blob16n[k] = nullptr; }
  }
  np.tc_tbl.clear(), np.tc_tbl_n.clear();
  if (!tc_padded_shape(np)) return DFB_OK;  // SIMT only
  // narrower networks: embedded in 8x256 with zeros (every mode of the kernel) ...
  const int rc = pack_tc_variant(np, tc_pad_params(np, P_in), 256, false);
  if (rc) return rc;
  // ... and, for the reference's default netwidth, additionally as the native 128-wide program (inference passes)
  if (np.W == 128 && fold_final()) return pack_tc_variant(np, P_in, 128, true);
  return DFB_OK;
}

// true when launch_mlp_tc_rays will run the native 128-wide program for this network (decides the layout of the per-ray
// bias: contiguous [dir 64 | transient 64] instead of the padded embedding's [dir | 0 | transient | 0])
bool tc_native128(const DfbNerf* n, int which, bool masks, bool split3) {
  const NetPack& np = n->net[which];
  return np.W == 128 && np.blob16n[0] != nullptr && !np.tc_tbl_n.empty() && !masks && !split3 && tc_cta_group_env() == 2 &&
         native128_enabled();
}

// 3-D tensor map over a packed weight image: [n_img][64][128 x u16], one box = one 16 KB image.
int make_weight_tmap(void* base, size_t bytes, CUtensorMap* out) {
  static PFN_cuTensorMapEncodeTiled encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    DFB_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    DFB_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, DFB_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    encode = (PFN_cuTensorMapEncodeTiled)fn;
  }
  const cuuint64_t dims[3] = {128, 64, (cuuint64_t)(bytes / tc::kChunkBytes)};
  const cuuint64_t strides[2] = {256, (cuuint64_t)tc::kChunkBytes};
  const cuuint32_t box[3] = {128, 64, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DFB_REQUIRE(r == CUDA_SUCCESS, DFB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return DFB_OK;
}
static unsigned long long* g_prof = nullptr;

// cta_group used by the tcgen05 kernel.  2 (default): CTA pairs share every weight chunk, which halves the
// L2 -> shared-memory weight stream per SM (the 1-CTA kernel waits ~25 % of its time for weight stages);
// measured 3.46 vs 3.39 M rays/s.  DFB_TC_CTA_GROUP=1 selects the 1-CTA kernel.
static int tc_cta_group() { return tc_cta_group_env(); }

int launch_mlp_tc_rays(const DfbNerf* nerf, int which, int mode, int kind, const float* rayrec, const float* z,
                       const float* raybias, int64_t n_rays, int S, float* raw, cudaStream_t st, uint32_t* masks, bool split3,
                       float* part, int part_k, const int* ert_rowmap, const int* ert_offsets) {
  const NetPack& np = nerf->net[which];
  DFB_REQUIRE(tc_supported(nerf, which, mode), DFB_ERR_UNSUPPORTED, "network shape not supported by the tcgen05 kernel");
  DFB_REQUIRE(kind == DFB_MMA_F16 || kind == DFB_MMA_BF16, DFB_ERR_INVALID, "bad mma kind");
  const bool full = mode == MLP_FULL;
  DFB_REQUIRE(!full || raybias, DFB_ERR_INVALID, "ray-constant inputs missing");
  DFB_REQUIRE(!split3 || (mode == MLP_SIGMA && kind == DFB_MMA_F16 && np.blob16x3), DFB_ERR_UNSUPPORTED,
              "the split-precision pass covers the sigma-only coarse network with fp16 operands");
  int* error_flag = nullptr;
  {
    const int rc = device_error_flag(&error_flag);
    if (rc) return rc;
  }
  const int cg = (split3 || ert_rowmap) ? 2 : tc_cta_group();
  const bool native = cg == 2 && tc_native128(nerf, which, masks != nullptr, split3);
  tc::TcArgs a = {};
  const std::vector<LStep> prog = build_program(full, native);
  a.hb = native ? 2 : 4, a.pe_panel0 = native ? 16 : tc::kHPanels;
  a.n_epi = (int)prog.size();
  int total_chunks = 0;
  for (const LStep& ls : prog) total_chunks += ls.K / (tc::kChunkBytes / ((ls.N / cg) * 2));
  int cb = 0, ns = 0;
  for (int e = 0; e < a.n_epi; ++e) {
    const LStep& ls = prog[e];
    const int kc = tc::kChunkBytes / ((ls.N / cg) * 2);
    tc::Step stp;
    stp.n_chunks = ls.K / kc, stp.ksteps = ls.kcap ? ls.kcap : kc / 16, stp.n = ls.N, stp.a_panel0 = ls.a_panel0, stp.chunk_base = cb;
    if (!split3) {
      a.steps[ns] = stp, a.first[ns] = 1, a.last[ns] = 1;
      if (e == 4) a.last_pe_step = ns;
      ++ns;
    } else {
      // A_hi W_hi, A_lo W_hi (lo operand: 40 panels further), A_hi W_lo (lo image: total_chunks further)
      a.steps[ns] = stp, a.first[ns] = 1, a.last[ns] = 0, ++ns;
      a.steps[ns] = stp, a.steps[ns].a_panel0 = ls.a_panel0 + tc::kHPanels + tc::kPePanels, a.first[ns] = 0, a.last[ns] = 0, ++ns;
      a.steps[ns] = stp, a.steps[ns].chunk_base = total_chunks + cb, a.first[ns] = 0, a.last[ns] = 1;
      if (e == 4) a.last_pe_step = ns;
      ++ns;
    }
    a.kind[e] = ls.kind;
    // mask layers (see TcArgs::masks): trunk 0..7, dir|transient.0 = 8, transient_encoding.{2,4,6} = 9..11
    a.mlayer[e] = ls.logical < 8 ? ls.logical : (ls.logical == 9 || ls.logical == 19) ? 8 : (ls.logical >= 10 && ls.logical <= 12) ? ls.logical - 1 : -1;
    cb += ls.K / kc;
  }
  a.n_steps = ns;
  a.bias32 = np.tc_bias32_dev;
  a.wimg = split3 ? np.blob16x3 : native ? np.blob16n[kind == DFB_MMA_F16 ? 0 : 1] : np.blob16[kind == DFB_MMA_F16 ? 0 : 1][cg - 1];
  if (cg == 2) {
    int rc = make_weight_tmap(const_cast<void*>(a.wimg), split3 ? np.blob16x3_bytes : native ? np.blob16n_bytes : np.blob16_bytes, &a.tmap);
    if (rc) return rc;
  }
  memcpy(a.tbl, (native ? np.tc_tbl_n : np.tc_tbl).data(), sizeof(a.tbl));
  for (int i = 0; i < tc::kMaxSteps * 128; ++i) {
    const float lo = a.tbl[2 * i], hi = a.tbl[2 * i + 1];
    a.btbl[i] = kind == DFB_MMA_F16 ? ((uint32_t)f2h(hi) << 16 | f2h(lo)) : ((uint32_t)f2b(hi) << 16 | f2b(lo));
  }
  a.rayrec = rayrec, a.z = z, a.raybias = raybias, a.S = S, a.P = n_rays * S, a.raw = raw;
  DFB_REQUIRE(!masks || full, DFB_ERR_INVALID, "ReLU masks are an output of the fine network only");
  DFB_REQUIRE(!part || (full && !masks && part_k >= 1), DFB_ERR_INVALID, "fused compositing is a mode of the fine pass without masks");
  DFB_REQUIRE(!ert_rowmap || (part && ert_offsets), DFB_ERR_INVALID, "early ray termination is a mode of the fused fine pass");
  a.masks = masks;
  a.part = part, a.part_k = part_k;
  a.rowmap = ert_rowmap, a.offsets = ert_offsets, a.P_dev = ert_offsets ? ert_offsets + n_rays : nullptr;
  a.error_flag = error_flag;
#ifdef DFB_TC_PROF
  if (!g_prof) {
    DFB_CHECK_CUDA(cudaMalloc(&g_prof, 512 * 16 * sizeof(unsigned long long)));
  }
  {
    const char* w = getenv("DFB_TC_PROF_WHICH");
    if (!w || atoi(w) == which) {
      DFB_CHECK_CUDA(cudaMemsetAsync(g_prof, 0, 512 * 16 * sizeof(unsigned long long), st));
      a.prof = g_prof;
    }
  }
#endif
  if (a.P == 0) return DFB_OK;
  const int64_t tiles = (a.P + tc::kTileM - 1) / tc::kTileM;
  const int nslot = split3 ? 1 : 2;
  a.n_pass = (tiles + nslot * cg - 1) / (nslot * cg);
  const int grid = cg * (int)std::min<int64_t>(a.n_pass, nerf->num_sms / cg);
  auto launch = [&](auto kern) -> int {
    DFB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemTotal));
    kern<<<grid, tc::kThreads, tc::kSmemTotal, st>>>(a);
    DFB_LAUNCH_CHECK();
    return DFB_OK;
  };
  const bool f16 = kind == DFB_MMA_F16;
  if (split3) return launch(tc::k_mlp_tc2_x3);
  if (ert_rowmap) return f16 ? launch(tc::k_mlp_tc2<__half, 4>) : launch(tc::k_mlp_tc2<__nv_bfloat16, 4>);
  if (part) {
    if (cg == 2) return f16 ? launch(tc::k_mlp_tc2<__half, 3>) : launch(tc::k_mlp_tc2<__nv_bfloat16, 3>);
    return f16 ? launch(tc::k_mlp_tc<__half, 3>) : launch(tc::k_mlp_tc<__nv_bfloat16, 3>);
  }
  if (cg == 2) {
    if (masks) return f16 ? launch(tc::k_mlp_tc2<__half, 2>) : launch(tc::k_mlp_tc2<__nv_bfloat16, 2>);
    if (f16) return full ? launch(tc::k_mlp_tc2<__half, 1>) : launch(tc::k_mlp_tc2<__half, 0>);
    return full ? launch(tc::k_mlp_tc2<__nv_bfloat16, 1>) : launch(tc::k_mlp_tc2<__nv_bfloat16, 0>);
  }
  if (masks) return f16 ? launch(tc::k_mlp_tc<__half, 2>) : launch(tc::k_mlp_tc<__nv_bfloat16, 2>);
  if (f16) return full ? launch(tc::k_mlp_tc<__half, 1>) : launch(tc::k_mlp_tc<__half, 0>);
  return full ? launch(tc::k_mlp_tc<__nv_bfloat16, 1>) : launch(tc::k_mlp_tc<__nv_bfloat16, 0>);
}

}  // namespace dfb

// Debug seam: cycle counters of the last tcgen05 MLP launch (DFB_TC_PROF builds only).
// out[cta][16]: producer {wait W_EMPTY,-,-,total}, MMA {wait W_FULL, wait W_FULLP, wait A_READY, total},
// epilogue slot0 {wait D_FULL,-,-,total}, epilogue slot1 {...}.
extern "C" int dfb_debug_tc_prof(unsigned long long* out_host, int n_cta) {
  if (!dfb::g_prof) return DFB_ERR_UNSUPPORTED;
  cudaDeviceSynchronize();
  cudaMemcpy(out_host, dfb::g_prof, (size_t)n_cta * 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);  // n_cta up to 512: rows 256.. hold the issuer's per-step A_READY waits
  return DFB_OK;
}

// Debug seam: tensor-pipe rate with the kernel's operand layout; returns mean cycles per MMA over all CTAs.
extern "C" int dfb_debug_umma_rate(int iters, int n, int grid, double* cycles_per_mma) {
  using namespace dfb;
  DFB_REQUIRE(cycles_per_mma && iters >= 1 && n >= 16 && n <= 256 && n % 16 == 0 && grid >= 1 && grid <= 1024, DFB_ERR_INVALID, "bad arguments");
  unsigned long long* d = nullptr;
  DFB_CHECK_CUDA(cudaMalloc(&d, grid * 8));
  const int smem = 65536 + 32768;
  DFB_CHECK_CUDA(cudaFuncSetAttribute(tc::k_umma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  tc::k_umma_rate<<<grid, 128, smem>>>(iters, n, d);
  DFB_LAUNCH_CHECK();
  std::vector<unsigned long long> h(grid);
  DFB_CHECK_CUDA(cudaMemcpy(h.data(), d, grid * 8, cudaMemcpyDeviceToHost));
  cudaFree(d);
  double s = 0;
  for (auto v : h) s += (double)v;
  *cycles_per_mma = s / grid / ((double)iters * 16.0);
  return DFB_OK;
}

// Debug seam: TMEM read rate; returns mean cycles per tcgen05.ld.32x32b.x32 (4 KB) per warp with `nwarps` warps reading.
extern "C" int dfb_debug_tmem_rate(int iters, int nwarps, int grid, double* cycles_per_ld) {
  using namespace dfb;
  DFB_REQUIRE(cycles_per_ld && iters >= 1 && nwarps >= 1 && nwarps <= 4 && grid >= 1 && grid <= 1024, DFB_ERR_INVALID, "bad arguments");
  unsigned long long* d = nullptr;
  DFB_CHECK_CUDA(cudaMalloc(&d, grid * 4 * 8));
  DFB_CHECK_CUDA(cudaMemset(d, 0, grid * 4 * 8));
  tc::k_tmem_rate<<<grid, 128>>>(iters, nwarps, d);
  DFB_LAUNCH_CHECK();
  std::vector<unsigned long long> h(grid * 4);
  DFB_CHECK_CUDA(cudaMemcpy(h.data(), d, grid * 4 * 8, cudaMemcpyDeviceToHost));
  cudaFree(d);
  double s = 0;
  for (int b = 0; b < grid; ++b)
    for (int w = 0; w < nwarps; ++w) s += (double)h[b * 4 + w];
  *cycles_per_ld = s / ((double)grid * nwarps) / ((double)iters * 8.0);
  return DFB_OK;
}

// Debug seam: the same while a fifth warp keeps the tensor pipe busy with `mma_iters` x 16 MMAs (M=128, N=256, K=16).
// cycles[0] = mean cycles per tcgen05.ld per warp, cycles[1] = cycles per MMA as seen by the issuing warp.
extern "C" int dfb_debug_tmem_rate_mma(int iters, int nwarps, int mma_iters, int grid, double* cycles) {
  using namespace dfb;
  DFB_REQUIRE(cycles && iters >= 1 && nwarps >= 1 && nwarps <= 4 && mma_iters >= 1 && grid >= 1 && grid <= 1024, DFB_ERR_INVALID, "bad arguments");
  unsigned long long* d = nullptr;
  DFB_CHECK_CUDA(cudaMalloc(&d, grid * 8 * 8));
  DFB_CHECK_CUDA(cudaMemset(d, 0, grid * 8 * 8));
  const int smem = 65536 + 32768;
  DFB_CHECK_CUDA(cudaFuncSetAttribute(tc::k_tmem_rate_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  tc::k_tmem_rate_mma<<<grid, 160, smem>>>(iters, nwarps, mma_iters, d);
  DFB_LAUNCH_CHECK();
  std::vector<unsigned long long> h(grid * 8);
  DFB_CHECK_CUDA(cudaMemcpy(h.data(), d, grid * 8 * 8, cudaMemcpyDeviceToHost));
  cudaFree(d);
  double s = 0, m = 0;
  for (int b = 0; b < grid; ++b) {
    for (int w = 0; w < nwarps; ++w) s += (double)h[b * 8 + w];
    m += (double)h[b * 8 + 4];
  }
  cycles[0] = s / ((double)grid * nwarps) / ((double)iters * 8.0);
  cycles[1] = m / grid / ((double)mma_iters * 16.0);
  return DFB_OK;
}

// Debug seam (not used by the product path): single-tile UMMA GEMM through the same descriptor
// and TMEM code as the MLP kernel.  variant 0 is the layout the kernel uses.
extern "C" int dfb_debug_umma_gemm(const float* A, const float* B, int N, int K, int kind, int variant, float* D,
                                   void* stream) {
  using namespace dfb;
  DFB_REQUIRE(A && B && D, DFB_ERR_INVALID, "null argument");
  DFB_REQUIRE(N % 32 == 0 && N >= 32 && N <= 256 && K % 16 == 0 && K >= 16 && K <= 320, DFB_ERR_INVALID, "bad N/K");
  int* flag = nullptr;
  DFB_CHECK_CUDA(cudaMalloc(&flag, 4));
  DFB_CHECK_CUDA(cudaMemset(flag, 0, 4));
  const size_t smem = (size_t)K * 128 * 2 + (size_t)K * N * 2;
  cudaStream_t st = (cudaStream_t)stream;
  if (kind == DFB_MMA_BF16) {
    DFB_CHECK_CUDA(cudaFuncSetAttribute(tc::k_umma_selftest<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc::k_umma_selftest<__nv_bfloat16><<<1, 128, smem, st>>>(A, B, N, K, variant, D, flag);
  } else {
    DFB_CHECK_CUDA(cudaFuncSetAttribute(tc::k_umma_selftest<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc::k_umma_selftest<__half><<<1, 128, smem, st>>>(A, B, N, K, variant, D, flag);
  }
  DFB_LAUNCH_CHECK();
  DFB_CHECK_CUDA(cudaStreamSynchronize(st));
  cudaFree(flag);
  return DFB_OK;
}
