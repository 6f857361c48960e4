// Weight-gradient convolution on tcgen05 for the DFNet pose regressor (reference
// feature/direct_feature_matching.py:378 `loss.backward()` through feature/dfnet.py:74-172):
//
//   dW[n, c, ky, kx] = sum_{b,y,x} gO[b,y,x,n] * X[b, y+ky-pad, x+kx-pad, c]
//
// Both operands are NHWC, i.e. the contraction index (pixel) is the SLOW one: they are MN-major
// tcgen05 operands.  With the im2col-free patch layout of conv_tc.cu,
// [8-channel panel][patch row][patch col][16 B], eight consecutive pixels of a patch row form one
// 8(K) x 16 B(MN) core matrix, the next channel panel is SBO away and the next patch row (the next
// 8 pixels of the K dimension) LBO away, so again every filter tap is only a shifted start address.
// Tiles (gO tile, X patch of one filter row incl. halo) arrive as two TMA boxes each through a 3-stage ring.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "tc_common.cuh"

namespace dfb {
namespace wg {

using namespace dfb::tc;

// instruction descriptor with explicit operand formats and major-ness (bit 15 / 16: 1 = MN-major)
__device__ __forceinline__ uint32_t make_idesc_ex(int fmt_a, int fmt_b, int a_mn, int b_mn, int n, int m) {
  return (1u << 4) | ((uint32_t)fmt_a << 7) | ((uint32_t)fmt_b << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ uint16_t cvt16(float v, int fmt) {
  if (fmt) return __bfloat16_as_ushort(__float2bfloat16_rn(v));
  return __half_as_ushort(__float2half_rn(v));
}


// ------------------------------------------------------------------------------------------
// weight-gradient kernel
// ------------------------------------------------------------------------------------------
constexpr int kTH = 16, kTW = 8;   // pixel tile = 128 contraction steps
constexpr int kMaxStages = 3;
constexpr int kThreads = 160;      // warps 0-3: epilogue (warp 0 first produces the tiles by TMA); warp 4: MMA issuer
constexpr uint32_t kGBytes = 16u * kTH * kTW * 16u;  // gO tile: 16 channel panels x 128 pixels x 16 B

enum Bar { FULL = 0, EMPTY = 3, DONE = 6, N_BARS = 7 };

struct WgArgs {
  // 5-D tensor maps over gO and X (NHWC seen as [B][C/8][H][W][8]): one box = the tile's shared-memory image
  alignas(64) CUtensorMap tmap_g;
  alignas(64) CUtensorMap tmap_x;
  int n_stages;        // tile ring depth (3 where shared memory allows, else 2)
  int sw128;           // 1: pixel-major 128-byte-swizzled tiles (1x1 layers with 64-channel multiples), see the producer
  int tap_major;       // 1: dW points to a scratch image [KH][KW][Cout][Cin_s] (input channel fastest) that takes 16-byte vector
  int Cin_s;           //    atomics; k_wgrad_unpack transposes it into [Cout][Cin][KH][KW] afterwards
  int atomic;          // 1: the epilogue adds to dW (pixel tiles split over CTAs, or the caller accumulates); 0: it stores
  const uint16_t* gO;  // NHWC [B,H,W,Cout]
  const uint16_t* X;   // NHWC [B,H,W,Cin_pad]
  float* dW;           // [Cout][Cin][KH][KW], accumulated with atomics (zeroed by the caller)
  int B, H, W, Cin, Cin_pad, Cout, KH, KW, pad;
  int NB;              // input-channel block = N of the MMA (multiple of 16, NB * KW <= 512 TMEM columns)
  int n_cib, n_cob, n_split;
  int tiles_x, tiles_y, n_tiles;
  int PW;              // patch columns incl. halo
  uint32_t x_bytes;    // X patch: NB/8 panels x 16 rows x PW cols x 16 B
  int fmt;
  int* error_flag;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One CTA = (128 output channels, NB input channels, one filter row ky, one share of the pixel tiles).
// D[co][kx*NB + ci] accumulates in TMEM over all the CTA's tiles; A = gO tile, B = X patch shifted by kx.
__global__ void __launch_bounds__(kThreads, 1) k_conv_wgrad(const __grid_constant__ WgArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int kStages = a.n_stages;
  const uint32_t sG = smem_u32(smem);
  const uint32_t sX = sG + kStages * kGBytes;
  const uint32_t sBar = sX + kStages * a.x_bytes;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kStages * kGBytes + kStages * a.x_bytes + N_BARS * 8);
  const int tid = threadIdx.x, warp = tid >> 5;
  auto bar = [&](int i) { return sBar + 8u * i; };

  int u = blockIdx.x;
  const int split = u % a.n_split; u /= a.n_split;
  const int ky = u % a.KH; u /= a.KH;
  const int cib = u % a.n_cib;
  const int cob = u / a.n_cib;
  const int t0 = (int)((int64_t)split * a.n_tiles / a.n_split), t1 = (int)((int64_t)(split + 1) * a.n_tiles / a.n_split);
  if (t0 >= t1) return;  // uniform per CTA

  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&a.tmap_g)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&a.tmap_x)) : "memory");
    for (int i = 0; i < kMaxStages; ++i) mbar_init(bar(FULL + i), 1), mbar_init(bar(EMPTY + i), 1);
    mbar_init(bar(DONE), 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc<1>(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();               // global memory is read and written only from here on
  pdl_launch_dependents();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_img = a.tiles_x * a.tiles_y;
  const int npan = a.NB / 8;
  const uint32_t xpanel = (uint32_t)kTH * a.PW * 16u;

  if (warp < 4) {
    if (warp == 0) {
      // ===== tile producer: two TMA boxes per tile (gO tile, X patch of filter row ky incl. halo; out-of-image pixels and
      // channels past the tensor are zero-filled by the TMA unit).  The first version gathered both with 16-byte cp.async
      // from 128 threads (~1 000 address-arithmetic instructions per thread and tile, two stages): the weight-gradient
      // kernels ran at ~200 TFLOP/s, loader-bound.
      uint32_t seq = 0;
      for (int t = t0; t < t1; ++t, ++seq) {
        const uint32_t st = seq % kStages, ph = (seq / kStages) & 1;
        mbar_wait(bar(EMPTY + st), ph ^ 1, a.error_flag);
        if (elect_one()) {
          const int b = t / tiles_img, ti = t % tiles_img;
          const int y0 = (ti / a.tiles_x) * kTH, x0 = (ti % a.tiles_x) * kTW;
          mbar_expect_tx(bar(FULL + st), kGBytes + a.x_bytes);
          if (a.sw128) {
            // 1x1 layers: a tile is 128 pixels x 64 channels per box, one 128-byte row per pixel, 128-byte swizzle - the
            // canonical MN-major SWIZZLE_128B operand (8 pixels x 128 B atoms).  An eighth of the rows the TMA unit has to
            // move compared with the 16-byte rows of the panel layout, which is what bounded this kernel (~5 000 cycles per
            // tile against 1 536 for its MMAs).
            for (int sl = 0; sl < 2; ++sl)
              tma_load_4d(sG + st * kGBytes + sl * 16384u, &a.tmap_g, cob * 128 + sl * 64, x0, y0, b, bar(FULL + st));
            for (int sl = 0; sl < a.NB / 64; ++sl)
              tma_load_4d(sX + st * a.x_bytes + sl * 16384u, &a.tmap_x, cib * a.NB + sl * 64, x0, y0, b, bar(FULL + st));
          } else {
            tma_load_5d<1>(sG + st * kGBytes, &a.tmap_g, 0, x0, y0, cob * 16, b, bar(FULL + st));
            tma_load_5d<1>(sX + st * a.x_bytes, &a.tmap_x, 0, x0 - a.pad, y0 + ky - a.pad, cib * npan, b, bar(FULL + st));
          }
        }
        __syncwarp();
      }
    }

    // ===== epilogue: thread = output channel; columns = (kx, ci) ==================================
    mbar_wait(bar(DONE), 0, a.error_flag);
    tc_fence_after();
    const int co = cob * 128 + tid;
    const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int kx = 0; kx < a.KW; ++kx)
      for (int c0 = 0; c0 < a.NB; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(t_row + kx * a.NB + c0, v);
        tmem_ld_wait(v);
        if (co < a.Cout && a.tap_major) {
          float* base = a.dW + ((int64_t)(ky * a.KW + kx) * a.Cout + co) * a.Cin_s;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const int ci = cib * a.NB + c0 + j;
            if (c0 + j < a.NB && ci < a.Cin_s)   // columns past Cin hold the zero padding channels of X
              atomicAdd(reinterpret_cast<float4*>(base + ci),
                        make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])));
          }
        } else if (co < a.Cout && a.atomic && a.KH == 1 && (a.Cin & 3) == 0) {
          // 1x1 layers: a thread's 32 columns are contiguous in dW -> 16-byte vector atomics (a quarter of the L2 atomic
          // operations; with the pixel tiles split over every SM each dW element receives ~148 of them)
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const int ci = cib * a.NB + c0 + j;
            if (c0 + j < a.NB && ci < a.Cin)   // NB and Cin are multiples of 4: a group is in range as a whole
              atomicAdd(reinterpret_cast<float4*>(a.dW + (int64_t)co * a.Cin + ci),
                        make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])));
          }
        } else if (co < a.Cout) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int ci = cib * a.NB + c0 + j;
            if (c0 + j < a.NB && ci < a.Cin) {
              float* o = a.dW + (((int64_t)co * a.Cin + ci) * a.KH + ky) * a.KW + kx;
              if (a.atomic) atomicAdd(o, __uint_as_float(v[j]));
              else *o = __uint_as_float(v[j]);   // this CTA covered every pixel tile: a plain store, no zero-fill needed
            }
          }
        }
      }
    tc_fence_before();
  } else {
    // ===== MMA issuer ===============================================================================
    const uint32_t idesc = make_idesc_ex(a.fmt, a.fmt, 1, 1, a.NB, 128);
    // MN-major, no swizzle: LBO = distance between 8-pixel K groups (next patch row), SBO = panel stride
    // MN-major, 128-byte swizzle (sw128): SBO = 1024 B between 8-pixel K groups, LBO = 16 KB between 64-channel slabs
    const uint32_t a_hi = a.sw128 ? ((1024u >> 4) | (1u << 14) | (2u << 29)) : ((2048u >> 4) | (1u << 14));
    const uint32_t a_lbo = (a.sw128 ? (16384u >> 4) : (128u >> 4)) << 16;
    const uint32_t b_hi = a.sw128 ? a_hi : ((xpanel >> 4) | (1u << 14));
    const uint32_t b_lbo = (a.sw128 ? (16384u >> 4) : ((uint32_t)(a.PW * 16) >> 4)) << 16;
    const uint32_t a_ks = a.sw128 ? (2048u >> 4) : 16u;                       // start-address step per K = 16 pixels
    const uint32_t b_ks = a.sw128 ? (2048u >> 4) : (uint32_t)(2 * a.PW);
    uint32_t seq = 0;
    for (int t = t0; t < t1; ++t, ++seq) {
      const uint32_t st = seq % kStages, ph = (seq / kStages) & 1;
      mbar_wait(bar(FULL + st), ph, a.error_flag);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t g_lo = ((sG + st * kGBytes) >> 4) | a_lbo;
        const uint32_t x_lo = ((sX + st * a.x_bytes) >> 4) | b_lbo;
        for (int kx = 0; kx < a.KW; ++kx) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)  // K = 16 pixels = two patch rows
            umma_f16<1>(tmem_base + kx * a.NB, mk64(g_lo + ks * a_ks, a_hi), mk64(x_lo + ks * b_ks + (uint32_t)kx, b_hi), idesc,
                        (uint32_t)(seq | ks));
        }
        umma_commit<1>(bar(EMPTY + st));
        if (t == t1 - 1) umma_commit<1>(bar(DONE));
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

// dW[co][ci][ky][kx] (+)= S[ky][kx][co][ci]: the tap-major scratch image of a 3x3 / 5x5 weight gradient into the parameter's
// layout.  With the pixel tiles split over the SMs every dW element receives n_split atomics; in the parameter's layout a
// thread's 32 accumulator columns are KH*KW floats apart (scalar atomics, ~75 M of them for the 13 encoder layers of a
// 480x640 training step), in the tap-major image they are contiguous and go out as 16-byte vector atomics.
__global__ void __launch_bounds__(256) k_wgrad_unpack(const float* __restrict__ S, float* __restrict__ dW, int Cout, int Cin, int Cin_s,
                                                      int taps, int accumulate) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= (int64_t)Cout * Cin) return;
  const int co = (int)(i / Cin), ci = (int)(i - (int64_t)co * Cin);
  float* o = dW + i * taps;
  for (int t = 0; t < taps; ++t) {
    const float v = S[((int64_t)t * Cout + co) * Cin_s + ci];
    o[t] = accumulate ? o[t] + v : v;
  }
}

// gB[n] += sum over pixels of gO[pix][n]   (bias gradient); grid (Cout/64, splits), 256 threads.  Generic fallback.
template <typename T>
__global__ void k_bias_grad_generic(const uint16_t* __restrict__ gO, int64_t npix, int C, float* __restrict__ gB) {
  const int c = blockIdx.x * 64 + (threadIdx.x & 63), part = threadIdx.x >> 6;
  pdl_wait();
  pdl_launch_dependents();
  float s = 0.f;
  for (int64_t p = (int64_t)blockIdx.y * 4 + part; p < npix; p += (int64_t)gridDim.y * 4)
    s += std::is_same<T, __nv_bfloat16>::value ? __uint_as_float((uint32_t)gO[p * C + c] << 16)
                                               : __half2float(__ushort_as_half(gO[p * C + c]));
  __shared__ float sm[256];
  sm[threadIdx.x] = s;
  __syncthreads();
  if (part == 0) atomicAdd(gB + c, sm[threadIdx.x] + sm[threadIdx.x + 64] + sm[threadIdx.x + 128] + sm[threadIdx.x + 192]);
}

// Same for C in {64, 128, 256, 512}: a thread reads 16 bytes (8 channels) of a pixel, a block strides over the pixels
// with 256 / (C/8) pixel lanes, fp32 partial sums are combined through shared memory and one atomic per channel and
// block.  HBM-bound (the first version read 2 bytes per thread from 64-512 blocks: 0.31 ms for the 13 encoder layers of
// the training step, 166 MB).
template <typename T>
__global__ void __launch_bounds__(256) k_bias_grad(const uint16_t* __restrict__ gO, int64_t npix, int C, float* __restrict__ gB) {
  const int cgs = C >> 3, g = threadIdx.x % cgs, lane = threadIdx.x / cgs, nl = 256 / cgs;
  pdl_wait();
  pdl_launch_dependents();
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  auto add8 = [&](const uint4& v) {
    const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (std::is_same<T, __nv_bfloat16>::value) {
        s[2 * e] += __uint_as_float(w4[e] << 16), s[2 * e + 1] += __uint_as_float(w4[e] & 0xffff0000u);
      } else {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w4[e]));
        s[2 * e] += f.x, s[2 * e + 1] += f.y;
      }
    }
  };
  // four independent 16-byte loads in flight per thread (one per trip left every load waiting for the previous add:
  // 1.4 TB/s on the 50 MB gradients of a NeRF layer)
  const int64_t step = (int64_t)gridDim.x * nl;
  int64_t p = (int64_t)blockIdx.x * nl + lane;
  for (; p + 3 * step < npix; p += 4 * step) {
    const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(gO + p * C) + g);
    const uint4 v1 = __ldg(reinterpret_cast<const uint4*>(gO + (p + step) * C) + g);
    const uint4 v2 = __ldg(reinterpret_cast<const uint4*>(gO + (p + 2 * step) * C) + g);
    const uint4 v3 = __ldg(reinterpret_cast<const uint4*>(gO + (p + 3 * step) * C) + g);
    add8(v0), add8(v1), add8(v2), add8(v3);
  }
  for (; p < npix; p += step) add8(__ldg(reinterpret_cast<const uint4*>(gO + p * C) + g));
  __shared__ float sm[256 * 8];
#pragma unroll
  for (int e = 0; e < 8; ++e) sm[lane * C + g * 8 + e] = s[e];
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float t = 0.f;
    for (int l = 0; l < nl; ++l) t += sm[l * C + c];
    atomicAdd(gB + c, t);
  }
}

// Single-tile self test of MN-major operands: D[128,N] = sum_k A[k][m] * B[k][n].
// smem image: [MN panel of 8][k][16 B]  (element (mn,k) at (mn/8)*K*16 + k*16 + (mn%8)*2).
__global__ void __launch_bounds__(128, 1) k_umma_selftest_mn(const float* A, const float* Bm, int N, int K, int fmt_a, int fmt_b,
                                                               int variant, float* D, int* error_flag) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint16_t* sAm = reinterpret_cast<uint16_t*>(smem);
  uint16_t* sBm = reinterpret_cast<uint16_t*>(smem + (size_t)K * 128 * 2);
  for (int i = tid; i < 128 * K; i += 128) {
    const int k = i / 128, m = i % 128;
    sAm[(size_t)(m / 8) * K * 8 + k * 8 + m % 8] = cvt16(A[i], fmt_a);
  }
  for (int i = tid; i < N * K; i += 128) {
    const int k = i / N, n = i % N;
    sBm[(size_t)(n / 8) * K * 8 + k * 8 + n % 8] = cvt16(Bm[i], fmt_b);
  }
  if (tid == 0) { mbar_init(smem_u32(&mbar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&tslot), 256);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tslot;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_ex(fmt_a, fmt_b, 1, 1, N, 128);
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint32_t a_addr = smem_u32(sAm) + ks * 256, b_addr = smem_u32(sBm) + ks * 256;
      uint64_t ad, bd;
      // variant 0: LBO = K-group stride (128 B), SBO = MN-panel stride (K*16 B); variant 1: swapped
      if (variant == 0) ad = make_desc(a_addr, 128, K * 16), bd = make_desc(b_addr, 128, K * 16);
      else ad = make_desc(a_addr, K * 16, 128), bd = make_desc(b_addr, K * 16, 128);
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

}  // namespace wg
}  // namespace dfb

using namespace dfb;


// dW [Cout,Cin,KH,KW] (+ optional dB [Cout]) from gO NHWC [B,H,W,Cout] and X NHWC [B,H,W,Cin_pad] (16-bit, fmt 0 f16 / 1 bf16).
// Both outputs are overwritten (accumulate != 0: added to, the caller zeroed them - one memset for a whole flat gradient
// buffer instead of two per layer).
static int conv_wgrad_impl(const void* gO, const void* X, int B, int H, int W, int Cin, int Cin_pad, int Cout, int KH, int fmt,
                           float* dW, float* dB, void* stream, int accumulate);

extern "C" int dfb_conv_wgrad(const void* gO, const void* X, int B, int H, int W, int Cin, int Cin_pad, int Cout, int KH, int fmt,
                              float* dW, float* dB, void* stream) {
  return conv_wgrad_impl(gO, X, B, H, W, Cin, Cin_pad, Cout, KH, fmt, dW, dB, stream, 0);
}

extern "C" int dfb_conv_wgrad_acc(const void* gO, const void* X, int B, int H, int W, int Cin, int Cin_pad, int Cout, int KH, int fmt,
                                  float* dW, float* dB, void* stream) {
  return conv_wgrad_impl(gO, X, B, H, W, Cin, Cin_pad, Cout, KH, fmt, dW, dB, stream, 1);
}

// Scratch image of the tap-major path: one buffer per (host thread, stream), grown on demand and kept for the life of the
// process (at most 9.4 MB for a 512 x 512 x 3 x 3 layer).  Launches on one stream are ordered, so a buffer is never shared
// by two gradients in flight; growing it goes through cudaFree, which waits for the device.
static int wgrad_scratch(cudaStream_t st, size_t bytes, float** out) {
  struct Entry { cudaStream_t st; int dev; float* p; size_t bytes; };
  static thread_local std::vector<Entry> cache;
  int dev = 0;
  DFB_CHECK_CUDA(cudaGetDevice(&dev));
  for (Entry& e : cache)
    if (e.st == st && e.dev == dev) {
      if (e.bytes < bytes) {
        DFB_CHECK_CUDA(cudaFree(e.p));
        e.p = nullptr, e.bytes = 0;
        DFB_CHECK_CUDA(cudaMalloc(&e.p, bytes));
        e.bytes = bytes;
      }
      *out = e.p;
      return DFB_OK;
    }
  Entry e = {st, dev, nullptr, std::max<size_t>(bytes, (size_t)512 * 512 * 9 * 4)};
  DFB_CHECK_CUDA(cudaMalloc(&e.p, e.bytes));
  cache.push_back(e);
  *out = e.p;
  return DFB_OK;
}

static int conv_wgrad_impl(const void* gO, const void* X, int B, int H, int W, int Cin, int Cin_pad, int Cout, int KH, int fmt,
                           float* dW, float* dB, void* stream, int accumulate) {
  DFB_REQUIRE(gO && X && dW, DFB_ERR_INVALID, "dfb_conv_wgrad: null argument");
  DFB_REQUIRE(B >= 1 && H >= 1 && W >= 1 && Cin >= 1 && Cin_pad % 8 == 0 && Cin_pad >= Cin && Cout % 64 == 0 && Cout >= 64,
              DFB_ERR_INVALID, "dfb_conv_wgrad: bad shape");
  DFB_REQUIRE(KH == 1 || KH == 3 || KH == 5, DFB_ERR_UNSUPPORTED, "kernel size must be 1, 3 or 5");
  int* error_flag = nullptr;
  {
    const int rc = device_error_flag(&error_flag);
    if (rc) return rc;
  }
  int dev = 0, sms = 0;
  DFB_CHECK_CUDA(cudaGetDevice(&dev));
  DFB_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  cudaStream_t st = (cudaStream_t)stream;
  wg::WgArgs a = {};
  a.gO = (const uint16_t*)gO, a.X = (const uint16_t*)X, a.dW = dW;
  a.B = B, a.H = H, a.W = W, a.Cin = Cin, a.Cin_pad = Cin_pad, a.Cout = Cout, a.KH = KH, a.KW = KH, a.pad = KH / 2;
  const int nb_max = KH <= 3 ? 128 : 64;
  a.NB = std::min(nb_max, round_up(Cin_pad, 16));
  a.n_cib = (Cin_pad + a.NB - 1) / a.NB;
  a.n_cob = (Cout + 127) / 128;
  a.tiles_x = (W + wg::kTW - 1) / wg::kTW, a.tiles_y = (H + wg::kTH - 1) / wg::kTH;
  a.n_tiles = a.tiles_x * a.tiles_y * B;
  const int units = a.n_cib * a.n_cob * KH;
  // Pixel tiles are split over as many CTAs as the GPU has SMs (fp32 atomics at the end; a CTA that covers every tile
  // stores directly).  Measured and dropped: running the deep layers unsplit (conv4_x / conv5_x of a 480x640 image have 40 /
  // 10 tiles) - a tile costs ~5 000 cycles (its two TMA boxes are 4 600 separate 16-byte rows; the MMAs are 1 536), so
  // fewer CTAs lose more than the atomics cost.
  a.n_split = std::max(1, std::min(a.n_tiles, sms / units));
  a.PW = wg::kTW + 2 * a.pad;
  a.x_bytes = (uint32_t)(a.NB / 8) * wg::kTH * a.PW * 16u;
  a.fmt = fmt, a.error_flag = error_flag;
  a.n_stages = (size_t)wg::kMaxStages * (wg::kGBytes + a.x_bytes) + 256 <= 232448 ? wg::kMaxStages : 2;
  const size_t smem = (size_t)a.n_stages * (wg::kGBytes + a.x_bytes) + 256;
  DFB_REQUIRE(smem <= 232448, DFB_ERR_UNSUPPORTED, "shared memory budget exceeded");
  DFB_REQUIRE(Cout % 8 == 0, DFB_ERR_INVALID, "dfb_conv_wgrad: Cout must be a multiple of 8");
  {
    const char* e = getenv("DFB_WGRAD_SW128");   // =0: the panel layout for every layer (A/B, tests)
    a.sw128 = (KH == 1 && a.NB % 64 == 0 && !(e && e[0] == '0')) ? 1 : 0;
    int rc = make_patch_tmap(gO, B, H, W, Cout, wg::kTH, wg::kTW, &a.tmap_g, a.sw128 ? -64 : 16);
    if (rc) return rc;
    rc = make_patch_tmap(X, B, H, W, Cin_pad, wg::kTH, a.PW, &a.tmap_x, a.sw128 ? -64 : a.NB / 8);
    if (rc) return rc;
  }
  a.atomic = (a.n_split > 1 || accumulate) ? 1 : 0;
  float* scratch = nullptr;
  {
    const char* e = getenv("DFB_WGRAD_TAP_MAJOR");   // =0: scalar atomics straight into dW (A/B, tests)
    if (KH > 1 && a.n_split > 1 && !(e && e[0] == '0')) {
      a.Cin_s = round_up(Cin, 4);
      const size_t bytes = (size_t)KH * KH * Cout * a.Cin_s * 4;
      const int rc = wgrad_scratch(st, bytes, &scratch);
      if (rc) return rc;
      DFB_CHECK_CUDA(cudaMemsetAsync(scratch, 0, bytes, st));
      a.tap_major = 1, a.dW = scratch;
    }
  }
  if (a.n_split > 1 && !accumulate && !a.tap_major) DFB_CHECK_CUDA(cudaMemsetAsync(dW, 0, (size_t)Cout * Cin * KH * KH * 4, st));
  {
    static thread_local uint64_t attr_set = 0;   // bit d: set on device d
    int cur = 0;
    DFB_CHECK_CUDA(cudaGetDevice(&cur));
    if (!((attr_set >> (cur & 63)) & 1)) {
      DFB_CHECK_CUDA(cudaFuncSetAttribute(wg::k_conv_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
      attr_set |= 1ull << (cur & 63);
    }
  }
  DFB_CHECK_CUDA(dfb_launch_pdl(wg::k_conv_wgrad, dim3(units * a.n_split), dim3(wg::kThreads), smem, st, true, a));
  DFB_LAUNCH_CHECK();
  if (a.tap_major) {
    const int64_t n = (int64_t)Cout * Cin;
    wg::k_wgrad_unpack<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(scratch, dW, Cout, Cin, a.Cin_s, KH * KH, accumulate);
    DFB_LAUNCH_CHECK();
  }
  if (dB) {
    if (!accumulate) DFB_CHECK_CUDA(cudaMemsetAsync(dB, 0, (size_t)Cout * 4, st));
    const int64_t npix = (int64_t)B * H * W;
    if (Cout == 64 || Cout == 128 || Cout == 256 || Cout == 512) {
      const int nl = 256 / (Cout / 8);
      const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((npix + 4 * nl - 1) / (4 * nl), 148 * 8));
      auto kern = fmt ? wg::k_bias_grad<__nv_bfloat16> : wg::k_bias_grad<__half>;
      DFB_CHECK_CUDA(dfb_launch_pdl(kern, dim3(blocks), dim3(256), 0, st, true, (const uint16_t*)gO, npix, Cout, dB));
    } else {
      const int splits = (int)std::max<int64_t>(1, std::min<int64_t>(64, npix / 256));
      auto kern = fmt ? wg::k_bias_grad_generic<__nv_bfloat16> : wg::k_bias_grad_generic<__half>;
      DFB_CHECK_CUDA(dfb_launch_pdl(kern, dim3(Cout / 64, splits), dim3(256), 0, st, true, (const uint16_t*)gO, npix, Cout, dB));
    }
    DFB_LAUNCH_CHECK();
  }
  return DFB_OK;
}

extern "C" int dfb_debug_umma_gemm_mn(const float* A, const float* B, int N, int K, int fmt_a, int fmt_b, int variant, float* D,
                                      void* stream) {
  DFB_REQUIRE(A && B && D, DFB_ERR_INVALID, "null argument");
  DFB_REQUIRE(N % 32 == 0 && N >= 32 && N <= 256 && K % 16 == 0 && K >= 16 && K <= 128, DFB_ERR_INVALID, "bad N / K");
  int* flag = nullptr;
  DFB_CHECK_CUDA(cudaMalloc(&flag, 4));
  DFB_CHECK_CUDA(cudaMemset(flag, 0, 4));
  const size_t smem = (size_t)K * 128 * 2 + (size_t)K * 256 * 2;
  cudaStream_t st = (cudaStream_t)stream;
  DFB_CHECK_CUDA(cudaFuncSetAttribute(wg::k_umma_selftest_mn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  wg::k_umma_selftest_mn<<<1, 128, smem, st>>>(A, B, N, K, fmt_a, fmt_b, variant, D, flag);
  DFB_LAUNCH_CHECK();
  DFB_CHECK_CUDA(cudaStreamSynchronize(st));
  cudaFree(flag);
  return DFB_OK;
}
