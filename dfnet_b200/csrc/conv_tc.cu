// Implicit-GEMM 2-D convolution on tcgen05 for the DFNet feature extractor
// (reference feature/dfnet.py:74-172: VGG-16 3x3 convs, 1x1 / 5x5 adaptation convs).
//
//   out[b,y,x,n] = act( bias[n] + sum_{ky,kx,c} in[b, y+ky-pad, x+kx-pad, c] * w[n,c,ky,kx] )
//
// GEMM view: M = B*H*W pixels (tiles of 128), N = Cout (tiles of <= 256), K = KH*KW*Cin ordered
// (ky,kx,c).  Activations are NHWC 16-bit with Cin % 8 == 0, so the 8 channels of one K panel of
// one pixel are 16 contiguous bytes: the A operand is gathered straight into the core-matrix
// panel layout of tc_common (no im2col buffer) with 16-byte cp.async (zero-filled at the image
// border), weights arrive as pre-packed chunks through cp.async.bulk, tcgen05.mma accumulates in
// TMEM (double buffered across tiles), and the epilogue applies bias / ReLU and writes NHWC 16-bit
// (next layer), an optional pre-activation tap, or fp32 NCHW (the layout the reference returns).
//
// Warps: 0-3 epilogue, 4-7 A gather (thread = pixel row), 8 weight producer, 9 MMA issuer.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "tc_common.cuh"

namespace dfb {
namespace conv {

using namespace dfb::tc;

constexpr int kTileM = 128;
constexpr int kStages = 4;
constexpr int kPanelsPerChunk = 8;               // K = 64 per stage
constexpr int kABytes = kPanelsPerChunk * 2048;  // 16 KB
constexpr int kThreads = 320;
constexpr int kLookahead = 2;                    // cp.async groups in flight per gather thread

enum Bar { A_FULL = 0, B_FULL = 4, EMPTY = 8, D_FULL = 12, D_EMPTY = 14, N_BARS = 16 };

struct ConvArgs {
  const __half* in;   // NHWC [B,H,W,Cin]
  const uint8_t* wimg;
  const float* bias;  // [Cout]
  __half* out;        // NHWC [B,H,W,Cout], after activation (nullable)
  __half* tap;        // NHWC [B,H,W,Cout], before activation (nullable)
  float* out_nchw;    // fp32 [B,Cout,H,W], before activation (nullable)
  int B, H, W, Cin, Cout, KH, KW, pad, relu;
  int nt;             // N tile (64, 128, 256)
  int n_ntiles, n_mtiles;
  int n_panels;       // K / 8 rounded up to even
  int real_panels;    // KH*KW*Cin/8
  int n_chunks;
  int64_t M;
  int* error_flag;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// bias / activation / stores of one 32-channel block of one pixel
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
  if (a.out_nchw) {
#pragma unroll
    for (int j = 0; j < 32; ++j) a.out_nchw[nchw_base + (int64_t)(n + j) * plane] = xv[j];
  }
  if (a.tap) {
    uint4* d = reinterpret_cast<uint4*>(a.tap + o);
#pragma unroll
    for (int q = 0; q < 4; ++q)
      d[q] = make_uint4(pack2<__half>(xv[8 * q], xv[8 * q + 1]), pack2<__half>(xv[8 * q + 2], xv[8 * q + 3]),
                        pack2<__half>(xv[8 * q + 4], xv[8 * q + 5]), pack2<__half>(xv[8 * q + 6], xv[8 * q + 7]));
  }
  if (a.out) {
    if (a.relu) {
#pragma unroll
      for (int j = 0; j < 32; ++j) xv[j] = fmaxf(xv[j], 0.f);
    }
    uint4* d = reinterpret_cast<uint4*>(a.out + o);
#pragma unroll
    for (int q = 0; q < 4; ++q)
      d[q] = make_uint4(pack2<__half>(xv[8 * q], xv[8 * q + 1]), pack2<__half>(xv[8 * q + 2], xv[8 * q + 3]),
                        pack2<__half>(xv[8 * q + 4], xv[8 * q + 5]), pack2<__half>(xv[8 * q + 6], xv[8 * q + 7]));
  }
}

__global__ void __launch_bounds__(kThreads, 1) k_conv_tc(const __grid_constant__ ConvArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int nt = a.nt;
  const uint32_t b_bytes = (uint32_t)nt * 16u * kPanelsPerChunk;
  const uint32_t stage_bytes = kABytes + b_bytes;
  const uint32_t s0 = smem_u32(smem);
  const uint32_t sBar = s0 + kStages * stage_bytes;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kStages * stage_bytes + N_BARS * 8);
  const int tid = threadIdx.x, warp = tid >> 5;
  auto bar = [&](int i) { return sBar + 8u * i; };

  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(bar(A_FULL + i), 128), mbar_init(bar(B_FULL + i), 1), mbar_init(bar(EMPTY + i), 1);
    }
    for (int i = 0; i < 2; ++i) mbar_init(bar(D_FULL + i), 1), mbar_init(bar(D_EMPTY + i), 128);
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc<1>(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_tiles = a.n_mtiles * a.n_ntiles;
  const int cpp = a.Cin >> 3;  // 8-channel panels per filter tap

  if (warp >= 4 && warp < 8) {
    // ===== A gather: thread = pixel row, 16-byte cp.async per (row, K panel) ==================
    const int r = tid - 128;
    uint32_t it = 0;  // global chunk counter -> stage / phase
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int mt = t / a.n_ntiles;
      const int64_t m = (int64_t)mt * kTileM + r;
      const bool mvalid = m < a.M;
      const int64_t mm = mvalid ? m : 0;
      const int x = (int)(mm % a.W);
      const int y = (int)((mm / a.W) % a.H);
      const int b = (int)(mm / ((int64_t)a.W * a.H));
      const __half* base = a.in + ((int64_t)b * a.H * a.W) * a.Cin;
      int tap = 0, c8 = 0;  // decomposition of the running panel index
      for (int c = 0; c < a.n_chunks + kLookahead; ++c) {
        if (c < a.n_chunks) {
          const uint32_t stage = (it + c) % kStages, ph = ((it + c) / kStages) & 1;
          mbar_wait(bar(EMPTY + stage), ph ^ 1, a.error_flag);
          const uint32_t dst = s0 + stage * stage_bytes + r * 16;
          const int np = min(kPanelsPerChunk, a.n_panels - c * kPanelsPerChunk);
          for (int p = 0; p < np; ++p) {
            const int ky = tap / a.KW, kx = tap - ky * a.KW;
            const int yy = y + ky - a.pad, xx = x + kx - a.pad;
            const bool ok = mvalid && tap < a.KH * a.KW && yy >= 0 && yy < a.H && xx >= 0 && xx < a.W;
            const __half* src = ok ? base + ((int64_t)yy * a.W + xx) * a.Cin + c8 * 8 : a.in;
            cp_async16(dst + p * 2048, src, ok ? 16u : 0u);
            if (++c8 == cpp) c8 = 0, ++tap;
          }
          cp_async_commit();
        } else {
          cp_async_commit();  // empty group keeps the wait_group arithmetic uniform
        }
        if (c >= kLookahead) {
          cp_async_wait<kLookahead>();
          fence_proxy_async();
          mbar_arrive(bar(A_FULL + (it + c - kLookahead) % kStages));
        }
      }
      it += a.n_chunks;
    }
  } else if (warp == 8) {
    // ===== weight producer ======================================================================
    uint32_t it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int ntile = t % a.n_ntiles;
      const uint8_t* src = a.wimg + (size_t)ntile * a.n_chunks * b_bytes;
      for (int c = 0; c < a.n_chunks; ++c, ++it) {
        const uint32_t stage = it % kStages, ph = (it / kStages) & 1;
        mbar_wait(bar(EMPTY + stage), ph ^ 1, a.error_flag);
        if (elect_one()) {
          mbar_expect_tx(bar(B_FULL + stage), b_bytes);
          bulk_g2s(s0 + stage * stage_bytes + kABytes, src + (size_t)c * b_bytes, b_bytes, bar(B_FULL + stage));
        }
        __syncwarp();
      }
    }
  } else if (warp == 9) {
    // ===== MMA issuer =============================================================================
    const uint32_t idesc = make_idesc(0, nt, kTileM);
    const uint32_t desc_hi = (128u >> 4) | (1u << 14);
    const uint32_t b_step = 2u * nt;
    uint32_t it = 0, tl = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tl) {
      const uint32_t buf = tl & 1;
      mbar_wait(bar(D_EMPTY + buf), ((tl >> 1) & 1) ^ 1, a.error_flag);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + buf * 256;
      uint32_t acc = 0;
      for (int c = 0; c < a.n_chunks; ++c, ++it) {
        const uint32_t stage = it % kStages, ph = (it / kStages) & 1;
        mbar_wait(bar(A_FULL + stage), ph, a.error_flag);
        mbar_wait(bar(B_FULL + stage), ph, a.error_flag);
        tc_fence_after();
        const uint32_t a_lo = ((s0 + stage * stage_bytes) >> 4) | ((2048u >> 4) << 16);
        const uint32_t b_lo = ((s0 + stage * stage_bytes + kABytes) >> 4) | ((uint32_t)nt << 16);
        const int ks_n = min(kPanelsPerChunk, a.n_panels - c * kPanelsPerChunk) >> 1;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            if (ks < ks_n) umma_f16<1>(d_tmem, mk64(a_lo + ks * 256, desc_hi), mk64(b_lo + ks * b_step, desc_hi), idesc, acc | ks);
          umma_commit<1>(bar(EMPTY + stage));
          if (c == a.n_chunks - 1) umma_commit<1>(bar(D_FULL + buf));
        }
        __syncwarp();
        acc = 1;
      }
    }
  } else if (warp < 4) {
    // ===== epilogue (thread = pixel) ===============================================================
    const int r = tid;
    uint32_t tl = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tl) {
      const uint32_t buf = tl & 1;
      const int mt = t / a.n_ntiles, n0 = (t % a.n_ntiles) * nt;
      const int64_t m = (int64_t)mt * kTileM + r;
      const bool valid = m < a.M;
      mbar_wait(bar(D_FULL + buf), (tl >> 1) & 1, a.error_flag);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16) + buf * 256;
      int64_t nchw_base = 0;
      if (a.out_nchw && valid) {
        const int x = (int)(m % a.W), y = (int)((m / a.W) % a.H);
        const int b = (int)(m / ((int64_t)a.W * a.H));
        nchw_base = ((int64_t)b * a.Cout * a.H + y) * a.W + x;
      }
      const int64_t plane = (int64_t)a.H * a.W;
      uint32_t v0[32], v1[32];
      tmem_ld32(t_row, v0);
#pragma unroll 1
      for (int cb = 0; cb < nt / 32; cb += 2) {  // nt is a multiple of 64: two blocks per trip
        tmem_ld_wait(v0);
        tmem_ld32(t_row + (cb + 1) * 32, v1);
        epi_store(a, v0, m, valid, n0 + cb * 32, nchw_base, plane);
        tmem_ld_wait(v1);
        if (cb + 2 < nt / 32) tmem_ld32(t_row + (cb + 2) * 32, v0);
        epi_store(a, v1, m, valid, n0 + (cb + 1) * 32, nchw_base, plane);
      }
      tc_fence_before();
      mbar_arrive(bar(D_EMPTY + buf));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

}  // namespace conv
}  // namespace dfb

// ------------------------------------------------------------------------------------------
// host: handle, weight packing, launch
// ------------------------------------------------------------------------------------------
struct DfbConv {
  int Cin, Cin_pad, Cout, KH, KW, pad, nt, n_ntiles, n_panels, real_panels, n_chunks;
  uint8_t* wimg = nullptr;
  float* bias = nullptr;
  int num_sms = 0;
};

using namespace dfb;

static uint16_t f2h16(float f) { __half h = __float2half_rn(f); uint16_t u; memcpy(&u, &h, 2); return u; }

extern "C" int dfb_conv_create(int Cin, int Cout, int KH, int KW, const float* weight, const float* bias,
                               const float* bn_scale, const float* bn_shift, DfbConv** out) {
  DFB_REQUIRE(weight && out && Cin >= 1 && Cout >= 64 && Cout % 64 == 0, DFB_ERR_INVALID,
              "dfb_conv_create: Cout must be a positive multiple of 64");
  DFB_REQUIRE(KH == KW && (KH == 1 || KH == 3 || KH == 5), DFB_ERR_UNSUPPORTED, "kernel size must be 1, 3 or 5");
  DfbConv* c = new DfbConv();
  c->Cin = Cin, c->Cin_pad = round_up(Cin, 8), c->Cout = Cout, c->KH = KH, c->KW = KW, c->pad = KH / 2;
  c->nt = Cout % 256 == 0 ? 256 : (Cout % 128 == 0 ? 128 : 64);
  c->n_ntiles = Cout / c->nt;
  c->real_panels = KH * KW * c->Cin_pad / 8;
  c->n_panels = round_up(c->real_panels, 2);
  c->n_chunks = (c->n_panels + conv::kPanelsPerChunk - 1) / conv::kPanelsPerChunk;
  int dev = 0;
  DFB_CHECK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  DFB_CHECK_CUDA(cudaGetDeviceProperties(&p, dev));
  DFB_REQUIRE(p.major == 10, DFB_ERR_UNSUPPORTED, "sm_100a code only");
  c->num_sms = p.multiProcessorCount;
  const size_t wn = (size_t)Cout * Cin * KH * KW;
  std::vector<float> w(wn), b(Cout, 0.f), sc(Cout, 1.f), sh(Cout, 0.f);
  DFB_CHECK_CUDA(cudaMemcpy(w.data(), weight, wn * 4, cudaMemcpyDefault));
  if (bias) DFB_CHECK_CUDA(cudaMemcpy(b.data(), bias, Cout * 4, cudaMemcpyDefault));
  if (bn_scale) DFB_CHECK_CUDA(cudaMemcpy(sc.data(), bn_scale, Cout * 4, cudaMemcpyDefault));
  if (bn_shift) DFB_CHECK_CUDA(cudaMemcpy(sh.data(), bn_shift, Cout * 4, cudaMemcpyDefault));
  // eval-mode BatchNorm folded into the conv: y = scale * (conv + bias) + shift
  std::vector<float> bf(Cout);
  for (int n = 0; n < Cout; ++n) bf[n] = sc[n] * b[n] + sh[n];
  const size_t b_bytes = (size_t)c->nt * 16 * conv::kPanelsPerChunk;
  std::vector<uint16_t> img((size_t)c->n_ntiles * c->n_chunks * b_bytes / 2, 0);
  for (int t = 0; t < c->n_ntiles; ++t)
    for (int ch = 0; ch < c->n_chunks; ++ch)
      for (int pp = 0; pp < conv::kPanelsPerChunk; ++pp) {
        const int P = ch * conv::kPanelsPerChunk + pp;
        if (P >= c->real_panels) continue;
        const int cpp = c->Cin_pad / 8, tap = P / cpp, c8 = P % cpp, ky = tap / KW, kx = tap % KW;
        for (int rr = 0; rr < c->nt; ++rr)
          for (int e = 0; e < 8; ++e) {
            const int ci = c8 * 8 + e, n = t * c->nt + rr;
            const float v = ci < Cin ? sc[n] * w[(((size_t)n * Cin + ci) * KH + ky) * KW + kx] : 0.f;
            img[((size_t)(t * c->n_chunks + ch) * b_bytes) / 2 + (size_t)pp * c->nt * 8 + (size_t)rr * 8 + e] = f2h16(v);
          }
      }
  DFB_CHECK_CUDA(cudaMalloc(&c->wimg, img.size() * 2));
  DFB_CHECK_CUDA(cudaMemcpy(c->wimg, img.data(), img.size() * 2, cudaMemcpyHostToDevice));
  DFB_CHECK_CUDA(cudaMalloc(&c->bias, Cout * 4));
  DFB_CHECK_CUDA(cudaMemcpy(c->bias, bf.data(), Cout * 4, cudaMemcpyHostToDevice));
  *out = c;
  return DFB_OK;
}

extern "C" void dfb_conv_destroy(DfbConv* c) {
  if (!c) return;
  if (c->wimg) cudaFree(c->wimg);
  if (c->bias) cudaFree(c->bias);
  delete c;
}

static int* g_conv_error_flag = nullptr;

extern "C" int dfb_conv_fwd(DfbConv* c, const void* in_nhwc16, int B, int H, int W, int relu, void* out_nhwc16,
                            void* tap_nhwc16, float* out_nchw32, void* stream) {
  DFB_REQUIRE(c && in_nhwc16 && (out_nhwc16 || tap_nhwc16 || out_nchw32), DFB_ERR_INVALID, "dfb_conv_fwd: null argument");
  DFB_REQUIRE(B >= 1 && H >= 1 && W >= 1, DFB_ERR_INVALID, "bad image size");
  if (!g_conv_error_flag) {
    DFB_CHECK_CUDA(cudaMalloc(&g_conv_error_flag, 4));
    DFB_CHECK_CUDA(cudaMemset(g_conv_error_flag, 0, 4));
  }
  conv::ConvArgs a = {};
  a.in = (const __half*)in_nhwc16, a.wimg = c->wimg, a.bias = c->bias;
  a.out = (__half*)out_nhwc16, a.tap = (__half*)tap_nhwc16, a.out_nchw = out_nchw32;
  a.B = B, a.H = H, a.W = W, a.Cin = c->Cin_pad, a.Cout = c->Cout, a.KH = c->KH, a.KW = c->KW, a.pad = c->pad, a.relu = relu;
  a.nt = c->nt, a.n_ntiles = c->n_ntiles, a.n_panels = c->n_panels, a.real_panels = c->real_panels, a.n_chunks = c->n_chunks;
  a.M = (int64_t)B * H * W;
  a.n_mtiles = (int)((a.M + conv::kTileM - 1) / conv::kTileM);
  a.error_flag = g_conv_error_flag;
  const int n_tiles = a.n_mtiles * a.n_ntiles;
  const int grid = std::min(n_tiles, c->num_sms);
  const size_t smem = (size_t)conv::kStages * (conv::kABytes + (size_t)c->nt * 16 * conv::kPanelsPerChunk) + 256;
  DFB_CHECK_CUDA(cudaFuncSetAttribute(conv::k_conv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  conv::k_conv_tc<<<grid, conv::kThreads, smem, (cudaStream_t)stream>>>(a);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}
