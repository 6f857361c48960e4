// Backward pass of DFNet for the direct-feature-matching training step (reference
// feature/direct_feature_matching.py:322-390 `train_on_batch`, `loss.backward()` at :378):
//
//   * feature net G (frozen, eval): gradient of the feature stacks w.r.t. the input images, i.e. the
//     data-gradient chain  upsample^T -> BatchNorm/5x5^T -> ReLU' -> 1x1^T -> (+ encoder chain) -> conv^T ... -> 1/std
//   * pose regressor F (trained): gradient of the pose w.r.t. every encoder / fc parameter, i.e.
//     fc^T -> avgpool^T -> maxpool^T -> per layer { weight gradient, bias gradient, data gradient }
//
// Gradients travel as NHWC bf16 (fp32 accumulation inside the tensor-core kernels); the data-gradient
// convolutions are conv_tc.cu run on transposed/flipped filters with the ReLU mask applied in the
// epilogue, the weight gradients are conv_bwd_tc.cu.  Everything reads the tape a forward with
// flags bit3 left behind (dfnet_kernels.cu).
#include <algorithm>

#include "dfnet_handle.cuh"
#include "tc_common.cuh"

namespace dfb {

__device__ __forceinline__ float bf16_bits_to_float(uint32_t u) { return __uint_as_float(u << 16); }

// fp32 NCHW [n,C,h,w] -> bf16 NHWC [n,h,w,C]; thread = (pixel, 8-channel group), pixels fastest so the reads coalesce
__global__ void k_nchw32_to_nhwc_bf16(const float* __restrict__ src, uint16_t* __restrict__ dst, int64_t npix, int64_t plane, int C) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int C8 = C / 8;
  if (i >= npix * C8) return;
  const int64_t pix = i % npix;
  const int c8 = (int)(i / npix);
  const int64_t b = pix / plane, p = pix % plane;
  const float* s = src + (b * C + c8 * 8) * plane + p;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = __ldg(s + j * plane);
  uint4 o;
  o.x = tc::pack2<__nv_bfloat16>(v[0], v[1]), o.y = tc::pack2<__nv_bfloat16>(v[2], v[3]);
  o.z = tc::pack2<__nv_bfloat16>(v[4], v[5]), o.w = tc::pack2<__nv_bfloat16>(v[6], v[7]);
  *reinterpret_cast<uint4*>(dst + pix * C + c8 * 8) = o;
}

// adjoint of k_resize_bilinear_ac (align_corners=True): scatter every output gradient to its 4 sources
__global__ void k_resize_bilinear_ac_bwd(const float* __restrict__ gdst, float* __restrict__ gsrc, int planes, int h, int w, int Ho,
                                         int Wo) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)planes * Ho * Wo) return;
  const int xo = (int)(i % Wo), yo = (int)((i / Wo) % Ho);
  const int64_t pl = i / ((int64_t)Wo * Ho);
  const float sy = Ho > 1 ? (float)(h - 1) / (float)(Ho - 1) : 0.f, sx = Wo > 1 ? (float)(w - 1) / (float)(Wo - 1) : 0.f;
  const float fy = sy * yo, fx = sx * xo;
  const int y0 = (int)fy, y1 = y0 + (y0 < h - 1), x0 = (int)fx, x1 = x0 + (x0 < w - 1);
  const float ly = fy - y0, lx = fx - x0;
  const float g = __ldg(gdst + i);
  float* r0 = gsrc + (pl * h + y0) * w;
  float* r1 = gsrc + (pl * h + y1) * w;
  atomicAdd(r0 + x0, (1.f - ly) * (1.f - lx) * g);
  atomicAdd(r0 + x1, (1.f - ly) * lx * g);
  atomicAdd(r1 + x0, ly * (1.f - lx) * g);
  atomicAdd(r1 + x1, ly * lx * g);
}

// out = addend (or 0): initialises the rows / columns a floor-mode 2x2 pooling never reads
__global__ void k_fill16(uint4* __restrict__ out, const uint4* __restrict__ addend, int64_t n16) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n16) out[i] = addend ? addend[i] : make_uint4(0, 0, 0, 0);
}

// 2x2/2 max-pool backward fused with the ReLU mask of the pooled activation:
//   gA[window argmax] = gP  (first maximum in scan order, like ATen), zero elsewhere and where act <= 0;  + addend
// act: post-ReLU NHWC 16-bit (fp16 or bf16 patterns order like their values), gP / addend / gA: bf16.
__global__ void k_maxpool2x2_bwd(const uint16_t* __restrict__ act, const uint16_t* __restrict__ gP, const uint16_t* __restrict__ addend,
                                 uint16_t* __restrict__ gA, int B, int H, int W, int C) {
  const int Ho = H / 2, Wo = W / 2, C8 = C / 8;
  const int64_t n = (int64_t)B * Ho * Wo * C8;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c8 = (int)(i % C8);
  const int xo = (int)((i / C8) % Wo), yo = (int)((i / ((int64_t)C8 * Wo)) % Ho), b = (int)(i / ((int64_t)C8 * Wo * Ho));
  const int64_t o00 = (((int64_t)b * H + 2 * yo) * W + 2 * xo) * C + c8 * 8;
  const int64_t offs[4] = {o00, o00 + C, o00 + (int64_t)W * C, o00 + (int64_t)W * C + C};
  uint4 av[4], gv = *reinterpret_cast<const uint4*>(gP + i * 8), ov[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) av[k] = *reinterpret_cast<const uint4*>(act + offs[k]);
  const uint32_t* g32 = reinterpret_cast<const uint32_t*>(&gv);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t res[4] = {0, 0, 0, 0};
#pragma unroll
    for (int hlf = 0; hlf < 2; ++hlf) {
      const int sh = hlf * 16;
      int best = 0;
      int32_t bv = -0x10000;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t p = (reinterpret_cast<const uint32_t*>(&av[k])[q] >> sh) & 0xffffu;
        const int32_t ps = (p & 0x8000u) ? -(int32_t)(p & 0x7fffu) : (int32_t)p;
        if (ps > bv) bv = ps, best = k;
      }
      if (bv > 0) res[best] |= ((g32[q] >> sh) & 0xffffu) << sh;  // ReLU': only a strictly positive maximum passes
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) reinterpret_cast<uint32_t*>(&ov[k])[q] = res[k];
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (addend) {
      const uint4 ad = *reinterpret_cast<const uint4*>(addend + offs[k]);
      const uint32_t* a32 = reinterpret_cast<const uint32_t*>(&ad);
      uint32_t* o32 = reinterpret_cast<uint32_t*>(&ov[k]);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        o32[q] = tc::pack2<__nv_bfloat16>(bf16_bits_to_float(o32[q] & 0xffffu) + bf16_bits_to_float(a32[q] & 0xffffu),
                                          bf16_bits_to_float(o32[q] >> 16) + bf16_bits_to_float(a32[q] >> 16));
    }
    *reinterpret_cast<uint4*>(gA + offs[k]) = ov[k];
  }
}

// fp16 -> bf16 copy of an activation (the weight-gradient MMA needs both operands in one format)
__global__ void k_f16_to_bf16(const uint4* __restrict__ in, uint4* __restrict__ out, int64_t n16) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n16) return;
  const uint4 v = in[i];
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  uint32_t o[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[q]));
    o[q] = tc::pack2<__nv_bfloat16>(f.x, f.y);
  }
  out[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

// AdaptiveAvgPool2d(1) backward: g[b,p,c] = g_pooled[b,c] / HW, bf16 NHWC
__global__ void k_avgpool_bwd(const float* __restrict__ g_pooled, uint16_t* __restrict__ out, int HW, int C, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = (int)(i % C);
  const int64_t b = i / ((int64_t)C * HW);
  out[i] = __bfloat16_as_ushort(__float2bfloat16_rn(g_pooled[b * C + c] / (float)HW));
}

// Linear(512,12) backward; one block of 512 threads (thread = input feature k)
__global__ void k_fc_bwd(const float* __restrict__ g_pose, const float* __restrict__ pooled, const float* __restrict__ w, int B,
                         float* __restrict__ gW, float* __restrict__ gb, float* __restrict__ g_pooled) {
  const int k = threadIdx.x;
  for (int o = 0; o < 12; ++o) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s = fmaf(g_pose[b * 12 + o], pooled[b * 512 + k], s);
    if (gW) gW[o * 512 + k] = s;
  }
  for (int b = 0; b < B; ++b) {
    float s = 0.f;
    for (int o = 0; o < 12; ++o) s = fmaf(g_pose[b * 12 + o], w[o * 512 + k], s);
    g_pooled[b * 512 + k] = s;
  }
  if (gb && k < 12) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += g_pose[b * 12 + k];
    gb[k] = s;
  }
}

// BatchNorm backward of the adaptation heads (y = gamma * zhat + beta, zhat = (z - mean) * rstd), fp32 NCHW [B,128,plane].
// Pass 1: per channel S1 = sum g_y, S2 = sum g_y * zhat (double partial sums, [128][kBnBwdSplits][2]).
constexpr int kBnBwdSplits = 32;
__global__ void __launch_bounds__(256) k_bn_bwd_partial(const float* __restrict__ gy, const float* __restrict__ z, int B, int64_t plane,
                                                        const float* __restrict__ stat, float eps, double* __restrict__ part) {
  __shared__ double sh[2][8];
  const int c = blockIdx.y, sp = blockIdx.x;
  const float mean = stat[c], rstd = rsqrtf(stat[128 + c] + eps);
  const int64_t n = (int64_t)B * plane;
  double s1 = 0.0, s2 = 0.0;
  for (int64_t i = (int64_t)sp * 256 + threadIdx.x; i < n; i += (int64_t)kBnBwdSplits * 256) {
    const int64_t b = i / plane, p = i - b * plane;
    const int64_t at = (b * 128 + c) * plane + p;
    const float g = gy[at];
    s1 += (double)g, s2 += (double)(g * ((z[at] - mean) * rstd));
  }
  for (int o = 16; o > 0; o >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, o), s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  if ((threadIdx.x & 31) == 0) sh[0][threadIdx.x >> 5] = s1, sh[1][threadIdx.x >> 5] = s2;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b2 = 0.0;
    for (int w = 0; w < 8; ++w) a += sh[0][w], b2 += sh[1][w];
    part[((size_t)c * kBnBwdSplits + sp) * 2] = a, part[((size_t)c * kBnBwdSplits + sp) * 2 + 1] = b2;
  }
}
// sums[c] = {S1, S2}; g_gamma = S2, g_beta = S1 (nullable outputs)
__global__ void k_bn_bwd_finalize(const double* __restrict__ part, float* __restrict__ sums, float* __restrict__ g_gamma,
                                  float* __restrict__ g_beta) {
  const int c = threadIdx.x;
  double s1 = 0.0, s2 = 0.0;
  for (int i = 0; i < kBnBwdSplits; ++i) s1 += part[((size_t)c * kBnBwdSplits + i) * 2], s2 += part[((size_t)c * kBnBwdSplits + i) * 2 + 1];
  sums[c] = (float)s1, sums[128 + c] = (float)s2;
  if (g_gamma) g_gamma[c] = (float)s2;
  if (g_beta) g_beta[c] = (float)s1;
}
// Pass 2: g_z = gamma * rstd * (g_y - [batch statistics] (S1 + zhat * S2) / M) -> bf16 NHWC [B,plane,128].
// gamma * rstd = scale (stat[256 + c]).  thread = (pixel, 8-channel group), pixels fastest.
__global__ void k_bn_bwd_apply(const float* __restrict__ gy, const float* __restrict__ z, int64_t npix, int64_t plane,
                               const float* __restrict__ stat, const float* __restrict__ sums, float eps, int batch_stats, float inv_m,
                               uint16_t* __restrict__ gz) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix * 16) return;
  const int64_t pix = i % npix;
  const int c8 = (int)(i / npix);
  const int64_t b = pix / plane, p = pix % plane;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c8 * 8 + j;
    const int64_t at = (b * 128 + c) * plane + p;
    float g = __ldg(gy + at);
    if (batch_stats) {
      const float zh = (__ldg(z + at) - stat[c]) * rsqrtf(stat[128 + c] + eps);
      g -= (sums[c] + zh * sums[128 + c]) * inv_m;
    }
    v[j] = g * stat[256 + c];
  }
  uint4 o;
  o.x = tc::pack2<__nv_bfloat16>(v[0], v[1]), o.y = tc::pack2<__nv_bfloat16>(v[2], v[3]);
  o.z = tc::pack2<__nv_bfloat16>(v[4], v[5]), o.w = tc::pack2<__nv_bfloat16>(v[6], v[7]);
  *reinterpret_cast<uint4*>(gz + pix * 128 + c8 * 8) = o;
}

// input normalisation backward: g_x[b,c] = g_norm[b,c] / std[c]
__global__ void k_unnorm_grad(float* __restrict__ g, int64_t plane, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = (int)((i / plane) % 3);
  g[i] = __fdiv_rn(g[i], c == 0 ? 0.229f : (c == 1 ? 0.224f : 0.225f));
}

}  // namespace dfb

using namespace dfb;

namespace {
struct BwdWs { size_t gA, gB, gC, gtap[3], gmid, g16, fstage, gpooled, xcvt, bnpart, bnsums, total; };

BwdWs bwd_ws(int nb, int H, int W) {
  BwdWs w = {};
  size_t off = 0;
  auto take = [&](size_t b) { size_t o = off; off += (b + 255) / 256 * 256; return o; };
  const size_t px = (size_t)nb * H * W;
  w.gA = take(px * 64 * 2), w.gB = take(px * 64 * 2), w.gC = take(px * 64 * 2);
  w.gtap[0] = take(px * 64 * 2);
  w.gtap[1] = take((size_t)nb * (H / 4) * (W / 4) * 256 * 2);
  w.gtap[2] = take((size_t)nb * (H / 16) * (W / 16) * 512 * 2);
  w.gmid = take(px * 64 * 2);
  w.g16 = take(px * 128 * 2);
  w.fstage = take(px * 128 * 4);  // level 0 is resampled too when upsampleH/W differ from the input size
  w.gpooled = take((size_t)nb * 512 * 4 + 256);
  w.xcvt = take(px * 64 * 2);
  w.bnpart = take((size_t)128 * kBnBwdSplits * 2 * sizeof(double));
  w.bnsums = take(256 * sizeof(float));
  w.total = off;
  return w;
}
}  // namespace

extern "C" int dfb_dfnet_bwd_workspace_bytes(const DfbDfnet* d, int B, int H, int W, size_t* out) {
  DFB_REQUIRE(d && out && B >= 1 && H >= 32 && W >= 32, DFB_ERR_INVALID, "bad arguments");
  *out = bwd_ws(B, H, W).total;
  return DFB_OK;
}

// flags: the forward's (bit0 return_feature, bit1 single_stream, bit2 return_pose, bit4 bf16 tape).
// g_feats_t / g_feats_r: gradients of the two feature stacks [L,Bs,128,upH,upW] (either may be null: that stream is
//   skipped entirely — train_on_batch only differentiates the rendered stream); level_mask bit l = level l carries gradient.
// g_pose [B,12] (nullable).  g_x [nb,3,H,W] fp32 (nullable): gradient w.r.t. the images of the differentiated sub-batch
//   (both streams: all B; one stream: its B/2 images).  g_params: n_params pointers in dfb_dfnet_load order (nullable
//   entries); only the encoder and fc_pose entries are written.
extern "C" int dfb_dfnet_bwd(DfbDfnet* d, int B, int H, int W, uint32_t flags, int upH, int upW, const float* g_feats_t,
                             const float* g_feats_r, uint32_t level_mask, const float* g_pose, const void* tape, float* g_x,
                             float* const* g_params, int n_params, void* scratch, size_t scratch_bytes, void* stream) {
  DFB_REQUIRE(d && d->loaded && tape && scratch, DFB_ERR_INVALID, "dfb_dfnet_bwd: null argument");
  const bool ret_feat = flags & 1, single = flags & 2, ret_pose = flags & 4, bf = flags & 16;
  const bool feat_grad = ret_feat && (g_feats_t || g_feats_r) && level_mask;
  const bool pose_grad = ret_pose && g_pose;
  DFB_REQUIRE(feat_grad || pose_grad, DFB_ERR_INVALID, "no gradient given");
  // flags bit6: the forward ran the heads un-folded and kept their pre-BatchNorm outputs: the heads (1x1, 5x5, BatchNorm
  // affine) are differentiated too (run_feature.py training); bit5 additionally means batch statistics
  const bool head_tape = (flags & 64) != 0, bn_batch = (flags & 32) != 0;
  DFB_REQUIRE(d->enc_dg[0], DFB_ERR_INVALID, "training variants not loaded (dfb_dfnet_load_ex flags bit0)");
  DFB_REQUIRE(!g_params || pose_grad || (feat_grad && head_tape), DFB_ERR_UNSUPPORTED,
              "parameter gradients of the feature path need a forward with the head tape (flags bit6)");
  DFB_REQUIRE(!(feat_grad && bn_batch) || head_tape, DFB_ERR_UNSUPPORTED,
              "the backward through train-mode BatchNorm needs a forward with the head tape (flags bit6)");
  DFB_REQUIRE(!head_tape || !feat_grad || single || (g_feats_t && g_feats_r), DFB_ERR_UNSUPPORTED,
              "training the heads differentiates both streams of a siamese forward");
  DFB_REQUIRE(!head_tape || !feat_grad || d->head5_raw_dg[0], DFB_ERR_INVALID,
              "head training variants not loaded (dfb_dfnet_load_ex flags bit3)");
  DFB_REQUIRE(!g_params || n_params == 26 + 8 * d->n_levels + 2, DFB_ERR_INVALID, "g_params has the wrong length");
  cudaStream_t st = (cudaStream_t)stream;
  const DfWs L = dfnet_ws(B, H, W, d->n_levels, upH, upW, true);
  // differentiated sub-batch
  int b0 = 0, nb = B;
  if (feat_grad && !single) {
    DFB_REQUIRE(B % 2 == 0, DFB_ERR_INVALID, "siamese mode needs an even batch");
    DFB_REQUIRE(!pose_grad || (g_feats_t && g_feats_r), DFB_ERR_UNSUPPORTED,
                "pose and feature gradients in one call differentiate the whole batch (both streams)");
    if (!g_feats_t) b0 = B / 2, nb = B / 2;
    else if (!g_feats_r) nb = B / 2;
  }
  const BwdWs S = bwd_ws(nb, H, W);
  DFB_REQUIRE(scratch_bytes >= S.total, DFB_ERR_WORKSPACE, "scratch too small: need %zu bytes", S.total);
  char* sc = (char*)scratch;
  const char* tp = (const char*)tape;
  auto act_ptr = [&](int i) { return tp + L.act[i] + (size_t)b0 * L.h[i] * L.w[i] * kEncCout[i] * 2; };

  const uint16_t* gtap[3] = {nullptr, nullptr, nullptr};
  int i_start = -1;
  const uint16_t* cur = nullptr;

  if (feat_grad) {
    const int Bs = single ? B : B / 2;
    for (int l = 0; l < d->n_levels; ++l) {
      if (!(level_mask & (1u << l))) continue;
      const int ci = kTapConv[l], fh = L.h[ci], fw = L.w[ci];
      const size_t lvl_stride = (size_t)Bs * 128 * upH * upW;
      // pieces of the sub-batch: (gradient stack, first image inside the sub-batch, images)
      struct Piece { const float* g; int at, n; } pieces[2];
      int np = 0;
      if (single) pieces[np++] = {g_feats_t, 0, B};
      else {
        if (g_feats_t) pieces[np++] = {g_feats_t, 0, Bs};
        if (g_feats_r) pieces[np++] = {g_feats_r, g_feats_t ? Bs : 0, Bs};
      }
      uint16_t* g16 = (uint16_t*)(sc + S.g16);
      if (head_tape) {
        // ---- heads being trained: BatchNorm backward on the kept pre-BatchNorm output, then weight + data gradients ----
        const int64_t fplane = (int64_t)fh * fw, npix_all = (int64_t)nb * fplane;
        float* gy = (float*)(sc + S.fstage);   // d loss / d y for ALL images of the call, fp32 NCHW [nb,128,fh,fw]
        for (int p = 0; p < np; ++p) {
          const float* src = pieces[p].g + l * lvl_stride;
          float* dstp = gy + (size_t)pieces[p].at * 128 * fplane;
          const int n = pieces[p].n;
          if (fh != upH || fw != upW) {
            DFB_CHECK_CUDA(cudaMemsetAsync(dstp, 0, (size_t)n * 128 * fplane * 4, st));
            const int64_t tot = (int64_t)n * 128 * upH * upW;
            k_resize_bilinear_ac_bwd<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(src, dstp, n * 128, fh, fw, upH, upW);
            DFB_LAUNCH_CHECK();
          } else {
            DFB_CHECK_CUDA(cudaMemcpyAsync(dstp, src, (size_t)n * 128 * fplane * 4, cudaMemcpyDeviceToDevice, st));
          }
        }
        const float* z = (const float*)(tp + L.zbn[l]);
        const float* stat = d->bn_stat + l * 4 * 128;
        double* part = (double*)(sc + S.bnpart);
        float* sums = (float*)(sc + S.bnsums);
        float* const* hp = g_params ? g_params + 26 + 8 * l : nullptr;
        k_bn_bwd_partial<<<dim3(kBnBwdSplits, 128), 256, 0, st>>>(gy, z, nb, fplane, stat, d->bn_eps, part);
        DFB_LAUNCH_CHECK();
        k_bn_bwd_finalize<<<1, 128, 0, st>>>(part, sums, hp ? hp[4] : nullptr, hp ? hp[5] : nullptr);
        DFB_LAUNCH_CHECK();
        k_bn_bwd_apply<<<(unsigned)((npix_all * 16 + 255) / 256), 256, 0, st>>>(gy, z, npix_all, fplane, stat, sums, d->bn_eps,
                                                                              bn_batch ? 1 : 0, 1.f / (float)npix_all, g16);
        DFB_LAUNCH_CHECK();
        const char* mid = tp + L.mid[l];
        if (hp && hp[2]) {   // 5x5 conv: weight / bias gradient (input = the ReLU output of the 1x1 conv)
          const int64_t n16 = npix_all * 64 / 8;
          k_f16_to_bf16<<<(unsigned)((n16 + 255) / 256), 256, 0, st>>>((const uint4*)mid, (uint4*)(sc + S.xcvt), n16);
          DFB_LAUNCH_CHECK();
          int rc = dfb_conv_wgrad(g16, sc + S.xcvt, nb, fh, fw, 64, 64, 128, 5, 1, hp[2], hp[3], stream);
          if (rc) return rc;
          // batch statistics remove the per-channel mean, so d loss / d (5x5 bias) = sum g_z is exactly 0; the sum of the
          // bf16-rounded g_z is only rounding residue
          if (bn_batch && hp[3]) DFB_CHECK_CUDA(cudaMemsetAsync(hp[3], 0, 128 * sizeof(float), st));
        }
        int rc = dfb_conv_run(d->head5_raw_dg[l], g16, nb, fh, fw, 0, sc + S.gmid, nullptr, nullptr, mid, nullptr, stream);
        if (rc) return rc;
        if (hp && hp[0]) {   // 1x1 conv: weight / bias gradient (input = the pre-ReLU tap of the encoder)
          const int C = kTapCh[l];
          const int64_t n16 = npix_all * C / 8;
          k_f16_to_bf16<<<(unsigned)((n16 + 255) / 256), 256, 0, st>>>((const uint4*)(tp + L.tap[l]), (uint4*)(sc + S.xcvt), n16);
          DFB_LAUNCH_CHECK();
          rc = dfb_conv_wgrad(sc + S.gmid, sc + S.xcvt, nb, fh, fw, C, C, 64, 1, 1, hp[0], hp[1], stream);
          if (rc) return rc;
        }
        rc = dfb_conv_run(d->head1_dg[l], sc + S.gmid, nb, fh, fw, 0, sc + S.gtap[l], nullptr, nullptr, nullptr, nullptr, stream);
        if (rc) return rc;
        gtap[l] = (const uint16_t*)(sc + S.gtap[l]);
        i_start = ci;
        continue;
      }
      for (int p = 0; p < np; ++p) {
        const float* src = pieces[p].g + l * lvl_stride;
        const int n = pieces[p].n;
        if (fh != upH || fw != upW) {
          float* fs = (float*)(sc + S.fstage);
          DFB_CHECK_CUDA(cudaMemsetAsync(fs, 0, (size_t)n * 128 * fh * fw * 4, st));
          const int64_t tot = (int64_t)n * 128 * upH * upW;
          k_resize_bilinear_ac_bwd<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(src, fs, n * 128, fh, fw, upH, upW);
          DFB_LAUNCH_CHECK();
          src = fs;
        }
        const int64_t npix = (int64_t)n * fh * fw;
        k_nchw32_to_nhwc_bf16<<<(unsigned)((npix * 16 + 255) / 256), 256, 0, st>>>(src, g16 + (size_t)pieces[p].at * fh * fw * 128, npix,
                                                                                  (int64_t)fh * fw, 128);
        DFB_LAUNCH_CHECK();
      }
      // BatchNorm(eval) + 5x5 conv transposed, ReLU mask of the 1x1 output, then the 1x1 conv transposed
      const char* mid = tp + L.mid[l] + (size_t)b0 * fh * fw * 64 * 2;
      int rc = dfb_conv_run(d->head5_dg[l], g16, nb, fh, fw, 0, sc + S.gmid, nullptr, nullptr, mid, nullptr, stream);
      if (rc) return rc;
      rc = dfb_conv_run(d->head1_dg[l], sc + S.gmid, nb, fh, fw, 0, sc + S.gtap[l], nullptr, nullptr, nullptr, nullptr, stream);
      if (rc) return rc;
      gtap[l] = (const uint16_t*)(sc + S.gtap[l]);
      i_start = ci;
    }
    cur = gtap[i_start == 12 ? 2 : (i_start == 6 ? 1 : 0)];  // gradient w.r.t. the pre-activation of the deepest tapped conv
  }

  void* bufs[2] = {sc + S.gA, sc + S.gB};
  int flip = 0;
  if (pose_grad) {
    const int o = 26 + 8 * d->n_levels;
    float* gpooled = (float*)(sc + S.gpooled);
    k_fc_bwd<<<1, 512, 0, st>>>(g_pose, (const float*)(tp + L.pooled), d->fc_w, B, g_params ? g_params[o] : nullptr,
                                g_params ? g_params[o + 1] : nullptr, gpooled);
    DFB_LAUNCH_CHECK();
    const int h5 = L.h[12] / 2, w5 = L.w[12] / 2;
    const int64_t n5 = (int64_t)B * h5 * w5 * 512;
    k_avgpool_bwd<<<(unsigned)((n5 + 255) / 256), 256, 0, st>>>(gpooled, (uint16_t*)(sc + S.gC), h5 * w5, 512, n5);
    DFB_LAUNCH_CHECK();
    uint16_t* g12 = (uint16_t*)bufs[flip];
    const uint16_t* add12 = gtap[2];   // feature gradient of conv5_3's pre-activation tap (combined feature + pose backward)
    if ((L.h[12] | L.w[12]) & 1) {
      const int64_t n16 = (int64_t)B * L.h[12] * L.w[12] * 512 / 8;
      k_fill16<<<(unsigned)((n16 + 255) / 256), 256, 0, st>>>((uint4*)g12, (const uint4*)add12, n16);
      DFB_LAUNCH_CHECK();
    }
    const int64_t nw = (int64_t)B * h5 * w5 * (512 / 8);
    k_maxpool2x2_bwd<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>((const uint16_t*)act_ptr(12), (const uint16_t*)(sc + S.gC), add12, g12,
                                                                  B, L.h[12], L.w[12], 512);
    DFB_LAUNCH_CHECK();
    cur = g12, flip ^= 1, i_start = 12;
  }

  for (int i = i_start; i >= 0; --i) {
    const int h = L.h[i], w = L.w[i];
    if (g_params && (g_params[2 * i] || g_params[2 * i + 1])) {
      DFB_REQUIRE(g_params[2 * i], DFB_ERR_INVALID, "bias gradient without weight gradient");
      const void* X = i == 0 ? tp + L.in8 : (kPoolAfter[i - 1] ? tp + L.pool[i - 1] : tp + L.act[i - 1]);
      if (!bf) {  // fp16 tape: the gradient is bf16, so hand the MMA a bf16 copy of the layer input
        const int64_t n16 = (int64_t)B * h * w * (i == 0 ? 8 : kEncCin[i]) / 8;
        k_f16_to_bf16<<<(unsigned)((n16 + 255) / 256), 256, 0, st>>>((const uint4*)X, (uint4*)(sc + S.xcvt), n16);
        DFB_LAUNCH_CHECK();
        X = sc + S.xcvt;
      }
      int rc = dfb_conv_wgrad(cur, X, B, h, w, kEncCin[i], i == 0 ? 8 : kEncCin[i], kEncCout[i], 3, 1, g_params[2 * i],
                              g_params[2 * i + 1], stream);
      if (rc) return rc;
    }
    if (g_params && d->bucket_event && i == d->bucket_first_layer)
      DFB_CHECK_CUDA(cudaEventRecord((cudaEvent_t)d->bucket_event, st));
    if (i == 0) {
      if (g_x) {
        int rc = dfb_conv_run(d->enc_dg[0], cur, nb, h, w, 0, nullptr, nullptr, g_x, nullptr, nullptr, stream);
        if (rc) return rc;
        const int64_t n = (int64_t)nb * 3 * h * w;
        k_unnorm_grad<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g_x, (int64_t)h * w, n);
        DFB_LAUNCH_CHECK();
      }
      break;
    }
    const int pv = i - 1;
    if (!kPoolAfter[pv]) {
      int rc = dfb_conv_run(d->enc_dg[i], cur, nb, h, w, 0, bufs[flip], nullptr, nullptr, act_ptr(pv), nullptr, stream);
      if (rc) return rc;
      cur = (const uint16_t*)bufs[flip], flip ^= 1;
    } else {
      // gradient w.r.t. the pooled tensor, then max-pool^T + ReLU' (+ the tap gradient of this conv)
      int rc = dfb_conv_run(d->enc_dg[i], cur, nb, h, w, 0, sc + S.gC, nullptr, nullptr, nullptr, nullptr, stream);
      if (rc) return rc;
      const int ph = L.h[pv], pw = L.w[pv], C = kEncCout[pv];
      const uint16_t* add = pv == 1 ? gtap[0] : (pv == 6 ? gtap[1] : nullptr);
      uint16_t* out = (uint16_t*)bufs[flip];
      if ((ph | pw) & 1) {
        const int64_t n16 = (int64_t)nb * ph * pw * C / 8;
        k_fill16<<<(unsigned)((n16 + 255) / 256), 256, 0, st>>>((uint4*)out, (const uint4*)add, n16);
        DFB_LAUNCH_CHECK();
      }
      const int64_t nw = (int64_t)nb * (ph / 2) * (pw / 2) * (C / 8);
      k_maxpool2x2_bwd<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>((const uint16_t*)act_ptr(pv), (const uint16_t*)(sc + S.gC), add, out, nb,
                                                                    ph, pw, C);
      DFB_LAUNCH_CHECK();
      cur = out, flip ^= 1;
    }
  }
  return DFB_OK;
}

// See DfbDfnet::bucket_event.  event = NULL switches the notification off.
extern "C" int dfb_dfnet_bwd_bucket_event(DfbDfnet* d, int first_layer, void* event) {
  DFB_REQUIRE(d && first_layer >= 0 && first_layer <= 12, DFB_ERR_INVALID, "dfb_dfnet_bwd_bucket_event: bad arguments");
  d->bucket_event = event, d->bucket_first_layer = first_layer;
  return DFB_OK;
}

// adjoint of dfb_resize_bilinear_ac: g_dst [P,Ho,Wo] -> g_src [P,h,w] (overwritten)
extern "C" int dfb_resize_bilinear_ac_bwd(const float* g_dst, int64_t planes, int h, int w, int Ho, int Wo, float* g_src, void* stream) {
  DFB_REQUIRE(g_dst && g_src && planes >= 1 && h >= 1 && w >= 1 && Ho >= 1 && Wo >= 1, DFB_ERR_INVALID, "bad arguments");
  DFB_CHECK_CUDA(cudaMemsetAsync(g_src, 0, (size_t)planes * h * w * 4, (cudaStream_t)stream));
  const int64_t tot = planes * Ho * Wo;
  k_resize_bilinear_ac_bwd<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(g_dst, g_src, (int)planes, h, w, Ho, Wo);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}
