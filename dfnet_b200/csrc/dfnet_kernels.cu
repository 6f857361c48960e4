// HBM-bound pieces of the DFNet feature path and the DFNet forward orchestration.
//
// Reference (paths relative to /root/reference/script):
//   input normalisation            feature/dfnet.py:120-122
//   VGG-16 encoder, pre-ReLU taps  feature/dfnet.py:124-136 (indices 2, 14, 28)
//   adaptation heads               feature/dfnet.py:42-72  (1x1 conv, ReLU, 5x5 conv, BatchNorm2d)
//   bilinear upsample + stack      feature/dfnet.py:145-160 (UpsamplingBilinear2d = align_corners=True)
//   pose head                      feature/dfnet.py:167-170 (AdaptiveAvgPool2d(1), Linear(512,12))
//   feature_loss                   feature/direct_feature_matching.py:114-136 (cosine similarity)
#include <algorithm>
#include <vector>

#include "common.cuh"

struct DfbConv;
extern "C" int dfb_conv_create(int Cin, int Cout, int KH, int KW, const float* weight, const float* bias,
                               const float* bn_scale, const float* bn_shift, DfbConv** out);
extern "C" void dfb_conv_destroy(DfbConv* c);
extern "C" int dfb_conv_fwd(DfbConv* c, const void* in_nhwc16, int B, int H, int W, int relu, void* out_nhwc16,
                            void* tap_nhwc16, float* out_nchw32, void* stream);

namespace dfb {

// x [B,3,H,W] fp32 in [0,1] -> NHWC fp16 [B,H,W,8]: (x - mean) / std, channels 3..7 zero
__global__ void k_input_norm_nhwc8(const float* __restrict__ x, __half* __restrict__ out, int64_t npix, int64_t plane) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const int64_t b = i / plane, p = i % plane;
  const float* s = x + b * 3 * plane + p;
  const float r = __fdiv_rn(__fsub_rn(s[0], 0.485f), 0.229f);
  const float g = __fdiv_rn(__fsub_rn(s[plane], 0.456f), 0.224f);
  const float bl = __fdiv_rn(__fsub_rn(s[2 * plane], 0.406f), 0.225f);
  __half2 h0 = __floats2half2_rn(r, g), h1 = __floats2half2_rn(bl, 0.f);
  uint4 v;
  v.x = *reinterpret_cast<uint32_t*>(&h0), v.y = *reinterpret_cast<uint32_t*>(&h1), v.z = 0u, v.w = 0u;
  reinterpret_cast<uint4*>(out)[i] = v;
}

// 2x2 / stride 2 max pool (floor), NHWC fp16, 8 channels (16 B) per thread
__global__ void k_maxpool2x2_nhwc(const __half* __restrict__ in, __half* __restrict__ out, int B, int H, int W, int C) {
  const int Ho = H / 2, Wo = W / 2, C8 = C / 8;
  const int64_t n = (int64_t)B * Ho * Wo * C8;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c8 = (int)(i % C8);
  const int xo = (int)((i / C8) % Wo), yo = (int)((i / ((int64_t)C8 * Wo)) % Ho), b = (int)(i / ((int64_t)C8 * Wo * Ho));
  const uint4* p = reinterpret_cast<const uint4*>(in + (((int64_t)b * H + 2 * yo) * W + 2 * xo) * C) + c8;
  const int64_t rs = (int64_t)W * C8;
  uint4 a = p[0], bq = p[C8], c = p[rs], d = p[rs + C8];
  auto mx = [](uint32_t u, uint32_t v) {
    __half2 r = __hmax2(*reinterpret_cast<__half2*>(&u), *reinterpret_cast<__half2*>(&v));
    return *reinterpret_cast<uint32_t*>(&r);
  };
  uint4 o;
  o.x = mx(mx(a.x, bq.x), mx(c.x, d.x)), o.y = mx(mx(a.y, bq.y), mx(c.y, d.y));
  o.z = mx(mx(a.z, bq.z), mx(c.z, d.z)), o.w = mx(mx(a.w, bq.w), mx(c.w, d.w));
  reinterpret_cast<uint4*>(out)[i] = o;
}

// bilinear resize, align_corners=True, fp32 NCHW planes: src [P,h,w] -> dst [P,Ho,Wo].
// grid (ceil(Wo/4/128), Ho, P/8): every thread owns 4 outputs along x and loops over 8 planes, so
// the interpolation coordinates are computed once per 32 outputs and every store is 16 bytes
// (HBM-write bound; the sources are L2-resident).  Weights follow ATen: src = scale * dst_index,
// scale = (in-1)/(out-1) in float, lambda = fractional part.
constexpr int kResizePlanes = 8;
__global__ void __launch_bounds__(128) k_resize_bilinear_ac(const float* __restrict__ src, float* __restrict__ dst, int planes,
                                                            int h, int w, int Ho, int Wo) {
  const int xq = (blockIdx.x * blockDim.x + threadIdx.x) * 4, yo = blockIdx.y;
  if (xq >= Wo) return;
  const float sy = Ho > 1 ? (float)(h - 1) / (float)(Ho - 1) : 0.f, sx = Wo > 1 ? (float)(w - 1) / (float)(Wo - 1) : 0.f;
  const float fy = sy * yo;
  const int y0 = (int)fy, y1 = y0 + (y0 < h - 1);
  const float ly = fy - y0;
  int x0[4], x1[4];
  float lx[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int xo = min(xq + k, Wo - 1);
    const float fx = sx * xo;
    x0[k] = (int)fx, x1[k] = x0[k] + (x0[k] < w - 1), lx[k] = fx - x0[k];
  }
  const bool vec = xq + 3 < Wo && (Wo & 3) == 0;
  const int p0 = blockIdx.z * kResizePlanes, p1 = min(planes, p0 + kResizePlanes);
  for (int pl = p0; pl < p1; ++pl) {
    const float* r0 = src + ((int64_t)pl * h + y0) * w;
    const float* r1 = src + ((int64_t)pl * h + y1) * w;
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      o[k] = (1.f - ly) * ((1.f - lx[k]) * __ldg(r0 + x0[k]) + lx[k] * __ldg(r0 + x1[k])) +
             ly * ((1.f - lx[k]) * __ldg(r1 + x0[k]) + lx[k] * __ldg(r1 + x1[k]));
    float* d = dst + ((int64_t)pl * Ho + yo) * Wo + xq;
    if (vec) __stcs(reinterpret_cast<float4*>(d), make_float4(o[0], o[1], o[2], o[3]));
    else
      for (int k = 0; k < 4 && xq + k < Wo; ++k) d[k] = o[k];
  }
}

// adaptive average pool to 1x1 over NHWC fp16 -> fp32 [B,C]; one block per (b, 64 channels)
__global__ void k_avgpool_nhwc(const __half* __restrict__ in, float* __restrict__ out, int HW, int C) {
  const int b = blockIdx.y, c = blockIdx.x * 64 + (threadIdx.x & 63), part = threadIdx.x >> 6;  // 256 threads: 4 pixel lanes
  float s = 0.f;
  for (int p = part; p < HW; p += 4) s += __half2float(in[((int64_t)b * HW + p) * C + c]);
  __shared__ float sm[256];
  sm[threadIdx.x] = s;
  __syncthreads();
  if (part == 0) out[b * C + c] = (sm[threadIdx.x] + sm[threadIdx.x + 64] + sm[threadIdx.x + 128] + sm[threadIdx.x + 192]) / (float)HW;
}

// y[b,o] = bias[o] + sum_k x[b,k] w[o,k]; one warp per output
__global__ void k_fc_small(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                           float* __restrict__ y, int K, int O) {
  const int b = blockIdx.x, o = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (o >= O) return;
  float s = 0.f;
  for (int k = lane; k < K; k += 32) s = fmaf(x[b * K + k], w[o * K + k], s);
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  if (lane == 0) y[b * O + o] = s + bias[o];
}

// ---- cosine feature loss (feature_loss, direct_feature_matching.py:114-136) ---------------------
// rows x cols matrix pair (row-major); cosine along `cols` for every row:
//   per_channel=False -> rows = C, cols = HW (contiguous)      [the reference default]
// partial sums {a.b, a.a, b.b} per (row, split) -> ws[row][split][3]
__global__ void k_cos_rows_partial(const float* __restrict__ fa, const float* __restrict__ fb, int64_t cols, int splits,
                                   float* __restrict__ ws) {
  const int row = blockIdx.y, sp = blockIdx.x;
  const int64_t per = (cols + splits - 1) / splits, c0 = sp * per, c1 = min(cols, c0 + per);
  const float* a = fa + (int64_t)row * cols;
  const float* b = fb + (int64_t)row * cols;
  float ab = 0.f, aa = 0.f, bb = 0.f;
  for (int64_t c = c0 + threadIdx.x; c < c1; c += blockDim.x) {
    const float x = a[c], y = b[c];
    ab = fmaf(x, y, ab), aa = fmaf(x, x, aa), bb = fmaf(y, y, bb);
  }
  __shared__ float sm[3][32];
  for (int d = 16; d > 0; d >>= 1) {
    ab += __shfl_xor_sync(0xffffffffu, ab, d), aa += __shfl_xor_sync(0xffffffffu, aa, d), bb += __shfl_xor_sync(0xffffffffu, bb, d);
  }
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sm[0][wid] = ab, sm[1][wid] = aa, sm[2][wid] = bb;
  __syncthreads();
  if (wid == 0) {
    const int nw = blockDim.x >> 5;
    ab = lane < nw ? sm[0][lane] : 0.f, aa = lane < nw ? sm[1][lane] : 0.f, bb = lane < nw ? sm[2][lane] : 0.f;
    for (int d = 16; d > 0; d >>= 1) {
      ab += __shfl_xor_sync(0xffffffffu, ab, d), aa += __shfl_xor_sync(0xffffffffu, aa, d), bb += __shfl_xor_sync(0xffffffffu, bb, d);
    }
    if (lane == 0) {
      float* o = ws + ((int64_t)row * splits + sp) * 3;
      o[0] = ab, o[1] = aa, o[2] = bb;
    }
  }
}

// loss = 1 - mean_rows( ab / (max(|a|,eps) * max(|b|,eps)) )   (torch >= 1.12 clamps each norm)
__global__ void k_cos_rows_final(const float* __restrict__ ws, int rows, int splits, float eps, float* __restrict__ loss) {
  float acc = 0.f;
  for (int r = threadIdx.x; r < rows; r += blockDim.x) {
    float ab = 0.f, aa = 0.f, bb = 0.f;
    for (int s = 0; s < splits; ++s) {
      const float* o = ws + ((int64_t)r * splits + s) * 3;
      ab += o[0], aa += o[1], bb += o[2];
    }
    acc += ab / (fmaxf(sqrtf(aa), eps) * fmaxf(sqrtf(bb), eps));
  }
  __shared__ float sm[32];
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.f;
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if (threadIdx.x == 0) *loss = 1.f - acc / (float)rows;
  }
}

// per_channel=True: cosine along the channel axis for every pixel (rows = C strided by HW);
// each thread owns one pixel; block partial sums of the cosines -> ws[block]
__global__ void k_cos_cols_partial(const float* __restrict__ fa, const float* __restrict__ fb, int C, int64_t HW, float eps,
                                   float* __restrict__ ws) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float cs = 0.f;
  if (p < HW) {
    float ab = 0.f, aa = 0.f, bb = 0.f;
    for (int c = 0; c < C; ++c) {
      const float x = fa[(int64_t)c * HW + p], y = fb[(int64_t)c * HW + p];
      ab = fmaf(x, y, ab), aa = fmaf(x, x, aa), bb = fmaf(y, y, bb);
    }
    cs = ab / (fmaxf(sqrtf(aa), eps) * fmaxf(sqrtf(bb), eps));
  }
  __shared__ float sm[32];
  for (int d = 16; d > 0; d >>= 1) cs += __shfl_xor_sync(0xffffffffu, cs, d);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = cs;
  __syncthreads();
  if (threadIdx.x < 32) {
    cs = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.f;
    for (int d = 16; d > 0; d >>= 1) cs += __shfl_xor_sync(0xffffffffu, cs, d);
    if (threadIdx.x == 0) ws[blockIdx.x] = cs;
  }
}
__global__ void k_sum_final(const float* __restrict__ ws, int n, float inv_count, float* __restrict__ loss) {
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += ws[i];
  __shared__ float sm[32];
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.f;
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if (threadIdx.x == 0) *loss = 1.f - acc * inv_count;
  }
}

}  // namespace dfb

using namespace dfb;

// ------------------------------------------------------------------------------------------
// DFNet handle
// ------------------------------------------------------------------------------------------
struct DfbDfnet {
  int n_levels = 3;
  DfbConv* enc[13] = {};
  DfbConv* head1[3] = {};
  DfbConv* head5[3] = {};
  float* fc_w = nullptr;
  float* fc_b = nullptr;
  bool loaded = false;
};

static const int kEncCin[13] = {3, 64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512};
static const int kEncCout[13] = {64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512, 512};
static const bool kPoolAfter[13] = {false, true, false, true, false, false, true, false, false, true, false, false, false};
static const int kTapConv[3] = {1, 6, 12};   // conv1_2, conv3_3, conv5_3
static const int kTapCh[3] = {64, 256, 512};

extern "C" int dfb_dfnet_create(int n_levels, DfbDfnet** out) {
  DFB_REQUIRE(out && (n_levels == 1 || n_levels == 3), DFB_ERR_INVALID, "n_levels must be 3 (DFNet) or 1 (DFNet_s)");
  DfbDfnet* d = new DfbDfnet();
  d->n_levels = n_levels;
  *out = d;
  return DFB_OK;
}

extern "C" void dfb_dfnet_destroy(DfbDfnet* d) {
  if (!d) return;
  for (auto c : d->enc) dfb_conv_destroy(c);
  for (auto c : d->head1) dfb_conv_destroy(c);
  for (auto c : d->head5) dfb_conv_destroy(c);
  if (d->fc_w) cudaFree(d->fc_w);
  if (d->fc_b) cudaFree(d->fc_b);
  delete d;
}

// params: 13 x (conv weight, bias), then per level (w1x1, b1x1, w5x5, b5x5, bn_weight, bn_bias,
// bn_running_mean, bn_running_var), then fc_pose weight, bias — fp32, host or device memory.
extern "C" int dfb_dfnet_load(DfbDfnet* d, const float* const* params, const int64_t* numel, int n_params, float bn_eps) {
  DFB_REQUIRE(d && params && numel, DFB_ERR_INVALID, "null argument");
  const int expect = 26 + 8 * d->n_levels + 2;
  DFB_REQUIRE(n_params == expect, DFB_ERR_INVALID, "expected %d tensors, got %d", expect, n_params);
  for (int i = 0; i < 13; ++i) {
    DFB_REQUIRE(numel[2 * i] == (int64_t)kEncCout[i] * kEncCin[i] * 9 && numel[2 * i + 1] == kEncCout[i], DFB_ERR_INVALID,
                "encoder conv %d has the wrong size", i);
    if (d->enc[i]) dfb_conv_destroy(d->enc[i]);
    int rc = dfb_conv_create(kEncCin[i], kEncCout[i], 3, 3, params[2 * i], params[2 * i + 1], nullptr, nullptr, &d->enc[i]);
    if (rc) return rc;
  }
  for (int l = 0; l < d->n_levels; ++l) {
    const float* const* p = params + 26 + 8 * l;
    const int64_t* ne = numel + 26 + 8 * l;
    DFB_REQUIRE(ne[0] == 64 * kTapCh[l] && ne[1] == 64 && ne[2] == 128 * 64 * 25 && ne[3] == 128 && ne[4] == 128 &&
                    ne[5] == 128 && ne[6] == 128 && ne[7] == 128,
                DFB_ERR_INVALID, "adaptation layer %d has the wrong size", l);
    std::vector<float> g(128), be(128), mu(128), var(128), sc(128), sh(128);
    DFB_CHECK_CUDA(cudaMemcpy(g.data(), p[4], 512, cudaMemcpyDefault));
    DFB_CHECK_CUDA(cudaMemcpy(be.data(), p[5], 512, cudaMemcpyDefault));
    DFB_CHECK_CUDA(cudaMemcpy(mu.data(), p[6], 512, cudaMemcpyDefault));
    DFB_CHECK_CUDA(cudaMemcpy(var.data(), p[7], 512, cudaMemcpyDefault));
    for (int c = 0; c < 128; ++c) {  // eval-mode BatchNorm2d: y = (x - mean) / sqrt(var + eps) * gamma + beta
      sc[c] = g[c] / sqrtf(var[c] + bn_eps);
      sh[c] = be[c] - mu[c] * sc[c];
    }
    if (d->head1[l]) dfb_conv_destroy(d->head1[l]);
    if (d->head5[l]) dfb_conv_destroy(d->head5[l]);
    int rc = dfb_conv_create(kTapCh[l], 64, 1, 1, p[0], p[1], nullptr, nullptr, &d->head1[l]);
    if (rc) return rc;
    rc = dfb_conv_create(64, 128, 5, 5, p[2], p[3], sc.data(), sh.data(), &d->head5[l]);
    if (rc) return rc;
  }
  const int o = 26 + 8 * d->n_levels;
  DFB_REQUIRE(numel[o] == 12 * 512 && numel[o + 1] == 12, DFB_ERR_INVALID, "fc_pose has the wrong size");
  if (!d->fc_w) DFB_CHECK_CUDA(cudaMalloc(&d->fc_w, 12 * 512 * 4));
  if (!d->fc_b) DFB_CHECK_CUDA(cudaMalloc(&d->fc_b, 12 * 4));
  DFB_CHECK_CUDA(cudaMemcpy(d->fc_w, params[o], 12 * 512 * 4, cudaMemcpyDefault));
  DFB_CHECK_CUDA(cudaMemcpy(d->fc_b, params[o + 1], 12 * 4, cudaMemcpyDefault));
  d->loaded = true;
  return DFB_OK;
}

static size_t al256(size_t x) { return (x + 255) / 256 * 256; }

struct DfWs { size_t in8, bufA, bufB, tap[3], mid, feat, pooled, total; };

static DfWs dfnet_ws(int B, int H, int W, int n_levels, int upH, int upW) {
  DfWs w;
  size_t off = 0;
  auto take = [&](size_t b) { size_t o = off; off += al256(b); return o; };
  const size_t px = (size_t)B * H * W;
  w.in8 = take(px * 8 * 2);
  w.bufA = take(px * 64 * 2);
  w.bufB = take(px * 64 * 2);
  int h = H, wd = W;
  int lv = 0;
  size_t stage = 256;  // fp32 NCHW staging of the largest level that needs resampling
  for (int i = 0; i < 13; ++i) {
    if (lv < 3 && kTapConv[lv] == i) {
      w.tap[lv] = take((size_t)B * h * wd * kTapCh[lv] * 2);
      if (lv < n_levels && (h != upH || wd != upW)) stage = std::max(stage, (size_t)B * h * wd * 128 * 4);
      ++lv;
    }
    if (kPoolAfter[i]) h /= 2, wd /= 2;
  }
  w.mid = take(px * 64 * 2);
  w.feat = take(stage);
  w.pooled = take((size_t)B * 512 * 4 + 256);
  w.total = off;
  return w;
}

extern "C" int dfb_dfnet_workspace_bytes(const DfbDfnet* d, int B, int H, int W, int upH, int upW, size_t* out) {
  DFB_REQUIRE(d && out && B >= 1 && H >= 32 && W >= 32, DFB_ERR_INVALID, "bad arguments (image must be at least 32x32)");
  *out = dfnet_ws(B, H, W, d->n_levels, upH, upW).total;
  return DFB_OK;
}

// flags: bit0 return_feature, bit1 single_stream, bit2 return_pose.
// feats_t / feats_r: [L, Bs, 128, upH, upW] fp32 with Bs = B (single stream, feats_r unused) or B/2.
extern "C" int dfb_dfnet_fwd(DfbDfnet* d, const float* x, int B, int H, int W, uint32_t flags, int upH, int upW,
                             float* feats_t, float* feats_r, float* pose, void* ws, size_t ws_bytes, void* stream) {
  DFB_REQUIRE(d && d->loaded && x, DFB_ERR_INVALID, "DFNet handle not loaded or null input");
  const bool ret_feat = flags & 1, single = flags & 2, ret_pose = flags & 4;
  DFB_REQUIRE(!ret_feat || feats_t, DFB_ERR_INVALID, "feature output missing");
  DFB_REQUIRE(!ret_feat || single || (feats_r && B % 2 == 0), DFB_ERR_INVALID, "siamese mode needs an even batch and feats_r");
  DFB_REQUIRE(!ret_pose || pose, DFB_ERR_INVALID, "pose output missing");
  DFB_REQUIRE(H >= 32 && W >= 32, DFB_ERR_INVALID, "image must be at least 32x32");
  const DfWs L = dfnet_ws(B, H, W, d->n_levels, upH, upW);
  DFB_REQUIRE(ws && ws_bytes >= L.total, DFB_ERR_WORKSPACE, "workspace too small: need %zu bytes", L.total);
  cudaStream_t st = (cudaStream_t)stream;
  char* base = (char*)ws;
  const int64_t plane = (int64_t)H * W, npix = (int64_t)B * plane;
  k_input_norm_nhwc8<<<(unsigned)((npix + 255) / 256), 256, 0, st>>>(x, (__half*)(base + L.in8), npix, plane);
  DFB_LAUNCH_CHECK();
  const void* cur = base + L.in8;
  void* bufs[2] = {base + L.bufA, base + L.bufB};
  int flip = 0, h = H, w = W, lv = 0;
  int tap_h[3] = {0, 0, 0}, tap_w[3] = {0, 0, 0};
  const int last_conv = d->n_levels == 1 && !ret_pose ? 1 : 12;  // DFNet_s stops after conv1_2 when no pose is needed
  for (int i = 0; i <= last_conv; ++i) {
    void* tap = nullptr;
    if (lv < d->n_levels && kTapConv[lv] == i) { tap = base + L.tap[lv]; tap_h[lv] = h, tap_w[lv] = w; ++lv; }
    const bool need_out = i < last_conv || ret_pose;
    void* o = need_out ? bufs[flip] : nullptr;
    int rc = dfb_conv_fwd(d->enc[i], cur, B, h, w, 1, o, tap, nullptr, stream);
    if (rc) return rc;
    if (!need_out) break;
    cur = o, flip ^= 1;
    if (kPoolAfter[i] || i == 12) {
      const int64_t n = (int64_t)B * (h / 2) * (w / 2) * (kEncCout[i] / 8);
      k_maxpool2x2_nhwc<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const __half*)cur, (__half*)bufs[flip], B, h, w, kEncCout[i]);
      DFB_LAUNCH_CHECK();
      cur = bufs[flip], flip ^= 1, h /= 2, w /= 2;
    }
  }
  if (ret_pose) {  // cur = pool5 output [B,h,w,512]
    float* pooled = (float*)(base + L.pooled);
    k_avgpool_nhwc<<<dim3(512 / 64, B), 256, 0, st>>>((const __half*)cur, pooled, h * w, 512);
    DFB_LAUNCH_CHECK();
    k_fc_small<<<B, 12 * 32, 0, st>>>(pooled, d->fc_w, d->fc_b, pose, 512, 12);
    DFB_LAUNCH_CHECK();
  }
  if (ret_feat) {
    const int Bs = single ? B : B / 2;
    for (int l = 0; l < d->n_levels; ++l) {
      const int fh = tap_h[l], fw = tap_w[l];
      int rc = dfb_conv_fwd(d->head1[l], base + L.tap[l], B, fh, fw, 1, base + L.mid, nullptr, nullptr, stream);
      if (rc) return rc;
      const size_t lvl_stride = (size_t)Bs * 128 * upH * upW;
      const __half* mid = (const __half*)(base + L.mid);
      if (fh == upH && fw == upW) {
        // align_corners resampling to the same size is the identity (level 0 at full resolution):
        // the 5x5 conv writes fp32 NCHW straight into the stacks; siamese split = two half batches
        rc = dfb_conv_fwd(d->head5[l], mid, Bs, fh, fw, 0, nullptr, nullptr, feats_t + l * lvl_stride, stream);
        if (rc) return rc;
        if (!single) {
          rc = dfb_conv_fwd(d->head5[l], mid + (size_t)Bs * fh * fw * 64, Bs, fh, fw, 0, nullptr, nullptr,
                            feats_r + l * lvl_stride, stream);
          if (rc) return rc;
        }
        continue;
      }
      float* featbuf = (float*)(base + L.feat);
      rc = dfb_conv_fwd(d->head5[l], mid, B, fh, fw, 0, nullptr, nullptr, featbuf, stream);
      if (rc) return rc;
      const int planes_s = Bs * 128;
      const dim3 rg((upW + 511) / 512, upH, (planes_s + kResizePlanes - 1) / kResizePlanes);
      k_resize_bilinear_ac<<<rg, 128, 0, st>>>(featbuf, feats_t + l * lvl_stride, planes_s, fh, fw, upH, upW);
      DFB_LAUNCH_CHECK();
      if (!single) {
        k_resize_bilinear_ac<<<rg, 128, 0, st>>>(featbuf + (size_t)planes_s * fh * fw, feats_r + l * lvl_stride, planes_s, fh, fw,
                                                 upH, upW);
        DFB_LAUNCH_CHECK();
      }
    }
  }
  return DFB_OK;
}

// feature_loss (direct_feature_matching.py:114-136): fr, ft fp32 [C, HW] -> *loss (device scalar).
// ws: at least max(C*64*3, ceil(HW/256)) floats.
extern "C" int dfb_cosine_loss(const float* fr, const float* ft, int C, int64_t HW, int per_channel, float eps, float* loss,
                               void* ws, size_t ws_bytes, void* stream) {
  DFB_REQUIRE(fr && ft && loss && ws && C >= 1 && HW >= 1, DFB_ERR_INVALID, "null or empty argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (!per_channel) {
    const int splits = (int)std::min<int64_t>(64, std::max<int64_t>(1, HW / 4096));
    DFB_REQUIRE(ws_bytes >= (size_t)C * splits * 3 * 4, DFB_ERR_WORKSPACE, "workspace too small");
    k_cos_rows_partial<<<dim3(splits, C), 256, 0, st>>>(fr, ft, HW, splits, (float*)ws);
    DFB_LAUNCH_CHECK();
    k_cos_rows_final<<<1, 256, 0, st>>>((const float*)ws, C, splits, eps, loss);
    DFB_LAUNCH_CHECK();
  } else {
    const int blocks = (int)((HW + 255) / 256);
    DFB_REQUIRE(ws_bytes >= (size_t)blocks * 4, DFB_ERR_WORKSPACE, "workspace too small");
    k_cos_cols_partial<<<blocks, 256, 0, st>>>(fr, ft, C, HW, eps, (float*)ws);
    DFB_LAUNCH_CHECK();
    k_sum_final<<<1, 256, 0, st>>>((const float*)ws, blocks, 1.f / (float)HW, loss);
    DFB_LAUNCH_CHECK();
  }
  return DFB_OK;
}

// ------------------------------------------------------------------------------------------
// triplet loss with in-triplet hard negative mining (feature/misc.py:399-435) and MSE
// ------------------------------------------------------------------------------------------
namespace dfb {

__device__ __forceinline__ float block_sum(float v, float* sm) {
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  v = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.f;
  if (threadIdx.x < 32)
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;  // valid in thread 0
}

// f1, f2: [L,B,inner]; negatives are the batch-rolled tensors (roll(shifts=1, dims=1): b -> b-1).
// partial sums of the four squared distances {|f1-roll f2|, |f2-roll f1|, |f1-roll f1|, |f2-roll f2|}
__global__ void k_triplet_mse_partial(const float* __restrict__ f1, const float* __restrict__ f2, int L, int B, int64_t inner,
                                      float* __restrict__ ws) {
  const int64_t n = (int64_t)L * B * inner;
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t in = i % inner, lb = i / inner;
    const int b = (int)(lb % B);
    const int64_t l = lb / B;
    const int64_t j = (l * B + (b + B - 1) % B) * inner + in;  // rolled index
    const float a = f1[i], p = f2[i], an = f1[j], ng = f2[j];
    s[0] = fmaf(a - ng, a - ng, s[0]), s[1] = fmaf(p - an, p - an, s[1]);
    s[2] = fmaf(a - an, a - an, s[2]), s[3] = fmaf(p - ng, p - ng, s[3]);
  }
  __shared__ float sm[32];
  for (int k = 0; k < 4; ++k) {
    const float v = block_sum(s[k], sm);
    if (threadIdx.x == 0) ws[blockIdx.x * 4 + k] = v;
  }
}

// sums[4] (deterministic order) and the chosen case (torch.argmin: first minimum)
__global__ void k_triplet_case(const float* __restrict__ ws, int nblocks, float* __restrict__ sums, int* __restrict__ chosen) {
  __shared__ float sm[32];
  float tot[4];
  for (int k = 0; k < 4; ++k) {
    float v = 0.f;
    for (int i = threadIdx.x; i < nblocks; i += blockDim.x) v += ws[i * 4 + k];
    tot[k] = block_sum(v, sm);
  }
  if (threadIdx.x == 0) {
    int best = 0;
    for (int k = 0; k < 4; ++k) { sums[k] = tot[k]; if (tot[k] < tot[best]) best = k; }
    *chosen = best;
  }
}

// TripletMarginLoss(margin, p=2, eps=1e-6, reduction='mean') on [L,B,C,H,W]: pairwise_distance over W.
// one warp per row (l,b,c,h); block partial sums of the hinge -> ws[block]
__global__ void k_triplet_hinge_partial(const float* __restrict__ f1, const float* __restrict__ f2, int B, int64_t rows_per_b,
                                        int W, int64_t n_rows, float margin, const int* __restrict__ chosen,
                                        float* __restrict__ ws) {
  const int cs = *chosen;
  // case 0: (a,p,n) = (f1, f2, roll f2); 1: (f2, f1, roll f1); 2: (f1, f2, roll f1); 3: (f2, f1, roll f2)
  const float* A = (cs == 0 || cs == 2) ? f1 : f2;
  const float* P = (cs == 0 || cs == 2) ? f2 : f1;
  const float* N = (cs == 0 || cs == 3) ? f2 : f1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  float acc = 0.f;
  for (int64_t row = (int64_t)blockIdx.x * wpb + warp; row < n_rows; row += (int64_t)gridDim.x * wpb) {
    const int64_t lb = row / rows_per_b, rin = row % rows_per_b;
    const int b = (int)(lb % B);
    const int64_t l = lb / B;
    const int64_t rrow = (l * B + (b + B - 1) % B) * rows_per_b + rin;
    const float* a = A + row * W;
    const float* p = P + row * W;
    const float* ng = N + rrow * W;
    float dap = 0.f, dan = 0.f;
    for (int x = lane; x < W; x += 32) {
      const float av = a[x];
      const float u = av - p[x] + 1e-6f, v = av - ng[x] + 1e-6f;
      dap = fmaf(u, u, dap), dan = fmaf(v, v, dan);
    }
    for (int d = 16; d > 0; d >>= 1) dap += __shfl_xor_sync(0xffffffffu, dap, d), dan += __shfl_xor_sync(0xffffffffu, dan, d);
    if (lane == 0) acc += fmaxf(sqrtf(dap) - sqrtf(dan) + margin, 0.f);
  }
  __shared__ float sm[32];
  const float v = block_sum(acc, sm);
  if (threadIdx.x == 0) ws[blockIdx.x] = v;
}

__global__ void k_scaled_sum(const float* __restrict__ ws, int n, float scale, float* __restrict__ out) {
  __shared__ float sm[32];
  float v = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) v += ws[i];
  v = block_sum(v, sm);
  if (threadIdx.x == 0) *out = v * scale;
}

__global__ void k_sqdiff_partial(const float* __restrict__ a, const float* __restrict__ b, int64_t n, float* __restrict__ ws) {
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = a[i] - b[i];
    s = fmaf(d, d, s);
  }
  __shared__ float sm[32];
  s = block_sum(s, sm);
  if (threadIdx.x == 0) ws[blockIdx.x] = s;
}

}  // namespace dfb

// triplet_loss_hard_negative_mining_plus (feature/misc.py:399-435).  f1, f2 fp32 [L,B,C,H,W];
// *loss and *chosen_case are device scalars; ws >= 8192 floats.
extern "C" int dfb_triplet_loss(const float* f1, const float* f2, int L, int B, int Cc, int H, int W, float margin, float* loss,
                                int* chosen_case, void* ws, size_t ws_bytes, void* stream) {
  DFB_REQUIRE(f1 && f2 && loss && chosen_case && ws, DFB_ERR_INVALID, "null argument");
  DFB_REQUIRE(L >= 1 && B >= 1 && Cc >= 1 && H >= 1 && W >= 1, DFB_ERR_INVALID, "empty feature stack");
  DFB_REQUIRE(ws_bytes >= 8192 * 4, DFB_ERR_WORKSPACE, "workspace too small (8192 floats)");
  cudaStream_t st = (cudaStream_t)stream;
  float* w = (float*)ws;
  const int64_t inner = (int64_t)Cc * H * W, n = (int64_t)L * B * inner;
  const int nb = (int)std::min<int64_t>(1024, (n + 255) / 256);
  k_triplet_mse_partial<<<nb, 256, 0, st>>>(f1, f2, L, B, inner, w);
  DFB_LAUNCH_CHECK();
  k_triplet_case<<<1, 256, 0, st>>>(w, nb, w + 4096, chosen_case);
  DFB_LAUNCH_CHECK();
  const int64_t rows_per_b = (int64_t)Cc * H, n_rows = (int64_t)L * B * rows_per_b;
  const int nb2 = (int)std::min<int64_t>(2048, (n_rows + 7) / 8);
  k_triplet_hinge_partial<<<nb2, 256, 0, st>>>(f1, f2, B, rows_per_b, W, n_rows, margin, chosen_case, w + 4104);
  DFB_LAUNCH_CHECK();
  k_scaled_sum<<<1, 256, 0, st>>>(w + 4104, nb2, 1.f / (float)n_rows, loss);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

// mean((a-b)^2): nn.MSELoss / img2mse (models/nerfw.py:11, feature/direct_feature_matching.py:138-142).
extern "C" int dfb_mse(const float* a, const float* b, int64_t n, float* out, void* ws, size_t ws_bytes, void* stream) {
  DFB_REQUIRE(a && b && out && ws && n >= 1, DFB_ERR_INVALID, "null or empty argument");
  DFB_REQUIRE(ws_bytes >= 1024 * 4, DFB_ERR_WORKSPACE, "workspace too small (1024 floats)");
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = (int)std::min<int64_t>(1024, (n + 255) / 256);
  k_sqdiff_partial<<<nb, 256, 0, st>>>(a, b, n, (float*)ws);
  DFB_LAUNCH_CHECK();
  k_scaled_sum<<<1, 256, 0, st>>>((const float*)ws, nb, 1.f / (float)n, out);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

// ------------------------------------------------------------------------------------------
// bicubic resize (torch.nn.Upsample(size, mode='bicubic'), align_corners=False, A = -0.75,
// border-clamped taps, no output clamp) — feature/direct_feature_matching.py:346, feature/misc.py:233
// ------------------------------------------------------------------------------------------
namespace dfb {

__device__ __forceinline__ void cubic_coeffs(float t, float (&c)[4]) {
  const float A = -0.75f;
  auto c1 = [&](float x) { return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; };           // |x| <= 1
  auto c2 = [&](float x) { return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; };    // 1 < |x| < 2
  c[0] = c2(t + 1.f), c[1] = c1(t), c[2] = c1(1.f - t), c[3] = c2(2.f - t);
}

// src [P,h,w] fp32 -> dst [P,Ho,Wo] fp32; one thread per output pixel, grid.z = plane
__global__ void k_resize_bicubic(const float* __restrict__ src, float* __restrict__ dst, int h, int w, int Ho, int Wo) {
  const int xo = blockIdx.x * blockDim.x + threadIdx.x, yo = blockIdx.y;
  if (xo >= Wo) return;
  const int64_t pl = blockIdx.z;
  // area_pixel_compute_source_index(scale, dst, align_corners=False, cubic=True): scale*(dst+0.5)-0.5, not clamped
  const float sy = (float)h / (float)Ho, sx = (float)w / (float)Wo;
  const float fy = sy * (yo + 0.5f) - 0.5f, fx = sx * (xo + 0.5f) - 0.5f;
  const int iy = (int)floorf(fy), ix = (int)floorf(fx);
  float cy[4], cx[4];
  cubic_coeffs(fy - iy, cy);
  cubic_coeffs(fx - ix, cx);
  const float* s = src + pl * h * w;
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int yy = min(max(iy - 1 + i, 0), h - 1);
    float row = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int xx = min(max(ix - 1 + j, 0), w - 1);
      row = fmaf(cx[j], __ldg(s + yy * w + xx), row);
    }
    acc = fmaf(cy[i], row, acc);
  }
  dst[(pl * Ho + yo) * Wo + xo] = acc;
}

}  // namespace dfb

// torch.nn.Upsample(size=(Ho,Wo), mode='bicubic') on fp32 [P,h,w] planes (P = B*C).
extern "C" int dfb_resize_bicubic(const float* src, int64_t planes, int h, int w, int Ho, int Wo, float* dst, void* stream) {
  DFB_REQUIRE(src && dst && planes >= 1 && h >= 1 && w >= 1 && Ho >= 1 && Wo >= 1, DFB_ERR_INVALID, "bad arguments");
  DFB_REQUIRE(planes <= 65535, DFB_ERR_INVALID, "too many planes");
  k_resize_bicubic<<<dim3((Wo + 127) / 128, Ho, (unsigned)planes), 128, 0, (cudaStream_t)stream>>>(src, dst, h, w, Ho, Wo);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

// torch.nn.UpsamplingBilinear2d(size=(Ho,Wo)) (align_corners=True) on fp32 [P,h,w] planes.
extern "C" int dfb_resize_bilinear_ac(const float* src, int64_t planes, int h, int w, int Ho, int Wo, float* dst, void* stream) {
  DFB_REQUIRE(src && dst && planes >= 1 && h >= 1 && w >= 1 && Ho >= 1 && Wo >= 1, DFB_ERR_INVALID, "bad arguments");
  const dim3 rg((Wo + 511) / 512, Ho, (unsigned)((planes + kResizePlanes - 1) / kResizePlanes));
  DFB_REQUIRE(rg.z <= 65535, DFB_ERR_INVALID, "too many planes");
  k_resize_bilinear_ac<<<rg, 128, 0, (cudaStream_t)stream>>>(src, dst, (int)planes, h, w, Ho, Wo);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}
