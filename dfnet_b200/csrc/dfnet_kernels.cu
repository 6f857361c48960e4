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

#include <stdlib.h>

#include "dfnet_handle.cuh"
#include "tc_common.cuh"

namespace dfb {

// x [B,3,H,W] fp32 in [0,1] -> NHWC fp16 [B,H,W,8]: (x - mean) / std, channels 3..7 zero
template <typename T>
__global__ void k_input_norm_nhwc8(const float* __restrict__ x, T* __restrict__ out, int64_t npix, int64_t plane) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const int64_t b = i / plane, p = i % plane;
  const float* s = x + b * 3 * plane + p;
  const float r = __fdiv_rn(__fsub_rn(s[0], 0.485f), 0.229f);
  const float g = __fdiv_rn(__fsub_rn(s[plane], 0.456f), 0.224f);
  const float bl = __fdiv_rn(__fsub_rn(s[2 * plane], 0.406f), 0.225f);
  uint4 v;
  v.x = tc::pack2<T>(r, g), v.y = tc::pack2<T>(bl, 0.f), v.z = 0u, v.w = 0u;
  reinterpret_cast<uint4*>(out)[i] = v;
}

// 2x2 / stride 2 max pool (floor), NHWC fp16, 8 channels (16 B) per thread
template <typename T>
__global__ void k_maxpool2x2_nhwc(const T* __restrict__ in, T* __restrict__ out, int B, int H, int W, int C) {
  const int Ho = H / 2, Wo = W / 2, C8 = C / 8;
  const int64_t n = (int64_t)B * Ho * Wo * C8;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c8 = (int)(i % C8);
  const int xo = (int)((i / C8) % Wo), yo = (int)((i / ((int64_t)C8 * Wo)) % Ho), b = (int)(i / ((int64_t)C8 * Wo * Ho));
  const uint4* p = reinterpret_cast<const uint4*>(in + (((int64_t)b * H + 2 * yo) * W + 2 * xo) * C) + c8;
  const int64_t rs = (int64_t)W * C8;
  uint4 a = p[0], bq = p[C8], c = p[rs], d = p[rs + C8];
  // post-ReLU activations are >= 0 (or -0 / NaN-free), so the 15-bit magnitude patterns order like the values
  auto mx = [](uint32_t u, uint32_t v) {
    const uint32_t ul = u & 0xffffu, vl = v & 0xffffu, uh = u >> 16, vh = v >> 16;
    auto m1 = [](uint32_t p, uint32_t q) {  // max of two 16-bit floats by value (sign-magnitude compare)
      const int32_t ps = (p & 0x8000u) ? -(int32_t)(p & 0x7fffu) : (int32_t)p;
      const int32_t qs = (q & 0x8000u) ? -(int32_t)(q & 0x7fffu) : (int32_t)q;
      return qs > ps ? q : p;
    };
    return m1(ul, vl) | (m1(uh, vh) << 16);
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

// Upsampling variant: one block = one plane x kRsRows output rows.  The <= kRsRows*sy + 2 source rows the block needs
// are staged in shared memory once (the first version gathered 16 scalars from L2 per 16-byte store and ran at 2.6
// TB/s); afterwards every thread produces float4 stores from shared memory, so the kernel is bound by its HBM writes.
// Same arithmetic as k_resize_bilinear_ac (bit-identical results).
constexpr int kRsRows = 16;
__global__ void __launch_bounds__(256) k_resize_bilinear_ac_up(const float* __restrict__ src, float* __restrict__ dst, int h, int w,
                                                               int Ho, int Wo, int nrows_max) {
  extern __shared__ float srows[];  // [nrows][w]
  const int pl = blockIdx.y, yo0 = blockIdx.x * kRsRows, yo1 = min(Ho, yo0 + kRsRows);
  const float sy = Ho > 1 ? (float)(h - 1) / (float)(Ho - 1) : 0.f, sx = Wo > 1 ? (float)(w - 1) / (float)(Wo - 1) : 0.f;
  const int ys0 = (int)(sy * yo0);
  const int ylast = (int)(sy * (yo1 - 1));
  const int ys1 = min(h - 1, ylast + 1);
  const int nrows = ys1 - ys0 + 1;
  const float* sp = src + ((int64_t)pl * h + ys0) * w;
  for (int i = threadIdx.x; i < nrows * w; i += 256) srows[i] = __ldg(sp + i);
  __syncthreads();
  const int nq = (Wo + 3) >> 2;
  for (int q = threadIdx.x; q < nq; q += 256) {
    const int xq = q * 4;
    int x0[4], x1[4];
    float lx[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int xo = min(xq + k, Wo - 1);
      const float fx = sx * xo;
      x0[k] = (int)fx, x1[k] = x0[k] + (x0[k] < w - 1), lx[k] = fx - x0[k];
    }
    const bool vec = xq + 3 < Wo && (Wo & 3) == 0;
    for (int yo = yo0; yo < yo1; ++yo) {
      const float fy = sy * yo;
      const int y0 = (int)fy, y1 = y0 + (y0 < h - 1);
      const float ly = fy - y0;
      const float* r0 = srows + (y0 - ys0) * w;
      const float* r1 = srows + (y1 - ys0) * w;
      float o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        o[k] = (1.f - ly) * ((1.f - lx[k]) * r0[x0[k]] + lx[k] * r0[x1[k]]) + ly * ((1.f - lx[k]) * r1[x0[k]] + lx[k] * r1[x1[k]]);
      float* d = dst + ((int64_t)pl * Ho + yo) * Wo + xq;
      if (vec) __stcs(reinterpret_cast<float4*>(d), make_float4(o[0], o[1], o[2], o[3]));
      else
        for (int k = 0; k < 4 && xq + k < Wo; ++k) d[k] = o[k];
    }
  }
}

// picks the staged kernel for upsampling (source rows of a block fit shared memory), the gather kernel otherwise
static int launch_resize_bilinear_ac(const float* src, float* dst, int64_t planes, int h, int w, int Ho, int Wo, cudaStream_t st) {
  const double sy = Ho > 1 ? (double)(h - 1) / (double)(Ho - 1) : 0.0;
  const int nrows_max = (int)(sy * (kRsRows - 1)) + 3;
  const size_t smem = (size_t)nrows_max * w * sizeof(float);
  if (Ho >= h && Wo >= w && smem <= 48 * 1024 && planes <= 65535) {
    const dim3 g((Ho + kRsRows - 1) / kRsRows, (unsigned)planes);
    k_resize_bilinear_ac_up<<<g, 256, smem, st>>>(src, dst, h, w, Ho, Wo, nrows_max);
    DFB_LAUNCH_CHECK();
    return DFB_OK;
  }
  const dim3 rg((Wo + 511) / 512, Ho, (unsigned)((planes + kResizePlanes - 1) / kResizePlanes));
  DFB_REQUIRE(rg.z <= 65535, DFB_ERR_INVALID, "too many planes");
  k_resize_bilinear_ac<<<rg, 128, 0, st>>>(src, dst, (int)planes, h, w, Ho, Wo);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

// adaptive average pool to 1x1 over NHWC fp16 -> fp32 [B,C]; one block per (b, 64 channels)
template <typename T>
__global__ void k_avgpool_nhwc(const T* __restrict__ in, float* __restrict__ out, int HW, int C) {
  const int b = blockIdx.y, c = blockIdx.x * 64 + (threadIdx.x & 63), part = threadIdx.x >> 6;  // 256 threads: 4 pixel lanes
  float s = 0.f;
  for (int p = part; p < HW; p += 4) s += (float)in[((int64_t)b * HW + p) * C + c];
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
// One warp per row: the lanes share the row's `splits` partial sums (with one thread per row and a serial loop over the
// splits this single-block launch took 17 us of the 78 us loss).  blockDim = 1024.
__global__ void __launch_bounds__(1024) k_cos_rows_final(const float* __restrict__ ws, int rows, int splits, float eps,
                                                         float* __restrict__ loss) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float acc = 0.f;
  for (int r = warp; r < rows; r += nw) {
    float ab = 0.f, aa = 0.f, bb = 0.f;
    for (int s = lane; s < splits; s += 32) {
      const float* o = ws + ((int64_t)r * splits + s) * 3;
      ab += o[0], aa += o[1], bb += o[2];
    }
    for (int d = 16; d > 0; d >>= 1) {
      ab += __shfl_xor_sync(0xffffffffu, ab, d);
      aa += __shfl_xor_sync(0xffffffffu, aa, d);
      bb += __shfl_xor_sync(0xffffffffu, bb, d);
    }
    acc += ab / (fmaxf(sqrtf(aa), eps) * fmaxf(sqrtf(bb), eps));   // identical in every lane
  }
  __shared__ float sm[32];
  if (lane == 0) sm[warp] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = threadIdx.x < nw ? sm[threadIdx.x] : 0.f;
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
  for (auto c : d->enc_n128) dfb_conv_destroy(c);
  for (auto c : d->enc_bf) dfb_conv_destroy(c);
  for (auto c : d->enc_dg) dfb_conv_destroy(c);
  for (auto c : d->head1) dfb_conv_destroy(c);
  for (auto c : d->head5) dfb_conv_destroy(c);
  for (auto c : d->head1_dg) dfb_conv_destroy(c);
  for (auto c : d->head5_dg) dfb_conv_destroy(c);
  for (auto c : d->head5_raw) dfb_conv_destroy(c);
  for (auto c : d->head5_raw_dg) dfb_conv_destroy(c);
  if (d->bn_gb) cudaFree(d->bn_gb);
  if (d->bn_stat) cudaFree(d->bn_stat);
  if (d->bn_part) cudaFree(d->bn_part);
  if (d->bn_sc) cudaFree(d->bn_sc);
  if (d->bn_sh) cudaFree(d->bn_sh);
  if (d->bn_stage) cudaFree(d->bn_stage);
  for (int i = 0; i < 2; ++i)
    if (d->side[i]) cudaStreamDestroy(d->side[i]);
  for (int i = 0; i < 5; ++i)
    if (d->ev[i]) cudaEventDestroy(d->ev[i]);
  if (d->fc_w) cudaFree(d->fc_w);
  if (d->fc_b) cudaFree(d->fc_b);
  delete d;
}

namespace dfb {
// eval-mode BatchNorm2d: y = (x - mean) / sqrt(var + eps) * gamma + beta  ->  scale, shift
__global__ void k_bn_fold(const float* g, const float* be, const float* mu, const float* var, float eps, float* sc, float* sh) {
  const int c = threadIdx.x;
  const float s = g[c] / sqrtf(var[c] + eps);
  sc[c] = s, sh[c] = be[c] - mu[c] * s;
}

// ---- train-mode BatchNorm2d of the adaptation heads (feature/dfnet.py:57-62 under model.train(), run_feature.py:133) ----
constexpr int kBnSplits = 64;
// x: [B,128,plane] fp32 (pre-BatchNorm conv output).  part[c][split] = {sum, sum of squares} in float64.
__global__ void __launch_bounds__(256) k_bn_partial(const float* __restrict__ x, int B, int64_t plane, double* __restrict__ part) {
  __shared__ double sh[2][8];
  const int c = blockIdx.y, sp = blockIdx.x;
  const int64_t n = (int64_t)B * plane;
  double s = 0.0, q = 0.0;
  for (int64_t i = (int64_t)sp * 256 + threadIdx.x; i < n; i += (int64_t)kBnSplits * 256) {
    const int64_t b = i / plane, p = i - b * plane;
    const double v = (double)x[(b * 128 + c) * plane + p];
    s += v, q += v * v;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o), q += __shfl_xor_sync(0xffffffffu, q, o);
  if ((threadIdx.x & 31) == 0) sh[0][threadIdx.x >> 5] = s, sh[1][threadIdx.x >> 5] = q;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b2 = 0.0;
    for (int w = 0; w < 8; ++w) a += sh[0][w], b2 += sh[1][w];
    part[((size_t)c * kBnSplits + sp) * 2] = a, part[((size_t)c * kBnSplits + sp) * 2 + 1] = b2;
  }
}
// stat[0] mean, stat[1] biased variance (what normalises the batch), stat[2] scale, stat[3] shift; one thread per channel
__global__ void k_bn_finalize(const double* __restrict__ part, double n, const float* __restrict__ gamma,
                              const float* __restrict__ beta, float eps, float* __restrict__ stat) {
  const int c = threadIdx.x;
  double s = 0.0, q = 0.0;
  for (int i = 0; i < kBnSplits; ++i) s += part[((size_t)c * kBnSplits + i) * 2], q += part[((size_t)c * kBnSplits + i) * 2 + 1];
  const double mean = s / n, var = fmax(q / n - mean * mean, 0.0);
  const float sc = gamma[c] / sqrtf((float)var + eps);
  stat[c] = (float)mean, stat[128 + c] = (float)var, stat[256 + c] = sc, stat[384 + c] = beta[c] - (float)mean * sc;
}
// eval-mode BatchNorm as "statistics": stat = {running_mean, running_var, scale, shift} (bn = gamma, beta, mean, var)
__global__ void k_bn_stat_from_running(const float* __restrict__ bn, float eps, float* __restrict__ stat) {
  const int c = threadIdx.x;
  const float sc = bn[c] / sqrtf(bn[384 + c] + eps);
  stat[c] = bn[256 + c], stat[128 + c] = bn[384 + c], stat[256 + c] = sc, stat[384 + c] = bn[128 + c] - bn[256 + c] * sc;
}
// y = x * scale[c] + shift[c]; images [0,Bs) go to out_t, [Bs,B) to out_r (siamese split; out_r unused when Bs == B)
__global__ void __launch_bounds__(256) k_bn_apply(const float* x, int B, int Bs, int64_t plane,
                                                  const float* __restrict__ stat, float* out_t, float* out_r) {  // x may alias out_t
  const int bc = blockIdx.y, b = bc / 128, c = bc % 128;
  const float sc = stat[256 + c], shf = stat[384 + c];
  const float* src = x + (size_t)bc * plane;
  float* dst = (b < Bs ? out_t + ((size_t)b * 128 + c) * plane : out_r + ((size_t)(b - Bs) * 128 + c) * plane);
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < plane; i += (int64_t)gridDim.x * 256) dst[i] = fmaf(src[i], sc, shf);
}
}  // namespace dfb

// create on first use, repack in place afterwards (every optimizer step re-loads the parameters)
static int conv_set(DfbConv** slot, int Cin, int Cout, int K, const float* w, const float* b, const float* sc, const float* sh,
                    int fmt, int dgrad, int nt_force = 0) {
  if (*slot) return dfb_conv_update_impl(*slot, w, b, sc, sh, nullptr);
  return dfb_conv_create_impl(Cin, Cout, K, K, w, b, sc, sh, fmt, dgrad, slot, nt_force);
}

// params: 13 x (conv weight, bias), then per level (w1x1, b1x1, w5x5, b5x5, bn_weight, bn_bias,
// bn_running_mean, bn_running_var), then fc_pose weight, bias — fp32, host or device memory.
// flags bit0: also (re)build the training variants (bf16 forward and data-gradient convolutions).
static int dfnet_load_impl(DfbDfnet* d, const float* const* params, const int64_t* numel, int n_params, float bn_eps, uint32_t flags);

extern "C" int dfb_dfnet_load_ex(DfbDfnet* d, const float* const* params, const int64_t* numel, int n_params, float bn_eps,
                                 uint32_t flags) {
  // every (re)packing request of the call goes out as one launch (conv_tc.cu, k_pack_conv_weights); the flush precedes
  // the copies / synchronisation at the end of dfnet_load_impl only in stream order, which is all they need
  dfb_conv_pack_batch_begin();
  const int rc = dfnet_load_impl(d, params, numel, n_params, bn_eps, flags);
  const int rc2 = dfb_conv_pack_batch_flush(rc != 0);
  if (!rc && !rc2 && !(flags & 2)) DFB_CHECK_CUDA(cudaStreamSynchronize(nullptr));
  return rc ? rc : rc2;
}

static int dfnet_load_impl(DfbDfnet* d, const float* const* params, const int64_t* numel, int n_params, float bn_eps, uint32_t flags) {
  DFB_REQUIRE(d && params && numel, DFB_ERR_INVALID, "null argument");
  const int expect = 26 + 8 * d->n_levels + 2;
  DFB_REQUIRE(n_params == expect, DFB_ERR_INVALID, "expected %d tensors, got %d", expect, n_params);
  const bool train = flags & 1;
  for (int i = 0; i < 13; ++i) {
    DFB_REQUIRE(numel[2 * i] == (int64_t)kEncCout[i] * kEncCin[i] * 9 && numel[2 * i + 1] == kEncCout[i], DFB_ERR_INVALID,
                "encoder conv %d has the wrong size", i);
    int rc = conv_set(&d->enc[i], kEncCin[i], kEncCout[i], 3, params[2 * i], params[2 * i + 1], nullptr, nullptr, 0, 0);
    if (rc) return rc;
    if (kEncCout[i] % 256 == 0) {  // second packing with 128-wide output-channel tiles (see the tile choice in dfnet_fwd_impl)
      rc = conv_set(&d->enc_n128[i], kEncCin[i], kEncCout[i], 3, params[2 * i], params[2 * i + 1], nullptr, nullptr, 0, 0, 128);
      if (rc) return rc;
    }
    if (train && !(flags & 32)) {   // bit5: the bf16 encoder forward is not going to run with these weights
      rc = conv_set(&d->enc_bf[i], kEncCin[i], kEncCout[i], 3, params[2 * i], params[2 * i + 1], nullptr, nullptr, 1, 0);
      if (rc) return rc;
    }
    if (train) {
      rc = conv_set(&d->enc_dg[i], kEncCin[i], kEncCout[i], 3, params[2 * i], nullptr, nullptr, nullptr, 1, 1);
      if (rc) return rc;
    }
  }
  if (!d->bn_sc) DFB_CHECK_CUDA(cudaMalloc(&d->bn_sc, 3 * 128 * 4));
  if (!d->bn_sh) DFB_CHECK_CUDA(cudaMalloc(&d->bn_sh, 3 * 128 * 4));
  // bit4: the adaptation heads are not going to be evaluated with these weights (a pose regressor re-loaded after every
  // optimizer step): their images are left as they are
  for (int l = 0; l < ((flags & 16) ? 0 : d->n_levels); ++l) {
    const float* const* p = params + 26 + 8 * l;
    const int64_t* ne = numel + 26 + 8 * l;
    DFB_REQUIRE(ne[0] == 64 * kTapCh[l] && ne[1] == 64 && ne[2] == 128 * 64 * 25 && ne[3] == 128 && ne[4] == 128 &&
                    ne[5] == 128 && ne[6] == 128 && ne[7] == 128,
                DFB_ERR_INVALID, "adaptation layer %d has the wrong size", l);
    // the four BatchNorm vectors may live on the host: stage them (persistent staging buffer, see dfb_conv_update_impl)
    if (!d->bn_stage) DFB_CHECK_CUDA(cudaMalloc(&d->bn_stage, 3 * 4 * 128 * 4));
    float* bn = d->bn_stage + (size_t)l * 4 * 128;
    for (int k = 0; k < 4; ++k) DFB_CHECK_CUDA(cudaMemcpyAsync(bn + 128 * k, p[4 + k], 512, cudaMemcpyDefault, nullptr));
    float* sc = d->bn_sc + 128 * l, *sh = d->bn_sh + 128 * l;
    k_bn_fold<<<1, 128>>>(bn, bn + 128, bn + 256, bn + 384, bn_eps, sc, sh);
    DFB_LAUNCH_CHECK();
    int rc = conv_set(&d->head1[l], kTapCh[l], 64, 1, p[0], p[1], nullptr, nullptr, 0, 0);
    if (rc) return rc;
    rc = conv_set(&d->head5[l], 64, 128, 5, p[2], p[3], sc, sh, 0, 0);
    if (rc) return rc;
    if (flags & 4) {  // train-mode BatchNorm: the 5x5 conv without the fold, gamma / beta on the device
      rc = conv_set(&d->head5_raw[l], 64, 128, 5, p[2], p[3], nullptr, nullptr, 0, 0);
      if (rc) return rc;
      if (!d->bn_gb) DFB_CHECK_CUDA(cudaMalloc(&d->bn_gb, 3 * 2 * 128 * 4));
      if (!d->bn_stat) DFB_CHECK_CUDA(cudaMalloc(&d->bn_stat, 3 * 4 * 128 * 4));
      if (!d->bn_part) DFB_CHECK_CUDA(cudaMalloc(&d->bn_part, (size_t)3 * 128 * dfb::kBnSplits * 2 * sizeof(double)));
      DFB_CHECK_CUDA(cudaMemcpyAsync(d->bn_gb + (l * 2 + 0) * 128, p[4], 512, cudaMemcpyDefault, nullptr));
      DFB_CHECK_CUDA(cudaMemcpyAsync(d->bn_gb + (l * 2 + 1) * 128, p[5], 512, cudaMemcpyDefault, nullptr));
      d->bn_eps = bn_eps;
    }
    if (flags & 8) {  // training of the heads themselves: data gradient of the un-folded 5x5 conv
      rc = conv_set(&d->head5_raw_dg[l], 64, 128, 5, p[2], nullptr, nullptr, nullptr, 1, 1);
      if (rc) return rc;
    }
    if (train) {
      rc = conv_set(&d->head1_dg[l], kTapCh[l], 64, 1, p[0], nullptr, nullptr, nullptr, 1, 1);
      if (rc) return rc;
      rc = conv_set(&d->head5_dg[l], 64, 128, 5, p[2], nullptr, sc, nullptr, 1, 1);
      if (rc) return rc;
    }
  }
  const int o = 26 + 8 * d->n_levels;
  DFB_REQUIRE(numel[o] == 12 * 512 && numel[o + 1] == 12, DFB_ERR_INVALID, "fc_pose has the wrong size");
  if (!d->fc_w) DFB_CHECK_CUDA(cudaMalloc(&d->fc_w, 12 * 512 * 4));
  if (!d->fc_b) DFB_CHECK_CUDA(cudaMalloc(&d->fc_b, 12 * 4));
  // flags bit 1: the caller's work is ordered on the legacy default stream (the stream every packing kernel above
  // ran on) and the sources stay alive in stream order: no host synchronisation, so a training loop that reloads
  // the pose regressor every step does not drain the GPU here
  if (flags & 2) {
    DFB_CHECK_CUDA(cudaMemcpyAsync(d->fc_w, params[o], 12 * 512 * 4, cudaMemcpyDefault, nullptr));
    DFB_CHECK_CUDA(cudaMemcpyAsync(d->fc_b, params[o + 1], 12 * 4, cudaMemcpyDefault, nullptr));
  } else {
    DFB_CHECK_CUDA(cudaMemcpy(d->fc_w, params[o], 12 * 512 * 4, cudaMemcpyDefault));
    DFB_CHECK_CUDA(cudaMemcpy(d->fc_b, params[o + 1], 12 * 4, cudaMemcpyDefault));
    DFB_CHECK_CUDA(cudaStreamSynchronize(nullptr));
  }
  d->loaded = true;
  return DFB_OK;
}

// Batch statistics of the last train-mode forward (flags bit5): out [n_levels][2][128] = mean, biased variance per head
// channel (device or host pointer; copied on `stream`).  The caller applies the running-statistics update
// (momentum, unbiased variance) to its own buffers, as torch.nn.BatchNorm2d does.
extern "C" int dfb_dfnet_bn_batch_stats(const DfbDfnet* d, float* out, void* stream) {
  DFB_REQUIRE(d && out && d->bn_stat, DFB_ERR_INVALID, "no train-mode BatchNorm forward has run");
  for (int l = 0; l < d->n_levels; ++l)
    DFB_CHECK_CUDA(cudaMemcpyAsync(out + (size_t)l * 256, d->bn_stat + (size_t)l * 512, 256 * sizeof(float), cudaMemcpyDefault,
                                   (cudaStream_t)stream));
  return DFB_OK;
}

extern "C" int dfb_dfnet_load(DfbDfnet* d, const float* const* params, const int64_t* numel, int n_params, float bn_eps) {
  return dfb_dfnet_load_ex(d, params, numel, n_params, bn_eps, 0);
}

// Workspace layout.  tape = false: two ping-pong activation buffers (inference).  tape = true: every
// activation has its own buffer, which is what dfb_dfnet_bwd reads back (training).
DfWs dfnet_ws(int B, int H, int W, int n_levels, int upH, int upW, bool tape) {
  DfWs w = {};
  size_t off = 0;
  auto al256 = [](size_t x) { return (x + 255) / 256 * 256; };
  auto take = [&](size_t b) { size_t o = off; off += al256(b); return o; };
  const size_t px = (size_t)B * H * W;
  w.in8 = take(px * 8 * 2);
  const size_t bufA = tape ? 0 : take(px * 64 * 2), bufB = tape ? 0 : take(px * 64 * 2);
  int h = H, wd = W, lv = 0, flip = 0;
  size_t stage = 256;  // fp32 NCHW staging of the largest level that needs resampling
  for (int i = 0; i < 13; ++i) {
    w.h[i] = h, w.w[i] = wd;
    w.act[i] = tape ? take((size_t)B * h * wd * kEncCout[i] * 2) : (flip ? bufB : bufA);
    flip ^= 1;
    if (lv < 3 && kTapConv[lv] == i) {
      w.tap[lv] = take((size_t)B * h * wd * kTapCh[lv] * 2);
      w.mid[lv] = (tape || lv == 0) ? take((size_t)B * h * wd * 64 * 2) : w.mid[0];
      // fp32 NCHW staging: levels that need resampling, and every level under train-mode BatchNorm (the batch
      // statistics need the whole pre-BatchNorm output before anything can be written to the stacks)
      if (lv < n_levels) stage = std::max(stage, (size_t)B * h * wd * 128 * 4);
      if (tape && lv < n_levels) w.zbn[lv] = take((size_t)B * h * wd * 128 * 4);
      ++lv;
    }
    if (kPoolAfter[i]) {
      h /= 2, wd /= 2;
      w.pool[i] = tape ? take((size_t)B * h * wd * kEncCout[i] * 2) : (flip ? bufB : bufA);
      flip ^= 1;
    }
  }
  w.feat = take(stage);
  w.pooled = take((size_t)B * 512 * 4 + 256);
  w.total = off;
  return w;
}

extern "C" int dfb_dfnet_workspace_bytes(const DfbDfnet* d, int B, int H, int W, int upH, int upW, size_t* out) {
  DFB_REQUIRE(d && out && B >= 1 && H >= 32 && W >= 32, DFB_ERR_INVALID, "bad arguments (image must be at least 32x32)");
  *out = dfnet_ws(B, H, W, d->n_levels, upH, upW, false).total;
  return DFB_OK;
}

// Debug seam: byte offsets inside the tape, so that tests can read the stored activations back.
// out[0] = in8; then for every encoder conv i: out[1+4i..] = {act offset, pool offset or -1, h, w}; then tap[3], mid[3], pooled.
extern "C" int dfb_debug_dfnet_tape_layout(const DfbDfnet* d, int B, int H, int W, int upH, int upW, int64_t* out) {
  DFB_REQUIRE(d && out && B >= 1 && H >= 32 && W >= 32, DFB_ERR_INVALID, "bad arguments");
  const DfWs L = dfnet_ws(B, H, W, d->n_levels, upH, upW, true);
  out[0] = (int64_t)L.in8;
  for (int i = 0; i < 13; ++i) {
    out[1 + 4 * i] = (int64_t)L.act[i], out[2 + 4 * i] = kPoolAfter[i] ? (int64_t)L.pool[i] : -1;
    out[3 + 4 * i] = L.h[i], out[4 + 4 * i] = L.w[i];
  }
  for (int l = 0; l < 3; ++l) out[53 + l] = (int64_t)L.tap[l], out[56 + l] = (int64_t)L.mid[l];
  out[59] = (int64_t)L.pooled;
  return DFB_OK;
}

extern "C" int dfb_dfnet_tape_bytes(const DfbDfnet* d, int B, int H, int W, int upH, int upW, size_t* out) {
  DFB_REQUIRE(d && out && B >= 1 && H >= 32 && W >= 32, DFB_ERR_INVALID, "bad arguments (image must be at least 32x32)");
  *out = dfnet_ws(B, H, W, d->n_levels, upH, upW, true).total;
  return DFB_OK;
}

// flags: bit0 return_feature, bit1 single_stream, bit2 return_pose, bit3 keep the tape (ws = tape buffer of
// dfb_dfnet_tape_bytes, read back by dfb_dfnet_bwd), bit4 bf16 operands in the encoder (pose-only training).
// feats_t / feats_r: [L, Bs, 128, upH, upW] fp32 with Bs = B (single stream, feats_r unused) or B/2.
template <typename T>
static int dfnet_fwd_impl(DfbDfnet* d, const float* x, int B, int H, int W, uint32_t flags, int upH, int upW, float* feats_t,
                          float* feats_r, float* pose, void* ws, size_t ws_bytes, void* stream) {
  const bool ret_feat = flags & 1, single = flags & 2, ret_pose = flags & 4, tape = flags & 8;
  const bool bf = std::is_same<T, __nv_bfloat16>::value;
  const DfWs L = dfnet_ws(B, H, W, d->n_levels, upH, upW, tape);
  DFB_REQUIRE(ws && ws_bytes >= L.total, DFB_ERR_WORKSPACE, "workspace too small: need %zu bytes", L.total);
  cudaStream_t st = (cudaStream_t)stream;
  char* base = (char*)ws;
  const int64_t plane = (int64_t)H * W, npix = (int64_t)B * plane;
  // (a direct fp32 kernel for conv1_1 fused with this normalisation was measured at the tensor-core kernel's speed once
  // that kernel loaded its patches by TMA - 110 vs 107 us under ncu - and was dropped)
  k_input_norm_nhwc8<T><<<(unsigned)((npix + 255) / 256), 256, 0, st>>>(x, (T*)(base + L.in8), npix, plane);
  DFB_LAUNCH_CHECK();
  // One adaptation head (1x1 conv + ReLU, 5x5 conv + BatchNorm, resampling into the stacks) on stream `st`.
  const int Bs = single ? B : B / 2;
  // flags bits 8..10: level l is not wanted (its head is not evaluated, its slice of the stacks is left untouched)
  const uint32_t skip = (flags >> 8) & 7u;
  auto run_head = [&](int l, cudaStream_t st) -> int {
    if (skip & (1u << l)) return DFB_OK;
    void* stream = (void*)st;
      const int fh = L.h[kTapConv[l]], fw = L.w[kTapConv[l]];
      int rc = dfb_conv_fwd(d->head1[l], base + L.tap[l], B, fh, fw, 1, base + L.mid[l], nullptr, nullptr, stream);
      if (rc) return rc;
      const size_t lvl_stride = (size_t)Bs * 128 * upH * upW;
      const __half* mid = (const __half*)(base + L.mid[l]);
      if (flags & (32 | 64)) {
        // bit5, train-mode BatchNorm: conv without the fold over the WHOLE batch -> batch statistics -> normalise.
        // bit6, heads being trained (tape): the same un-folded sequence with the pre-BatchNorm output z kept on the tape
        // (the BatchNorm backward needs it); in eval mode the "statistics" are the running ones.
        DFB_REQUIRE(d->head5_raw[l] && d->bn_stat, DFB_ERR_INVALID, "train-mode BatchNorm variants not loaded (dfb_dfnet_load_ex flags bit2)");
        const bool keep_z = (flags & 64) != 0;
        DFB_REQUIRE(!keep_z || tape, DFB_ERR_INVALID, "flags bit6 (head tape) needs bit3 (tape)");
        float* stagebuf = (float*)(base + L.feat);
        float* featbuf = keep_z ? (float*)(base + L.zbn[l]) : stagebuf;
        rc = dfb_conv_fwd(d->head5_raw[l], mid, B, fh, fw, 0, nullptr, nullptr, featbuf, stream);
        if (rc) return rc;
        const int64_t fplane = (int64_t)fh * fw;
        float* stat = d->bn_stat + l * 4 * 128;
        if (flags & 32) {
          // (one partial-sum buffer for all levels: the levels' heads run on different streams, so each level gets its
          // own third of it)
          double* part = d->bn_part + (size_t)l * 128 * dfb::kBnSplits * 2;
          dfb::k_bn_partial<<<dim3(dfb::kBnSplits, 128), 256, 0, st>>>(featbuf, B, fplane, part);
          DFB_LAUNCH_CHECK();
          dfb::k_bn_finalize<<<1, 128, 0, st>>>(part, (double)B * (double)fplane, d->bn_gb + (l * 2) * 128,
                                                d->bn_gb + (l * 2 + 1) * 128, d->bn_eps, stat);
          DFB_LAUNCH_CHECK();
        } else {
          dfb::k_bn_stat_from_running<<<1, 128, 0, st>>>(d->bn_stage + (size_t)l * 512, d->bn_eps, stat);
          DFB_LAUNCH_CHECK();
        }
        const dim3 ag((unsigned)std::min<int64_t>((fplane + 255) / 256, 64), B * 128);
        if (fh == upH && fw == upW) {
          dfb::k_bn_apply<<<ag, 256, 0, st>>>(featbuf, B, Bs, fplane, stat, feats_t + l * lvl_stride,
                                              single ? nullptr : feats_r + l * lvl_stride);
          DFB_LAUNCH_CHECK();
          return DFB_OK;
        }
        dfb::k_bn_apply<<<ag, 256, 0, st>>>(featbuf, B, B, fplane, stat, stagebuf, nullptr);  // (in place unless z is kept), then resample
        DFB_LAUNCH_CHECK();
        featbuf = stagebuf;
        const int planes_s = Bs * 128;
        rc = launch_resize_bilinear_ac(featbuf, feats_t + l * lvl_stride, planes_s, fh, fw, upH, upW, st);
        if (rc) return rc;
        if (!single) {
          rc = launch_resize_bilinear_ac(featbuf + (size_t)planes_s * fh * fw, feats_r + l * lvl_stride, planes_s, fh, fw, upH, upW, st);
          if (rc) return rc;
        }
        return DFB_OK;
      }
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
        return DFB_OK;
      }
      float* featbuf = (float*)(base + L.feat);
      rc = dfb_conv_fwd(d->head5[l], mid, B, fh, fw, 0, nullptr, nullptr, featbuf, stream);
      if (rc) return rc;
      const int planes_s = Bs * 128;
      rc = launch_resize_bilinear_ac(featbuf, feats_t + l * lvl_stride, planes_s, fh, fw, upH, upW, st);
      if (rc) return rc;
      if (!single) {
        rc = launch_resize_bilinear_ac(featbuf + (size_t)planes_s * fh * fw, feats_r + l * lvl_stride, planes_s, fh, fw, upH, upW, st);
        if (rc) return rc;
      }
    return DFB_OK;
  };
  // The heads only depend on their tap, so they run on side streams NEXT TO the rest of the encoder: every kernel
  // here is a persistent grid of <= 148 CTAs whose tile counts rarely divide by 148 (conv3: 300 tiles, conv4: 160,
  // conv5: 40-80), and the CTAs of a concurrent kernel fill the SMs that finish early.  Level 0 (the 5x5 conv at full
  // resolution, 40 % of the forward's FLOPs) gets its own stream; levels 1 and 2 share one (and the fp32 staging buffer).
  // DFB_DFNET_STREAMS=0 serialises everything on the caller's stream.
  static const bool use_streams = [] { const char* e = getenv("DFB_DFNET_STREAMS"); return !(e && e[0] == '0'); }();
  const bool fork = ret_feat && use_streams;
  if (fork && !d->side[0]) {
    for (int i = 0; i < 2; ++i) DFB_CHECK_CUDA(cudaStreamCreateWithFlags(&d->side[i], cudaStreamNonBlocking));
    for (int i = 0; i < 5; ++i) DFB_CHECK_CUDA(cudaEventCreateWithFlags(&d->ev[i], cudaEventDisableTiming));
  }
  const void* cur = base + L.in8;
  int h = H, w = W, lv = 0;
  // without a pose the encoder stops after the deepest hyper-column that is wanted (DFNet_s: conv1_2; DFNet with only
  // level 0 wanted, e.g. the feature net of train_on_batch with feature_matching_lvl = [0]: conv1_2 as well)
  int last_conv = 12;
  if (!ret_pose && ret_feat) {
    last_conv = 0;
    for (int l = 0; l < d->n_levels; ++l)
      if (!(skip & (1u << l))) last_conv = std::max(last_conv, kTapConv[l]);
  }
  for (int i = 0; i <= last_conv; ++i) {
    void* tap = nullptr;
    if (ret_feat && lv < d->n_levels && kTapConv[lv] == i) { tap = (skip & (1u << lv)) ? nullptr : base + L.tap[lv]; ++lv; }
    const bool need_out = i < last_conv || ret_pose || tape;
    void* o = need_out ? base + L.act[i] : nullptr;
    DfbConv* cv = bf ? d->enc_bf[i] : d->enc[i];
    DFB_REQUIRE(cv, DFB_ERR_INVALID, "training variants not loaded (dfb_dfnet_load_ex flags bit0)");
    // Tile choice by wave count: the persistent grid runs ceil(tiles / SMs) rounds.  A 128-wide tile is half the MMA work
    // of a 256-wide one plus a second load of the input patch (~0.6 of its time, measured); it wins where the 256-wide
    // tiling lands just above a multiple of the SM count (480x640 pair: conv3_x 300 tiles = 3 rounds -> 600 = 5 half
    // rounds; conv4_x 160 = 2 rounds -> 320 = 3 half rounds; conv5_x 40 -> 80 tiles in one round).
    if (!bf && d->enc_n128[i]) {
      const double c256 = (double)dfb_conv_rounds(cv, B, h, w), c128 = 0.62 * (double)dfb_conv_rounds(d->enc_n128[i], B, h, w);
      if (c128 < c256) cv = d->enc_n128[i];
    }
    int rc = dfb_conv_fwd(cv, cur, B, h, w, 1, o, tap, nullptr, stream);
    if (rc) return rc;
    if (tap && fork) {  // level lv-1 is ready: its head starts on a side stream while the encoder continues
      const int l = lv - 1;
      // level 0 has its own stream when it writes straight into the stacks (no resampling: the usual case); a level that
      // is resampled goes through the shared fp32 staging buffer and therefore shares the levels-1/2 stream
      cudaStream_t hs = d->side[(l == 0 && L.h[kTapConv[0]] == upH && L.w[kTapConv[0]] == upW) ? 0 : 1];
      DFB_CHECK_CUDA(cudaEventRecord(d->ev[l], st));
      DFB_CHECK_CUDA(cudaStreamWaitEvent(hs, d->ev[l], 0));
      rc = run_head(l, hs);
      if (rc) return rc;
    }
    if (i == last_conv && !ret_pose) break;
    cur = o;
    if (kPoolAfter[i]) {
      const int64_t n = (int64_t)B * (h / 2) * (w / 2) * (kEncCout[i] / 8);
      k_maxpool2x2_nhwc<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const T*)cur, (T*)(base + L.pool[i]), B, h, w, kEncCout[i]);
      DFB_LAUNCH_CHECK();
      cur = base + L.pool[i], h /= 2, w /= 2;
    }
  }
  if (ret_pose) {  // cur = pool5 output [B,h,w,512]
    float* pooled = (float*)(base + L.pooled);
    k_avgpool_nhwc<T><<<dim3(512 / 64, B), 256, 0, st>>>((const T*)cur, pooled, h * w, 512);
    DFB_LAUNCH_CHECK();
    k_fc_small<<<B, 12 * 32, 0, st>>>(pooled, d->fc_w, d->fc_b, pose, 512, 12);
    DFB_LAUNCH_CHECK();
  }
  if (ret_feat && !fork)
    for (int l = 0; l < d->n_levels; ++l) {
      const int rc = run_head(l, st);
      if (rc) return rc;
    }
  if (fork)  // join: the caller's stream continues when both side streams are done
    for (int i = 0; i < 2; ++i) {
      DFB_CHECK_CUDA(cudaEventRecord(d->ev[3 + i], d->side[i]));
      DFB_CHECK_CUDA(cudaStreamWaitEvent(st, d->ev[3 + i], 0));
    }
  return DFB_OK;
}

extern "C" int dfb_dfnet_fwd(DfbDfnet* d, const float* x, int B, int H, int W, uint32_t flags, int upH, int upW,
                             float* feats_t, float* feats_r, float* pose, void* ws, size_t ws_bytes, void* stream) {
  DFB_REQUIRE(d && d->loaded && x, DFB_ERR_INVALID, "DFNet handle not loaded or null input");
  const bool ret_feat = flags & 1, single = flags & 2, ret_pose = flags & 4, bf = flags & 16;
  DFB_REQUIRE(!ret_feat || feats_t, DFB_ERR_INVALID, "feature output missing");
  DFB_REQUIRE(!ret_feat || single || (feats_r && B % 2 == 0), DFB_ERR_INVALID, "siamese mode needs an even batch and feats_r");
  DFB_REQUIRE(!ret_pose || pose, DFB_ERR_INVALID, "pose output missing");
  DFB_REQUIRE(H >= 32 && W >= 32, DFB_ERR_INVALID, "image must be at least 32x32");
  DFB_REQUIRE(!(bf && ret_feat), DFB_ERR_UNSUPPORTED, "bf16 operands cover the pose-only path (the adaptation heads run in fp16)");
  if (bf) return dfnet_fwd_impl<__nv_bfloat16>(d, x, B, H, W, flags, upH, upW, feats_t, feats_r, pose, ws, ws_bytes, stream);
  return dfnet_fwd_impl<__half>(d, x, B, H, W, flags, upH, upW, feats_t, feats_r, pose, ws, ws_bytes, stream);
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
    k_cos_rows_final<<<1, 1024, 0, st>>>((const float*)ws, C, splits, eps, loss);
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

// adjoint of k_triplet_hinge_partial: d loss / d f1, d f2 for the chosen case (the case selection itself is a
// no-grad block in the reference, feature/misc.py:414-421).  One warp per row; every element receives at most two
// contributions (as anchor or positive of its own row, as negative of the row that rolls onto it), so the
// atomic adds onto the zeroed outputs are order independent.
__global__ void k_triplet_hinge_bwd(const float* __restrict__ f1, const float* __restrict__ f2, int B, int64_t rows_per_b, int W,
                                    int64_t n_rows, float margin, const int* __restrict__ chosen, const float* __restrict__ g_loss,
                                    float* __restrict__ g_f1, float* __restrict__ g_f2) {
  const int cs = *chosen;
  const bool a1 = (cs == 0 || cs == 2), n2 = (cs == 0 || cs == 3);
  const float* A = a1 ? f1 : f2;
  const float* P = a1 ? f2 : f1;
  const float* N = n2 ? f2 : f1;
  float* gA = a1 ? g_f1 : g_f2;
  float* gP = a1 ? g_f2 : g_f1;
  float* gN = n2 ? g_f2 : g_f1;
  const float coef = *g_loss / (float)n_rows;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
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
    const float sap = sqrtf(dap), san = sqrtf(dan);
    if (!(sap - san + margin > 0.f)) continue;   // clamp_min(., 0): no gradient on the flat side
    const float iap = sap > 0.f ? coef / sap : 0.f, ian = san > 0.f ? coef / san : 0.f;
    for (int x = lane; x < W; x += 32) {
      const float av = a[x];
      const float u = (av - p[x] + 1e-6f) * iap, v = (av - ng[x] + 1e-6f) * ian;
      atomicAdd(gA + row * W + x, u - v);
      atomicAdd(gP + row * W + x, -u);
      atomicAdd(gN + rrow * W + x, v);
    }
  }
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

// Backward of dfb_triplet_loss: chosen_case is the forward's device scalar, g_loss the upstream gradient (device
// scalar); g_f1, g_f2 [L,B,C,H,W] are overwritten.
extern "C" int dfb_triplet_loss_bwd(const float* f1, const float* f2, int L, int B, int Cc, int H, int W, float margin,
                                    const int* chosen_case, const float* g_loss, float* g_f1, float* g_f2, void* stream) {
  DFB_REQUIRE(f1 && f2 && chosen_case && g_loss && g_f1 && g_f2, DFB_ERR_INVALID, "null argument");
  DFB_REQUIRE(L >= 1 && B >= 1 && Cc >= 1 && H >= 1 && W >= 1, DFB_ERR_INVALID, "empty feature stack");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t rows_per_b = (int64_t)Cc * H, n_rows = (int64_t)L * B * rows_per_b;
  DFB_CHECK_CUDA(cudaMemsetAsync(g_f1, 0, (size_t)n_rows * W * 4, st));
  DFB_CHECK_CUDA(cudaMemsetAsync(g_f2, 0, (size_t)n_rows * W * 4, st));
  const int nb = (int)std::min<int64_t>(4096, (n_rows + 7) / 8);
  dfb::k_triplet_hinge_bwd<<<nb, 256, 0, st>>>(f1, f2, B, rows_per_b, W, n_rows, margin, chosen_case, g_loss, g_f1, g_f2);
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
  return launch_resize_bilinear_ac(src, dst, planes, h, w, Ho, Wo, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------
// backward of the losses and of the bicubic resampling (train_on_batch, direct_feature_matching.py:342-378)
// ------------------------------------------------------------------------------------------
namespace dfb {

// d/d fa of  1 - mean_rows cos(fa_r, fb_r):  -(g/rows) * ( fb/(na*nb) - ab*fa/(na^3*nb) ), with the eps clamps of the
// forward (a clamped norm is a constant).  stats: ws[row][split][3] from k_cos_rows_partial.
__global__ void k_cos_rows_bwd(const float* __restrict__ fa, const float* __restrict__ fb, const float* __restrict__ ws, int splits,
                               int64_t cols, int rows, float eps, const float* __restrict__ g_loss, float* __restrict__ g_fa) {
  const int row = blockIdx.y;
  float ab = 0.f, aa = 0.f, bb = 0.f;
  for (int s = 0; s < splits; ++s) {
    const float* o = ws + ((int64_t)row * splits + s) * 3;
    ab += o[0], aa += o[1], bb += o[2];
  }
  const float na = sqrtf(aa), nb = sqrtf(bb);
  const float ca = fmaxf(na, eps), cb = fmaxf(nb, eps);
  const float g = -__ldg(g_loss) / (float)rows;
  const float k1 = g / (ca * cb);
  const float k2 = na > eps ? g * ab / (ca * ca * ca * cb) : 0.f;
  const float4* a4 = reinterpret_cast<const float4*>(fa + (int64_t)row * cols);
  const float4* b4 = reinterpret_cast<const float4*>(fb + (int64_t)row * cols);
  float4* o4 = reinterpret_cast<float4*>(g_fa + (int64_t)row * cols);
  const int64_t n4 = (cols % 4 == 0 && ((int64_t)row * cols) % 4 == 0) ? cols / 4 : 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 x = __ldg(a4 + i), y = __ldg(b4 + i);
    o4[i] = make_float4(k1 * y.x - k2 * x.x, k1 * y.y - k2 * x.y, k1 * y.z - k2 * x.z, k1 * y.w - k2 * x.w);
  }
  for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cols; i += (int64_t)gridDim.x * blockDim.x)
    g_fa[(int64_t)row * cols + i] = k1 * fb[(int64_t)row * cols + i] - k2 * fa[(int64_t)row * cols + i];
}

// g_a = g * 2 (a - b) / n
__global__ void k_mse_bwd(const float* __restrict__ a, const float* __restrict__ b, int64_t n, const float* __restrict__ g_loss,
                          float* __restrict__ g_a) {
  const float k = 2.f * __ldg(g_loss) / (float)n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    g_a[i] = k * (a[i] - b[i]);
}

// adjoint of k_resize_bicubic: every output gradient is scattered to its 16 (border-clamped) taps
__global__ void k_resize_bicubic_bwd(const float* __restrict__ gdst, float* __restrict__ gsrc, int h, int w, int Ho, int Wo) {
  const int xo = blockIdx.x * blockDim.x + threadIdx.x, yo = blockIdx.y;
  if (xo >= Wo) return;
  const int64_t pl = blockIdx.z;
  const float sy = (float)h / (float)Ho, sx = (float)w / (float)Wo;
  const float fy = sy * (yo + 0.5f) - 0.5f, fx = sx * (xo + 0.5f) - 0.5f;
  const int iy = (int)floorf(fy), ix = (int)floorf(fx);
  float cy[4], cx[4];
  cubic_coeffs(fy - iy, cy);
  cubic_coeffs(fx - ix, cx);
  float* s = gsrc + pl * h * w;
  const float g = __ldg(gdst + (pl * Ho + yo) * Wo + xo);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int yy = min(max(iy - 1 + i, 0), h - 1);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int xx = min(max(ix - 1 + j, 0), w - 1);
      atomicAdd(s + yy * w + xx, cy[i] * cx[j] * g);
    }
  }
}

}  // namespace dfb

// gradient of dfb_cosine_loss w.r.t. fr; g_loss: device scalar (upstream gradient); ws as in the forward.
extern "C" int dfb_cosine_loss_bwd(const float* fr, const float* ft, int C, int64_t HW, int per_channel, float eps,
                                   const float* g_loss, float* g_fr, void* ws, size_t ws_bytes, void* stream) {
  DFB_REQUIRE(fr && ft && g_loss && g_fr && ws && C >= 1 && HW >= 1, DFB_ERR_INVALID, "null or empty argument");
  // per_channel bit 1: ws still holds the row statistics dfb_cosine_loss left there for the same fr / ft - the pass that
  // recomputes them (a second read of both 157 MB stacks at 480x640) is skipped
  const bool have_stats = (per_channel & 2) != 0;
  per_channel &= 1;
  DFB_REQUIRE(!per_channel, DFB_ERR_UNSUPPORTED, "the backward covers the reference default per_channel=False");
  cudaStream_t st = (cudaStream_t)stream;
  const int splits = (int)std::min<int64_t>(64, std::max<int64_t>(1, HW / 4096));
  DFB_REQUIRE(ws_bytes >= (size_t)C * splits * 3 * 4, DFB_ERR_WORKSPACE, "workspace too small");
  if (!have_stats) {
    k_cos_rows_partial<<<dim3(splits, C), 256, 0, st>>>(fr, ft, HW, splits, (float*)ws);
    DFB_LAUNCH_CHECK();
  }
  const int bx = (int)std::min<int64_t>(64, std::max<int64_t>(1, HW / 2048));
  k_cos_rows_bwd<<<dim3(bx, C), 256, 0, st>>>(fr, ft, (const float*)ws, splits, HW, C, eps, g_loss, g_fr);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

// gradient of dfb_mse w.r.t. a
extern "C" int dfb_mse_bwd(const float* a, const float* b, int64_t n, const float* g_loss, float* g_a, void* stream) {
  DFB_REQUIRE(a && b && g_loss && g_a && n >= 1, DFB_ERR_INVALID, "null or empty argument");
  const int blocks = (int)std::min<int64_t>(1184, (n + 255) / 256);
  k_mse_bwd<<<blocks, 256, 0, (cudaStream_t)stream>>>(a, b, n, g_loss, g_a);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

// adjoint of dfb_resize_bicubic: g_dst [P,Ho,Wo] -> g_src [P,h,w] (overwritten)
extern "C" int dfb_resize_bicubic_bwd(const float* g_dst, int64_t planes, int h, int w, int Ho, int Wo, float* g_src, void* stream) {
  DFB_REQUIRE(g_dst && g_src && planes >= 1 && h >= 1 && w >= 1 && Ho >= 1 && Wo >= 1, DFB_ERR_INVALID, "bad arguments");
  DFB_REQUIRE(planes <= 65535, DFB_ERR_INVALID, "too many planes");
  DFB_CHECK_CUDA(cudaMemsetAsync(g_src, 0, (size_t)planes * h * w * 4, (cudaStream_t)stream));
  k_resize_bicubic_bwd<<<dim3((Wo + 127) / 128, Ho, (unsigned)planes), 128, 0, (cudaStream_t)stream>>>(g_dst, g_src, h, w, Ho, Wo);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}
