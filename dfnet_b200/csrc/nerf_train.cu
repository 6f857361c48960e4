// Kernels of the NeRF-Hist TRAINING step (reference script/run_nerf.py:32-80, models/losses.py:19-57).
//
// The training forward / backward of the NeRF-W MLPs runs layer by layer on the tcgen05 convolution machinery of the
// DFNet path: a Linear layer over P samples is a 1x1 convolution over a P-pixel image, so k_conv_tc (forward and
// data-gradient with the ReLU-mask epilogue) and k_conv_wgrad (weight gradient, MN-major operands) apply unchanged; the
// orchestration lives in dfnet_b200/nerf_train.py.  This file holds what is specific to NeRF around those layers:
//
//   k_embed_xyz16          pts = o + d*z and the positional encoding (models/nerfw.py:105-133) as the first layer's
//                          NHWC 16-bit operand [P, 64]
//   k_rows_expand16        per-ray pre-activation contribution (view direction / appearance / transient codes through
//                          their slice of dir_encoding.0 / transient_encoding.0) broadcast to the ray's samples
//   k_rows_reduce          its adjoint: sum of a [P, C] gradient over the samples of each ray
//   k_heads_fwd / _bwd     Softplus / Sigmoid heads (models/nerfw.py:275-295): fp32 pre-activations -> raw, and
//                          d raw -> d pre-activation from the stored outputs (sigmoid' = o (1 - o), softplus' = 1 - e^-o)
//   k_raw2outputs_bwd      adjoint of raw2outputs_NeRFW (models/rendering.py:132-243) in train mode w.r.t. raw, for the
//                          outputs the NeRF-W loss reads: rgb (coarse / fine), beta, transient_sigmas
//   k_cast_f16_bf16        forward activations (fp16) as bf16 operands of the weight-gradient kernel
#include <algorithm>

#include "common.cuh"

namespace dfb {

// Eight threads per sample, one 16-byte store each (a warp writes 512 contiguous bytes; the first version had one thread
// per sample writing its 64 halves one at a time, 128 bytes apart from its neighbours: 115 us for 25 MB).  out_bf (nullable):
// the same values as bf16 - the operand type of the first layer's weight gradient - instead of a separate conversion pass.
__global__ void __launch_bounds__(256) k_embed_xyz16(const float* __restrict__ rays, int ray_stride, const float* __restrict__ z,
                                                     int64_t P, int S, int L, int ld, __half* __restrict__ out,
                                                     __nv_bfloat16* __restrict__ out_bf) {
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int per = ld >> 3;                  // 16-byte chunks per sample
  const int64_t g = t / per;
  if (g >= P) return;
  const int c0 = (int)(t - g * per) * 8;
  const float* r = rays + (g / S) * ray_stride;
  const float zz = z[g];
  float pt[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) pt[c] = __fadd_rn(r[c], __fmul_rn(r[3 + c], zz));
  const int n_ch = 3 + 6 * L;
  __align__(16) __half h[8];
  __align__(16) __nv_bfloat16 hb[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = c0 + e;
    float v = 0.f;
    if (c < 3) {
      v = pt[c];
    } else if (c < n_ch) {
      const int q = c - 3, l = q / 6, rr = q - 6 * l, comp = rr >= 3 ? rr - 3 : rr;
      float sn, cs;
      sincosf(__fmul_rn(pt[comp], (float)(1 << l)), &sn, &cs);
      v = rr >= 3 ? cs : sn;
    }
    h[e] = __float2half_rn(v);
    hb[e] = __float2bfloat16_rn(__half2float(h[e]));
  }
  *reinterpret_cast<uint4*>(out + g * ld + c0) = *reinterpret_cast<const uint4*>(h);
  if (out_bf) *reinterpret_cast<uint4*>(out_bf + g * ld + c0) = *reinterpret_cast<const uint4*>(hb);
}

// out[p, c] = fp16(rb[ray(p), c])
__global__ void __launch_bounds__(256) k_rows_expand16(const float* __restrict__ rb, int64_t P, int S, int C, __half* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;   // one thread per 8 channels
  const int c8 = C / 8;
  if (i >= P * c8) return;
  const int64_t p = i / c8;
  const int c = (int)(i % c8) * 8;
  const float* s = rb + (p / S) * C + c;
  __half2 h[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) h[k] = __floats2half2_rn(s[2 * k], s[2 * k + 1]);
  *reinterpret_cast<uint4*>(out + p * C + c) = *reinterpret_cast<uint4*>(h);
}

// out[r, c] = sum_{s < S} g[(r*S + s), c]  (g bf16, accumulation fp32); one block per ray, one thread per channel
__global__ void k_rows_reduce(const __nv_bfloat16* __restrict__ g, int S, int C, float* __restrict__ out) {
  const int64_t r = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    const __nv_bfloat16* p = g + (r * S) * C + c;
    for (int s = 0; s < S; ++s) acc += __bfloat162float(p[(int64_t)s * C]);
    out[r * C + c] = acc;
  }
}

__device__ __forceinline__ float nt_softplus(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float nt_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

// pre-activations are fp32 planes [channel][P] (the NCHW output of the head convolutions, plane stride `ps`)
__global__ void __launch_bounds__(256) k_heads_fwd(const float* __restrict__ sig, const float* __restrict__ rgb,
                                                   const float* __restrict__ tr, int64_t P, int64_t ps, int C,
                                                   float* __restrict__ raw) {
  const int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (p >= P) return;
  float* o = raw + p * C;
  for (int c = 0; c < 3; ++c) o[c] = nt_sigmoid(rgb[c * ps + p]);
  o[3] = nt_softplus(sig[p]);
  if (C == 9) {   // transient head planes: rgb(3), sigma, beta
    for (int c = 0; c < 3; ++c) o[4 + c] = nt_sigmoid(tr[c * ps + p]);
    o[7] = nt_softplus(tr[3 * ps + p]);
    o[8] = nt_softplus(tr[4 * ps + p]);
  }
}

// d raw -> d pre-activation as bf16 NHWC [P, 64] operands (columns past the head's width are zero):
// g_sig[:,0]; g_rgb[:,0:3]; g_tr[:,0:5] = transient rgb(3), sigma, beta
// Eight threads per sample: thread j writes the 16-byte chunk j of each [P, 64] row (a warp stores 512 contiguous bytes per
// array); chunk 0 carries the values, the others are zeros.  (One thread per sample writing its 3 x 64 elements one by one
// took 372 us for the fine pass of a 1536-ray step, 75 MB.)
__global__ void __launch_bounds__(256) k_heads_bwd(const float* __restrict__ raw, const float* __restrict__ g_raw, int64_t P, int C,
                                                   __nv_bfloat16* __restrict__ g_sig, __nv_bfloat16* __restrict__ g_rgb,
                                                   __nv_bfloat16* __restrict__ g_tr) {
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t p = t >> 3;
  if (p >= P) return;
  const int j = (int)(t & 7);
  const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);
  __align__(16) __nv_bfloat16 a[8], b[8], tr[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) a[c] = zero, b[c] = zero, tr[c] = zero;
  if (j == 0) {
    const float* o = raw + p * C;
    const float* g = g_raw + p * C;
    for (int c = 0; c < 3; ++c) b[c] = __float2bfloat16_rn(g[c] * o[c] * (1.f - o[c]));
    a[0] = __float2bfloat16_rn(g[3] * (1.f - expf(-o[3])));
    if (C == 9) {
      for (int c = 4; c < 7; ++c) tr[c - 4] = __float2bfloat16_rn(g[c] * o[c] * (1.f - o[c]));
      tr[3] = __float2bfloat16_rn(g[7] * (1.f - expf(-o[7])));
      tr[4] = __float2bfloat16_rn(g[8] * (1.f - expf(-o[8])));
    }
  }
  const int64_t off = p * 64 + j * 8;
  *reinterpret_cast<uint4*>(g_sig + off) = *reinterpret_cast<const uint4*>(a);
  *reinterpret_cast<uint4*>(g_rgb + off) = *reinterpret_cast<const uint4*>(b);
  if (C == 9) *reinterpret_cast<uint4*>(g_tr + off) = *reinterpret_cast<const uint4*>(tr);
}

// One thread per ray.  C == 9 (fine, train mode): loss inputs rgb = sum T (a_s c_s + a_t c_t), beta = sum T a_t b + beta_min,
// transient_sigmas = sigma_t.  C == 4 (coarse, train mode): rgb = sum T a c, a = 1 - exp(-delta relu(sigma + noise std)).
// With E_i the upstream-weighted emission of sample i, d/d a_j of sum_{i>j} T_i E_i is -T_j R_j with the backward
// recurrence R_j = E_{j+1} + (1 - a_{j+1}) R_{j+1} (no division by 1 - a).
__global__ void __launch_bounds__(128) k_raw2outputs_bwd(const float* __restrict__ raw, const float* __restrict__ z, int64_t N, int S,
                                                         int C, const float* __restrict__ noise, float noise_std,
                                                         const float* __restrict__ g_rgb, const float* __restrict__ g_beta,
                                                         const float* __restrict__ g_tsig, float* __restrict__ g_raw) {
  const int64_t ray = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (ray >= N) return;
  const float* rw = raw + ray * S * C;
  const float* zz = z + ray * S;
  float* go = g_raw + ray * S * C;
  const float g0 = g_rgb ? g_rgb[ray * 3] : 0.f, g1 = g_rgb ? g_rgb[ray * 3 + 1] : 0.f, g2 = g_rgb ? g_rgb[ray * 3 + 2] : 0.f;
  const float gb = g_beta ? g_beta[ray] : 0.f;
  auto delta_at = [&](int i) { return (i + 1 < S) ? __fsub_rn(zz[i + 1], zz[i]) : 1e2f; };
  // pass 1 (front to back): transmittance in front of every sample, parked in g_raw's sigma slot
  double T = 1.0;
  for (int i = 0; i < S; ++i) {
    const float d = delta_at(i);
    float a;
    if (C == 9) a = 1.f - expf(-d * (rw[i * 9 + 3] + rw[i * 9 + 7]));
    else a = 1.f - expf(-d * fmaxf(rw[i * C + 3] + (noise ? noise[ray * S + i] * noise_std : 0.f), 0.f));
    go[i * C + 3] = (float)T;
    T *= (double)(1.f - a);
  }
  // pass 2 (back to front)
  double R = 0.0;
  for (int i = S - 1; i >= 0; --i) {
    const float d = delta_at(i), Ti = go[i * C + 3];
    if (C == 9) {
      const float ss = rw[i * 9 + 3], st = rw[i * 9 + 7];
      const float es = expf(-d * ss), et = expf(-d * st), ea = expf(-d * (ss + st));
      const float as = 1.f - es, at = 1.f - et;
      const float gcs = g0 * rw[i * 9] + g1 * rw[i * 9 + 1] + g2 * rw[i * 9 + 2];
      const float gct = g0 * rw[i * 9 + 4] + g1 * rw[i * 9 + 5] + g2 * rw[i * 9 + 6] + gb * rw[i * 9 + 8];
      const float E = as * gcs + at * gct;
      const float dA = -Ti * (float)R;             // through the transmittance of the samples behind
      go[i * 9 + 0] = g0 * as * Ti, go[i * 9 + 1] = g1 * as * Ti, go[i * 9 + 2] = g2 * as * Ti;
      go[i * 9 + 4] = g0 * at * Ti, go[i * 9 + 5] = g1 * at * Ti, go[i * 9 + 6] = g2 * at * Ti;
      go[i * 9 + 8] = gb * at * Ti;
      go[i * 9 + 3] = d * (es * Ti * gcs + ea * dA);
      go[i * 9 + 7] = d * (et * Ti * gct + ea * dA) + (g_tsig ? g_tsig[ray * S + i] : 0.f);
      R = (double)E + (double)ea * R;
    } else {
      const float pre = rw[i * C + 3] + (noise ? noise[ray * S + i] * noise_std : 0.f);
      const float sg = fmaxf(pre, 0.f);
      const float ea = expf(-d * sg), a = 1.f - ea;
      const float gc = g0 * rw[i * C] + g1 * rw[i * C + 1] + g2 * rw[i * C + 2];
      const float E = a * gc;
      go[i * C + 0] = g0 * a * Ti, go[i * C + 1] = g1 * a * Ti, go[i * C + 2] = g2 * a * Ti;
      go[i * C + 3] = pre > 0.f ? d * ea * (Ti * gc - Ti * (float)R) : 0.f;
      R = (double)E + (double)ea * R;
    }
  }
}

// The same adjoint with one WARP per ray (S <= 256): lane l owns the K = ceil(S / 32) consecutive samples from l K on.  The
// transmittance is an exclusive prefix product (per-lane products, then a shuffle scan), the backward recurrence
// R_j = E_{j+1} + ea_{j+1} R_{j+1} a suffix scan of affine maps (A, B): R -> B + A R; both in double like the sequential
// version, whose results this reproduces up to the association of those products.  A ray's raw / g_raw are contiguous and
// every lane reads and writes a contiguous piece.  (One thread per ray: 1536 threads on the whole GPU, every load a
// dependent, uncoalesced one - 330 us for the fine pass of a 1536-ray step; this one is ~10x faster.)
template <int C>
__global__ void __launch_bounds__(128) k_raw2outputs_bwd_warp(const float* __restrict__ raw, const float* __restrict__ z, int64_t N, int S,
                                                              const float* __restrict__ noise, float noise_std,
                                                              const float* __restrict__ g_rgb, const float* __restrict__ g_beta,
                                                              const float* __restrict__ g_tsig, float* __restrict__ g_raw) {
  constexpr unsigned kFull = 0xffffffffu;
  constexpr int kMaxK = 8;
  const int64_t ray = ((int64_t)blockIdx.x * 128 + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (ray >= N) return;   // uniform per warp
  const int K = (S + 31) >> 5;
  const int i0 = lane * K;
  const float* rw = raw + ray * S * C;
  const float* zz = z + ray * S;
  float* go = g_raw + ray * S * C;
  const float g0 = g_rgb ? g_rgb[ray * 3] : 0.f, g1 = g_rgb ? g_rgb[ray * 3 + 1] : 0.f, g2 = g_rgb ? g_rgb[ray * 3 + 2] : 0.f;
  const float gb = g_beta ? g_beta[ray] : 0.f;
  float dl[kMaxK], Ti[kMaxK];
  // pass 1: transmittance in front of every sample
  double prod = 1.0;
#pragma unroll
  for (int k = 0; k < kMaxK; ++k) {
    const int i = i0 + k;
    if (k < K && i < S) {
      const float d = (i + 1 < S) ? __fsub_rn(zz[i + 1], zz[i]) : 1e2f;
      float a;
      if (C == 9) a = 1.f - expf(-d * (rw[i * 9 + 3] + rw[i * 9 + 7]));
      else a = 1.f - expf(-d * fmaxf(rw[i * C + 3] + (noise ? noise[ray * S + i] * noise_std : 0.f), 0.f));
      dl[k] = d;
      Ti[k] = __double2float_rn(prod);     // product of this lane's earlier samples, scaled by the lanes in front below
      prod *= (double)(1.f - a);
    } else {
      dl[k] = 0.f, Ti[k] = 0.f;
    }
  }
  double incl = prod;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const double o = __shfl_up_sync(kFull, incl, d);
    if (lane >= d) incl *= o;
  }
  double excl = __shfl_up_sync(kFull, incl, 1);
  if (lane == 0) excl = 1.0;
  // (redo the per-lane products in double so that T_i = excl * prod_k is rounded once, like the sequential version)
  {
    double pk = 1.0;
#pragma unroll
    for (int k = 0; k < kMaxK; ++k) {
      const int i = i0 + k;
      if (k < K && i < S) {
        float a;
        if (C == 9) a = 1.f - expf(-dl[k] * (rw[i * 9 + 3] + rw[i * 9 + 7]));
        else a = 1.f - expf(-dl[k] * fmaxf(rw[i * C + 3] + (noise ? noise[ray * S + i] * noise_std : 0.f), 0.f));
        Ti[k] = (float)(excl * pk);
        pk *= (double)(1.f - a);
      }
    }
  }
  // pass 2: this lane's affine map (applied back to front), suffix scan over the lanes, then the samples
  float Ek[kMaxK], eak[kMaxK];
  double A = 1.0, B = 0.0;
#pragma unroll
  for (int k = kMaxK - 1; k >= 0; --k) {
    const int i = i0 + k;
    Ek[k] = 0.f, eak[k] = 1.f;
    if (k < K && i < S) {
      const float d = dl[k];
      if (C == 9) {
        const float ss = rw[i * 9 + 3], st = rw[i * 9 + 7];
        const float es = expf(-d * ss), et = expf(-d * st), ea = expf(-d * (ss + st));
        const float gcs = g0 * rw[i * 9] + g1 * rw[i * 9 + 1] + g2 * rw[i * 9 + 2];
        const float gct = g0 * rw[i * 9 + 4] + g1 * rw[i * 9 + 5] + g2 * rw[i * 9 + 6] + gb * rw[i * 9 + 8];
        Ek[k] = (1.f - es) * gcs + (1.f - et) * gct, eak[k] = ea;
      } else {
        const float sg = fmaxf(rw[i * C + 3] + (noise ? noise[ray * S + i] * noise_std : 0.f), 0.f);
        const float ea = expf(-d * sg);
        const float gc = g0 * rw[i * C] + g1 * rw[i * C + 1] + g2 * rw[i * C + 2];
        Ek[k] = (1.f - ea) * gc, eak[k] = ea;
      }
      B = (double)Ek[k] + (double)eak[k] * B;
      A = (double)eak[k] * A;
    }
  }
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const double A2 = __shfl_down_sync(kFull, A, d), B2 = __shfl_down_sync(kFull, B, d);
    if (lane + d < 32) B = B + A * B2, A = A * A2;
  }
  double R = __shfl_down_sync(kFull, B, 1);
  if (lane == 31) R = 0.0;
#pragma unroll
  for (int k = kMaxK - 1; k >= 0; --k) {
    const int i = i0 + k;
    if (k < K && i < S) {
      const float d = dl[k], T = Ti[k];
      if (C == 9) {
        const float ss = rw[i * 9 + 3], st = rw[i * 9 + 7];
        const float es = expf(-d * ss), et = expf(-d * st), ea = eak[k];
        const float as = 1.f - es, at = 1.f - et;
        const float gcs = g0 * rw[i * 9] + g1 * rw[i * 9 + 1] + g2 * rw[i * 9 + 2];
        const float gct = g0 * rw[i * 9 + 4] + g1 * rw[i * 9 + 5] + g2 * rw[i * 9 + 6] + gb * rw[i * 9 + 8];
        const float dA = -T * (float)R;
        go[i * 9 + 0] = g0 * as * T, go[i * 9 + 1] = g1 * as * T, go[i * 9 + 2] = g2 * as * T;
        go[i * 9 + 4] = g0 * at * T, go[i * 9 + 5] = g1 * at * T, go[i * 9 + 6] = g2 * at * T;
        go[i * 9 + 8] = gb * at * T;
        go[i * 9 + 3] = d * (es * T * gcs + ea * dA);
        go[i * 9 + 7] = d * (et * T * gct + ea * dA) + (g_tsig ? g_tsig[ray * S + i] : 0.f);
      } else {
        const float pre = rw[i * C + 3] + (noise ? noise[ray * S + i] * noise_std : 0.f);
        const float ea = eak[k], a = 1.f - ea;
        const float gc = g0 * rw[i * C] + g1 * rw[i * C + 1] + g2 * rw[i * C + 2];
        go[i * C + 0] = g0 * a * T, go[i * C + 1] = g1 * a * T, go[i * C + 2] = g2 * a * T;
        go[i * C + 3] = pre > 0.f ? d * ea * (T * gc - T * (float)R) : 0.f;
      }
      R = (double)Ek[k] + (double)eak[k] * R;
    }
  }
}

__global__ void __launch_bounds__(256) k_cast_f16_bf16(const __half* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
  const int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 2;
  if (i + 1 < n) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(src + i));
    *reinterpret_cast<__nv_bfloat162*>(dst + i) = __floats2bfloat162_rn(f.x, f.y);
  } else if (i < n) {
    dst[i] = __float2bfloat16_rn(__half2float(src[i]));
  }
}

// ---- NeRF-W loss (models/losses.py:42-57), all four terms in one pass ---------------------------------------------------
//   c_l = coef * 0.5 mean((rgb_c - t)^2)      f_l = coef * mean((rgb_f - t)^2 / (2 beta^2))
//   b_l = coef * (3 + mean(log beta))         s_l = coef * lambda_u * mean(transient_sigmas)
// As tensor expressions these are ~25 launches forward and ~35 backward on a few thousand values, all host time in a
// launch-bound training step.  Partial sums in double per block, summed in block order by a second, single-block launch
// (deterministic).  out[4] = the four terms, out[4] (fifth value) = mean((rgb_f - t)^2) for the PSNR read-out.
constexpr int kLossBlocks = 64;

__global__ void __launch_bounds__(256) k_nerfw_loss_partial(const float* __restrict__ rc, const float* __restrict__ rf,
                                                            const float* __restrict__ beta, const float* __restrict__ ts,
                                                            const float* __restrict__ tg, int64_t N, int S, double* __restrict__ ws) {
  double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  const int64_t stride = (int64_t)gridDim.x * 256, i0 = (int64_t)blockIdx.x * 256 + threadIdx.x;
  for (int64_t n = i0; n < N; n += stride) {
    const float b = beta[n];
    float sc = 0.f, sf = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float t = tg[n * 3 + c], dc = rc[n * 3 + c] - t, df = rf[n * 3 + c] - t;
      sc += dc * dc, sf += df * df;
    }
    acc[0] += (double)sc, acc[1] += (double)(sf / (2.f * b * b)), acc[2] += (double)logf(b), acc[4] += (double)sf;
  }
  for (int64_t i = i0; i < N * S; i += stride) acc[3] += (double)ts[i];
  __shared__ double sm[5][256];
#pragma unroll
  for (int k = 0; k < 5; ++k) sm[k][threadIdx.x] = acc[k];
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) {
#pragma unroll
      for (int k = 0; k < 5; ++k) sm[k][threadIdx.x] += sm[k][threadIdx.x + w];
    }
    __syncthreads();
  }
  if (threadIdx.x < 5) ws[blockIdx.x * 5 + threadIdx.x] = sm[threadIdx.x][0];
}

__global__ void k_nerfw_loss_final(const double* __restrict__ ws, int blocks, int64_t N, int S, float coef, float lambda_u,
                                   float* __restrict__ out) {
  if (threadIdx.x >= 5) return;
  double s = 0.0;
  for (int b = 0; b < blocks; ++b) s += ws[b * 5 + threadIdx.x];
  const double n3 = 3.0 * (double)N, ns = (double)N * (double)S;
  double v;
  switch (threadIdx.x) {
    case 0: v = coef * 0.5 * s / n3; break;
    case 1: v = coef * s / n3; break;
    case 2: v = coef * (3.0 + s / (double)N); break;
    case 3: v = coef * lambda_u * s / ns; break;
    default: v = s / n3; break;
  }
  out[threadIdx.x] = (float)v;
}

// adjoint: g[4] are the upstream gradients of the four terms (device scalars, nullable = 0)
__global__ void __launch_bounds__(256) k_nerfw_loss_bwd(const float* __restrict__ rc, const float* __restrict__ rf,
                                                        const float* __restrict__ beta, const float* __restrict__ tg, int64_t N, int S,
                                                        float coef, float lambda_u, const float* __restrict__ gc_p,
                                                        const float* __restrict__ gf_p, const float* __restrict__ gb_p,
                                                        const float* __restrict__ gs_p, float* __restrict__ g_rc,
                                                        float* __restrict__ g_rf, float* __restrict__ g_beta, float* __restrict__ g_ts) {
  const float gc = gc_p ? *gc_p : 0.f, gf = gf_p ? *gf_p : 0.f, gb = gb_p ? *gb_p : 0.f, gs = gs_p ? *gs_p : 0.f;
  const float inv3n = 1.f / (3.f * (float)N), invn = 1.f / (float)N;
  const int64_t stride = (int64_t)gridDim.x * 256, i0 = (int64_t)blockIdx.x * 256 + threadIdx.x;
  for (int64_t n = i0; n < N; n += stride) {
    const float b = beta[n], ib2 = 1.f / (b * b);
    float sf = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float t = tg[n * 3 + c], dc = rc[n * 3 + c] - t, df = rf[n * 3 + c] - t;
      if (g_rc) g_rc[n * 3 + c] = gc * coef * dc * inv3n;
      if (g_rf) g_rf[n * 3 + c] = gf * coef * df * ib2 * inv3n;
      sf += df * df;
    }
    if (g_beta) g_beta[n] = -gf * coef * sf * ib2 / b * inv3n + gb * coef * invn / b;
  }
  if (g_ts) {
    const float v = gs * coef * lambda_u / ((float)N * (float)S);
    for (int64_t i = i0; i < N * S; i += stride) g_ts[i] = v;
  }
}

}  // namespace dfb

using namespace dfb;

extern "C" size_t dfb_nerfw_loss_workspace_bytes(void) { return (size_t)kLossBlocks * 5 * sizeof(double); }

extern "C" int dfb_nerfw_loss_fwd(const float* rgb_coarse, const float* rgb_fine, const float* beta, const float* transient_sigmas,
                                  const float* targets, int64_t N, int S, float coef, float lambda_u, void* ws, float* out5, void* stream) {
  DFB_REQUIRE(rgb_coarse && rgb_fine && beta && transient_sigmas && targets && ws && out5 && N >= 1 && S >= 1, DFB_ERR_INVALID,
              "dfb_nerfw_loss_fwd: bad arguments");
  const int blocks = (int)std::min<int64_t>(kLossBlocks, (N * S + 255) / 256);
  k_nerfw_loss_partial<<<blocks, 256, 0, (cudaStream_t)stream>>>(rgb_coarse, rgb_fine, beta, transient_sigmas, targets, N, S, (double*)ws);
  DFB_LAUNCH_CHECK();
  k_nerfw_loss_final<<<1, 32, 0, (cudaStream_t)stream>>>((const double*)ws, blocks, N, S, coef, lambda_u, out5);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

extern "C" int dfb_nerfw_loss_bwd(const float* rgb_coarse, const float* rgb_fine, const float* beta, const float* targets, int64_t N, int S,
                                  float coef, float lambda_u, const float* g_c, const float* g_f, const float* g_b, const float* g_s,
                                  float* g_rgb_coarse, float* g_rgb_fine, float* g_beta, float* g_transient_sigmas, void* stream) {
  DFB_REQUIRE(rgb_coarse && rgb_fine && beta && targets && N >= 1 && S >= 1, DFB_ERR_INVALID, "dfb_nerfw_loss_bwd: bad arguments");
  const int blocks = (int)std::min<int64_t>(148 * 4, ((g_transient_sigmas ? N * S : N) + 255) / 256);
  k_nerfw_loss_bwd<<<blocks, 256, 0, (cudaStream_t)stream>>>(rgb_coarse, rgb_fine, beta, targets, N, S, coef, lambda_u, g_c, g_f, g_b, g_s,
                                                              g_rgb_coarse, g_rgb_fine, g_beta, g_transient_sigmas);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

static int embed_xyz16_impl(const float* rays, int ray_stride, const float* z, int64_t N, int S, int L, int ld, void* out, void* out_bf16,
                           void* stream) {
  DFB_REQUIRE(rays && z && out && N >= 0 && S >= 1 && L >= 0 && L <= 20 && ray_stride >= 6, DFB_ERR_INVALID, "dfb_embed_xyz16: bad arguments");
  DFB_REQUIRE(ld >= 3 + 6 * L && ld % 8 == 0, DFB_ERR_INVALID, "dfb_embed_xyz16: ld must be a multiple of 8 and >= 3 + 6 L");
  const int64_t P = N * S;
  if (P == 0) return DFB_OK;
  const int64_t threads = P * (ld / 8);
  k_embed_xyz16<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rays, ray_stride, z, P, S, L, ld, (__half*)out,
                                                                                     (__nv_bfloat16*)out_bf16);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

extern "C" int dfb_embed_xyz16(const float* rays, int ray_stride, const float* z, int64_t N, int S, int L, int ld, void* out,
                               void* stream) {
  return embed_xyz16_impl(rays, ray_stride, z, N, S, L, ld, out, nullptr, stream);
}

extern "C" int dfb_embed_xyz16_ex(const float* rays, int ray_stride, const float* z, int64_t N, int S, int L, int ld, void* out,
                                  void* out_bf16, void* stream) {
  return embed_xyz16_impl(rays, ray_stride, z, N, S, L, ld, out, out_bf16, stream);
}

extern "C" int dfb_rows_expand16(const float* rb, int64_t N, int S, int C, void* out, void* stream) {
  DFB_REQUIRE(rb && out && N >= 0 && S >= 1 && C >= 8 && C % 8 == 0, DFB_ERR_INVALID, "dfb_rows_expand16: bad arguments");
  const int64_t n = N * S * (C / 8);
  if (n == 0) return DFB_OK;
  k_rows_expand16<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rb, N * S, S, C, (__half*)out);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

extern "C" int dfb_rows_reduce_bf16(const void* g, int64_t N, int S, int C, float* out, void* stream) {
  DFB_REQUIRE(g && out && N >= 0 && S >= 1 && C >= 1, DFB_ERR_INVALID, "dfb_rows_reduce_bf16: bad arguments");
  if (N == 0) return DFB_OK;
  k_rows_reduce<<<(unsigned)N, 128, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)g, S, C, out);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

extern "C" int dfb_nerf_heads_fwd(const float* sig_pre, const float* rgb_pre, const float* tr_pre, int64_t P, int64_t plane_stride,
                                  int C, float* raw, void* stream) {
  DFB_REQUIRE(sig_pre && rgb_pre && raw && (C == 4 || (C == 9 && tr_pre)), DFB_ERR_INVALID, "dfb_nerf_heads_fwd: bad arguments");
  if (P == 0) return DFB_OK;
  k_heads_fwd<<<(unsigned)((P + 255) / 256), 256, 0, (cudaStream_t)stream>>>(sig_pre, rgb_pre, tr_pre, P, plane_stride, C, raw);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

extern "C" int dfb_nerf_heads_bwd(const float* raw, const float* g_raw, int64_t P, int C, void* g_sig16, void* g_rgb16, void* g_tr16,
                                  void* stream) {
  DFB_REQUIRE(raw && g_raw && g_sig16 && g_rgb16 && (C == 4 || (C == 9 && g_tr16)), DFB_ERR_INVALID, "dfb_nerf_heads_bwd: bad arguments");
  if (P == 0) return DFB_OK;
  k_heads_bwd<<<(unsigned)((P * 8 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(raw, g_raw, P, C, (__nv_bfloat16*)g_sig16,
                                                                            (__nv_bfloat16*)g_rgb16, (__nv_bfloat16*)g_tr16);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

extern "C" int dfb_raw2outputs_bwd(const float* raw, const float* z_vals, int64_t N, int S, int C, const float* noise,
                                   float raw_noise_std, const float* g_rgb, const float* g_beta, const float* g_tsig, float* g_raw,
                                   void* stream) {
  DFB_REQUIRE(raw && z_vals && g_raw && (C == 4 || C == 9), DFB_ERR_INVALID, "dfb_raw2outputs_bwd: raw must be [N,S,4] or [N,S,9]");
  DFB_REQUIRE(raw_noise_std == 0.f || (noise && C == 4), DFB_ERR_INVALID, "dfb_raw2outputs_bwd: noise applies to the coarse pass");
  if (N == 0) return DFB_OK;
  const float* nz = raw_noise_std != 0.f ? noise : nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  const char* seq = getenv("DFB_R2O_BWD_SEQ");   // =1: the one-thread-per-ray version (A/B, tests)
  if (S <= 256 && !(seq && seq[0] == '1')) {
    const unsigned blocks = (unsigned)((N * 32 + 127) / 128);
    if (C == 9) k_raw2outputs_bwd_warp<9><<<blocks, 128, 0, st>>>(raw, z_vals, N, S, nz, raw_noise_std, g_rgb, g_beta, g_tsig, g_raw);
    else k_raw2outputs_bwd_warp<4><<<blocks, 128, 0, st>>>(raw, z_vals, N, S, nz, raw_noise_std, g_rgb, g_beta, g_tsig, g_raw);
  } else {
    k_raw2outputs_bwd<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(raw, z_vals, N, S, C, nz, raw_noise_std, g_rgb, g_beta, g_tsig, g_raw);
  }
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

extern "C" int dfb_cast_f16_bf16(const void* src, void* dst, int64_t n, void* stream) {
  DFB_REQUIRE(src && dst && n >= 0, DFB_ERR_INVALID, "dfb_cast_f16_bf16: bad arguments");
  if (n == 0) return DFB_OK;
  k_cast_f16_bf16<<<(unsigned)((n / 2 + 256) / 256), 256, 0, (cudaStream_t)stream>>>((const __half*)src, (__nv_bfloat16*)dst, n);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}
