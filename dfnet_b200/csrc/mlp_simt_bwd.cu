// Backward of the test-time render w.r.t. the rays (rays_o, rays_d, viewdirs) — what train.py needs
// from the renderer (reference feature/direct_feature_matching.py:342-378: loss -> rgb -> render ->
// c2w -> pose net; the NeRF weights are frozen and z_samples are detached, models/rendering.py:302,
// so the gradient flows through the FINE network's inputs only).
//
//   k_composite_bwd   d rgb -> d raw[N,S,9]   (adjoint of models/rendering.py:169-212; suffix sums)
//   k_mlp_simt_bwd    recomputes the fine MLP forward per 64-sample tile keeping only the ReLU masks,
//                     then runs the input-gradient chain g_in = W^T (g_out . mask) through heads,
//                     transient branch, dir/transient layer, xyz_encoding_final, the trunk (skip
//                     layer splits into [pe | h]) and the positional encoding -> d pts, d dirPE
//   k_ray_grad        per ray: d rays_o = sum d pts, d rays_d = sum z * d pts, d viewdirs from d dirPE
// fp32 FFMA version (exact-order reference implementation of the backward; the tensor-core backward is
// a later step).
#include "common.cuh"

namespace dfb {

constexpr int kBT = 64;        // samples per CTA
constexpr int kBThreads = 256; // 8 warps x 8 samples
constexpr int kBRows = 8;

__device__ __forceinline__ float bw_softplus(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float bw_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

// ---------------------------------------------------------------------------------------------
// compositing backward (test_time fine compositing: rgb = sum_i T_i (a_s,i c_s,i + a_t,i c_t,i))
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_composite_bwd(const float* __restrict__ raw, const float* __restrict__ z,
                                                       const float* __restrict__ g_rgb, int64_t N, int S,
                                                       float* __restrict__ g_raw) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * 4 + warp;
  if (ray >= N) return;
  float* oma = sm + (size_t)warp * 3 * S;  // 1 - alpha
  float* T = oma + S;                      // transmittance
  float* tq = T + S;                       // T_i * q_i, then suffix sums R_i
  const float* rw = raw + ray * S * 9;
  const float* zz = z + ray * S;
  const float g0 = g_rgb[ray * 3], g1 = g_rgb[ray * 3 + 1], g2 = g_rgb[ray * 3 + 2];
  for (int i = lane; i < S; i += 32) {
    const float delta = (i + 1 < S) ? __fsub_rn(zz[i + 1], zz[i]) : 1e2f;
    oma[i] = expf(-delta * (rw[i * 9 + 3] + rw[i * 9 + 7]));
  }
  __syncwarp();
  if (lane == 0) {
    double t = 1.0;
    for (int i = 0; i < S; ++i) { T[i] = (float)t; t *= (double)oma[i]; }
  }
  __syncwarp();
  for (int i = lane; i < S; i += 32) {
    const float delta = (i + 1 < S) ? __fsub_rn(zz[i + 1], zz[i]) : 1e2f;
    const float as = 1.f - expf(-delta * rw[i * 9 + 3]), at = 1.f - expf(-delta * rw[i * 9 + 7]);
    const float gcs = g0 * rw[i * 9] + g1 * rw[i * 9 + 1] + g2 * rw[i * 9 + 2];
    const float gct = g0 * rw[i * 9 + 4] + g1 * rw[i * 9 + 5] + g2 * rw[i * 9 + 6];
    tq[i] = T[i] * (as * gcs + at * gct);
  }
  __syncwarp();
  if (lane == 0) {  // R_i = sum_{j>i} T_j q_j
    double acc = 0.0;
    for (int i = S - 1; i >= 0; --i) { const float v = tq[i]; tq[i] = (float)acc; acc += (double)v; }
  }
  __syncwarp();
  float* go = g_raw + ray * S * 9;
  for (int i = lane; i < S; i += 32) {
    const float delta = (i + 1 < S) ? __fsub_rn(zz[i + 1], zz[i]) : 1e2f;
    const float es = expf(-delta * rw[i * 9 + 3]), et = expf(-delta * rw[i * 9 + 7]);
    const float as = 1.f - es, at = 1.f - et, Ti = T[i], R = tq[i];
    const float gcs = g0 * rw[i * 9] + g1 * rw[i * 9 + 1] + g2 * rw[i * 9 + 2];
    const float gct = g0 * rw[i * 9 + 4] + g1 * rw[i * 9 + 5] + g2 * rw[i * 9 + 6];
    go[i * 9 + 0] = g0 * as * Ti, go[i * 9 + 1] = g1 * as * Ti, go[i * 9 + 2] = g2 * as * Ti;
    go[i * 9 + 3] = delta * (es * Ti * gcs - R);
    go[i * 9 + 4] = g0 * at * Ti, go[i * 9 + 5] = g1 * at * Ti, go[i * 9 + 6] = g2 * at * Ti;
    go[i * 9 + 7] = delta * (et * Ti * gct - R);
    go[i * 9 + 8] = 0.f;  // beta does not reach rgb
  }
}

// ---------------------------------------------------------------------------------------------
// fused forward-recompute + input-gradient backward of the fine network
// ---------------------------------------------------------------------------------------------
struct BwdArgs {
  const float* rayrec;   // [n_rays,12]
  const float* z;        // [n_rays,S]
  const float* raybias;  // [n_rays, W]
  const float* raw;      // [P,9] forward outputs (for the head derivatives)
  const float* g_raw;    // [P,9]
  int S;
  int64_t P;
  int D, skip, pek, in_xyz;
  const float* blob;     // forward layout ([K][N])
  const float* blobb;    // backward layout ([N][K])
  uint32_t trunk_w[16], trunk_b[16], bw_trunk[16];
  uint32_t sigma_w, final_w, final_b, dt_w, rgb_w, t_w[3], t_b[3], tsig_w, trgb_w, tbeta_w;
  uint32_t bw_final, bw_dt, bw_dtx, bw_t[3];
  float* g_samp;         // [P,32]: d pts (3), d dirPE (27), pad
  uint32_t* mask_dump;   // debug seam: [P][D+4][8] ballot words of the forward recompute, or null
};

// out[s][n] = act(bias[n] + rb[s][n] + sum_k in[s][k] W[k][n]); optionally records the ReLU mask
// (one ballot word per (sample, 32-column group)) or applies a recorded mask to the result.
template <int NJ>
__device__ __forceinline__ void bgemm(const float* in0, int ld0, int K0, const float* in1, int ld1, int K1,
                                      const float* __restrict__ Wm, int ldw, const float* __restrict__ bias,
                                      const float* const* rbrow, bool relu, uint32_t* mask_out, const uint32_t* mask_in,
                                      int mask_ld, float* out, int ldo, int lane) {
  float acc[kBRows][NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const float b = bias ? __ldg(bias + lane + 32 * j) : 0.f;
#pragma unroll
    for (int s = 0; s < kBRows; ++s) acc[s][j] = b + (rbrow ? __ldg(rbrow[s] + lane + 32 * j) : 0.f);
  }
  for (int seg = 0; seg < 2; ++seg) {
    const float* in = seg == 0 ? in0 : in1;
    const int ld = seg == 0 ? ld0 : ld1, K = seg == 0 ? K0 : K1;
    if (!in || K == 0) continue;
    const float* w = Wm + (seg == 0 ? 0 : (size_t)K0 * ldw);
    for (int k0 = 0; k0 < K; k0 += 4) {
      float4 a[kBRows];
#pragma unroll
      for (int s = 0; s < kBRows; ++s) a[s] = *reinterpret_cast<const float4*>(in + s * ld + k0);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        float wv[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) wv[j] = __ldg(w + (size_t)(k0 + kk) * ldw + lane + 32 * j);
#pragma unroll
        for (int s = 0; s < kBRows; ++s) {
          const float av = kk == 0 ? a[s].x : kk == 1 ? a[s].y : kk == 2 ? a[s].z : a[s].w;
#pragma unroll
          for (int j = 0; j < NJ; ++j) acc[s][j] = fmaf(av, wv[j], acc[s][j]);
        }
      }
    }
  }
#pragma unroll
  for (int s = 0; s < kBRows; ++s)
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      float v = acc[s][j];
      if (mask_out) {
        const uint32_t m = __ballot_sync(0xffffffffu, v > 0.f);
        if (lane == 0) mask_out[s * mask_ld + j] = m;
      }
      if (relu) v = fmaxf(v, 0.f);
      if (mask_in) v = ((mask_in[s * mask_ld + j] >> lane) & 1u) ? v : 0.f;
      out[s * ldo + lane + 32 * j] = v;
    }
}

__device__ __forceinline__ void bgemm_n(int N, const float* in0, int ld0, int K0, const float* in1, int ld1, int K1,
                                        const float* Wm, int ldw, const float* bias, const float* const* rbrow, bool relu,
                                        uint32_t* mask_out, const uint32_t* mask_in, int mask_ld, float* out, int ldo, int lane) {
  switch (N / 32) {
    case 1: bgemm<1>(in0, ld0, K0, in1, ld1, K1, Wm, ldw, bias, rbrow, relu, mask_out, mask_in, mask_ld, out, ldo, lane); break;
    case 2: bgemm<2>(in0, ld0, K0, in1, ld1, K1, Wm, ldw, bias, rbrow, relu, mask_out, mask_in, mask_ld, out, ldo, lane); break;
    case 3: bgemm<3>(in0, ld0, K0, in1, ld1, K1, Wm, ldw, bias, rbrow, relu, mask_out, mask_in, mask_ld, out, ldo, lane); break;
    case 4: bgemm<4>(in0, ld0, K0, in1, ld1, K1, Wm, ldw, bias, rbrow, relu, mask_out, mask_in, mask_ld, out, ldo, lane); break;
    case 6: bgemm<6>(in0, ld0, K0, in1, ld1, K1, Wm, ldw, bias, rbrow, relu, mask_out, mask_in, mask_ld, out, ldo, lane); break;
    case 8: bgemm<8>(in0, ld0, K0, in1, ld1, K1, Wm, ldw, bias, rbrow, relu, mask_out, mask_in, mask_ld, out, ldo, lane); break;
    default: break;
  }
}

template <int W>
__global__ void __launch_bounds__(kBThreads) k_mlp_simt_bwd(BwdArgs a) {
  extern __shared__ __align__(16) float sm[];
  constexpr int Hh = W / 2, MW = W / 32;  // mask words per (layer, sample)
  const int pek = a.pek;
  float* pe = sm;                    // [kBT][pek]   positional encoding (kept for its own backward)
  float* bufA = pe + kBT * pek;      // [kBT][W]
  float* bufB = bufA + kBT * W;      // [kBT][W]
  float* gpe = bufB + kBT * W;       // [kBT][pek]   gradient w.r.t. the encoding
  uint32_t* masks = reinterpret_cast<uint32_t*>(gpe + kBT * pek);  // [D + 4][kBT][MW]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t g0 = (int64_t)blockIdx.x * kBT;
  const float* B = a.blob;
  const float* BB = a.blobb;

  for (int i = tid; i < kBT * pek; i += kBThreads) {
    const int s = i / pek, c = i % pek;
    const int64_t g = min(g0 + s, a.P - 1);
    float v = 0.f;
    if (c < a.in_xyz) {
      const int64_t ray = g / a.S;
      const float* rr = a.rayrec + ray * kRayRec;
      const int comp = c < 3 ? c : (c - 3) % 3;
      const float p = __fadd_rn(rr[comp], __fmul_rn(rr[3 + comp], a.z[g]));
      if (c < 3) v = p;
      else {
        const float xf = __fmul_rn(p, (float)(1 << ((c - 3) / 6)));
        v = ((c - 3) % 6) < 3 ? sinf(xf) : cosf(xf);
      }
    }
    pe[i] = v;
    gpe[i] = 0.f;
  }
  __syncthreads();

  const int s0 = warp * kBRows;
  float* mype = pe + s0 * pek;
  float* mygpe = gpe + s0 * pek;
  float* cur = bufA + s0 * W;
  float* nxt = bufB + s0 * W;
  auto mk = [&](int layer) { return masks + ((size_t)layer * kBT + s0) * MW; };
  const float* rbrow[kBRows];
#pragma unroll
  for (int s = 0; s < kBRows; ++s) rbrow[s] = a.raybias + (min(g0 + s0 + s, a.P - 1) / a.S) * W;

  // ---- forward recompute, keeping only the ReLU masks (layers 0..D-1 trunk, D: dir|transient0, D+1..D+3) ----
  for (int i = 0; i < a.D; ++i) {
    if (i == 0) bgemm_n(W, mype, pek, pek, nullptr, 0, 0, B + a.trunk_w[0], W, B + a.trunk_b[0], nullptr, true, mk(0), nullptr, MW, cur, W, lane);
    else {
      if (i == a.skip) bgemm_n(W, mype, pek, pek, cur, W, W, B + a.trunk_w[i], W, B + a.trunk_b[i], nullptr, true, mk(i), nullptr, MW, nxt, W, lane);
      else bgemm_n(W, cur, W, W, nullptr, 0, 0, B + a.trunk_w[i], W, B + a.trunk_b[i], nullptr, true, mk(i), nullptr, MW, nxt, W, lane);
      float* t = cur; cur = nxt; nxt = t;
    }
    __syncwarp();
  }
  bgemm_n(W, cur, W, W, nullptr, 0, 0, B + a.final_w, W, B + a.final_b, nullptr, false, nullptr, nullptr, MW, nxt, W, lane);
  __syncwarp();
  { float* t = cur; cur = nxt; nxt = t; }
  bgemm_n(W, cur, W, W, nullptr, 0, 0, B + a.dt_w, W, nullptr, rbrow, true, mk(a.D), nullptr, MW, nxt, W, lane);
  __syncwarp();
  { float* t = cur; cur = nxt; nxt = t; }  // cur = [dir_enc | transient0]
  bgemm_n(Hh, cur + Hh, W, Hh, nullptr, 0, 0, B + a.t_w[0], Hh, B + a.t_b[0], nullptr, true, mk(a.D + 1), nullptr, MW, nxt, W, lane);
  __syncwarp();
  bgemm_n(Hh, nxt, W, Hh, nullptr, 0, 0, B + a.t_w[1], Hh, B + a.t_b[1], nullptr, true, mk(a.D + 2), nullptr, MW, cur, W, lane);
  __syncwarp();
  bgemm_n(Hh, cur, W, Hh, nullptr, 0, 0, B + a.t_w[2], Hh, B + a.t_b[2], nullptr, true, mk(a.D + 3), nullptr, MW, nxt, W, lane);
  __syncwarp();

  if (a.mask_dump) {
    for (int i = lane; i < (a.D + 4) * kBRows * MW; i += 32) {
      const int layer = i / (kBRows * MW), s = (i / MW) % kBRows, j = i % MW;
      if (g0 + s0 + s < a.P) a.mask_dump[((g0 + s0 + s) * (a.D + 4) + layer) * 8 + j] = mk(layer)[s * MW + j];
    }
  }
  // ---- head derivatives from the saved forward outputs ----------------------------------------------
  // lane s (< 8) owns sample s of this warp: d pre-activation of rgb(3), sigma, t_rgb(3), t_sigma, t_beta
  float gh[9];
  {
    const int64_t g = min(g0 + s0 + (lane & 7), a.P - 1);
    const bool live = (g0 + s0 + (lane & 7)) < a.P;
#pragma unroll
    for (int c = 0; c < 9; ++c) {
      const float o = a.raw[g * 9 + c], gr = live ? a.g_raw[g * 9 + c] : 0.f;
      const bool sg = c < 3 || (c >= 4 && c < 7);               // sigmoid outputs; the others are softplus
      gh[c] = sg ? gr * o * (1.f - o) : gr * (1.f - expf(-o));  // softplus'(v) = sigmoid(v) = 1 - exp(-softplus(v))
    }
  }
  // ---- transient heads -> d T3 output (masked), into cur[:, :Hh] --------------------------------------
  {
    const uint32_t* m = mk(a.D + 3);
#pragma unroll
    for (int s = 0; s < kBRows; ++s) {
      const float t0 = __shfl_sync(0xffffffffu, gh[4], s), t1 = __shfl_sync(0xffffffffu, gh[5], s);
      const float t2 = __shfl_sync(0xffffffffu, gh[6], s), tv = __shfl_sync(0xffffffffu, gh[7], s);
      const float tb = __shfl_sync(0xffffffffu, gh[8], s);
      for (int k = lane; k < Hh; k += 32) {
        float v = t0 * __ldg(B + a.trgb_w + k) + t1 * __ldg(B + a.trgb_w + Hh + k) + t2 * __ldg(B + a.trgb_w + 2 * Hh + k) +
                  tv * __ldg(B + a.tsig_w + k) + tb * __ldg(B + a.tbeta_w + k);
        v = ((m[s * MW + (k >> 5)] >> (k & 31)) & 1u) ? v : 0.f;
        cur[s * W + k] = v;
      }
    }
  }
  __syncwarp();
  // T3 -> T2 -> T1 -> transient0: g_in = W^T g_out, masked with the producing layer's ReLU mask
  bgemm_n(Hh, cur, W, Hh, nullptr, 0, 0, BB + a.bw_t[2], Hh, nullptr, nullptr, false, nullptr, mk(a.D + 2), MW, nxt, W, lane);
  __syncwarp();
  bgemm_n(Hh, nxt, W, Hh, nullptr, 0, 0, BB + a.bw_t[1], Hh, nullptr, nullptr, false, nullptr, mk(a.D + 1), MW, cur, W, lane);
  __syncwarp();
  // g wrt transient0 (second half of the dir|transient layer) -> nxt[:, Hh:W]; its mask is the second half of mk(D)
  bgemm_n(Hh, cur, W, Hh, nullptr, 0, 0, BB + a.bw_t[0], Hh, nullptr, nullptr, false, nullptr, mk(a.D) + Hh / 32, MW, nxt + Hh, W, lane);
  // g wrt dir_enc (first half) from the rgb head -> nxt[:, 0:Hh]
  {
    const uint32_t* m = mk(a.D);
#pragma unroll
    for (int s = 0; s < kBRows; ++s) {
      const float u0 = __shfl_sync(0xffffffffu, gh[0], s), u1 = __shfl_sync(0xffffffffu, gh[1], s);
      const float u2 = __shfl_sync(0xffffffffu, gh[2], s);
      for (int k = lane; k < Hh; k += 32) {
        float v = u0 * __ldg(B + a.rgb_w + k) + u1 * __ldg(B + a.rgb_w + Hh + k) + u2 * __ldg(B + a.rgb_w + 2 * Hh + k);
        v = ((m[s * MW + (k >> 5)] >> (k & 31)) & 1u) ? v : 0.f;
        nxt[s * W + k] = v;
      }
    }
  }
  __syncwarp();
  // d dirPE = dir_encoding[:, W:W+27]^T g_dir  -> g_samp[:, 3:30] (via cur[:, 0:32] as scratch)
  bgemm_n(32, nxt, W, Hh, nullptr, 0, 0, BB + a.bw_dtx, 32, nullptr, nullptr, false, nullptr, nullptr, MW, cur, W, lane);
  __syncwarp();
#pragma unroll
  for (int s = 0; s < kBRows; ++s) {
    const int64_t g = g0 + s0 + s;
    if (g < a.P && lane >= 3 && lane < 30) a.g_samp[g * 32 + lane] = cur[s * W + lane - 3];
  }
  __syncwarp();
  // d xyz_encoding_final = [W_dir[:, :W]; W_t0[:, :W]]^T g  -> cur
  bgemm_n(W, nxt, W, W, nullptr, 0, 0, BB + a.bw_dt, W, nullptr, nullptr, false, nullptr, nullptr, MW, cur, W, lane);
  __syncwarp();
  // d h_{D-1} = W_f^T g_final + d sigma_pre * w_sigma, masked with the last trunk layer's mask -> nxt
  bgemm_n(W, cur, W, W, nullptr, 0, 0, BB + a.bw_final, W, nullptr, nullptr, false, nullptr, nullptr, MW, nxt, W, lane);
  __syncwarp();
  {
    const uint32_t* m = mk(a.D - 1);
#pragma unroll
    for (int s = 0; s < kBRows; ++s) {
      const float gv = __shfl_sync(0xffffffffu, gh[3], s);
      for (int k = lane; k < W; k += 32) {
        float v = nxt[s * W + k] + gv * __ldg(B + a.sigma_w + k);
        nxt[s * W + k] = ((m[s * MW + (k >> 5)] >> (k & 31)) & 1u) ? v : 0.f;
      }
    }
  }
  __syncwarp();
  { float* t = cur; cur = nxt; nxt = t; }  // cur = masked gradient w.r.t. trunk layer D-1 output
  // ---- trunk backward ------------------------------------------------------------------------------------
  for (int i = a.D - 1; i >= 1; --i) {
    const uint32_t* mprev = mk(i - 1);
    if (i == a.skip) {
      // input was [pe | h]: the first pek gradient columns accumulate into gpe, the rest continue down
      bgemm_n(pek, cur, W, W, nullptr, 0, 0, BB + a.bw_trunk[i], pek + W, nullptr, nullptr, false, nullptr, nullptr, MW, nxt, W, lane);
      __syncwarp();
#pragma unroll
      for (int s = 0; s < kBRows; ++s)
        for (int k = lane; k < pek; k += 32) mygpe[s * pek + k] += nxt[s * W + k];
      __syncwarp();
      bgemm_n(W, cur, W, W, nullptr, 0, 0, BB + a.bw_trunk[i] + pek, pek + W, nullptr, nullptr, false, nullptr, mprev, MW, nxt, W, lane);
    } else {
      bgemm_n(W, cur, W, W, nullptr, 0, 0, BB + a.bw_trunk[i], W, nullptr, nullptr, false, nullptr, mprev, MW, nxt, W, lane);
    }
    __syncwarp();
    float* t = cur; cur = nxt; nxt = t;
  }
  bgemm_n(pek, cur, W, W, nullptr, 0, 0, BB + a.bw_trunk[0], pek, nullptr, nullptr, false, nullptr, nullptr, MW, nxt, W, lane);
  __syncwarp();
  // ---- positional-encoding backward: d p_c = g[c] + sum_l 2^l (cos(2^l p_c) g_sin - sin(2^l p_c) g_cos) ---
  if (lane < 24) {
    const int s = lane / 3, c = lane % 3;
    const int64_t g = g0 + s0 + s;
    const float* ge = nxt + s * W;
    const float* ga = mygpe + s * pek;
    const float* pv = mype + s * pek;
    float acc = ge[c] + ga[c];
    const int L = (a.in_xyz - 3) / 6;
    for (int l = 0; l < L; ++l) {
      const int is = 3 + 6 * l + c, ic = is + 3;
      acc += (float)(1 << l) * (pv[ic] * (ge[is] + ga[is]) - pv[is] * (ge[ic] + ga[ic]));
    }
    if (g < a.P) a.g_samp[g * 32 + c] = acc;
  }
}

// per ray: d rays_o, d rays_d (through pts = o + d z) and d viewdirs (through the direction encoding)
__global__ void __launch_bounds__(128) k_ray_grad(const float* __restrict__ g_samp, const float* __restrict__ z,
                                                  const float* __restrict__ rayrec, int64_t N, int S, float* __restrict__ g_o,
                                                  float* __restrict__ g_d, float* __restrict__ g_vd) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * 4 + warp;
  if (ray >= N) return;
  float so[3] = {0, 0, 0}, sd[3] = {0, 0, 0}, dpe = 0.f;  // lane j (3..29) accumulates d dirPE[j-3]
  for (int i = 0; i < S; ++i) {
    const float* g = g_samp + (ray * S + i) * 32;
    if (lane >= 3 && lane < 30) dpe += g[lane];
  }
  for (int i = lane; i < S; i += 32) {
    const float* g = g_samp + (ray * S + i) * 32;
    const float zi = z[ray * S + i];
#pragma unroll
    for (int c = 0; c < 3; ++c) so[c] += g[c], sd[c] += zi * g[c];
  }
#pragma unroll
  for (int c = 0; c < 3; ++c)
    for (int o = 16; o > 0; o >>= 1) so[c] += __shfl_xor_sync(0xffffffffu, so[c], o), sd[c] += __shfl_xor_sync(0xffffffffu, sd[c], o);
  // direction encoding backward: entry j of dirPE: j<3 identity, else band l=(j-3)/6, sin for (j-3)%6<3
  const float* rr = rayrec + ray * kRayRec;
  float contrib = 0.f;
  int comp = -1;
  if (lane >= 3 && lane < 30) {
    const int j = lane - 3;
    if (j < 3) comp = j, contrib = dpe;
    else {
      const int l = (j - 3) / 6, r = (j - 3) % 6;
      comp = r % 3;
      const float x = __fmul_rn(rr[8 + comp], (float)(1 << l));
      contrib = (float)(1 << l) * (r < 3 ? cosf(x) : -sinf(x)) * dpe;
    }
  }
  float gv[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v = comp == c ? contrib : 0.f;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    gv[c] = v;
  }
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < 3; ++c) g_o[ray * 3 + c] = so[c], g_d[ray * 3 + c] = sd[c], g_vd[ray * 3 + c] = gv[c];
  }
}

uint32_t* g_dbg_simt_mask_dump = nullptr;

template <int W>
static int launch_bwd_w(const BwdArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)kBT * (2 * a.pek + 2 * W) * sizeof(float) + (size_t)(a.D + 4) * kBT * (W / 32) * sizeof(uint32_t);
  DFB_CHECK_CUDA(cudaFuncSetAttribute(k_mlp_simt_bwd<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t blocks = (a.P + kBT - 1) / kBT;
  DFB_REQUIRE(blocks < (1ll << 31), DFB_ERR_INVALID, "too many samples in one launch");
  k_mlp_simt_bwd<W><<<(unsigned)blocks, kBThreads, smem, st>>>(a);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

int launch_render_bwd(const DfbNerf* nerf, const float* rayrec, const float* z, const float* raybias, const float* raw,
                      const float* g_rgb, int64_t n_rays, int S, float* g_raw, float* g_samp, float* g_o, float* g_d,
                      float* g_vd, int kind, cudaStream_t st, const uint32_t* saved_masks) {
  const NetPack& np = nerf->net[1];
  DFB_REQUIRE(np.loaded && np.fine && np.blob32b, DFB_ERR_INVALID, "fine network not loaded");
  DFB_REQUIRE((size_t)4 * 3 * S * sizeof(float) <= 48 * 1024, DFB_ERR_UNSUPPORTED, "too many samples per ray for the backward");
  k_composite_bwd<<<(unsigned)((n_rays + 3) / 4), 128, (size_t)4 * 3 * S * sizeof(float), st>>>(raw, z, g_rgb, n_rays, S, g_raw);
  DFB_LAUNCH_CHECK();
  if (kind != DFB_MMA_FP32_SIMT && tc_bwd_supported(nerf)) {
    // 8x256 network: forward recompute + input-gradient chain on tcgen05 (mlp_tc_bwd.cu)
    int rc = launch_mlp_tc_bwd(nerf, kind, rayrec, z, raybias, raw, g_raw, n_rays, S, g_samp, st, saved_masks);
    if (rc) return rc;
    k_ray_grad<<<(unsigned)((n_rays + 3) / 4), 128, 0, st>>>(g_samp, z, rayrec, n_rays, S, g_o, g_d, g_vd);
    DFB_LAUNCH_CHECK();
    return DFB_OK;
  }
  DFB_REQUIRE(!saved_masks, DFB_ERR_UNSUPPORTED, "saved ReLU masks belong to the tcgen05 path");
  BwdArgs a = {};
  a.rayrec = rayrec, a.z = z, a.raybias = raybias, a.raw = raw, a.g_raw = g_raw, a.S = S, a.P = n_rays * S;
  a.D = np.D, a.skip = np.skip, a.pek = np.pek, a.in_xyz = np.in_xyz, a.blob = np.blob32, a.blobb = np.blob32b;
  for (int i = 0; i < np.D; ++i)
    a.trunk_w[i] = (uint32_t)np.trunk_w[i], a.trunk_b[i] = (uint32_t)np.trunk_b[i], a.bw_trunk[i] = (uint32_t)np.bw_trunk[i];
  a.sigma_w = np.sigma_w, a.final_w = np.final_w, a.final_b = np.final_b, a.dt_w = np.dt_w, a.rgb_w = np.rgb_w;
  for (int i = 0; i < 3; ++i) a.t_w[i] = np.t_w[i], a.t_b[i] = np.t_b[i], a.bw_t[i] = (uint32_t)np.bw_t[i];
  a.tsig_w = np.tsig_w, a.trgb_w = np.trgb_w, a.tbeta_w = np.tbeta_w;
  a.bw_final = np.bw_final, a.bw_dt = np.bw_dt, a.bw_dtx = np.bw_dtx;
  a.g_samp = g_samp;
  a.mask_dump = g_dbg_simt_mask_dump;
  int rc;
  switch (np.W) {
    case 64: rc = launch_bwd_w<64>(a, st); break;
    case 128: rc = launch_bwd_w<128>(a, st); break;
    case 192: rc = launch_bwd_w<192>(a, st); break;
    case 256: rc = launch_bwd_w<256>(a, st); break;
    default: set_error("netwidth %d unsupported", np.W); return DFB_ERR_UNSUPPORTED;
  }
  if (rc) return rc;
  k_ray_grad<<<(unsigned)((n_rays + 3) / 4), 128, 0, st>>>(g_samp, z, rayrec, n_rays, S, g_o, g_d, g_vd);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

}  // namespace dfb
