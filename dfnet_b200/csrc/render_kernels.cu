// Ray generation, ray-constant MLP inputs, volumetric compositing and hierarchical
// sampling kernels (everything on the render path except the 256-wide contractions).
//
// Reference arithmetic (paths relative to /root/reference/script):
//   k_prep_rays      models/ray_utils.py:5-15 (get_rays), models/rendering.py:366-389 (ray record),
//                    :269-287 (z_vals, stratified jitter), models/nerfw.py:69-78 (hist -> embedding rows)
//   k_raybias        the ray-constant columns of dir_encoding / transient_encoding.0
//                    (models/nerfw.py:337-345: cat([xyz_encoding_final, input_dir_a]) etc.)
//   k_composite      models/rendering.py:132-243 (raw2outputs_NeRFW)
//   k_sample_pdf     models/rendering.py:24-65 (sample_pdf) and :300-304 (z_mid, sort(cat))
// Index parity: the float32 sum uses ATen's vector order, the cdf / transmittance scans run
// sequentially in float64 like ATen-CPU, products and sums are rounded separately (no FMA).
#include "common.cuh"

namespace dfb {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// --------------------------------------------------------------------------------------
// k_prep_rays
// --------------------------------------------------------------------------------------

constexpr int kPrepThreads = 128;

__global__ void __launch_bounds__(kPrepThreads) k_prep_rays(PrepArgs a) {
  extern __shared__ float sm[];
  float* rec = sm;                                  // [128][12]
  int* hidx = (int*)(sm + kPrepThreads * kRayRec);  // [128][hb]
  const int tid = threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.x * kPrepThreads;
  const int64_t r = r0 + tid;
  const int nloc = (int)min((int64_t)kPrepThreads, a.N - r0);
  if (r < a.N) {
    float o[3], d[3], vd[3], nr, fr;
    if (a.c2w) {
      const int64_t rg0 = a.pix0 + r;
      const int64_t hw = (int64_t)a.H * a.W;
      const int img = a.n_pose > 1 ? (int)(rg0 / hw) : 0;       // batched multi-pose render: image-major rays
      const int64_t rg = a.n_pose > 1 ? rg0 - (int64_t)img * hw : rg0;
      const float* c2w = a.c2w + (size_t)img * 3 * a.c2w_ld;
      const float* hist = a.hist + (size_t)img * a.hb;
      const int pj = (int)(rg / a.W), pi = (int)(rg % a.W);
      const float dx = __fdiv_rn(__fsub_rn((float)pi, (float)(a.W * 0.5)), a.focal);
      const float dy = -__fdiv_rn(__fsub_rn((float)pj, (float)(a.H * 0.5)), a.focal);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float* R = c2w + k * a.c2w_ld;
        d[k] = __fadd_rn(__fadd_rn(__fmul_rn(dx, R[0]), __fmul_rn(dy, R[1])), __fmul_rn(-1.0f, R[2]));
        o[k] = R[3];
      }
      const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
#pragma unroll
      for (int k = 0; k < 3; ++k) vd[k] = __fdiv_rn(d[k], nrm);
      nr = a.near, fr = a.far;
      for (int b = 0; b < a.hb; ++b) hidx[tid * a.hb + b] = min(max((int)hist[b], 0), a.n_vocab - 1);
    } else {
      const float* p = a.rays + r * (11 + a.hb);
#pragma unroll
      for (int k = 0; k < 3; ++k) o[k] = p[k], d[k] = p[3 + k], vd[k] = p[8 + k];
      nr = p[6], fr = p[7];
      for (int b = 0; b < a.hb; ++b) hidx[tid * a.hb + b] = min(max((int)p[11 + b], 0), a.n_vocab - 1);
    }
    float* q = rec + tid * kRayRec;
    q[0] = o[0], q[1] = o[1], q[2] = o[2], q[3] = d[0], q[4] = d[1], q[5] = d[2];
    q[6] = nr, q[7] = fr, q[8] = vd[0], q[9] = vd[1], q[10] = vd[2], q[11] = 0.f;
  }
  __syncthreads();
  for (int i = tid; i < nloc * kRayRec; i += kPrepThreads) a.rayrec[r0 * kRayRec + i] = rec[i];
  // z_vals (rendering.py:269-285)
  for (int i = tid; i < nloc * a.Nc; i += kPrepThreads) {
    const int rl = i / a.Nc, s = i % a.Nc;
    const float nr = rec[rl * kRayRec + 6], fr = rec[rl * kRayRec + 7];
    auto zat = [&](int k) {
      const float t = a.t_vals[k];
      if (!a.lindisp) return __fadd_rn(__fmul_rn(nr, __fsub_rn(1.f, t)), __fmul_rn(fr, t));
      return __fdiv_rn(1.f, __fadd_rn(__fmul_rn(__fdiv_rn(1.f, nr), __fsub_rn(1.f, t)), __fmul_rn(__fdiv_rn(1.f, fr), t)));
    };
    float zv = zat(s);
    if (a.t_rand) {
      const float lower = s == 0 ? zv : __fmul_rn(0.5f, __fadd_rn(zv, zat(s - 1)));
      const float upper = s == a.Nc - 1 ? zv : __fmul_rn(0.5f, __fadd_rn(zat(s + 1), zv));
      zv = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), a.t_rand[(r0 + rl) * a.Nc + s]));
    }
    a.z[r0 * a.Nc + i] = zv;
  }
  if (a.extra) {
    const int ne = a.n_extra;
    for (int i = tid; i < nloc * ne; i += kPrepThreads) {
      const int rl = i / ne, c = i % ne;
      float v;
      if (c < 27) {
        if (c < 3) v = rec[rl * kRayRec + 8 + c];
        else {
          const int l = (c - 3) / 6, rr = (c - 3) % 6;
          const float x = __fmul_rn(rec[rl * kRayRec + 8 + rr % 3], (float)(1 << l));
          v = rr < 3 ? sinf(x) : cosf(x);
        }
      } else if (c < 27 + a.a_dim) {
        const int e = c - 27;
        v = a.emb_a[hidx[rl * a.hb + e / 5] * 5 + e % 5];
      } else {
        const int e = c - 27 - a.a_dim;
        v = a.emb_t[hidx[rl * a.hb + e / 2] * 2 + e % 2];
      }
      a.extra[r0 * ne + i] = v;
    }
  }
}

// --------------------------------------------------------------------------------------
// k_raybias: rb[r][0:H] = b_dir + extra[r][0:nd] . Wdx ; rb[r][H:2H] = b_t0 + extra[r][nd:nd+nt] . Wtx
// --------------------------------------------------------------------------------------
__global__ void k_raybias(const float* __restrict__ extra, int ld, int64_t N, int nd, int nt, int Hh,
                          const float* __restrict__ dirx_w, const float* __restrict__ dirx_b,
                          const float* __restrict__ tx_w, const float* __restrict__ tx_b, float* __restrict__ rb,
                          int n_rb, int rb_ld, const float* __restrict__ add_bias, int pack_kind, int pad_h) {
  // block = n_rb threads (one output column each), 8 rays per block.  Every weight is loaded once per block and
  // applied to the 8 rays from registers; the ray-constant inputs are read from shared memory as float4 over k
  // (per (ray, column) the accumulation still runs over k in order, so results are unchanged).
  extern __shared__ float sm[];
  const int64_t r0 = (int64_t)blockIdx.x * 8;
  const int nloc = (int)min((int64_t)8, N - r0);
  const int ndp = (nd + 3) & ~3, ntp = (nt + 3) & ~3;
  float* smd = sm;             // [8][ndp] inputs of dir_encoding (zero padded)
  float* smt = sm + 8 * ndp;   // [8][ntp] inputs of transient_encoding.0
  for (int i = threadIdx.x; i < 8 * (ndp + ntp); i += blockDim.x) {
    const bool t = i >= 8 * ndp;
    const int ii = t ? i - 8 * ndp : i, wdt = t ? ntp : ndp, rl = ii / wdt, k = ii % wdt;
    sm[i] = (rl < nloc && k < (t ? nt : nd)) ? extra[(r0 + rl) * ld + (t ? nd : 0) + k] : 0.f;
  }
  __syncthreads();
  const int n = min((int)threadIdx.x, n_rb - 1);
  const bool tr = n >= Hh;
  const int col = tr ? n - Hh : n;
  const float* w = tr ? tx_w : dirx_w;
  const float* xin = tr ? smt : smd;
  const int kn = tr ? nt : nd, kp = tr ? ntp : ndp;
  const float b = tr ? tx_b[col] : dirx_b[col];
  float accs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int k = 0; k < kn; k += 4) {
    float wk[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) wk[e] = k + e < kn ? __ldg(w + (size_t)(k + e) * Hh + col) : 0.f;
#pragma unroll
    for (int rl = 0; rl < 8; ++rl) {
      const float4 x = *reinterpret_cast<const float4*>(xin + rl * kp + k);
      accs[rl] = fmaf(x.x, wk[0], accs[rl]);
      if (k + 1 < kn) accs[rl] = fmaf(x.y, wk[1], accs[rl]);
      if (k + 2 < kn) accs[rl] = fmaf(x.z, wk[2], accs[rl]);
      if (k + 3 < kn) accs[rl] = fmaf(x.w, wk[3], accs[rl]);
    }
  }
#pragma unroll
  for (int rl = 0; rl < 8; ++rl) {
    if (rl >= nloc) break;
    const float acc = accs[rl];
    float v = acc + b;
    // pad_h > 0: output layout of the zero-padded 8x256 embedding (tcgen05 path for narrower networks): the
    // transient_encoding.0 half starts at column pad_h; the columns in between stay zero (the caller clears rb)
    const int oc = pad_h > 0 && tr ? pad_h + col : n;
    if (add_bias) v += add_bias[oc];  // constant part of the consuming layer's bias (tcgen05 path, see mlp_tc.cu)
    if (pack_kind == 0) {
      if ((int)threadIdx.x < n_rb) rb[(r0 + rl) * rb_ld + oc] = v;
    } else {
      // packed 16-bit pairs {column 2j, column 2j+1} in word j of the row: the tcgen05 epilogue adds them with
      // one HFMA2.RELU per column pair, exactly like the constant biases of the hidden layers
      const float hi = __shfl_down_sync(0xffffffffu, v, 1);
      if ((threadIdx.x & 1) == 0 && (int)threadIdx.x < n_rb) {
        uint32_t w16;
        if (pack_kind == 1) { __half2 h = __floats2half2_rn(v, hi); w16 = *reinterpret_cast<uint32_t*>(&h); }
        else { __nv_bfloat162 h = __floats2bfloat162_rn(v, hi); w16 = *reinterpret_cast<uint32_t*>(&h); }
        reinterpret_cast<uint32_t*>(rb + (r0 + rl) * rb_ld)[oc >> 1] = w16;
      }
    }
  }
}

// --------------------------------------------------------------------------------------
// k_composite: raw2outputs_NeRFW, one warp per ray
// --------------------------------------------------------------------------------------

constexpr int kWarpsPerBlock = 4;

__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_composite(CompositeArgs a) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  if (ray >= a.N) return;
  const int S = a.S, C = a.C;
  float* oma = sm + (size_t)warp * 4 * S;  // 1 - alpha
  float* omas = oma + S;                   // 1 - static alpha
  float* T = omas + S;                     // transmittance
  float* Ts = T + S;                       // static-only transmittance
  const float* raw = a.raw + ray * S * C;
  const float* z = a.z + ray * S;
  const bool full = (C == 9);
  const bool static_depth = full && a.test_time;  // rendering.py:214-230
  for (int i = lane; i < S; i += 32) {
    const float delta = (i + 1 < S) ? __fsub_rn(z[i + 1], z[i]) : 1e2f;
    const float nd = -delta;
    if (full) {
      const float ss = raw[i * 9 + 3], st = raw[i * 9 + 7];
      oma[i] = __fsub_rn(1.f, __fsub_rn(1.f, expf(__fmul_rn(nd, __fadd_rn(ss, st)))));
      omas[i] = __fsub_rn(1.f, __fsub_rn(1.f, expf(__fmul_rn(nd, ss))));
    } else {
      // relu(sigma + noise * raw_noise_std)
      const float ss = fmaxf(a.noise ? __fadd_rn(raw[i * C + (C - 1)], __fmul_rn(a.noise[ray * S + i], a.noise_std)) : raw[i * C + (C - 1)], 0.f);
      oma[i] = __fsub_rn(1.f, __fsub_rn(1.f, expf(__fmul_rn(nd, ss))));
    }
  }
  __syncwarp();
  // exclusive cumprod, float64 running product rounded per element (ATen-CPU cumprod)
  if (lane == 0) {
    double t = 1.0;
    for (int i = 0; i < S; ++i) { T[i] = (float)t; t *= (double)oma[i]; }
  } else if (lane == 1 && static_depth) {
    double t = 1.0;
    for (int i = 0; i < S; ++i) { Ts[i] = (float)t; t *= (double)omas[i]; }
  }
  __syncwarp();
  float s_acc = 0.f, s_depth = 0.f, s_beta = 0.f, s_r[3] = {0, 0, 0}, t_r[3] = {0, 0, 0};
  for (int i = lane; i < S; i += 32) {
    const float delta = (i + 1 < S) ? __fsub_rn(z[i + 1], z[i]) : 1e2f;
    const float nd = -delta;
    const float zi = z[i], Ti = T[i];
    if (full) {
      const float ss = raw[i * 9 + 3], st = raw[i * 9 + 7];
      const float al = __fsub_rn(1.f, expf(__fmul_rn(nd, __fadd_rn(ss, st))));
      const float als = __fsub_rn(1.f, expf(__fmul_rn(nd, ss)));
      const float alt = __fsub_rn(1.f, expf(__fmul_rn(nd, st)));
      const float w = __fmul_rn(al, Ti), sw = __fmul_rn(als, Ti), tw = __fmul_rn(alt, Ti);
      if (a.weights) a.weights[ray * S + i] = w;
      if (a.tsig) a.tsig[ray * S + i] = st;
      s_acc += w;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        s_r[c] += __fmul_rn(sw, raw[i * 9 + c]);
        t_r[c] += __fmul_rn(tw, raw[i * 9 + 4 + c]);
      }
      s_beta += __fmul_rn(tw, raw[i * 9 + 8]);
      s_depth += static_depth ? __fmul_rn(__fmul_rn(als, Ts[i]), zi) : __fmul_rn(w, zi);
    } else {
      const float ss = fmaxf(a.noise ? __fadd_rn(raw[i * C + (C - 1)], __fmul_rn(a.noise[ray * S + i], a.noise_std)) : raw[i * C + (C - 1)], 0.f);
      const float al = __fsub_rn(1.f, expf(__fmul_rn(nd, ss)));
      const float w = __fmul_rn(al, Ti);
      if (a.weights) a.weights[ray * S + i] = w;
      s_acc += w;
      if (C == 4) {
#pragma unroll
        for (int c = 0; c < 3; ++c) s_r[c] += __fmul_rn(w, raw[i * 4 + c]);
        s_depth += __fmul_rn(w, zi);
      }
    }
  }
  s_acc = warp_sum(s_acc);
  s_depth = warp_sum(s_depth);
  s_beta = warp_sum(s_beta);
#pragma unroll
  for (int c = 0; c < 3; ++c) s_r[c] = warp_sum(s_r[c]), t_r[c] = warp_sum(t_r[c]);
  if (lane == 0) {
    if (a.acc) a.acc[ray] = s_acc;
    if (C != 1) {
      if (a.rgb)
        for (int c = 0; c < 3; ++c) a.rgb[ray * 3 + c] = full ? __fadd_rn(s_r[c], t_r[c]) : s_r[c];
      if (a.depth) a.depth[ray] = s_depth;
      if (a.disp) a.disp[ray] = __fdiv_rn(1.f, fmaxf(1e-10f, __fdiv_rn(s_depth, s_acc)));
      if (a.beta) a.beta[ray] = full ? __fadd_rn(s_beta, a.beta_min) : 0.f;
    }
  }
}

// --------------------------------------------------------------------------------------
// k_sample_pdf: inverse-CDF sampling (+ optional merge-sort with the coarse depths)
// --------------------------------------------------------------------------------------
__device__ __forceinline__ int ceil_log2_i(int x) { return x <= 1 ? 0 : 32 - __clz(x - 1); }

// torch.sum over a contiguous float32 row in ATen-CPU's order (see oracle aten_sum_lastdim).
__device__ float aten_sum_warp(const float* x, int n, int lane) {
  const int vec_size = n >> 3, size_ilp = vec_size >> 2;
  float p = 0.f;
  if (lane < 8) {
    float acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[j][k] = 0.f;
    const int lp = max(4, ceil_log2_i(size_ilp) / 4), lstep = 1 << lp, lmask = lstep - 1;
    int i = 0;
    while (i + lstep <= size_ilp) {
      for (int j = 0; j < lstep; ++j, ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[0][k] += x[(i * 4 + k) * 8 + lane];
      for (int j = 1; j < 4; ++j) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { acc[j][k] += acc[j - 1][k]; acc[j - 1][k] = 0.f; }
        if ((i & (lmask << (j * lp))) != 0) break;
      }
    }
    for (; i < size_ilp; ++i)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[0][k] += x[(i * 4 + k) * 8 + lane];
    for (int j = 1; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[0][k] += acc[j][k];
    p = acc[0][0];
    for (int v = size_ilp * 4; v < vec_size; ++v) p += x[v * 8 + lane];
#pragma unroll
    for (int k = 1; k < 4; ++k) p += acc[0][k];
  }
  float fin = 0.f;
  for (int k = vec_size * 8; k < n; ++k) fin += x[k];
#pragma unroll
  for (int l = 0; l < 8; ++l) fin += __shfl_sync(0xffffffffu, p, l);
  return fin;
}


__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_sample_pdf(SampleArgs a) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  if (ray >= a.N) return;
  const bool modeA = a.z_c != nullptr;
  const int nb = modeA ? a.Nc - 1 : a.nb, nw = nb - 1, Nf = a.Nf;
  const int S = modeA ? a.Nc + Nf : 0;
  const int per_warp = 3 * nb + 8 + (modeA ? a.Nc + Nf : Nf);
  float* bins = sm + (size_t)warp * per_warp;
  float* pdf = bins + nb;
  float* cdf = pdf + nb;
  float* zall = cdf + nb + 8;  // [Nc coarse | Nf samples] (mode A) or [Nf]
  float* smp = modeA ? zall + a.Nc : zall;
  if (modeA) {
    const float* z = a.z_c + ray * a.Nc;
    const float* w = a.w_c + ray * a.Nc;
    for (int i = lane; i < a.Nc; i += 32) zall[i] = z[i];
    for (int i = lane; i < nb; i += 32) bins[i] = __fmul_rn(0.5f, __fadd_rn(z[i + 1], z[i]));
    for (int i = lane; i < nw; i += 32) pdf[i] = __fadd_rn(w[i + 1], 1e-5f);
  } else {
    for (int i = lane; i < nb; i += 32) bins[i] = a.bins[ray * nb + i];
    for (int i = lane; i < nw; i += 32) pdf[i] = __fadd_rn(a.weights[ray * nw + i], 1e-5f);
  }
  __syncwarp();
  const float tot = aten_sum_warp(pdf, nw, lane);
  __syncwarp();
  for (int i = lane; i < nw; i += 32) pdf[i] = __fdiv_rn(pdf[i], tot);
  __syncwarp();
  if (lane == 0) {  // float64 running sum rounded per element (ATen-CPU cumsum)
    double c = 0.0;
    cdf[0] = 0.f;
    for (int i = 0; i < nw; ++i) { c += (double)pdf[i]; cdf[i + 1] = (float)c; }
  }
  __syncwarp();
  float s1 = 0.f;
  for (int k = lane; k < Nf; k += 32) {
    const float u = a.u ? a.u[ray * Nf + k] : a.u_lin[k];
    int lo = 0, hi = nb;  // searchsorted(right=True): first index with cdf > u
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
    }
    const int below = max(0, lo - 1), above = min(nb - 1, lo);
    const float c0 = cdf[below], c1 = cdf[above];
    float denom = __fsub_rn(c1, c0);
    if (denom < 1e-5f) denom = 1.f;
    const float t = __fdiv_rn(__fsub_rn(u, c0), denom);
    const float v = __fadd_rn(bins[below], __fmul_rn(t, __fsub_rn(bins[above], bins[below])));
    smp[k] = v;
    s1 += v;
    if (a.samples) a.samples[ray * Nf + k] = v;
    if (a.inds) a.inds[ray * Nf + k] = lo;
  }
  __syncwarp();
  if (a.z_std) {  // torch.std(z_samples, -1, unbiased=False) (rendering.py:327)
    const float mean = warp_sum(s1) / (float)Nf;
    float s2 = 0.f;
    for (int k = lane; k < Nf; k += 32) { const float dlt = smp[k] - mean; s2 = fmaf(dlt, dlt, s2); }
    s2 = warp_sum(s2);
    if (lane == 0) a.z_std[ray] = sqrtf(s2 / (float)Nf);
  }
  if (modeA && a.n_live) {
    // early ray termination: the coarse transmittance after sample j is 1 - sum_{i<=j} w_i; everything behind the coarse
    // sample that follows the first j with T < eps is dropped from the fine pass (one coarse interval of margin)
    float zt = 0.f;
    if (lane == 0) {
      const float* w = a.w_c + ray * a.Nc;
      float cum = 0.f;
      int j = 0;
      for (; j < a.Nc; ++j) { cum += w[j]; if (1.f - cum < a.ert_eps) break; }
      zt = zall[min(j + 1, a.Nc - 1)];
    }
    zt = __shfl_sync(0xffffffffu, zt, 0);
    int cnt = 0;
    for (int i = lane; i < a.Nc; i += 32) cnt += zall[i] <= zt;
    for (int k = lane; k < Nf; k += 32) cnt += smp[k] <= zt;
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) a.n_live[ray] = max(cnt, 1);
  }
  if (modeA && a.z_vals) {  // torch.sort(torch.cat([z_vals, z_samples], -1), -1) values
    // Both lists are normally already sorted (coarse depths always; samples when u is the
    // deterministic grid): then the union is a merge, each element's rank is its own index plus a
    // binary search in the other list.  Otherwise (random u) fall back to an O(S^2) rank sort.
    const int Nc = a.Nc;
    bool sorted = true;
    for (int k = lane + 1; k < Nf; k += 32) sorted &= smp[k - 1] <= smp[k];
    for (int i = lane + 1; i < Nc; i += 32) sorted &= zall[i - 1] <= zall[i];
    sorted = __all_sync(0xffffffffu, sorted);
    if (sorted) {
      for (int i = lane; i < Nc; i += 32) {  // coarse element: samples strictly below it come first
        const float v = zall[i];
        int lo = 0, hi = Nf;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (smp[mid] < v) lo = mid + 1; else hi = mid; }
        a.z_vals[ray * S + i + lo] = v;
      }
      for (int k = lane; k < Nf; k += 32) {  // sample: coarse elements <= it come first
        const float v = smp[k];
        int lo = 0, hi = Nc;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (zall[mid] <= v) lo = mid + 1; else hi = mid; }
        a.z_vals[ray * S + k + lo] = v;
      }
    } else {
      for (int i = lane; i < S; i += 32) {
        const float v = zall[i];
        int rank = 0;
        for (int j = 0; j < S; ++j) {
          const float o = zall[j];
          rank += (o < v) || (o == v && j < i);
        }
        a.z_vals[ray * S + rank] = v;
      }
    }
  }
}

// --------------------------------------------------------------------------------------
// host launchers
// --------------------------------------------------------------------------------------
int launch_prep(const PrepArgs& a, cudaStream_t st) {
  const int blocks = (int)((a.N + kPrepThreads - 1) / kPrepThreads);
  const size_t smem = kPrepThreads * kRayRec * sizeof(float) + kPrepThreads * a.hb * sizeof(int);
  k_prep_rays<<<blocks, kPrepThreads, smem, st>>>(a);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

int launch_raybias(const float* extra, int ld, int64_t N, const NetPack& np, bool with_transient, float* rb,
                   int rb_ld, cudaStream_t st, const float* add_bias, int pack_kind, int pad_h) {
  const int Hh = np.W / 2, nd = np.in_dir + np.a_dim, nt = with_transient ? np.t_dim : 0;
  const int n_rb = with_transient ? 2 * Hh : Hh;
  const int blocks = (int)((N + 7) / 8);
  const int threads = round_up(n_rb, 32);
  const size_t smem = 8 * (((nd + 3) & ~3) + ((nt + 3) & ~3)) * sizeof(float);
  const float* b = np.blob32;
  if (pad_h > Hh) DFB_CHECK_CUDA(cudaMemsetAsync(rb, 0, (size_t)N * rb_ld * sizeof(float), st));  // zero-padded embedding
  k_raybias<<<blocks, threads, smem, st>>>(extra, ld, N, nd, nt, Hh, b + np.dirx_w, b + np.dirx_b,
                                            with_transient ? b + np.tx_w : nullptr,
                                            with_transient ? b + np.tx_b : nullptr, rb, n_rb, rb_ld, add_bias, pack_kind,
                                            pad_h > Hh ? pad_h : 0);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

// Test-time fine compositing (C == 9, rgb / disp / acc only: the render_path step, rendering.py:196-242), one warp per
// ray.  The ray's raw row (S x 9 floats) is staged in shared memory with coalesced loads and read once; the two
// exclusive transmittance products (all densities / static only) are float64 warp scans over contiguous per-lane
// segments instead of a serial loop on one lane.  The association of the products therefore differs from ATen's
// serial cumprod by float64 rounding; nothing downstream of this kernel decides a sample index (the coarse pass,
// which does, stays on k_composite), and rgb / disp / acc are gated at 1e-4 relative.
__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_composite_fine_tt(CompositeArgs a) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  if (ray >= a.N) return;
  const int S = a.S;
  float* rw = sm + (size_t)warp * (10 * S + 4);  // raw [S][9]
  float* zz = rw + 9 * S;                        // z [S]
  {
    const float* src = a.raw + ray * S * 9;
    if ((S & 3) == 0) {  // 16-byte copies: the row (S x 36 B) and the shared-memory slot are 16-byte aligned
      const float4* s4 = reinterpret_cast<const float4*>(src);
      float4* d4 = reinterpret_cast<float4*>(rw);
      for (int i = lane; i < 9 * S / 4; i += 32) d4[i] = __ldcs(s4 + i);
    } else {
      for (int i = lane; i < 9 * S; i += 32) rw[i] = __ldcs(src + i);
    }
    const float* zs = a.z + ray * S;
    for (int i = lane; i < S; i += 32) zz[i] = zs[i];
  }
  __syncwarp();
  const int per = (S + 31) / 32, i0 = min(lane * per, S), i1 = min(i0 + per, S);
  // pass 1: this lane's segment products of (1 - alpha) and (1 - alpha_static)
  double p_all = 1.0, p_st = 1.0;
  for (int i = i0; i < i1; ++i) {
    const float delta = (i + 1 < S) ? __fsub_rn(zz[i + 1], zz[i]) : 1e2f;
    const float ss = rw[i * 9 + 3], st = rw[i * 9 + 7];
    p_all *= (double)__fsub_rn(1.f, __fsub_rn(1.f, expf(__fmul_rn(-delta, __fadd_rn(ss, st)))));
    p_st *= (double)__fsub_rn(1.f, __fsub_rn(1.f, expf(__fmul_rn(-delta, ss))));
  }
  // exclusive scan over lanes
  double e_all = p_all, e_st = p_st;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double u = __shfl_up_sync(0xffffffffu, e_all, o), v = __shfl_up_sync(0xffffffffu, e_st, o);
    if (lane >= o) e_all *= u, e_st *= v;
  }
  e_all = __shfl_up_sync(0xffffffffu, e_all, 1), e_st = __shfl_up_sync(0xffffffffu, e_st, 1);
  if (lane == 0) e_all = 1.0, e_st = 1.0;
  // pass 2: weights and sums
  float s_acc = 0.f, s_depth = 0.f, s_r[3] = {0, 0, 0}, t_r[3] = {0, 0, 0};
  double t_all = e_all, t_st = e_st;
  for (int i = i0; i < i1; ++i) {
    const float delta = (i + 1 < S) ? __fsub_rn(zz[i + 1], zz[i]) : 1e2f;
    const float ss = rw[i * 9 + 3], st = rw[i * 9 + 7];
    const float ea = expf(__fmul_rn(-delta, __fadd_rn(ss, st))), es = expf(__fmul_rn(-delta, ss));
    const float al = __fsub_rn(1.f, ea), als = __fsub_rn(1.f, es), alt = __fsub_rn(1.f, expf(__fmul_rn(-delta, st)));
    const float Ti = (float)t_all, Tsi = (float)t_st;
    const float w = __fmul_rn(al, Ti), sw = __fmul_rn(als, Ti), tw = __fmul_rn(alt, Ti);
    s_acc += w;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      s_r[c] += __fmul_rn(sw, rw[i * 9 + c]);
      t_r[c] += __fmul_rn(tw, rw[i * 9 + 4 + c]);
    }
    s_depth += __fmul_rn(__fmul_rn(als, Tsi), zz[i]);
    t_all *= (double)__fsub_rn(1.f, al), t_st *= (double)__fsub_rn(1.f, als);
  }
  s_acc = warp_sum(s_acc);
  s_depth = warp_sum(s_depth);
#pragma unroll
  for (int c = 0; c < 3; ++c) s_r[c] = warp_sum(s_r[c]), t_r[c] = warp_sum(t_r[c]);
  if (lane == 0) {
    if (a.acc) a.acc[ray] = s_acc;
    if (a.rgb)
      for (int c = 0; c < 3; ++c) a.rgb[ray * 3 + c] = __fadd_rn(s_r[c], t_r[c]);
    if (a.disp) a.disp[ray] = __fdiv_rn(1.f, fmaxf(1e-10f, __fdiv_rn(s_depth, s_acc)));
  }
}

// Early ray termination, compaction: exclusive scan of n_live over the rays of a chunk (<= 65 536, one block) and the
// row map of the compacted sample list.
__global__ void __launch_bounds__(1024) k_ert_scan(const int* __restrict__ n_live, int n, int* __restrict__ offsets) {
  __shared__ int part[1024];
  const int per = (n + 1023) / 1024, i0 = min(threadIdx.x * per, n), i1 = min(i0 + per, n);
  int s = 0;
  for (int i = i0; i < i1; ++i) s += n_live[i];
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const int v = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  int run = threadIdx.x ? part[threadIdx.x - 1] : 0;
  for (int i = i0; i < i1; ++i) { offsets[i] = run; run += n_live[i]; }
  if (threadIdx.x == 1023) offsets[n] = part[1023];
}

__global__ void __launch_bounds__(128) k_ert_rowmap(const int* __restrict__ offsets, int64_t N, int S, int* __restrict__ rowmap) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = (int64_t)blockIdx.x * 4 + warp;
  if (ray >= N) return;
  const int o0 = offsets[ray], n = offsets[ray + 1] - o0;
  for (int i = lane; i < n; i += 32) rowmap[o0 + i] = (int)(ray * S + i);
}

int launch_ert_compact(const int* n_live, int64_t n_rays, int S, int* offsets, int* rowmap, cudaStream_t st) {
  DFB_REQUIRE(n_rays <= (1 << 20), DFB_ERR_INVALID, "ERT scan: too many rays in one chunk");
  k_ert_scan<<<1, 1024, 0, st>>>(n_live, (int)n_rays, offsets);
  DFB_LAUNCH_CHECK();
  k_ert_rowmap<<<(unsigned)((n_rays + 3) / 4), 128, 0, st>>>(offsets, n_rays, S, rowmap);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

// Chains the per-warp segment records of the fused fine pass (mlp_tc.cu, fused_composite) front to back:
// rgb += T * C_k, acc += T * A_k, depth += T_static * D_k, T *= prod_k (rendering.py:196-242, test_time).
__global__ void __launch_bounds__(128) k_composite_partials(const float* __restrict__ part, int part_k, int64_t N, int S,
                                                             float* __restrict__ rgb, float* __restrict__ disp,
                                                             float* __restrict__ acc, const int* __restrict__ offsets) {
  const int64_t ray = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (ray >= N) return;
  // rows of the ray in the tile grid: [ray*S, ray*S + S), or its live rows in the compacted list (ERT)
  const int64_t g0 = offsets ? offsets[ray] : ray * S, g1 = offsets ? offsets[ray + 1] : g0 + S;
  const int K = (int)(((g1 - 1) >> 5) - (g0 >> 5)) + 1;
  const float4* p = reinterpret_cast<const float4*>(part + (size_t)ray * part_k * 8);
  float T = 1.f, Ts = 1.f, r0 = 0.f, r1 = 0.f, r2 = 0.f, ac = 0.f, dep = 0.f;
  for (int k = 0; k < K; ++k) {
    const float4 u = __ldcs(p + 2 * k), v = __ldcs(p + 2 * k + 1);
    r0 = fmaf(T, u.z, r0), r1 = fmaf(T, u.w, r1), r2 = fmaf(T, v.x, r2);
    ac = fmaf(T, v.y, ac);
    dep = fmaf(Ts, v.z, dep);
    T *= u.x, Ts *= u.y;
  }
  rgb[ray * 3 + 0] = r0, rgb[ray * 3 + 1] = r1, rgb[ray * 3 + 2] = r2;
  acc[ray] = ac;
  disp[ray] = __fdiv_rn(1.f, fmaxf(1e-10f, __fdiv_rn(dep, ac)));
}

int launch_composite_partials(const float* part, int part_k, int64_t n_rays, int S, float* rgb, float* disp, float* acc,
                              cudaStream_t st, const int* ert_offsets) {
  k_composite_partials<<<(unsigned)((n_rays + 127) / 128), 128, 0, st>>>(part, part_k, n_rays, S, rgb, disp, acc, ert_offsets);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

int launch_composite(const CompositeArgs& a, cudaStream_t st) {
  // render_path step: fine pass at test time with only rgb / disp / acc wanted
  if (a.C == 9 && a.typ_fine && a.test_time && !a.weights && !a.tsig && !a.beta && !a.depth) {
    const size_t smem_f = (size_t)kWarpsPerBlock * (10 * a.S + 4) * sizeof(float);
    if (smem_f <= 48 * 1024) {
      k_composite_fine_tt<<<(unsigned)((a.N + kWarpsPerBlock - 1) / kWarpsPerBlock), kWarpsPerBlock * 32, smem_f, st>>>(a);
      DFB_LAUNCH_CHECK();
      return DFB_OK;
    }
  }
  const int blocks = (int)((a.N + kWarpsPerBlock - 1) / kWarpsPerBlock);
  const size_t smem = (size_t)kWarpsPerBlock * 4 * a.S * sizeof(float);
  DFB_REQUIRE(smem <= 48 * 1024, DFB_ERR_UNSUPPORTED, "samples per ray %d too large for the compositing kernel", a.S);
  k_composite<<<blocks, kWarpsPerBlock * 32, smem, st>>>(a);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

int launch_sample(const SampleArgs& a, cudaStream_t st) {
  const bool modeA = a.z_c != nullptr;
  const int nb = modeA ? a.Nc - 1 : a.nb;
  DFB_REQUIRE(nb >= 2 && a.Nf >= 1, DFB_ERR_INVALID, "sample_pdf needs >= 2 bins and >= 1 sample");
  const int per_warp = 3 * nb + 8 + (modeA ? a.Nc + a.Nf : a.Nf);
  const size_t smem = (size_t)kWarpsPerBlock * per_warp * sizeof(float);
  DFB_REQUIRE(smem <= 48 * 1024, DFB_ERR_UNSUPPORTED, "N_samples/N_importance too large for the sampling kernel");
  const int blocks = (int)((a.N + kWarpsPerBlock - 1) / kWarpsPerBlock);
  k_sample_pdf<<<blocks, kWarpsPerBlock * 32, smem, st>>>(a);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

}  // namespace dfb
