// fp32 SIMT NeRF-W MLP (any width in {64,128,192,256}); positional encoding computed
// in-kernel from the ray record and sample depth, all layers fused, activations in shared
// memory.  This is the exact-order (no tensor-core rounding) variant used for odd network
// sizes, for the NeRFW.forward seam and as the fp32 end of the precision ladder; the
// 256-wide production path is mlp_tc.cu.
//
// Reference: models/nerfw.py:105-133 (Embedder.embed), :297-354 (NeRFW.forward),
//            :15-95 (run_network_NeRFW), models/rendering.py:287,305 (pts = o + d*z).
#include "common.cuh"

namespace dfb {

constexpr int kTile = 64;      // samples per CTA
constexpr int kThreads = 256;  // 8 warps x 8 samples
constexpr int kRows = 8;       // samples per warp

struct SimtArgs {
  // geometry source A: rays + depths
  const float* rayrec;   // [n_rays,12]
  const float* z;        // [n_rays,S]
  int S;
  // geometry source B: embedded inputs
  const float* x;        // [P, ldx]; first in_xyz columns are the xyz encoding
  int ldx;
  int64_t P;             // total samples
  const float* raybias;  // [n_rays or P, n_rb] (null for sigma-only)
  int n_rb;
  int mode;              // MlpMode
  int D, skip, pek, in_xyz;
  const float* blob;
  uint32_t trunk_w[16], trunk_b[16];
  uint32_t sigma_w, sigma_b, final_w, final_b, dt_w, rgb_w, rgb_b;
  uint32_t t_w[3], t_b[3], tsig_w, tsig_b, trgb_w, trgb_b, tbeta_w, tbeta_b;
  float* raw;            // [P, C]
};

__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

// out[s][n] = act(bias[n] + rb[s][n] + sum_k in[s][k] * Wt[k][n]) for this warp's 8 samples.
template <int NJ>
__device__ __forceinline__ void gemm_step(const float* in0, int ld0, int K0, const float* in1, int ld1, int K1,
                                          const float* __restrict__ Wt, int ldw, const float* __restrict__ bias,
                                          const float* const* rbrow, bool relu, float* out, int ldo, int lane) {
  float acc[kRows][NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const float b = bias ? __ldg(bias + lane + 32 * j) : 0.f;
#pragma unroll
    for (int s = 0; s < kRows; ++s) acc[s][j] = b + (rbrow ? __ldg(rbrow[s] + lane + 32 * j) : 0.f);
  }
  for (int seg = 0; seg < 2; ++seg) {
    const float* in = seg == 0 ? in0 : in1;
    const int ld = seg == 0 ? ld0 : ld1, K = seg == 0 ? K0 : K1;
    if (!in || K == 0) continue;
    const float* w = Wt + (seg == 0 ? 0 : (size_t)K0 * ldw);
    for (int k0 = 0; k0 < K; k0 += 4) {
      float4 a[kRows];
#pragma unroll
      for (int s = 0; s < kRows; ++s) a[s] = *reinterpret_cast<const float4*>(in + s * ld + k0);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        float wv[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) wv[j] = __ldg(w + (size_t)(k0 + kk) * ldw + lane + 32 * j);
#pragma unroll
        for (int s = 0; s < kRows; ++s) {
          const float av = kk == 0 ? a[s].x : kk == 1 ? a[s].y : kk == 2 ? a[s].z : a[s].w;
#pragma unroll
          for (int j = 0; j < NJ; ++j) acc[s][j] = fmaf(av, wv[j], acc[s][j]);
        }
      }
    }
  }
#pragma unroll
  for (int s = 0; s < kRows; ++s)
#pragma unroll
    for (int j = 0; j < NJ; ++j) out[s * ldo + lane + 32 * j] = relu ? fmaxf(acc[s][j], 0.f) : acc[s][j];
}

__device__ __forceinline__ void gemm_dispatch(int N, const float* in0, int ld0, int K0, const float* in1, int ld1,
                                              int K1, const float* Wt, int ldw, const float* bias,
                                              const float* const* rbrow, bool relu, float* out, int ldo, int lane) {
  switch (N / 32) {
    case 1: gemm_step<1>(in0, ld0, K0, in1, ld1, K1, Wt, ldw, bias, rbrow, relu, out, ldo, lane); break;
    case 2: gemm_step<2>(in0, ld0, K0, in1, ld1, K1, Wt, ldw, bias, rbrow, relu, out, ldo, lane); break;
    case 3: gemm_step<3>(in0, ld0, K0, in1, ld1, K1, Wt, ldw, bias, rbrow, relu, out, ldo, lane); break;
    case 4: gemm_step<4>(in0, ld0, K0, in1, ld1, K1, Wt, ldw, bias, rbrow, relu, out, ldo, lane); break;
    case 6: gemm_step<6>(in0, ld0, K0, in1, ld1, K1, Wt, ldw, bias, rbrow, relu, out, ldo, lane); break;
    case 8: gemm_step<8>(in0, ld0, K0, in1, ld1, K1, Wt, ldw, bias, rbrow, relu, out, ldo, lane); break;
    default: break;
  }
}

// dot(h[s][0:K], w) for this warp's 8 samples; result valid in all lanes.
// (returned value is sample `lane`'s dot product for lane < 8)
__device__ __forceinline__ float head_dot(const float* h, int ld, int K, const float* __restrict__ w, int lane) {
  float p[kRows];
#pragma unroll
  for (int s = 0; s < kRows; ++s) p[s] = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float wv = __ldg(w + k);
#pragma unroll
    for (int s = 0; s < kRows; ++s) p[s] = fmaf(h[s * ld + k], wv, p[s]);
  }
  float mine = 0.f;
#pragma unroll
  for (int s = 0; s < kRows; ++s) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) p[s] += __shfl_xor_sync(0xffffffffu, p[s], o);
    if (lane == s) mine = p[s];
  }
  return mine;
}

template <int W>
__global__ void __launch_bounds__(kThreads) k_mlp_simt(SimtArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int pek = a.pek;
  float* pe = sm;                    // [kTile][pek]
  float* bufA = pe + kTile * pek;    // [kTile][W]
  float* bufB = bufA + kTile * W;    // [kTile][W]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t g0 = (int64_t)blockIdx.x * kTile;
  const float* B = a.blob;
  constexpr int Hh = W / 2;

  // ---- positional encoding of this tile (nerfw.py:128-133) ------------------------------
  for (int i = tid; i < kTile * pek; i += kThreads) {
    const int s = i / pek, c = i % pek;
    const int64_t g = g0 + s;
    float v = 0.f;
    if (g < a.P && c < a.in_xyz) {
      if (a.x) {
        v = a.x[g * a.ldx + c];
      } else {
        const int64_t ray = g / a.S;
        const float* rr = a.rayrec + ray * kRayRec;
        const float zz = a.z[g];
        const int comp = c < 3 ? c : (c - 3) % 3;
        const float p = __fadd_rn(rr[comp], __fmul_rn(rr[3 + comp], zz));  // pts = o + d*z
        if (c < 3) v = p;
        else {
          const int l = (c - 3) / 6;
          const float xf = __fmul_rn(p, (float)(1 << l));
          v = ((c - 3) % 6) < 3 ? sinf(xf) : cosf(xf);
        }
      }
    }
    pe[i] = v;
  }
  __syncthreads();

  const int s0 = warp * kRows;  // this warp's rows; rows are warp-private from here on
  float* mype = pe + s0 * pek;
  float* cur = bufA + s0 * W;
  float* nxt = bufB + s0 * W;
  const float* rbrow[kRows];
#pragma unroll
  for (int s = 0; s < kRows; ++s) {
    int64_t g = min(g0 + s0 + s, a.P - 1);
    const int64_t row = a.x ? g : g / a.S;
    rbrow[s] = a.raybias ? a.raybias + row * a.n_rb : nullptr;
  }

  // ---- trunk (nerfw.py:326-330) -----------------------------------------------------------
  for (int i = 0; i < a.D; ++i) {
    if (i == 0) gemm_dispatch(W, mype, pek, pek, nullptr, 0, 0, B + a.trunk_w[0], W, B + a.trunk_b[0], nullptr, true, cur, W, lane);
    else {
      if (i == a.skip) gemm_dispatch(W, mype, pek, pek, cur, W, W, B + a.trunk_w[i], W, B + a.trunk_b[i], nullptr, true, nxt, W, lane);
      else gemm_dispatch(W, cur, W, W, nullptr, 0, 0, B + a.trunk_w[i], W, B + a.trunk_b[i], nullptr, true, nxt, W, lane);
      float* t = cur; cur = nxt; nxt = t;
    }
    __syncwarp();
  }
  const float sig = softplus_f(head_dot(cur, W, W, B + a.sigma_w, lane) + __ldg(B + a.sigma_b));
  const int C = a.mode == MLP_SIGMA ? 1 : (a.mode == MLP_STATIC ? 4 : 9);
  const int64_t gmine = g0 + s0 + lane;  // lanes 0..7 own one sample each for the heads
  const bool wr = lane < kRows && gmine < a.P;
  if (a.mode == MLP_SIGMA) {
    if (wr) a.raw[gmine] = sig;
    return;
  }
  // ---- xyz_encoding_final, dir_encoding (+ transient_encoding.0) (nerfw.py:336-345) ---------
  gemm_dispatch(W, cur, W, W, nullptr, 0, 0, B + a.final_w, W, B + a.final_b, nullptr, false, nxt, W, lane);
  __syncwarp();
  { float* t = cur; cur = nxt; nxt = t; }
  const int ndt = a.mode == MLP_FULL ? W : Hh;
  gemm_dispatch(ndt, cur, W, W, nullptr, 0, 0, B + a.dt_w, a.n_rb, nullptr, rbrow, true, nxt, W, lane);
  __syncwarp();
  { float* t = cur; cur = nxt; nxt = t; }  // cur = [dir_enc (Hh) | transient0 (Hh)]
  float out[9];
#pragma unroll
  for (int c = 0; c < 3; ++c)
    out[c] = sigmoid_f(head_dot(cur, W, Hh, B + a.rgb_w + c * Hh, lane) + __ldg(B + a.rgb_b + c));
  out[3] = sig;
  if (a.mode == MLP_FULL) {
    // ---- transient branch (nerfw.py:345-354) ------------------------------------------------
    gemm_dispatch(Hh, cur + Hh, W, Hh, nullptr, 0, 0, B + a.t_w[0], Hh, B + a.t_b[0], nullptr, true, nxt, W, lane);
    __syncwarp();
    gemm_dispatch(Hh, nxt, W, Hh, nullptr, 0, 0, B + a.t_w[1], Hh, B + a.t_b[1], nullptr, true, cur, W, lane);
    __syncwarp();
    gemm_dispatch(Hh, cur, W, Hh, nullptr, 0, 0, B + a.t_w[2], Hh, B + a.t_b[2], nullptr, true, nxt, W, lane);
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 3; ++c)
      out[4 + c] = sigmoid_f(head_dot(nxt, W, Hh, B + a.trgb_w + c * Hh, lane) + __ldg(B + a.trgb_b + c));
    out[7] = softplus_f(head_dot(nxt, W, Hh, B + a.tsig_w, lane) + __ldg(B + a.tsig_b));
    out[8] = softplus_f(head_dot(nxt, W, Hh, B + a.tbeta_w, lane) + __ldg(B + a.tbeta_b));
  }
  if (wr) {
#pragma unroll
    for (int c = 0; c < 9; ++c)
      if (c < C) a.raw[gmine * C + c] = out[c];
  }
}

static int fill_args(const DfbNerf* nerf, int which, int mode, SimtArgs& a) {
  const NetPack& np = nerf->net[which];
  DFB_REQUIRE(np.loaded, DFB_ERR_INVALID, "network %d has no parameters loaded", which);
  DFB_REQUIRE(mode != MLP_FULL || np.fine, DFB_ERR_INVALID, "full (transient) output needs the fine network");
  a.mode = mode, a.D = np.D, a.skip = np.skip, a.pek = np.pek, a.in_xyz = np.in_xyz;
  a.blob = np.blob32;
  for (int i = 0; i < np.D; ++i) a.trunk_w[i] = (uint32_t)np.trunk_w[i], a.trunk_b[i] = (uint32_t)np.trunk_b[i];
  a.sigma_w = np.sigma_w, a.sigma_b = np.sigma_b, a.final_w = np.final_w, a.final_b = np.final_b;
  a.dt_w = np.dt_w, a.rgb_w = np.rgb_w, a.rgb_b = np.rgb_b;
  for (int i = 0; i < 3; ++i) a.t_w[i] = np.t_w[i], a.t_b[i] = np.t_b[i];
  a.tsig_w = np.tsig_w, a.tsig_b = np.tsig_b, a.trgb_w = np.trgb_w, a.trgb_b = np.trgb_b;
  a.tbeta_w = np.tbeta_w, a.tbeta_b = np.tbeta_b;
  a.n_rb = np.n_dt;
  return DFB_OK;
}

template <int W>
static int launch_w(const SimtArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)kTile * (a.pek + 2 * W) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    DFB_CHECK_CUDA(cudaFuncSetAttribute(k_mlp_simt<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int64_t blocks = (a.P + kTile - 1) / kTile;
  DFB_REQUIRE(blocks < (1ll << 31), DFB_ERR_INVALID, "too many samples in one launch");
  k_mlp_simt<W><<<(unsigned)blocks, kThreads, smem, st>>>(a);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

static int launch_any(int W, const SimtArgs& a, cudaStream_t st) {
  switch (W) {
    case 64: return launch_w<64>(a, st);
    case 128: return launch_w<128>(a, st);
    case 192: return launch_w<192>(a, st);
    case 256: return launch_w<256>(a, st);
  }
  set_error("netwidth %d unsupported", W);
  return DFB_ERR_UNSUPPORTED;
}

int launch_mlp_simt_rays(const DfbNerf* nerf, int which, int mode, const float* rayrec, const float* z,
                         const float* raybias, int64_t n_rays, int S, float* raw, cudaStream_t st) {
  SimtArgs a = {};
  int rc = fill_args(nerf, which, mode, a);
  if (rc) return rc;
  a.rayrec = rayrec, a.z = z, a.S = S, a.x = nullptr, a.ldx = 0, a.P = n_rays * S, a.raybias = raybias, a.raw = raw;
  DFB_REQUIRE(mode == MLP_SIGMA || raybias, DFB_ERR_INVALID, "ray-constant inputs missing");
  if (a.P == 0) return DFB_OK;
  return launch_any(nerf->net[which].W, a, st);
}

int launch_mlp_simt_embedded(const DfbNerf* nerf, int which, int mode, const float* x, int64_t P, float* out,
                             cudaStream_t st) {
  SimtArgs a = {};
  int rc = fill_args(nerf, which, mode, a);
  if (rc) return rc;
  const NetPack& np = nerf->net[which];
  const int ldx = mode == MLP_SIGMA ? np.in_xyz
                                    : np.in_xyz + np.in_dir + np.a_dim + (mode == MLP_FULL ? np.t_dim : 0);
  a.x = x, a.ldx = ldx, a.P = P, a.S = 1, a.raw = out, a.raybias = nullptr;
  if (P == 0) return DFB_OK;
  float* rb = nullptr;
  if (mode != MLP_SIGMA) {
    // per-point "ray constants": the direction/appearance/transient columns of x
    DFB_CHECK_CUDA(cudaMallocAsync(&rb, (size_t)P * np.n_dt * sizeof(float), st));
    rc = launch_raybias(x + np.in_xyz, ldx, P, np, mode == MLP_FULL, rb, np.n_dt, st);
    if (rc) { cudaFreeAsync(rb, st); return rc; }
    a.raybias = rb;
    a.n_rb = np.n_dt;
  }
  rc = launch_any(np.W, a, st);
  if (rb) cudaFreeAsync(rb, st);
  return rc;
}

}  // namespace dfb
