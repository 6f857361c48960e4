// Data-side and evaluation kernels around the hot path.
//
// Reference arithmetic (paths relative to /root/reference):
//   k_luma_hist       dataset_loaders/seven_scenes.py:346-352 (histogram of the Y channel, as percentages, rounded)
//                     with dataset_loaders/utils/color.py:29-35 (rgb_to_yuv, Y = 0.299 r + 0.587 g + 0.114 b)
//   k_resize_area     dataset_loaders/seven_scenes.py:328-332 (cv2.resize(..., interpolation=cv2.INTER_AREA) of the
//                     float image when df != 1); also feature/direct_feature_matching.py:149
//   k_pose_error      script/feature/misc.py:49-107 (compute_error_in_q: SVD-orthogonalised rotation, quaternions,
//                     angular error in degrees, translation error)
#include <algorithm>

#include "common.cuh"

namespace dfb {

// ---------------------------------------------------------------------------------------------------------------------
// luma histogram: one block row per image, shared-memory integer bins (exact, order independent)
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kMaxBins = 256;

__global__ void __launch_bounds__(256) k_luma_hist_count(const float* __restrict__ img, int64_t plane, int bins,
                                                         unsigned int* __restrict__ counts) {
  __shared__ unsigned int sh[kMaxBins];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < bins; i += 256) sh[i] = 0;
  __syncthreads();
  const float* r = img + (size_t)b * 3 * plane;
  const float* g = r + plane;
  const float* bl = g + plane;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < plane; i += (int64_t)gridDim.x * 256) {
    // 0.299 * r + 0.587 * g + 0.114 * b, every product and sum rounded to fp32 like the tensor expression
    const float y = __fadd_rn(__fadd_rn(__fmul_rn(0.299f, r[i]), __fmul_rn(0.587f, g[i])), __fmul_rn(0.114f, bl[i]));
    if (y >= 0.f && y <= 1.f) {   // torch.histc ignores values outside [min, max]
      // ATen histc (CPU): pos = (int64)((x - min) / (max - min) * bins), the maximum goes to the last bin
      int pos = (int)__fmul_rn(__fdiv_rn(__fsub_rn(y, 0.f), 1.f), (float)bins);
      pos = pos < bins - 1 ? pos : bins - 1;
      atomicAdd(&sh[pos], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < bins; i += 256)
    if (sh[i]) atomicAdd(&counts[(size_t)b * bins + i], sh[i]);
}

// hist / hist.sum() * 100, torch.round (half to even); the float32 sum of integer counts is exact below 2^24 pixels
__global__ void k_luma_hist_final(const unsigned int* __restrict__ counts, int bins, float* __restrict__ out) {
  const int b = blockIdx.x;
  __shared__ float tot;
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < bins; ++i) s = __fadd_rn(s, (float)counts[(size_t)b * bins + i]);
    tot = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < bins; i += blockDim.x)
    out[(size_t)b * bins + i] = rintf(__fmul_rn(__fdiv_rn((float)counts[(size_t)b * bins + i], tot), 100.f));
}

// ---------------------------------------------------------------------------------------------------------------------
// INTER_AREA downscale of an HWC float image (cv2's area tables: fractional coverage at both ends of a cell)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void area_span(int d, double scale, int ssize, int& s0, int& s1, float& w_first, float& w_full,
                                          float& w_last) {
  // cv2 computeResizeAreaTab: destination cell [d*scale, (d+1)*scale) over the source axis
  const double fsx1 = d * scale, fsx2 = fsx1 + scale;
  const double cell = fmin(scale, (double)ssize - fsx1);
  int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
  sx2 = sx2 < ssize - 1 ? sx2 : ssize - 1;
  sx1 = sx1 < sx2 ? sx1 : sx2;
  w_first = (sx1 - fsx1 > 1e-3) ? (float)((sx1 - fsx1) / cell) : 0.f;
  w_full = (float)(1.0 / cell);
  w_last = (fsx2 - sx2 > 1e-3) ? (float)(fmin(fmin(fsx2 - sx2, 1.0), cell) / cell) : 0.f;
  s0 = sx1, s1 = sx2;
}

__global__ void __launch_bounds__(256) k_resize_area(const float* __restrict__ src, int H, int W, int C, int h, int w,
                                                     float* __restrict__ dst) {
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t total = (int64_t)h * w * C;
  if (idx >= total) return;
  const int c = (int)(idx % C);
  const int dx = (int)((idx / C) % w), dy = (int)(idx / ((int64_t)C * w));
  const double sx = (double)W / w, sy = (double)H / h;
  int x0, x1, y0, y1;
  float wxf, wx, wxl, wyf, wy, wyl;
  area_span(dx, sx, W, x0, x1, wxf, wx, wxl);
  area_span(dy, sy, H, y0, y1, wyf, wy, wyl);
  auto row = [&](int y) {
    const float* p = src + ((int64_t)y * W) * C + c;
    float s = 0.f;
    if (wxf > 0.f) s = fmaf(p[(int64_t)(x0 - 1) * C], wxf, s);
    for (int x = x0; x < x1; ++x) s = fmaf(p[(int64_t)x * C], wx, s);
    if (wxl > 0.f) s = fmaf(p[(int64_t)x1 * C], wxl, s);
    return s;
  };
  float acc = 0.f;
  if (wyf > 0.f) acc = fmaf(row(y0 - 1), wyf, acc);
  for (int y = y0; y < y1; ++y) acc = fmaf(row(y), wy, acc);
  if (wyl > 0.f) acc = fmaf(row(y1), wyl, acc);
  dst[idx] = acc;
}

// ---------------------------------------------------------------------------------------------------------------------
// pose error: one thread per pose pair, double precision inside (a 3x3 problem)
// ---------------------------------------------------------------------------------------------------------------------
// R -> U V^T (the orthogonal polar factor, what torch.svd + matmul(u, v^T) produces) by Jacobi eigen-decomposition of
// R^T R = V diag(s^2) V^T and U V^T = R V diag(1/s) V^T.
// Vout / s2out (nullable): the eigenvectors and eigenvalues of R^T R, i.e. V and the squared singular values of R.
__device__ void polar_orthogonal(const double R[9], double Q[9], double* Vout = nullptr, double* s2out = nullptr) {
  double A[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += R[k * 3 + i] * R[k * 3 + j];
      A[i * 3 + j] = s;
    }
  // cyclic Jacobi converges quadratically: 5-6 sweeps reach 1e-32 of the trace, far below double precision (running the
  // full 30 sweeps - dependent double sqrt / div chains on one thread - made this a 40 us kernel on the step's critical path)
  const double tiny = 1e-32 * (fabs(A[0]) + fabs(A[4]) + fabs(A[8]));
  for (int sweep = 0; sweep < 30; ++sweep) {
    const double off = fabs(A[1]) + fabs(A[2]) + fabs(A[5]);
    if (off <= tiny) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        const double apq = A[p * 3 + q];
        if (fabs(apq) < 1e-300) continue;
        const double theta = (A[q * 3 + q] - A[p * 3 + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {  // A <- A J
          const double akp = A[k * 3 + p], akq = A[k * 3 + q];
          A[k * 3 + p] = c * akp - s * akq, A[k * 3 + q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {  // A <- J^T A
          const double apk = A[p * 3 + k], aqk = A[q * 3 + k];
          A[p * 3 + k] = c * apk - s * aqk, A[q * 3 + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          const double vkp = V[k * 3 + p], vkq = V[k * 3 + q];
          V[k * 3 + p] = c * vkp - s * vkq, V[k * 3 + q] = s * vkp + c * vkq;
        }
      }
  }
  double M[9];  // V diag(1/s) V^T
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += V[i * 3 + k] * V[j * 3 + k] / sqrt(fmax(A[k * 3 + k], 1e-300));
      M[i * 3 + j] = s;
    }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += R[i * 3 + k] * M[k * 3 + j];
      Q[i * 3 + j] = s;
    }
  if (Vout)
    for (int k = 0; k < 9; ++k) Vout[k] = V[k];
  if (s2out)
    for (int k = 0; k < 3; ++k) s2out[k] = fmax(A[k * 3 + k], 1e-300);
}

// svd_reg of the pose regressor (feature/direct_feature_matching.py:81-86: u, s, v = torch.svd(R); R <- u v^T) as one
// kernel without the host synchronisation torch.svd's error check costs a training step.  One thread per 3x3 matrix,
// double precision.  aux [n][21] = U (9) | V (9) | s (3) for the backward.
__global__ void k_polar3x3_fwd(const float* __restrict__ A, int n, float* __restrict__ Q, double* __restrict__ aux) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double R[9], Qd[9], V[9], s2[3];
  for (int k = 0; k < 9; ++k) R[k] = A[(size_t)i * 9 + k];
  polar_orthogonal(R, Qd, V, s2);
  for (int k = 0; k < 9; ++k) Q[(size_t)i * 9 + k] = (float)Qd[k];
  if (aux) {
    double* a = aux + (size_t)i * 21;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {   // U = R V diag(1/s)
        double u = 0;
        for (int k = 0; k < 3; ++k) u += R[r * 3 + k] * V[k * 3 + c];
        a[r * 3 + c] = u / sqrt(s2[c]);
      }
    for (int k = 0; k < 9; ++k) a[9 + k] = V[k];
    for (int k = 0; k < 3; ++k) a[18 + k] = sqrt(s2[k]);
  }
}

// Adjoint of R -> U V^T: with M = U^T G V, dR = U [ (M - M^T)_ij / (s_i + s_j) ] V^T  (the differential of the orthogonal
// polar factor is dQ = U [ (X - X^T)_ij / (s_i + s_j) ] V^T, X = U^T dR V).
__global__ void k_polar3x3_bwd(const double* __restrict__ aux, const float* __restrict__ G, int n, float* __restrict__ dA) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* a = aux + (size_t)i * 21;
  const double *U = a, *V = a + 9, *s = a + 18;
  double g[9], T[9], M[9], N[9];
  for (int k = 0; k < 9; ++k) g[k] = G[(size_t)i * 9 + k];
  for (int r = 0; r < 3; ++r)       // T = U^T G
    for (int c = 0; c < 3; ++c) T[r * 3 + c] = U[0 * 3 + r] * g[0 * 3 + c] + U[1 * 3 + r] * g[1 * 3 + c] + U[2 * 3 + r] * g[2 * 3 + c];
  for (int r = 0; r < 3; ++r)       // M = T V
    for (int c = 0; c < 3; ++c) M[r * 3 + c] = T[r * 3 + 0] * V[0 * 3 + c] + T[r * 3 + 1] * V[1 * 3 + c] + T[r * 3 + 2] * V[2 * 3 + c];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) N[r * 3 + c] = (M[r * 3 + c] - M[c * 3 + r]) / (s[r] + s[c]);
  for (int r = 0; r < 3; ++r)       // T = U N
    for (int c = 0; c < 3; ++c) T[r * 3 + c] = U[r * 3 + 0] * N[0 * 3 + c] + U[r * 3 + 1] * N[1 * 3 + c] + U[r * 3 + 2] * N[2 * 3 + c];
  for (int r = 0; r < 3; ++r)       // dA = T V^T
    for (int c = 0; c < 3; ++c)
      dA[(size_t)i * 9 + r * 3 + c] = (float)(T[r * 3 + 0] * V[c * 3 + 0] + T[r * 3 + 1] * V[c * 3 + 1] + T[r * 3 + 2] * V[c * 3 + 2]);
}

// pytorch3d 0.3.0 transforms.matrix_to_quaternion (requirements.txt:76; not vendored): w = sqrt(max(0, 1 + m00 + m11 +
// m22)) / 2 and x, y, z alike with copysign from the antisymmetric part; float32 like torch.Tensor(...)
__device__ void mat_to_quat(const float m[9], float q[4]) {
  const float m00 = m[0], m11 = m[4], m22 = m[8];
  const float o0 = 0.5f * sqrtf(fmaxf(0.f, 1.f + m00 + m11 + m22));
  const float x = 0.5f * sqrtf(fmaxf(0.f, 1.f + m00 - m11 - m22));
  const float y = 0.5f * sqrtf(fmaxf(0.f, 1.f - m00 + m11 - m22));
  const float z = 0.5f * sqrtf(fmaxf(0.f, 1.f - m00 - m11 + m22));
  q[0] = o0;
  q[1] = copysignf(x, m[7] - m[5]);
  q[2] = copysignf(y, m[2] - m[6]);
  q[3] = copysignf(z, m[3] - m[1]);
}

__global__ void k_pose_error(const float* __restrict__ pred, const float* __restrict__ gt, int n, int use_svd,
                             float* __restrict__ out, float* __restrict__ pred_fixed) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = pred + (size_t)i * 12;
  const float* g = gt + (size_t)i * 12;
  float Rp[9], Rg[9];
  double Rd[9], Qd[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) Rd[r * 3 + c] = p[r * 4 + c], Rg[r * 3 + c] = g[r * 4 + c];
  if (use_svd) {
    polar_orthogonal(Rd, Qd);
    for (int k = 0; k < 9; ++k) Rp[k] = (float)Qd[k];
  } else {
    for (int k = 0; k < 9; ++k) Rp[k] = (float)Rd[k];
  }
  if (pred_fixed)
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) pred_fixed[(size_t)i * 12 + r * 4 + c] = Rp[r * 3 + c];
      pred_fixed[(size_t)i * 12 + r * 4 + 3] = p[r * 4 + 3];
    }
  float q1[4], q2[4];
  mat_to_quat(Rg, q1), mat_to_quat(Rp, q2);
  const float n1 = sqrtf(q1[0] * q1[0] + q1[1] * q1[1] + q1[2] * q1[2] + q1[3] * q1[3]);
  const float n2 = sqrtf(q2[0] * q2[0] + q2[1] * q2[1] + q2[2] * q2[2] + q2[3] * q2[3]);
  float d = 0.f;
  for (int k = 0; k < 4; ++k) d += (q1[k] / n1) * (q2[k] / n2);
  d = fminf(fmaxf(fabsf(d), -1.f), 1.f);
  const float theta = 2.f * acosf(d) * 180.f / 3.14159265358979323846f;
  const float dx = g[3] - p[3], dy = g[7] - p[7], dz = g[11] - p[11];
  out[(size_t)i * 2 + 0] = sqrtf(dx * dx + dy * dy + dz * dz);
  out[(size_t)i * 2 + 1] = theta;
}

}  // namespace dfb

using namespace dfb;

extern "C" int dfb_luma_hist(const float* img, int B, int H, int W, int bins, float* hist, void* ws, size_t ws_bytes,
                             void* stream) {
  DFB_REQUIRE(img && hist && ws && B >= 1 && H >= 1 && W >= 1, DFB_ERR_INVALID, "dfb_luma_hist: bad arguments");
  DFB_REQUIRE(bins >= 1 && bins <= kMaxBins, DFB_ERR_INVALID, "dfb_luma_hist: 1 <= bins <= %d", kMaxBins);
  DFB_REQUIRE((int64_t)H * W < (1 << 24), DFB_ERR_UNSUPPORTED, "dfb_luma_hist: images of 2^24 pixels or more");
  DFB_REQUIRE(ws_bytes >= (size_t)B * bins * 4, DFB_ERR_WORKSPACE, "dfb_luma_hist: workspace needs B*bins*4 bytes");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned int* counts = (unsigned int*)ws;
  DFB_CHECK_CUDA(cudaMemsetAsync(counts, 0, (size_t)B * bins * 4, st));
  const int64_t plane = (int64_t)H * W;
  const int gx = (int)std::min<int64_t>((plane + 255) / 256, 296);
  k_luma_hist_count<<<dim3(gx, B), 256, 0, st>>>(img, plane, bins, counts);
  DFB_LAUNCH_CHECK();
  k_luma_hist_final<<<B, 64, 0, st>>>(counts, bins, hist);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

extern "C" int dfb_resize_area(const float* src, int H, int W, int C, int h, int w, float* dst, void* stream) {
  DFB_REQUIRE(src && dst && H >= 1 && W >= 1 && C >= 1 && h >= 1 && w >= 1, DFB_ERR_INVALID, "dfb_resize_area: bad arguments");
  DFB_REQUIRE(h <= H && w <= W, DFB_ERR_UNSUPPORTED, "dfb_resize_area: INTER_AREA is implemented for downscaling (cv2 switches to "
                                                     "bilinear interpolation when enlarging)");
  const int64_t total = (int64_t)h * w * C;
  k_resize_area<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, H, W, C, h, w, dst);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

extern "C" int dfb_pose_error(const float* pred, const float* gt, int n, int use_svd, float* out, float* pred_fixed,
                              void* stream) {
  DFB_REQUIRE(pred && gt && out && n >= 0, DFB_ERR_INVALID, "dfb_pose_error: bad arguments");
  if (n == 0) return DFB_OK;
  k_pose_error<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(pred, gt, n, use_svd, out, pred_fixed);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

extern "C" int dfb_polar3x3_fwd(const float* A, int n, float* Q, double* aux, void* stream) {
  DFB_REQUIRE(A && Q && n >= 0, DFB_ERR_INVALID, "dfb_polar3x3_fwd: bad arguments");
  if (n == 0) return DFB_OK;
  k_polar3x3_fwd<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(A, n, Q, aux);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

extern "C" int dfb_polar3x3_bwd(const double* aux, const float* G, int n, float* dA, void* stream) {
  DFB_REQUIRE(aux && G && dA && n >= 0, DFB_ERR_INVALID, "dfb_polar3x3_bwd: bad arguments");
  if (n == 0) return DFB_OK;
  k_polar3x3_bwd<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(aux, G, n, dA);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

// ---- batched strided copy ------------------------------------------------------------------------------------------------
// Every parameter of a network goes into its padded fp32 staging tensor before the weight images are re-packed (the NeRF-W
// trainer re-loads 13 / 18 layers per optimizer step).  Done with one tensor.copy_ per weight and bias that was 60+ launches
// of 1-2 us with ~8 us of host time each; here the whole list is one launch: block = (item, row range), thread = column.
namespace {
constexpr int kMaxCopy = 96;
struct CopyBatchArgs {
  DfbCopy2d it[kMaxCopy];
  int n;
};
__global__ void __launch_bounds__(256) k_copy2d_batch(const __grid_constant__ CopyBatchArgs b) {
  const DfbCopy2d& c = b.it[blockIdx.x];
  const int64_t total = (int64_t)c.rows * c.cols;
  for (int64_t i = (int64_t)blockIdx.y * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.y * 256) {
    const int r = (int)(i / c.cols), col = (int)(i - (int64_t)r * c.cols);
    c.dst[(int64_t)r * c.dst_ld + col] = c.src[(int64_t)r * c.src_ld + col];
  }
}
}  // namespace

extern "C" int dfb_copy2d_batch(const DfbCopy2d* items, int n, void* stream) {
  DFB_REQUIRE(n >= 0 && (items || n == 0), DFB_ERR_INVALID, "dfb_copy2d_batch: bad arguments");
  for (int i0 = 0; i0 < n; i0 += kMaxCopy) {
    CopyBatchArgs b;
    b.n = std::min(kMaxCopy, n - i0);
    int64_t big = 1;
    for (int i = 0; i < b.n; ++i) {
      const DfbCopy2d& c = items[i0 + i];
      DFB_REQUIRE(c.src && c.dst && c.rows >= 0 && c.cols >= 0 && c.src_ld >= c.cols && c.dst_ld >= c.cols, DFB_ERR_INVALID,
                  "dfb_copy2d_batch: item %d is malformed", i0 + i);
      b.it[i] = c;
      big = std::max<int64_t>(big, (int64_t)c.rows * c.cols);
    }
    const int gy = (int)std::min<int64_t>(32, (big + 2047) / 2048);
    k_copy2d_batch<<<dim3(b.n, gy), 256, 0, (cudaStream_t)stream>>>(b);
    DFB_LAUNCH_CHECK();
  }
  return DFB_OK;
}

// ---- rays of a pose and their adjoint ---------------------------------------------------------------------------------------
// get_rays (models/ray_utils.py:5-15) + viewdirs = rays_d / |rays_d| (rendering.py:366-370) for the differentiable render of
// train_on_batch, where the pose carries gradient: as tensor expressions that is ~20 launches forward and ~25 backward in
// the host-bound first 2 ms of the step.  Same arithmetic as k_prep_rays.  Adjoint: with dir_i = (dx, dy, -1) the pixel
// direction, g_R[k][b] = sum_i gd_i[k] dir_i[b], g_t[k] = sum_i g_o_i[k], gd = g_d + (g_v - v (v . g_v)) / |d|.
namespace {
__device__ __forceinline__ void pixel_dir(int64_t r, int H, int W, float focal, float& dx, float& dy) {
  const int pj = (int)(r / W), pi = (int)(r % W);
  dx = __fdiv_rn(__fsub_rn((float)pi, (float)(W * 0.5)), focal);
  dy = -__fdiv_rn(__fsub_rn((float)pj, (float)(H * 0.5)), focal);
}

__global__ void __launch_bounds__(256) k_pose_rays_fwd(const float* __restrict__ c2w, int ld, int H, int W, float focal,
                                                       float* __restrict__ ro, float* __restrict__ rd, float* __restrict__ vd) {
  const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (r >= (int64_t)H * W) return;
  float dx, dy, d[3];
  pixel_dir(r, H, W, focal, dx, dy);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float* R = c2w + k * ld;
    d[k] = __fadd_rn(__fadd_rn(__fmul_rn(dx, R[0]), __fmul_rn(dy, R[1])), __fmul_rn(-1.0f, R[2]));
    ro[r * 3 + k] = R[3];
    rd[r * 3 + k] = d[k];
  }
  const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
#pragma unroll
  for (int k = 0; k < 3; ++k) vd[r * 3 + k] = __fdiv_rn(d[k], nrm);
}

constexpr int kPoseBlocks = 64;

__global__ void __launch_bounds__(256) k_pose_rays_bwd(const float* __restrict__ c2w, int ld, int H, int W, float focal,
                                                       const float* __restrict__ g_o, const float* __restrict__ g_d,
                                                       const float* __restrict__ g_v, double* __restrict__ ws) {
  double acc[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) acc[k] = 0.0;
  const int64_t N = (int64_t)H * W;
  for (int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x; r < N; r += (int64_t)gridDim.x * 256) {
    float dx, dy, d[3], gd[3];
    pixel_dir(r, H, W, focal, dx, dy);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float* R = c2w + k * ld;
      d[k] = dx * R[0] + dy * R[1] - R[2];
      gd[k] = g_d ? g_d[r * 3 + k] : 0.f;
    }
    if (g_v) {
      const float inv = rsqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      const float v[3] = {d[0] * inv, d[1] * inv, d[2] * inv};
      const float gv[3] = {g_v[r * 3], g_v[r * 3 + 1], g_v[r * 3 + 2]};
      const float dot = v[0] * gv[0] + v[1] * gv[1] + v[2] * gv[2];
#pragma unroll
      for (int k = 0; k < 3; ++k) gd[k] += (gv[k] - v[k] * dot) * inv;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      acc[k * 4 + 0] += (double)(gd[k] * dx), acc[k * 4 + 1] += (double)(gd[k] * dy), acc[k * 4 + 2] -= (double)gd[k];
      if (g_o) acc[k * 4 + 3] += (double)g_o[r * 3 + k];
    }
  }
  __shared__ double sm[8][12];
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    double v = acc[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 12) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += sm[w][threadIdx.x];
    ws[blockIdx.x * 12 + threadIdx.x] = v;
  }
}

__global__ void k_pose_rays_bwd_final(const double* __restrict__ ws, int blocks, float* __restrict__ out) {
  if (threadIdx.x >= 12) return;
  double v = 0.0;
  for (int b = 0; b < blocks; ++b) v += ws[b * 12 + threadIdx.x];
  out[threadIdx.x] = (float)v;
}
}  // namespace

extern "C" int dfb_pose_rays_fwd(const float* c2w, int row_stride, int H, int W, float focal, float* rays_o, float* rays_d,
                                 float* viewdirs, void* stream) {
  DFB_REQUIRE(c2w && rays_o && rays_d && viewdirs && H > 0 && W > 0 && row_stride >= 4, DFB_ERR_INVALID, "dfb_pose_rays_fwd: bad arguments");
  const int64_t N = (int64_t)H * W;
  k_pose_rays_fwd<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(c2w, row_stride, H, W, focal, rays_o, rays_d, viewdirs);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

extern "C" size_t dfb_pose_rays_workspace_bytes(void) { return (size_t)kPoseBlocks * 12 * sizeof(double); }

extern "C" int dfb_pose_rays_bwd(const float* c2w, int row_stride, int H, int W, float focal, const float* g_rays_o, const float* g_rays_d,
                                 const float* g_viewdirs, void* ws, float* g_c2w12, void* stream) {
  DFB_REQUIRE(c2w && ws && g_c2w12 && H > 0 && W > 0 && row_stride >= 4, DFB_ERR_INVALID, "dfb_pose_rays_bwd: bad arguments");
  const int64_t N = (int64_t)H * W;
  const int blocks = (int)std::min<int64_t>(kPoseBlocks, (N + 255) / 256);
  k_pose_rays_bwd<<<blocks, 256, 0, (cudaStream_t)stream>>>(c2w, row_stride, H, W, focal, g_rays_o, g_rays_d, g_viewdirs, (double*)ws);
  DFB_LAUNCH_CHECK();
  k_pose_rays_bwd_final<<<1, 32, 0, (cudaStream_t)stream>>>((const double*)ws, blocks, g_c2w12);
  DFB_LAUNCH_CHECK();
  return DFB_OK;
}

