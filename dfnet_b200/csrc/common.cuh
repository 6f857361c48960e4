// Internal shared declarations for libdfnet_b200 (not part of the public ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/dfnet_b200.h"

namespace dfb {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
// Device-side "mbarrier wait timed out" flag of the CURRENT device (one int per device, allocated on first use;
// the kernels only ever write it on a protocol failure, so sharing it between launches is benign).
int device_error_flag(int** out);

#define DFB_CHECK_CUDA(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      dfb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return DFB_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

#define DFB_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      dfb::set_error(__VA_ARGS__);    \
      return (code);                  \
    }                                 \
  } while (0)

#define DFB_LAUNCH_CHECK()                                                      \
  do {                                                                          \
    dfb::count_launch();                                                        \
    cudaError_t _e = cudaGetLastError();                                        \
    if (_e != cudaSuccess) {                                                    \
      dfb::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return DFB_ERR_CUDA;                                                      \
    }                                                                           \
  } while (0)


// Launch with programmatic stream serialization when `pdl` (the kernel calls pdl_wait() before it touches global memory,
// see tc_common.cuh); an ordinary launch otherwise.  DFB_PDL=0 in the environment turns it off everywhere.
inline bool dfb_pdl_env() {
  const char* e = getenv("DFB_PDL");
  return !(e && e[0] == '0');
}
template <typename... KArgs, typename... Args>
inline cudaError_t dfb_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = (pdl && dfb_pdl_env()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

constexpr int kRayRec = 12;  // o3 d3 near far vd3 pad
constexpr size_t kReluMaskWordsPerTile = DFB_RELU_MASK_WORDS_PER_TILE;

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// One network's parameters in kernel layouts (device memory unless noted).
struct NetPack {
  bool loaded = false;
  bool fine = false;  // has appearance input + transient branch
  int D = 0, W = 0, skip = -1, pek = 0, in_xyz = 0, in_dir = 0, a_dim = 0, t_dim = 0;

  // ---- fp32 SIMT layout: every matrix transposed to [K][N] (N contiguous) ----------
  float* blob32 = nullptr;          // single allocation, offsets below are in floats
  size_t blob32_floats = 0;
  std::vector<size_t> trunk_w;      // D entries; K = pek (layer 0), pek+W (skip layer: [pe|h]) or W
  std::vector<size_t> trunk_b;
  size_t sigma_w = 0, sigma_b = 0;  // [W], [1]
  size_t final_w = 0, final_b = 0;  // [W][W], [W]
  size_t dt_w = 0;                  // [W][Ndt]: dir_encoding[:, :W] (cols 0..W/2) | transient_encoding.0[:, :W]
  int n_dt = 0;                     // W/2 (static only) or W (fine)
  size_t dirx_w = 0, dirx_b = 0;    // ray-constant part of dir_encoding: [in_dir+a_dim][W/2], bias [W/2]
  size_t tx_w = 0, tx_b = 0;        // ray-constant part of transient_encoding.0: [t_dim][W/2], bias [W/2]
  size_t t_w[3] = {0, 0, 0}, t_b[3] = {0, 0, 0};  // transient_encoding.{2,4,6}: [W/2][W/2]
  size_t rgb_w = 0, rgb_b = 0;      // static_rgb [3][W/2] (row-major, as in torch), [3]
  size_t tsig_w = 0, tsig_b = 0;    // [W/2], [1]
  size_t trgb_w = 0, trgb_b = 0;    // [3][W/2], [3]
  size_t tbeta_w = 0, tbeta_b = 0;  // [W/2], [1]

  // ---- fp32 backward layout (fine network): torch's [out][in] rows, `in` padded so that the input-
  // gradient GEMMs g_in = W^T g_out run with the same kernel structure (mlp_simt_bwd.cu) -------------
  float* blob32b = nullptr;
  std::vector<size_t> bw_trunk;     // D entries: [W][in_pad]; layer 0: in_pad = pek; skip layer: [pe(pek) | h(W)]
  size_t bw_final = 0;              // [W][W]
  size_t bw_dt = 0;                 // [W][W]: rows 0..W/2 dir_encoding[:, :W], rows W/2..W transient_encoding.0[:, :W]
  size_t bw_dtx = 0;                // [W/2][32]: dir_encoding[:, W:W+27] (view-direction encoding columns)
  size_t bw_t[3] = {0, 0, 0};       // transient_encoding.{2,4,6}: [W/2][W/2]

  // ---- tcgen05 layout (W == 256 only): 16-bit core-matrix panels, see mlp_tc.cu ------
  void* blob16[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // [fp16|bf16][cta_group 1|2 chunking]
  size_t blob16_bytes = 0;
  std::vector<float> tc_tbl;  // host copy of the bias / head-weight table passed as kernel parameter
  float* tc_dtbias_dev = nullptr;  // [W] constant bias of the dir|transient.0 step (W_dt b_final when folded), added by k_raybias
  // split-precision coarse pass (DFB_MMA_F16_SPLIT_COARSE): [hi image | lo image] of the sigma-only program in the
  // cta_group::2 chunking (lo = rn16(w - rn16(w))), and the fp32 biases [step][256] in global memory
  void* blob16x3 = nullptr;
  size_t blob16x3_bytes = 0;
  float* tc_bias32_dev = nullptr;
  // native 128-wide program (W == 128, the reference's default netwidth; mlp_tc.cu build_program): cta_group::2 images
  void* blob16n[2] = {nullptr, nullptr};  // [fp16|bf16]
  size_t blob16n_bytes = 0;
  std::vector<float> tc_tbl_n;
  float* tc_dtbias_n_dev = nullptr;       // [128] contiguous [dir 64 | transient 64]

  // ---- tcgen05 backward layout (fine 8x256 network): the 26-step image of mlp_tc_bwd.cu ------
  void* blob16b[2] = {nullptr, nullptr};  // [fp16|bf16]
  size_t blob16b_bytes = 0;
  void* blob16b2[2] = {nullptr, nullptr}; // the same program in the cta_group::2 chunking (16 KB half-chunk images)
  size_t blob16b2_bytes = 0;
  std::vector<float> tcb_tbl;
};

}  // namespace dfb

struct DfbNerf {
  DfbNerfDesc desc;
  int device = 0;
  dfb::NetPack net[2];
  float* emb_a = nullptr;  // [n_vocab,5]
  float* emb_t = nullptr;  // [n_vocab,2]
  bool has_emb = false;
  int num_sms = 0;
  float* lin_dev = nullptr;  // [Nc | Nf] linspace grids of the last render config
  int lin_nc = -1, lin_nf = -1;
  // per-handle (= per-device) launch scratch: the handle's launches are ordered on one stream at a time
  mutable uint32_t* bwd_scratch = nullptr;  // tcgen05 backward: L2-resident per-CTA scratch (mlp_tc_bwd.cu)
  mutable int bwd_scratch_ctas = 0;
};

namespace dfb {

// ---- kernels / launchers defined across translation units ------------------------------
enum MlpMode { MLP_SIGMA = 0, MLP_STATIC = 1, MLP_FULL = 2 };

// rays+z driven MLP (SIMT fp32).  raw: [P, C] with C = 1/4/9 for the three modes.
int launch_mlp_simt_rays(const DfbNerf* nerf, int which, int mode, const float* rayrec, const float* z,
                         const float* raybias, int64_t n_rays, int S, float* raw, cudaStream_t st);
// embedded-input MLP (SIMT fp32), the NeRFW.forward seam.
int launch_mlp_simt_embedded(const DfbNerf* nerf, int which, int mode, const float* x, int64_t P, float* out,
                             cudaStream_t st);
// tcgen05 MLP (W == 256).  kind: DFB_MMA_F16 / DFB_MMA_BF16.
// masks (nullable, fine network only): ReLU masks for the tcgen05 backward, see TcArgs::masks in mlp_tc.cu
// split3: the split-precision variant of the sigma-only pass (three MMA sub-steps per layer on hi/lo fp16 operands)
int launch_mlp_tc_rays(const DfbNerf* nerf, int which, int mode, int kind, const float* rayrec, const float* z,
                       const float* raybias, int64_t n_rays, int S, float* raw, cudaStream_t st,
                       uint32_t* masks = nullptr, bool split3 = false, float* part = nullptr, int part_k = 0,
                       const int* ert_rowmap = nullptr, const int* ert_offsets = nullptr);
// early ray termination: offsets [N+1] = exclusive scan of n_live (offsets[N] = live rows), rowmap[row] = ray*S + i
int launch_ert_compact(const int* n_live, int64_t n_rays, int S, int* offsets, int* rowmap, cudaStream_t st);
// fused compositing (fine pass at test time): records per ray of launch_mlp_tc_rays(part = ...) and the kernel that
// chains them into rgb / disp / acc
inline int composite_part_k(int S) { return (S + 30) / 32 + 1; }
int launch_composite_partials(const float* part, int part_k, int64_t n_rays, int S, float* rgb, float* disp, float* acc,
                              cudaStream_t st, const int* ert_offsets = nullptr);
bool tc_supported(const DfbNerf* nerf, int which, int mode);
// 3-D tensor map over a packed image of 16 KB chunks: [n][64][128 x u16], one box = one chunk (2-SM TMA weight loads)
int make_weight_tmap(void* base, size_t bytes, CUtensorMap* out);
// 5-D tensor map over an NHWC 16-bit tensor seen as [B][C/8 panels][H][W][8 channels]; one box = npanels x PH x PW pixels
// x 8 channels = the shared-memory image [panel][row][col][16 B] (conv_tc.cu); out-of-bounds coordinates are zero-filled
int make_patch_tmap(const void* in, int B, int H, int W, int Cpad, int PH, int PW, CUtensorMap* out, int npanels = 8);
// Networks narrower than 256 run on the 256-wide tcgen05 kernels EXACTLY, embedded with zero weights / zero biases
// (a ReLU unit with zero input weights and bias stays at 0 and feeds nothing): tc_pad_params returns the state dict
// of the equivalent 8x256 network (state-dict order, hidden units 0..W-1 / 0..W/2-1 live).
bool tc_padded_shape(const dfb::NetPack& np);
int tc_cta_group_env();
bool tc_native128(const DfbNerf* nerf, int which, bool masks, bool split3);
std::vector<std::vector<float>> tc_pad_params(const dfb::NetPack& np, const std::vector<std::vector<float>>& P);
int pack_tc_weights(DfbNerf* nerf, int which, const std::vector<std::vector<float>>& P);
// tcgen05 backward of the fine network w.r.t. its inputs (mlp_tc_bwd.cu)
int pack_tc_bwd_weights(DfbNerf* nerf, int which, const std::vector<std::vector<float>>& P);
bool tc_bwd_supported(const DfbNerf* nerf);
// saved_masks (nullable): relu_masks of the training forward for these rays -> no forward recompute
int launch_mlp_tc_bwd(const DfbNerf* nerf, int kind, const float* rayrec, const float* z, const float* raybias,
                      const float* raw, const float* g_raw, int64_t n_rays, int S, float* g_samp, cudaStream_t st,
                      const uint32_t* saved_masks = nullptr);


// ---- argument blocks of the non-MLP render kernels (render_kernels.cu) ----------------
struct PrepArgs {
  const float* rays;    // [N, 11+hb] or null
  const float* c2w;     // [3,4] (row stride c2w_ld) or null; n_pose > 1: [n_pose][3][c2w_ld] back to back, image-major rays
  int c2w_ld;
  int n_pose;           // c2w mode: number of poses (images of H x W rays each); hist is [n_pose][hb] then
  int H, W;
  float focal, near, far;
  const float* hist;    // [hb] (c2w mode)
  int hb, n_vocab;
  const float* emb_a;   // [n_vocab,5] or null
  const float* emb_t;   // [n_vocab,2] or null
  int64_t N;
  int64_t pix0;         // c2w mode: pixel index of ray 0 of this launch
  int Nc;
  const float* t_vals;  // [Nc] device
  const float* t_rand;  // [N,Nc] or null
  int lindisp;
  float* rayrec;        // [N,12]
  float* extra;         // [N, n_extra] or null : dirPE(27) | a(5*hb) | t(2*hb)
  int n_extra, a_dim, t_dim;
  float* z;             // [N,Nc]
};

struct CompositeArgs {
  const float* raw;  // [N,S,C]
  const float* z;    // [N,S]
  int64_t N;
  int S, C;          // C: 1 (coarse+test: sigma), 4 (rgb,sigma), 9 (full)
  int typ_fine, test_time;
  float beta_min;
  float *rgb, *disp, *acc, *weights, *depth, *tsig, *beta;  // any may be null
  const float* noise;  // [N,S] standard-normal draws (coarse pass only, rendering.py:173-174) or null
  float noise_std;
};

struct SampleArgs {
  // mode A (render path): z_c [N,Nc] and coarse weights [N,Nc]; bins = mids, weights[1:-1]
  const float* z_c;
  const float* w_c;
  int Nc;
  // mode B (seam): bins [N,nb], weights [N,nb-1]
  const float* bins;
  const float* weights;
  int nb;
  const float* u;      // [N,Nf] or null
  const float* u_lin;  // [Nf] linspace(0,1,Nf) when u == null
  int64_t N;
  int Nf;
  float* samples;      // [N,Nf] or null
  int32_t* inds;       // [N,Nf] or null
  float* z_vals;       // [N,Nc+Nf] sorted union, or null (mode A only)
  float* z_std;        // [N] or null
  // opt-in early ray termination (mode A): n_live[ray] = number of sorted depths up to the coarse sample behind which
  // the COARSE transmittance 1 - cumsum(w_c) has fallen below ert_eps (>= 1); null = off
  float ert_eps;
  int* n_live;
};

int launch_prep(const PrepArgs& a, cudaStream_t st);
// add_bias [n_rb] (nullable): constant added to every row; pack_kind: 0 fp32 rows, 1 / 2 = fp16 / bf16 pairs
// packed into the first n_rb/2 words of each row (what the tcgen05 forward kernel consumes)
// pad_h (> W/2): row layout of the zero-padded 8x256 embedding: transient half at column pad_h, zeros elsewhere
int launch_raybias(const float* extra, int ld, int64_t N, const NetPack& np, bool with_transient, float* rb,
                   int rb_ld, cudaStream_t st, const float* add_bias = nullptr, int pack_kind = 0, int pad_h = 0);
int launch_composite(const CompositeArgs& a, cudaStream_t st);
int launch_sample(const SampleArgs& a, cudaStream_t st);

}  // namespace dfb
