/*
 * dfnet_b200 — C ABI of the B200-native NeRF-Hist render / DFNet feature hot path.
 *
 * The reference (ActiveVisionLab/DFNet) is pure Python/PyTorch and has no FFI; its
 * boundary for this path is the Python call surface listed in SURVEY.md §8(b).  Every
 * entry point below cites the reference function (path relative to script/) whose
 * arithmetic it replaces.  The Python shims that keep the reference's signatures live in
 * dfnet_b200/ (rendering.py, nerfw.py, ...) and bind these symbols through ctypes; see
 * INTEGRATION.md for the stub a reference maintainer would add.
 *
 * Conventions
 *   - every call returns 0 on success, <0 on error; dfb_last_error() gives the message
 *     (thread-local, valid until the next failing call on the thread);
 *   - no exceptions, no allocations and no torch types cross the boundary;
 *   - device buffers (inputs, outputs, workspace) are owned by the caller and passed as
 *     raw pointers + sizes; the library owns only the opaque handles (repacked weights);
 *   - launches are asynchronous on the `stream` argument (a cudaStream_t passed as void*);
 *   - handles are bound to the device that was current at creation; not thread-safe.
 */
#ifndef DFNET_B200_H_
#define DFNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFB_OK 0
#define DFB_ERR_INVALID (-1)
#define DFB_ERR_CUDA (-2)
#define DFB_ERR_UNSUPPORTED (-3)
#define DFB_ERR_WORKSPACE (-4)

/* How the 256-wide contractions are computed. */
#define DFB_MMA_FP32_SIMT 0 /* fp32 FFMA, any width (exact-order fallback for odd sizes)   */
#define DFB_MMA_F16 1       /* tcgen05.mma kind::f16, fp16 operands, fp32 accumulate in TMEM */
#define DFB_MMA_BF16 2      /* tcgen05.mma kind::f16, bf16 operands, fp32 accumulate in TMEM */
/* DFB_MMA_F16 with a SPLIT-PRECISION coarse pass: the sigma-only coarse network - the pass that decides where the fine
 * samples are placed - runs on hi + lo fp16 operand pairs (A_hi W_hi + A_lo W_hi + A_hi W_lo, fp32 accumulate; ~22
 * mantissa bits, 3x the coarse tensor work); the fine network runs as DFB_MMA_F16.  On fields with sharp density steps
 * the fp16 rounding of the coarse pass moves fine samples across surfaces and dominates the image error;
 * this kind removes that term.  (kind::tf32 would not: it has fp16's 10-bit mantissa.) */
#define DFB_MMA_F16_SPLIT_COARSE 3

const char* dfb_last_error(void);
int dfb_version(void);
/* 1 when a CUDA device of compute capability 10.x is current, else 0 (never fails). */
int dfb_device_ok(void);

/* torch.linspace(start,end,steps) float32 exactly as ATen-CPU evaluates it (host helper;
 * models/rendering.py:32,269 build u and t_vals with it). */
int dfb_linspace_f32(float start, float end, int steps, float* out_host);

/* ------------------------------------------------------------------------------------
 * NeRF-W / NeRF-Hist networks  (models/nerfw.py:220-354 NeRFW, :356-502 create_nerf)
 * ---------------------------------------------------------------------------------- */
typedef struct DfbNerf DfbNerf;

typedef struct DfbNerfDesc {
  int32_t D;        /* trunk depth   (args.netdepth, 8)                      */
  int32_t W;        /* trunk width   (args.netwidth, 256 / 128 / 64)         */
  int32_t skip;     /* trunk layer index that concatenates input_xyz (4); <0 = none */
  int32_t L_xyz;    /* positional-encoding bands for xyz (multires, 10)      */
  int32_t L_dir;    /* bands for view directions (multires_views, 4)         */
  int32_t a_dim;    /* appearance code width = hist_bin*5 (in_channels_a, 50) */
  int32_t t_dim;    /* transient code width  = hist_bin*2 (in_channels_t, 20) */
  int32_t hist_bin; /* histogram bins (10)                                   */
  int32_t n_vocab;  /* rows of embedding_a / embedding_t (1000)              */
  float beta_min;   /* NeRFW.beta_min (0.1)                                  */
  int32_t has_fine; /* 1: coarse + fine networks, 0: coarse only (N_importance == 0) */
} DfbNerfDesc;

int dfb_nerf_create(const DfbNerfDesc* desc, DfbNerf** out);
void dfb_nerf_destroy(DfbNerf* nerf);

/* Load one network's parameters.  which: 0 = network_fn (coarse), 1 = network_fine.
 * params[i] points to the i-th tensor of NeRFW.state_dict() (fp32, contiguous, host or
 * device memory), numel[i] is its element count; order and sizes are checked against the
 * descriptor.  Weights are repacked into the kernels' layouts (fp32 K-major for SIMT,
 * fp16/bf16 core-matrix panels for tcgen05). */
int dfb_nerf_load(DfbNerf* nerf, int which, const float* const* params, const int64_t* numel, int n_params);
/* embedding_a.weight [n_vocab,5] and embedding_t.weight [n_vocab,2] (nerfw.py:386-394). */
int dfb_nerf_set_embeddings(DfbNerf* nerf, const float* emb_a, const float* emb_t);

/* NeRFW.forward on already-embedded points (models/nerfw.py:297-354).
 * mode: 0 sigma_only (x [P,63] -> out [P,1]); 1 static (x [P,63+27+a] -> [P,4], a = a_dim
 * for the fine net, 0 for coarse); 2 full (x [P,63+27+a+t] -> [P,9], fine net only).
 * fp32 SIMT path; op-level seam used by NeRFW.forward and the parity tests. */
int dfb_nerfw_forward(DfbNerf* nerf, int which, int mode, const float* x, int64_t P, float* out, void* stream);

/* ------------------------------------------------------------------------------------
 * Rendering  (models/rendering.py:245-400 render_rays / batchify_rays / render,
 *             models/ray_utils.py:5-15 get_rays, models/nerfw.py:15-95 run_network_NeRFW)
 * ---------------------------------------------------------------------------------- */
typedef struct DfbRenderCfg {
  int32_t N_samples;    /* coarse samples per ray                                     */
  int32_t N_importance; /* extra fine samples per ray (0 = coarse only)               */
  int32_t test_time;    /* render_kwargs_test: sigma-only coarse pass, static-only depth */
  int32_t perturb;      /* 1: stratified jitter; t_rand and u must be supplied         */
  int32_t mma_kind;     /* DFB_MMA_*                                                   */
  int32_t lindisp;      /* sample linearly in disparity (rendering.py:272-273)         */
  float raw_noise_std;  /* std of the density noise (rendering.py:173-174); != 0 needs the `noise` draws  */
  int32_t ray_stride;   /* rays mode: floats per ray record as laid out by the caller; must be
                           11 + hist_bin (the [o3,d3,near,far,viewdir3,hist] row of rendering.py:382-389)  */
  int32_t hist_len;     /* c2w mode: number of floats behind `hist`; must be hist_bin            */
  float ert_eps;        /* opt-in early ray termination (0 = off, the parity path): fine samples behind the
                           depth where the COARSE transmittance falls below ert_eps are not evaluated   */
} DfbRenderCfg;

/* Optional outputs (NULL = not wanted).  Train-mode extras follow rendering.py:318-331. */
typedef struct DfbRenderExtras {
  float* rgb0;             /* [N,3]  coarse composite (train mode)                      */
  float* disp0;            /* [N]                                                       */
  float* acc0;             /* [N]                                                       */
  float* z_std;            /* [N]    std of the fine samples, unbiased=False            */
  float* beta;             /* [N]                                                       */
  float* transient_sigmas; /* [N,S]                                                     */
  float* raw;              /* [N,S,9] fine raw (or [N,Nc,4] when N_importance == 0)     */
  float* weights_coarse;   /* [N,Nc]  seam: coarse weights fed to sample_pdf            */
  float* z_vals;           /* [N,S]   seam: sorted union of coarse and fine depths      */
  float* z_samples;        /* [N,Nf]  seam: sample_pdf output                           */
  int32_t* inds;           /* [N,Nf]  seam: searchsorted indices (int64 in the reference) */
  float* depth;            /* [N]                                                       */
  uint32_t* relu_masks;    /* [ceil(N*S/128), 12, 8, 128] ReLU masks of the fine network's 12 hidden layers, one bit per
                              activation (tcgen05 path only): input of dfb_render_bwd_saved               */
  int32_t* n_live;         /* [N]   early ray termination (cfg->ert_eps > 0): fine samples evaluated per ray */
} DfbRenderExtras;
#define DFB_RELU_MASK_WORDS_PER_TILE (12 * 8 * 128)

/* Bytes of device workspace dfb_render_fwd needs for n_rays rays. */
int dfb_render_workspace_bytes(const DfbNerf* nerf, const DfbRenderCfg* cfg, int64_t n_rays, size_t* out);

/* render() (rendering.py:353-400) for N rays.
 * Ray source, exactly one of:
 *   rays  != NULL : device [N, 11+hist_bin] records [o3,d3,near,far,viewdir3,hist] (:382-389)
 *   c2w   != NULL : device [3,4] (or [4,4]) pose; rays are generated in-kernel like get_rays
 *                   for an H x W image with N == H*W; near/far scalars; hist device [hist_bin].
 * t_rand [N,Nc] / u [N,Nf]: uniform draws the reference takes from torch.rand when
 * perturb > 0 (:282, :36); NULL when perturb == 0.
 * noise [N,Nc]: the standard-normal draws of the coarse compositing (torch.randn_like(static_sigmas), :173);
 * required when cfg->raw_noise_std != 0, else NULL (the reference multiplies them by 0).
 * Outputs rgb [N,3], disp [N], acc [N] (device). */
int dfb_render_fwd(DfbNerf* nerf, const DfbRenderCfg* cfg, const float* rays, const float* c2w, int H, int W,
                   float focal, float near, float far, const float* hist, int64_t N, const float* t_rand,
                   const float* u, const float* noise, float* rgb, float* disp, float* acc,
                   const DfbRenderExtras* extras, void* ws, size_t ws_bytes, void* stream);

/* Batched multi-pose render: n_pose poses c2w [n_pose,3,4] with histograms hist [n_pose,hist_bin] (device), one H x W
 * image each in ONE call; outputs [n_pose*H*W, ...] image-major.  Random view synthesis renders hundreds of small virtual
 * views per refresh (feature/misc.py:249-289 `render_virtual_imgs`, one render() call per view in the reference).
 * cfg->hist_len = hist_bin.  Workspace as dfb_render_workspace_bytes for n_pose*H*W rays. */
int dfb_render_poses_fwd(DfbNerf* nerf, const DfbRenderCfg* cfg, const float* c2w, int n_pose, int H, int W, float focal,
                         float near, float far, const float* hist, float* rgb, float* disp, float* acc, void* ws, size_t ws_bytes,
                         void* stream);

/* Same as dfb_render_fwd with c2w/hist and the three outputs in HOST memory (pinned for
 * true asynchrony): the pose/histogram upload and the image download are enqueued on
 * `stream` around the kernels.  This is the call render_path() makes per image
 * (rendering.py:420-424: render + .cpu()). The caller synchronises the stream. */
int dfb_render_image_host(DfbNerf* nerf, const DfbRenderCfg* cfg, const float* c2w_host, int H, int W, float focal,
                          float near, float far, const float* hist_host, float* rgb_host, float* disp_host,
                          float* acc_host, void* ws, size_t ws_bytes, void* stream);

/* Backward of the test-time render w.r.t. the rays — what train.py needs from the renderer
 * (feature/direct_feature_matching.py:342-378: the NeRF weights are frozen and z_samples detached,
 * models/rendering.py:302, so only the fine network's inputs carry gradient).
 * rays [N,ray_stride] as in the forward (ray_stride must be 11+hist_bin); z_vals [N,S] and raw [N,S,9] are the forward's extras
 * (S = N_samples + N_importance); g_rgb [N,3] is dLoss/d rgb_map.  Outputs: gradients w.r.t. rays_o,
 * rays_d (through pts = o + d*z) and the view directions (through the direction encoding), each [N,3].
 * fp32 kernels (forward recompute + input-gradient chain); all device pointers. */
int dfb_render_bwd_workspace_bytes(const DfbNerf* nerf, int64_t n_rays, int S, size_t* out);
int dfb_render_bwd(DfbNerf* nerf, const float* rays, int ray_stride, int64_t N, int S, const float* z_vals, const float* raw,
                   const float* g_rgb, float* g_rays_o, float* g_rays_d, float* g_viewdirs, void* ws, size_t ws_bytes,
                   void* stream);
/* Same with the fine network's forward recompute and input-gradient chain on the tensor cores (tcgen05, 8x256
 * networks; mma_kind as in DfbRenderCfg: the kind the forward ran with).  DFB_MMA_FP32_SIMT = dfb_render_bwd;
 * other network shapes fall back to the fp32 kernels. */
int dfb_render_bwd_mma(DfbNerf* nerf, int mma_kind, const float* rays, int ray_stride, int64_t N, int S, const float* z_vals,
                       const float* raw, const float* g_rgb, float* g_rays_o, float* g_rays_d, float* g_viewdirs, void* ws,
                       size_t ws_bytes, void* stream);

/* Same as dfb_render_bwd_mma for a forward that saved its ReLU masks (DfbRenderExtras::relu_masks, same rays, same
 * mma_kind): the backward kernel skips the forward recompute (15 instead of 26 MMA steps per tile).  relu_masks == NULL
 * recomputes.  N*S must keep every internal 16384-ray chunk aligned to 128 samples (any S that is a multiple of 1/128
 * of the chunk, e.g. every S when N <= 16384). */
int dfb_render_bwd_saved(DfbNerf* nerf, int mma_kind, const float* rays, int ray_stride, int64_t N, int S, const float* z_vals,
                         const float* raw, const uint32_t* relu_masks, const float* g_rgb, float* g_rays_o, float* g_rays_d,
                         float* g_viewdirs, void* ws, size_t ws_bytes, void* stream);

/* Op-level seams (same arguments as the reference functions). */
/* sample_pdf (rendering.py:24-65): bins [N,nb], weights [N,nb-1], u [N,Nf] or NULL (det). */
int dfb_sample_pdf(const float* bins, const float* weights, const float* u, int64_t N, int n_bins, int Nf,
                   float* samples, int32_t* inds, void* stream);
/* raw2outputs_NeRFW (rendering.py:132-243).  typ: 0 coarse, 1 fine.  raw [N,S,C] with
 * C = 1 (coarse+test), 4 (coarse train) or 9 (fine).  Any output may be NULL.
 * noise [N,S] (nullable) with raw_noise_std: the coarse pass' density noise (:173-174; the fine pass has none). */
int dfb_raw2outputs(const float* raw, const float* z_vals, int64_t N, int S, int C, int typ, int test_time,
                    float beta_min, float* rgb, float* disp, float* acc, float* weights, float* depth,
                    float* transient_sigmas, float* beta, const float* noise, float raw_noise_std, void* stream);
/* get_rays (ray_utils.py:5-15): c2w device [3,4] -> rays_o, rays_d device [H*W,3]. */
int dfb_get_rays(const float* c2w, int row_stride, int H, int W, float focal, float* rays_o, float* rays_d,
                 void* stream);

/* ------------------------------------------------------------------------------------
 * DFNet feature extractor and feature losses
 *   (feature/dfnet.py:42-172 AdaptLayers + DFNet/DFNet_s, feature/direct_feature_matching.py:114-136)
 * ---------------------------------------------------------------------------------- */
typedef struct DfbConv DfbConv;
typedef struct DfbDfnet DfbDfnet;

/* One convolution layer (nn.Conv2d, stride 1, "same" padding, kernel 1/3/5, Cout % 64 == 0) with an
 * optional folded eval-mode BatchNorm (y = scale * (conv + bias) + shift).  weight [Cout,Cin,KH,KW]. */
int dfb_conv_create(int Cin, int Cout, int KH, int KW, const float* weight, const float* bias, const float* bn_scale,
                    const float* bn_shift, DfbConv** out);
void dfb_conv_destroy(DfbConv* conv);
/* in: NHWC fp16 [B,H,W,round_up(Cin,8)].  Outputs (any subset, NULL = skip): out NHWC fp16 after the
 * optional ReLU, tap NHWC fp16 before it, out_nchw32 fp32 [B,Cout,H,W] before it. */
int dfb_conv_fwd(DfbConv* conv, const void* in_nhwc16, int B, int H, int W, int relu, void* out_nhwc16, void* tap_nhwc16,
                 float* out_nchw32, void* stream);

/* n_levels: 3 = DFNet (taps conv1_2, conv3_3, conv5_3), 1 = DFNet_s (conv1_2 only). */
int dfb_dfnet_create(int n_levels, DfbDfnet** out);
void dfb_dfnet_destroy(DfbDfnet* net);
/* params (fp32, host or device): 13 x (encoder conv weight, bias) in VGG-16 order, then per level
 * (conv1x1 w, b, conv5x5 w, b, bn weight, bn bias, bn running_mean, bn running_var), then fc_pose w, b.
 * BatchNorm is folded in eval mode (freezeBN / model.eval(), feature/dfnet.py heads). */
int dfb_dfnet_load(DfbDfnet* net, const float* const* params, const int64_t* numel, int n_params, float bn_eps);
int dfb_dfnet_workspace_bytes(const DfbDfnet* net, int B, int H, int W, int upH, int upW, size_t* out);
/* DFNet.forward (feature/dfnet.py:106-172).  x [B,3,H,W] fp32 in [0,1].
 * flags: bit0 return_feature, bit1 isSingleStream, bit2 return_pose, bit3 keep the tape (training), bit4 bf16
 * encoder (pose-only training), bit5 train-mode BatchNorm in the heads: batch statistics over the whole batch of this
 * call (run_feature.py:133,204 without freezeBN; needs dfb_dfnet_load_ex flags bit2; the statistics are read back with
 * dfb_dfnet_bn_batch_stats for the caller's running-statistics update), bits 8..10 level 0 / 1 / 2 not wanted (its head is
 * skipped and its slice of the stacks left untouched; without a pose the encoder stops after the deepest wanted level).
 * feats_t / feats_r: [L, Bs, 128, upH, upW] fp32, Bs = B (single stream; feats_r unused) or B/2
 * (siamese: first half of the batch -> feats_t, second half -> feats_r).  pose: [B,12]. */
int dfb_dfnet_fwd(DfbDfnet* net, const float* x, int B, int H, int W, uint32_t flags, int upH, int upW, float* feats_t,
                  float* feats_r, float* pose, void* ws, size_t ws_bytes, void* stream);
/* feature_loss: fr, ft fp32 [C,HW]; 1 - mean(cosine) with the cosine over HW per channel
 * (per_channel = 0, the reference default) or over C per pixel (per_channel = 1).  *loss is a device
 * scalar; ws needs max(C*64*3, ceil(HW/256)) floats. */
int dfb_cosine_loss(const float* fr, const float* ft, int C, int64_t HW, int per_channel, float eps, float* loss, void* ws,
                    size_t ws_bytes, void* stream);

/* Resampling of fp32 [planes,h,w] -> [planes,Ho,Wo]:
 *   bicubic     = torch.nn.Upsample(size, mode='bicubic') (align_corners=False, A=-0.75, clamped taps,
 *                 output not clamped) used on the rendered image, feature/direct_feature_matching.py:346;
 *   bilinear_ac = torch.nn.UpsamplingBilinear2d(size) (align_corners=True), feature/dfnet.py:145. */
int dfb_resize_bicubic(const float* src, int64_t planes, int h, int w, int Ho, int Wo, float* dst, void* stream);
int dfb_resize_bilinear_ac(const float* src, int64_t planes, int h, int w, int Ho, int Wo, float* dst, void* stream);

/* triplet_loss_hard_negative_mining_plus (feature/misc.py:399-435): f1, f2 fp32 [L,B,C,H,W];
 * negatives are the batch-rolled stacks, the in-triplet case is argmin of four MSE distances,
 * TripletMarginLoss(margin, p=2, eps=1e-6, mean) reduces over W.  *loss, *chosen_case: device
 * scalars; ws >= 8192 floats. */
int dfb_triplet_loss(const float* f1, const float* f2, int L, int B, int C, int H, int W, float margin, float* loss,
                     int* chosen_case, void* ws, size_t ws_bytes, void* stream);
/* Backward of dfb_triplet_loss (the case selection is a no-grad block in the reference): chosen_case = the forward's
 * device scalar, g_loss = upstream gradient (device scalar); g_f1, g_f2 [L,B,C,H,W] are overwritten. */
int dfb_triplet_loss_bwd(const float* f1, const float* f2, int L, int B, int C, int H, int W, float margin,
                         const int* chosen_case, const float* g_loss, float* g_f1, float* g_f2, void* stream);
/* mean((a-b)^2): nn.MSELoss in PoseLoss (feature/direct_feature_matching.py:138-142) and img2mse
 * (models/nerfw.py:11).  ws >= 1024 floats. */
int dfb_mse(const float* a, const float* b, int64_t n, float* out, void* ws, size_t ws_bytes, void* stream);

/* Number of kernel launches issued by this library since load (bench.py gpu_launches). */
int64_t dfb_launch_count(void);

/* Per-kernel timing of the coarse / fine MLP launches with CUDA events on the launching
 * stream (measurement hook for bench.py's roofline line; off by default).  dfb_profile_read
 * waits for the recorded events, returns summed milliseconds and launch counts since the
 * previous read, and clears the records. */
int dfb_profile_enable(int on);
int dfb_profile_read(double* coarse_ms, double* fine_ms, int64_t* coarse_launches, int64_t* fine_launches);

/* Debug seam, not on the product path: D[128,N] = A[128,K] * B[N,K]^T on one CTA through the
 * same shared-memory descriptors, tcgen05.mma and TMEM loads as the MLP kernel (fp32 in/out,
 * operands rounded to `kind`).  variant 0 = the descriptor convention the kernel uses. */
/* Debug seam: per-CTA cycle counters of the last tcgen05 MLP launch; only libraries built with
 * -DDFB_TC_PROF record them (returns DFB_ERR_UNSUPPORTED otherwise). out_host: [n_cta][16] u64. */
int dfb_debug_tc_prof(unsigned long long* out_host, int n_cta);
/* Same for the tcgen05 render backward (k_mlp_tc_bwd): out_host [n_cta][64] u64 = producer {wait W_EMPTY,-,-,total},
 * issuer {wait W_FULL, wait A_READY / PASS_DONE, wait PE_READY, total}, epilogue slot 0 / slot 1 {wait D_FULL,-,-,total},
 * then the busy cycles of slot 0's epilogue per program step (tools/tcb_prof.py). */
int dfb_debug_tcb_prof(unsigned long long* out_host, int n_cta);
/* Debug seam (tests only): ReLU masks of the render backward's forward recompute, device buffers [P][12][8] uint32
 * (fine 8x256 network, N <= 16384 rays).  simt_dump <- fp32 kernels (word j bit l = column 32j+l); tc_out <- tcgen05
 * kernel (bit 16*(c&1) + (c>>1)%16 of word c/32 = column c); tc_in replaces the tcgen05 kernel's own masks so that its
 * gradient chain can be compared with the fp32 chain on identical ReLU patterns.  NULLs switch the seam off. */
int dfb_debug_bwd_masks(uint32_t* simt_dump, const uint32_t* tc_in, uint32_t* tc_out);

/* Debug seam: measured tensor-pipe cycles per tcgen05.mma (M=128, N=n, K=16) with the kernels' no-swizzle
 * panel layout, `grid` CTAs issuing back to back. */
int dfb_debug_umma_rate(int iters, int n, int grid, double* cycles_per_mma);
/* Debug seam: TMEM read rate, mean cycles per tcgen05.ld.32x32b.x32 (4 KB) per warp with nwarps (1..4) warps of a CTA reading. */
int dfb_debug_tmem_rate(int iters, int nwarps, int grid, double* cycles_per_ld);
/* Debug seam: the same under tensor-pipe load (a fifth warp issues mma_iters x 16 MMAs M=128 N=256 K=16 meanwhile):
 * cycles[0] = cycles per tcgen05.ld per warp, cycles[1] = cycles per MMA. */
int dfb_debug_tmem_rate_mma(int iters, int nwarps, int mma_iters, int grid, double* cycles);

int dfb_debug_umma_gemm(const float* A, const float* B, int N, int K, int kind, int variant, float* D, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Training step of the direct-feature-matching loop (row a15; reference
 * feature/direct_feature_matching.py:322-390 `train_on_batch`, `loss.backward()` at :378).
 * Gradients travel as NHWC bf16 with fp32 accumulation.
 * ---------------------------------------------------------------------------------------------- */

/* dfb_conv_create with explicit operand format (fmt 0 fp16 / 1 bf16).  dgrad != 0 builds the DATA-GRADIENT convolution
 * of the layer (Cin, Cout, weight [Cout,Cin,KH,KW], bn_scale): its input has Cout channels, its output round_up(Cin,64)
 * (replaces the conv backward of torch autograd w.r.t. the input). */
int dfb_conv_create_ex(int Cin, int Cout, int KH, int KW, const float* weight, const float* bias, const float* bn_scale,
                       const float* bn_shift, int fmt, int dgrad, DfbConv** out);
/* dfb_conv_fwd plus the backward epilogue: result zeroed where mask_nhwc16 <= 0 (ReLU'), then addend_nhwc16 added. */
int dfb_conv_fwd_ex(DfbConv* c, const void* in_nhwc16, int B, int H, int W, int relu, void* out_nhwc16, void* tap_nhwc16,
                    float* out_nchw32, const void* mask_nhwc16, const void* addend_nhwc16, void* stream);
/* dfb_conv_fwd_ex with a bf16 copy of the 16-bit output (out_bf16, NHWC like out_nhwc16; the operand type of
 * dfb_conv_wgrad), written by the same epilogue instead of a separate conversion kernel. */
int dfb_conv_fwd_ex2(DfbConv* conv, const void* in_nhwc16, int B, int H, int W, int relu, void* out_nhwc16, void* tap_nhwc16,
                     float* out_nchw32, const void* mask_nhwc16, const void* addend_nhwc16, void* out_bf16, void* stream);
/* Weight (and optional bias) gradient of a KHxKH convolution, stride 1, pad KH/2 (torch autograd conv backward w.r.t.
 * weight): gO NHWC [B,H,W,Cout], X NHWC [B,H,W,Cin_pad] 16-bit (fmt 0 f16 / 1 bf16) -> dW fp32 [Cout,Cin,KH,KH], dB [Cout]. */
int dfb_conv_wgrad(const void* gO, const void* X, int B, int H, int W, int Cin, int Cin_pad, int Cout, int KH, int fmt,
                   float* dW, float* dB, void* stream);
/* Same, ADDING to dW / dB (the caller zeroed them: one memset for a flat gradient buffer instead of two per layer). */
int dfb_conv_wgrad_acc(const void* gO, const void* X, int B, int H, int W, int Cin, int Cin_pad, int Cout, int KH, int fmt,
                   float* dW, float* dB, void* stream);

/* dfb_dfnet_load, flags bit0: also build the training variants (bf16 encoder, data-gradient convolutions); bit1: the
 * caller's work is ordered on the legacy default stream and the sources stay alive in stream order (no host
 * synchronisation); bit2: also build the train-mode BatchNorm variants of the heads (5x5 convs without the fold); bit3: the
 * heads' own training variants; bit4: leave the adaptation heads' images as they are (they are not going to be evaluated
 * with these weights: a pose regressor re-loaded after every optimizer step); bit5: likewise the bf16 encoder variant. */
int dfb_dfnet_load_ex(DfbDfnet* d, const float* const* params, const int64_t* numel, int n_params, float bn_eps,
                      uint32_t flags);
/* Batch statistics of the last forward with flags bit5: out [n_levels][2][128] = mean, biased variance per channel. */
int dfb_dfnet_bn_batch_stats(const DfbDfnet* d, float* out, void* stream);
/* Bytes of the tape a forward with flags bit3 writes (every activation kept) and of the backward scratch. */
int dfb_dfnet_tape_bytes(const DfbDfnet* d, int B, int H, int W, int upH, int upW, size_t* out);
/* Debug seam: byte offsets inside the tape (out[60]): in8; per encoder conv {act, pool or -1, h, w}; tap[3]; mid[3]; pooled. */
int dfb_debug_dfnet_tape_layout(const DfbDfnet* d, int B, int H, int W, int upH, int upW, int64_t* out);
int dfb_dfnet_bwd_workspace_bytes(const DfbDfnet* d, int B, int H, int W, size_t* out);
/* Backward of dfb_dfnet_fwd (flags as in that call: bit0 return_feature, bit1 single_stream, bit2 return_pose, bit3 tape,
 * bit4 bf16 encoder operands).
 *   g_feats_t / g_feats_r: gradients of the feature stacks [L,Bs,128,upH,upW]; a null stack skips that stream
 *   level_mask: bit l set = level l carries gradient;  g_pose [B,12] (nullable)
 *   g_x: gradient w.r.t. the images of the differentiated sub-batch, fp32 [nb,3,H,W] (nullable)
 *   g_params: n_params pointers in dfb_dfnet_load order (nullable entries); encoder and fc_pose entries are written. */
int dfb_dfnet_bwd(DfbDfnet* d, int B, int H, int W, uint32_t flags, int upH, int upW, const float* g_feats_t,
                  const float* g_feats_r, uint32_t level_mask, const float* g_pose, const void* tape, float* g_x,
                  float* const* g_params, int n_params, void* scratch, size_t scratch_bytes, void* stream);

/* Data-parallel training ("one NCCL allreduce ... overlapped as the tail of F's backward"): from now on
 * dfb_dfnet_bwd records `event` (a cudaEvent_t) on its stream as soon as the gradients of fc_pose and of the encoder
 * layers >= first_layer (0..12, VGG order) are complete.  The backward walks the encoder from conv5_3 down to conv1_1,
 * so with first_layer = 7 (conv4_1) 88 % of the gradient bytes are ready while 90 % of the backward's work is still to
 * run; the caller all-reduces g_params[2*first_layer ..] on another stream behind this event.  NULL switches it off. */
int dfb_dfnet_bwd_bucket_event(DfbDfnet* d, int first_layer, void* event);

/* Backward of dfb_cosine_loss w.r.t. fr (g_loss: device scalar), of dfb_mse w.r.t. a, and the adjoints of the two
 * resampling operators (outputs overwritten).  dfb_cosine_loss_bwd: per_channel bit 1 (value 2) = `ws` is the workspace
 * dfb_cosine_loss was called with for the same fr / ft and still holds its row statistics (they are not recomputed). */
int dfb_cosine_loss_bwd(const float* fr, const float* ft, int C, int64_t HW, int per_channel, float eps, const float* g_loss,
                        float* g_fr, void* ws, size_t ws_bytes, void* stream);
int dfb_mse_bwd(const float* a, const float* b, int64_t n, const float* g_loss, float* g_a, void* stream);
int dfb_resize_bicubic_bwd(const float* g_dst, int64_t planes, int h, int w, int Ho, int Wo, float* g_src, void* stream);
int dfb_resize_bilinear_ac_bwd(const float* g_dst, int64_t planes, int h, int w, int Ho, int Wo, float* g_src, void* stream);

/* Debug seam (tools/conv_prof.py): per-CTA cycle counters of the convolution kernel's roles, see csrc/conv_tc.cu. */
int dfb_debug_conv_prof(int on, unsigned long long* out_host, int max_cta, int* grid);

/* Debug seam: one tcgen05 tile with MN-major operands, D[128,N] = sum_k A[k][m] * B[k][n] (A [K,128], B [K,N] fp32 on
 * the device, rounded to fmt_a / fmt_b: 0 = f16, 1 = bf16).  Pins the descriptor convention the weight-gradient
 * kernel relies on. */
int dfb_debug_umma_gemm_mn(const float* A, const float* B, int N, int K, int fmt_a, int fmt_b, int variant, float* D,
                           void* stream);

/* ------------------------------------------------------------------------------------------------
 * NeRF-Hist training step (run_nerf.py:32-80, models/losses.py:19-57).  The MLP layers run as 1x1
 * convolutions over the P = N*S samples on the dfb_conv_* kernels (forward, data gradient with ReLU mask, weight
 * gradient); these are the NeRF-specific pieces around them (all device pointers, see csrc/nerf_train.cu).
 * ---------------------------------------------------------------------------------------------- */
/* Re-pack an existing convolution handle from new fp32 parameters (after an optimizer step). */
int dfb_conv_update(DfbConv* conv, const float* weight, const float* bias, const float* bn_scale, const float* bn_shift, void* stream);
/* Bracket for re-loading many convolutions in a row: between _begin and _end every dfb_conv_update of device-resident
 * tensors is queued and all of them go out as ONE packing launch on `stream` (the sources must stay unchanged until then). */
int dfb_conv_pack_begin(void);
int dfb_conv_pack_end(void* stream);
/* dfb_conv_pack_begin, dfb_conv_update(convs[i], weights[i], biases[i], NULL, NULL) for i < n, dfb_conv_pack_end in one call. */
int dfb_conv_update_many(DfbConv* const* convs, const float* const* weights, const float* const* biases, int n, void* stream);
/* pts = o + d*z, positional encoding with L bands (nerfw.py:105-133) -> fp16 [N*S, ld], columns >= 3+6L zero.
 * rays: [N, ray_stride] with o at 0..2 and d at 3..5. */
int dfb_embed_xyz16(const float* rays, int ray_stride, const float* z, int64_t N, int S, int L, int ld, void* out, void* stream);
/* the same with a bf16 copy of the output (out_bf16, nullable): the first layer's weight-gradient operand */
int dfb_embed_xyz16_ex(const float* rays, int ray_stride, const float* z, int64_t N, int S, int L, int ld, void* out, void* out_bf16,
                       void* stream);
/* out[p, :] = fp16(rb[p / S, :]) for p < N*S (C % 8 == 0), and its adjoint out[r, :] = sum_s g[r*S + s, :] (g bf16). */
int dfb_rows_expand16(const float* rb, int64_t N, int S, int C, void* out, void* stream);
int dfb_rows_reduce_bf16(const void* g, int64_t N, int S, int C, float* out, void* stream);
/* Heads (nerfw.py:275-295): fp32 pre-activation planes [channel][P] (plane stride given) -> raw [P,C], C = 4 (static rgb,
 * sigma) or 9 (+ transient rgb, sigma, beta; tr_pre planes in that order).  Backward: d raw -> d pre-activation as bf16
 * [P,64] operands (zero padded) for the head layers' data / weight gradient convolutions. */
int dfb_nerf_heads_fwd(const float* sig_pre, const float* rgb_pre, const float* tr_pre, int64_t P, int64_t plane_stride, int C,
                       float* raw, void* stream);
int dfb_nerf_heads_bwd(const float* raw, const float* g_raw, int64_t P, int C, void* g_sig16, void* g_rgb16, void* g_tr16,
                       void* stream);
/* Adjoint of raw2outputs_NeRFW (rendering.py:132-243) in train mode w.r.t. raw [N,S,C] for the outputs NerfWLoss reads:
 * C = 9: g_rgb [N,3], g_beta [N], g_tsig [N,S] (any may be NULL); C = 4: g_rgb (+ the coarse pass' noise draws). */
int dfb_raw2outputs_bwd(const float* raw, const float* z_vals, int64_t N, int S, int C, const float* noise, float raw_noise_std,
                        const float* g_rgb, const float* g_beta, const float* g_tsig, float* g_raw, void* stream);
/* NerfWLoss (models/losses.py:42-57) in one pass: out5 = {c_l, f_l, b_l, s_l, mean((rgb_fine - targets)^2)} with
 * c_l = coef 0.5 mean((rgb_coarse - t)^2), f_l = coef mean((rgb_fine - t)^2 / (2 beta^2)), b_l = coef (3 + mean(log beta)),
 * s_l = coef lambda_u mean(transient_sigmas).  rgb [N,3], beta [N], transient_sigmas [N,S], ws: dfb_nerfw_loss_workspace_bytes().
 * _bwd: g_c..g_s = upstream gradients of the four terms (device scalars, NULL = 0); every output is nullable. */
size_t dfb_nerfw_loss_workspace_bytes(void);
int dfb_nerfw_loss_fwd(const float* rgb_coarse, const float* rgb_fine, const float* beta, const float* transient_sigmas,
                       const float* targets, int64_t N, int S, float coef, float lambda_u, void* ws, float* out5, void* stream);
int dfb_nerfw_loss_bwd(const float* rgb_coarse, const float* rgb_fine, const float* beta, const float* targets, int64_t N, int S,
                       float coef, float lambda_u, const float* g_c, const float* g_f, const float* g_b, const float* g_s,
                       float* g_rgb_coarse, float* g_rgb_fine, float* g_beta, float* g_transient_sigmas, void* stream);
int dfb_cast_f16_bf16(const void* src, void* dst, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Data side and evaluation around the hot path
 * ---------------------------------------------------------------------------------------------- */
/* Luma histogram of the loaders (dataset_loaders/seven_scenes.py:346-352 with utils/color.py:29-35): img [B,3,H,W] fp32
 * in [0,1] -> hist [B,bins] = round(histc(0.299 r + 0.587 g + 0.114 b, bins, 0, 1) / count * 100), the integer-valued
 * percentages that index embedding_a / embedding_t.  ws: B*bins*4 bytes. */
int dfb_luma_hist(const float* img, int B, int H, int W, int bins, float* hist, void* ws, size_t ws_bytes, void* stream);
/* cv2.resize(img, (w, h), interpolation=cv2.INTER_AREA) of an HWC fp32 image, downscaling only
 * (dataset_loaders/seven_scenes.py:328-332, feature/direct_feature_matching.py:149). */
int dfb_resize_area(const float* src, int H, int W, int C, int h, int w, float* dst, void* stream);
/* compute_error_in_q (feature/misc.py:49-107) for n pose pairs at once: pred, gt [n,12] row-major 3x4; use_svd: the
 * predicted rotation is replaced by U V^T first (:70-77).  out [n,2] = {|t_gt - t_pred|, angle between the rotations
 * in degrees via quaternions (pytorch3d 0.3.0 matrix_to_quaternion)}; pred_fixed [n,12] (nullable): the predicted
 * pose with the orthogonalised rotation. */
int dfb_pose_error(const float* pred, const float* gt, int n, int use_svd, float* out, float* pred_fixed, void* stream);
/* svd_reg of the pose regressor (feature/direct_feature_matching.py:81-86: u, s, v = torch.svd(R); R <- u v^T): the
 * orthogonal polar factor of n 3x3 matrices A [n,9] -> Q [n,9] on the device (no host synchronisation; torch.svd checks
 * its status on the host), and its adjoint G [n,9] -> dA [n,9].  aux [n,21] doubles = U | V | s, written by the forward. */
int dfb_polar3x3_fwd(const float* A, int n, float* Q, double* aux, void* stream);
int dfb_polar3x3_bwd(const double* aux, const float* G, int n, float* dA, void* stream);
/* n strided fp32 copies dst[r*dst_ld + c] = src[r*src_ld + c] (r < rows, c < cols) in one launch: the parameter ->
 * padded staging step of the NeRF-W training executor (what `param.data.copy_` does per tensor in the reference's
 * optimizer loop has no counterpart there; this replaces the host-side python copies of dfnet_b200/nerf_train.py). */
typedef struct DfbCopy2d {
  const float* src;
  float* dst;
  int rows, cols, src_ld, dst_ld;
} DfbCopy2d;
int dfb_copy2d_batch(const DfbCopy2d* items, int n, void* stream);
/* get_rays (models/ray_utils.py:5-15) with viewdirs = rays_d / |rays_d| (rendering.py:366-370) for a pose that carries
 * gradient, and the adjoint: rays_o / rays_d / viewdirs [H*W,3]; g_c2w12 = d loss / d c2w[:3,:4] (row-major) from the
 * gradients of the three outputs (each nullable).  ws: dfb_pose_rays_workspace_bytes(). */
int dfb_pose_rays_fwd(const float* c2w, int row_stride, int H, int W, float focal, float* rays_o, float* rays_d, float* viewdirs,
                      void* stream);
size_t dfb_pose_rays_workspace_bytes(void);
int dfb_pose_rays_bwd(const float* c2w, int row_stride, int H, int W, float focal, const float* g_rays_o, const float* g_rays_d,
                      const float* g_viewdirs, void* ws, float* g_c2w12, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DFNET_B200_H_ */
