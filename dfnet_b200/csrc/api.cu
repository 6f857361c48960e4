// C ABI entry points for rendering (see include/dfnet_b200.h).  Orchestrates the kernels of
// render_kernels.cu / mlp_simt.cu / mlp_tc.cu the way models/rendering.py:245-400 chains
// render -> batchify_rays -> render_rays -> {network_query_fn, raw2outputs_NeRFW, sample_pdf}.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

using namespace dfb;

namespace {

constexpr int64_t kChunkRays = 1 << 16;  // internal batchify (bounds the workspace, not a tuning knob of the ABI)

bool fuse_composite() {  // read per call so that tests can compare both paths in one process
  const char* e = getenv("DFB_TC_FUSE_COMPOSITE");
  return !(e && e[0] == '0');
}

// operand kind of everything but the coarse sigma-only pass
int eff_kind(int k) { return k == DFB_MMA_F16_SPLIT_COARSE ? DFB_MMA_F16 : k; }

struct WsLayout {
  size_t rayrec, extra, z_c, raw_c, rb_c, w_c, z_s, z_all, raw_f, rb_f, lin, total;
};

size_t align256(size_t x) { return (x + 255) / 256 * 256; }

WsLayout ws_layout(const DfbNerf* n, const DfbRenderCfg* c, int64_t rays) {
  const DfbNerfDesc& d = n->desc;
  const int Nc = c->N_samples, Nf = c->N_importance, S = Nc + Nf, W = d.W;
  const int n_extra = 27 + d.a_dim + d.t_dim;
  const int Cc = c->test_time ? 1 : 4;
  WsLayout L;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return o; };
  L.lin = take((size_t)(Nc + std::max(Nf, 1)) * 4);
  L.rayrec = take((size_t)rays * kRayRec * 4);
  L.extra = take((size_t)rays * n_extra * 4);
  L.z_c = take((size_t)rays * Nc * 4);
  L.raw_c = take((size_t)rays * Nc * Cc * 4);
  L.rb_c = take((size_t)rays * (W / 2) * 4);
  L.w_c = take((size_t)rays * Nc * 4);
  L.z_s = take((size_t)rays * std::max(Nf, 1) * 4);
  L.z_all = take((size_t)rays * S * 4);
  L.raw_f = take(Nf > 0 ? (size_t)rays * S * 9 * 4 : 0);
  L.rb_f = take(Nf > 0 ? (size_t)rays * std::max(W, 256) * 4 : 0);  // 256-wide rows on the tcgen05 path
  L.total = off;
  return L;
}

int check_cfg(const DfbNerf* n, const DfbRenderCfg* c) {
  DFB_REQUIRE(n && c, DFB_ERR_INVALID, "null handle or config");
  DFB_REQUIRE(c->N_samples >= 4 && c->N_samples <= 1024, DFB_ERR_INVALID, "N_samples %d out of range", c->N_samples);
  DFB_REQUIRE(c->N_importance >= 0 && c->N_importance <= 1024, DFB_ERR_INVALID, "N_importance out of range");
  DFB_REQUIRE(!(c->N_importance == 0 && c->test_time), DFB_ERR_INVALID,
              "N_importance == 0 with test_time yields rgb_map = None in the reference (rendering.py:188-193)");
  DFB_REQUIRE(c->N_importance == 0 || n->desc.has_fine, DFB_ERR_INVALID, "N_importance > 0 needs network_fine");
  DFB_REQUIRE(n->net[0].loaded && (c->N_importance == 0 || n->net[1].loaded), DFB_ERR_INVALID, "parameters not loaded");
  DFB_REQUIRE(c->N_importance == 0 || n->has_emb, DFB_ERR_INVALID, "embedding_a / embedding_t not set");
  DFB_REQUIRE(c->mma_kind >= 0 && c->mma_kind <= 3, DFB_ERR_INVALID, "unknown mma_kind %d", c->mma_kind);
  return DFB_OK;
}

// ---- optional per-kernel timing of the two MLP launches (bench.py roofline) ----------------
struct ProfRec { cudaEvent_t a, b; int which; };
bool g_prof_on = false;
std::vector<ProfRec> g_prof;

int run_mlp_inner(const DfbNerf* n, const DfbRenderCfg* c, int which, int mode, const float* rayrec, const float* z,
                  const float* rb, int64_t rays, int S, float* raw, cudaStream_t st, uint32_t* masks, float* part,
                  const int* ert_rowmap = nullptr, const int* ert_offsets = nullptr) {
  // The tcgen05 kernel covers the 8x256 networks (sigma-only coarse pass and full fine pass); every
  // other shape or mode (other widths, the train-mode coarse pass) runs on the fp32 CUDA kernel.
  if (c->mma_kind != DFB_MMA_FP32_SIMT && tc_supported(n, which, mode))
    return launch_mlp_tc_rays(n, which, mode, eff_kind(c->mma_kind), rayrec, z, rb, rays, S, raw, st, masks,
                              c->mma_kind == DFB_MMA_F16_SPLIT_COARSE && which == 0 && mode == MLP_SIGMA, part,
                              part ? composite_part_k(S) : 0, ert_rowmap, ert_offsets);
  DFB_REQUIRE(!part, DFB_ERR_UNSUPPORTED, "fused compositing is a mode of the tcgen05 fine pass");
  DFB_REQUIRE(!masks, DFB_ERR_UNSUPPORTED, "relu_masks are an output of the tcgen05 path (8x256 fine network, mma f16 / bf16)");
  return launch_mlp_simt_rays(n, which, mode, rayrec, z, rb, rays, S, raw, st);
}

int run_mlp(const DfbNerf* n, const DfbRenderCfg* c, int which, int mode, const float* rayrec, const float* z,
            const float* rb, int64_t rays, int S, float* raw, cudaStream_t st, uint32_t* masks = nullptr,
            float* part = nullptr, const int* ert_rowmap = nullptr, const int* ert_offsets = nullptr) {
  if (!g_prof_on) return run_mlp_inner(n, c, which, mode, rayrec, z, rb, rays, S, raw, st, masks, part, ert_rowmap, ert_offsets);
  ProfRec r;
  r.which = which;
  DFB_CHECK_CUDA(cudaEventCreate(&r.a));
  DFB_CHECK_CUDA(cudaEventCreate(&r.b));
  DFB_CHECK_CUDA(cudaEventRecord(r.a, st));
  int rc = run_mlp_inner(n, c, which, mode, rayrec, z, rb, rays, S, raw, st, masks, part, ert_rowmap, ert_offsets);
  DFB_CHECK_CUDA(cudaEventRecord(r.b, st));
  g_prof.push_back(r);
  return rc;
}

}  // namespace

extern "C" int dfb_render_workspace_bytes(const DfbNerf* n, const DfbRenderCfg* c, int64_t n_rays, size_t* out) {
  int rc = check_cfg(n, c);
  if (rc) return rc;
  DFB_REQUIRE(out && n_rays >= 0, DFB_ERR_INVALID, "bad arguments");
  *out = ws_layout(n, c, std::min<int64_t>(std::max<int64_t>(n_rays, 1), kChunkRays)).total;
  return DFB_OK;
}

static int render_fwd_impl(DfbNerf* n, const DfbRenderCfg* c, const float* rays, const float* c2w, int n_pose, int H, int W,
                           float focal, float near, float far, const float* hist, int64_t N, const float* t_rand,
                           const float* u, const float* noise, float* rgb, float* disp, float* acc,
                           const DfbRenderExtras* ex, void* ws, size_t ws_bytes, void* stream);

extern "C" int dfb_render_fwd(DfbNerf* n, const DfbRenderCfg* c, const float* rays, const float* c2w, int H, int W,
                              float focal, float near, float far, const float* hist, int64_t N, const float* t_rand,
                              const float* u, const float* noise, float* rgb, float* disp, float* acc,
                              const DfbRenderExtras* ex, void* ws, size_t ws_bytes, void* stream) {
  return render_fwd_impl(n, c, rays, c2w, 1, H, W, focal, near, far, hist, N, t_rand, u, noise, rgb, disp, acc, ex, ws, ws_bytes, stream);
}

// Batched multi-pose render (random view synthesis, feature/misc.py:249-289 renders hundreds of small views one by one):
// n_pose poses [n_pose,3,4] with their histograms [n_pose,hist_bin], one H x W image each, rays image-major.
extern "C" int dfb_render_poses_fwd(DfbNerf* n, const DfbRenderCfg* c, const float* c2w, int n_pose, int H, int W, float focal,
                                    float near, float far, const float* hist, float* rgb, float* disp, float* acc, void* ws,
                                    size_t ws_bytes, void* stream) {
  DFB_REQUIRE(c2w && hist && n_pose >= 1, DFB_ERR_INVALID, "dfb_render_poses_fwd: bad arguments");
  return render_fwd_impl(n, c, nullptr, c2w, n_pose, H, W, focal, near, far, hist, (int64_t)n_pose * H * W, nullptr, nullptr, nullptr,
                         rgb, disp, acc, nullptr, ws, ws_bytes, stream);
}

static int render_fwd_impl(DfbNerf* n, const DfbRenderCfg* c, const float* rays, const float* c2w, int n_pose, int H, int W,
                           float focal, float near, float far, const float* hist, int64_t N, const float* t_rand,
                           const float* u, const float* noise, float* rgb, float* disp, float* acc,
                           const DfbRenderExtras* ex, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_cfg(n, c);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  DFB_REQUIRE((rays != nullptr) != (c2w != nullptr), DFB_ERR_INVALID, "pass exactly one of rays / c2w");
  DFB_REQUIRE(!c2w || ((int64_t)H * W * n_pose == N && hist), DFB_ERR_INVALID, "c2w mode needs N == n_pose*H*W and a histogram");
  DFB_REQUIRE(rgb && disp && acc, DFB_ERR_INVALID, "rgb/disp/acc outputs are required");
  const DfbNerfDesc& d = n->desc;
  // the kernels index rays as r*(11+hist_bin) and hist[0..hist_bin): a record of any other width would be read
  // misaligned / out of bounds (the reference raises a shape error in torch.cat / Embedding instead)
  DFB_REQUIRE(!rays || c->ray_stride == 11 + d.hist_bin, DFB_ERR_INVALID,
              "ray records are %d floats wide, expected 11 + hist_bin = %d ([o3,d3,near,far,viewdir3,hist])", c->ray_stride,
              11 + d.hist_bin);
  DFB_REQUIRE(!c2w || c->hist_len == d.hist_bin, DFB_ERR_INVALID, "histogram has %d entries, expected hist_bin = %d",
              c->hist_len, d.hist_bin);
  const int Nc = c->N_samples, Nf = c->N_importance, S = Nc + Nf;
  DFB_REQUIRE(!c->perturb || (t_rand && (Nf == 0 || u)), DFB_ERR_INVALID,
              "perturb > 0 needs the uniform draws t_rand [N,Nc] and u [N,Nf] (rendering.py:282,36)");
  DFB_REQUIRE(c->raw_noise_std == 0.f || noise, DFB_ERR_INVALID,
              "raw_noise_std != 0 needs the standard-normal draws noise [N,Nc] (rendering.py:173)");
  if (N == 0) return DFB_OK;
  DFB_CHECK_CUDA(cudaSetDevice(n->device));
  const int64_t chunk = std::min<int64_t>(N, kChunkRays);
  const WsLayout L = ws_layout(n, c, chunk);
  DFB_REQUIRE(ws && ws_bytes >= L.total, DFB_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", L.total, ws_bytes);
  char* base = (char*)ws;
  auto P = [&](size_t off) { return (float*)(base + off); };

  // t_vals / u grids exactly as torch.linspace builds them (host, ATen order); cached on the handle.
  if (n->lin_nc != Nc || n->lin_nf != Nf) {
    std::vector<float> lin(Nc + std::max(Nf, 1));
    dfb_linspace_f32(0.f, 1.f, Nc, lin.data());
    if (Nf > 0) dfb_linspace_f32(0.f, 1.f, Nf, lin.data() + Nc);
    if (n->lin_dev) cudaFree(n->lin_dev);
    n->lin_dev = nullptr;
    DFB_CHECK_CUDA(cudaMalloc(&n->lin_dev, lin.size() * 4));
    DFB_CHECK_CUDA(cudaMemcpy(n->lin_dev, lin.data(), lin.size() * 4, cudaMemcpyHostToDevice));
    n->lin_nc = Nc, n->lin_nf = Nf;
  }
  const float* d_lin = n->lin_dev;

  const int n_extra = 27 + d.a_dim + d.t_dim;
  const bool train = !c->test_time;
  const int Cc = train ? 4 : 1;
  const int hb = d.hist_bin;

  for (int64_t r0 = 0; r0 < N; r0 += chunk) {
    const int64_t nr = std::min<int64_t>(chunk, N - r0);
    PrepArgs pa = {};
    pa.rays = rays ? rays + r0 * (11 + hb) : nullptr;
    pa.c2w = c2w, pa.c2w_ld = 4, pa.n_pose = n_pose, pa.H = H, pa.W = W, pa.focal = focal, pa.near = near, pa.far = far, pa.hist = hist;
    pa.hb = hb, pa.n_vocab = d.n_vocab, pa.emb_a = n->emb_a, pa.emb_t = n->emb_t;
    pa.N = nr, pa.Nc = Nc, pa.t_vals = d_lin, pa.t_rand = c->perturb ? t_rand + r0 * Nc : nullptr;
    pa.lindisp = c->lindisp, pa.rayrec = P(L.rayrec), pa.z = P(L.z_c);
    const bool need_extra = train || Nf > 0;
    pa.extra = need_extra ? P(L.extra) : nullptr;
    pa.n_extra = n_extra, pa.a_dim = d.a_dim, pa.t_dim = d.t_dim;
    if (need_extra && !n->has_emb) {
      // coarse-only train mode never reads the embeddings; give the kernel zeros via hist idx 0
      DFB_REQUIRE(Nf == 0, DFB_ERR_INVALID, "embeddings not set");
      pa.a_dim = 0, pa.t_dim = 0, pa.n_extra = 27;
    }
    pa.pix0 = r0;  // c2w mode: global pixel index of this chunk's first ray
    rc = launch_prep(pa, st);
    if (rc) return rc;

    const float* rayrec = P(L.rayrec);
    // ---- coarse network (rendering.py:289-295) ---------------------------------------------
    float* raw_c = P(L.raw_c);
    float* rb_c = nullptr;
    if (train) {
      rb_c = P(L.rb_c);
      rc = launch_raybias(pa.extra, pa.n_extra, nr, n->net[0], false, rb_c, n->net[0].n_dt, st);
      if (rc) return rc;
    }
    rc = run_mlp(n, c, 0, train ? MLP_STATIC : MLP_SIGMA, rayrec, P(L.z_c), rb_c, nr, Nc, raw_c, st);
    if (rc) return rc;
    CompositeArgs ca = {};
    ca.raw = raw_c, ca.z = P(L.z_c), ca.N = nr, ca.S = Nc, ca.C = Cc, ca.typ_fine = 0, ca.test_time = c->test_time;
    ca.beta_min = d.beta_min;
    ca.weights = P(L.w_c);
    if (c->raw_noise_std != 0.f) ca.noise = noise + r0 * Nc, ca.noise_std = c->raw_noise_std;
    if (Nf == 0) {
      ca.rgb = rgb + r0 * 3, ca.disp = disp + r0, ca.acc = acc + r0;
      ca.depth = ex && ex->depth ? ex->depth + r0 : nullptr;
    } else if (train && ex) {
      ca.rgb = ex->rgb0 ? ex->rgb0 + r0 * 3 : nullptr;
      ca.disp = ex->disp0 ? ex->disp0 + r0 : nullptr;
      ca.acc = ex->acc0 ? ex->acc0 + r0 : nullptr;
    }
    rc = launch_composite(ca, st);
    if (rc) return rc;
    if (ex && ex->weights_coarse)
      DFB_CHECK_CUDA(cudaMemcpyAsync(ex->weights_coarse + r0 * Nc, P(L.w_c), (size_t)nr * Nc * 4, cudaMemcpyDeviceToDevice, st));
    if (Nf == 0) {
      if (ex && ex->raw)
        DFB_CHECK_CUDA(cudaMemcpyAsync(ex->raw + r0 * Nc * Cc, raw_c, (size_t)nr * Nc * Cc * 4, cudaMemcpyDeviceToDevice, st));
      if (ex && ex->z_vals)
        DFB_CHECK_CUDA(cudaMemcpyAsync(ex->z_vals + r0 * Nc, P(L.z_c), (size_t)nr * Nc * 4, cudaMemcpyDeviceToDevice, st));
      continue;
    }
    // ---- hierarchical sampling (rendering.py:300-305) ----------------------------------------
    SampleArgs sa = {};
    sa.z_c = P(L.z_c), sa.w_c = P(L.w_c), sa.Nc = Nc, sa.N = nr, sa.Nf = Nf;
    sa.u = c->perturb ? u + r0 * Nf : nullptr, sa.u_lin = d_lin + Nc;
    sa.samples = ex && ex->z_samples ? ex->z_samples + r0 * Nf : nullptr;
    sa.inds = ex && ex->inds ? ex->inds + r0 * Nf : nullptr;
    sa.z_vals = ex && ex->z_vals ? ex->z_vals + r0 * S : P(L.z_all);
    sa.z_std = train && ex && ex->z_std ? ex->z_std + r0 : nullptr;
    // opt-in early ray termination (test-time fused path only): workspace carved from the unused raw buffer, behind the
    // segment records: n_live [nr], offsets [nr+1], rowmap [nr*S]
    const bool tc_fine = c->mma_kind != DFB_MMA_FP32_SIMT && tc_supported(n, 1, MLP_FULL);
    const bool fuse_ok = tc_fine && c->test_time && fuse_composite() &&
                         !(ex && (ex->raw || ex->depth || ex->beta || ex->transient_sigmas || ex->relu_masks));
    DFB_REQUIRE(c->ert_eps == 0.f || (fuse_ok && c->ert_eps > 0.f && c->ert_eps < 1.f), DFB_ERR_INVALID,
                "ert_eps (early ray termination) needs the fused test-time tensor-core path (test_time, mma f16 / bf16, no raw / "
                "depth / beta extras) and 0 < ert_eps < 1");
    int *ert_nlive = nullptr, *ert_off = nullptr, *ert_map = nullptr;
    if (c->ert_eps > 0.f) {
      char* eb = base + L.raw_f + align256((size_t)nr * composite_part_k(S) * 32);
      ert_nlive = (int*)eb, ert_off = (int*)(eb + align256((size_t)nr * 4)), ert_map = (int*)(eb + 2 * align256((size_t)(nr + 1) * 4));
      sa.ert_eps = c->ert_eps, sa.n_live = ert_nlive;
    }
    rc = launch_sample(sa, st);
    if (rc) return rc;
    if (ert_nlive) {
      rc = launch_ert_compact(ert_nlive, nr, S, ert_off, ert_map, st);
      if (rc) return rc;
    }
    const float* z_all = sa.z_vals;
    // ---- fine network (rendering.py:307-316) -------------------------------------------------
    float* rb_f = P(L.rb_f);
    // tcgen05 path: the per-ray bias carries the step's constant bias and is stored as packed 16-bit pairs
    const bool tc_f = c->mma_kind != DFB_MMA_FP32_SIMT && tc_supported(n, 1, MLP_FULL);
    // (the native 128-wide program reads the two halves contiguously, the padded embedding at columns 0 / 128)
    const bool nat_f = tc_f && tc_native128(n, 1, ex && ex->relu_masks, false);
    rc = launch_raybias(pa.extra, pa.n_extra, nr, n->net[1], true, rb_f, tc_f ? 256 : n->net[1].n_dt, st,
                        tc_f ? (nat_f ? n->net[1].tc_dtbias_n_dev : n->net[1].tc_dtbias_dev) : nullptr,
                        tc_f ? (eff_kind(c->mma_kind) == DFB_MMA_F16 ? 1 : 2) : 0, tc_f && !nat_f ? 128 : 0);
    if (rc) return rc;
    float* raw_f = ex && ex->raw ? ex->raw + r0 * S * 9 : P(L.raw_f);
    uint32_t* masks = nullptr;
    if (ex && ex->relu_masks) {
      DFB_REQUIRE((r0 * S) % 128 == 0, DFB_ERR_INVALID, "relu_masks: chunk start not aligned to a 128-sample tile");
      masks = ex->relu_masks + (size_t)(r0 * S / 128) * kReluMaskWordsPerTile;
    }
    // Test-time render_path step on the tcgen05 path with nothing but rgb / disp / acc wanted: compositing is fused
    // into the heads epilogue of the fine kernel (no [P,9] raw tensor; one 32-byte record per 32 samples instead) and
    // k_composite_partials chains the records.  DFB_TC_FUSE_COMPOSITE=0 keeps the raw round trip (A/B measurement).
    const bool fuse = tc_f && c->test_time && fuse_composite() &&
                      !(ex && (ex->raw || ex->depth || ex->beta || ex->transient_sigmas || ex->relu_masks));
    if (fuse) {
      float* part = P(L.raw_f);  // the raw buffer's space: rays * part_k * 32 B << rays * S * 36 B
      rc = run_mlp(n, c, 1, MLP_FULL, rayrec, z_all, rb_f, nr, S, nullptr, st, nullptr, part, ert_map, ert_off);
      if (rc) return rc;
      if (ex && ex->n_live && ert_nlive)
        DFB_CHECK_CUDA(cudaMemcpyAsync(ex->n_live + r0, ert_nlive, (size_t)nr * 4, cudaMemcpyDeviceToDevice, st));
      rc = launch_composite_partials(part, composite_part_k(S), nr, S, rgb + r0 * 3, disp + r0, acc + r0, st, ert_off);
      if (rc) return rc;
      continue;
    }
    rc = run_mlp(n, c, 1, MLP_FULL, rayrec, z_all, rb_f, nr, S, raw_f, st, masks);
    if (rc) return rc;
    CompositeArgs cf = {};
    cf.raw = raw_f, cf.z = z_all, cf.N = nr, cf.S = S, cf.C = 9, cf.typ_fine = 1, cf.test_time = c->test_time;
    cf.beta_min = d.beta_min;
    cf.rgb = rgb + r0 * 3, cf.disp = disp + r0, cf.acc = acc + r0;
    if (ex) {
      cf.depth = ex->depth ? ex->depth + r0 : nullptr;
      cf.beta = ex->beta ? ex->beta + r0 : nullptr;
      cf.tsig = ex->transient_sigmas ? ex->transient_sigmas + r0 * S : nullptr;
    }
    rc = launch_composite(cf, st);
    if (rc) return rc;
  }
  return DFB_OK;
}

extern "C" int dfb_render_image_host(DfbNerf* n, const DfbRenderCfg* c, const float* c2w_host, int H, int W, float focal,
                                     float near, float far, const float* hist_host, float* rgb_host, float* disp_host,
                                     float* acc_host, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_cfg(n, c);
  if (rc) return rc;
  DFB_REQUIRE(c2w_host && hist_host && rgb_host && disp_host && acc_host, DFB_ERR_INVALID, "null host buffer");
  DFB_REQUIRE(c->hist_len == n->desc.hist_bin, DFB_ERR_INVALID, "histogram has %d entries, expected hist_bin = %d", c->hist_len,
              n->desc.hist_bin);
  DFB_REQUIRE(!c->perturb && c->raw_noise_std == 0.f, DFB_ERR_INVALID, "dfb_render_image_host renders with render_kwargs_test (no perturb, no noise)");
  cudaStream_t st = (cudaStream_t)stream;
  DFB_CHECK_CUDA(cudaSetDevice(n->device));
  const int64_t N = (int64_t)H * W;
  // the image-sized staging lives at the tail of the caller's workspace
  size_t need = 0;
  rc = dfb_render_workspace_bytes(n, c, N, &need);
  if (rc) return rc;
  const size_t stage = align256(16 * 4) + align256(n->desc.hist_bin * 4) + align256((size_t)N * 5 * 4);
  DFB_REQUIRE(ws && ws_bytes >= need + stage, DFB_ERR_WORKSPACE,
              "workspace too small: dfb_render_image_host needs %zu bytes (render %zu + staging %zu)", need + stage, need, stage);
  char* tail = (char*)ws + need;
  float* d_c2w = (float*)tail;
  float* d_hist = (float*)(tail + align256(16 * 4));
  float* d_out = (float*)(tail + align256(16 * 4) + align256(n->desc.hist_bin * 4));
  DFB_CHECK_CUDA(cudaMemcpyAsync(d_c2w, c2w_host, 12 * 4, cudaMemcpyHostToDevice, st));
  DFB_CHECK_CUDA(cudaMemcpyAsync(d_hist, hist_host, n->desc.hist_bin * 4, cudaMemcpyHostToDevice, st));
  rc = dfb_render_fwd(n, c, nullptr, d_c2w, H, W, focal, near, far, d_hist, N, nullptr, nullptr, nullptr, d_out,
                      d_out + 3 * N, d_out + 4 * N, nullptr, ws, need, stream);
  if (rc) return rc;
  DFB_CHECK_CUDA(cudaMemcpyAsync(rgb_host, d_out, (size_t)N * 3 * 4, cudaMemcpyDeviceToHost, st));
  DFB_CHECK_CUDA(cudaMemcpyAsync(disp_host, d_out + 3 * N, (size_t)N * 4, cudaMemcpyDeviceToHost, st));
  DFB_CHECK_CUDA(cudaMemcpyAsync(acc_host, d_out + 4 * N, (size_t)N * 4, cudaMemcpyDeviceToHost, st));
  return DFB_OK;
}

extern "C" int dfb_sample_pdf(const float* bins, const float* weights, const float* u, int64_t N, int n_bins, int Nf,
                              float* samples, int32_t* inds, void* stream) {
  if (N == 0) return DFB_OK;
  DFB_REQUIRE(bins && weights && (samples || inds), DFB_ERR_INVALID, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  float* d_lin = nullptr;
  if (!u) {
    std::vector<float> lin(Nf);
    dfb_linspace_f32(0.f, 1.f, Nf, lin.data());
    DFB_CHECK_CUDA(cudaMallocAsync(&d_lin, Nf * 4, st));
    DFB_CHECK_CUDA(cudaMemcpyAsync(d_lin, lin.data(), Nf * 4, cudaMemcpyHostToDevice, st));
    DFB_CHECK_CUDA(cudaStreamSynchronize(st));
  }
  SampleArgs sa = {};
  sa.bins = bins, sa.weights = weights, sa.nb = n_bins, sa.u = u, sa.u_lin = d_lin, sa.N = N, sa.Nf = Nf;
  sa.samples = samples, sa.inds = inds;
  int rc = launch_sample(sa, st);
  if (d_lin) cudaFreeAsync(d_lin, st);
  return rc;
}

extern "C" int dfb_raw2outputs(const float* raw, const float* z_vals, int64_t N, int S, int C, int typ, int test_time,
                               float beta_min, float* rgb, float* disp, float* acc, float* weights, float* depth,
                               float* transient_sigmas, float* beta, const float* noise, float raw_noise_std,
                               void* stream) {
  if (N == 0) return DFB_OK;
  DFB_REQUIRE(raw && z_vals, DFB_ERR_INVALID, "null argument");
  DFB_REQUIRE(C == 1 || C == 4 || C == 9, DFB_ERR_INVALID, "raw must have 1, 4 or 9 channels");
  DFB_REQUIRE((typ == 1) == (C == 9), DFB_ERR_INVALID, "fine compositing takes 9 channels, coarse 1 or 4");
  if (N == 0) return DFB_OK;
  CompositeArgs ca = {};
  ca.raw = raw, ca.z = z_vals, ca.N = N, ca.S = S, ca.C = C, ca.typ_fine = typ, ca.test_time = test_time;
  ca.beta_min = beta_min, ca.rgb = rgb, ca.disp = disp, ca.acc = acc, ca.weights = weights, ca.depth = depth;
  ca.tsig = transient_sigmas, ca.beta = beta;
  DFB_REQUIRE(raw_noise_std == 0.f || (noise && C != 9), DFB_ERR_INVALID, "raw_noise_std != 0 needs noise [N,S] (coarse pass only)");
  if (raw_noise_std != 0.f) ca.noise = noise, ca.noise_std = raw_noise_std;
  return launch_composite(ca, (cudaStream_t)stream);
}

extern "C" int dfb_get_rays(const float* c2w, int row_stride, int H, int W, float focal, float* rays_o, float* rays_d,
                            void* stream) {
  DFB_REQUIRE(c2w && rays_o && rays_d && H > 0 && W > 0, DFB_ERR_INVALID, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t N = (int64_t)H * W;
  float *rec = nullptr, *z = nullptr, *lin = nullptr;
  DFB_CHECK_CUDA(cudaMallocAsync(&rec, N * kRayRec * 4, st));
  DFB_CHECK_CUDA(cudaMallocAsync(&z, N * 4, st));
  DFB_CHECK_CUDA(cudaMallocAsync(&lin, 4, st));
  DFB_CHECK_CUDA(cudaMemsetAsync(lin, 0, 4, st));
  float hist0 = 0.f;
  float* d_hist = nullptr;
  DFB_CHECK_CUDA(cudaMallocAsync(&d_hist, 4, st));
  DFB_CHECK_CUDA(cudaMemcpyAsync(d_hist, &hist0, 4, cudaMemcpyHostToDevice, st));
  PrepArgs pa = {};
  pa.c2w = c2w, pa.c2w_ld = row_stride, pa.H = H, pa.W = W, pa.focal = focal, pa.near = 0.f, pa.far = 1.f;
  pa.hist = d_hist, pa.hb = 1, pa.n_vocab = 1, pa.N = N, pa.Nc = 1, pa.t_vals = lin, pa.rayrec = rec, pa.z = z;
  int rc = launch_prep(pa, st);
  if (rc == DFB_OK) {
    DFB_CHECK_CUDA(cudaMemcpy2DAsync(rays_o, 12, rec, kRayRec * 4, 12, N, cudaMemcpyDeviceToDevice, st));
    DFB_CHECK_CUDA(cudaMemcpy2DAsync(rays_d, 12, rec + 3, kRayRec * 4, 12, N, cudaMemcpyDeviceToDevice, st));
  }
  DFB_CHECK_CUDA(cudaStreamSynchronize(st));  // hist0 is a stack variable
  cudaFreeAsync(rec, st), cudaFreeAsync(z, st), cudaFreeAsync(lin, st), cudaFreeAsync(d_hist, st);
  return rc;
}

extern "C" int dfb_nerfw_forward(DfbNerf* n, int which, int mode, const float* x, int64_t P, float* out, void* stream) {
  DFB_REQUIRE(n && x && out, DFB_ERR_INVALID, "null argument");
  DFB_REQUIRE(which == 0 || which == 1, DFB_ERR_INVALID, "which must be 0 or 1");
  DFB_REQUIRE(mode >= 0 && mode <= 2, DFB_ERR_INVALID, "mode must be 0 (sigma), 1 (static) or 2 (full)");
  DFB_CHECK_CUDA(cudaSetDevice(n->device));
  return launch_mlp_simt_embedded(n, which, mode, x, P, out, (cudaStream_t)stream);
}

extern "C" int dfb_profile_enable(int on) {
  g_prof_on = on != 0;
  return DFB_OK;
}

extern "C" int dfb_profile_read(double* coarse_ms, double* fine_ms, int64_t* coarse_launches, int64_t* fine_launches) {
  double ms[2] = {0, 0};
  int64_t cnt[2] = {0, 0};
  for (auto& r : g_prof) {
    DFB_CHECK_CUDA(cudaEventSynchronize(r.b));
    float t = 0.f;
    DFB_CHECK_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
    ms[r.which] += t, cnt[r.which] += 1;
    cudaEventDestroy(r.a), cudaEventDestroy(r.b);
  }
  g_prof.clear();
  if (coarse_ms) *coarse_ms = ms[0];
  if (fine_ms) *fine_ms = ms[1];
  if (coarse_launches) *coarse_launches = cnt[0];
  if (fine_launches) *fine_launches = cnt[1];
  return DFB_OK;
}

// ------------------------------------------------------------------------------------------
// render backward w.r.t. the rays (test-time compositing, gradient of rgb only)
// ------------------------------------------------------------------------------------------
namespace dfb {
int launch_render_bwd(const DfbNerf* nerf, const float* rayrec, const float* z, const float* raybias, const float* raw,
                      const float* g_rgb, int64_t n_rays, int S, float* g_raw, float* g_samp, float* g_o, float* g_d,
                      float* g_vd, int kind, cudaStream_t st, const uint32_t* saved_masks);
}

namespace dfb {
extern uint32_t* g_dbg_simt_mask_dump;
extern const uint32_t* g_dbg_tc_mask_in;
extern uint32_t* g_dbg_tc_mask_out;
}
// Debug seam (tests only): ReLU masks of the render backward's forward recompute, per sample [P][12][8] words
// (fine 8x256 network, one workspace chunk: N <= 16384 rays).  simt_dump: written by the fp32 kernels (ballot
// layout: word j bit l = column 32j+l).  tc_out: written by the tcgen05 kernel (bit 16*(c&1) + (c>>1)%16 of word
// c/32 = column c).  tc_in: replaces the tcgen05 kernel's own masks, so that its gradient chain can be compared
// with the fp32 chain on IDENTICAL ReLU patterns.  Null pointers switch the seam off.
extern "C" int dfb_debug_bwd_masks(uint32_t* simt_dump, const uint32_t* tc_in, uint32_t* tc_out) {
  dfb::g_dbg_simt_mask_dump = simt_dump, dfb::g_dbg_tc_mask_in = tc_in, dfb::g_dbg_tc_mask_out = tc_out;
  return DFB_OK;
}

namespace {
constexpr int64_t kBwdChunk = 1 << 14;
struct BwdWs { size_t rayrec, extra, zdummy, rb, g_raw, g_samp, total; };
BwdWs bwd_layout(const DfbNerf* n, int64_t rays, int S) {
  const DfbNerfDesc& d = n->desc;
  BwdWs L;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return o; };
  L.rayrec = take((size_t)rays * kRayRec * 4);
  L.extra = take((size_t)rays * (27 + d.a_dim + d.t_dim) * 4);
  L.zdummy = take((size_t)rays * 4 + 256);
  L.rb = take((size_t)rays * std::max(d.W, 256) * 4);
  L.g_raw = take((size_t)rays * S * 9 * 4);
  L.g_samp = take((size_t)rays * S * 32 * 4);
  L.total = off;
  return L;
}
}  // namespace

extern "C" int dfb_render_bwd_workspace_bytes(const DfbNerf* n, int64_t n_rays, int S, size_t* out) {
  DFB_REQUIRE(n && out && n_rays >= 0 && S >= 1, DFB_ERR_INVALID, "bad arguments");
  *out = bwd_layout(n, std::min<int64_t>(std::max<int64_t>(n_rays, 1), kBwdChunk), S).total;
  return DFB_OK;
}

extern "C" int dfb_render_bwd(DfbNerf* n, const float* rays, int ray_stride, int64_t N, int S, const float* z_vals, const float* raw,
                              const float* g_rgb, float* g_rays_o, float* g_rays_d, float* g_viewdirs, void* ws,
                              size_t ws_bytes, void* stream) {
  return dfb_render_bwd_mma(n, DFB_MMA_FP32_SIMT, rays, ray_stride, N, S, z_vals, raw, g_rgb, g_rays_o, g_rays_d, g_viewdirs, ws, ws_bytes,
                            stream);
}

extern "C" int dfb_render_bwd_mma(DfbNerf* n, int mma_kind, const float* rays, int ray_stride, int64_t N, int S, const float* z_vals,
                                  const float* raw, const float* g_rgb, float* g_rays_o, float* g_rays_d, float* g_viewdirs,
                                  void* ws, size_t ws_bytes, void* stream) {
  return dfb_render_bwd_saved(n, mma_kind, rays, ray_stride, N, S, z_vals, raw, nullptr, g_rgb, g_rays_o, g_rays_d, g_viewdirs, ws, ws_bytes,
                              stream);
}

extern "C" int dfb_render_bwd_saved(DfbNerf* n, int mma_kind, const float* rays, int ray_stride, int64_t N, int S, const float* z_vals,
                                    const float* raw, const uint32_t* relu_masks, const float* g_rgb, float* g_rays_o,
                                    float* g_rays_d, float* g_viewdirs, void* ws, size_t ws_bytes, void* stream) {
  DFB_REQUIRE(!relu_masks || (mma_kind != DFB_MMA_FP32_SIMT && n && tc_bwd_supported(n)), DFB_ERR_UNSUPPORTED,
              "saved ReLU masks need the tcgen05 backward (8x256 fine network, mma f16 / bf16)");
  mma_kind = eff_kind(mma_kind);
  DFB_REQUIRE(mma_kind == DFB_MMA_FP32_SIMT || mma_kind == DFB_MMA_F16 || mma_kind == DFB_MMA_BF16, DFB_ERR_INVALID, "bad mma kind");
  DFB_REQUIRE(n && rays && z_vals && raw && g_rgb && g_rays_o && g_rays_d && g_viewdirs, DFB_ERR_INVALID, "null argument");
  DFB_REQUIRE(n->desc.has_fine && n->net[1].loaded && n->has_emb, DFB_ERR_INVALID, "fine network / embeddings not loaded");
  DFB_REQUIRE(ray_stride == 11 + n->desc.hist_bin, DFB_ERR_INVALID,
              "ray records are %d floats wide, expected 11 + hist_bin = %d ([o3,d3,near,far,viewdir3,hist])", ray_stride,
              11 + n->desc.hist_bin);
  if (N == 0) return DFB_OK;
  DFB_CHECK_CUDA(cudaSetDevice(n->device));
  cudaStream_t st = (cudaStream_t)stream;
  const DfbNerfDesc& d = n->desc;
  const int64_t chunk = std::min<int64_t>(N, kBwdChunk);
  const BwdWs L = bwd_layout(n, chunk, S);
  DFB_REQUIRE(ws && ws_bytes >= L.total, DFB_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", L.total, ws_bytes);
  char* base = (char*)ws;
  auto P = [&](size_t off) { return (float*)(base + off); };
  const int hb = d.hist_bin, n_extra = 27 + d.a_dim + d.t_dim;
  for (int64_t r0 = 0; r0 < N; r0 += chunk) {
    const int64_t nr = std::min<int64_t>(chunk, N - r0);
    PrepArgs pa = {};
    pa.rays = rays + r0 * (11 + hb), pa.hb = hb, pa.n_vocab = d.n_vocab, pa.emb_a = n->emb_a, pa.emb_t = n->emb_t;
    pa.N = nr, pa.Nc = 1, pa.t_vals = P(L.zdummy) + nr, pa.rayrec = P(L.rayrec), pa.z = P(L.zdummy);
    pa.extra = P(L.extra), pa.n_extra = n_extra, pa.a_dim = d.a_dim, pa.t_dim = d.t_dim;
    DFB_CHECK_CUDA(cudaMemsetAsync(P(L.zdummy) + nr, 0, 4, st));
    int rc = launch_prep(pa, st);
    if (rc) return rc;
    const bool tc_b = mma_kind != DFB_MMA_FP32_SIMT && tc_bwd_supported(n);  // 256-wide (zero-padded) rows for the tcgen05 kernel
    rc = launch_raybias(pa.extra, n_extra, nr, n->net[1], true, P(L.rb), tc_b ? 256 : n->net[1].n_dt, st, nullptr, 0,
                        tc_b ? 128 : 0);
    if (rc) return rc;
    rc = launch_render_bwd(n, P(L.rayrec), z_vals + r0 * S, P(L.rb), raw + r0 * S * 9, g_rgb + r0 * 3, nr, S, P(L.g_raw),
                           P(L.g_samp), g_rays_o + r0 * 3, g_rays_d + r0 * 3, g_viewdirs + r0 * 3, mma_kind, st,
                           relu_masks ? relu_masks + (size_t)(r0 * S / 128) * kReluMaskWordsPerTile : nullptr);
    if (rc) return rc;
  }
  return DFB_OK;
}
