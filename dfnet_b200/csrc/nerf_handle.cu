// DfbNerf handle: creation, parameter loading / repacking, destruction.
// Replaces the parameter ownership of models/nerfw.py:220-295 (NeRFW) and the embedding
// tables of create_nerf (nerfw.py:386-394) on the kernel side.
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace dfb {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches += n; }

int device_error_flag(int** out) {
  static int* flags[64] = {nullptr};
  int dev = 0;
  DFB_CHECK_CUDA(cudaGetDevice(&dev));
  DFB_REQUIRE(dev >= 0 && dev < 64, DFB_ERR_UNSUPPORTED, "device index %d out of range", dev);
  if (!flags[dev]) {
    // first use of this device by the library: keep the stream-ordered allocator's pool resident (the default release
    // threshold of 0 returns the memory to the driver at every synchronisation point, and re-acquiring it costs ms)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t keep = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
    DFB_CHECK_CUDA(cudaMalloc(&flags[dev], sizeof(int)));
    DFB_CHECK_CUDA(cudaMemset(flags[dev], 0, sizeof(int)));
  }
  *out = flags[dev];
  return DFB_OK;
}

}  // namespace dfb

using namespace dfb;

extern "C" const char* dfb_last_error(void) { return dfb::g_err; }
extern "C" int dfb_version(void) { return 100; }
extern "C" int64_t dfb_launch_count(void) { return dfb::g_launches.load(); }

extern "C" int dfb_device_ok(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  return p.major == 10 ? 1 : 0;
}

extern "C" int dfb_linspace_f32(float start, float end, int steps, float* out) {
  DFB_REQUIRE(steps >= 1 && out, DFB_ERR_INVALID, "dfb_linspace_f32: bad arguments");
  if (steps == 1) { out[0] = start; return DFB_OK; }
  // ATen-CPU: step in float32; lower half fma(step,i,start), upper half fma(-step,steps-1-i,end).
  volatile float stepv = (end - start) / (float)(steps - 1);
  float step = stepv;
  int half = steps / 2;
  for (int i = 0; i < steps; ++i)
    out[i] = i < half ? fmaf(step, (float)i, start) : fmaf(-step, (float)(steps - 1 - i), end);
  return DFB_OK;
}

extern "C" int dfb_nerf_create(const DfbNerfDesc* d, DfbNerf** out) {
  DFB_REQUIRE(d && out, DFB_ERR_INVALID, "dfb_nerf_create: null argument");
  DFB_REQUIRE(d->D >= 1 && d->D <= 16, DFB_ERR_UNSUPPORTED, "netdepth %d outside [1,16]", d->D);
  DFB_REQUIRE(d->W >= 64 && d->W <= 256 && d->W % 64 == 0, DFB_ERR_UNSUPPORTED,
              "netwidth %d unsupported (64, 128, 192 or 256)", d->W);
  DFB_REQUIRE(d->L_xyz == 10 && d->L_dir == 4, DFB_ERR_UNSUPPORTED,
              "only the paper-default embedding (multires=10, multires_views=4) is on the hot path");
  DFB_REQUIRE(d->a_dim >= 0 && d->a_dim <= 64 && d->t_dim >= 0 && d->t_dim <= 32, DFB_ERR_UNSUPPORTED,
              "in_channels_a/in_channels_t (%d,%d) too large", d->a_dim, d->t_dim);
  DFB_REQUIRE(d->hist_bin >= 1 && d->hist_bin <= 16 && d->a_dim == d->hist_bin * 5 && d->t_dim == d->hist_bin * 2,
              DFB_ERR_UNSUPPORTED, "encode_hist layout requires in_channels_a = 5*hist_bin and in_channels_t = 2*hist_bin");
  DFB_REQUIRE(d->skip < 0 || (d->skip >= 1), DFB_ERR_INVALID, "skip layer index must be >= 1");
  DfbNerf* n = new DfbNerf();
  n->desc = *d;
  DFB_CHECK_CUDA(cudaGetDevice(&n->device));
  cudaDeviceProp p;
  DFB_CHECK_CUDA(cudaGetDeviceProperties(&p, n->device));
  n->num_sms = p.multiProcessorCount;
  DFB_REQUIRE(p.major == 10, DFB_ERR_UNSUPPORTED,
              "device is sm_%d%d; libdfnet_b200 contains sm_100a code only", p.major, p.minor);
  *out = n;
  return DFB_OK;
}

extern "C" void dfb_nerf_destroy(DfbNerf* n) {
  if (!n) return;
  for (int i = 0; i < 2; ++i) {
    if (n->net[i].blob32) cudaFree(n->net[i].blob32);
    if (n->net[i].blob32b) cudaFree(n->net[i].blob32b);
    for (int k = 0; k < 2; ++k)
      for (int g = 0; g < 2; ++g)
        if (n->net[i].blob16[k][g]) cudaFree(n->net[i].blob16[k][g]);
    for (int k = 0; k < 2; ++k) {
      if (n->net[i].blob16b[k]) cudaFree(n->net[i].blob16b[k]);
      if (n->net[i].blob16b2[k]) cudaFree(n->net[i].blob16b2[k]);
    }
    if (n->net[i].tc_dtbias_dev) cudaFree(n->net[i].tc_dtbias_dev);
    if (n->net[i].tc_dtbias_n_dev) cudaFree(n->net[i].tc_dtbias_n_dev);
    for (int k = 0; k < 2; ++k)
      if (n->net[i].blob16n[k]) cudaFree(n->net[i].blob16n[k]);
  }
  if (n->emb_a) cudaFree(n->emb_a);
  if (n->emb_t) cudaFree(n->emb_t);
  if (n->lin_dev) cudaFree(n->lin_dev);
  if (n->bwd_scratch) cudaFree(n->bwd_scratch);
  delete n;
}

extern "C" int dfb_nerf_set_embeddings(DfbNerf* n, const float* emb_a, const float* emb_t) {
  DFB_REQUIRE(n && emb_a && emb_t, DFB_ERR_INVALID, "dfb_nerf_set_embeddings: null argument");
  DFB_CHECK_CUDA(cudaSetDevice(n->device));
  size_t na = (size_t)n->desc.n_vocab * 5, nt = (size_t)n->desc.n_vocab * 2;
  if (!n->emb_a) DFB_CHECK_CUDA(cudaMalloc(&n->emb_a, na * sizeof(float)));
  if (!n->emb_t) DFB_CHECK_CUDA(cudaMalloc(&n->emb_t, nt * sizeof(float)));
  DFB_CHECK_CUDA(cudaMemcpy(n->emb_a, emb_a, na * sizeof(float), cudaMemcpyDefault));
  DFB_CHECK_CUDA(cudaMemcpy(n->emb_t, emb_t, nt * sizeof(float), cudaMemcpyDefault));
  n->has_emb = true;
  return DFB_OK;
}

extern "C" int dfb_nerf_load(DfbNerf* n, int which, const float* const* params, const int64_t* numel, int n_params) {
  DFB_REQUIRE(n && params && numel, DFB_ERR_INVALID, "dfb_nerf_load: null argument");
  DFB_REQUIRE(which == 0 || which == 1, DFB_ERR_INVALID, "which must be 0 (coarse) or 1 (fine)");
  DFB_REQUIRE(which == 0 || n->desc.has_fine, DFB_ERR_INVALID, "handle was created without a fine network");
  DFB_CHECK_CUDA(cudaSetDevice(n->device));
  const DfbNerfDesc& d = n->desc;
  const int D = d.D, W = d.W, H = W / 2;
  const int in_xyz = 3 + 6 * d.L_xyz, in_dir = 3 + 6 * d.L_dir;
  const bool fine = which == 1;
  const int a_dim = fine ? d.a_dim : 0, t_dim = d.t_dim;

  // Expected state_dict order and sizes (models/nerfw.py:258-295).
  std::vector<int64_t> expect;
  for (int i = 0; i < D; ++i) {
    int kin = i == 0 ? in_xyz : (i == d.skip ? W + in_xyz : W);
    expect.push_back((int64_t)W * kin);
    expect.push_back(W);
  }
  expect.push_back((int64_t)W * W), expect.push_back(W);                       // xyz_encoding_final
  expect.push_back((int64_t)H * (W + in_dir + a_dim)), expect.push_back(H);    // dir_encoding.0
  expect.push_back(W), expect.push_back(1);                                    // static_sigma.0
  expect.push_back(3 * H), expect.push_back(3);                                // static_rgb.0
  if (fine) {
    expect.push_back((int64_t)H * (W + t_dim)), expect.push_back(H);           // transient_encoding.0
    for (int i = 0; i < 3; ++i) expect.push_back((int64_t)H * H), expect.push_back(H);
    expect.push_back(H), expect.push_back(1);                                  // transient_sigma.0
    expect.push_back(3 * H), expect.push_back(3);                              // transient_rgb.0
    expect.push_back(H), expect.push_back(1);                                  // transient_beta.0
  }
  DFB_REQUIRE((int)expect.size() == n_params, DFB_ERR_INVALID,
              "network %d: expected %d state_dict tensors, got %d", which, (int)expect.size(), n_params);
  std::vector<std::vector<float>> P(n_params);
  for (int i = 0; i < n_params; ++i) {
    DFB_REQUIRE(numel[i] == expect[i], DFB_ERR_INVALID, "network %d: tensor %d has %lld elements, expected %lld",
                which, i, (long long)numel[i], (long long)expect[i]);
    P[i].resize(numel[i]);
    DFB_CHECK_CUDA(cudaMemcpy(P[i].data(), params[i], numel[i] * sizeof(float), cudaMemcpyDefault));
  }

  NetPack& np = n->net[which];
  np.fine = fine;
  np.D = D, np.W = W, np.skip = d.skip, np.in_xyz = in_xyz, np.in_dir = in_dir, np.a_dim = a_dim, np.t_dim = t_dim;
  np.pek = round_up(in_xyz, 16);
  const int pek = np.pek;

  std::vector<float> blob;
  auto alloc = [&](size_t cnt) {
    size_t off = (blob.size() + 3) / 4 * 4;  // 16-byte aligned rows for float4 loads
    blob.resize(off + cnt, 0.f);
    return off;
  };
  np.trunk_w.assign(D, 0), np.trunk_b.assign(D, 0);
  for (int i = 0; i < D; ++i) {
    const std::vector<float>&w = P[2 * i], &b = P[2 * i + 1];
    bool is_skip = (i == d.skip);
    int kin = i == 0 ? in_xyz : (is_skip ? W + in_xyz : W);
    int kp = i == 0 ? pek : (is_skip ? pek + W : W);
    size_t o = alloc((size_t)kp * W);
    for (int nn = 0; nn < W; ++nn)
      for (int k = 0; k < kin; ++k) {
        int kk = k;                                   // layer 0 / plain layers
        if (is_skip) kk = k < in_xyz ? k : pek + (k - in_xyz);  // cat([input_xyz, h]) -> [pe(padded) | h]
        blob[o + (size_t)kk * W + nn] = w[(size_t)nn * kin + k];
      }
    np.trunk_w[i] = o;
    size_t ob = alloc(W);
    memcpy(&blob[ob], b.data(), W * sizeof(float));
    np.trunk_b[i] = ob;
  }
  int pi = 2 * D;
  const std::vector<float>&wf = P[pi], &bf = P[pi + 1], &wd = P[pi + 2], &bd = P[pi + 3], &ws = P[pi + 4],
                          &bs = P[pi + 5], &wr = P[pi + 6], &br = P[pi + 7];
  np.final_w = alloc((size_t)W * W);
  for (int nn = 0; nn < W; ++nn)
    for (int k = 0; k < W; ++k) blob[np.final_w + (size_t)k * W + nn] = wf[(size_t)nn * W + k];
  np.final_b = alloc(W);
  memcpy(&blob[np.final_b], bf.data(), W * sizeof(float));
  np.sigma_w = alloc(W), np.sigma_b = alloc(1);
  memcpy(&blob[np.sigma_w], ws.data(), W * sizeof(float));
  blob[np.sigma_b] = bs[0];
  np.rgb_w = alloc(3 * H), np.rgb_b = alloc(3);
  memcpy(&blob[np.rgb_w], wr.data(), 3 * H * sizeof(float));
  memcpy(&blob[np.rgb_b], br.data(), 3 * sizeof(float));

  np.n_dt = fine ? W : H;
  const int ndt = np.n_dt, kd = W + in_dir + a_dim;
  np.dt_w = alloc((size_t)W * ndt);
  for (int nn = 0; nn < H; ++nn)
    for (int k = 0; k < W; ++k) blob[np.dt_w + (size_t)k * ndt + nn] = wd[(size_t)nn * kd + k];
  np.dirx_w = alloc((size_t)(in_dir + a_dim) * H);
  for (int nn = 0; nn < H; ++nn)
    for (int j = 0; j < in_dir + a_dim; ++j) blob[np.dirx_w + (size_t)j * H + nn] = wd[(size_t)nn * kd + W + j];
  np.dirx_b = alloc(H);
  memcpy(&blob[np.dirx_b], bd.data(), H * sizeof(float));
  if (fine) {
    const std::vector<float>&wt0 = P[pi + 8], &bt0 = P[pi + 9];
    const int kt = W + t_dim;
    for (int nn = 0; nn < H; ++nn)
      for (int k = 0; k < W; ++k) blob[np.dt_w + (size_t)k * ndt + H + nn] = wt0[(size_t)nn * kt + k];
    np.tx_w = alloc((size_t)t_dim * H);
    for (int nn = 0; nn < H; ++nn)
      for (int j = 0; j < t_dim; ++j) blob[np.tx_w + (size_t)j * H + nn] = wt0[(size_t)nn * kt + W + j];
    np.tx_b = alloc(H);
    memcpy(&blob[np.tx_b], bt0.data(), H * sizeof(float));
    for (int i = 0; i < 3; ++i) {
      const std::vector<float>&w = P[pi + 10 + 2 * i], &b = P[pi + 11 + 2 * i];
      np.t_w[i] = alloc((size_t)H * H);
      for (int nn = 0; nn < H; ++nn)
        for (int k = 0; k < H; ++k) blob[np.t_w[i] + (size_t)k * H + nn] = w[(size_t)nn * H + k];
      np.t_b[i] = alloc(H);
      memcpy(&blob[np.t_b[i]], b.data(), H * sizeof(float));
    }
    np.tsig_w = alloc(H), np.tsig_b = alloc(1);
    memcpy(&blob[np.tsig_w], P[pi + 16].data(), H * sizeof(float));
    blob[np.tsig_b] = P[pi + 17][0];
    np.trgb_w = alloc(3 * H), np.trgb_b = alloc(3);
    memcpy(&blob[np.trgb_w], P[pi + 18].data(), 3 * H * sizeof(float));
    memcpy(&blob[np.trgb_b], P[pi + 19].data(), 3 * sizeof(float));
    np.tbeta_w = alloc(H), np.tbeta_b = alloc(1);
    memcpy(&blob[np.tbeta_w], P[pi + 20].data(), H * sizeof(float));
    blob[np.tbeta_b] = P[pi + 21][0];
  }
  if (np.blob32) cudaFree(np.blob32);
  np.blob32 = nullptr;
  np.blob32_floats = blob.size();
  DFB_CHECK_CUDA(cudaMalloc(&np.blob32, blob.size() * sizeof(float)));
  DFB_CHECK_CUDA(cudaMemcpy(np.blob32, blob.data(), blob.size() * sizeof(float), cudaMemcpyHostToDevice));

  if (fine) {
    // backward layout: rows = outputs (torch layout), input dimension padded / reordered like the forward K order
    std::vector<float> bb;
    auto balloc = [&](size_t cnt) { size_t off = (bb.size() + 3) / 4 * 4; bb.resize(off + cnt, 0.f); return off; };
    np.bw_trunk.assign(D, 0);
    for (int i = 0; i < D; ++i) {
      const std::vector<float>& w = P[2 * i];
      const bool is_skip = (i == d.skip);
      const int kin = i == 0 ? in_xyz : (is_skip ? W + in_xyz : W);
      const int kp = i == 0 ? pek : (is_skip ? pek + W : W);
      const size_t o = balloc((size_t)W * kp);
      for (int nn = 0; nn < W; ++nn)
        for (int k = 0; k < kin; ++k) {
          int kk = k;
          if (is_skip) kk = k < in_xyz ? k : pek + (k - in_xyz);
          bb[o + (size_t)nn * kp + kk] = w[(size_t)nn * kin + k];
        }
      np.bw_trunk[i] = o;
    }
    np.bw_final = balloc((size_t)W * W);
    memcpy(&bb[np.bw_final], wf.data(), (size_t)W * W * sizeof(float));
    const std::vector<float>& wt0 = P[pi + 8];
    const int kdd = W + in_dir + a_dim, ktt = W + t_dim;
    np.bw_dt = balloc((size_t)W * W);
    for (int nn = 0; nn < H; ++nn)
      for (int k = 0; k < W; ++k) {
        bb[np.bw_dt + (size_t)nn * W + k] = wd[(size_t)nn * kdd + k];
        bb[np.bw_dt + (size_t)(H + nn) * W + k] = wt0[(size_t)nn * ktt + k];
      }
    np.bw_dtx = balloc((size_t)H * 32);
    for (int nn = 0; nn < H; ++nn)
      for (int j = 0; j < in_dir; ++j) bb[np.bw_dtx + (size_t)nn * 32 + j] = wd[(size_t)nn * kdd + W + j];
    for (int i = 0; i < 3; ++i) {
      np.bw_t[i] = balloc((size_t)H * H);
      memcpy(&bb[np.bw_t[i]], P[pi + 10 + 2 * i].data(), (size_t)H * H * sizeof(float));
    }
    if (np.blob32b) cudaFree(np.blob32b);
    np.blob32b = nullptr;
    DFB_CHECK_CUDA(cudaMalloc(&np.blob32b, bb.size() * sizeof(float)));
    DFB_CHECK_CUDA(cudaMemcpy(np.blob32b, bb.data(), bb.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  int rc = pack_tc_weights(n, which, P);
  if (rc != DFB_OK) return rc;
  rc = pack_tc_bwd_weights(n, which, P);
  if (rc != DFB_OK) return rc;
  np.loaded = true;
  return DFB_OK;
}
