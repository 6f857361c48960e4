// DFNet handle shared by the forward (dfnet_kernels.cu) and training (dfnet_train.cu) paths.
#pragma once
#include "common.cuh"

struct DfbConv;
extern "C" int dfb_conv_create(int Cin, int Cout, int KH, int KW, const float* weight, const float* bias,
                               const float* bn_scale, const float* bn_shift, DfbConv** out);
extern "C" void dfb_conv_destroy(DfbConv* c);
extern "C" int dfb_conv_fwd(DfbConv* c, const void* in_nhwc16, int B, int H, int W, int relu, void* out_nhwc16,
                            void* tap_nhwc16, float* out_nchw32, void* stream);
extern "C" int dfb_conv_wgrad(const void* gO, const void* X, int B, int H, int W, int Cin, int Cin_pad, int Cout, int KH, int fmt,
                              float* dW, float* dB, void* stream);
int dfb_conv_create_impl(int Cin0, int Cout0, int KH, int KW, const float* weight, const float* bias, const float* bn_scale,
                         const float* bn_shift, int fmt, int dgrad, DfbConv** out, int nt_force = 0);
int64_t dfb_conv_tiles(const DfbConv* c, int B, int H, int W);
int dfb_conv_num_sms(const DfbConv* c);
int64_t dfb_conv_rounds(const DfbConv* c, int B, int H, int W);
int dfb_conv_update_impl(DfbConv* c, const float* weight, const float* bias, const float* bn_scale, const float* bn_shift,
                         void* stream);
void dfb_conv_pack_batch_begin();
int dfb_conv_pack_batch_flush(bool discard, void* stream = nullptr);
int dfb_conv_run(DfbConv* c, const void* in_nhwc16, int B, int H, int W, int relu, void* out_nhwc16, void* tap_nhwc16,
                 float* out_nchw32, const void* mask_nhwc16, const void* addend_nhwc16, void* stream, void* out_bf16 = nullptr);

struct DfbDfnet {
  int n_levels = 3;
  // inference: fp16 operands
  DfbConv* enc[13] = {};
  DfbConv* enc_n128[13] = {};   // 128-wide output-channel tiles of the wide layers, used when 256-wide tiles under-fill the GPU
  DfbConv* head1[3] = {};
  DfbConv* head5[3] = {};
  // training: bf16 forward of the encoder (activations double as wgrad operands) and bf16 data-gradient convs
  DfbConv* enc_bf[13] = {};
  DfbConv* enc_dg[13] = {};
  DfbConv* head1_dg[3] = {};
  DfbConv* head5_dg[3] = {};
  float* bn_stage = nullptr;  // [3][4][128] staging of the BatchNorm vectors handed to dfb_dfnet_load
  float* bn_sc = nullptr;  // [3][128] eval-mode BatchNorm scale / shift of the heads (device)
  float* bn_sh = nullptr;
  // train-mode BatchNorm of the heads (run_feature.py without freezeBN: batch statistics over the whole call's batch)
  DfbConv* head5_raw[3] = {};   // the 5x5 convs without the BatchNorm folded in
  DfbConv* head5_raw_dg[3] = {};  // their data-gradient convolutions (training of the heads: BatchNorm differentiated separately)
  float* bn_gb = nullptr;       // [3][2][128] gamma, beta (device)
  float* bn_stat = nullptr;     // [3][4][128] batch mean, biased variance, scale, shift of the last train-mode forward
  double* bn_part = nullptr;    // [128][64][2] partial sums
  float bn_eps = 1e-5f;
  float* fc_w = nullptr;
  float* fc_b = nullptr;
  bool loaded = false;
  // data-parallel training: cudaEvent_t recorded by dfb_dfnet_bwd on its stream as soon as the parameter
  // gradients of fc_pose and of the encoder layers >= bucket_first_layer are complete (the deep, parameter-heavy
  // layers are differentiated FIRST), so that the caller can all-reduce that bucket while the backward continues
  void* bucket_event = nullptr;
  int bucket_first_layer = 0;
  // forward: side streams for the adaptation heads (fork / join around the caller's stream, see dfnet_fwd_impl)
  cudaStream_t side[2] = {nullptr, nullptr};
  cudaEvent_t ev[5] = {};
};

static const int kEncCin[13] = {3, 64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512};
static const int kEncCout[13] = {64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512, 512};
static const bool kPoolAfter[13] = {false, true, false, true, false, false, true, false, false, true, false, false, true};
static const int kTapConv[3] = {1, 6, 12};   // conv1_2, conv3_3, conv5_3
static const int kTapCh[3] = {64, 256, 512};

struct DfWs {
  size_t in8, act[13], pool[13], tap[3], mid[3], zbn[3], feat, pooled, total;   // zbn: pre-BatchNorm head outputs (tape only)
  int h[13], w[13];  // input resolution of every encoder conv
};
DfWs dfnet_ws(int B, int H, int W, int n_levels, int upH, int upW, bool tape);
