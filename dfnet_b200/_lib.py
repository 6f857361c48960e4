"""ctypes binding of libdfnet_b200.so (the C ABI declared in include/dfnet_b200.h).

There is no fallback: if the shared library is missing the import fails loudly, and every
compute entry point refuses to run without a CUDA device.
"""
import ctypes as C
import os

import torch

_raw_stream = torch._C._cuda_getCurrentRawStream

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DFB_LIB_PATH", os.path.join(_HERE, "libdfnet_b200.so"))  # override: profiling builds only

# every symbol include/dfnet_b200.h declares (tests check that the library exports them all)
SYMBOLS = [
    "dfb_last_error", "dfb_version", "dfb_device_ok", "dfb_linspace_f32", "dfb_nerf_create", "dfb_nerf_destroy",
    "dfb_nerf_load", "dfb_nerf_set_embeddings", "dfb_nerfw_forward", "dfb_render_workspace_bytes",
    "dfb_render_fwd", "dfb_render_image_host", "dfb_render_bwd", "dfb_render_bwd_mma", "dfb_render_bwd_saved", "dfb_render_bwd_workspace_bytes", "dfb_sample_pdf", "dfb_raw2outputs", "dfb_get_rays",
    "dfb_launch_count", "dfb_profile_enable", "dfb_profile_read", "dfb_debug_umma_gemm", "dfb_debug_umma_gemm_mn", "dfb_debug_tc_prof", "dfb_debug_tcb_prof", "dfb_debug_bwd_masks", "dfb_debug_umma_rate", "dfb_debug_tmem_rate", "dfb_debug_tmem_rate_mma", "dfb_conv_create", "dfb_conv_destroy", "dfb_conv_fwd",
    "dfb_dfnet_create", "dfb_dfnet_destroy", "dfb_dfnet_load", "dfb_dfnet_workspace_bytes", "dfb_dfnet_fwd",
    "dfb_cosine_loss", "dfb_triplet_loss", "dfb_triplet_loss_bwd", "dfb_mse", "dfb_resize_bicubic", "dfb_resize_bilinear_ac",
    "dfb_conv_create_ex", "dfb_conv_fwd_ex", "dfb_conv_fwd_ex2", "dfb_conv_wgrad", "dfb_conv_wgrad_acc", "dfb_dfnet_load_ex", "dfb_dfnet_bn_batch_stats", "dfb_dfnet_tape_bytes", "dfb_debug_dfnet_tape_layout",
    "dfb_dfnet_bwd_workspace_bytes", "dfb_dfnet_bwd", "dfb_cosine_loss_bwd", "dfb_mse_bwd", "dfb_resize_bicubic_bwd",
    "dfb_resize_bilinear_ac_bwd", "dfb_dfnet_bwd_bucket_event", "dfb_luma_hist", "dfb_resize_area", "dfb_pose_error", "dfb_polar3x3_fwd", "dfb_polar3x3_bwd", "dfb_debug_conv_prof",
    "dfb_conv_update", "dfb_conv_update_many", "dfb_conv_pack_begin", "dfb_conv_pack_end", "dfb_embed_xyz16", "dfb_embed_xyz16_ex", "dfb_rows_expand16", "dfb_rows_reduce_bf16", "dfb_nerf_heads_fwd", "dfb_nerf_heads_bwd",
    "dfb_raw2outputs_bwd", "dfb_cast_f16_bf16", "dfb_render_poses_fwd", "dfb_copy2d_batch", "dfb_nerfw_loss_workspace_bytes", "dfb_nerfw_loss_fwd", "dfb_nerfw_loss_bwd",
    "dfb_pose_rays_fwd", "dfb_pose_rays_workspace_bytes", "dfb_pose_rays_bwd",
]

MMA_FP32_SIMT, MMA_F16, MMA_BF16, MMA_F16_SPLIT_COARSE = 0, 1, 2, 3
# "f16s": fp16 tensor-core path with the split-precision (hi + lo operands) coarse pass, see include/dfnet_b200.h
MMA_KINDS = {"fp32": MMA_FP32_SIMT, "f16": MMA_F16, "bf16": MMA_BF16, "f16s": MMA_F16_SPLIT_COARSE}


class Copy2d(C.Structure):
    """DfbCopy2d (include/dfnet_b200.h): one strided fp32 copy of dfb_copy2d_batch."""
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("rows", C.c_int), ("cols", C.c_int), ("src_ld", C.c_int), ("dst_ld", C.c_int)]


class NerfDesc(C.Structure):
    _fields_ = [("D", C.c_int32), ("W", C.c_int32), ("skip", C.c_int32), ("L_xyz", C.c_int32), ("L_dir", C.c_int32),
                ("a_dim", C.c_int32), ("t_dim", C.c_int32), ("hist_bin", C.c_int32), ("n_vocab", C.c_int32),
                ("beta_min", C.c_float), ("has_fine", C.c_int32)]


class RenderCfg(C.Structure):
    _fields_ = [("N_samples", C.c_int32), ("N_importance", C.c_int32), ("test_time", C.c_int32),
                ("perturb", C.c_int32), ("mma_kind", C.c_int32), ("lindisp", C.c_int32),
                ("raw_noise_std", C.c_float), ("ray_stride", C.c_int32), ("hist_len", C.c_int32), ("ert_eps", C.c_float)]


class RenderExtras(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in
                ("rgb0", "disp0", "acc0", "z_std", "beta", "transient_sigmas", "raw", "weights_coarse", "z_vals",
                 "z_samples", "inds", "depth", "relu_masks", "n_live")]


class DfbError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C dfnet_b200/csrc` (there is no non-CUDA fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float
    lib.dfb_last_error.restype = C.c_char_p
    lib.dfb_last_error.argtypes = []
    lib.dfb_version.argtypes = []
    lib.dfb_device_ok.argtypes = []
    lib.dfb_launch_count.restype = i64
    lib.dfb_launch_count.argtypes = []
    lib.dfb_linspace_f32.argtypes = [f32, f32, i32, C.POINTER(C.c_float)]
    lib.dfb_nerf_create.argtypes = [C.POINTER(NerfDesc), C.POINTER(vp)]
    lib.dfb_nerf_destroy.argtypes = [vp]
    lib.dfb_nerf_destroy.restype = None
    lib.dfb_nerf_load.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(i64), i32]
    lib.dfb_nerf_set_embeddings.argtypes = [vp, vp, vp]
    lib.dfb_nerfw_forward.argtypes = [vp, i32, i32, vp, i64, vp, vp]
    lib.dfb_render_workspace_bytes.argtypes = [vp, C.POINTER(RenderCfg), i64, C.POINTER(C.c_size_t)]
    lib.dfb_render_fwd.argtypes = [vp, C.POINTER(RenderCfg), vp, vp, i32, i32, f32, f32, f32, vp, i64, vp, vp, vp, vp, vp,
                                   vp, C.POINTER(RenderExtras), vp, C.c_size_t, vp]
    lib.dfb_render_poses_fwd.argtypes = [vp, C.POINTER(RenderCfg), vp, i32, i32, i32, f32, f32, f32, vp, vp, vp, vp, vp, C.c_size_t, vp]
    lib.dfb_render_image_host.argtypes = [vp, C.POINTER(RenderCfg), vp, i32, i32, f32, f32, f32, vp, vp, vp, vp, vp,
                                          C.c_size_t, vp]
    lib.dfb_render_bwd_workspace_bytes.argtypes = [vp, i64, i32, C.POINTER(C.c_size_t)]
    lib.dfb_render_bwd.argtypes = [vp, vp, i32, i64, i32, vp, vp, vp, vp, vp, vp, vp, C.c_size_t, vp]
    lib.dfb_debug_bwd_masks.argtypes = [vp, vp, vp]
    lib.dfb_render_bwd_saved.argtypes = [vp, i32, vp, i32, i64, i32, vp, vp, vp, vp, vp, vp, vp, vp, C.c_size_t, vp]
    lib.dfb_render_bwd_mma.argtypes = [vp, i32, vp, i32, i64, i32, vp, vp, vp, vp, vp, vp, vp, C.c_size_t, vp]
    lib.dfb_sample_pdf.argtypes = [vp, vp, vp, i64, i32, i32, vp, vp, vp]
    lib.dfb_raw2outputs.argtypes = [vp, vp, i64, i32, i32, i32, i32, f32, vp, vp, vp, vp, vp, vp, vp, vp, f32, vp]
    lib.dfb_get_rays.argtypes = [vp, i32, i32, i32, f32, vp, vp, vp]
    lib.dfb_debug_umma_gemm.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp]
    lib.dfb_debug_umma_gemm_mn.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp, vp]
    lib.dfb_debug_tc_prof.argtypes = [vp, i32]
    lib.dfb_debug_tcb_prof.argtypes = [vp, i32]
    lib.dfb_debug_umma_rate.argtypes = [i32, i32, i32, C.POINTER(C.c_double)]
    lib.dfb_debug_tmem_rate.argtypes = [i32, i32, i32, C.POINTER(C.c_double)]
    lib.dfb_debug_tmem_rate_mma.argtypes = [i32, i32, i32, i32, C.POINTER(C.c_double)]
    lib.dfb_conv_create.argtypes = [i32, i32, i32, i32, vp, vp, vp, vp, C.POINTER(vp)]
    lib.dfb_conv_destroy.argtypes = [vp]
    lib.dfb_conv_destroy.restype = None
    lib.dfb_conv_fwd.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp, vp, vp]
    lib.dfb_conv_create_ex.argtypes = [i32, i32, i32, i32, vp, vp, vp, vp, i32, i32, C.POINTER(vp)]
    lib.dfb_conv_fwd_ex.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp]
    lib.dfb_conv_wgrad.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp]
    lib.dfb_conv_wgrad_acc.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp]
    lib.dfb_conv_fwd_ex2.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp]
    lib.dfb_dfnet_load_ex.argtypes = [vp, C.POINTER(vp), C.POINTER(i64), i32, f32, C.c_uint32]
    lib.dfb_dfnet_bn_batch_stats.argtypes = [vp, vp, vp]
    lib.dfb_dfnet_tape_bytes.argtypes = [vp, i32, i32, i32, i32, i32, C.POINTER(C.c_size_t)]
    lib.dfb_debug_dfnet_tape_layout.argtypes = [vp, i32, i32, i32, i32, i32, C.POINTER(i64)]
    lib.dfb_dfnet_bwd_workspace_bytes.argtypes = [vp, i32, i32, i32, C.POINTER(C.c_size_t)]
    lib.dfb_dfnet_bwd.argtypes = [vp, i32, i32, i32, C.c_uint32, i32, i32, vp, vp, C.c_uint32, vp, vp, vp, C.POINTER(vp), i32,
                                  vp, C.c_size_t, vp]
    lib.dfb_dfnet_bwd_bucket_event.argtypes = [vp, i32, vp]
    lib.dfb_conv_update.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.dfb_conv_update_many.argtypes = [vp, vp, vp, i32, vp]
    lib.dfb_embed_xyz16.argtypes = [vp, i32, vp, i64, i32, i32, i32, vp, vp]
    lib.dfb_embed_xyz16_ex.argtypes = [vp, i32, vp, i64, i32, i32, i32, vp, vp, vp]
    lib.dfb_rows_expand16.argtypes = [vp, i64, i32, i32, vp, vp]
    lib.dfb_rows_reduce_bf16.argtypes = [vp, i64, i32, i32, vp, vp]
    lib.dfb_nerf_heads_fwd.argtypes = [vp, vp, vp, i64, i64, i32, vp, vp]
    lib.dfb_nerf_heads_bwd.argtypes = [vp, vp, i64, i32, vp, vp, vp, vp]
    lib.dfb_raw2outputs_bwd.argtypes = [vp, vp, i64, i32, i32, vp, f32, vp, vp, vp, vp, vp]
    lib.dfb_cast_f16_bf16.argtypes = [vp, vp, i64, vp]
    lib.dfb_debug_conv_prof.argtypes = [i32, vp, i32, C.POINTER(i32)]
    lib.dfb_luma_hist.argtypes = [vp, i32, i32, i32, i32, vp, vp, C.c_size_t, vp]
    lib.dfb_resize_area.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp]
    lib.dfb_pose_error.argtypes = [vp, vp, i32, i32, vp, vp, vp]
    lib.dfb_conv_pack_begin.argtypes = []
    lib.dfb_conv_pack_end.argtypes = [vp]
    lib.dfb_polar3x3_fwd.argtypes = [vp, i32, vp, vp, vp]
    lib.dfb_polar3x3_bwd.argtypes = [vp, vp, i32, vp, vp]
    lib.dfb_copy2d_batch.argtypes = [vp, i32, vp]
    lib.dfb_pose_rays_fwd.argtypes = [vp, i32, i32, i32, f32, vp, vp, vp, vp]
    lib.dfb_pose_rays_workspace_bytes.restype = C.c_size_t
    lib.dfb_pose_rays_workspace_bytes.argtypes = []
    lib.dfb_pose_rays_bwd.argtypes = [vp, i32, i32, i32, f32, vp, vp, vp, vp, vp, vp]
    lib.dfb_nerfw_loss_workspace_bytes.restype = C.c_size_t
    lib.dfb_nerfw_loss_workspace_bytes.argtypes = []
    lib.dfb_nerfw_loss_fwd.argtypes = [vp, vp, vp, vp, vp, i64, i32, f32, f32, vp, vp, vp]
    lib.dfb_nerfw_loss_bwd.argtypes = [vp, vp, vp, vp, i64, i32, f32, f32, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.dfb_cosine_loss_bwd.argtypes = [vp, vp, i32, i64, i32, f32, vp, vp, vp, C.c_size_t, vp]
    lib.dfb_mse_bwd.argtypes = [vp, vp, i64, vp, vp, vp]
    lib.dfb_resize_bicubic_bwd.argtypes = [vp, i64, i32, i32, i32, i32, vp, vp]
    lib.dfb_resize_bilinear_ac_bwd.argtypes = [vp, i64, i32, i32, i32, i32, vp, vp]
    lib.dfb_dfnet_create.argtypes = [i32, C.POINTER(vp)]
    lib.dfb_dfnet_destroy.argtypes = [vp]
    lib.dfb_dfnet_destroy.restype = None
    lib.dfb_dfnet_load.argtypes = [vp, C.POINTER(vp), C.POINTER(i64), i32, f32]
    lib.dfb_dfnet_workspace_bytes.argtypes = [vp, i32, i32, i32, i32, i32, C.POINTER(C.c_size_t)]
    lib.dfb_dfnet_fwd.argtypes = [vp, vp, i32, i32, i32, C.c_uint32, i32, i32, vp, vp, vp, vp, C.c_size_t, vp]
    lib.dfb_cosine_loss.argtypes = [vp, vp, i32, i64, i32, f32, vp, vp, C.c_size_t, vp]
    lib.dfb_triplet_loss.argtypes = [vp, vp, i32, i32, i32, i32, i32, f32, vp, vp, vp, C.c_size_t, vp]
    lib.dfb_triplet_loss_bwd.argtypes = [vp, vp, i32, i32, i32, i32, i32, f32, vp, vp, vp, vp, vp]
    lib.dfb_mse.argtypes = [vp, vp, i64, vp, vp, C.c_size_t, vp]
    lib.dfb_resize_bicubic.argtypes = [vp, i64, i32, i32, i32, i32, vp, vp]
    lib.dfb_resize_bilinear_ac.argtypes = [vp, i64, i32, i32, i32, i32, vp, vp]
    lib.dfb_profile_enable.argtypes = [i32]
    lib.dfb_profile_read.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(i64), C.POINTER(i64)]
    return lib


lib = _load()


def raw_stream():
    """The current torch CUDA stream of the current device as a stream argument of the C ABI (a `void*`). The raw accessor
    costs ~0.2 us against ~1.5 us for `torch.cuda.current_stream().cuda_stream`, which adds up over the ~500 launches of a
    training step."""
    return C.c_void_p(_raw_stream(-1))


def check(rc):
    if rc != 0:
        raise DfbError(f"libdfnet_b200 error {rc}: {lib.dfb_last_error().decode()}")
