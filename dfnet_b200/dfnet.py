"""Host-side mirror of the reference's `feature/dfnet.py` (DFNet / DFNet_s).

Same module tree and state_dict keys as the reference (`encoder.{0..28}.*`,
`adaptation_layers.adapt_layer_{i}.{0,2,3}.*`, `fc_pose.*`, reference feature/dfnet.py:74-172), so
its checkpoints load unchanged; the constructor never touches the network (the reference
downloads ImageNet VGG-16 weights, feature/dfnet.py:90).  forward() runs on the sm_100a kernels
(implicit-GEMM tcgen05 convolutions, fused bias/ReLU, eval-mode BatchNorm folded into the 5x5
convs); there is no eager fallback.
"""
import ctypes as C
import weakref

import torch
import torch.nn as nn

from . import _lib
from ._lib import check, lib, raw_stream

_VGG16_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M"]
_TAP_CHANNELS = {"conv1_2": 64, "conv3_3": 256, "conv5_3": 512}


def _vgg16_features():
    layers, c = [], 3
    for v in _VGG16_CFG:
        if v == "M":
            layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
        else:
            layers += [nn.Conv2d(c, v, kernel_size=3, padding=1), nn.ReLU(inplace=True)]
            c = v
    return nn.Sequential(*layers)


class AdaptLayers(nn.Module):
    """Adaptation heads (reference feature/dfnet.py:42-72): Conv1x1 -> ReLU -> Conv5x5 -> BatchNorm2d."""

    def __init__(self, hypercolumn_layers, output_dim=128):
        super().__init__()
        for i, name in enumerate(hypercolumn_layers):
            self.add_module(f"adapt_layer_{i}", nn.Sequential(
                nn.Conv2d(_TAP_CHANNELS[name], 64, kernel_size=1, stride=1, padding=0), nn.ReLU(),
                nn.Conv2d(64, output_dim, kernel_size=5, stride=1, padding=2), nn.BatchNorm2d(output_dim)))


class _DfnetHandle:
    def __init__(self, module):
        h = C.c_void_p()
        check(lib.dfb_dfnet_create(len(module.hypercolumn_layers), C.byref(h)))
        self._h = h
        self._fin = weakref.finalize(self, lib.dfb_dfnet_destroy, h)
        self._versions = None
        self._ws = None
        self._bwd_ws = None
        self.n_levels = len(module.hypercolumn_layers)
        self.grad_sync = None      # parallel.GradSync: set by a data-parallel training step (train_on_batch)
        self._bucket_ev = None
        self.last_flat_grad = self.last_grad_views = None

    def refresh(self, module, train=False, bn_train=False, head_train=False, need_feats=True, need_bf16=True):
        """need_feats / need_bf16 = False: the adaptation heads / the bf16 encoder variant are not (re)packed by this load
        (a pose regressor that is re-loaded after every optimizer step never evaluates its heads, and only
        train_dtype="bf16" runs the bf16 encoder); they are brought up to date by the first call that needs them."""
        # state_dict() walks and renames every tensor (~ms): cache the tensors themselves, keyed on their identity,
        # and poll their versions
        key = (id(module), tuple(id(t) for t in module.parameters()), tuple(id(t) for t in module.buffers()))
        if getattr(self, "_sd_key", None) != key:
            self._sd, self._sd_key = dict(module.state_dict(keep_vars=True)), key
        sd = self._sd
        # (the BatchNorm running statistics are buffers: a train-mode forward updates them, which bumps their versions)
        v = [(t.data_ptr(), t._version) for t in sd.values()] + [bool(train)]
        bn_train = bn_train or head_train      # both need the un-folded 5x5 convs and the BatchNorm vectors on the device
        have_bn, have_ht = getattr(self, "_bn_loaded", False), getattr(self, "_ht_loaded", False)
        heads_ok, bf_ok = getattr(self, "_heads_current", False), getattr(self, "_bf_current", False)
        if v[:-1] == (self._versions or [None])[:-1] and (self._versions[-1] or not train) and (have_bn or not bn_train) and \
                (have_ht or not head_train) and (heads_ok or not need_feats) and (bf_ok or not (need_bf16 and train)):
            return
        same = v[:-1] == (self._versions or [None])[:-1]
        # variants that are current stay current only if the weights did not change; what this load packs becomes current
        need_feats = need_feats or bn_train or head_train or (same and heads_ok)
        need_bf16 = need_bf16 or (same and bf_ok)
        # the argument table (tensor list, pointer and size arrays) is rebuilt only when a tensor has moved: with fp32
        # contiguous parameters - the normal case - detach().float().contiguous() is the tensor itself, and building the
        # list anew on every optimizer step was ~0.2 ms of host time in front of the step's first kernel
        tbl_key = (key, tuple(pv[0] for pv in v[:-1]))
        tbl = getattr(self, "_load_tbl", None)
        if tbl is None or tbl[0] != tbl_key:
            names = [f"encoder.{i}" for i, m in enumerate(module.encoder) if isinstance(m, nn.Conv2d)]
            src = []
            for n in names:
                src += [sd[n + ".weight"], sd[n + ".bias"]]
            for l in range(len(module.hypercolumn_layers)):
                p = f"adaptation_layers.adapt_layer_{l}."
                src += [sd[p + "0.weight"], sd[p + "0.bias"], sd[p + "2.weight"], sd[p + "2.bias"], sd[p + "3.weight"],
                        sd[p + "3.bias"], sd[p + "3.running_mean"], sd[p + "3.running_var"]]
            src += [sd["fc_pose.weight"], sd["fc_pose.bias"]]
            direct = all(t.dtype == torch.float32 and t.is_contiguous() for t in src)
            tbl = [tbl_key, src, direct, None, None, None, all(t.is_cuda for t in src)]
            self._load_tbl = tbl
        if tbl[2] and tbl[3] is not None:
            ts, ptrs, numel = tbl[3], tbl[4], tbl[5]
        else:
            ts = [t.detach().float().contiguous() for t in tbl[1]]
            ptrs = (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
            numel = (C.c_int64 * len(ts))(*[t.numel() for t in ts])
            if tbl[2]:
                tbl[3], tbl[4], tbl[5] = ts, ptrs, numel
        eps = module.adaptation_layers.adapt_layer_0[3].eps
        # bit 1: everything is ordered on the legacy default stream -> the library skips its host synchronisation
        on_default = tbl[6] and torch._C._cuda_getCurrentRawStream(ts[0].device.index) == 0
        check(lib.dfb_dfnet_load_ex(self._h, ptrs, numel, len(ts), eps,
                                    (1 if train else 0) | (2 if on_default else 0) | (4 if bn_train else 0) | (8 if head_train else 0)
                                    | (0 if need_feats else 16) | (0 if need_bf16 else 32)))
        self._bn_loaded, self._ht_loaded = bool(bn_train), bool(head_train)
        self._heads_current, self._bf_current = bool(need_feats), bool(need_bf16 and train)
        self._versions = v
        self.n_params = len(ts)

    def invalidate(self):
        """Force a re-upload on the next forward.  Change detection compares (data_ptr, _version) of every tensor; writes
        through `.data` do not bump `_version`, so call this after such writes."""
        self._versions = None

    def forward(self, x, return_feature, single, return_pose, upH, upW, tape=False, bf16=False, bn_train=False, head_train=False,
                skip_levels=0):
        """tape=True keeps every activation in a fresh buffer (returned as 4th value) for `backward`.
        bn_train=True: train-mode BatchNorm in the heads (batch statistics, see `bn_batch_stats`).
        skip_levels: bit l set = feature level l is not wanted (not computed; its slice of the stacks is uninitialised)."""
        if not x.is_cuda:
            raise _lib.DfbError("DFNet input must be a CUDA tensor: the dfnet_b200 hot path has no CPU fallback")
        x = x.detach().float().contiguous()
        B, _, H, W = x.shape
        dev = x.device
        need = C.c_size_t()
        flags = (1 if return_feature else 0) | (2 if single else 0) | (4 if return_pose else 0) | (32 if bn_train else 0) \
            | ((int(skip_levels) & 7) << 8)
        if tape:
            check(lib.dfb_dfnet_tape_bytes(self._h, B, H, W, upH, upW, C.byref(need)))
            ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
            flags |= 8 | (16 if bf16 else 0) | (64 if head_train and return_feature else 0)
        else:
            check(lib.dfb_dfnet_workspace_bytes(self._h, B, H, W, upH, upW, C.byref(need)))
            if self._ws is None or self._ws.numel() < need.value or self._ws.device != dev:
                self._ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
            ws = self._ws
        ft = fr = pose = None
        if return_feature:
            Bs = B if single else B // 2
            ft = torch.empty(self.n_levels, Bs, 128, upH, upW, device=dev)
            fr = None if single else torch.empty(self.n_levels, Bs, 128, upH, upW, device=dev)
        if return_pose:
            pose = torch.empty(B, 12, device=dev)

        def p(t):
            return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
        check(lib.dfb_dfnet_fwd(self._h, p(x), B, H, W, flags, upH, upW, p(ft), p(fr), p(pose), p(ws),
                                ws.numel(), raw_stream()))
        if tape:
            self.last_tape = (ws, flags, (B, H, W, upH, upW))  # debug: tests read the stored activations back
            return ft, fr, pose, (ws, flags)
        return ft, fr, pose

    def bn_batch_stats(self, device):
        """[L,2,128]: batch mean and biased variance of every head's BatchNorm input in the last bn_train forward."""
        out = torch.empty(self.n_levels, 2, 128, device=device)
        check(lib.dfb_dfnet_bn_batch_stats(self._h, C.c_void_p(out.data_ptr()), raw_stream()))
        return out

    def tape_activations(self):
        """Debug: the stored activations of the last taped forward as NCHW fp32 tensors
        {'in', 'act{i}', 'tap{l}', 'mid{l}'} (dfb_debug_dfnet_tape_layout)."""
        ws, flags, (B, H, W, upH, upW) = self.last_tape
        out = (C.c_int64 * 60)()
        check(lib.dfb_debug_dfnet_tape_layout(self._h, B, H, W, upH, upW, out))
        dt = torch.bfloat16 if flags & 16 else torch.float16
        cout = [64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512, 512]

        def view(off, h, w, c, dtype):
            n = B * h * w * c * 2
            return ws[off:off + n].view(dtype).view(B, h, w, c).permute(0, 3, 1, 2).float()
        acts = {"in": view(out[0], H, W, 8, dt)[:, :3]}
        for i in range(13):
            acts[f"act{i}"] = view(out[1 + 4 * i], out[3 + 4 * i], out[4 + 4 * i], cout[i], dt)
        for l, (ci, c) in enumerate(((1, 64), (6, 256), (12, 512))[: self.n_levels]):
            acts[f"tap{l}"] = view(out[53 + l], out[3 + 4 * ci], out[4 + 4 * ci], c, torch.float16)
            acts[f"mid{l}"] = view(out[56 + l], out[3 + 4 * ci], out[4 + 4 * ci], 64, torch.float16)
        return acts

    def backward(self, tape, shape, upH, upW, g_ft, g_fr, level_mask, g_pose, want_gx, param_shapes):
        """dfb_dfnet_bwd.  Returns (g_x [B,3,H,W] or None, list of parameter gradients in load order or None)."""
        ws, flags = tape
        B, H, W = shape
        dev = ws.device
        need = C.c_size_t()
        check(lib.dfb_dfnet_bwd_workspace_bytes(self._h, B, H, W, C.byref(need)))
        if self._bwd_ws is None or self._bwd_ws.numel() < need.value or self._bwd_ws.device != dev:
            self._bwd_ws = torch.empty(need.value, dtype=torch.uint8, device=dev)

        def p(t):
            return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
        g_x = gx_sub = None
        if want_gx:
            g_x = torch.zeros(B, 3, H, W, device=dev)
            gx_sub = g_x
            if not (flags & 2) and (g_ft is None) != (g_fr is None):  # siamese, one stream differentiated
                gx_sub = g_x[B // 2:] if g_ft is None else g_x[:B // 2]
        grads = ptrs = None
        if param_shapes is not None:
            # one flat buffer: a data-parallel step all-reduces it in a single NCCL call (parallel.py)
            sizes = [int(torch.Size(sh).numel()) for sh in param_shapes]
            flat = torch.zeros(sum(sizes), device=dev)
            grads, off = [], 0
            for sh, n in zip(param_shapes, sizes):
                grads.append(flat[off:off + n].view(sh))
                off += n
            # BatchNorm running statistics (entries 6, 7 of every head) are buffers, not parameters
            skip = {26 + 8 * l + k for l in range(self.n_levels) for k in ((6, 7) if flags & 64 else range(8))}
            ptrs = (C.c_void_p * len(grads))(*[None if i in skip else g.data_ptr() for i, g in enumerate(grads)])
            self.last_flat_grad, self.last_grad_views = flat, grads
        sync = self.grad_sync if grads is not None and self.grad_sync is not None and self.grad_sync.world() > 1 else None
        if sync is not None:
            # conv4_1 (encoder layer 7) and everything deeper + fc_pose: 88 % of the bucket, complete early
            if self._bucket_ev is None:
                self._bucket_ev = torch.cuda.Event()
                self._bucket_ev.record()      # creates the cudaEvent_t the library re-records
            check(lib.dfb_dfnet_bwd_bucket_event(self._h, 7, C.c_void_p(self._bucket_ev.cuda_event)))
        check(lib.dfb_dfnet_bwd(self._h, B, H, W, flags, upH, upW, p(g_ft), p(g_fr), level_mask, p(g_pose), p(ws), p(gx_sub),
                                ptrs, 0 if grads is None else len(grads), p(self._bwd_ws), self._bwd_ws.numel(),
                                raw_stream()))
        if sync is not None:
            check(lib.dfb_dfnet_bwd_bucket_event(self._h, 0, None))
            sync.launch(flat, sum(sizes[:14]), self._bucket_ev)
        return g_x, grads


class _DfnetFn(torch.autograd.Function):
    """DFNet forward with the hand-written backward (dfb_dfnet_bwd): gradient of the feature stacks w.r.t. the input
    images (feature net of train_on_batch, frozen) or of the pose w.r.t. the encoder / fc_pose parameters (pose
    regressor).  Reference: autograd through feature/dfnet.py:106-172."""

    @staticmethod
    def forward(ctx, x, handle, cfg, *params):
        return_feature, single, return_pose, upH, upW, level_mask, _bf16, bn_train, head_train, skip = cfg
        need_p = any(t.requires_grad for t in params)
        ft, fr, pose, tape = handle.forward(x, return_feature, single, return_pose, upH, upW, tape=True, bf16=bool(cfg[6]),
                                            bn_train=bn_train, head_train=head_train, skip_levels=skip)
        ctx.handle, ctx.tape, ctx.cfg, ctx.need_p = handle, tape, cfg, need_p
        ctx.xshape = (x.shape[0], x.shape[2], x.shape[3])
        ctx.pshapes = [t.shape for t in params]
        ctx.want_gx = x.requires_grad
        ctx.set_materialize_grads(False)
        outs = [t for t in (ft, fr, pose) if t is not None]
        ctx.slots = [n for n, t in (("ft", ft), ("fr", fr), ("pose", pose)) if t is not None]
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gs):
        g = dict(zip(ctx.slots, gs))
        g_ft, g_fr, g_pose = (None if g.get(k) is None else g[k].float().contiguous() for k in ("ft", "fr", "pose"))
        return_feature, single, return_pose, upH, upW, level_mask, _bf16, bn_train, head_train, _skip = ctx.cfg
        n_out = 3 + len(ctx.pshapes)
        feat_g = g_ft is not None or g_fr is not None
        if g_ft is None and g_fr is None and g_pose is None:
            return (None,) * n_out
        if feat_g and (bn_train or ctx.need_p) and not head_train:
            raise NotImplementedError("feature-path gradients through train-mode BatchNorm / w.r.t. parameters need the head tape "
                                      "(a forward with trainable adaptation heads)")
        if feat_g and head_train and not single and (g_ft is None or g_fr is None):
            # a siamese forward whose loss reads one stream only: the other stream's gradient is zero
            z = torch.zeros_like(g_ft if g_ft is not None else g_fr)
            g_ft, g_fr = (g_ft if g_ft is not None else z), (g_fr if g_fr is not None else z)
        if feat_g and g_pose is not None and not single and (g_ft is None or g_fr is None):
            z = torch.zeros_like(g_ft if g_ft is not None else g_fr)
            g_ft, g_fr = (g_ft if g_ft is not None else z), (g_fr if g_fr is not None else z)
        g_x, grads = ctx.handle.backward(ctx.tape, ctx.xshape, upH, upW, g_ft, g_fr, level_mask, g_pose, ctx.want_gx,
                                         ctx.pshapes if ctx.need_p else None)
        ctx.tape = None
        pg = [None] * len(ctx.pshapes)
        if grads is not None:
            n_lv = (len(grads) - 28) // 8
            skip = {26 + 8 * l + k for l in range(n_lv) for k in ((6, 7) if head_train and feat_g else range(8))}
            pg = [None if i in skip else t for i, t in enumerate(grads)]
        return (g_x, None, None, *pg)


class DFNet(nn.Module):
    """DFNet (reference feature/dfnet.py:74-172): VGG-16 encoder, three hyper-column adaptation heads
    (pre-ReLU conv1_2 / conv3_3 / conv5_3), pose head AdaptiveAvgPool2d(1) -> Linear(512, feat_dim)."""
    hypercolumn_layers = ["conv1_2", "conv3_3", "conv5_3"]
    mean = [0.485, 0.456, 0.406]
    std = [0.229, 0.224, 0.225]

    def __init__(self, feat_dim=12, places365_model_path=""):
        super().__init__()
        if feat_dim != 12:
            raise NotImplementedError("the pose head is a 3x4 matrix (feat_dim=12) everywhere in the reference")
        self.encoder = _vgg16_features()
        self.scales = [1, 4, 16][: len(self.hypercolumn_layers)]
        self.adaptation_layers = AdaptLayers(self.hypercolumn_layers, 128)
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.fc_pose = nn.Linear(512, feat_dim)
        self._handle = None

    def forward(self, x, return_feature=False, isSingleStream=False, return_pose=True, upsampleH=240, upsampleW=427):
        """Reference feature/dfnet.py:106-172 -> (feature_maps, predict): feature_maps is None,
        [stack [L,B,128,H,W]] (single stream) or [target_stack, render_stack] (siamese)."""
        bns = [getattr(self.adaptation_layers, f"adapt_layer_{l}")[3] for l in range(len(self.hypercolumn_layers))]
        # train-mode BatchNorm (run_feature.py without freezeBN, feature/dfnet.py heads under model.train()): batch
        # statistics over the whole batch of this call (target and render images together), running statistics updated
        bn_train = bool(return_feature and bns[0].training)
        if bn_train and (any(b.training != bns[0].training for b in bns) or any(b.momentum is None or not b.track_running_stats for b in bns)):
            raise NotImplementedError("train-mode BatchNorm: all heads in the same mode, momentum set, running statistics tracked")
        if self._handle is None:
            self._handle = _DfnetHandle(self)
        params = self._load_order_params()
        train = torch.is_grad_enabled() and (x.requires_grad or any(t.requires_grad for t in params))
        # the adaptation heads themselves are trained (run_feature.py): their parameters require grad and features are returned
        head_params = [t for l in range(len(self.hypercolumn_layers)) for t in params[26 + 8 * l:26 + 8 * l + 6]]
        head_train = bool(train and return_feature and any(t.requires_grad for t in head_params))
        self._handle.refresh(self, train=train, bn_train=bn_train, head_train=head_train, need_feats=bool(return_feature),
                             need_bf16=getattr(self, "train_dtype", "f16") == "bf16" and not return_feature)
        # want_levels (optional attribute, None = all): hyper-column levels the caller is going to read.  The others are not
        # computed - their slices of the returned stacks are uninitialised - and without a pose the encoder stops after the
        # deepest wanted level.  The reference always evaluates all three; train_on_batch reads feature_matching_lvl only.
        n_lv = len(self.hypercolumn_layers)
        want = getattr(self, "want_levels", None)
        skip = 0 if (want is None or bn_train or not return_feature) else sum(1 << l for l in range(n_lv) if l not in set(want))
        if train:
            levels = getattr(self, "grad_levels", None)
            if skip and levels is None:
                levels = [l for l in range(n_lv) if not (skip >> l) & 1]
            if skip and any((skip >> l) & 1 for l in levels):
                raise ValueError("grad_levels must be a subset of want_levels")
            mask = sum(1 << l for l in (range(len(self.hypercolumn_layers)) if levels is None else levels))
            # train_dtype: "f16" (default: the inference kernels' fp16 forward, bf16 gradients) or "bf16" (BASELINE config[3]:
            # bf16 storage in the pose regressor's forward as well)
            bf16 = getattr(self, "train_dtype", "f16") == "bf16" and not return_feature
            cfg = (bool(return_feature), bool(isSingleStream), bool(return_pose), int(upsampleH), int(upsampleW), mask, bf16, bn_train,
                   head_train, skip)
            outs = list(_DfnetFn.apply(x, self._handle, cfg, *params))
            ft = outs.pop(0) if return_feature else None
            fr = outs.pop(0) if return_feature and not isSingleStream else None
            pose = outs.pop(0) if return_pose else None
        else:
            ft, fr, pose = self._handle.forward(x, return_feature, isSingleStream, return_pose, int(upsampleH), int(upsampleW),
                                                bn_train=bn_train, skip_levels=skip)
        if bn_train:
            # torch.nn.BatchNorm2d bookkeeping: running = (1 - m) running + m stat, with the UNBIASED batch variance
            stats = self._handle.bn_batch_stats(x.device)
            B = x.shape[0]
            with torch.no_grad():
                for l, b in enumerate(bns):
                    sc = self.scales[l]
                    h, w = x.shape[2], x.shape[3]
                    for _ in range({1: 0, 4: 2, 16: 4}[sc]):
                        h, w = h // 2, w // 2
                    n = B * h * w
                    b.running_mean.mul_(1 - b.momentum).add_(stats[l, 0].to(b.running_mean.device), alpha=b.momentum)
                    b.running_var.mul_(1 - b.momentum).add_(stats[l, 1].to(b.running_var.device) * (n / max(n - 1, 1)), alpha=b.momentum)
                    b.num_batches_tracked += 1
        if not return_feature:
            feature_maps = None
        elif isSingleStream:
            feature_maps = [ft]
        else:
            feature_maps = [ft, fr]
        return feature_maps, pose


    def _load_order_params(self):
        """Parameters in dfb_dfnet_load order (BatchNorm running statistics are passed as plain tensors)."""
        ts = []
        for m in self.encoder:
            if isinstance(m, nn.Conv2d):
                ts += [m.weight, m.bias]
        for l in range(len(self.hypercolumn_layers)):
            seq = getattr(self.adaptation_layers, f"adapt_layer_{l}")
            ts += [seq[0].weight, seq[0].bias, seq[2].weight, seq[2].bias, seq[3].weight, seq[3].bias, seq[3].running_mean,
                   seq[3].running_var]
        return ts + [self.fc_pose.weight, self.fc_pose.bias]



class DFNet_s(DFNet):
    """DFNet_s (reference feature/dfnet.py:174-273): only the conv1_2 hyper-column."""
    hypercolumn_layers = ["conv1_2"]


class _CosineLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fr, ft, per_channel):
        a, b = fr.detach().float().contiguous(), ft.detach().float().contiguous()
        Cc = a.shape[0]
        HW = a.numel() // Cc
        loss = torch.empty((), device=a.device)
        ws = torch.empty(max(Cc * 64 * 3, (HW + 255) // 256), device=a.device)
        check(lib.dfb_cosine_loss(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), Cc, HW, int(bool(per_channel)), 1e-6,
                                  C.c_void_p(loss.data_ptr()), C.c_void_p(ws.data_ptr()), ws.numel() * 4,
                                  raw_stream()))
        # the workspace keeps the per-row statistics: the backward reads them instead of recomputing them
        ctx.save_for_backward(a, b, ws)
        ctx.per_channel = bool(per_channel)
        return loss

    @staticmethod
    def backward(ctx, g):
        a, b, ws = ctx.saved_tensors
        Cc = a.shape[0]
        HW = a.numel() // Cc
        g = g.float().contiguous()
        ga = torch.empty_like(a)
        check(lib.dfb_cosine_loss_bwd(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), Cc, HW, int(ctx.per_channel) | 2, 1e-6,
                                      C.c_void_p(g.data_ptr()), C.c_void_p(ga.data_ptr()), C.c_void_p(ws.data_ptr()),
                                      ws.numel() * 4, raw_stream()))
        return ga, None, None  # the target stream is a constant of the step (its inputs carry no gradient)


def feature_loss(feature_rgb, feature_target, img_in=True, per_channel=False):
    """Cosine feature loss (reference feature/direct_feature_matching.py:114-136).  With the default
    per_channel=False the cosine runs over the pixels of each channel (the reference's naming is
    inverted w.r.t. its behaviour); per_channel=True gives the per-pixel cosine."""
    if not feature_rgb.is_cuda:
        raise _lib.DfbError("feature_loss inputs must be CUDA tensors")
    if torch.is_grad_enabled() and feature_rgb.requires_grad:
        return _CosineLossFn.apply(feature_rgb, feature_target, per_channel)
    fr = feature_rgb.detach().float().contiguous()
    ft = feature_target.detach().float().contiguous()
    Cc = fr.shape[0]
    HW = fr.numel() // Cc
    loss = torch.empty((), device=fr.device)
    ws = torch.empty(max(Cc * 64 * 3, (HW + 255) // 256), device=fr.device)
    check(lib.dfb_cosine_loss(C.c_void_p(fr.data_ptr()), C.c_void_p(ft.data_ptr()), Cc, HW, int(bool(per_channel)), 1e-6,
                              C.c_void_p(loss.data_ptr()), C.c_void_p(ws.data_ptr()), ws.numel() * 4,
                              raw_stream()))
    return loss


def preprocess_features_for_loss(feature):
    """[L,B,C,H,W] -> [B,L*C,H,W] (reference feature/direct_feature_matching.py:41-50)."""
    feature = feature.permute(1, 0, 2, 3, 4)
    B, L, Cc, H, W = feature.size()
    return feature.reshape((B, L * Cc, H, W))
