"""NeRF-Hist training: `render(...)` differentiable w.r.t. the NeRF-W parameters and the histogram embeddings
(reference script/run_nerf.py:32-80 `train_on_epoch_nerfw`, script/models/rendering.py:245-337 in train mode).

The inference path evaluates the MLPs in one fused persistent kernel; for training every Linear layer has to expose its
input and its output gradient, so the layers run one by one - each as a 1x1 convolution over the P = N_rays * N_samples
samples on the tcgen05 convolution kernels of the DFNet path: `k_conv_tc` forward (bias / ReLU / pre-activation epilogue),
`k_conv_tc` on the transposed filter with the ReLU-mask epilogue for the data gradient, `k_conv_wgrad` (MN-major
tcgen05 operands, fp32 accumulation) for the weight and bias gradients.  Activations are fp16, gradients bf16, as in the
`train.py` path.  The NeRF-specific kernels around the layers (positional encoding, heads, compositing adjoint) are in
csrc/nerf_train.cu.  The per-RAY part of dir_encoding.0 / transient_encoding.0 (view direction, appearance and transient
codes: 0.1 % of the FLOPs) stays in torch autograd and enters the per-sample layers as a pre-activation addend.

torch supplies memory, the autograd graph edges, the optimizer and torch.sort of the [N, S] depths."""
import ctypes as C

import torch

from . import ops
from ._lib import Copy2d, check, lib, raw_stream


def _p(t):
    # a plain int / None converts to the ABI's void* through argtypes; cheaper than building a c_void_p per argument
    return t.data_ptr() if t is not None else None


def _st():
    return raw_stream()


def _copy_table(items):
    """ctypes table of DfbCopy2d items (src, dst, rows, cols, src_ld, dst_ld)."""
    tbl = (Copy2d * len(items))()
    for k, it in enumerate(items):
        tbl[k] = Copy2d(*it)
    return tbl


def _copy_batch(tbl):
    check(lib.dfb_copy2d_batch(tbl, len(tbl), _st()))


class _Layer:
    """One Linear layer as a pair of 1x1 convolution handles: forward (fp16 operands) and data gradient (bf16)."""

    def __init__(self, cin, cout, dev, dgrad=True):
        self.cin, self.cout = cin, cout
        self.cout_pad = (cout + 63) // 64 * 64
        self.cin_pad = (cin + 7) // 8 * 8
        self.w = torch.zeros(self.cout_pad, cin, device=dev)       # padded fp32 staging of the parameters
        self.b = torch.zeros(self.cout_pad, device=dev)
        self.fwd = C.c_void_p()
        check(lib.dfb_conv_create_ex(cin, self.cout_pad, 1, 1, _p(self.w), _p(self.b), None, None, 0, 0, C.byref(self.fwd)))
        self.dg = None
        if dgrad:
            self.dg = C.c_void_p()
            check(lib.dfb_conv_create_ex(cin, self.cout_pad, 1, 1, _p(self.w), None, None, None, 1, 1, C.byref(self.dg)))
        self.dg_out = (cin + 63) // 64 * 64                        # channels of the data gradient's output

    def stage(self, items, w, b, row0=0):
        """Queue the copies of w [rows, cols] (a parameter or a column slice of one; cols <= cin, the rest of the staging
        row stays zero) and b [rows] or None into the padded staging tensors, from row `row0` on."""
        w = w.detach()
        assert w.dtype == torch.float32 and w.stride(1) == 1 and row0 + w.shape[0] <= self.cout_pad and w.shape[1] <= self.cin
        items.append((w.data_ptr(), self.w.data_ptr() + 4 * row0 * self.cin, w.shape[0], w.shape[1], w.stride(0), self.cin))
        if b is not None:
            b = b.detach()
            assert b.dtype == torch.float32 and b.is_contiguous()
            items.append((b.data_ptr(), self.b.data_ptr() + 4 * row0, 1, b.shape[0], b.shape[0], b.shape[0]))

    def update(self):
        """Re-pack the weight images from the staging tensors (queued while a dfb_conv_pack_begin batch is open)."""
        check(lib.dfb_conv_update(self.fwd, _p(self.w), _p(self.b), None, None, _st()))
        if self.dg is not None:
            check(lib.dfb_conv_update(self.dg, _p(self.w), None, None, None, _st()))

    def load(self, w, b):
        """w [cout, cin] (a parameter or a column slice of one), b [cout] or None (one layer, outside a refresh)."""
        items = []
        self.stage(items, w, b)
        _copy_batch(_copy_table(items))
        self.update()

    def __del__(self):
        try:
            lib.dfb_conv_destroy(self.fwd)
            if self.dg is not None:
                lib.dfb_conv_destroy(self.dg)
        except Exception:  # noqa: BLE001
            pass


def _conv(handle, x, Hh, relu, out=None, tap=None, nchw=None, mask=None, addend=None, out_bf=None):
    """out_bf: bf16 copy of `out`, written by the same epilogue (the weight-gradient kernel's operand type)."""
    if out_bf is not None:
        check(lib.dfb_conv_fwd_ex2(handle, _p(x), 1, Hh, 8, int(relu), _p(out), _p(tap), _p(nchw), _p(mask), _p(addend), _p(out_bf), _st()))
    else:
        check(lib.dfb_conv_fwd_ex(handle, _p(x), 1, Hh, 8, int(relu), _p(out), _p(tap), _p(nchw), _p(mask), _p(addend), _st()))


def trainable_shape(net):
    """True for the network shapes the training executor covers (8 layers, netwidth 128 / 256, skip at 4, 10 xyz bands)."""
    return net.D == 8 and list(net.skips) == [4] and net.W % 128 == 0 and net.in_channels_xyz == 63


class NetTrainer:
    """Layer-wise training executor of one NeRFW module (coarse: static branch; fine: static + transient)."""

    def __init__(self, net):
        dev = next(net.parameters()).device
        W, H2 = net.W, net.W // 2
        if not trainable_shape(net):
            raise NotImplementedError("NeRF training on the B200 path covers the 8-layer networks with netwidth 128 or 256 "
                                      "(the reference default and the benchmark shape), skip at layer 4, 10 xyz bands")
        self.net, self.dev, self.W, self.H2, self.fine = net, dev, W, H2, net.typ == "fine"
        self.L = {}
        for i in range(8):
            self.L[f"x{i}"] = _Layer(64 if i == 0 else W, W, dev, dgrad=i > 0)
        self.L["xpe"] = _Layer(64, W, dev, dgrad=False)              # the positional-encoding columns of the skip layer
        self.L["sigma"] = _Layer(W, 1, dev)
        self.L["final"] = _Layer(W, W, dev)
        self.L["dir"] = _Layer(W, H2, dev)
        self.L["rgb"] = _Layer(H2, 3, dev)
        if self.fine:
            self.L["t0"] = _Layer(W, H2, dev)
            for k in (1, 2, 3):
                self.L[f"t{k}"] = _Layer(H2, H2, dev)
            self.L["th"] = _Layer(H2, 5, dev)                        # transient rgb(3), sigma, beta
        self._versions = None
        self._copy_tbl = None  # (parameter pointers, DfbCopy2d table of the parameter -> staging copies)
        self._update_tbl = None
        self._bufs = {}
        self._live = False     # a tape of this executor is waiting for its backward

    # ---- parameters -------------------------------------------------------------------------------------------------
    def params(self):
        return list(self.net.parameters())

    def refresh(self):
        n = self.net
        v = [(p.data_ptr(), p._version) for p in n.parameters()]
        if v == self._versions:
            return
        # parameters -> padded staging tensors as ONE batched strided copy (the table is rebuilt only when a parameter has
        # moved), then every layer's (re)packing request as ONE launch (the staging tensors stay untouched until
        # dfb_conv_pack_end).  Issued tensor by tensor this was ~70 copies and 116 packing launches per training step.
        ptrs = tuple(pv[0] for pv in v)
        if self._copy_tbl is None or self._copy_tbl[0] != ptrs:
            items = []
            self._stage_layers(n, items)
            self._copy_tbl = (ptrs, _copy_table(items))
        _copy_batch(self._copy_tbl[1])
        if self._update_tbl is None:      # handles and staging tensors never move: (forward, data-gradient) x layers, built once
            hs, ws, bs = [], [], []
            for layer in self.L.values():
                hs.append(layer.fwd.value), ws.append(layer.w.data_ptr()), bs.append(layer.b.data_ptr())
                if layer.dg is not None:
                    hs.append(layer.dg.value), ws.append(layer.w.data_ptr()), bs.append(None)
            arr = C.c_void_p * len(hs)
            self._update_tbl = (arr(*hs), arr(*ws), arr(*bs), len(hs))
        t = self._update_tbl
        check(lib.dfb_conv_update_many(t[0], t[1], t[2], t[3], _st()))
        self._versions = v

    def _stage_layers(self, n, items):
        W = self.W
        for i in range(8):
            lin = getattr(n, f"xyz_encoding_{i + 1}")[0]
            if i == 4:
                self.L["xpe"].stage(items, lin.weight[:, :63], None)     # 63 of 64 staging columns, the last stays zero
                self.L["x4"].stage(items, lin.weight[:, 63:], lin.bias)
            else:
                self.L[f"x{i}"].stage(items, lin.weight, lin.bias)       # layer 0: 63 of 64 columns
        self.L["sigma"].stage(items, n.static_sigma[0].weight, n.static_sigma[0].bias)
        self.L["final"].stage(items, n.xyz_encoding_final.weight, n.xyz_encoding_final.bias)
        self.L["dir"].stage(items, n.dir_encoding[0].weight[:, :W], n.dir_encoding[0].bias)
        self.L["rgb"].stage(items, n.static_rgb[0].weight, n.static_rgb[0].bias)
        if self.fine:
            te = n.transient_encoding
            self.L["t0"].stage(items, te[0].weight[:, :W], te[0].bias)
            for k, idx in ((1, 2), (2, 4), (3, 6)):
                self.L[f"t{k}"].stage(items, te[idx].weight, te[idx].bias)
            row = 0
            for head in (n.transient_rgb[0], n.transient_sigma[0], n.transient_beta[0]):   # rows 0..2 | 3 | 4 of one layer
                self.L["th"].stage(items, head.weight, head.bias, row0=row)
                row += head.weight.shape[0]

    # ---- buffers ----------------------------------------------------------------------------------------------------
    def begin(self):
        """Start of a forward.  A second forward before the first one's backward (chunked ray batches) gets fresh buffers;
        the outstanding tape keeps the old ones alive."""
        if self._live:
            self._bufs = {}

    def buf(self, name, rows, cols, dtype):
        key = (name, rows, cols, dtype)
        t = self._bufs.get(key)
        if t is None:
            t = torch.zeros(rows, cols, device=self.dev, dtype=dtype)   # zero: rows past P never carry gradient
            self._bufs[key] = t
        return t

    # ---- forward ----------------------------------------------------------------------------------------------------
    def forward(self, pe16, P, rb_dir, rb_t, S, twin=False, pe_bf=None):
        """pe16 [Pp, 64] fp16 (Pp = P rounded up to 8), rb_dir / rb_t [N, H2] fp32 per-ray pre-activation addends
        (rb_t: fine only) -> raw [P, 4 | 9] fp32 and the tape."""
        self.refresh()
        self._live = True
        W, H2, Pp = self.W, self.H2, pe16.shape[0]
        Hh = Pp // 8
        f16 = torch.float16
        N = rb_dir.shape[0]
        # with gradients wanted every activation that is a weight-gradient operand gets a bf16 twin from the same epilogue
        # (26 separate conversion launches per step otherwise)
        bfd = torch.bfloat16
        tw = {} if pe_bf is None else {"pe": pe_bf}

        def bf(name, cols):
            if not twin:
                return None
            tw[name] = self.buf("bf_" + name, Pp, cols, bfd)
            return tw[name]
        h = []
        x = pe16
        skip = self.buf("skip", Pp, W, f16)
        for i in range(8):
            o = self.buf(f"h{i}", Pp, W, f16)
            if i == 4:
                _conv(self.L["xpe"].fwd, pe16, Hh, 0, tap=skip)                       # W_pe . pe (pre-activation part)
                _conv(self.L["x4"].fwd, x, Hh, 1, out=o, addend=skip, out_bf=bf(f"h{i}", W))
            else:
                _conv(self.L[f"x{i}"].fwd, x, Hh, 1, out=o, out_bf=bf(f"h{i}", W))
            h.append(o)
            x = o
        planes = self.buf("planes", 3 * 64, Pp, torch.float32)                        # fp32 head pre-activations [ch][P]
        _conv(self.L["sigma"].fwd, x, Hh, 0, nchw=planes[0:64])
        final = self.buf("final", Pp, W, f16)
        _conv(self.L["final"].fwd, x, Hh, 0, out=final, out_bf=bf("final", W))
        add_d = self.buf("add_d", Pp, H2, f16)
        check(lib.dfb_rows_expand16(_p(rb_dir.contiguous()), N, S, H2, _p(add_d), _st()))
        dirh = self.buf("dirh", Pp, H2, f16)
        _conv(self.L["dir"].fwd, final, Hh, 1, out=dirh, addend=add_d, out_bf=bf("dirh", H2))
        _conv(self.L["rgb"].fwd, dirh, Hh, 0, nchw=planes[64:128])
        t = []
        Cc = 4
        if self.fine:
            Cc = 9
            add_t = self.buf("add_t", Pp, H2, f16)
            check(lib.dfb_rows_expand16(_p(rb_t.contiguous()), N, S, H2, _p(add_t), _st()))
            o = self.buf("t0", Pp, H2, f16)
            _conv(self.L["t0"].fwd, final, Hh, 1, out=o, addend=add_t, out_bf=bf("t0", H2))
            t.append(o)
            for k in (1, 2, 3):
                o2 = self.buf(f"t{k}", Pp, H2, f16)
                _conv(self.L[f"t{k}"].fwd, t[-1], Hh, 1, out=o2, out_bf=bf(f"t{k}", H2))
                t.append(o2)
            _conv(self.L["th"].fwd, t[-1], Hh, 0, nchw=planes[128:192])
        raw = torch.empty(P, Cc, device=self.dev)
        check(lib.dfb_nerf_heads_fwd(_p(planes[0:64]), _p(planes[64:128]), _p(planes[128:192]) if self.fine else None, P, Pp, Cc,
                                     _p(raw), _st()))
        return raw, dict(pe=pe16, h=h, final=final, dirh=dirh, t=t, P=P, Pp=Pp, S=S, N=N, bf=tw)

    # ---- backward ---------------------------------------------------------------------------------------------------
    def backward(self, tape, raw, g_raw):
        """-> (g_rb_dir [N,H2], g_rb_t [N,H2] or None, parameter gradients in net.parameters() order)."""
        W, H2, P, Pp, S, N = self.W, self.H2, tape["P"], tape["Pp"], tape["S"], tape["N"]
        self._live = False
        Hh = Pp // 8
        bf = torch.bfloat16
        Cc = 9 if self.fine else 4
        st = _st()

        twins = tape.get("bf", {})

        def cast(x, name):
            if name in twins:        # written by the forward's epilogue
                return twins[name]
            o = self.buf("bf_" + name, x.shape[0], x.shape[1], bf)
            check(lib.dfb_cast_f16_bf16(_p(x), _p(o), x.numel(), st))
            return o

        # all weight / bias gradients of this backward live in ONE zeroed buffer (one memset instead of two per layer)
        need = sum(l.cout_pad * l.cin + l.cout_pad + 4 for l in self.L.values())
        flat = torch.zeros(need, device=self.dev)
        cursor = [0]

        def wgrad(gO, X, cin, cout_pad):
            o0 = cursor[0]
            dW = flat[o0:o0 + cout_pad * cin].view(cout_pad, cin)
            dB = flat[o0 + cout_pad * cin:o0 + cout_pad * cin + cout_pad]
            cursor[0] = o0 + (cout_pad * cin + cout_pad + 3) // 4 * 4
            assert cursor[0] <= need
            check(lib.dfb_conv_wgrad_acc(_p(gO), _p(X), 1, Hh, 8, cin, X.shape[1], cout_pad, 1, 1, _p(dW), _p(dB), st))
            return dW, dB

        gs = self.buf("g_sig", Pp, 64, bf)
        gr = self.buf("g_rgb", Pp, 64, bf)
        gt = self.buf("g_tr", Pp, 64, bf) if self.fine else None
        check(lib.dfb_nerf_heads_bwd(_p(raw), _p(g_raw.contiguous()), P, Cc, _p(gs), _p(gr), _p(gt), st))
        G = {}
        final_bf = cast(tape["final"], "final")
        h7_bf = cast(tape["h"][7], "h7")
        # ---- transient branch ----
        g_final_t = None
        g_rb_t = None
        if self.fine:
            t = tape["t"]
            G["th"] = wgrad(gt, cast(t[3], "t3"), H2, 64)
            g = self.buf("g_t3", Pp, H2, bf)
            _conv(self.L["th"].dg, gt, Hh, 0, out=g, mask=t[3])
            for k in (3, 2, 1):
                G[f"t{k}"] = wgrad(g, cast(t[k - 1], f"t{k - 1}"), H2, H2)
                g2 = self.buf(f"g_t{k - 1}", Pp, H2, bf)
                _conv(self.L[f"t{k}"].dg, g, Hh, 0, out=g2, mask=t[k - 1])
                g = g2
            G["t0"] = wgrad(g, final_bf, W, H2)
            g_rb_t = torch.empty(N, H2, device=self.dev)
            check(lib.dfb_rows_reduce_bf16(_p(g), N, S, H2, _p(g_rb_t), st))
            g_final_t = self.buf("g_final_t", Pp, W, bf)
            _conv(self.L["t0"].dg, g, Hh, 0, out=g_final_t)
        # ---- static colour branch ----
        G["rgb"] = wgrad(gr, cast(tape["dirh"], "dirh"), H2, 64)
        g_dir = self.buf("g_dir", Pp, H2, bf)
        _conv(self.L["rgb"].dg, gr, Hh, 0, out=g_dir, mask=tape["dirh"])
        G["dir"] = wgrad(g_dir, final_bf, W, H2)
        g_rb_d = torch.empty(N, H2, device=self.dev)
        check(lib.dfb_rows_reduce_bf16(_p(g_dir), N, S, H2, _p(g_rb_d), st))
        g_final = self.buf("g_final", Pp, W, bf)
        _conv(self.L["dir"].dg, g_dir, Hh, 0, out=g_final, addend=g_final_t)
        # ---- xyz_encoding_final, sigma, trunk ----
        G["final"] = wgrad(g_final, h7_bf, W, W)
        G["sigma"] = wgrad(gs, h7_bf, W, 64)
        ga = self.buf("g_h7a", Pp, W, bf)
        _conv(self.L["sigma"].dg, gs, Hh, 0, out=ga, mask=tape["h"][7])
        g = self.buf("g_h7", Pp, W, bf)
        _conv(self.L["final"].dg, g_final, Hh, 0, out=g, mask=tape["h"][7], addend=ga)
        pe_bf = cast(tape["pe"], "pe")
        for i in range(7, 0, -1):
            xin = cast(tape["h"][i - 1], f"h{i - 1}") if i - 1 != 7 else h7_bf
            G[f"x{i}"] = wgrad(g, xin, W, W)
            if i == 4:
                G["xpe"] = wgrad(g, pe_bf, 64, W)
            g2 = self.buf(f"g_h{i - 1}", Pp, W, bf)
            _conv(self.L[f"x{i}"].dg, g, Hh, 0, out=g2, mask=tape["h"][i - 1])
            g = g2
        G["x0"] = wgrad(g, pe_bf, 64, W)
        # ---- scatter into net.parameters() order ----
        n = self.net
        out = {}
        for i in range(8):
            lin = getattr(n, f"xyz_encoding_{i + 1}")[0]
            dW, dB = G[f"x{i}"]
            if i == 0:
                out[lin.weight] = dW[:, :63]
            elif i == 4:
                out[lin.weight] = torch.cat([G["xpe"][0][:, :63], dW], 1)
            else:
                out[lin.weight] = dW
            out[lin.bias] = dB
        out[n.xyz_encoding_final.weight], out[n.xyz_encoding_final.bias] = G["final"]
        wd = n.dir_encoding[0].weight
        gd = torch.zeros_like(wd)
        gd[:, :W] = G["dir"][0]
        out[wd], out[n.dir_encoding[0].bias] = gd, G["dir"][1]
        out[n.static_sigma[0].weight], out[n.static_sigma[0].bias] = G["sigma"][0][:1], G["sigma"][1][:1]
        out[n.static_rgb[0].weight], out[n.static_rgb[0].bias] = G["rgb"][0][:3], G["rgb"][1][:3]
        if self.fine:
            te = n.transient_encoding
            wt = te[0].weight
            gtw = torch.zeros_like(wt)
            gtw[:, :W] = G["t0"][0]
            out[wt], out[te[0].bias] = gtw, G["t0"][1]
            for k, idx in ((1, 2), (2, 4), (3, 6)):
                out[te[idx].weight], out[te[idx].bias] = G[f"t{k}"]
            dW, dB = G["th"]
            out[n.transient_rgb[0].weight], out[n.transient_rgb[0].bias] = dW[0:3], dB[0:3]
            out[n.transient_sigma[0].weight], out[n.transient_sigma[0].bias] = dW[3:4], dB[3:4]
            out[n.transient_beta[0].weight], out[n.transient_beta[0].bias] = dW[4:5], dB[4:5]
        return g_rb_d, g_rb_t, [out[p].contiguous() for p in n.parameters()]


def trainer_for(net):
    t = net.__dict__.get("_dfb_trainer")
    if t is None:
        t = NetTrainer(net)
        net.__dict__["_dfb_trainer"] = t
    return t


class _MLPFn(torch.autograd.Function):
    """raw = NeRFW(samples) with the hand-written layer-wise backward; differentiable w.r.t. the per-ray addends and the
    module's parameters (the sample positions are constants: rays are data, depths are detached, rendering.py:302)."""

    @staticmethod
    def forward(ctx, tr, pe, P, S, rb_dir, rb_t, *params):
        # grad mode is always off inside Function.forward: whether a backward will follow is what needs_input_grad says
        pe16, pe_bf = pe
        raw, tape = tr.forward(pe16, P, rb_dir.detach(), None if rb_t is None else rb_t.detach(), S, twin=any(ctx.needs_input_grad),
                               pe_bf=pe_bf)
        ctx.tr, ctx.tape, ctx.has_t = tr, tape, rb_t is not None
        ctx.save_for_backward(raw)
        return raw

    @staticmethod
    def backward(ctx, g_raw):
        (raw,) = ctx.saved_tensors
        g_d, g_t, gp = ctx.tr.backward(ctx.tape, raw, g_raw.float())
        ctx.tape = None
        return (None, None, None, None, g_d, g_t if ctx.has_t else None, *gp)


class _CompositeFn(torch.autograd.Function):
    """raw2outputs_NeRFW in train mode (rendering.py:132-243) with dfb_raw2outputs_bwd: the outputs NerfWLoss reads carry
    gradient (rgb; fine: beta, transient_sigmas), disp / acc / weights / depth do not."""

    @staticmethod
    def forward(ctx, raw, z, typ, beta_min, noise, noise_std):
        o = ops.raw2outputs(raw.detach(), z, typ, False, beta_min, noise=noise, raw_noise_std=noise_std)
        ctx.save_for_backward(raw.detach(), z, noise if noise is not None else torch.empty(0))
        ctx.typ, ctx.noise_std = typ, float(noise_std)
        outs = (o["rgb"], o["disp"], o["acc"], o["weights"], o["beta"],
                o["transient_sigmas"] if typ == "fine" else torch.empty(0, device=raw.device))
        ctx.mark_non_differentiable(outs[1], outs[2], outs[3])
        return outs

    @staticmethod
    def backward(ctx, g_rgb, g_disp, g_acc, g_w, g_beta, g_tsig):
        raw, z, noise = ctx.saved_tensors
        N, S, Cc = raw.shape
        fine = ctx.typ == "fine"
        g_raw = torch.empty_like(raw)

        def c(t):
            return None if t is None else t.float().contiguous()
        g_rgb, g_beta, g_tsig = c(g_rgb), c(g_beta) if fine else None, c(g_tsig) if fine and g_tsig is not None and g_tsig.numel() else None
        check(lib.dfb_raw2outputs_bwd(_p(raw), _p(z), N, S, Cc, _p(noise) if noise.numel() else None, ctx.noise_std, _p(g_rgb),
                                      _p(g_beta), _p(g_tsig), _p(g_raw), _st()))
        return g_raw, None, None, None, None, None


def _embed(x, L):
    """[x, sin(x 2^0), cos(x 2^0), ..., sin(x 2^(L-1)), cos(x 2^(L-1))] along the last axis (models/nerfw.py Embedding with
    log-spaced bands), all bands in one pass: 5 launches instead of 4 L + 1, same values element for element."""
    xb = x[..., None, :] * (2.0 ** torch.arange(L, device=x.device, dtype=x.dtype))[:, None]      # [..., L, C]
    sc = torch.stack([torch.sin(xb), torch.cos(xb)], -2)                                            # [..., L, 2, C]
    return torch.cat([x, sc.reshape(*x.shape[:-1], -1)], -1)


_T_VALS = {}


def render_rays_train(ray_batch, network_fn, network_fine, embedding_a, embedding_t, N_samples, N_importance, perturb=0.,
                      raw_noise_std=0., lindisp=False, retraw=False, pytest=False):
    """render_rays (rendering.py:245-337) with test_time=False, differentiable w.r.t. the networks and embeddings."""
    dev = ray_batch.device
    N = ray_batch.shape[0]
    rays = ray_batch.detach().float().contiguous()
    near, far, viewdirs, hist = rays[:, 6:7], rays[:, 7:8], rays[:, 8:11], rays[:, 11:]
    t_vals = _T_VALS.get((N_samples, dev))           # constant of the run: one upload instead of a pageable copy per step
    if t_vals is None:
        t_vals = _T_VALS[(N_samples, dev)] = ops.linspace(0., 1., N_samples).to(dev)
    z = near * (1. - t_vals) + far * t_vals if not lindisp else 1. / (1. / near * (1. - t_vals) + 1. / far * t_vals)
    z = z.expand(N, N_samples)
    if perturb > 0.:
        mids = .5 * (z[..., 1:] + z[..., :-1])
        upper, lower = torch.cat([mids, z[..., -1:]], -1), torch.cat([z[..., :1], mids], -1)
        z = lower + (upper - lower) * torch.rand(z.shape, device=dev)
    z = z.contiguous()
    noise = torch.randn(N, N_samples, device=dev)     # drawn even when raw_noise_std == 0 (rendering.py:173)
    dir_pe = _embed(viewdirs, 4)

    def run(net, zz, fine):
        tr = trainer_for(net)
        tr.begin()
        S = zz.shape[1]
        P = N * S
        Pp = (P + 7) // 8 * 8
        pe16 = tr.buf("pe_in", Pp, 64, torch.float16)
        # with a backward to come the encoding is also written as bf16, the first layer's weight-gradient operand
        pe_bf = tr.buf("bf_pe", Pp, 64, torch.bfloat16) if torch.is_grad_enabled() else None
        check(lib.dfb_embed_xyz16_ex(_p(rays), rays.shape[1], _p(zz), N, S, 10, 64, _p(pe16), _p(pe_bf), _st()))
        W = net.W
        if fine:
            ts = hist.long()
            a = embedding_a(ts).reshape(N, -1)
            t = embedding_t(ts).reshape(N, -1)
            rb_d = torch.cat([dir_pe, a], -1) @ net.dir_encoding[0].weight[:, W:].t()
            rb_t = t @ net.transient_encoding[0].weight[:, W:].t()
        else:
            rb_d = dir_pe @ net.dir_encoding[0].weight[:, W:].t()
            rb_t = None
        raw = _MLPFn.apply(tr, (pe16, pe_bf), P, S, rb_d, rb_t, *tr.params())
        return raw.reshape(N, S, -1)

    raw_c = run(network_fn, z, False)
    rgb0, disp0, acc0, w_c, _, _ = _CompositeFn.apply(raw_c, z, "coarse", 0.1, noise if raw_noise_std > 0. else None, raw_noise_std)
    ret = {"rgb_map": rgb0, "disp_map": disp0, "acc_map": acc0}
    raw = raw_c
    if N_importance > 0:
        z_mid = .5 * (z[..., 1:] + z[..., :-1])
        u = None
        if perturb > 0.:
            if pytest:
                import numpy as np
                np.random.seed(0)
                u = torch.tensor(np.random.rand(N, N_importance), dtype=torch.float32)
            else:
                u = torch.rand(N, N_importance, device=dev)
        z_samples, _ = ops.sample_pdf(z_mid.contiguous(), w_c[..., 1:-1].contiguous(), N_importance, det=(perturb == 0.), u=u)
        z_all, _ = torch.sort(torch.cat([z, z_samples], -1), -1)
        raw = run(network_fine, z_all.contiguous(), True)
        rgb, disp, acc, _, beta, tsig = _CompositeFn.apply(raw, z_all.contiguous(), "fine", float(network_fine.beta_min), None, 0.)
        ret = {"rgb_map": rgb, "disp_map": disp, "acc_map": acc, "rgb0": rgb0, "disp0": disp0, "acc0": acc0,
               "z_std": torch.std(z_samples, dim=-1, unbiased=False), "transient_sigmas": tsig, "beta": beta}
    if retraw:
        ret["raw"] = raw
    return ret


def train_on_batch_nerfw(args, target, pose, img_idx, H, W, focal, N_rand, optimizer, loss_func, global_step, render_kwargs_train,
                         near=0., far=1., select_inds=None):
    """One iteration of the reference's `train_on_epoch_nerfw` (script/run_nerf.py:33-77): N_rand random rays of one image,
    render in train mode, NeRF-W loss, backward, Adam step, exponential learning-rate decay.
    target [3,H,W] (or [H,W,3]) in [0,1], pose [3,4] / [12] c2w, img_idx [1,hist_bin].  The pixel choice follows the
    reference (np.random.choice without replacement on the host); `select_inds` overrides it (tests).
    -> (loss, psnr) device scalars."""
    import numpy as np
    from .rendering import render
    dev = next(render_kwargs_train["network_fn"].parameters()).device
    target = target.to(dev)
    if target.shape[0] == 3 and target.dim() == 3:
        target = target.permute(1, 2, 0)
    pose = pose.reshape(3, 4).to(dev)
    rays_o, rays_d = ops.get_rays(int(H), int(W), float(focal), pose)
    if N_rand is not None:
        if select_inds is None:
            select_inds = np.random.choice(int(H) * int(W), size=[N_rand], replace=False)
        sel = torch.as_tensor(select_inds, device=dev, dtype=torch.long)
        rays_o, rays_d = rays_o.reshape(-1, 3)[sel], rays_d.reshape(-1, 3)[sel]
        target_s = target.reshape(-1, 3)[sel]
    else:
        rays_o, rays_d, target_s = rays_o.reshape(-1, 3), rays_d.reshape(-1, 3), target.reshape(-1, 3)
    rgb, disp, acc, extras = render(H, W, focal, chunk=args.chunk, rays=(rays_o, rays_d), retraw=True, img_idx=img_idx.to(dev),
                                    near=near, far=far, **render_kwargs_train)
    optimizer.zero_grad()
    results = {"rgb_fine": rgb, "rgb_coarse": extras["rgb0"], "beta": extras["beta"], "transient_sigmas": extras["transient_sigmas"]}
    loss_d = loss_func(results, target_s)
    loss = sum(l for l in loss_d.values())
    with torch.no_grad():
        mse_f = getattr(loss_func, "last_mse_fine", None)       # NerfWLoss' fused pass already has it
        psnr = -10. * torch.log10(mse_f if mse_f is not None else torch.mean((rgb - target_s) ** 2))
    loss.backward()
    optimizer.step()
    decay_rate, decay_steps = 0.1, args.lrate_decay * 1000
    new_lrate = args.lrate * (decay_rate ** (global_step / decay_steps))
    for param_group in optimizer.param_groups:
        param_group["lr"] = new_lrate
    return loss.detach(), psnr
