"""Losses of the NeRF-Hist training loop (reference script/models/losses.py:5-57): same classes, constructor arguments and
return structure.  They act on per-ray tensors ([N,3], [N], [N,S]) - a few thousand values per step - and are plain tensor
expressions; the renderer they differentiate into is dfnet_b200.nerf_train."""
import torch
from torch import nn


class ColorLoss(nn.Module):
    def __init__(self, coef=1):
        super().__init__()
        self.coef = coef
        self.loss = nn.MSELoss(reduction="mean")

    def forward(self, inputs, targets):
        loss = self.loss(inputs["rgb_coarse"], targets)
        if "rgb_fine" in inputs:
            loss = loss + self.loss(inputs["rgb_fine"], targets)
        return self.coef * loss


class _NerfWLossFn(torch.autograd.Function):
    """The four NeRF-W terms (and the fine pass' mean squared error) from dfb_nerfw_loss_fwd / _bwd: 3 launches for what is
    ~60 as tensor expressions.  CUDA fp32 inputs; NerfWLoss keeps the tensor expressions for everything else."""

    @staticmethod
    def forward(ctx, rgb_c, rgb_f, beta, tsig, targets, coef, lambda_u):
        from ._lib import check, lib, raw_stream
        rgb_c, rgb_f, beta, tsig, targets = (t.detach().contiguous() for t in (rgb_c, rgb_f, beta, tsig, targets))
        N, S = tsig.shape
        ws = torch.empty(lib.dfb_nerfw_loss_workspace_bytes() // 8, device=rgb_c.device, dtype=torch.float64)
        out = torch.empty(5, device=rgb_c.device)
        check(lib.dfb_nerfw_loss_fwd(rgb_c.data_ptr(), rgb_f.data_ptr(), beta.data_ptr(), tsig.data_ptr(), targets.data_ptr(), N, S,
                                     float(coef), float(lambda_u), ws.data_ptr(), out.data_ptr(), raw_stream()))
        ctx.save_for_backward(rgb_c, rgb_f, beta, targets)
        ctx.dims, ctx.coef, ctx.lambda_u = (N, S), float(coef), float(lambda_u)
        c_l, f_l, b_l, s_l, mse_f = out.unbind(0)
        ctx.mark_non_differentiable(mse_f)
        return c_l, f_l, b_l, s_l, mse_f

    @staticmethod
    def backward(ctx, g_c, g_f, g_b, g_s, _g_mse):
        from ._lib import check, lib, raw_stream
        rgb_c, rgb_f, beta, targets = ctx.saved_tensors
        N, S = ctx.dims
        need = ctx.needs_input_grad
        g_c, g_f, g_b, g_s = (g.float().contiguous() for g in (g_c, g_f, g_b, g_s))
        o_c = torch.empty_like(rgb_c) if need[0] else None
        o_f = torch.empty_like(rgb_f) if need[1] else None
        o_b = torch.empty_like(beta) if need[2] else None
        o_s = torch.empty(N, S, device=beta.device) if need[3] else None

        def p(t):
            return None if t is None else t.data_ptr()
        check(lib.dfb_nerfw_loss_bwd(rgb_c.data_ptr(), rgb_f.data_ptr(), beta.data_ptr(), targets.data_ptr(), N, S, ctx.coef, ctx.lambda_u,
                                     g_c.data_ptr(), g_f.data_ptr(), g_b.data_ptr(), g_s.data_ptr(), p(o_c), p(o_f), p(o_b), p(o_s),
                                     raw_stream()))
        return o_c, o_f, o_b, o_s, None, None, None


def _fused_ok(inputs, targets):
    keys = ("rgb_coarse", "rgb_fine", "beta", "transient_sigmas")
    if not all(k in inputs for k in keys):
        return False
    ts = [inputs[k] for k in keys] + [targets]
    if not all(t.is_cuda and t.dtype == torch.float32 for t in ts):
        return False
    n = targets.shape[0]
    return (targets.dim() == 2 and targets.shape[1] == 3 and inputs["rgb_coarse"].shape == targets.shape and
            inputs["rgb_fine"].shape == targets.shape and inputs["beta"].shape == (n,) and inputs["transient_sigmas"].dim() == 2 and
            inputs["transient_sigmas"].shape[0] == n and n >= 1 and inputs["transient_sigmas"].shape[1] >= 1)


class NerfWLoss(nn.Module):
    """Equation 13 of NeRF-W: c_l coarse colour, f_l fine colour weighted by 1 / (2 beta^2), b_l = 3 + mean(log beta),
    s_l = lambda_u * mean(transient_sigmas)."""

    def __init__(self, coef=1, lambda_u=0.01):
        super().__init__()
        self.coef = coef
        self.lambda_u = lambda_u

    def forward(self, inputs, targets, use_hier_rgbs=False, rgb_h=None, rgb_w=None):
        if _fused_ok(inputs, targets):
            c_l, f_l, b_l, s_l, mse_f = _NerfWLossFn.apply(inputs["rgb_coarse"], inputs["rgb_fine"], inputs["beta"],
                                                           inputs["transient_sigmas"], targets, self.coef, self.lambda_u)
            self.last_mse_fine = mse_f      # mean((rgb_fine - targets)^2) of this call: the PSNR read-out of the training loop
            return {"c_l": c_l, "f_l": f_l, "b_l": b_l, "s_l": s_l}
        self.last_mse_fine = None
        ret = {"c_l": 0.5 * ((inputs["rgb_coarse"] - targets) ** 2).mean()}
        if "rgb_fine" in inputs:
            if "beta" not in inputs:
                ret["f_l"] = 0.5 * ((inputs["rgb_fine"] - targets) ** 2).mean()
            else:
                ret["f_l"] = ((inputs["rgb_fine"] - targets) ** 2 / (2 * inputs["beta"].unsqueeze(1) ** 2)).mean()
                ret["b_l"] = 3 + torch.log(inputs["beta"]).mean()
                ret["s_l"] = self.lambda_u * inputs["transient_sigmas"].mean()
        return {k: self.coef * v for k, v in ret.items()}


loss_dict = {"color": ColorLoss, "nerfw": NerfWLoss}
