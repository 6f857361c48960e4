"""Host-side mirror of the reference's `models/nerfw.py` for the render hot path.

`NeRFW` keeps the reference's parameter names, registration order and RNG behaviour
(reference models/nerfw.py:220-295) so that reference checkpoints load unchanged and a
freshly constructed module holds bit-identical weights.  Its arithmetic runs in the
sm_100a kernels behind the C ABI (include/dfnet_b200.h); there is no eager fallback.
"""
import os

import torch
import torch.nn as nn

img2mse = lambda x, y: torch.mean((x - y) ** 2)  # noqa: E731  (reference models/nerfw.py:11; plain tensors, host side)
mse2psnr = lambda x: -10. * torch.log(x) / torch.log(torch.tensor([10.], device=x.device))  # noqa: E731  (:12)
to8b = lambda x: (255 * __import__("numpy").clip(x, 0, 1)).astype("uint8")  # noqa: E731  (:13)


class NeRFW(nn.Module):
    """NeRF-W / NeRF-Hist MLP container (reference models/nerfw.py:220-295).

    state_dict keys: xyz_encoding_{1..D}.0.{weight,bias}, xyz_encoding_final,
    dir_encoding.0, static_sigma.0, static_rgb.0 and, for the fine network,
    transient_encoding.{0,2,4,6}, transient_{sigma,rgb,beta}.0.
    """

    def __init__(self, typ, D=8, W=256, skips=(4,), in_channels_xyz=63, in_channels_dir=27,
                 encode_appearance=False, in_channels_a=48, encode_transient=False,
                 in_channels_t=16, beta_min=0.1, out_ch_size=3):
        super().__init__()
        # The reference reseeds the global RNG inside __init__ (nerfw.py:245); coarse and
        # fine trunks therefore share their initial weights.  Kept for parity.
        torch.manual_seed(0)
        if out_ch_size != 3:
            raise NotImplementedError("feature-output NeRFW (out_ch_size != 3) is outside the hot path")
        self.typ = typ
        self.D, self.W, self.skips = D, W, list(skips)
        self.in_channels_xyz, self.in_channels_dir = in_channels_xyz, in_channels_dir
        self.encode_appearance = False if typ == "coarse" else encode_appearance
        self.in_channels_a = in_channels_a if encode_appearance else 0
        self.encode_transient = False if typ == "coarse" else encode_transient
        self.in_channels_t = in_channels_t
        self.beta_min = beta_min

        for i in range(D):
            fan_in = in_channels_xyz if i == 0 else (W + in_channels_xyz if i in self.skips else W)
            setattr(self, f"xyz_encoding_{i + 1}", nn.Sequential(nn.Linear(fan_in, W), nn.ReLU(True)))
        self.xyz_encoding_final = nn.Linear(W, W)
        self.dir_encoding = nn.Sequential(
            nn.Linear(W + in_channels_dir + self.in_channels_a, W // 2), nn.ReLU(True))
        self.static_sigma = nn.Sequential(nn.Linear(W, 1), nn.Softplus())
        self.static_rgb = nn.Sequential(nn.Linear(W // 2, 3), nn.Sigmoid())
        if self.encode_transient:
            self.transient_encoding = nn.Sequential(
                nn.Linear(W + in_channels_t, W // 2), nn.ReLU(True),
                nn.Linear(W // 2, W // 2), nn.ReLU(True),
                nn.Linear(W // 2, W // 2), nn.ReLU(True),
                nn.Linear(W // 2, W // 2), nn.ReLU(True))
            self.transient_sigma = nn.Sequential(nn.Linear(W // 2, 1), nn.Softplus())
            self.transient_rgb = nn.Sequential(nn.Linear(W // 2, 3), nn.Sigmoid())
            self.transient_beta = nn.Sequential(nn.Linear(W // 2, 1), nn.Softplus())

    def forward(self, x, sigma_only=False, output_transient=True):
        """Embedded points -> raw outputs (reference nerfw.py:297-354) on the CUDA path."""
        from . import ops
        return ops.nerfw_forward(self, x, sigma_only=sigma_only, output_transient=output_transient)


def synthetic_init_(model, gain=1.6, sigma_bias=-1.0):
    """Benchmark/test initialisation from SURVEY.md §8(d): default init, every Linear
    weight scaled by `gain`, sigma-head bias set to `sigma_bias` (gives the random field
    structure and acc ~ 1)."""
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, nn.Linear):
                m.weight.mul_(gain)
        model.static_sigma[0].bias.fill_(sigma_bias)
    return model


def make_synthetic_nerf(D=8, W=256, in_channels_a=50, in_channels_t=20, n_vocab=1000, seed=0,
                        gain=1.6, sigma_bias=-1.0, fine=True):
    """Coarse + fine NeRFW and the two histogram embeddings, seeded like
    run_nerf.py:24-27 then create_nerf (reference nerfw.py:386-419)."""
    torch.manual_seed(seed)
    emb_a = nn.Embedding(n_vocab, 5)
    emb_t = nn.Embedding(n_vocab, 2)
    coarse = synthetic_init_(NeRFW("coarse", D=D, W=W), gain, sigma_bias)
    net_fine = None
    if fine:
        net_fine = synthetic_init_(
            NeRFW("fine", D=D, W=W, encode_appearance=True, encode_transient=True,
                  in_channels_a=in_channels_a, in_channels_t=in_channels_t), gain, sigma_bias)
    return coarse, net_fine, emb_a, emb_t


def _arg(args, name, default):
    return getattr(args, name, default)


def create_nerf(args):
    """Instantiate the NeRF-Hist networks, optimizer and render kwargs and reload the newest checkpoint (reference
    models/nerfw.py:356-502) -> (render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer).

    Same `args` attribute names as models/options.py.  Checkpoint format (run_nerf.py:150-167): a torch-saved dict
    with global_step, network_fn_state_dict, network_fine_state_dict, embedding_a_state_dict, embedding_t_state_dict
    (+ optimizer_state_dict, which the reference does not reload either); the newest `*tar*` file of
    basedir/expname is taken unless args.ft_path names one.  The kwargs carry the modules; the arithmetic of
    `network_query_fn` (positional encoding + MLP, nerfw.py:15-95) lives inside the fused kernels, so the entry is kept
    only for signature compatibility and render() ignores it."""
    if not _arg(args, "NeRFH", True):
        raise NotImplementedError("only the NeRF-Hist / NeRF-W networks (--NeRFH) are on the B200 path")
    if _arg(args, "multi_gpu", False):
        raise NotImplementedError("--multi_gpu (nn.DataParallel) is replaced by one process per GPU (bench.py --gpus N)")
    if _arg(args, "reduce_embedding", -1) not in (-1, None) or _arg(args, "i_embed", 0) != 0:
        raise NotImplementedError("only the paper-default positional encoding (reduce_embedding=-1, i_embed=0) is on the B200 path")
    if not _arg(args, "use_viewdirs", True):
        raise NotImplementedError("NeRF-Hist always renders with use_viewdirs=True")
    if not _arg(args, "encode_hist", True):
        raise NotImplementedError("NeRF-Hist needs --encode_hist (the reference leaves embedding_a undefined without it, nerfw.py:384-390)")
    multires, multires_views = _arg(args, "multires", 10), _arg(args, "multires_views", 4)
    input_ch, input_ch_views = 3 + 6 * multires, 3 + 6 * multires_views
    if not torch.cuda.is_available():
        raise RuntimeError("create_nerf needs a CUDA device (the reference hard-codes torch.device('cuda'), nerfw.py:380)")
    device = torch.device("cuda", torch.cuda.current_device())
    embedding_a = nn.Embedding(args.N_vocab, 5).to(device)
    embedding_t = nn.Embedding(args.N_vocab, 2).to(device)
    model = NeRFW("coarse", D=args.netdepth, W=args.netwidth, skips=[4], in_channels_xyz=input_ch,
                  in_channels_dir=input_ch_views).to(device)
    grad_vars = list(model.parameters())
    model_fine = None
    if args.N_importance > 0:
        model_fine = NeRFW("fine", D=args.netdepth, W=args.netwidth, skips=[4], in_channels_xyz=input_ch,
                           in_channels_dir=input_ch_views, encode_appearance=True, encode_transient=True,
                           in_channels_a=args.in_channels_a, in_channels_t=args.in_channels_t).to(device)
        grad_vars += list(model_fine.parameters())
        grad_vars += list(embedding_a.parameters())
        grad_vars += list(embedding_t.parameters())
    if _arg(args, "no_grad_update", False):
        grad_vars, optimizer = None, None
    else:
        optimizer = torch.optim.Adam(params=grad_vars, lr=args.lrate, betas=(0.9, 0.999))

    start = 0
    ft_path = _arg(args, "ft_path", None)
    if ft_path is not None and ft_path != "None":
        ckpts = [ft_path]
    else:
        d = os.path.join(args.basedir, args.expname)
        ckpts = [os.path.join(d, f) for f in sorted(os.listdir(d)) if "tar" in f]
    print("Found ckpts", ckpts)
    if len(ckpts) > 0 and not _arg(args, "no_reload", False):
        print("Reloading from", ckpts[-1])
        ckpt = torch.load(ckpts[-1], map_location=device, weights_only=False)
        start = ckpt["global_step"]
        model.load_state_dict(ckpt["network_fn_state_dict"])
        if model_fine is not None:
            model_fine.load_state_dict(ckpt["network_fine_state_dict"])
            embedding_a.load_state_dict(ckpt["embedding_a_state_dict"])
            embedding_t.load_state_dict(ckpt["embedding_t_state_dict"])

    def network_query_fn(*a, **k):  # noqa: ARG001
        raise RuntimeError("network_query_fn is fused into the render kernels; call dfnet_b200.rendering.render "
                           "(or NeRFW.forward on embedded points)")

    render_kwargs_train = {
        "network_query_fn": network_query_fn, "perturb": args.perturb, "N_importance": args.N_importance,
        "network_fine": model_fine, "N_samples": args.N_samples, "network_fn": model,
        "use_viewdirs": _arg(args, "use_viewdirs", True), "white_bkgd": _arg(args, "white_bkgd", False),
        "raw_noise_std": _arg(args, "raw_noise_std", 0.), "embedding_a": embedding_a, "embedding_t": embedding_t,
        "test_time": False}
    if _arg(args, "dataset_type", "7Scenes") != "llff" or _arg(args, "no_ndc", False):
        render_kwargs_train["ndc"] = False
        render_kwargs_train["lindisp"] = _arg(args, "lindisp", False)
    render_kwargs_test = dict(render_kwargs_train)
    render_kwargs_test["perturb"] = False
    render_kwargs_test["raw_noise_std"] = 0.
    render_kwargs_test["test_time"] = True
    return render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer


def save_checkpoint(path, global_step, render_kwargs, optimizer=None):
    """The checkpoint dict run_nerf.py:150-167 writes (same keys; create_nerf above and the reference's create_nerf
    both load it)."""
    d = {"global_step": global_step, "network_fn_state_dict": render_kwargs["network_fn"].state_dict()}
    if render_kwargs.get("network_fine") is not None:
        d["network_fine_state_dict"] = render_kwargs["network_fine"].state_dict()
        d["embedding_a_state_dict"] = render_kwargs["embedding_a"].state_dict()
        d["embedding_t_state_dict"] = render_kwargs["embedding_t"].state_dict()
    if optimizer is not None:
        d["optimizer_state_dict"] = optimizer.state_dict()
    torch.save(d, path)
    return path
