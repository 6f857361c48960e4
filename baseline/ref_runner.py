"""Drive the UNMODIFIED reference (baseline/_ref via ref_shims) for bench.py --impl reference and the drop-in tests.

`reference_render_kwargs` builds the dict create_nerf returns (models/nerfw.py:425-434,476-500) by hand: create_nerf
itself hard-codes torch.device("cuda") and lists a log directory, which a CPU-timed arm cannot use."""
import torch

from . import ref_shims


def reference_render_kwargs(mods, test_time=True, netchunk=65536, N_samples=64, N_importance=128, device="cpu"):
    """mods = (coarse, fine, emb_a, emb_t) dfnet_b200 modules (their state_dicts are the reference's, key for key).
    Returns the reference's render kwargs with REFERENCE NeRFW modules holding the same weights."""
    ref_shims.activate()
    from models import nerfw as RN
    c, f, ea, et = mods
    embed_fn, input_ch, _ = RN.get_embedder(10, 0, -1)
    embeddirs_fn, input_ch_views, _ = RN.get_embedder(4, 0, -1)
    with torch.random.fork_rng(devices=[]):
        coarse = RN.NeRFW("coarse", D=c.D, W=c.W, skips=[4], in_channels_xyz=input_ch, in_channels_dir=input_ch_views)
        coarse.load_state_dict(c.state_dict())
        fine = None
        if f is not None:
            fine = RN.NeRFW("fine", D=f.D, W=f.W, skips=[4], in_channels_xyz=input_ch, in_channels_dir=input_ch_views,
                            encode_appearance=True, encode_transient=True, in_channels_a=f.in_channels_a,
                            in_channels_t=f.in_channels_t)
            fine.load_state_dict(f.state_dict())
        emb_a = torch.nn.Embedding(*ea.weight.shape)
        emb_t = torch.nn.Embedding(*et.weight.shape)
        emb_a.load_state_dict(ea.state_dict()), emb_t.load_state_dict(et.state_dict())
    mods_ref = [m.to(device) for m in (coarse, fine, emb_a, emb_t) if m is not None]
    for m in mods_ref:
        for p in m.parameters():
            p.requires_grad_(False)

    def network_query_fn(inputs, viewdirs, ts, network_fn, typ, embedding_a, embedding_t, output_transient, test_time):
        return RN.run_network_NeRFW(inputs, viewdirs, ts, network_fn, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn, typ=typ,
                                    embedding_a=embedding_a, embedding_t=embedding_t, output_transient=output_transient,
                                    netchunk=netchunk, test_time=test_time)
    return {"network_query_fn": network_query_fn, "perturb": False if test_time else 1.0,
            "N_importance": N_importance, "network_fine": fine, "N_samples": N_samples, "network_fn": coarse,
            "use_viewdirs": True, "white_bkgd": False, "raw_noise_std": 0., "embedding_a": emb_a, "embedding_t": emb_t,
            "test_time": test_time, "ndc": False, "lindisp": False}


def reference_render_rays(kwargs, rays_o, rays_d, near, far, hist, H=1, W=1, focal=1.0, chunk=32768):
    """models.rendering.render on explicit rays (CPU or GPU tensors) -> (rgb, disp, acc, extras)."""
    ref_shims.activate()
    from models import rendering as RR
    with torch.no_grad():
        return RR.render(H, W, focal, chunk=chunk, rays=(rays_o, rays_d), near=near, far=far, img_idx=hist, **kwargs)
