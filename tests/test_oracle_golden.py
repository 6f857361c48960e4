"""Pin oracle/nerf_oracle.py against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import nerf_oracle as O
from helpers import sd_checksum, synthetic_nets, rel_err

TOL = 2e-5  # fp32 reassociation noise between numpy/OpenBLAS and ATen/MKL


def test_linspace(golden):
    i = 0
    while f"linspace_{i}" in golden:
        a, b, n = golden[f"linspace_{i}_args"]
        assert np.array_equal(O.linspace_f32(a, b, int(n)), golden[f"linspace_{i}"]), (a, b, n)
        i += 1
    assert i >= 7


@pytest.mark.parametrize("n", [62, 64, 14, 192, 1000])
def test_aten_sum_order_bit_exact(golden, n):
    assert np.array_equal(O.aten_sum_lastdim(golden[f"sum_{n}_x"]), golden[f"sum_{n}"])


def test_scans_bit_exact(golden):
    assert np.array_equal(O.cumsum_f64acc(golden["cumsum_x"]), golden["cumsum"])
    assert np.array_equal(O.cumprod_f64acc(golden["cumprod_x"]), golden["cumprod"])


def test_get_rays(golden):
    H, W, f = golden["rays_hwf"]
    o, d = O.get_rays(int(H), int(W), float(f), golden["rays_c2w"])
    assert np.array_equal(o, golden["rays_o"])
    assert np.array_equal(d, golden["rays_d"])  # bit-exact: same op order


def test_embed(golden):
    x = golden["embed_x"]
    e10, e4 = O.embed(x, 10), O.embed(x, 4)
    assert e10.shape == (32, 63) and e4.shape == (32, 27)
    assert np.abs(e10 - golden["embed_L10"]).max() < 1e-6
    assert np.abs(e4 - golden["embed_L4"]).max() < 1e-6


@pytest.mark.parametrize("tag,D,W", [("s", 4, 64), ("b", 8, 256)])
def test_nerfw_forward(golden, tag, D, W):
    (c, f, _, _), nets = synthetic_nets(D, W)
    assert sd_checksum(c.state_dict()) == bytes(golden[f"mlp_{tag}_coarse_sha"]).decode()
    assert sd_checksum(f.state_dict()) == bytes(golden[f"mlp_{tag}_fine_sha"]).decode()
    x = golden[f"mlp_{tag}_x"]
    s = O.nerfw_forward(nets["coarse"], x[:, :63], D, sigma_only=True)
    assert rel_err(s, golden[f"mlp_{tag}_sigma_only"]) < TOL
    st = O.nerfw_forward(nets["coarse"], x[:, :90], D, output_transient=False, in_a=0)
    assert rel_err(st, golden[f"mlp_{tag}_coarse_static"]) < TOL
    full = O.nerfw_forward(nets["fine"], x, D, output_transient=True, in_a=50, in_t=20)
    assert full.shape == (40, 9)
    assert rel_err(full, golden[f"mlp_{tag}_fine_full"]) < TOL


@pytest.mark.parametrize("case", ["coarse_test", "coarse_train", "fine_test", "fine_train"])
def test_raw2outputs(golden, case):
    raw, z = golden["r2o_raw"], golden["r2o_z"]
    kw = dict(coarse_test=dict(raw=raw[..., 3:4], output_transient=False, test_time=True, typ="coarse"),
              coarse_train=dict(raw=raw[..., :4], output_transient=False, test_time=False, typ="coarse"),
              fine_test=dict(raw=raw, output_transient=True, test_time=True, typ="fine"),
              fine_train=dict(raw=raw, output_transient=True, test_time=False, typ="fine"))[case]
    out = O.raw2outputs_nerfw(kw.pop("raw"), z, **kw)
    n = 0
    for nm in ["rgb", "disp", "acc", "weights", "depth", "transient_sigmas", "beta"]:
        key = f"r2o_{case}_{nm}"
        if key in golden:
            assert out[nm] is not None, nm
            # alpha = 1-exp(-x) cancels for small x: 1-ulp exp differences are absolute, not relative
            assert np.allclose(out[nm], golden[key], rtol=TOL, atol=2e-6), nm
            n += 1
        else:
            assert out[nm] is None or nm == "beta", nm
    assert n >= 2


def test_sample_pdf_bit_exact_indices(golden):
    bins, w = golden["pdf_bins"], golden["pdf_w"]
    s, inds = O.sample_pdf(bins, w, 128, det=True)
    assert np.array_equal(inds, golden["pdf_det_inds"])
    assert np.array_equal(s, golden["pdf_det_samples"])  # same op order => bit-exact values too
    s, inds = O.sample_pdf(bins, w, 128, det=False, u=golden["pdf_u_rand"])
    assert np.array_equal(inds, golden["pdf_rand_inds"])
    assert np.array_equal(s, golden["pdf_rand_samples"])


def test_render_cfg1_shape(golden):
    _, nets = synthetic_nets(4, 64, fine=False)
    r = O.render(8, 8, 8.0, nets, 64, 0, 0.0, 2.5, c2w=golden["e2e_a_c2w"], hist=golden["hist"], test_time=False)
    assert rel_err(r["rgb_map"], golden["e2e_a_rgb"]) < 1e-4
    assert rel_err(r["disp_map"], golden["e2e_a_disp"]) < 1e-4
    assert rel_err(r["acc_map"], golden["e2e_a_acc"]) < 1e-4


def test_render_cfg2_shape(golden):
    _, nets = synthetic_nets(8, 256)
    r = O.render(6, 8, 7.3125, nets, 64, 128, 0.0, 2.5, c2w=golden["e2e_b_c2w"], hist=golden["hist"],
                 test_time=True, return_internals=True)
    assert rel_err(r["rgb_map"], golden["e2e_b_rgb"]) < 1e-4
    assert rel_err(r["disp_map"], golden["e2e_b_disp"]) < 1e-4
    assert rel_err(r["acc_map"], golden["e2e_b_acc"]) < 1e-4
    # op-level bit-exactness: oracle sampler on the REFERENCE's coarse weights
    _, inds = O.sample_pdf(0.5 * (r["_z_coarse"].reshape(-1, 64)[:, 1:] + r["_z_coarse"].reshape(-1, 64)[:, :-1]),
                           golden["e2e_b_w_coarse"], 128, det=True)
    assert np.array_equal(inds, golden["e2e_b_inds"])
    # end-to-end: indices may flip only where u sits within rounding noise of a cdf knot
    flips = (r["_inds"].reshape(-1, 128) != golden["e2e_b_inds"]).mean()
    assert flips < 0.01, flips
    assert np.abs(r["_z_samples"].reshape(-1, 128) - golden["e2e_b_z_samples"]).max() < 1e-4


def test_render_train_mode_extras(golden):
    _, nets = synthetic_nets(8, 64)
    rays = golden["e2e_c_rays"]
    r = O.render(4, 6, 5.0, nets, 16, 24, 0.0, 2.5, rays=(rays[0], rays[1]), hist=golden["hist"],
                 test_time=False, retraw=True)
    for k, g in [("rgb_map", "rgb"), ("disp_map", "disp"), ("acc_map", "acc"), ("rgb0", "rgb0"), ("disp0", "disp0"),
                 ("acc0", "acc0"), ("z_std", "z_std"), ("transient_sigmas", "transient_sigmas"), ("beta", "beta"),
                 ("raw", "raw")]:
        assert rel_err(r[k], golden[f"e2e_c_{g}"], floor=1e-3) < 1e-4, k


def test_render_stratified(golden):
    _, nets = synthetic_nets(8, 64)
    rays = golden["e2e_c_rays"]
    r = O.render(4, 6, 5.0, nets, 16, 24, 0.0, 2.5, rays=(rays[0], rays[1]), hist=golden["hist"],
                 test_time=False, perturb=1.0, t_rand=golden["e2e_d_t_rand"], u=golden["e2e_d_u"])
    for k, g in [("rgb_map", "rgb"), ("disp_map", "disp"), ("acc_map", "acc"), ("rgb0", "rgb0"), ("beta", "beta"),
                 ("z_std", "z_std")]:
        assert rel_err(r[k], golden[f"e2e_d_{g}"], floor=1e-3) < 1e-4, k


@pytest.mark.parametrize("W", [64, 128, 192])
def test_narrow_network_embedded_in_256_is_the_same_function(W):
    """The tcgen05 kernels run networks narrower than 256 (the reference's default netwidth is 128) embedded in
    8x256 with zero weights (csrc/mlp_tc.cu::tc_pad_params).  The claim that this is EXACTLY the narrow network's
    function is checked here on the CPU oracle: same outputs bit for bit, coarse (sigma only) and fine (all 9)."""
    from helpers import pad_nerfw_state_dict, synthetic_nets
    _, nets = synthetic_nets(8, W)
    rng = np.random.RandomState(W)
    x = rng.randn(257, 63 + 27 + 50 + 20).astype(np.float32)
    fine, coarse = nets["fine"], nets["coarse"]
    a = O.nerfw_forward(fine, x, 8, in_a=50, in_t=20)
    b = O.nerfw_forward(pad_nerfw_state_dict(fine, W), x, 8, in_a=50, in_t=20)
    assert a.shape == (257, 9) and np.array_equal(a, b)
    s = O.nerfw_forward(coarse, x[:, :63], 8, sigma_only=True)
    t = O.nerfw_forward(pad_nerfw_state_dict(coarse, W), x[:, :63], 8, sigma_only=True)
    assert np.array_equal(s, t)
