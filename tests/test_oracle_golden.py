"""Pins oracle/nerfpp_oracle.py against outputs of the UNMODIFIED reference (tests/golden/*.npz,
made by oracle/gen_golden.py).  CPU only."""
import hashlib
import os

import numpy as np
import pytest
import torch

import nerfpp_oracle as O

CASES = ["c1_coarse_det", "c2_train_dense", "c2_train_init", "c2_det_dense"]
FLOAT_RTOL = 2e-6   # oracle vs reference on the same CPU: same ATen ops => normally bit-exact


def load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, "nerfpp_%s.npz" % name))
    return {k: z[k] for k in z.files}


def T(a):
    return torch.from_numpy(np.asarray(a))


def digest(params):
    h = hashlib.sha256()
    for k in sorted(params):
        h.update(k.encode())
        h.update(params[k].numpy().astype(np.float32).tobytes())
    return h.hexdigest()


def close(a, b, rtol=FLOAT_RTOL, atol=0.0):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    scale = np.abs(b).max() if b.size else 1.0
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol + rtol * scale * 1e-2)


def test_param_init_matches_reference_rng_stream(golden_dir):
    g = load(golden_dir, "c2_train_dense")
    levels = O.make_params_levels(2)
    assert [digest(p) for p in levels] == list(g["meta_param_digest"])
    assert sum(v.numel() for v in levels[0].values()) == 1202440


@pytest.mark.parametrize("name", CASES)
def test_cascade_matches_reference(golden_dir, name):
    g = load(golden_dir, name)
    cascade = tuple(int(x) for x in g["meta_cascade"])
    train = bool(g["meta_train"])
    levels = O.make_params_levels(len(cascade))
    if float(g["meta_sigma_bias"]):
        levels = [O.densify(p, float(g["meta_sigma_bias"])) for p in levels]
    rand = None
    if train:
        rand = {k[5:]: T(v) for k, v in g.items() if k.startswith("rand_")}
    with torch.no_grad():
        out, fg_far = O.cascade_forward(levels, T(g["ray_o"]), T(g["ray_d"]), T(g["min_depth"]), cascade, rand)
    assert np.array_equal(fg_far.numpy(), g["fg_far"])
    for m, (ret, fg_z, bg_z) in enumerate(out):
        assert np.array_equal(fg_z.numpy(), g["fg_z_%d" % m]), "fg depths level %d" % m
        assert np.array_equal(bg_z.numpy(), g["bg_z_%d" % m]), "bg depths level %d" % m
        assert list(ret.keys()) == ["rgb", "fg_weights", "bg_weights", "fg_dists", "fg_rgb", "fg_depth",
                                    "bg_rgb", "bg_depth", "bg_lambda", "depth"]
        for k, v in ret.items():
            close(v.numpy(), g["ret%d_%s" % (m, k)])
        sig = float(g["meta_depth_sigma"]) * float(g["meta_depth_scale"])
        close(O.img2mse(ret["rgb"], T(g["rgb_gt"])), g["loss%d_rgb" % m])
        close(O.depth_mse(T(g["depth_sup"]), ret["depth"]), g["loss%d_mse" % m])
        close(O.depth_l1(T(g["depth_sup"]), ret["depth"]), g["loss%d_l1" % m])
        close(O.depth_kl(ret["fg_weights"], T(g["depth_sup"]), fg_z, ret["fg_dists"], sig, fg_far),
              g["loss%d_kl" % m])


@pytest.mark.parametrize("name", ["c2_train_dense", "c2_train_init", "c2_det_dense"])
def test_inverse_cdf_indices_bit_exact(golden_dir, name):
    """The reference's own gather indices (ddp_train_nerf.py:111-121) given its own weights."""
    g = load(golden_dir, name)
    train = bool(g["meta_train"])
    for side in ("fg", "bg"):
        w = T(g["ret0_%s_weights" % side])[..., 1:-1]
        z = T(g["%s_z_0" % side])
        mids = 0.5 * (z[..., 1:] + z[..., :-1])
        u = T(g["rand_u_%s_1" % side]) if train else O.det_u(z.shape[0], 128)
        samples, cdf, above = O.sample_pdf(mids, w, u, return_aux=True)
        assert np.array_equal(cdf.numpy(), g["%s_cdf_1" % side])
        assert np.array_equal(above.numpy().astype(np.int32), g["%s_inds_1" % side])
        assert np.array_equal(samples.numpy(), g["%s_new_1" % side])
        # contract used by the CUDA kernel: count == searchsorted(right=True) on cdf[:M]
        M = cdf.shape[-1] - 1
        ss = torch.searchsorted(cdf[..., :M].contiguous(), u.contiguous(), right=True)
        assert torch.equal(ss, above)
        assert int(above.min()) >= 1 and int(above.max()) <= M


def test_gradients_match_reference(golden_dir):
    g = load(golden_dir, "c2_train_dense")
    cascade = (64, 128)
    levels = [O.densify(p, 5.0) for p in O.make_params_levels(2)]
    sig = float(g["meta_depth_sigma"]) * float(g["meta_depth_scale"])
    for m in range(2):
        params = {k: v.clone().requires_grad_(True) for k, v in levels[m].items()}
        fg_z, bg_z = T(g["fg_z_%d" % m]), T(g["bg_z_%d" % m])
        for lt in ("mse", "l1", "kl"):
            ret = O.nerfpp_forward(params, T(g["ray_o"]), T(g["ray_d"]), T(g["fg_far"]), fg_z, bg_z)
            loss, _, _ = O.level_loss(ret, T(g["rgb_gt"]), T(g["depth_sup"]), fg_z, T(g["fg_far"]), True, lt, 0.1, sig)
            grads = torch.autograd.grad(loss, list(params.values()))
            for (pname, _), gr in zip(params.items(), grads):
                short = pname.replace("nerf_net.", "").replace("_layers", "").replace(".weight", ".w").replace(".bias", ".b")
                ref_norm = float(g["grad%d_%s_norm/%s" % (m, lt, short)])
                assert abs(float(gr.norm()) - ref_norm) <= 1e-5 * max(ref_norm, 1e-12), (m, lt, short)
                close(gr.reshape(-1)[:16].numpy(), g["grad%d_%s_head/%s" % (m, lt, short)], rtol=1e-4)


def test_intersect_sphere_raises_outside_unit_sphere():
    o = torch.tensor([[0.0, 0.0, 2.0]])
    d = torch.tensor([[1.0, 0.0, 0.0]])
    with pytest.raises(O.UnboundedCameraError):
        O.intersect_sphere(o, d)


def test_depth_losses_edge_cases():
    gt = torch.zeros(7)
    pred = torch.rand(7)
    assert torch.isnan(O.depth_mse(gt, pred)) and torch.isnan(O.depth_l1(gt, pred))   # empty mean => NaN
    w = torch.rand(7, 5)
    assert float(O.depth_kl(w, gt, torch.rand(7, 5), torch.rand(7, 5), 0.01, torch.ones(7))) == 0.0


def test_ray_generation_matches_reference(golden_dir):
    """N1: oracle get_rays_single_image vs rays recorded from the unmodified reference sampler (gen_golden_rays.py)."""
    g = dict(np.load(os.path.join(golden_dir, "rays_tat_truck.npz")))
    for ci in range(2):
        H, W = (int(x) for x in g["hw%d" % ci])
        ro, rd, dp = O.get_rays_single_image(H, W, g["K%d" % ci], g["c2w%d" % ci])
        ids = g["ids%d" % ci]
        assert np.array_equal(rd[ids].astype(np.float32), g["ray_d%d" % ci]) and np.array_equal(ro[ids].astype(np.float32), g["ray_o%d" % ci])
        assert np.array_equal(dp[ids].astype(np.float32), g["depth%d" % ci])
