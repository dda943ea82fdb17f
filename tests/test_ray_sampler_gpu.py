"""N1: device ray generation / batch gather vs rays recorded from the UNMODIFIED reference sampler
(oracle/gen_golden_rays.py -> tests/golden/rays_tat_truck.npz) and vs the oracle restatement."""
import os

import numpy as np
import pytest
import torch

import nerfpp_oracle as O

pytestmark = pytest.mark.gpu


def test_rays_match_reference_golden(golden_dir):
    from nerfpp_b200.ray_sampler import DeviceRaySampler
    g = dict(np.load(os.path.join(golden_dir, "rays_tat_truck.npz")))
    for ci in range(2):
        H, W = (int(x) for x in g["hw%d" % ci])
        s = DeviceRaySampler(H, W, g["K%d" % ci], g["c2w%d" % ci])
        ret = s.random_sample(2048, select_inds=g["ids%d" % ci])
        # np.dot runs the two 3x3 products through BLAS sgemm (summation order / FMA use unspecified): 1e-6 relative
        np.testing.assert_allclose(ret["ray_d"].cpu().numpy(), g["ray_d%d" % ci], rtol=2e-6, atol=1e-7)
        assert np.array_equal(ret["ray_o"].cpu().numpy(), g["ray_o%d" % ci])
        np.testing.assert_allclose(ret["depth"].cpu().numpy(), g["depth%d" % ci], rtol=1e-6)
        assert ret["rgb"] is None and torch.all(ret["min_depth"] == 1e-4) and "depth_sup" not in ret


def test_sampler_gathers_and_get_all_match_oracle(golden_dir):
    from nerfpp_b200.ray_sampler import DeviceRaySampler
    g = dict(np.load(os.path.join(golden_dir, "rays_tat_truck.npz")))
    H, W = 24, 40
    K = g["K0"].copy(); K[0, 2], K[1, 2] = W / 2, H / 2
    rng = np.random.default_rng(0)
    img = rng.random((H, W, 3), dtype=np.float32)
    dsup = rng.random((H, W), dtype=np.float32); dsup[::3] = 0
    mind = rng.random((H, W), dtype=np.float32) * 0.1
    s = DeviceRaySampler(H, W, K, g["c2w0"], img=img, depth_sup=dsup, min_depth=mind, img_path="a/b/c.png", depth_scale=0.05)
    ro, rd, dp = O.get_rays_single_image(H, W, K, g["c2w0"])
    allr = s.get_all()
    np.testing.assert_allclose(allr["ray_d"].cpu().numpy(), rd, rtol=2e-6, atol=1e-7)
    assert np.array_equal(allr["rgb"].cpu().numpy(), img.reshape(-1, 3)) and np.array_equal(allr["depth_sup"].cpu().numpy(), dsup.reshape(-1))
    np.random.seed(3)
    want_ids = np.random.choice(H * W, size=(100,), replace=False)
    np.random.seed(3)
    ret = s.random_sample(100)                      # same numpy draw as the reference (nerf_sample_ray_split.py:178)
    ref = O.sample_ray_batch(H, W, K, g["c2w0"], want_ids, img, dsup, mind)
    for k in ("rgb", "depth_sup", "min_depth"):
        assert np.array_equal(ret[k].cpu().numpy(), ref[k]), k
    np.testing.assert_allclose(ret["ray_d"].cpu().numpy(), ref["ray_d"], rtol=2e-6, atol=1e-7)
    assert ret["img_name"] == "a/b/c.png" and s.get_depth_scale() == 0.05
    assert list(ret.keys()) == ["ray_o", "ray_d", "depth", "rgb", "mask", "min_depth", "depth_sup", "img_name"]
    cc = s.random_sample(16, center_crop=True)
    assert cc["ray_d"].shape == (16, 3)
    dv = s.random_sample(50, device_rng=True)
    assert dv["ray_d"].shape == (50, 3)


def test_image_metrics_match_oracle():
    """N3: PSNR / RMSE / AbsRel ... of the reference's test loop (ddp_train_nerf.py:556-600)."""
    from nerfpp_b200 import ops
    rng = np.random.default_rng(0)
    H, W, scale = 64, 96, 0.05
    im, gt = rng.random((H, W, 3), dtype=np.float32), rng.random((H, W, 3), dtype=np.float32)
    dgt = (rng.random((H, W), dtype=np.float32) * 100 * scale).astype(np.float32)     # some beyond the 80 m cap
    dgt[::4] = 0                                                                       # sparse LiDAR
    dpred = (dgt + rng.normal(0, 0.2, (H, W)).astype(np.float32) * scale).astype(np.float32)
    ref = O.image_metrics(im, gt, dpred, dgt, scale)
    got = ops.image_metrics(torch.from_numpy(im).cuda(), torch.from_numpy(gt).cuda(), torch.from_numpy(dpred).cuda(),
                            torch.from_numpy(dgt).cuda(), scale)
    for k, v in ref.items():
        assert abs(got[k] - v) <= 2e-5 * max(abs(v), 1e-6), (k, got[k], v)
