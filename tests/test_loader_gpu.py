"""N2: the drop-in ``data_loader_split.load_data_split`` (raw PNG pixels uploaded, decoded by
``nerfpp_decode_pixels``, rays by ``nerfpp_gen_rays``) against arrays recorded from the UNMODIFIED reference loader
(tests/golden/loader_scene.npz).  Decoded pixel data bit-exact; ray directions to BLAS-sgemm tolerance."""
import os

import numpy as np
import pytest
import torch

import loader_scene

pytestmark = pytest.mark.gpu

CASES = (("a", "train", 1, "gt"), ("b", "train", 2, "mono"), ("c", "test", 1, "mono"))
EXACT = ("ray_o", "rgb", "min_depth", "depth_gt", "depth_sup")


@pytest.fixture(scope="module")
def scene(tmp_path_factory):
    base = str(tmp_path_factory.mktemp("scene"))
    loader_scene.write_scene(base, "synth", seed=0)
    return base


def test_loader_matches_reference_golden(scene, golden_dir):
    import data_loader_split as DL
    g = dict(np.load(os.path.join(golden_dir, "loader_scene.npz")))
    for tag, split, skip, typ in CASES:
        samplers = DL.load_data_split(scene, "synth", split, skip=skip, try_load_min_depth=True, depth_sup_type=typ)
        assert len(samplers) == int(g["%s_n" % tag])
        for i, s in enumerate(samplers):
            ret = s.get_all()
            assert os.path.basename(s.img_path) == str(g["%s_%d_name" % (tag, i)])
            assert s.get_depth_scale() == float(g["%s_%d_scale" % (tag, i)])
            for k in EXACT:
                assert np.array_equal(ret[k].cpu().numpy(), g["%s_%d_%s" % (tag, i, k)]), (tag, i, k)
            np.testing.assert_allclose(ret["ray_d"].cpu().numpy(), g["%s_%d_ray_d" % (tag, i)], rtol=2e-6, atol=1e-7)
            np.testing.assert_allclose(ret["depth"].cpu().numpy(), g["%s_%d_depth" % (tag, i)], rtol=1e-6)
            # the accessors the trainer's numpy metric code reads (ddp_train_nerf.py:557-570): host arrays, like the reference
            img, dgt = s.get_img(), s.get_gt_depth_img()
            assert isinstance(img, np.ndarray) and img.shape == (s.H, s.W, 3) and dgt.shape == (s.H, s.W)
            assert np.array_equal(img.reshape(-1, 3), g["%s_%d_rgb" % (tag, i)]) and s.resolution_level == 1
            assert np.array_equal(s.get_sup_depth_img().reshape(-1), g["%s_%d_depth_sup" % (tag, i)])
        if tag == "a":
            np.random.seed(3)
            r = samplers[1].random_sample(64, center_crop=False)       # same numpy draw as the reference (:178)
            for k in EXACT:
                assert np.array_equal(r[k].cpu().numpy(), g["a_rs_%s" % k]), k
            np.testing.assert_allclose(r["ray_d"].cpu().numpy(), g["a_rs_ray_d"], rtol=2e-6, atol=1e-7)


def test_decode_pixels_exhaustive():
    """Every uint8 and every uint16 code through the three decode forms, bit-exact against numpy's evaluation order."""
    from nerfpp_b200.ray_sampler import decode_pixels
    u8 = np.arange(256, dtype=np.uint8)
    u16 = np.arange(65536, dtype=np.uint16)
    assert np.array_equal(decode_pixels(u8, 255.0).cpu().numpy(), u8.astype(np.float32) / 255.)
    md = np.float32(3.5)
    assert np.array_equal(decode_pixels(u8, 255.0, 3.5, 1e-4).cpu().numpy(), (u8.astype(np.float32) / 255. * 3.5 + 1e-4).astype(np.float32))
    for scale in (0.0125, 0.05, 1.0 / 3.0):
        want = (scale * (u16.astype(np.float32) / 256.0)).astype(np.float32)
        assert np.array_equal(decode_pixels(u16, 256.0, scale).cpu().numpy(), want), scale
    assert decode_pixels(np.zeros((0,), np.uint8), 255.0).numel() == 0


def test_loader_without_min_depth_and_with_cuda_batches(tmp_path):
    import data_loader_split as DL
    base = str(tmp_path)
    loader_scene.write_scene(base, "synth", seed=2, with_min_depth=False)
    s = DL.load_data_split(base, "synth", "train", skip=4, depth_sup_type="mono")
    assert len(s) == 2
    r = s[0].random_sample(32)
    assert torch.all(r["min_depth"] == 1e-4) and r["depth_sup"].is_cuda and r["depth_sup"].shape == (32,)
    assert list(r.keys())[:8] == ["ray_o", "ray_d", "depth", "rgb", "mask", "min_depth", "depth_gt", "depth_sup"] or "depth_gt" in r
