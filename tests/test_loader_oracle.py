"""N2: the numpy restatement of the reference's on-disk loader (oracle/data_loader_oracle.py) against
arrays recorded from the UNMODIFIED reference loader (oracle/gen_golden_loader.py -> tests/golden/loader_scene.npz) on
the deterministic synthetic scene of oracle/loader_scene.py; plus the host half of the product loader (file discovery,
order, skip, scale / max_depth parsing), which needs no GPU."""
import os

import numpy as np
import pytest

import data_loader_oracle as DO
import loader_scene

CASES = (("a", "train", 1, "gt"), ("b", "train", 2, "mono"), ("c", "test", 1, "mono"))
KEYS = ("ray_o", "ray_d", "depth", "rgb", "min_depth", "depth_gt", "depth_sup")


@pytest.fixture(scope="module")
def scene(tmp_path_factory):
    base = str(tmp_path_factory.mktemp("scene"))
    loader_scene.write_scene(base, "synth", seed=0)
    return base


@pytest.fixture(scope="module")
def golden(golden_dir):
    return dict(np.load(os.path.join(golden_dir, "loader_scene.npz")))


def test_oracle_loader_matches_reference_golden(scene, golden):
    for tag, split, skip, typ in CASES:
        cams = DO.load_data_split(scene, "synth", split, skip=skip, depth_sup_type=typ)
        assert len(cams) == int(golden["%s_n" % tag])
        for i, c in enumerate(cams):
            assert os.path.basename(c["img_path"]) == str(golden["%s_%d_name" % (tag, i)])
            assert c["depth_scale"] == float(golden["%s_%d_scale" % (tag, i)])
            for k in KEYS:
                want = golden["%s_%d_%s" % (tag, i, k)]
                if k in ("ray_d", "depth"):          # np.dot -> BLAS sgemm in both; identical here, tolerance for other BLAS builds
                    np.testing.assert_allclose(c[k], want, rtol=2e-6, atol=1e-7, err_msg="%s %d %s" % (tag, i, k))
                else:
                    assert np.array_equal(c[k], want), (tag, i, k)


def test_oracle_random_sample_matches_reference_golden(scene, golden):
    cams = DO.load_data_split(scene, "synth", "train", skip=1, depth_sup_type="gt")
    c = cams[1]
    np.random.seed(3)
    ids = np.random.choice(c["H"] * c["W"], size=(64,), replace=False)          # nerf_sample_ray_split.py:178
    for k in KEYS:
        want = golden["a_rs_%s" % k]
        if k in ("ray_d", "depth"):
            np.testing.assert_allclose(c[k][ids], want, rtol=2e-6, atol=1e-7)
        else:
            assert np.array_equal(c[k][ids], want), k


def test_list_split_host_logic(scene):
    import data_loader_split as DL
    assert DL.__file__.startswith(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "outdoor-nerf-depth_b200"))
    ls = DL.list_split(scene + "/", "synth", "train", skip=2, depth_sup_type="mono")
    assert ls["cam_cnt"] == 3 and (ls["H"], ls["W"]) == (loader_scene.H, loader_scene.W)
    assert [os.path.basename(p) for p in ls["rgb"]] == ["000000.png", "000006.png", "000012.png"]
    assert [os.path.basename(p) for p in ls["pose"]] == ["000000.txt", "000006.txt", "000012.txt"]
    assert all("/depth_mono/" in p for p in ls["depth_sup"]) and all("/depth/" in p for p in ls["depth_gt"])
    assert ls["depth_scale"] == loader_scene.SCALE and ls["max_depth"] == loader_scene.MAX_DEPTH
    assert ls["mask"] == [None] * 3
    assert DL.list_split(scene, "synth", "train", try_load_min_depth=False)["min_depth"] == [None] * 5
    files = DL.load_data_split(scene, "synth", "test", only_img_files=True)
    assert [os.path.basename(p) for p in files] == ["000001.png", "000004.png"]
    K = DL.parse_txt(ls["intrinsics"][0])
    assert K.dtype == np.float32 and K.shape == (4, 4) and K[3, 3] == 1


def test_split_without_depth_dir(tmp_path):
    """The reference raises IndexError here (data_loader_split.py:83-89 assigns to ``depth_files`` but reads
    ``depth_gt_files``); the drop-in returns samplers without depth instead (documented difference)."""
    import shutil
    import data_loader_split as DL
    base = str(tmp_path)
    loader_scene.write_scene(base, "synth", seed=1, with_min_depth=False)
    for split in ("train", "test"):
        shutil.rmtree(os.path.join(base, "synth", split, "depth"))
        shutil.rmtree(os.path.join(base, "synth", split, "depth_mono"))
    ls = DL.list_split(base, "synth", "train")
    assert ls["depth_gt"] == [None] * 5 and ls["depth_sup"] == [None] * 5 and ls["depth_scale"] is None and ls["max_depth"] is None


def test_decode_requires_cuda():
    from nerfpp_b200 import _lib
    from nerfpp_b200.ray_sampler import decode_pixels
    with pytest.raises(_lib.NerfppError):
        decode_pixels(np.zeros((2, 2), np.uint8), 255.0, device="cpu")
    with pytest.raises(ValueError):
        decode_pixels(np.zeros((2, 2), np.float32), 255.0)
