"""CPU-side checks of round 2's host logic and test infrastructure (no GPU): bench.py's config planning and traffic stamp,
the launcher's stand-in packages, the synthetic scene, the training harness' determinism, the isolated reference import."""
import json
import os
import sys
import tempfile
from types import SimpleNamespace

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref", "nerfplusplus")


def _bench():
    sys.path.insert(0, ROOT)
    import bench
    return bench


def test_bench_plan_configs():
    """BASELINE.json:configs restated: c2 weak = 4096 rays per GPU whatever N; strong / c4 / c5 shard a fixed global batch."""
    B = _bench()
    A = lambda **k: SimpleNamespace(config=k.get("config", "c2"), scaling=k.get("scaling", "weak"), rays_per_gpu=k.get("rays", 0))
    assert B.plan(A(), 8)[:4] == (4096, 32768, ("mse",), "weak")
    assert B.plan(A(scaling="strong"), 8)[:4] == (512, 4096, ("mse",), "strong")
    assert B.plan(A(config="c4"), 4)[:4] == (2048, 8192, ("l1",), "strong")
    assert B.plan(A(config="c5"), 8)[:4] == (2048, 16384, ("mse", "l1", "kl"), "strong")
    assert B.plan(A(config="c5", rays=100), 8)[0] == 100
    with pytest.raises(SystemExit):
        B.plan(A(config="c4"), 3)               # ddp_train_nerf.py:137-139: the pixel count must divide by the world size


def test_bench_traffic_stamp():
    """roofline.traffic comes from the committed ncu summary and carries the file's name and hash."""
    B = _bench()
    t = B.ncu_traffic("field_tc_kernel")
    assert t is not None and t["bytes_per_launch"] > 1e6 and t["source"].endswith("traffic.json") and len(t["sha16"]) == 16
    assert B.ncu_traffic("no_such_kernel") is None


def test_compat_configargparse_reads_trainer_config(tmp_path):
    """The stand-in for configargparse (used only where the real package is missing): `key = value` files, comments,
    booleans for store_true options, `None` skipped, command line wins."""
    sys.path.insert(0, os.path.join(ROOT, "outdoor-nerf-depth_b200", "compat"))
    import importlib
    ca = importlib.import_module("configargparse")
    if not ca.__file__.startswith(os.path.join(ROOT, "outdoor-nerf-depth_b200", "compat")):
        pytest.skip("a real configargparse is installed")
    cfg = tmp_path / "c.txt"
    cfg.write_text("### INPUT\ndatadir = /data\nscene = s1\nckpt_path = None\nuse_viewdirs = True\nuse_depth = False\n"
                   "cascade_samples = 64,128\nN_iters = 500001   # comment\n")
    p = ca.ArgumentParser()
    p.add_argument("--config", is_config_file=True)
    p.add_argument("--datadir", type=str, default=None)
    p.add_argument("--scene", type=str, default=None)
    p.add_argument("--ckpt_path", type=str, default=None)
    p.add_argument("--use_viewdirs", action="store_true")
    p.add_argument("--use_depth", action="store_true")
    p.add_argument("--cascade_samples", type=str, default="64,64")
    p.add_argument("--N_iters", type=int, default=1)
    a = p.parse_args(["--config", str(cfg), "--scene", "s2", "--use_depth"])
    assert (a.datadir, a.scene, a.ckpt_path, a.use_viewdirs, a.use_depth, a.cascade_samples, a.N_iters) == \
        ("/data", "s2", None, True, True, "64,128", 500001)
    assert "config file" in p.format_values()


def test_compat_imageio_roundtrip(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "outdoor-nerf-depth_b200", "compat"))
    import importlib
    io = importlib.import_module("imageio")
    rgb = (np.arange(4 * 5 * 3).reshape(4, 5, 3) * 4).astype(np.uint8)
    d16 = (np.arange(20).reshape(4, 5) * 3000).astype(np.uint16)
    io.imwrite(str(tmp_path / "a.png"), rgb)
    io.imwrite(str(tmp_path / "d.png"), d16)
    assert np.array_equal(io.imread(str(tmp_path / "a.png")), rgb)          # RGB order survives the OpenCV round trip
    got = io.imread(str(tmp_path / "d.png"))
    assert got.dtype == np.uint16 and np.array_equal(got, d16)


def test_synth_scene_is_loadable_and_bounded():
    """oracle/synth_scene.py writes the reference's on-disk layout; every camera sits inside the unit sphere
    (intersect_sphere's precondition), images have >= 1024 pixels (the trainer forces N_rand = 1024), sky pixels carry no depth."""
    import nerfpp_oracle as O
    import synth_scene
    import train_harness as TH
    base = tempfile.mkdtemp()
    synth_scene.write_scene(base)
    tr = TH.load_views(base, "synth_learnable", "train")
    te = TH.load_views(base, "synth_learnable", "test")
    assert len(tr) == synth_scene.N_TRAIN and len(te) == synth_scene.N_TEST and tr[0]["H"] * tr[0]["W"] >= 1024
    for v in tr + te:
        O.intersect_sphere(torch.from_numpy(v["ray_o"]), torch.from_numpy(v["ray_d"]))      # raises if unbounded
        assert 0.05 < float((v["depth_sup"] > 0).mean()) < 0.6 and float(v["depth_gt"].max()) < 1.2
        assert v["rgb"].shape == (v["H"] * v["W"], 3) and 0.0 <= v["rgb"].min() and v["rgb"].max() <= 1.0


def test_train_harness_draws_are_deterministic_and_oracle_steps():
    """Both arms of the convergence test consume train_harness.step_draws: same (seed, step) -> same image, pixels and
    uniform draws; different steps differ.  Two tiny oracle steps run on the CPU and change the weights."""
    import nerfpp_oracle as O
    import synth_scene
    import train_harness as TH
    a, b = TH.step_draws(3, 7, 12, 3072, 64), TH.step_draws(3, 7, 12, 3072, 64)
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and all(torch.equal(a[2][k], b[2][k]) for k in a[2])
    c = TH.step_draws(3, 8, 12, 3072, 64)
    assert not np.array_equal(a[1], c[1]) and len(set(a[1].tolist())) == 64            # without replacement
    base = tempfile.mkdtemp()
    synth_scene.write_scene(base)
    views = TH.load_views(base, "synth_learnable", "train")
    init = [O.densify(p, 5.0) for p in O.make_params_levels(2)]
    w, hist = TH.train_oracle(init, views, 2, "cpu", n_rand=16)
    assert len(hist) == 2 and all(np.isfinite(x) for r in hist for x in r)
    k = "nerf_net.fg_net.base_layers.0.0.weight"
    assert not torch.equal(w[0][k], init[0][k]) and not torch.equal(w[1][k], init[1][k])


def test_reference_import_is_isolated():
    """oracle/_refload.load_reference imports the reference's ddp_model / depth_loss while the product's same-named
    modules stay bound in sys.modules (bench.py's product arm and its cpu_baseline leg share a process)."""
    if not os.path.isdir(REF) and not os.path.isdir("/root/reference/nerf-methods/nerfplusplus"):
        pytest.skip("no reference files here")
    import ddp_model as ours
    import ref_harness as RH
    nets = RH.build_nets(2, sigma_bias=0.0)
    import ddp_model as again
    assert again is ours and ours.__file__.endswith(os.path.join("outdoor-nerf-depth_b200", "ddp_model.py"))
    assert type(nets[0]).__module__ == "ddp_model" and type(nets[0]) is not ours.NerfNetWithAutoExpo
    import nerfpp_oracle as O
    rays = O.synthetic_rays(8, seed=2)
    out = RH.reference_step(nets, rays, cascade=(16, 16))
    assert len(out) == 2 and all(torch.isfinite(l) for _, l in out) and out[1][0]["rgb"].shape == (8, 3)
    # ... and the reference's own forward equals the oracle port on the same inputs (the port is pinned by goldens; this ties
    # the timed "reference" arm to it)
    p = O.make_params_levels(1)[0]
    far = O.intersect_sphere(rays["ray_o"], rays["ray_d"])
    fg = O.coarse_fg_depths(rays["min_depth"], far, 16)
    bg = O.coarse_bg_depths(8, 16).contiguous()
    with torch.no_grad():
        a = nets[0](rays["ray_o"], rays["ray_d"], far, fg, bg)
        b = O.nerfpp_forward(p, rays["ray_o"], rays["ray_d"], far, fg, bg)
    assert torch.allclose(a["rgb"], b["rgb"], atol=2e-6) and torch.allclose(a["depth"], b["depth"], rtol=2e-6)
