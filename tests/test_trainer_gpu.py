"""The UNMODIFIED reference trainer (baseline/_ref/nerfplusplus/ddp_train_nerf.py -- the git-ignored copy made by
oracle/install_reference.py, which travels to the GPU box) running on the CUDA kernels through
outdoor-nerf-depth_b200/launch_ddp_train_nerf.py (BASELINE.json north_star: "so ddp_train_nerf.py ...
invoke[s] it unchanged").  Skipped when the reference files are not present."""
import os
import re
import socket
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch

import nerfpp_oracle as O

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref", "nerfplusplus")
LAUNCHER = os.path.join(ROOT, "outdoor-nerf-depth_b200", "launch_ddp_train_nerf.py")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run_trainer(base, world, iters, extra=()):
    cmd = [sys.executable, LAUNCHER, "--reference", REF, "--config", os.path.join(REF, "configs", "kitti.txt"),
           "--datadir", base, "--scene", "synth_learnable", "--expname", "run_w%d" % world, "--basedir", os.path.join(base, "logs"),
           "--N_iters", str(iters), "--use_depth", "--lambda_depth", "0.1", "--depth_loss_type", "mse", "--depth_sup_type", "gt",
           "--trainskip", "1", "--testskip", "1", "--world_size", str(world), "--i_print", "1", "--i_weights", "1000000",
           "--port", str(_free_port())] + list(extra)
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=base)
    return p


def _scalars(log, step):
    """{name: value} of the trainer's log line for ``step`` (ddp_train_nerf.py:508-513)."""
    for line in log.splitlines():
        m = re.search(r" step: %d  (.*)$" % step, line)
        if m:
            it = re.findall(r"(\S+): (-?[\d.]+(?:e[+-]?\d+)?|nan|inf)", m.group(1))
            return {k: float(v) for k, v in it}
    return None


@pytest.fixture(scope="module")
def scene():
    if not os.path.isdir(REF):
        pytest.skip("reference files not present (baseline/_ref is made by oracle/install_reference.py in the build container)")
    import synth_scene
    base = tempfile.mkdtemp()
    synth_scene.write_scene(base)
    return base


def test_unmodified_trainer_runs_on_the_kernels(scene):
    """50 optimisation steps of the reference's own loop (ddp_train_nerf.py:417-503), world size 1: it must finish, log
    finite losses that go down, and its first logged losses must equal the oracle's for the same pixels and draws."""
    import train_harness as TH
    p = _run_trainer(scene, 1, 50)
    log = p.stdout + p.stderr
    assert p.returncode == 0, log[-3000:]
    s0, s49 = _scalars(log, 0), _scalars(log, 49)
    assert s0 is not None and s49 is not None, log[-3000:]
    assert all(np.isfinite(v) for v in s49.values()), s49
    assert s49["level_1/rgb_loss"] < 1e-3 * s0["level_1/rgb_loss"], (s0, s49)        # from ~1e7 (untrained background depth) to O(0.1)
    # the trainer's own wall clock per iteration (:418,504-505: 1024 rays, both levels, its .item() syncs included)
    its = [(_scalars(log, k) or {}).get("iter_time") for k in range(20, 50)]
    its = sorted(t for t in its if t is not None)
    print("unmodified trainer: median iter_time %.4f s over steps 20-49 (1024 rays/step => %.0f rays/s)" % (its[len(its) // 2], 1024 / its[len(its) // 2]))
    assert its[len(its) // 2] < 0.25
    # ---- step 0 against the oracle: replay the trainer's RNG (ddp_train_nerf.py:404-408, 423-424; nerf_sample_ray_split.py:178)
    views = TH.load_views(scene, "synth_learnable", "train")
    np.random.seed(777)
    i = np.random.randint(low=0, high=len(views))
    sel = np.random.choice(views[0]["H"] * views[0]["W"], size=(1024,), replace=False)
    dev = torch.device("cuda:0")
    torch.manual_seed(777)
    rand = {"t_fg": torch.rand(1024, 64, device=dev), "t_bg": torch.rand(1024, 64, device=dev),
            "u_fg_1": torch.rand(1024, 128, device=dev), "u_bg_1": torch.rand(1024, 128, device=dev)}
    b = TH.batch_of(views[i], sel, dev)
    levels = [{k: v.to(dev) for k, v in lv.items()} for lv in O.make_params_levels(2)]
    with torch.no_grad():
        # level 1 of step 0 already runs on level 0's UPDATED... no: the two levels are independent nets; level 1 only takes
        # level 0's weights-along-the-ray as its sampling pdf, computed before any optimiser step of that net
        out, far = TH.oracle_cascade(levels, b, rand, dev)
    ret, fg_z, _ = out[0]
    depth_loss = float(O.depth_mse(b["depth_sup"], ret["depth"]))
    total = float(O.img2mse(ret["rgb"], b["rgb"])) + 0.1 * depth_loss
    assert abs(s0["level_0/loss_depth"] - depth_loss) <= 2e-3 * abs(depth_loss), (s0, depth_loss)
    # the trainer logs `rgb_loss` AFTER `loss += lambda * depth_loss` on the same tensor (:482,493-495): it is the total
    assert abs(s0["level_0/rgb_loss"] - total) <= 2e-3 * abs(total), (s0, total)


def test_unmodified_trainer_two_ranks(scene):
    """The same with world size 2: the trainer's own DDP wrapping (gloo process group, :298,323) around the drop-in module,
    one process per GPU spawned by the launcher."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    p = _run_trainer(scene, 2, 20)
    log = p.stdout + p.stderr
    assert p.returncode == 0, log[-3000:]
    s0, s19 = _scalars(log, 0), _scalars(log, 19)
    assert s0 is not None and s19 is not None, log[-3000:]
    assert all(np.isfinite(v) for v in s19.values()) and s19["level_1/rgb_loss"] < 1e-2 * s0["level_1/rgb_loss"], (s0, s19)
