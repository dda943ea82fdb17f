"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the
header declares, the shim modules mirror the reference's import surface / state dict / RNG
stream, and the product path refuses to run without CUDA (no fallback)."""
import hashlib
import os
import re

import numpy as np
import pytest
import torch

import nerfpp_oracle as O
from conftest import ROOT, reference_args


def header_functions():
    src = open(os.path.join(ROOT, "include", "nerfpp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nerfpp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from nerfpp_b200 import _lib
    L = _lib.lib()
    names = header_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), "libnerfpp_b200.so lacks %s" % n
    bound = set(_lib.SIGNATURES) | set(_lib.OPTIONAL)
    assert set(names) <= bound, "ctypes binding lacks %s" % (set(names) - bound)
    assert L.nerfpp_abi_version() == 1
    assert L.nerfpp_packed_bytes(0, _lib.FIELD_SIMT) > 4 * 590000
    assert L.nerfpp_packed_bytes(1, _lib.FIELD_TC) > 2 * 590000
    assert L.nerfpp_forward_workspace_bytes(4096, 192, 192) >= 4096 * 192 * 4 * 9


def test_argument_errors_return_codes_not_exceptions():
    from nerfpp_b200 import _lib
    L = _lib.lib()
    rc = L.nerfpp_intersect_sphere(None, None, 4, None, None, None)
    assert rc < 0 and b"intersect_sphere" in L.nerfpp_last_error()
    assert L.nerfpp_packed_bytes(0, 7) == -1


def test_library_is_sm100a_tcgen05():
    """The shipped binary holds sm_100a SASS with tensor-core MMA (UTCHMMA) and bulk-copy (UBLKCP) ops."""
    import shutil
    import subprocess
    from nerfpp_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", "-fun", "field_tc_kernel", _lib.LIB_PATH], capture_output=True, text=True).stdout
    if "UTCHMMA" not in sass:
        sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "UTCHMMA" in sass and "LDTM" in sass and "UBLKCP" in sass


def test_shim_surface_and_state_dict_match_reference_layout():
    import ddp_model
    import depth_loss
    for name in ("NerfNetWithAutoExpo", "NerfNet", "depth2pts_outside", "remap_name"):
        assert hasattr(ddp_model, name)
    for name in ("depth_mse", "depth_l1", "depth_kl", "depth_light_of_sight", "depth_gaussian_log_likelihood"):
        assert hasattr(depth_loss, name)
    torch.manual_seed(777)
    nets = [ddp_model.NerfNetWithAutoExpo(reference_args()) for _ in range(2)]
    ref_levels = O.make_params_levels(2)
    for net, ref in zip(nets, ref_levels):
        sd = net.state_dict()
        assert list(sd.keys()) == list(ref.keys())
        for k in ref:
            assert sd[k].shape == ref[k].shape
            assert torch.equal(sd[k], ref[k]), k     # same RNG stream as the reference's create_nerf
    assert sum(p.numel() for p in nets[0].parameters()) == 1202440
    # golden digest of the unmodified reference's parameters
    g = np.load(os.path.join(ROOT, "tests", "golden", "nerfpp_c2_train_dense.npz"))
    h = hashlib.sha256()
    sd = nets[0].state_dict()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].numpy().astype(np.float32).tobytes())
    assert h.hexdigest() == str(g["meta_param_digest"][0])
    # a reference-style checkpoint (DDP 'module.' prefix stripped by the loader) loads strictly
    nets[1].load_state_dict({k: v.clone() for k, v in ref_levels[0].items()}, strict=True)


def test_autoexpo_and_remap_name():
    import ddp_model
    assert ddp_model.remap_name("/data/kitti/seq00/train/rgb/000001.png") == "train/rgb/000001-png"
    net = ddp_model.NerfNetWithAutoExpo(reference_args(), optim_autoexpo=True, img_names=["a/b/c/d.png"])
    assert list(net.autoexpo_params.keys()) == ["b/c/d-png"]


def test_unsupported_network_shape_is_refused():
    import ddp_model
    a = reference_args()
    a.netwidth = 128
    with pytest.raises(ValueError):
        ddp_model.NerfNetWithAutoExpo(a)


def test_no_cpu_fallback():
    """CPU tensors must fail loudly: the path exists only as CUDA kernels."""
    from nerfpp_b200 import NerfppError, ops
    import ddp_model
    import depth_loss
    with pytest.raises(NerfppError):
        ops.intersect_sphere(torch.zeros(4, 3), torch.ones(4, 3))
    net = ddp_model.NerfNetWithAutoExpo(reference_args())
    n, S = 4, 8
    with pytest.raises(NerfppError):
        with torch.no_grad():
            net(torch.zeros(n, 3), torch.ones(n, 3), torch.ones(n), torch.rand(n, S), torch.rand(n, S))
    with pytest.raises(NerfppError):
        depth_loss.depth_mse(torch.ones(4), torch.ones(4))
    # the graph-captured step and the device-resident sampler refuse a CPU device as well
    from nerfpp_b200 import GraphedRenderStep
    from nerfpp_b200.ray_sampler import decode_pixels
    with pytest.raises(NerfppError):
        GraphedRenderStep([net, net], 16, device="cpu")
    with pytest.raises(NerfppError):
        decode_pixels(np.zeros((2, 2), np.uint8), 255.0, device="cpu")


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "outdoor-nerf-depth_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "nerfpp_oracle" not in src and "oracle/" not in src.replace("oracle/_ref", ""), f


def test_launcher_patches_trainer_functions():
    """launch_ddp_train_nerf.patch_trainer_module rebinds the three functions the reference defines inside its
    trainer file (ddp_train_nerf.py:51-130) -- checked on a stub module, the real trainer's deps are not in this image."""
    import types
    import launch_ddp_train_nerf as LN
    from nerfpp_b200 import ops
    stub = types.ModuleType("ddp_train_nerf")
    stub.intersect_sphere = stub.perturb_samples = stub.sample_pdf = lambda *a, **k: None
    LN.patch_trainer_module(stub)
    assert stub.intersect_sphere is ops.intersect_sphere
    assert stub.perturb_samples is ops.perturb_samples
    assert stub.sample_pdf is ops.sample_pdf


def test_launcher_can_keep_the_reference_loader(tmp_path):
    """NERFPP_REFERENCE_LOADER=1: the name ``data_loader_split`` resolves to the reference directory's file although the
    drop-in directory (which holds a module of the same name) is first on sys.path."""
    import sys
    import launch_ddp_train_nerf as LN
    (tmp_path / "data_loader_split.py").write_text("WHO = 'reference stub'\ndef load_data_split(*a, **k):\n    return WHO\n")
    saved = sys.modules.pop("data_loader_split", None)
    try:
        mod = LN._use_reference_loader(str(tmp_path))
        import data_loader_split as again
        assert again is mod and again.load_data_split() == "reference stub"
    finally:
        sys.modules.pop("data_loader_split", None)
        if saved is not None:
            sys.modules["data_loader_split"] = saved
    import data_loader_split as ours
    assert ours.__file__.endswith(os.path.join("outdoor-nerf-depth_b200", "data_loader_split.py"))


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference (the reference's algorithm on the host cores; the one leg that needs no GPU): stdout
    carries exactly one JSON line with the contract's keys, whatever libraries print."""
    import json
    import subprocess
    import sys
    env = dict(os.environ, WORLD_SIZE="1", RANK="0")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "config",
              "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "rays/s" and d["value"] > 0
    # "reference" = the reference's own files were found (baseline/_ref or /root/reference) and timed; "port" = the oracle port
    assert d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["config"]["rays_per_gpu"] == 4096 and d["steps"] == 1 and d["warmup"] == 1
    # a rank other than 0 does no work and prints nothing
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1"],
                       capture_output=True, text=True, timeout=600, env=dict(env, RANK="1", WORLD_SIZE="2"))
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_missing_library_fails_loudly():
    """No silent fallback: with the shared library absent, the first call through the binding raises NerfppError that
    says how to build it (checked in a child process so this process's loaded library is left alone)."""
    import subprocess
    import sys
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from nerfpp_b200 import _lib\n"
        "_lib.LIB_PATH = '/nonexistent/libnerfpp_b200.so'\n"
        "_lib._lib = None\n"
        "try:\n"
        "    _lib.lib()\n"
        "except _lib.NerfppError as e:\n"
        "    assert 'not built' in str(e) and 'no CPU or PyTorch fallback' in str(e), str(e)\n"
        "    print('raised')\n" % os.path.join(ROOT, "outdoor-nerf-depth_b200"))
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and p.stdout.strip() == "raised", p.stderr[-1500:]
