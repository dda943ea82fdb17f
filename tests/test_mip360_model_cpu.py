"""CPU-side checks of config 3's host logic (no GPU): layer shapes, the C ABI's pure host queries, the no-CPU-fallback rule,
and the reference arm's JSON contract for --config c3."""
import json
import os
import subprocess
import sys

import pytest

import mip360_model_oracle as MM

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dense_shapes_agree_with_the_oracle():
    from nerfpp_b200.mip360_model import dense_shapes
    for depth, width, rgb in ((4, 256, False), (8, 1024, True), (8, 256, True)):
        assert dense_shapes(depth, width, rgb) == MM.dense_shapes(depth, width, rgb)


def test_packed_and_workspace_queries_are_host_only():
    """mip360_mlp_packed_bytes / mip360_field_workspace_bytes are plain host arithmetic: callable without a device."""
    from nerfpp_b200 import _lib
    L = _lib.lib()
    prop = L.mip360_mlp_packed_bytes(4, 256, 0, 0)
    nerf = L.mip360_mlp_packed_bytes(8, 1024, 1, 0)
    nerf_split = L.mip360_mlp_packed_bytes(8, 1024, 1, 1)
    # fp16 images of the padded Dense kernels + fp32 biases and heads
    assert prop >= 2 * (512 * 256 + 3 * 256 * 256) and prop < 2 * (512 * 256 + 3 * 256 * 256) + 16384
    assert nerf >= 2 * (512 * 1024 + 6 * 1024 * 1024 + 1536 * 1024 + 1024 * 256 + 320 * 128)
    assert nerf_split > 1.9 * nerf - 65536
    assert L.mip360_mlp_packed_bytes(3, 100, 0, 0) == -1 and b"unsupported" in L.nerfpp_last_error()
    w1, w2 = L.mip360_field_workspace_bytes(1000, 8, 1024, 1, 0), L.mip360_field_workspace_bytes(2000, 8, 1024, 1, 0)
    assert 0 < w1 < w2 and w1 >= 1000 * (512 + 2 * 1024 + 256 + 64 + 128) * 2


def test_no_cpu_fallback():
    from nerfpp_b200 import NerfppError
    from nerfpp_b200.mip360_model import MLP, Model
    with pytest.raises(NerfppError):
        MLP(4, 256, True, "cpu")
    with pytest.raises(NerfppError):
        Model("cpu")


def test_reference_arm_c3_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "c3", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600)
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stderr[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "rays/s" and d["value"] > 0 and d["config"]["name"] == "c3"
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
