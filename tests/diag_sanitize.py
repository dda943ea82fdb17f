"""Small-shape driver for compute-sanitizer (memcheck / racecheck / synccheck): one inference cascade (fast and split-precision
field), one training step per backward variant.  Run on the GPU box:

    compute-sanitizer --tool racecheck python tests/diag_sanitize.py
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa: F401
import nerfpp_oracle as O
import depth_loss as DL
from nerfpp_b200 import FIELD_TC, FIELD_TC_SPLIT, _lib, cascade_forward, ops
from test_parity_gpu import make_models

n = int(os.environ.get("RAYS", 24))
levels = [O.densify(p, 5.0) for p in O.make_params_levels(2)]
nets = make_models(levels)
rays = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in O.synthetic_rays(n, seed=3).items()}
with torch.no_grad():
    for impl in (FIELD_TC, FIELD_TC_SPLIT):
        out, far = cascade_forward(nets, rays["ray_o"], rays["ray_d"], rays["min_depth"], (64, 128), train=False, impl=impl)
        torch.cuda.synchronize()
        print("forward impl %d ok: rgb mean %.5f" % (impl, float(out[-1][0]["rgb"].mean())))
L = _lib.lib()
L.nerfpp_debug_set_bwd_mode.argtypes = [ctypes.c_int]
fg_z, bg_z = out[-1][1], out[-1][2]
for mode in (2, 0, 1):
    L.nerfpp_debug_set_bwd_mode(mode)
    net = nets[1]
    net.zero_grad()
    o = net(rays["ray_o"], rays["ray_d"], far, fg_z, bg_z)
    loss = torch.mean((o["rgb"] - rays["rgb"]) ** 2) + 0.1 * DL.depth_mse(rays["depth_sup"], o["depth"])
    loss.backward()
    torch.cuda.synchronize()
    print("backward mode %d ok: loss %.5f |grad| %.4e" % (mode, float(loss), float(sum(p.grad.abs().sum() for p in net.parameters()))))
L.nerfpp_debug_set_bwd_mode(2)
