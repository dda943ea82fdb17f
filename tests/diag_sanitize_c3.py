"""Small-shape driver for compute-sanitizer on config 3's kernels (gemm_tc_kernel one-CTA / CTA-pair / split precision,
prop_chain_kernel, cast_encode_kernel, heads): one three-level model forward per precision mode on 48 rays, plus the
layer-by-layer PropMLP path and the N-split variant of the NeRF++ field kernel.  Run on the GPU box:

    compute-sanitizer --tool racecheck python tests/diag_sanitize_c3.py
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa: F401
import mip360_model_oracle as MM
from nerfpp_b200 import _lib
from nerfpp_b200.mip360_model import Model, Rays

dev = torch.device("cuda:0")
n = int(os.environ.get("RAYS", 48))
rays = MM.synthetic_rays(n, seed=4)
R = Rays(*(torch.from_numpy(rays[k]).to(dev) for k in ("origins", "directions", "viewdirs", "radii", "near", "far")))
L = _lib.lib()
L.mip360_debug_set_chain.argtypes = [ctypes.c_int]
for prec in (False, True):
    model = Model(dev, prec=prec).init(1)
    for chain in ((1, 0) if not prec else (1,)):
        L.mip360_debug_set_chain(chain)
        with torch.no_grad():
            rend, hist = model(None, R, train_frac=0.5)
        torch.cuda.synchronize()
        print("prec %d chain %d ok: rgb mean %.5f depth mean %.4f" % (prec, chain, float(rend[-1]["rgb"].mean()), float(rend[-1]["depth"].mean())))
L.mip360_debug_set_chain(1)
# the NeRF++ field kernel's column-half issue order (field_tc.cu: NSPLIT), fg and bg, on a small ragged batch
import nerfpp_oracle as O
from nerfpp_b200 import FIELD_TC, ops
from test_parity_gpu import make_models
L.nerfpp_debug_set_tc_nsplit.argtypes = [ctypes.c_int]
net = make_models([O.densify(O.make_params(), 5.0)])[0].nerf_net
r2 = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in O.synthetic_rays(40, seed=3).items()}
far = ops.intersect_sphere(r2["ray_o"], r2["ray_d"])
for is_bg, sub in ((0, net.fg_net), (1, net.bg_net)):
    z = torch.sort(torch.rand(40, 37, device="cuda"), -1)[0] * (1.0 if is_bg else far[:, None])
    pk = net._packed[is_bg].get(sub.tensors(), FIELD_TC)
    L.nerfpp_debug_set_tc_nsplit(1)
    o = ops.field_forward(pk, is_bg, r2["ray_o"], r2["ray_d"], z, FIELD_TC)
    torch.cuda.synchronize()
    L.nerfpp_debug_set_tc_nsplit(0)
    print("nsplit field %s ok: sigma mean %.5f" % ("bg" if is_bg else "fg", float(o[0].mean())))
