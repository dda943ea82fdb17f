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
