"""Diagnostic (not a test): prints the error table of the CUDA path against the golden vectors. Run on the GPU box."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa
import nerfpp_oracle as O
from test_parity_gpu import CASES, KEYS, G, golden_levels, load, make_models, relerr, impl_id

gd = conftest.GOLDEN
for name in CASES:
    g = load(gd, name)
    cascade, levels = golden_levels(g)
    nets = make_models(levels)
    for impl in ("simt", "tc"):
        for m in range(len(cascade)):
            with torch.no_grad():
                ret = nets[m](G(g["ray_o"]), G(g["ray_d"]), G(g["fg_far"]), G(g["fg_z_%d" % m]), G(g["bg_z_%d" % m]), impl=impl_id(impl))
            print(name, impl, m, " ".join("%s=%.1e" % (k, relerr(ret[k].cpu().numpy(), g["ret%d_%s" % (m, k)])) for k in KEYS))
