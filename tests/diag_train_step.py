"""Diagnostic (not a test): times one trainer step (ddp_train_nerf.py:432-498: per cascade level forward + loss + backward
+ Adam) on 4096 rays through the drop-in modules, and prints the split.  Run on the GPU box."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa
import nerfpp_oracle as O
import depth_loss as DL
from nerfpp_b200 import ops
from test_parity_gpu import make_models

import ctypes
from nerfpp_b200 import _lib as _L
if os.environ.get("BWD_MODE"):
    _L.lib().nerfpp_debug_set_bwd_mode.argtypes = [ctypes.c_int]
    _L.lib().nerfpp_debug_set_bwd_mode(int(os.environ["BWD_MODE"]))
if os.environ.get("BWD_CONS"):
    _L.lib().nerfpp_debug_set_bwd_consumers.argtypes = [ctypes.c_int]
    _L.lib().nerfpp_debug_set_bwd_consumers(int(os.environ["BWD_CONS"]))
if os.environ.get("HEADS_FOLDED"):
    _L.lib().nerfpp_debug_set_heads_folded.argtypes = [ctypes.c_int]
    _L.lib().nerfpp_debug_set_heads_folded(int(os.environ["HEADS_FOLDED"]))
dev = torch.device("cuda:0")
n = int(os.environ.get("RAYS", 4096))
levels = [O.densify(p, 5.0) for p in O.make_params_levels(2)]
nets = make_models(levels)
opts = [torch.optim.Adam(net.parameters(), lr=5e-4) for net in nets]
rays = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in O.synthetic_rays(n, seed=0).items()}


def step(timers=None):
    far = ops.intersect_sphere(rays["ray_o"], rays["ray_d"])
    fg_z = bg_z = ret = None
    for m, S in enumerate((64, 128)):
        if m == 0:
            fg_z, bg_z = ops.coarse_depths(rays["min_depth"], far, S, torch.rand(n, S, device=dev), torch.rand(n, S, device=dev))
        else:
            fg_z, bg_z = ops.resample_merge_pair(fg_z, ret["fg_weights"].detach(), bg_z, ret["bg_weights"].detach(), S)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        opts[m].zero_grad()
        ev[0].record()
        ret = nets[m](rays["ray_o"], rays["ray_d"], far, fg_z, bg_z)
        ev[1].record()
        loss = torch.mean((ret["rgb"] - rays["rgb"]) ** 2) + 0.1 * DL.depth_mse(rays["depth_sup"], ret["depth"])
        loss.backward()
        ev[2].record()
        opts[m].step()
        ev[3].record()
        if timers is not None:
            timers.append(ev)
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
T = []
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
K = 5
for _ in range(K):
    l = step(T)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / K
f = sum(e[0].elapsed_time(e[1]) for e in T) / K
bw = sum(e[1].elapsed_time(e[2]) for e in T) / K
ad = sum(e[2].elapsed_time(e[3]) for e in T) / K
print("train step (both levels) %.2f ms = %.0f rays/s | forward %.2f  loss+backward %.2f  adam %.2f | loss %.5f | peak mem %.1f GB"
      % (ms, n / ms * 1e3, f, bw, ad, float(l), torch.cuda.max_memory_allocated() / 2**30))
