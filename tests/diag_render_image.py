"""Diagnostic (not a test): times the drop-in render_single_image (ddp_train_nerf.py:133-249) on one Tanks&Temples-Truck
camera (980 x 546 = 535 080 rays, the reference's shipped camera, tests/golden/rays_tat_truck.npz) through the
device-resident sampler, and splits the time into render (device) and the final D2H.  Run on the GPU box."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa
import nerfpp_oracle as O
from nerfpp_b200 import render_single_image
from nerfpp_b200.ray_sampler import DeviceRaySampler
from test_parity_gpu import make_models

g = dict(np.load(os.path.join(conftest.GOLDEN, "rays_tat_truck.npz")))
H, W = (int(x) for x in g["hw0"])
levels = [O.densify(p, 5.0) for p in O.make_params_levels(2)]
nets = make_models(levels)
models = {"cascade_level": 2, "cascade_samples": [64, 128], "net_0": nets[0], "net_1": nets[1]}
s = DeviceRaySampler(H, W, g["K0"], g["c2w0"])
for chunk in (int(x) for x in os.environ.get("CHUNKS", "8192,32768").split(",")):
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = render_single_image(0, 1, models, s, chunk)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    n = H * W
    t1 = time.perf_counter()
    fd = out[-1]["fg_dists"]
    t_fd = time.perf_counter() - t1
    print("chunk %6d: %.3f s per %dx%d image = %.2f M rays/s (both levels, all 8 keys returned as CPU tensors); rgb %s; fetching fg_dists %s on demand: %.3f s"
          % (chunk, dt, W, H, n / dt / 1e6, tuple(out[-1]["rgb"].shape), tuple(fd.shape), t_fd))
