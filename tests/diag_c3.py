"""Diagnostic (not collected): config 3's step -- eager vs graphed timings, per-kernel breakdown with torch profiler off."""
import os, sys, time
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "outdoor-nerf-depth_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import numpy as np
import mip360_model_oracle as MM
from nerfpp_b200.mip360_model import Model, Rays, GraphedModelStep, IN_KEYS

dev = torch.device("cuda:0")
n = int(os.environ.get("N", 4096))
prec = bool(int(os.environ.get("PREC", 0)))
rays = MM.synthetic_rays(n, seed=0)
model = Model(dev, prec=prec).init(0)
R = Rays(*(torch.from_numpy(rays[k]).to(dev) for k in ("origins", "directions", "viewdirs", "radii", "near", "far")))
for _ in range(2):
    rend, hist = model(True, R, 0.5)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(10):
    rend, hist = model(True, R, 0.5)
e1.record(); torch.cuda.synchronize()
print("eager Model forward: %.3f ms / step" % (e0.elapsed_time(e1) / 10))
# per-level field time
for name, mlp, S in (("prop", model.prop_mlp, 64), ("nerf", model.nerf_mlp, 32)):
    sd = hist[0 if name == "prop" else 2]["sdist"]
    for _ in range(2): mlp.level(sd, R)
    e0.record()
    for _ in range(10): mlp.level(sd, R)
    e1.record(); torch.cuda.synchronize()
    print("  %s level (%d samples/ray): %.3f ms" % (name, S, e0.elapsed_time(e1) / 10))
step = GraphedModelStep(model, n, train_frac=0.5, host_io=False)
batch = {k: torch.from_numpy(rays[k]).to(dev) for k in ("origins", "directions", "viewdirs", "radii", "near", "far")}
batch["rgb"] = torch.rand(n, 3, device=dev); batch["disps_sup"] = torch.rand(n, 1, device=dev) * 5
for _ in range(3): out = step(batch)
torch.cuda.synchronize()
e0.record()
for _ in range(20): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print("graphed step (forward + losses): %.3f ms = %.2f M rays/s; kernels/replay %d; losses %s" % (ms, n / ms / 1e3, step.kernels_per_replay, out["losses"].tolist()))
