"""Diagnostic (not collected): a few eager Model forwards of config 3 for an ncu launch list."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "outdoor-nerf-depth_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import mip360_model_oracle as MM
from nerfpp_b200.mip360_model import Model, Rays
dev = torch.device("cuda:0")
n = 4096
rays = MM.synthetic_rays(n, seed=0)
model = Model(dev, prec=bool(int(os.environ.get("PREC", 0)))).init(0)
R = Rays(*(torch.from_numpy(rays[k]).to(dev) for k in ("origins", "directions", "viewdirs", "radii", "near", "far")))
for _ in range(3):
    model(True, R, 0.5)
torch.cuda.synchronize()
