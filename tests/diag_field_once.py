"""Driver for `ncu --set full`: the four inference field launches of one step (coarse / fine x fg / bg) and one
split-precision launch, nothing else."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa: F401
import nerfpp_oracle as O
from nerfpp_b200 import FIELD_TC, FIELD_TC_SPLIT, cascade_forward, ops
from test_parity_gpu import make_models

n = 4096
nets = make_models([O.densify(p, 5.0) for p in O.make_params_levels(2)])
rays = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in O.synthetic_rays(n, seed=0).items()}
with torch.no_grad():
    for _ in range(int(os.environ.get("REPS", 3))):
        out, far = cascade_forward(nets, rays["ray_o"], rays["ray_d"], rays["min_depth"], (64, 128), train=True)
    net = nets[1].nerf_net
    pk = net._packed[0].get(net.fg_net.tensors(), FIELD_TC_SPLIT)
    ops.field_forward(pk, 0, rays["ray_o"], rays["ray_d"], out[-1][1], FIELD_TC_SPLIT)
torch.cuda.synchronize()
print("done")
