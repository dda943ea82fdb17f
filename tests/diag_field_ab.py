"""A/B of the fast field kernel between two builds of the library (NERFPP_B200_LIB selects one): 4096 x 192 fg and bg."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa
import nerfpp_oracle as O
from test_parity_gpu import make_models
from nerfpp_b200 import ops, FIELD_TC
net = make_models([O.densify(O.make_params(), 5.0)])[0].nerf_net
rays = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in O.synthetic_rays(4096, seed=0).items()}
far = ops.intersect_sphere(rays["ray_o"], rays["ray_d"])
res = []
for is_bg, sub in ((0, net.fg_net), (1, net.bg_net)):
    z = torch.sort(torch.rand(4096, 192, device="cuda"), -1)[0]
    if not is_bg:
        z = z * far[:, None]
    pk = net._packed[is_bg].get(sub.tensors(), FIELD_TC)
    for _ in range(5):
        ops.field_forward(pk, is_bg, rays["ray_o"], rays["ray_d"], z, FIELD_TC)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(40):
        ops.field_forward(pk, is_bg, rays["ray_o"], rays["ray_d"], z, FIELD_TC)
    b.record()
    torch.cuda.synchronize()
    res.append(a.elapsed_time(b) / 40)
print(os.environ.get("NERFPP_B200_LIB", "current"), "fg %.4f ms  bg %.4f ms" % tuple(res))
