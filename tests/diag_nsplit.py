"""Diagnostic (not collected): the fast field kernel with and without the column-half issue order -- outputs must be equal,
timings side by side in one process (4096 x 192 and 4096 x 64 samples, fg and bg)."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa
import nerfpp_oracle as O
from test_parity_gpu import make_models
from nerfpp_b200 import ops, FIELD_TC, _lib
L = _lib.lib()
L.nerfpp_debug_set_tc_nsplit.argtypes = [ctypes.c_int]
net = make_models([O.densify(O.make_params(), 5.0)])[0].nerf_net
rays = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in O.synthetic_rays(4096, seed=0).items()}
far = ops.intersect_sphere(rays["ray_o"], rays["ray_d"])
for S in (192, 64, 7):
    for is_bg, sub in ((0, net.fg_net), (1, net.bg_net)):
        z = torch.sort(torch.rand(4096, S, device="cuda"), -1)[0]
        if not is_bg:
            z = z * far[:, None]
        pk = net._packed[is_bg].get(sub.tensors(), FIELD_TC)
        outs, times = [], []
        for mode in (0, 1, 0, 1):
            L.nerfpp_debug_set_tc_nsplit(mode)
            for _ in range(3):
                o = ops.field_forward(pk, is_bg, rays["ray_o"], rays["ray_d"], z, FIELD_TC)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(30):
                o = ops.field_forward(pk, is_bg, rays["ray_o"], rays["ray_d"], z, FIELD_TC)
            b.record()
            torch.cuda.synchronize()
            times.append(a.elapsed_time(b) / 30)
            outs.append([t.clone() for t in o if torch.is_tensor(t)])
        same = all(torch.equal(x, y) for x, y in zip(outs[0], outs[1]))
        md = max(float((x - y).abs().max()) for x, y in zip(outs[0], outs[1]))
        print("S %3d %s  plain %.4f / %.4f ms   nsplit %.4f / %.4f ms   equal %s (max diff %.2e)" % (S, "bg" if is_bg else "fg", times[0], times[2], times[1], times[3], same, md), flush=True)
L.nerfpp_debug_set_tc_nsplit(0)
