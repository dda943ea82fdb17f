"""Backward parity on the GPU: the CUDA backward kernels (through the C ABI) against torch autograd of the oracle."""
import numpy as np
import pytest
import torch

import nerfpp_oracle as O

pytestmark = pytest.mark.gpu


def relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def _composite_case(n, sf, sb, seed):
    g = torch.Generator().manual_seed(seed)
    rays = O.synthetic_rays(n, seed=seed)
    far = O.intersect_sphere(rays["ray_o"], rays["ray_d"])
    fg_z = torch.sort(torch.rand(n, sf, generator=g), -1)[0] * far[:, None]
    bg_z = torch.sort(torch.rand(n, sb, generator=g), -1)[0]
    t = dict(fg_sigma=torch.rand(n, sf, generator=g) * 30, fg_rgb=torch.rand(n, sf, 3, generator=g),
             bg_sigma=torch.rand(n, sb, generator=g) * 30, bg_rgb=torch.rand(n, sb, 3, generator=g),
             bg_dr=torch.rand(n, sb, generator=g) * 5 + 1)
    t["fg_sigma"][:, ::7] = 0.05      # a few nearly transparent samples
    return rays, far, fg_z, bg_z, t


@pytest.mark.parametrize("n,sf,sb", [(64, 64, 64), (33, 192, 192), (5, 70, 40)])
def test_composite_backward_matches_autograd(n, sf, sb):
    from nerfpp_b200 import ops
    rays, far, fg_z, bg_z, t = _composite_case(n, sf, sb, seed=sf + n)
    leaves = {k: v.clone().requires_grad_(True) for k, v in t.items() if k != "bg_dr"}
    ret = O.composite(leaves["fg_sigma"], leaves["fg_rgb"], leaves["bg_sigma"], leaves["bg_rgb"], t["bg_dr"], rays["ray_d"], far, fg_z, bg_z)
    g = torch.Generator().manual_seed(3)
    up = {k: torch.randn(ret[k].shape, generator=g) for k in ret if k != "fg_dists"}
    loss = sum((ret[k] * up[k]).sum() for k in up)
    ref = torch.autograd.grad(loss, [leaves[k] for k in ("fg_sigma", "fg_rgb", "bg_sigma", "bg_rgb")], retain_graph=True)
    C = lambda x: x.cuda()
    got = ops.composite_backward(C(rays["ray_d"]), C(far), C(fg_z), C(bg_z), C(t["fg_sigma"]), C(t["fg_rgb"]), C(t["bg_sigma"]),
                                 C(t["bg_rgb"]), C(t["bg_dr"]), C(ret["bg_lambda"].detach()), {k: C(v) for k, v in up.items()})
    for name, a, b in zip(("d_fg_sigma", "d_fg_rgb", "d_bg_sigma", "d_bg_rgb"), got, ref):
        assert relerr(a.cpu().numpy(), b.numpy()) <= 2e-5, name
    # only the gradients the trainer actually produces (rgb + depth, ddp_train_nerf.py:481-493)
    up2 = {"rgb": up["rgb"], "depth": up["depth"]}
    loss2 = sum((ret[k] * up2[k]).sum() for k in up2)
    ref2 = torch.autograd.grad(loss2, [leaves[k] for k in ("fg_sigma", "fg_rgb", "bg_sigma", "bg_rgb")])
    got2 = ops.composite_backward(C(rays["ray_d"]), C(far), C(fg_z), C(bg_z), C(t["fg_sigma"]), C(t["fg_rgb"]), C(t["bg_sigma"]),
                                  C(t["bg_rgb"]), C(t["bg_dr"]), C(ret["bg_lambda"].detach()), {k: C(v) for k, v in up2.items()})
    for name, a, b in zip(("d_fg_sigma", "d_fg_rgb", "d_bg_sigma", "d_bg_rgb"), got2, ref2):
        assert relerr(a.cpu().numpy(), b.numpy()) <= 2e-5, name
