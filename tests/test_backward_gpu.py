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


def _field_case(n, S, seed, is_bg):
    params = O.densify(O.make_params(), 5.0)
    rays = O.synthetic_rays(n, seed=seed)
    far = O.intersect_sphere(rays["ray_o"], rays["ray_d"])
    g = torch.Generator().manual_seed(seed)
    z = torch.sort(torch.rand(n, S, generator=g), -1)[0]
    if not is_bg:
        z = z * far[:, None]
    return params, rays, z


def _oracle_acts(params, rays, z, is_bg):
    n, S = z.shape
    o = rays["ray_o"][:, None, :].expand(n, S, 3)
    d = rays["ray_d"][:, None, :].expand(n, S, 3)
    v = (rays["ray_d"] / torch.norm(rays["ray_d"], dim=-1, keepdim=True))[:, None, :].expand(n, S, 3)
    if is_bg:
        pts, _ = O.inverted_sphere_points(o, d, z)
        pe, ve = torch.flip(O.posenc(pts, 10), dims=[-2]), torch.flip(O.posenc(v, 4), dims=[-2])
    else:
        pe, ve = O.posenc(o + z[..., None] * d, 10), O.posenc(v, 4)
    return O.mlp_field_acts(params, "bg_net" if is_bg else "fg_net", pe, ve), pe, ve


@pytest.mark.parametrize("is_bg", [False, True])
def test_training_forward_saves_activations(is_bg):
    """nerfpp_field_forward_train: same outputs as the inference call, and the saved fp16 activations / encoded inputs /
    raw sigma agree with the oracle's intermediates (fp16 rounding of the operands: 1e-3 of the layer's scale)."""
    from test_parity_gpu import make_models
    from nerfpp_b200 import ops, FIELD_TC
    n, S = 37, 64                       # 2368 samples = 18.5 tiles: exercises the ragged last tile
    params, rays, z = _field_case(n, S, 11, is_bg)
    net = make_models([params])[0].nerf_net
    sub = net.bg_net if is_bg else net.fg_net
    packed = net._packed[int(is_bg)].get(sub.tensors(), FIELD_TC)
    C = lambda x: x.cuda()
    sig0, rgb0, _ = ops.field_forward(packed, is_bg, C(rays["ray_o"]), C(rays["ray_d"]), C(z), FIELD_TC)
    sig, rgb, _, ws = ops.field_forward_train(packed, is_bg, C(rays["ray_o"]), C(rays["ray_d"]), C(z))
    assert torch.equal(sig, sig0) and torch.equal(rgb, rgb0)
    with torch.no_grad():
        (acts, raw_sigma, _), pe, ve = _oracle_acts(params, rays, z, is_bg)
    total = n * S
    T = (total + 127) // 128
    off = 0
    for l in range(10):
        nch = 2 if l == 9 else 4
        nb = T * nch * 16384
        got = ops.unpack_chunks(ws[off:off + nb], T, nch)[:total].float().cpu()
        want = acts[l].reshape(total, -1)
        assert got.shape == want.shape
        assert float((got - want).abs().max()) <= 2e-3 * max(float(want.abs().max()), 1.0), l
        off += nb
    e = ops.unpack_chunks(ws[off:off + T * 2 * 16384], T, 2)[:total].float().cpu()
    off += T * 2 * 16384
    emb = pe.shape[-1]
    assert float((e[:, :emb] - pe.reshape(total, emb)).abs().max()) <= 1.5e-3
    assert float((e[:, 96:123] - ve.reshape(total, 27)).abs().max()) <= 1.5e-3
    assert torch.all(e[:, 123:125] == 1) and torch.all(e[:, 125:] == 0)
    rs = ws[off:off + T * 128 * 4].view(torch.float32)[:total].cpu()
    assert relerr(rs.numpy(), raw_sigma.reshape(-1).numpy()) <= 1e-4
    assert torch.equal(rs.abs(), sig.reshape(-1).cpu())
