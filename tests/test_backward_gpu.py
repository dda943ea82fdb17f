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


@pytest.mark.parametrize("loss_type", ["mse", "kl"])
def test_full_backward_matches_oracle_autograd(loss_type):
    """loss.backward() through the drop-in module (nerfpp_forward_train + nerfpp_backward) vs torch autograd of the oracle,
    for the trainer's loss (ddp_train_nerf.py:481-493).  Operands inside the kernels are fp16 under two power-of-two loss
    scales per net (backward.cu): every gradient tensor must have cosine >= 0.999 and a norm within 2 % of autograd's,
    and its worst single entry must lie within 6 % (mse; measured 4.0 %) / 20 % (kl; measured 14 %) of the tensor's
    LARGEST entry -- the KL prior's gradient -1/(w + 1e-5) spans eight decades across the samples of a batch, and what is
    small next to the largest sample's contribution is carried with few fp16 bits."""
    from test_parity_gpu import make_models
    import depth_loss as DL
    n, cascade = 96, (64, 128)
    params = O.densify(O.make_params(), 5.0)
    rays = O.synthetic_rays(n, seed=21)
    far = O.intersect_sphere(rays["ray_o"], rays["ray_d"])
    g = torch.Generator().manual_seed(4)
    fg_z = torch.sort(torch.rand(n, 96, generator=g), -1)[0] * far[:, None]
    bg_z = torch.sort(torch.rand(n, 80, generator=g), -1)[0]
    sig = 0.01 * 0.05
    # oracle
    p_ref = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    ret = O.nerfpp_forward(p_ref, rays["ray_o"], rays["ray_d"], far, fg_z, bg_z)
    loss_ref, _, _ = O.level_loss(ret, rays["rgb"], rays["depth_sup"], fg_z, far, True, loss_type, 0.1, sig)
    g_ref = dict(zip(p_ref.keys(), torch.autograd.grad(loss_ref, list(p_ref.values()))))
    # CUDA
    net = make_models([params])[0]
    C = lambda x: x.cuda()
    out = net(C(rays["ray_o"]), C(rays["ray_d"]), C(far), C(fg_z), C(bg_z))
    rgb_loss = torch.mean((out["rgb"] - C(rays["rgb"])) ** 2)
    if loss_type == "kl":
        dl = DL.depth_kl(out["fg_weights"], C(rays["depth_sup"]), C(fg_z), out["fg_dists"], sig, C(far))
    else:
        dl = DL.depth_mse(C(rays["depth_sup"]), out["depth"])
    loss = rgb_loss + 0.1 * dl
    assert abs(float(loss) - float(loss_ref)) <= 1e-4 * abs(float(loss_ref))
    loss.backward()
    worst, errs = 0.0, {}
    for name, p in net.named_parameters():
        ref = g_ref[name]
        assert p.grad is not None, name
        err = float((p.grad.cpu() - ref).abs().max() / max(float(ref.abs().max()), 1e-20))
        worst = max(worst, err)
        cos = float((p.grad.cpu().double() * ref.double()).sum() / (p.grad.cpu().double().norm() * ref.double().norm() + 1e-300))
        print("%-50s max-rel-err %.2e  cos %.6f  |ref|max %.2e" % (name, err, cos, float(ref.abs().max())))
        errs[name] = err
        assert cos >= 0.999, (name, cos)        # a layout / indexing bug destroys the direction; fp16 noise does not
        nr = float(p.grad.norm()) / max(float(ref.norm()), 1e-30)
        assert abs(nr - 1) <= 2e-2, (name, nr)
    print("worst relative gradient error", worst)
    # fp16 operands (11-bit significand) through nine layers of sums with heavy cancellation: individual entries of a
    # gradient tensor carry noise of a few percent of the tensor's largest entry, like any mixed-precision backward
    assert worst <= (0.2 if loss_type == "kl" else 0.06), max(errs, key=errs.get)


def test_backward_matches_reference_golden_gradients(golden_dir):
    """Gradient norms and leading entries recorded from the UNMODIFIED reference's autograd (oracle/gen_golden.py,
    case c2_train_dense) vs loss.backward() through the CUDA path, for both cascade levels and all three depth losses."""
    import os
    from test_parity_gpu import make_models
    import depth_loss as DL
    g = dict(np.load(os.path.join(golden_dir, "nerfpp_c2_train_dense.npz")))
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    levels = [O.densify(p, 5.0) for p in O.make_params_levels(2)]
    sig = float(g["meta_depth_sigma"]) * float(g["meta_depth_scale"])
    for m in range(2):
        for lt in ("mse", "l1", "kl"):
            net = make_models([levels[m]])[0]
            fg_z, bg_z = T(g["fg_z_%d" % m]), T(g["bg_z_%d" % m])
            out = net(T(g["ray_o"]), T(g["ray_d"]), T(g["fg_far"]), fg_z, bg_z)
            loss = torch.mean((out["rgb"] - T(g["rgb_gt"])) ** 2)
            if lt == "kl":
                loss = loss + 0.1 * DL.depth_kl(out["fg_weights"], T(g["depth_sup"]), fg_z, out["fg_dists"], sig, T(g["fg_far"]))
            else:
                loss = loss + 0.1 * (DL.depth_mse if lt == "mse" else DL.depth_l1)(T(g["depth_sup"]), out["depth"])
            loss.backward()
            for name, p in net.named_parameters():
                short = name.replace("nerf_net.", "").replace("_layers", "").replace(".weight", ".w").replace(".bias", ".b")
                ref_norm = float(g["grad%d_%s_norm/%s" % (m, lt, short)])
                assert abs(float(p.grad.norm()) - ref_norm) <= 3e-2 * max(ref_norm, 1e-12), (m, lt, short, float(p.grad.norm()), ref_norm)
                head = g["grad%d_%s_head/%s" % (m, lt, short)]
                got = p.grad.reshape(-1)[:16].cpu().numpy()
                assert np.abs(got - head).max() <= 0.1 * max(np.abs(head).max(), 1e-3 * ref_norm), (m, lt, short)


def test_ddp_wrapped_training_step_reduces_loss():
    """The drop-in module under the trainer's own wrapping (ddp_train_nerf.py:322-325: .to(rank), DistributedDataParallel(
    find_unused_parameters=True), Adam): a few optimiser steps on a fixed batch must lower the loss, gradients of the
    unused autoexpo-free module must all be populated, and reference-named checkpoints must round-trip."""
    import os
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    from test_parity_gpu import make_models
    import depth_loss as DL
    from nerfpp_b200 import ops
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", str(29600 + os.getpid() % 300))
    dist.init_process_group("nccl", rank=0, world_size=1)
    try:
        torch.manual_seed(777)
        net = make_models([O.densify(O.make_params(), 5.0)])[0]
        ddp = DDP(net, device_ids=[0], output_device=0, find_unused_parameters=True)
        optim = torch.optim.Adam(ddp.parameters(), lr=5e-4)
        n = 512
        rays = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in O.synthetic_rays(n, seed=5).items()}
        far = ops.intersect_sphere(rays["ray_o"], rays["ray_d"])
        fg_z, bg_z = ops.coarse_depths(rays["min_depth"], far, 64)
        losses = []
        for _ in range(12):
            optim.zero_grad()
            ret = ddp(rays["ray_o"], rays["ray_d"], far, fg_z, bg_z, "img.png")
            loss = torch.mean((ret["rgb"] - rays["rgb"]) ** 2) + 0.1 * DL.depth_mse(rays["depth_sup"], ret["depth"])
            loss.backward()
            optim.step()
            losses.append(float(loss.detach()))
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in ddp.parameters())
        assert losses[-1] < 0.9 * losses[0], losses
        sd = ddp.state_dict()
        assert "module.nerf_net.fg_net.base_layers.0.0.weight" in sd        # the reference's checkpoint key (ddp_train_nerf.py:642-652)
        make_models([O.make_params()])[0].load_state_dict({k[len("module."):]: v for k, v in sd.items()})
    finally:
        dist.destroy_process_group()


def test_backward_handles_ragged_and_tiny_batches():
    from test_parity_gpu import make_models
    net = make_models([O.densify(O.make_params(), 5.0)])[0]
    for n, sf, sb in ((1, 64, 64), (3, 17, 5)):
        rays = O.synthetic_rays(n, seed=n)
        far = O.intersect_sphere(rays["ray_o"], rays["ray_d"])
        g = torch.Generator().manual_seed(n)
        fg_z = (torch.sort(torch.rand(n, sf, generator=g), -1)[0] * far[:, None]).cuda()
        bg_z = torch.sort(torch.rand(n, sb, generator=g), -1)[0].cuda()
        net.zero_grad()
        out = net(rays["ray_o"].cuda(), rays["ray_d"].cuda(), far.cuda(), fg_z, bg_z)
        out["rgb"].sum().backward()
        p_ref = {k: v.clone().requires_grad_(True) for k, v in O.densify(O.make_params(), 5.0).items()}
        ref = O.nerfpp_forward(p_ref, rays["ray_o"], rays["ray_d"], far, fg_z.cpu(), bg_z.cpu())
        g_ref = dict(zip(p_ref.keys(), torch.autograd.grad(ref["rgb"].sum(), list(p_ref.values()))))
        for name, p in net.named_parameters():
            r = g_ref[name]
            cos = float((p.grad.cpu().double() * r.double()).sum() / (p.grad.cpu().double().norm() * r.double().norm() + 1e-300))
            assert cos >= 0.99, (n, name, cos)      # 64 samples: no averaging of the fp16 rounding noise over a batch


def _set_bwd_mode(mode, consumers=None):
    import ctypes
    from nerfpp_b200 import _lib
    L = _lib.lib()
    L.nerfpp_debug_set_bwd_mode.argtypes = [ctypes.c_int]
    L.nerfpp_debug_set_bwd_mode(mode)
    if consumers is not None:
        L.nerfpp_debug_set_bwd_consumers.argtypes = [ctypes.c_int]
        L.nerfpp_debug_set_bwd_consumers(consumers)


@pytest.mark.parametrize("n,sf,sb", [(3, 17, 5), (96, 64, 64), (700, 192, 192), (2048, 192, 192)])
def test_backward_kernel_variants_agree(n, sf, sb):
    """Three ways to run the field's backward (backward.cu: nerfpp_debug_set_bwd_mode) must produce the same gradients:
    2 = the default (dgrad with storer-warp staging, then wgrad), 1 = round 1's dgrad kernel, 0 = the fused
    producer/consumer kernel (dZ through an L2-resident slot ring, bwd_fused.cu).  The same fp16 dZ values meet the same
    fp16 activations; only the order of the fp32 split-K accumulation (atomics) differs, so every gradient tensor must
    agree to 5e-4 of its largest entry -- from a one-tile batch (fewer tiles than producers) over ragged tile counts to
    thousands of tiles per launch."""
    from test_parity_gpu import make_models
    net = make_models([O.densify(O.make_params(), 5.0)])[0]
    rays = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in O.synthetic_rays(n, seed=n).items()}
    far = O.intersect_sphere(rays["ray_o"].cpu(), rays["ray_d"].cpu()).cuda()
    g = torch.Generator().manual_seed(n)
    fg_z = (torch.sort(torch.rand(n, sf, generator=g), -1)[0]).cuda() * far[:, None]
    bg_z = torch.sort(torch.rand(n, sb, generator=g), -1)[0].cuda()
    grads = {}
    try:
        for mode in (1, 2, 0):
            _set_bwd_mode(mode)
            net.zero_grad()
            out = net(rays["ray_o"], rays["ray_d"], far, fg_z, bg_z)
            loss = torch.mean((out["rgb"] - rays["rgb"]) ** 2) + 0.1 * torch.mean((out["depth"] - rays["depth_sup"]) ** 2)
            loss.backward()
            torch.cuda.synchronize()
            grads[mode] = {k: p.grad.detach().clone() for k, p in net.named_parameters()}
    finally:
        _set_bwd_mode(2)
    for mode in (2, 0):
        for k in grads[mode]:
            a, b = grads[mode][k].double(), grads[1][k].double()
            assert torch.isfinite(a).all(), (mode, k)
            err = float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)
            assert err <= 5e-4, (mode, k, err)


def test_graphed_train_step():
    """nerfpp_b200.GraphedTrainStep: the trainer's whole step (both levels: forward, loss, backward, Adam) as one CUDA graph.
    Construction (warm-up passes + capture) must leave weights and optimizer state untouched; replays must train --
    the first replay's loss is that of an eager step on the same batch (different uniform draws move the depth term a lot:
    within a factor of 3), the losses fall over 30 replays, parameters move, and the weights the graph re-packs on every replay are the ones it updates."""
    import copy
    from test_parity_gpu import make_models
    import depth_loss as DL
    from nerfpp_b200 import GraphedTrainStep, ops
    levels = [O.densify(p, 5.0) for p in O.make_params_levels(2)]
    nets = make_models(levels)
    before = [{k: v.detach().clone() for k, v in n.state_dict().items()} for n in nets]
    n = 512
    rays = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in O.synthetic_rays(n, seed=9).items()}
    step = GraphedTrainStep(nets, n, (64, 128), depth_loss_type="mse", lambda_depth=0.1, depth_scale=rays["depth_scale"], batch=rays)
    for net, sd in zip(nets, before):
        for k, v in net.state_dict().items():
            assert torch.equal(v, sd[k]), k                    # construction did not train
    for opt in step.optimizers:
        assert all(float(v.abs().sum()) == 0.0 for st in opt.state.values() for v in st.values() if torch.is_tensor(v))
    # an eager step of level 0 on the same batch, for scale
    ref = copy.deepcopy(nets[0])
    far = ops.intersect_sphere(rays["ray_o"], rays["ray_d"])
    fg_z, bg_z = ops.coarse_depths(rays["min_depth"], far, 64, torch.rand(n, 64, device="cuda"), torch.rand(n, 64, device="cuda"))
    out = ref(rays["ray_o"], rays["ray_d"], far, fg_z, bg_z)
    eager0 = float(torch.mean((out["rgb"] - rays["rgb"]) ** 2) + 0.1 * DL.depth_mse(rays["depth_sup"], out["depth"]))
    first = step().clone()
    assert eager0 / 3.0 <= float(first[0]) <= 3.0 * eager0, (float(first[0]), eager0)
    for _ in range(30):
        last = step()
    step.check_unbounded()
    last = last.clone()
    assert torch.isfinite(last).all() and float(last[1]) < 0.9 * float(first[1]), (first, last)
    moved = nets[1].state_dict()["nerf_net.fg_net.base_layers.3.0.weight"]
    assert not torch.equal(moved, before[1]["nerf_net.fg_net.base_layers.3.0.weight"])
    # eager inference with the trained weights after invalidating the host-side cache keys == the graph's own view
    step.invalidate_inference_caches()
    with torch.no_grad():
        a = nets[1](rays["ray_o"], rays["ray_d"], far, fg_z, bg_z)["rgb"]
        nets[1].nerf_net.invalidate_packed()
        b = nets[1](rays["ray_o"], rays["ray_d"], far, fg_z, bg_z)["rgb"]
    assert torch.equal(a, b)


def test_heads_folded_into_wgrad_agree():
    """Job::head (wgrad_tc.cu): the sigma / rgb.2 head gradients taken by the base_remap / view-direction jobs' idle
    epilogue warps instead of wgrad_small_kernel -- off by default (measured slower), must give the same gradients."""
    import ctypes
    from nerfpp_b200 import _lib
    from test_parity_gpu import make_models
    L = _lib.lib()
    L.nerfpp_debug_set_heads_folded.argtypes = [ctypes.c_int]
    net = make_models([O.densify(O.make_params(), 5.0)])[0]
    n = 300
    rays = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in O.synthetic_rays(n, seed=5).items()}
    far = O.intersect_sphere(rays["ray_o"].cpu(), rays["ray_d"].cpu()).cuda()
    g = torch.Generator().manual_seed(2)
    fg_z = (torch.sort(torch.rand(n, 100, generator=g), -1)[0]).cuda() * far[:, None]
    bg_z = torch.sort(torch.rand(n, 70, generator=g), -1)[0].cuda()
    res = {}
    try:
        for folded in (0, 1):
            L.nerfpp_debug_set_heads_folded(folded)
            net.zero_grad()
            out = net(rays["ray_o"], rays["ray_d"], far, fg_z, bg_z)
            (torch.mean((out["rgb"] - rays["rgb"]) ** 2) + 0.1 * torch.mean((out["depth"] - rays["depth_sup"]) ** 2)).backward()
            torch.cuda.synchronize()
            res[folded] = {k: p.grad.detach().clone() for k, p in net.named_parameters()}
    finally:
        L.nerfpp_debug_set_heads_folded(0)
    for k in res[0]:
        a, b = res[1][k].double(), res[0][k].double()
        tol = 1e-5 if ("sigma_layers" in k or "rgb_layers.2" in k) else 0.0     # the heads: same products, another summation order
        assert float((a - b).abs().max()) <= tol * max(float(b.abs().max()), 1e-30), k


def test_backward_is_bit_reproducible():
    """The weight-gradient kernels write per-CTA partials that are summed in a fixed order (wgrad_reduce_kernel,
    heads_reduce_kernel) instead of red.global.add-ing into the gradient: two backward passes over the same batch give
    bit-identical gradients for all 48 tensors, so a training run is reproducible."""
    from test_parity_gpu import make_models
    net = make_models([O.densify(O.make_params(), 5.0)])[0]
    n = 1500
    rays = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in O.synthetic_rays(n, seed=77).items()}
    far = O.intersect_sphere(rays["ray_o"].cpu(), rays["ray_d"].cpu()).cuda()
    g = torch.Generator().manual_seed(1)
    fg_z = (torch.sort(torch.rand(n, 192, generator=g), -1)[0]).cuda() * far[:, None]
    bg_z = torch.sort(torch.rand(n, 192, generator=g), -1)[0].cuda()
    runs = []
    for _ in range(3):
        net.zero_grad()
        out = net(rays["ray_o"], rays["ray_d"], far, fg_z, bg_z)
        loss = torch.mean((out["rgb"] - rays["rgb"]) ** 2) + 0.1 * torch.mean((out["depth"] - rays["depth_sup"]) ** 2)
        loss.backward()
        torch.cuda.synchronize()
        runs.append({k: p.grad.detach().clone() for k, p in net.named_parameters()})
    for k in runs[0]:
        assert float(runs[0][k].abs().max()) > 0, k
        assert torch.equal(runs[0][k], runs[1][k]) and torch.equal(runs[0][k], runs[2][k]), k
