"""Training harness for the convergence / PSNR-parity tests  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Runs the reference trainer's optimisation loop (ddp_train_nerf.py:417-503: pick an image, draw N_rand pixels, per cascade
level sample -> NerfNet.forward -> rgb MSE + lambda * depth loss -> backward -> Adam) in two arms that see IDENTICAL
initial weights, pixels and uniform draws:

  * ``train_oracle``: the fp32 oracle (oracle/nerfpp_oracle.py, torch autograd, any device; TF32 off), and
  * ``train_ours``:   the product (drop-in ``ddp_model`` / ``depth_loss`` modules over libnerfpp_b200.so, CUDA only),

and renders held-out views with each arm's weights (``render_oracle`` / ``render_ours``) so their PSNR / depth RMSE can be
compared (ddp_train_nerf.py:556-600).  The random draws come from numpy / a CPU torch.Generator keyed on (seed, step) and
are moved to the device, so neither arm depends on a device RNG stream.
"""
import os
from collections import OrderedDict
from types import SimpleNamespace

import numpy as np
import torch

import data_loader_oracle as DLO
import nerfpp_oracle as O

CASCADE = (64, 128)
NET_ARGS = SimpleNamespace(max_freq_log2=10, max_freq_log2_viewdirs=4, netdepth=8, netwidth=256, use_viewdirs=True)


def load_views(base, scene, split, depth_sup_type="gt"):
    """Per-image dicts of flat float32 arrays (ray_o, ray_d, rgb, depth_sup, depth_gt, min_depth) + H, W, depth_scale."""
    return DLO.load_data_split(base, scene, split, depth_sup_type=depth_sup_type)


def step_draws(seed, step, n_views, n_pixels, n_rand, cascade=CASCADE):
    """Everything random in one optimisation step: the image (ddp_train_nerf.py:423), its pixels without replacement
    (nerf_sample_ray_split.py:178) and the uniform draws of perturb_samples / sample_pdf (:75, :107)."""
    rs = np.random.RandomState((seed * 1000003 + step) % (2 ** 31 - 1))
    img = int(rs.randint(0, n_views))
    sel = rs.choice(n_pixels, size=(n_rand,), replace=False)
    g = torch.Generator().manual_seed(seed * 7919 + step)
    rand = {"t_fg": torch.rand(n_rand, cascade[0], generator=g), "t_bg": torch.rand(n_rand, cascade[0], generator=g)}
    for m in range(1, len(cascade)):
        rand["u_fg_%d" % m] = torch.rand(n_rand, cascade[m], generator=g)
        rand["u_bg_%d" % m] = torch.rand(n_rand, cascade[m], generator=g)
    return img, sel, rand


def batch_of(view, sel, device):
    """RaySamplerSingleImage.random_sample (nerf_sample_ray_split.py:155-221) for the pixels ``sel``: rows of the view's flat
    per-pixel arrays, as float32 tensors on ``device``."""
    return {k: torch.from_numpy(np.ascontiguousarray(view[k][sel], np.float32)).to(device)
            for k in ("ray_o", "ray_d", "rgb", "min_depth", "depth_sup") if view.get(k) is not None}


def oracle_cascade(levels, b, rand, device, detach_prev=True):
    """O.cascade_forward with every constant created on ``device`` (the oracle's own helper builds its linspace on the CPU)."""
    n = b["ray_o"].shape[0]
    fg_far = O.intersect_sphere(b["ray_o"], b["ray_d"])
    out = []
    fg_z = bg_z = ret = None
    for m, S in enumerate(CASCADE):
        if m == 0:
            fg_z = O.coarse_fg_depths(b["min_depth"], fg_far, S)
            bg_z = torch.linspace(0.0, 1.0, S).to(device).view(1, S).expand(n, S)
            if rand is not None:
                fg_z = O.perturb_samples(fg_z, rand["t_fg"])
                bg_z = O.perturb_samples(bg_z, rand["t_bg"])
        else:
            det = torch.linspace(0.0, 1.0, S).to(device).view(1, S).expand(n, S)
            u_fg = det if rand is None else rand["u_fg_%d" % m]
            u_bg = det if rand is None else rand["u_bg_%d" % m]
            fg_z = O.resample_level(fg_z, ret["fg_weights"].detach(), u_fg)
            bg_z = O.resample_level(bg_z, ret["bg_weights"].detach(), u_bg)
        ret = O.nerfpp_forward(levels[m], b["ray_o"], b["ray_d"], fg_far, fg_z, bg_z)
        out.append((ret, fg_z, bg_z))
    return out, fg_far


def train_oracle(levels, views, steps, device, n_rand=1024, seed=0, loss_type="mse", lambda_depth=0.1, depth_sigma=0.01,
                 lr=5e-4, log=None):
    """fp32 reference arm.  ``levels``: list of {state_dict_name: tensor}; trained in place (moved to ``device``)."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    levels = [OrderedDict((k, v.detach().clone().to(device).requires_grad_(True)) for k, v in p.items()) for p in levels]
    opts = [torch.optim.Adam(list(p.values()), lr=lr) for p in levels]
    npix = views[0]["H"] * views[0]["W"]
    scale = views[0]["depth_scale"] or 1.0
    hist = []
    for step in range(steps):
        img, sel, rand = step_draws(seed, step, len(views), npix, n_rand)
        b = batch_of(views[img], sel, device)
        rand = {k: v.to(device) for k, v in rand.items()}
        fg_far = O.intersect_sphere(b["ray_o"], b["ray_d"])
        n = n_rand
        fg_z = bg_z = ret = None
        rec = []
        for m, S in enumerate(CASCADE):
            if m == 0:
                fg_z = O.perturb_samples(O.coarse_fg_depths(b["min_depth"], fg_far, S), rand["t_fg"])
                bg_z = O.perturb_samples(torch.linspace(0.0, 1.0, S).to(device).view(1, S).expand(n, S), rand["t_bg"])
            else:
                fg_z = O.resample_level(fg_z, ret["fg_weights"].detach(), rand["u_fg_%d" % m])
                bg_z = O.resample_level(bg_z, ret["bg_weights"].detach(), rand["u_bg_%d" % m])
            opts[m].zero_grad()
            ret = O.nerfpp_forward(levels[m], b["ray_o"], b["ray_d"], fg_far, fg_z, bg_z)
            loss, rgb_only, dl = O.level_loss(ret, b["rgb"], b.get("depth_sup"), fg_z, fg_far, "depth_sup" in b and loss_type is not None,
                                              loss_type, lambda_depth, depth_sigma * scale)
            loss.backward()
            opts[m].step()
            rec.append(float(loss.detach()))
        hist.append(rec)
        if log and (step % log == 0 or step == steps - 1):
            print("oracle step %d loss %s" % (step, rec), flush=True)
    return [OrderedDict((k, v.detach()) for k, v in p.items()) for p in levels], hist


def make_ours(levels, device):
    import ddp_model
    nets = []
    for p in levels:
        net = ddp_model.NerfNetWithAutoExpo(NET_ARGS)
        net.load_state_dict(p)
        nets.append(net.to(device))
    return nets


def train_ours(levels, views, steps, device, n_rand=1024, seed=0, loss_type="mse", lambda_depth=0.1, depth_sigma=0.01, lr=5e-4,
               log=None):
    """Product arm: the same loop through the drop-in modules (CUDA kernels forward and backward) + torch Adam."""
    import depth_loss as DL
    from nerfpp_b200 import ops
    nets = make_ours(levels, device)
    opts = [torch.optim.Adam(net.parameters(), lr=lr) for net in nets]
    npix = views[0]["H"] * views[0]["W"]
    scale = views[0]["depth_scale"] or 1.0
    hist = []
    for step in range(steps):
        img, sel, rand = step_draws(seed, step, len(views), npix, n_rand)
        b = batch_of(views[img], sel, device)
        rand = {k: v.to(device) for k, v in rand.items()}
        far = ops.intersect_sphere(b["ray_o"], b["ray_d"])
        fg_z = bg_z = ret = None
        rec = []
        for m, S in enumerate(CASCADE):
            if m == 0:
                fg_z, bg_z = ops.coarse_depths(b["min_depth"], far, S, rand["t_fg"], rand["t_bg"])
            else:
                fg_z, bg_z = ops.resample_merge_pair(fg_z, ret["fg_weights"].detach(), bg_z, ret["bg_weights"].detach(), S,
                                                     u_fg=rand["u_fg_%d" % m], u_bg=rand["u_bg_%d" % m])
            opts[m].zero_grad()
            ret = nets[m](b["ray_o"], b["ray_d"], far, fg_z, bg_z)
            loss = torch.mean((ret["rgb"] - b["rgb"]) * (ret["rgb"] - b["rgb"]))
            if "depth_sup" in b and loss_type is not None:
                if loss_type == "kl":
                    dl = DL.depth_kl(ret["fg_weights"], b["depth_sup"], fg_z, ret["fg_dists"], depth_sigma * scale, far)
                else:
                    dl = (DL.depth_mse if loss_type == "mse" else DL.depth_l1)(b["depth_sup"], ret["depth"])
                loss = loss + lambda_depth * dl
            loss.backward()
            opts[m].step()
            rec.append(float(loss.detach()))
        hist.append(rec)
        if log and (step % log == 0 or step == steps - 1):
            print("ours   step %d loss %s" % (step, rec), flush=True)
    return nets, hist


def view_batch(view, device):
    n = view["H"] * view["W"]
    return batch_of(view, np.arange(n), device)


def render_oracle(levels, view, device, chunk=1024):
    """Deterministic test-time cascade (ddp_train_nerf.py:156-221) with the fp32 oracle; finest level's rgb / depth."""
    b = view_batch(view, device)
    levels = [OrderedDict((k, v.to(device)) for k, v in p.items()) for p in levels]
    rgb, depth = [], []
    with torch.no_grad():
        for s in range(0, b["ray_o"].shape[0], chunk):
            cb = {k: v[s:s + chunk] for k, v in b.items()}
            out, _ = oracle_cascade(levels, cb, None, device)
            rgb.append(out[-1][0]["rgb"])
            depth.append(out[-1][0]["depth"])
    return torch.cat(rgb), torch.cat(depth)


def render_ours(nets, view, device, chunk=8192):
    from nerfpp_b200 import cascade_forward
    b = view_batch(view, device)
    rgb, depth = [], []
    with torch.no_grad():
        for s in range(0, b["ray_o"].shape[0], chunk):
            out, _ = cascade_forward(nets, b["ray_o"][s:s + chunk], b["ray_d"][s:s + chunk], b["min_depth"][s:s + chunk], CASCADE,
                                     train=False)
            rgb.append(out[-1][0]["rgb"])
            depth.append(out[-1][0]["depth"])
    return torch.cat(rgb), torch.cat(depth)


def psnr_rmse(rgb, depth, view):
    """The test loop's two headline metrics (ddp_train_nerf.py:556-588) in numpy, for either arm."""
    m = O.image_metrics(rgb.cpu().numpy().reshape(view["H"], view["W"], 3), view["rgb"].reshape(view["H"], view["W"], 3),
                        depth.cpu().numpy().reshape(view["H"], view["W"]), view["depth_gt"].reshape(view["H"], view["W"]),
                        view["depth_scale"])
    return m["psnr"], m["rmse"]


def per_ray_rel(a, b, floor):
    """|a - b| / max(|b|, floor) per ray (rows of a [n] or [n, c] tensor; a row's error is its worst channel).
    Returns dict(max, p99, p50)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    e = (a - b).abs() / b.abs().clamp_min(floor)
    if e.dim() > 1:
        e = e.reshape(e.shape[0], -1).max(dim=1).values
    q = torch.quantile(e, torch.tensor([0.5, 0.99], dtype=torch.float64))
    return {"max": float(e.max()), "p99": float(q[1]), "p50": float(q[0])}
