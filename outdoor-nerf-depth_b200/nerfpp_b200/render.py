"""The "render_rays" body: the cascade loop the reference inlines in its trainer
(ddp_train_nerf.py:432-498) and in render_single_image (:156-221), as one call."""
from collections import OrderedDict

import torch

from . import ops


def cascade_forward(models, ray_o, ray_d, min_depth, cascade_samples=(64, 128), train=False, rand=None, impl=None,
                    fg_far=None):
    """Runs every cascade level. ``models[m]`` is a ddp_model.NerfNetWithAutoExpo (or anything with
    the same forward signature). ``train`` selects the stochastic path (perturbed coarse depths,
    random inverse-CDF draws, :440-465) vs the deterministic test path (:166-196).  ``rand`` may
    carry explicit draws {'t_fg','t_bg','u_fg_1','u_bg_1',...} (tests).
    Returns ([(ret, fg_z, bg_z) per level], fg_far)."""
    if fg_far is None:
        fg_far = ops.intersect_sphere(ray_o, ray_d)
    out = []
    fg_z = bg_z = ret = None
    n = ray_o.shape[0]
    for m, S in enumerate(cascade_samples):
        if m == 0:
            t_fg = t_bg = None
            if train:
                t_fg = rand["t_fg"] if rand else torch.rand(n, S, device=ray_o.device)
                t_bg = rand["t_bg"] if rand else torch.rand(n, S, device=ray_o.device)
            fg_z, bg_z = ops.coarse_depths(min_depth, fg_far, S, t_fg, t_bg)
        else:
            u_fg = rand["u_fg_%d" % m] if (train and rand) else None
            u_bg = rand["u_bg_%d" % m] if (train and rand) else None
            fg_z = ops.resample_merge(fg_z, ret["fg_weights"], S, det=not train, u=u_fg)
            bg_z = ops.resample_merge(bg_z, ret["bg_weights"], S, det=not train, u=u_bg)
        ret = models[m](ray_o, ray_d, fg_far, fg_z, bg_z) if impl is None else models[m](ray_o, ray_d, fg_far, fg_z, bg_z, impl=impl)
        out.append((ret, fg_z, bg_z))
    return out, fg_far


def render_rays(models, ray_batch, cascade_samples=(64, 128), train=False, depth_loss_type=None, lambda_depth=0.0,
                depth_sigma=0.01, rand=None, impl=None):
    """One pass of the whole hot path over a ray batch (a dict like RaySamplerSingleImage yields:
    ray_o, ray_d, min_depth, and optionally rgb, depth_sup, depth_scale).  Returns an OrderedDict
    with the finest level's outputs, the per-level depths, and -- when ``rgb`` is present -- the
    loss vector [rgb_loss, depth_loss, total, n_valid] per level (ddp_train_nerf.py:481-493)."""
    out, fg_far = cascade_forward(models, ray_batch["ray_o"], ray_batch["ray_d"], ray_batch["min_depth"],
                                  cascade_samples, train, rand, impl)
    res = OrderedDict(levels=out, fg_far=fg_far)
    if "rgb" in ray_batch:
        losses = []
        scale = ray_batch.get("depth_scale", 1.0)
        scale = float(scale) if not torch.is_tensor(scale) else float(scale.item())
        for ret, fg_z, _ in out:
            use_depth = depth_loss_type not in (None, "none") and "depth_sup" in ray_batch
            losses.append(ops.fused_loss(ret["rgb"], ray_batch["rgb"], ret["depth"], ray_batch.get("depth_sup"),
                                         depth_loss_type if use_depth else None, lambda_depth, ret["fg_weights"], fg_z,
                                         ret["fg_dists"], fg_far, depth_sigma * scale))
        res["losses"] = losses
    return res
