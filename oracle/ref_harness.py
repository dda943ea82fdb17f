"""Drives the UNMODIFIED reference functions through the trainer's cascade loop  --  TEST INFRASTRUCTURE.

``reference_step`` wires ``intersect_sphere`` / ``perturb_samples`` / ``sample_pdf`` (ddp_train_nerf.py:51-130),
``NerfNetWithAutoExpo.forward`` (ddp_model.py:175), ``img2mse`` (utils.py:12) and ``depth_mse|l1|kl`` (depth_loss.py:4-44)
exactly as ddp_train_nerf.py:432-493 does, on CPU tensors.  It is what ``bench.py --impl reference`` and the
``cpu_baseline`` leg time when the reference files are present (``cpu_baseline.kind = "reference"``)."""
from types import SimpleNamespace

import torch

from _refload import available, load_reference  # noqa: F401

NET_ARGS = SimpleNamespace(max_freq_log2=10, max_freq_log2_viewdirs=4, netdepth=8, netwidth=256, use_viewdirs=True)


def build_nets(n_levels=2, sigma_bias=5.0):
    """create_nerf's construction (ddp_train_nerf.py:308,315-322) without DDP / the device move; density bias as in the
    synthetic workload."""
    _, M, _, _ = load_reference()
    torch.manual_seed(777)
    nets = [M.NerfNetWithAutoExpo(NET_ARGS) for _ in range(n_levels)]
    if sigma_bias:
        with torch.no_grad():
            for n in nets:
                n.nerf_net.fg_net.sigma_layers[0].bias += sigma_bias
                n.nerf_net.bg_net.sigma_layers[0].bias += sigma_bias
    return nets


def reference_step(nets, rays, cascade=(64, 128), depth_loss_type="mse", lambda_depth=0.1, depth_sigma=0.01, grad=False):
    """One pass of ddp_train_nerf.py:432-493 (training path: perturbed coarse depths, random inverse-CDF draws) over a ray
    batch dict (ray_o, ray_d, min_depth, rgb, depth_sup, depth_scale).  ``grad=False``: forward + losses under no_grad.
    Returns [(ret, loss)] per level."""
    T, _, D, U = load_reference()
    losses = {"mse": D.depth_mse, "l1": D.depth_l1}
    ray_o, ray_d = rays["ray_o"], rays["ray_d"]
    dots_sh = list(ray_d.shape[:-1])
    out = []
    ret = fg_depth = bg_depth = None
    ctx = torch.enable_grad() if grad else torch.no_grad()
    with ctx:
        for m, N_samples in enumerate(cascade):
            if m == 0:
                fg_far_depth = T.intersect_sphere(ray_o, ray_d)
                fg_near_depth = rays["min_depth"]
                step = (fg_far_depth - fg_near_depth) / (N_samples - 1)
                fg_depth = torch.stack([fg_near_depth + i * step for i in range(N_samples)], dim=-1)
                fg_depth = T.perturb_samples(fg_depth)
                bg_depth = torch.linspace(0., 1., N_samples).view([1, ] * len(dots_sh) + [N_samples, ]).expand(dots_sh + [N_samples, ])
                bg_depth = T.perturb_samples(bg_depth)
            else:
                fg_weights = ret['fg_weights'].clone().detach()[..., 1:-1]
                fg_mid = .5 * (fg_depth[..., 1:] + fg_depth[..., :-1])
                fg_new = T.sample_pdf(bins=fg_mid, weights=fg_weights, N_samples=N_samples, det=False)
                fg_depth, _ = torch.sort(torch.cat((fg_depth, fg_new), dim=-1))
                bg_weights = ret['bg_weights'].clone().detach()[..., 1:-1]
                bg_mid = .5 * (bg_depth[..., 1:] + bg_depth[..., :-1])
                bg_new = T.sample_pdf(bins=bg_mid, weights=bg_weights, N_samples=N_samples, det=False)
                bg_depth, _ = torch.sort(torch.cat((bg_depth, bg_new), dim=-1))
            ret = nets[m](ray_o, ray_d, fg_far_depth, fg_depth, bg_depth, img_name=None)
            loss = U.img2mse(ret['rgb'], rays["rgb"])
            if depth_loss_type == "kl":
                loss = loss + lambda_depth * D.depth_kl(ret['fg_weights'], rays["depth_sup"], fg_depth, ret['fg_dists'],
                                                        depth_sigma * rays["depth_scale"], fg_far_depth)
            elif depth_loss_type in losses:
                loss = loss + lambda_depth * losses[depth_loss_type](rays["depth_sup"], ret['depth'])
            out.append((ret, loss))
    return out
