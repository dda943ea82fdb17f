"""CPU oracle for the NeRF++ ray-marching hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A from-scratch restatement (torch CPU, fp32) of the algorithm of the reference
``cwchenwang/outdoor-nerf-depth`` path named in BASELINE.json: hierarchical sampling ->
positional-encoded 8x256 MLP -> transmittance composite -> rgb + depth-prior losses.  Every function
cites the reference ``file:line`` (relative to ``nerf-methods/nerfplusplus/``) it follows.

Who may import this: ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs -- only as the checker or the timed CPU baseline.  The product package
(``outdoor-nerf-depth_b200/``) never imports it and has no CPU fallback.

Pinning: the reference ships NO tests or golden vectors for this path,
so the oracle is pinned against outputs of the unmodified reference code run in the build
container: ``oracle/gen_golden.py`` -> ``tests/golden/nerfpp_*.npz`` and
``tests/test_oracle_golden.py`` (bit-exact for sampling indices, <=2e-6 for floats).

The arithmetic is deliberately kept as the same sequence of ATen ops the reference issues (torch is
the third-party arithmetic of the reference, section 8(c)) so that CPU quirks recorded in SURVEY R7
(fp64-accumulating ``cumsum``/``cumprod``) are inherited, but the code is organised functionally:
parameters are a flat ``{state_dict_name: tensor}`` mapping, no nn.Module.
"""
from collections import OrderedDict

import torch

TINY_NUMBER = 1e-6   # utils.py:8
HUGE_NUMBER = 1e10   # utils.py:7


class UnboundedCameraError(Exception):
    """ddp_train_nerf.py:62-63 raises a bare Exception with this message."""


# ----------------------------------------------------------------------------------------------
# parameters (nerf_network.py:71-118, ddp_model.py:49-72, ddp_train_nerf.py:308,322)
# ----------------------------------------------------------------------------------------------
def layer_shapes(pos_dim, n_freq_pos=10, n_freq_view=4, depth=8, width=256, skips=(4,)):
    """[(state-dict suffix, out, in)] in construction order for one MLPNet (nerf_network.py:89-117)."""
    emb = pos_dim * (1 + 2 * n_freq_pos)          # nerf_network.py:30-34
    vemb = 3 * (1 + 2 * n_freq_view)
    shapes = []
    dim = emb
    for i in range(depth):
        shapes.append(("base_layers.%d.0" % i, width, dim))
        dim = width
        if i in skips and i != depth - 1:
            dim += emb
    shapes.append(("sigma_layers.0", 1, dim))
    shapes.append(("base_remap_layers.0", 256, dim))
    shapes.append(("rgb_layers.0", width // 2, 256 + vemb))
    shapes.append(("rgb_layers.2", 3, width // 2))
    return shapes


def make_params_levels(n_levels=2, seed=777, max_freq_log2=10, max_freq_log2_viewdirs=4, netdepth=8,
                       netwidth=256):
    """Random-init parameters exactly as ``create_nerf`` does: ONE ``torch.manual_seed(777)`` and
    then, per cascade level, construct fg_net followed by bg_net with default ``nn.Linear`` init
    (ddp_train_nerf.py:308,315-322; xavier init is commented out, nerf_network.py:98,102,108,118).
    Building ``nn.Linear`` objects in the same order consumes the same RNG stream, so the values
    equal the reference's (checked against the reference in tests/golden)."""
    torch.manual_seed(seed)
    levels = []
    for _ in range(n_levels):
        params = OrderedDict()
        for net, pos_dim in (("fg_net", 3), ("bg_net", 4)):
            for suffix, n_out, n_in in layer_shapes(pos_dim, max_freq_log2, max_freq_log2_viewdirs,
                                                    netdepth, netwidth):
                lin = torch.nn.Linear(n_in, n_out)
                params["nerf_net.%s.%s.weight" % (net, suffix)] = lin.weight.detach().clone()
                params["nerf_net.%s.%s.bias" % (net, suffix)] = lin.bias.detach().clone()
        levels.append(params)
    return levels


def make_params(seed=777, **kw):
    return make_params_levels(1, seed, **kw)[0]


def densify(params, sigma_bias=5.0):
    """Variant used by tests/bench: raise the density bias so that weights are not ~0 (an untrained
    net gives sigma ~ 0.1)."""
    out = OrderedDict((k, v.clone()) for k, v in params.items())
    for net in ("fg_net", "bg_net"):
        out["nerf_net.%s.sigma_layers.0.bias" % net] += sigma_bias
    return out


# ----------------------------------------------------------------------------------------------
# A1-A5 sampling (ddp_train_nerf.py:51-130, 432-465)
# ----------------------------------------------------------------------------------------------
def intersect_sphere(ray_o, ray_d):
    """ddp_train_nerf.py:51-66: depth at which the ray leaves the unit sphere."""
    d1 = -(ray_d * ray_o).sum(-1) / (ray_d * ray_d).sum(-1)
    p = ray_o + d1[..., None] * ray_d
    inv_len = 1.0 / torch.norm(ray_d, dim=-1)
    p_sq = (p * p).sum(-1)
    if bool((p_sq >= 1.0).any()):
        raise UnboundedCameraError("Not all your cameras are bounded by the unit sphere")
    return d1 + torch.sqrt(1.0 - p_sq) * inv_len


def coarse_fg_depths(near, far, n_samples):
    """ddp_train_nerf.py:441-443 / 168-171: ``near + i*step`` (i*step form, not a running add)."""
    step = (far - near) / (n_samples - 1)
    return torch.stack([near + i * step for i in range(n_samples)], dim=-1)


def coarse_bg_depths(n_rays, n_samples):
    """ddp_train_nerf.py:447-448 / 174-175: ``linspace(0,1,S)`` broadcast over rays."""
    return torch.linspace(0.0, 1.0, n_samples).view(1, n_samples).expand(n_rays, n_samples)


def perturb_samples(z_vals, t_rand):
    """ddp_train_nerf.py:69-78 with the ``rand_like`` draw (:75) passed in as ``t_rand``."""
    mids = 0.5 * (z_vals[..., 1:] + z_vals[..., :-1])
    upper = torch.cat([mids, z_vals[..., -1:]], -1)
    lower = torch.cat([z_vals[..., :1], mids], -1)
    return lower + (upper - lower) * t_rand


def weights_to_cdf(weights):
    """ddp_train_nerf.py:90-93: +1e-6, normalise, cumsum, prepend 0 -> [..., M+1]."""
    w = weights + TINY_NUMBER
    pdf = w / torch.sum(w, dim=-1, keepdim=True)
    cdf = torch.cumsum(pdf, dim=-1)
    return torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)


def det_u(n_rays, n_samples):
    """ddp_train_nerf.py:102-104."""
    return torch.linspace(0.0, 1.0, n_samples).view(1, n_samples).expand(n_rays, n_samples)


def cdf_indices(cdf, u):
    """ddp_train_nerf.py:111: above = #{j < M : u >= cdf[j]}  (== searchsorted(cdf[:M], u, right))."""
    M = cdf.shape[-1] - 1
    return (u[..., :, None] >= cdf[..., None, :M]).sum(-1).long()


def sample_pdf(bins, weights, u, return_aux=False):
    """ddp_train_nerf.py:81-130 with ``u`` (:102-107) supplied by the caller.
    bins [..., M+1], weights [..., M], u [..., Ns] -> samples [..., Ns]."""
    cdf = weights_to_cdf(weights)
    above = cdf_indices(cdf, u)
    below = torch.clamp(above - 1, min=0)                                    # :114
    cdf_lo, cdf_hi = torch.gather(cdf, -1, below), torch.gather(cdf, -1, above)   # :117-118
    bin_lo, bin_hi = torch.gather(bins, -1, below), torch.gather(bins, -1, above)  # :120-121
    denom = cdf_hi - cdf_lo
    denom = torch.where(denom < TINY_NUMBER, torch.ones_like(denom), denom)  # :124-125
    t = (u - cdf_lo) / denom
    samples = bin_lo + t * (bin_hi - bin_lo + TINY_NUMBER)                   # :128
    if return_aux:
        return samples, cdf, above
    return samples


def resample_level(z_prev, weights_prev, u):
    """ddp_train_nerf.py:452-457 (fg) / 460-465 (bg): mid-point bins, interior weights, inverse-CDF
    draw, then the sorted union of old and new depths (values only)."""
    mids = 0.5 * (z_prev[..., 1:] + z_prev[..., :-1])
    new = sample_pdf(mids, weights_prev[..., 1:-1], u)
    merged, _ = torch.sort(torch.cat((z_prev, new), -1))
    return merged


# ----------------------------------------------------------------------------------------------
# A6-A8 field
# ----------------------------------------------------------------------------------------------
def posenc(x, n_freqs):
    """nerf_network.py:42-60: [x, sin(2^0 x), cos(2^0 x), sin(2^1 x), ...], freq 2^k (:36)."""
    out = [x]
    for k in range(n_freqs):
        f = float(2.0 ** k)
        out.append(torch.sin(x * f))
        out.append(torch.cos(x * f))
    return torch.cat(out, -1)


def mlp_field(params, net, pos_emb, view_emb, depth=8, skips=(4,)):
    """nerf_network.py:120-142.  Returns (rgb [...,3], sigma [...])."""
    def lin(name, x):
        p = "nerf_net.%s.%s" % (net, name)
        return torch.nn.functional.linear(x, params[p + ".weight"], params[p + ".bias"])

    h = torch.relu(lin("base_layers.0.0", pos_emb))
    for i in range(depth - 1):
        if i in skips:
            h = torch.cat((pos_emb, h), -1)                       # :129-130
        h = torch.relu(lin("base_layers.%d.0" % (i + 1), h))
    sigma = torch.abs(lin("sigma_layers.0", h)).squeeze(-1)        # :133-134
    remap = lin("base_remap_layers.0", h)                          # :136 (no activation)
    c = torch.relu(lin("rgb_layers.0", torch.cat((remap, view_emb), -1)))
    rgb = torch.sigmoid(lin("rgb_layers.2", c))                    # :110-117
    return rgb, sigma


def mlp_field_acts(params, net, pos_emb, view_emb, depth=8, skips=(4,)):
    """mlp_field with its intermediates (for the backward tests): ([h0..h7, remap, rgb_hidden], raw_sigma, rgb)."""
    def lin(name, x):
        p = "nerf_net.%s.%s" % (net, name)
        return torch.nn.functional.linear(x, params[p + ".weight"], params[p + ".bias"])

    acts = []
    h = torch.relu(lin("base_layers.0.0", pos_emb))
    acts.append(h)
    for i in range(depth - 1):
        x = torch.cat((pos_emb, h), -1) if i in skips else h
        h = torch.relu(lin("base_layers.%d.0" % (i + 1), x))
        acts.append(h)
    raw_sigma = lin("sigma_layers.0", h).squeeze(-1)
    remap = lin("base_remap_layers.0", h)
    acts.append(remap)
    c = torch.relu(lin("rgb_layers.0", torch.cat((remap, view_emb), -1)))
    acts.append(c)
    return acts, raw_sigma, torch.sigmoid(lin("rgb_layers.2", c))


def inverted_sphere_points(ray_o, ray_d, depth):
    """ddp_model.py:16-45 (depth2pts_outside).  ray_o/ray_d [..., 3] (already expanded per sample),
    depth [...] in [0,1] = 1/r.  Returns pts [...,4] = (x', y', z', 1/r) and real depth [...]."""
    d1 = -(ray_d * ray_o).sum(-1) / (ray_d * ray_d).sum(-1)
    p_mid = ray_o + d1[..., None] * ray_d
    p_mid_norm = torch.norm(p_mid, dim=-1)
    inv_len = 1.0 / torch.norm(ray_d, dim=-1)
    d2 = torch.sqrt(1.0 - p_mid_norm * p_mid_norm) * inv_len
    p_sph = ray_o + (d1 + d2)[..., None] * ray_d
    axis = torch.cross(ray_o, p_sph, dim=-1)
    axis = axis / torch.norm(axis, dim=-1, keepdim=True)
    phi = torch.asin(p_mid_norm)
    theta = torch.asin(p_mid_norm * depth)
    ang = (phi - theta)[..., None]
    rot = (p_sph * torch.cos(ang)
           + torch.cross(axis, p_sph, dim=-1) * torch.sin(ang)
           + axis * (axis * p_sph).sum(-1, keepdim=True) * (1.0 - torch.cos(ang)))
    rot = rot / torch.norm(rot, dim=-1, keepdim=True)
    pts = torch.cat((rot, depth[..., None]), -1)
    depth_real = 1.0 / (depth + TINY_NUMBER) * torch.cos(theta) * inv_len + d1
    return pts, depth_real


# ----------------------------------------------------------------------------------------------
# A9-A11 composite
# ----------------------------------------------------------------------------------------------
def composite(fg_sigma, fg_rgb_raw, bg_sigma, bg_rgb_raw, bg_depth_real_flipped, ray_d, fg_z_max, fg_z, bg_z):
    """The "raw2outputs" inlined in ddp_model.py:95-105 (fg), :118-128 (bg, flipped sample order) and :131-134 (merge).
    Per-sample inputs of the background are in the flipped order the reference evaluates them in."""
    d_norm = torch.norm(ray_d, dim=-1, keepdim=True)
    dz = fg_z[..., 1:] - fg_z[..., :-1]
    fg_dists = d_norm * torch.cat((dz, fg_z_max[:, None] - fg_z[..., -1:]), -1)
    alpha = 1.0 - torch.exp(-fg_sigma * fg_dists)
    T = torch.cumprod(1.0 - alpha + TINY_NUMBER, dim=-1)
    bg_lambda = T[..., -1]
    T = torch.cat((torch.ones_like(T[..., :1]), T[..., :-1]), -1)
    fg_weights = alpha * T
    fg_rgb = (fg_weights[..., None] * fg_rgb_raw).sum(-2)
    fg_depth = (fg_weights * fg_z).sum(-1)
    zf = torch.flip(bg_z, dims=[-1])                                # :117  1 -> 0
    bg_dists = torch.cat((zf[..., :-1] - zf[..., 1:],
                          HUGE_NUMBER * torch.ones_like(zf[..., :1])), -1)   # :118-119
    alpha_b = 1.0 - torch.exp(-bg_sigma * bg_dists)
    Tb = torch.cumprod(1.0 - alpha_b + TINY_NUMBER, dim=-1)[..., :-1]
    Tb = torch.cat((torch.ones_like(Tb[..., :1]), Tb), -1)
    bg_weights = alpha_b * Tb
    bg_rgb = (bg_weights[..., None] * bg_rgb_raw).sum(-2)
    bg_depth = (bg_weights * bg_depth_real_flipped).sum(-1)
    bg_rgb = bg_lambda[..., None] * bg_rgb                           # :131-134
    bg_depth = bg_lambda * bg_depth
    return OrderedDict([("rgb", fg_rgb + bg_rgb), ("fg_weights", fg_weights),
                        ("bg_weights", bg_weights), ("fg_dists", fg_dists), ("fg_rgb", fg_rgb),
                        ("fg_depth", fg_depth), ("bg_rgb", bg_rgb), ("bg_depth", bg_depth),
                        ("bg_lambda", bg_lambda), ("depth", fg_depth + bg_depth)])


def nerfpp_forward(params, ray_o, ray_d, fg_z_max, fg_z, bg_z, n_freq_pos=10, n_freq_view=4,
                   return_raw=False):
    """ddp_model.py:74-147 (NerfNet.forward).  Returns the 10-key OrderedDict in reference order."""
    d_norm = torch.norm(ray_d, dim=-1, keepdim=True)
    viewdir = ray_d / d_norm
    N = ray_d.shape[0]

    # foreground, ddp_model.py:86-105
    S = fg_z.shape[-1]
    o = ray_o[:, None, :].expand(N, S, 3)
    d = ray_d[:, None, :].expand(N, S, 3)
    v = viewdir[:, None, :].expand(N, S, 3)
    pts = o + fg_z[..., None] * d
    fg_rgb_raw, fg_sigma = mlp_field(params, "fg_net", posenc(pts, n_freq_pos), posenc(v, n_freq_view))
    # background, ddp_model.py:107-128
    S = bg_z.shape[-1]
    o = ray_o[:, None, :].expand(N, S, 3)
    d = ray_d[:, None, :].expand(N, S, 3)
    v = viewdir[:, None, :].expand(N, S, 3)
    bg_pts, bg_depth_real = inverted_sphere_points(o, d, bg_z)
    pe = torch.flip(posenc(bg_pts, n_freq_pos), dims=[-2])          # :116
    ve = torch.flip(posenc(v, n_freq_view), dims=[-2])
    bg_rgb_raw, bg_sigma = mlp_field(params, "bg_net", pe, ve)
    ret = composite(fg_sigma, fg_rgb_raw, bg_sigma, bg_rgb_raw, torch.flip(bg_depth_real, dims=[-1]), ray_d, fg_z_max, fg_z, bg_z)
    if return_raw:
        ret["_fg_sigma"], ret["_fg_rgb_raw"] = fg_sigma, fg_rgb_raw
        ret["_bg_sigma"], ret["_bg_rgb_raw"] = bg_sigma, bg_rgb_raw     # flipped order
    return ret


# ----------------------------------------------------------------------------------------------
# A12-A14 losses
# ----------------------------------------------------------------------------------------------
def img2mse(x, y):
    """utils.py:12-14 (mask=None branch)."""
    return torch.mean((x - y) * (x - y))


def mse2psnr(x):
    """utils.py:31."""
    import math
    return -10.0 * math.log(x + TINY_NUMBER) / math.log(10.0)


def depth_mse(depth_gt, depth_pred):
    """depth_loss.py:4-10: mean squared error over rays with gt > 0 (NaN if none)."""
    m = depth_gt > 0.0
    return torch.mean((depth_gt[m] - depth_pred[m]) ** 2)


def depth_l1(depth_gt, depth_pred):
    """depth_loss.py:12-18."""
    m = depth_gt > 0.0
    return torch.mean(torch.abs(depth_gt[m] - depth_pred[m]))


def depth_kl(weights, termination_depth, steps, lengths, sigma, fg_far_depth=None):
    """depth_loss.py:20-44: note /(2*sigma) (not sigma^2) and the mean over SAMPLES of a sum over
    valid RAYS (total / S)."""
    m = termination_depth > 0
    if fg_far_depth is not None:
        m = torch.logical_and(m, termination_depth < fg_far_depth)
    per = (-torch.log(weights + 1e-5)
           * torch.exp(-((steps - termination_depth[:, None]) ** 2) / (2 * sigma)) * lengths)
    return torch.mean(per[m].sum(-2))


# ----------------------------------------------------------------------------------------------
# the "render_rays" body: the cascade loop of ddp_train_nerf.py:432-498 / 156-221
# ----------------------------------------------------------------------------------------------
def cascade_forward(params_levels, ray_o, ray_d, min_depth, cascade_samples=(64, 128),
                    rand=None):
    """Runs every cascade level's forward.  ``params_levels[m]`` = parameters of level m's net
    (independent nets, ddp_train_nerf.py:313-323).  ``rand`` = None for the deterministic test-time
    path (:166-196, det=True, no perturbation) or a dict with 't_fg','t_bg' [N,S0] and per level
    m>0 'u_fg_m','u_bg_m' [N,S_m] for the training path (:440-465).
    Returns list of (ret, fg_z, bg_z) per level and fg_far."""
    N = ray_o.shape[0]
    out = []
    fg_far = intersect_sphere(ray_o, ray_d)
    fg_z = bg_z = ret = None
    for m, S in enumerate(cascade_samples):
        if m == 0:
            fg_z = coarse_fg_depths(min_depth, fg_far, S)
            bg_z = coarse_bg_depths(N, S)
            if rand is not None:
                fg_z = perturb_samples(fg_z, rand["t_fg"])
                bg_z = perturb_samples(bg_z, rand["t_bg"])
        else:
            u_fg = det_u(N, S) if rand is None else rand["u_fg_%d" % m]
            u_bg = det_u(N, S) if rand is None else rand["u_bg_%d" % m]
            fg_z = resample_level(fg_z, ret["fg_weights"].detach(), u_fg)
            bg_z = resample_level(bg_z, ret["bg_weights"].detach(), u_bg)
        ret = nerfpp_forward(params_levels[m], ray_o, ray_d, fg_far, fg_z, bg_z)
        out.append((ret, fg_z, bg_z))
    return out, fg_far


def level_loss(ret, rgb_gt, depth_sup, fg_z, fg_far, use_depth, depth_loss_type, lambda_depth,
               depth_sigma_scaled):
    """ddp_train_nerf.py:481-493.  Returns (loss, rgb_loss_value_before_inplace_add, depth_loss)."""
    rgb_loss = img2mse(ret["rgb"], rgb_gt)
    rgb_only = rgb_loss.detach().clone()
    loss = rgb_loss
    dl = None
    if use_depth:
        if depth_loss_type == "kl":
            dl = depth_kl(ret["fg_weights"], depth_sup, fg_z, ret["fg_dists"], depth_sigma_scaled, fg_far)
        elif depth_loss_type == "mse":
            dl = depth_mse(depth_sup, ret["depth"])
        elif depth_loss_type == "l1":
            dl = depth_l1(depth_sup, ret["depth"])
        else:
            raise ValueError(depth_loss_type)
        loss = loss + lambda_depth * dl
    return loss, rgb_only, dl


# ----------------------------------------------------------------------------------------------
# synthetic workload; shared by tests and bench so both sides see the same rays
# ----------------------------------------------------------------------------------------------
def synthetic_rays(n_rays, seed=0, valid_every=5, depth_scale=0.05):
    g = torch.Generator().manual_seed(seed)
    dirs = torch.randn(n_rays, 3, generator=g)
    dirs = dirs / dirs.norm(dim=-1, keepdim=True)
    ray_o = dirs * (0.5 * torch.rand(n_rays, 1, generator=g) ** (1.0 / 3.0))
    d = torch.randn(n_rays, 3, generator=g)
    ray_d = d / d.norm(dim=-1, keepdim=True) * (1.0 + 0.2 * torch.rand(n_rays, 1, generator=g))
    min_depth = torch.full((n_rays,), 1e-4)
    rgb_gt = torch.rand(n_rays, 3, generator=g)
    far = intersect_sphere(ray_o, ray_d)
    depth_sup = far * torch.rand(n_rays, generator=g)
    if valid_every:
        depth_sup[::valid_every] = 0.0
    return dict(ray_o=ray_o.contiguous(), ray_d=ray_d.contiguous(), min_depth=min_depth,
                rgb=rgb_gt, depth_sup=depth_sup, depth_scale=depth_scale)


def synthetic_rand(n_rays, cascade_samples=(64, 128), seed=1):
    g = torch.Generator().manual_seed(seed)
    r = {"t_fg": torch.rand(n_rays, cascade_samples[0], generator=g),
         "t_bg": torch.rand(n_rays, cascade_samples[0], generator=g)}
    for m in range(1, len(cascade_samples)):
        r["u_fg_%d" % m] = torch.rand(n_rays, cascade_samples[m], generator=g)
        r["u_bg_%d" % m] = torch.rand(n_rays, cascade_samples[m], generator=g)
    return r


# ----------------------------------------------------------------------------------------------
# N1: ray generation and batch sampling, nerf_sample_ray_split.py
# ----------------------------------------------------------------------------------------------
def get_rays_single_image(H, W, intrinsics, c2w):
    """nerf_sample_ray_split.py:10-34 (numpy, float32 inputs): pixel centres -> K^-1 -> camera-to-world rotation.
    Returns rays_o [H*W,3], rays_d [H*W,3] (NOT normalised), depth [H*W] (= inv(c2w)[2,3])."""
    import numpy as np
    u, v = np.meshgrid(np.arange(W), np.arange(H))
    u = u.reshape(-1).astype(dtype=np.float32) + 0.5
    v = v.reshape(-1).astype(dtype=np.float32) + 0.5
    pixels = np.stack((u, v, np.ones_like(u)), axis=0)
    rays_d = np.dot(np.linalg.inv(intrinsics[:3, :3]), pixels)
    rays_d = np.dot(c2w[:3, :3], rays_d).transpose((1, 0))
    rays_o = np.tile(c2w[:3, 3].reshape((1, 3)), (rays_d.shape[0], 1))
    depth = np.linalg.inv(c2w)[2, 3] * np.ones((rays_o.shape[0],), dtype=np.float32)
    return rays_o, rays_d, depth


def sample_ray_batch(H, W, intrinsics, c2w, select_inds, img=None, depth_sup=None, min_depth=None):
    """RaySamplerSingleImage.random_sample (nerf_sample_ray_split.py:155-221) for given pixel indices: a gather of the
    per-image arrays; min_depth defaults to 1e-4 (:194-197)."""
    import numpy as np
    ro, rd, dp = get_rays_single_image(H, W, intrinsics, c2w)
    ret = OrderedDict(ray_o=ro[select_inds], ray_d=rd[select_inds], depth=dp[select_inds])
    ret["rgb"] = img.reshape(-1, 3)[select_inds] if img is not None else None
    ret["min_depth"] = min_depth.reshape(-1)[select_inds] if min_depth is not None else 1e-4 * np.ones_like(ret["ray_d"][..., 0])
    if depth_sup is not None:
        ret["depth_sup"] = depth_sup.reshape(-1)[select_inds]
    return ret


def image_metrics(im, gt_im, depth_pred, depth_gt, depth_scale, cap=80):
    """N3: the test loop's per-image metrics, ddp_train_nerf.py:556-600 (numpy, as the reference computes them)."""
    import numpy as np
    r = {}
    mse = np.mean((gt_im - im) * (gt_im - im))
    r["mse"], r["psnr"] = float(mse), float(-10. * np.log(mse + TINY_NUMBER) / np.log(10.))
    gt_sparse = depth_gt / depth_scale
    pred = depth_pred / depth_scale
    valid = (gt_sparse < cap) & (gt_sparse > 1e-3)
    vg = gt_sparse[valid].clip(1e-3, cap)
    vp = pred[valid].clip(1e-3, cap)
    r["n_valid"] = float(valid.sum())
    r["rmse"] = float(np.sqrt(np.mean((vg - vp) ** 2)))
    r["rmse_log"] = float(np.sqrt(np.mean((np.log(vg) - np.log(vp)) ** 2)))
    r["abs_diff"] = float(np.mean(np.abs(vg - vp)))
    r["abs_rel"] = float(np.mean(np.abs(vg - vp) / vg))
    r["sq_rel"] = float(np.mean(((vg - vp) ** 2) / vg))
    return r
