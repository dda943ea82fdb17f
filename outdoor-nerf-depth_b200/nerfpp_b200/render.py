"""The "render_rays" body: the cascade loop the reference inlines in its trainer
(ddp_train_nerf.py:432-498) and in render_single_image (:156-221), as one call."""
from collections import OrderedDict

import torch

from . import ops


def cascade_forward(models, ray_o, ray_d, min_depth, cascade_samples=(64, 128), train=False, rand=None, impl=None,
                    fg_far=None, unbounded_check=None):
    """Runs every cascade level. ``models[m]`` is a ddp_model.NerfNetWithAutoExpo (or anything with
    the same forward signature). ``train`` selects the stochastic path (perturbed coarse depths,
    random inverse-CDF draws, :440-465) vs the deterministic test path (:166-196).  ``rand`` may
    carry explicit draws {'t_fg','t_bg','u_fg_1','u_bg_1',...} (tests).
    Returns ([(ret, fg_z, bg_z) per level], fg_far)."""
    flag = None
    if fg_far is None:
        # the reference raises inside intersect_sphere (a host sync before anything else is enqueued); here the whole
        # cascade is enqueued first and the flag is read once at the end: same exception, no pipeline bubble
        fg_far, flag = ops.intersect_sphere(ray_o, ray_d, deferred=True)
    out = []
    fg_z = bg_z = ret = None
    n = ray_o.shape[0]
    draws = None
    if train and not rand:
        # every uniform draw of the step from ONE torch.rand call (one launch instead of one per level): per level a
        # contiguous [2, n, S] block, row 0 foreground, row 1 background
        flat = torch.rand(2 * n * sum(cascade_samples), device=ray_o.device)
        draws, off = [], 0
        for S in cascade_samples:
            draws.append(flat[off:off + 2 * n * S].view(2, n, S))
            off += 2 * n * S
    for m, S in enumerate(cascade_samples):
        if m == 0:
            t_fg = t_bg = None
            if train:
                t_fg, t_bg = (rand["t_fg"], rand["t_bg"]) if rand else draws[0].unbind(0)
            fg_z, bg_z = ops.coarse_depths(min_depth, fg_far, S, t_fg, t_bg)
        else:
            u_fg = rand["u_fg_%d" % m] if (train and rand) else (draws[m][0] if draws is not None else None)
            u_bg = rand["u_bg_%d" % m] if (train and rand) else (draws[m][1] if draws is not None else None)
            if fg_z.shape == bg_z.shape:
                fg_z, bg_z = ops.resample_merge_pair(fg_z, ret["fg_weights"], bg_z, ret["bg_weights"], S, det=not train,
                                                     u_fg=u_fg, u_bg=u_bg)
            else:
                fg_z = ops.resample_merge(fg_z, ret["fg_weights"], S, det=not train, u=u_fg)
                bg_z = ops.resample_merge(bg_z, ret["bg_weights"], S, det=not train, u=u_bg)
        ret = models[m](ray_o, ray_d, fg_far, fg_z, bg_z) if impl is None else models[m](ray_o, ray_d, fg_far, fg_z, bg_z, impl=impl)
        out.append((ret, fg_z, bg_z))
    if flag is not None:
        if unbounded_check is None:
            flag.raise_if_set()
        else:
            unbounded_check.append(flag)      # the caller reads it when convenient (e.g. one step later)
    return out, fg_far


def render_rays(models, ray_batch, cascade_samples=(64, 128), train=False, depth_loss_type=None, lambda_depth=0.0,
                depth_sigma=0.01, rand=None, impl=None, defer_unbounded_check=False):
    """One pass of the whole hot path over a ray batch (a dict like RaySamplerSingleImage yields:
    ray_o, ray_d, min_depth, and optionally rgb, depth_sup, depth_scale).  Returns an OrderedDict
    with the finest level's outputs, the per-level depths, and -- when ``rgb`` is present -- the
    loss vector [rgb_loss, depth_loss, total, n_valid] per level (ddp_train_nerf.py:481-493)."""
    pending = []
    out, fg_far = cascade_forward(models, ray_batch["ray_o"], ray_batch["ray_d"], ray_batch["min_depth"],
                                  cascade_samples, train, rand, impl, unbounded_check=pending)
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
    # ddp_train_nerf.py:62-63 raises as soon as a camera is outside the unit sphere; here everything has been enqueued
    # first.  With defer_unbounded_check the flag is handed to the caller (res["unbounded"].raise_if_set()).
    if defer_unbounded_check:
        res["unbounded"] = pending[0] if pending else None
    elif pending:
        pending[0].raise_if_set()
    return res


# ------------------------------------------------------------------------------------------------
# A15: render_single_image (ddp_train_nerf.py:133-249)
# ------------------------------------------------------------------------------------------------
# every tensor key of NerfNet.forward's dict except the two weight arrays (:210-211), in dict order
RENDER_KEYS = ("rgb", "fg_dists", "fg_rgb", "fg_depth", "bg_rgb", "bg_depth", "bg_lambda", "depth")
_KEY_WIDTH = dict(rgb=3, fg_rgb=3, bg_rgb=3, fg_depth=1, bg_depth=1, bg_lambda=1, depth=1)   # fg_dists: the level's sample count


class _LazyRenderDict(OrderedDict):
    """The per-level result of render_single_image: an OrderedDict with the reference's keys in the reference's order,
    whose bulky entries (``fg_dists``: one float per SAMPLE, 93 % of the bytes, read by nothing in the trainer or the
    tester) stay on the device until somebody asks for them.  Any access that could see the value -- ``d[k]``, ``get``,
    ``values()``, ``items()``, ``pop`` -- copies it to the host first, so it behaves like the plain dict of CPU tensors."""

    def __init__(self, eager, lazy):
        super().__init__(eager)
        self._lazy = dict(lazy)

    def _force(self, k=None):
        for key in ([k] if k is not None else list(self._lazy)):
            if key in self._lazy:
                super().__setitem__(key, self._lazy.pop(key)())

    def __getitem__(self, k):
        self._force(k)
        return super().__getitem__(k)

    def get(self, k, default=None):
        self._force(k)
        return super().get(k, default)

    def pop(self, k, *a):
        self._force(k)
        return super().pop(k, *a)

    def values(self):
        self._force()
        return super().values()

    def items(self):
        self._force()
        return super().items()

    def copy(self):
        self._force()
        return OrderedDict(super().items())


def band_sizes(n_rays, world_size):
    """ddp_train_nerf.py:137-143: contiguous row-major bands, pixel count must divide by the world size."""
    if (n_rays // world_size) * world_size != n_rays:
        raise Exception("Number of pixels in the image is not divisible by the number of GPUs!\n\t# pixels: {}\n\t# GPUs: {}"
                        .format(n_rays, world_size))
    sizes = [n_rays // world_size] * world_size
    sizes[-1] = n_rays - sum(sizes[:-1])
    return sizes


def _level_widths(models):
    """Packed channel layout of one ray: for every cascade level the RENDER_KEYS side by side."""
    layout, off = [], 0
    samples, tot = models["cascade_samples"], 0
    for m in range(models["cascade_level"]):
        tot += samples[m]                      # level m evaluates the union of all depths so far (:457)
        cols = OrderedDict()
        for k in RENDER_KEYS:
            w = tot if k == "fg_dists" else _KEY_WIDTH[k]
            cols[k] = (off, w)
            off += w
        layout.append(cols)
    return layout, off


def _render_chunk_cuda(models, chunk):
    nets = [models["net_%d" % m] for m in range(models["cascade_level"])]
    with torch.no_grad():
        out, _ = cascade_forward(nets, chunk["ray_o"], chunk["ray_d"], chunk["min_depth"], tuple(models["cascade_samples"]),
                                 train=False)
    return [ret for ret, _, _ in out]


def render_single_image(rank, world_size, models, ray_sampler, chunk_size, render_chunk=None, device=None):
    """Drop-in for ddp_train_nerf.render_single_image (:133-249): same arguments, same return -- on rank 0 a list
    (one entry per cascade level) of OrderedDicts of CPU tensors reshaped to (H, W, -1).squeeze(), None elsewhere.

    What changed underneath: each rank renders its band chunk by chunk straight into ONE packed device buffer
    [rays_of_rank, channels_of_all_levels] (no per-chunk device->host copies, no empty_cache()), and the ranks are
    merged by ONE all-gather of that buffer (NCCL over NVLink when the process group is NCCL) instead of one CPU
    gloo gather per key per level (:229-243).  Rank 0 copies the narrow keys (rgb, depths, lambda: 20 floats per ray
    and level) to the host at once; ``fg_dists`` (one float per sample -- 93 % of the bytes, read by nothing downstream)
    is fetched when first accessed (_LazyRenderDict).  ``render_chunk(models, chunk_dict) -> [ret per level]`` is the
    per-chunk renderer (default: the CUDA cascade); tests inject a stub to exercise the sharding logic on CPU."""
    import torch.distributed as dist
    ray_batch = ray_sampler.get_all()
    n = ray_batch["ray_d"].shape[0]
    sizes = band_sizes(n, world_size)
    if device is None:
        device = torch.device("cuda", rank) if torch.cuda.is_available() else torch.device("cpu")
    render_chunk = render_chunk or _render_chunk_cuda
    local = {k: torch.split(v, sizes)[rank].to(device) for k, v in ray_batch.items() if torch.is_tensor(v)}
    layout, channels = _level_widths(models)
    n_local = sizes[rank]
    packed = torch.empty(n_local, channels, device=device, dtype=torch.float32)
    for s0 in range(0, n_local, chunk_size):
        chunk = {k: v[s0:s0 + chunk_size] for k, v in local.items()}
        rets = render_chunk(models, chunk)
        for m, ret in enumerate(rets):
            for k, (off, w) in layout[m].items():
                packed[s0:s0 + chunk_size, off:off + w] = ret[k].reshape(-1, w)
    if world_size > 1:
        gathered = torch.empty(world_size * n_local, channels, device=device, dtype=torch.float32)
        dist.all_gather_into_tensor(gathered, packed)       # bands are equal-sized (divisibility check above)
    else:
        gathered = packed
    if rank != 0:
        return None
    # one D2H of the narrow keys (20 floats per ray and level); fg_dists stays on the device until it is read
    out = []
    H, W = ray_sampler.H, ray_sampler.W
    small = [(k, off, w) for cols in layout for k, (off, w) in cols.items() if k != "fg_dists"]
    host = torch.cat([gathered[:, off:off + w] for _, off, w in small], dim=1).cpu()
    pos = 0
    for cols in layout:
        eager, lazy = [], {}
        for k, (off, w) in cols.items():
            if k == "fg_dists":
                eager.append((k, None))
                lazy[k] = (lambda o=off, ww=w: gathered[:, o:o + ww].cpu().reshape(H, W, -1).squeeze())
            else:
                eager.append((k, host[:, pos:pos + w].reshape(H, W, -1).squeeze()))
                pos += w
        out.append(_LazyRenderDict(eager, lazy))
    return out
