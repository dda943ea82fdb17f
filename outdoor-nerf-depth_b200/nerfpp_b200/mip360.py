"""Torch-facing wrappers for the mipnerf360 twins, named and shaped like the reference's JAX
functions in ``nerf-methods/mipnerf360/internal/`` (``stepfun.sample_intervals``, ``render.compute_alpha_weights``,
``render.volumetric_rendering``, ``depth_loss.depth_loss`` + the mse/l1 branch of ``train_utils.compute_data_loss``).
All arithmetic happens in libnerfpp_b200.so; CUDA tensors only."""
import math

import torch

from . import _lib
from ._lib import DEPTH_KL, DEPTH_L1, DEPTH_MSE
from .ops import _c, _p, _stream, check      # ops.check counts the library's kernel launches

EPS = float(torch.finfo(torch.float32).eps)


_CONST = {}       # (kind, num_samples, device) -> device tensor: built once on the host (no H2D copy inside a graph capture)


def centers_u(num_samples, device):
    """stepfun.py:193-199 (rng=None, deterministic_center=True), built in float64 then rounded like jnp.linspace."""
    key = ("centers", int(num_samples), str(device))
    if key not in _CONST:
        pad = 1 / (2 * num_samples)
        _CONST[key] = torch.linspace(pad, 1. - pad - EPS, num_samples, dtype=torch.float64).float().to(device)
    return _CONST[key]


def max_jitter(num_samples):
    """stepfun.py:203-206."""
    u_max = EPS + (1 - EPS) / num_samples
    return (1 - u_max) / (num_samples - 1) - EPS


def jitter_base(num_samples, device):
    """stepfun.py:207: linspace(0, 1 - u_max, Ns), built in float64 then rounded like jnp.linspace (cached per device)."""
    key = ("jitter", int(num_samples), str(device))
    if key not in _CONST:
        u_max = EPS + (1 - EPS) / num_samples
        _CONST[key] = torch.linspace(0, 1 - u_max, num_samples, dtype=torch.float64).float().to(device)
    return _CONST[key]


def jittered_u(shape_prefix, num_samples, single_jitter, device, generator=None):
    """stepfun.py:203-209: linspace(0, 1-u_max, Ns) + U[0, max_jitter).  The draw is torch's, not jax.random's."""
    u_max = EPS + (1 - EPS) / num_samples
    max_jitter = (1 - u_max) / (num_samples - 1) - EPS
    d = 1 if single_jitter else num_samples
    key = ("jitter", int(num_samples), str(device))
    if key not in _CONST:
        _CONST[key] = torch.linspace(0, 1 - u_max, num_samples, dtype=torch.float64).float().to(device)
    base = _CONST[key]
    return base + torch.rand(tuple(shape_prefix) + (d,), device=device, generator=generator) * max_jitter


def sample_intervals(u, t, w_logits, num_samples, single_jitter=False, domain=(-math.inf, math.inf)):
    """stepfun.sample_intervals (stepfun.py:214-263).  ``u``: None (the reference's rng=None path) or the
    inverse-CDF ordinates [..., num_samples] (see ``jittered_u``).  t [..., M+1], w_logits [..., M] -> [..., Ns+1]."""
    if num_samples <= 1:
        raise ValueError(f"num_samples must be > 1, is {num_samples}.")
    tt = _c(t, "t")
    lead = tuple(tt.shape[:-1])
    tt = tt.reshape(-1, tt.shape[-1])
    lg = _c(w_logits, "w_logits").reshape(-1, w_logits.shape[-1])
    n, M = lg.shape
    if tt.shape != (n, M + 1):
        raise ValueError("t must be [..., M+1] for w_logits [..., M]")
    if u is None:
        uu, u_ld = centers_u(num_samples, tt.device), 0
    else:
        uu = _c(u, "u").reshape(-1, num_samples)
        u_ld = num_samples
        if uu.shape[0] != n:
            raise ValueError("u must be [..., num_samples]")
    out = torch.empty(n, num_samples + 1, device=tt.device, dtype=torch.float32)
    with torch.cuda.device(tt.device):
        check(_lib.lib().mip360_sample_intervals(_p(tt), _p(lg), _p(uu), u_ld, n, M, num_samples, float(domain[0]),
                                                 float(domain[1]), _p(out), _stream()), "mip360_sample_intervals")
    return out.reshape(lead + (num_samples + 1,))


def compute_alpha_weights(density, tdist, dirs, opaque_background=False, weights_only=False):
    """render.compute_alpha_weights (render.py:130-151) -> (weights, alpha, trans); ``weights_only``: alpha and trans are not
    materialised (None) -- Model.__call__ keeps only the weights (models.py:233-238)."""
    dn = _c(density, "density")
    lead = tuple(dn.shape[:-1])
    S = dn.shape[-1]
    dn = dn.reshape(-1, S)
    td = _c(tdist, "tdist").reshape(-1, S + 1)
    dr = _c(dirs, "dirs").reshape(-1, 3)
    n = dn.shape[0]
    w = torch.empty(n, S, device=dn.device, dtype=torch.float32)
    a, tr = (None, None) if weights_only else (torch.empty_like(w), torch.empty_like(w))
    with torch.cuda.device(dn.device):
        check(_lib.lib().mip360_compute_alpha_weights(_p(dn), _p(td), _p(dr), n, S, int(bool(opaque_background)), _p(w), _p(a),
                                                      _p(tr), _stream()), "mip360_compute_alpha_weights")
    return tuple((x.reshape(lead + (S,)) if x is not None else None) for x in (w, a, tr))


def volumetric_rendering(rgbs, weights, tdist, bg_rgbs, t_far, compute_extras=True, extras=None):
    """render.volumetric_rendering (render.py:154-216).  Returns the reference's dict: rgb (+ acc, distance_mean, depth,
    distance_percentile_5, distance_median, distance_percentile_95 when compute_extras)."""
    if extras is not None:
        raise NotImplementedError("extras (normals etc.) are ref-NeRF features outside the hot path")
    w = _c(weights, "weights")
    lead = tuple(w.shape[:-1])
    S = w.shape[-1]
    w = w.reshape(-1, S)
    n = w.shape[0]
    c = _c(rgbs, "rgbs").reshape(n, S, 3)
    td = _c(tdist, "tdist").reshape(n, S + 1)
    if isinstance(bg_rgbs, (int, float)):          # a constant background: built once per value (no H2D copy inside a graph capture)
        key = ("bg", float(bg_rgbs), str(w.device))
        if key not in _CONST:
            _CONST[key] = torch.full((3,), float(bg_rgbs), dtype=torch.float32, device=w.device)
        bg = _CONST[key]
    else:
        bg = torch.as_tensor(bg_rgbs, dtype=torch.float32, device=w.device)
        bg = bg.expand(3).contiguous() if bg.dim() <= 1 else bg.reshape(n, 3).contiguous()
    tf = _c(t_far, "t_far").reshape(n)
    rgb = torch.empty(n, 3, device=w.device, dtype=torch.float32)
    sc = torch.empty(n, 6, device=w.device, dtype=torch.float32)
    with torch.cuda.device(w.device):
        check(_lib.lib().mip360_volumetric_rendering(_p(c), _p(w), _p(td), _p(bg), 0 if bg.dim() == 1 else 3, _p(tf), n, S,
                                                     _p(rgb), _p(sc), _stream()), "mip360_volumetric_rendering")
    r = {"rgb": rgb.reshape(lead + (3,))}
    if compute_extras:
        for i, k in enumerate(("acc", "distance_mean", "depth", "distance_percentile_5", "distance_median", "distance_percentile_95")):
            r[k] = sc[:, i].reshape(lead)
    return r


def _loss(typ, weights, tdist, prior, pred, dirs, sigma):
    pr = _c(prior, "termination_depth").reshape(-1)
    n = pr.shape[0]
    S = weights.shape[-1] if weights is not None else 1
    args = [(_c(x, nm).reshape(shape) if x is not None else None) for x, nm, shape in
            ((weights, "weights", (n, S)), (tdist, "tdist", (n, S + 1)), (pred, "predicted_depth", (n,)), (dirs, "dirs", (n, 3)))]
    L = _lib.lib()
    out = torch.empty(1, device=pr.device, dtype=torch.float32)
    ws = torch.empty(int(L.mip360_depth_loss_workspace_bytes(n)), device=pr.device, dtype=torch.uint8)
    with torch.cuda.device(pr.device):
        check(L.mip360_depth_loss(_p(args[0]), _p(args[1]), _p(pr), _p(args[2]), _p(args[3]), n, S, typ, float(sigma), _p(out),
                                  _p(ws), _stream()), "mip360_depth_loss")
    return out[0]


def depth_loss(weights, tdist, termination_depth, predicted_depth, sigma, dirs, depth_loss_type):
    """depth_loss.depth_loss (depth_loss.py:66-103); 'kl' only -- 'urf' is not used by any script of the reference."""
    if depth_loss_type == "kl":
        return _loss(DEPTH_KL, weights, tdist, termination_depth, None, dirs, sigma)
    raise NotImplementedError("Provided depth loss type not implemented.")


def depth_point_loss(distance_mean, disps_sup, depth_loss_type):
    """The mse / l1 branch of train_utils.compute_data_loss (train_utils.py:109-121)."""
    typ = {"mse": DEPTH_MSE, "l1": DEPTH_L1}[depth_loss_type]
    return _loss(typ, None, None, disps_sup, distance_mean, None, 1.0)


# ------------------------------------------------------------------------------------------------
# N4 (partial): the two regularisers of the trainer (train_utils.py:160-180), with autograd
# ------------------------------------------------------------------------------------------------
class _LossfunOuter(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t, w, t_env, w_env, eps):
        S, Pn = w.shape[-1], w_env.shape[-1]
        lead = tuple(w.shape[:-1])
        tt, ww = _c(t, "t").reshape(-1, S + 1), _c(w, "w").reshape(-1, S)
        te, we = _c(t_env, "t_env").reshape(-1, Pn + 1), _c(w_env, "w_env").reshape(-1, Pn)
        n = ww.shape[0]
        if te.shape[0] != n or tt.shape[0] != n:
            raise ValueError("lossfun_outer: leading dimensions of t, w, t_env, w_env must agree")
        out = torch.empty(n, S, device=ww.device, dtype=torch.float32)
        with torch.cuda.device(ww.device):
            check(_lib.lib().mip360_lossfun_outer(_p(tt), _p(ww), _p(te), _p(we), n, S, Pn, float(eps), None, _p(out), None,
                                                  _stream()), "mip360_lossfun_outer")
        ctx.save_for_backward(tt, ww, te, we)
        ctx.eps, ctx.env_shape = float(eps), tuple(w_env.shape)
        return out.reshape(lead + (S,))

    @staticmethod
    def backward(ctx, g):
        tt, ww, te, we = ctx.saved_tensors
        n, S = ww.shape
        Pn = we.shape[1]
        gg = g.contiguous().float().reshape(n, S)
        dwe = torch.empty(n, Pn, device=ww.device, dtype=torch.float32)
        with torch.cuda.device(ww.device):
            check(_lib.lib().mip360_lossfun_outer(_p(tt), _p(ww), _p(te), _p(we), n, S, Pn, ctx.eps, _p(gg), None, _p(dwe),
                                                  _stream()), "mip360_lossfun_outer")
        # interlevel_loss stops the gradient into (t, w) (train_utils.py:164-165) and the indices are piecewise constant
        # in t_env: w_env is the one differentiable input
        return None, None, None, dwe.reshape(ctx.env_shape), None


def lossfun_outer(t, w, t_env, w_env, eps=EPS):
    """stepfun.lossfun_outer (stepfun.py:82-89): [..., S] excess of the NeRF histogram (t, w) over the outer measure of the
    proposal histogram (t_env, w_env).  Differentiable w.r.t. ``w_env``."""
    return _LossfunOuter.apply(t, w, t_env, w_env, eps)


class _LossfunDistortion(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t, w):
        S = w.shape[-1]
        lead = tuple(w.shape[:-1])
        tt, ww = _c(t, "t").reshape(-1, S + 1), _c(w, "w").reshape(-1, S)
        n = ww.shape[0]
        out = torch.empty(n, device=ww.device, dtype=torch.float32)
        with torch.cuda.device(ww.device):
            check(_lib.lib().mip360_lossfun_distortion(_p(tt), _p(ww), n, S, None, _p(out), None, None, _stream()),
                  "mip360_lossfun_distortion")
        ctx.save_for_backward(tt, ww)
        ctx.shapes = (tuple(t.shape), tuple(w.shape))
        return out.reshape(lead)

    @staticmethod
    def backward(ctx, g):
        tt, ww = ctx.saved_tensors
        n, S = ww.shape
        gg = g.contiguous().float().reshape(n)
        dt = torch.empty(n, S + 1, device=ww.device, dtype=torch.float32)
        dw = torch.empty(n, S, device=ww.device, dtype=torch.float32)
        with torch.cuda.device(ww.device):
            check(_lib.lib().mip360_lossfun_distortion(_p(tt), _p(ww), n, S, _p(gg), None, _p(dt), _p(dw), _stream()),
                  "mip360_lossfun_distortion")
        return dt.reshape(ctx.shapes[0]), dw.reshape(ctx.shapes[1])


def lossfun_distortion(t, w):
    """stepfun.lossfun_distortion (stepfun.py:266-276): [...] = iint w_i w_j |t_i - t_j|.  Differentiable w.r.t. t and w."""
    return _LossfunDistortion.apply(t, w)


def interlevel_loss(ray_history, interlevel_loss_mult=1.0):
    """train_utils.interlevel_loss (train_utils.py:160-171): ``ray_history`` is the list of per-level dicts with 'sdist' and
    'weights'; the last level is the NeRF, the others the proposals."""
    c, w = ray_history[-1]["sdist"].detach(), ray_history[-1]["weights"].detach()
    loss = 0.
    for rr in ray_history[:-1]:
        loss = loss + lossfun_outer(c, w, rr["sdist"], rr["weights"]).mean()
    return interlevel_loss_mult * loss


def distortion_loss(ray_history, distortion_loss_mult=0.01):
    """train_utils.distortion_loss (train_utils.py:174-180)."""
    last = ray_history[-1]
    return distortion_loss_mult * lossfun_distortion(last["sdist"], last["weights"]).mean()


def _max_dilate(t, w, dilation, domain, weights_mode, renormalize, eps):
    ww = _c(w, "w")
    lead = tuple(ww.shape[:-1])
    M = ww.shape[-1]
    ww = ww.reshape(-1, M)
    tt = _c(t, "t").reshape(-1, M + 1)
    n = ww.shape[0]
    out_t = torch.empty(n, 3 * M + 1, device=ww.device, dtype=torch.float32)
    out_w = torch.empty(n, 3 * M, device=ww.device, dtype=torch.float32)
    with torch.cuda.device(ww.device):
        check(_lib.lib().mip360_max_dilate(_p(tt), _p(ww), n, M, float(dilation), float(domain[0]), float(domain[1]), int(weights_mode),
                                           int(bool(renormalize)), float(eps), _p(out_t), _p(out_w), _stream()), "mip360_max_dilate")
    return out_t.reshape(lead + (3 * M + 1,)), out_w.reshape(lead + (3 * M,))


def max_dilate(t, w, dilation, domain=(-math.inf, math.inf)):
    """stepfun.max_dilate (stepfun.py:99-113): dilate (max-pool) a non-negative step function.  No gradient (the proposal
    resampling path is not differentiated)."""
    return _max_dilate(t, w, dilation, domain, 0, False, EPS ** 2)


def max_dilate_weights(t, w, dilation, domain=(-math.inf, math.inf), renormalize=False, eps=EPS ** 2):
    """stepfun.max_dilate_weights (stepfun.py:116-128), as models.py:150-169 applies it between proposal levels."""
    return _max_dilate(t, w, dilation, domain, 1, renormalize, eps)
