"""CPU oracle for the mipnerf360 twins of the hot path -- TEST INFRASTRUCTURE.

numpy fp32 restatement of the JAX functions of ``nerf-methods/mipnerf360/internal/`` that config 3 puts on the path:
hierarchical interval resampling, alpha-compositing weights, volumetric rendering (incl. this fork's ``depth``
output) and the depth-prior losses.  Every function cites the reference ``file:line`` (relative to
``nerf-methods/mipnerf360/internal/``).  JAX is not installed in the build image, so the reference itself cannot
run here; the restatement is pinned by the reference's own known-answer and property tests, ported in
``tests/test_mip360_oracle.py``:
  stepfun_test.py:579-586 (single dominant interval resamples to linspace(3,4,11)),
  math_test.py:157-180    (sorted_interp == interp, 1e-5),
  render_test.py:408-463  (alpha weights finite over e^+-100; one 1e10-density bin -> one-hot weights),
  stepfun_test.py weighted_percentile == interp of the integrated weights.
The fork's own additions -- ``depth_loss.py`` and the ``depth`` key of ``volumetric_rendering`` -- have no reference
test: for those two the parity is UNPINNED (restated from the source only).

Only ``tests/`` may import this module.
"""
import numpy as np

F32 = np.float32
EPS = np.finfo(np.float32).eps
FMAX = np.finfo(np.float32).max


def nan_to_num(x):
    """jnp.nan_to_num(x, <positional>) as the reference calls it: the positional argument is ``copy``, so NaN -> 0
    and +-inf -> +-float32 max (math.py:125, render.py:196-200)."""
    x = np.array(x, dtype=F32)
    x[np.isnan(x)] = 0
    x[np.isposinf(x)] = FMAX
    x[np.isneginf(x)] = -FMAX
    return x


def softmax(x):
    x = np.asarray(x, F32)
    e = np.exp(x - x.max(-1, keepdims=True), dtype=F32)
    return (e / e.sum(-1, keepdims=True, dtype=F32)).astype(F32)


def integrate_weights(w):
    """stepfun.py:131-150: [0, min(1, cumsum(w[:-1])), 1]."""
    w = np.asarray(w, F32)
    cw = np.minimum(F32(1), np.cumsum(w[..., :-1], axis=-1, dtype=F32))
    shape = cw.shape[:-1] + (1,)
    return np.concatenate([np.zeros(shape, F32), cw, np.ones(shape, F32)], axis=-1)


def sorted_interp(x, xp, fp):
    """math.py:108-127 (the O(S*M) masked max/min form, evaluated literally)."""
    x, xp, fp = (np.asarray(a, F32) for a in (x, xp, fp))
    mask = x[..., None, :] >= xp[..., :, None]

    def find_interval(v):
        v0 = np.max(np.where(mask, v[..., None], v[..., :1, None]), -2)
        v1 = np.min(np.where(~mask, v[..., None], v[..., -1:, None]), -2)
        return v0, v1

    fp0, fp1 = find_interval(fp)
    xp0, xp1 = find_interval(xp)
    with np.errstate(divide="ignore", invalid="ignore"):
        offset = np.clip(nan_to_num((x - xp0) / (xp1 - xp0)), 0, 1).astype(F32)
    return (fp0 + offset * (fp1 - fp0)).astype(F32)


def invert_cdf(u, t, w_logits):
    """stepfun.py:153-161 with use_gpu_resampling=False (models.py:71 default)."""
    return sorted_interp(u, integrate_weights(softmax(w_logits)), np.asarray(t, F32))


def centers_u(num_samples):
    """stepfun.py:193-199: rng=None, deterministic_center=True."""
    pad = 1 / (2 * num_samples)
    return np.linspace(pad, 1. - pad - EPS, num_samples).astype(F32)


def jitter_base_u(num_samples):
    """stepfun.py:203-209: the linspace the random jitter is added to; returns (base, max_jitter)."""
    u_max = EPS + (1 - EPS) / num_samples
    max_jitter = (1 - u_max) / (num_samples - 1) - EPS
    return np.linspace(0, 1 - u_max, num_samples).astype(F32), F32(max_jitter)


def sample_intervals(u, t, w_logits, num_samples, domain=(-np.inf, np.inf)):
    """stepfun.py:214-263.  ``u`` [..., num_samples] are the inverse-CDF ordinates (None = the deterministic
    centres); the reference draws the jitter with jax.random, which no other RNG reproduces."""
    t = np.asarray(t, F32)
    if num_samples <= 1:
        raise ValueError(f"num_samples must be > 1, is {num_samples}.")
    if u is None:
        u = np.broadcast_to(centers_u(num_samples), t.shape[:-1] + (num_samples,))
    centers = invert_cdf(u, t, w_logits)
    mid = ((centers[..., 1:] + centers[..., :-1]) / 2).astype(F32)
    first = np.maximum(F32(domain[0]), 2 * centers[..., :1] - mid[..., :1])
    last = np.minimum(F32(domain[1]), 2 * centers[..., -1:] - mid[..., -1:])
    return np.concatenate([first, mid, last], axis=-1).astype(F32)


def compute_alpha_weights(density, tdist, dirs, opaque_background=False):
    """render.py:130-151."""
    density, tdist, dirs = (np.asarray(a, F32) for a in (density, tdist, dirs))
    t_delta = tdist[..., 1:] - tdist[..., :-1]
    delta = t_delta * np.linalg.norm(dirs[..., None, :], axis=-1).astype(F32)
    dd = (density * delta).astype(F32)
    if opaque_background:
        dd = np.concatenate([dd[..., :-1], np.full_like(dd[..., -1:], np.inf)], axis=-1)
    with np.errstate(over="ignore", invalid="ignore"):
        alpha = (1 - np.exp(-dd)).astype(F32)
        trans = np.exp(-np.concatenate([np.zeros_like(dd[..., :1]), np.cumsum(dd[..., :-1], axis=-1, dtype=F32)], axis=-1)).astype(F32)
    return (alpha * trans).astype(F32), alpha, trans


def weighted_percentile(t, w, ps):
    """stepfun.py:298-308."""
    cw = integrate_weights(w)
    t = np.asarray(t, F32)
    flat_cw, flat_t = cw.reshape(-1, cw.shape[-1]), t.reshape(-1, t.shape[-1])
    out = np.stack([np.interp(np.array(ps, F32) / 100, c, tt) for c, tt in zip(flat_cw, flat_t)])
    return out.reshape(cw.shape[:-1] + (len(ps),)).astype(F32)


def volumetric_rendering(rgbs, weights, tdist, bg_rgbs, t_far):
    """render.py:154-216 with compute_extras=True, extras=None.  Keys: rgb, acc, distance_mean, depth (the fork's
    addition, :199-201), distance_percentile_5, distance_median, distance_percentile_95."""
    rgbs, weights, tdist, t_far = (np.asarray(a, F32) for a in (rgbs, weights, tdist, t_far))
    bg_rgbs = np.asarray(bg_rgbs, F32)
    r = {}
    acc = weights.sum(-1, dtype=F32)
    bg_w = np.maximum(F32(0), 1 - acc[..., None])
    r["rgb"] = ((weights[..., None] * rgbs).sum(-2, dtype=F32) + bg_w * bg_rgbs).astype(F32)
    r["acc"] = acc
    t_mids = (0.5 * (tdist[..., :-1] + tdist[..., 1:])).astype(F32)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        expectation = (weights * np.log(t_mids)).sum(-1, dtype=F32) / np.maximum(EPS, acc)
        r["distance_mean"] = np.clip(nan_to_num(np.exp(expectation)), tdist[..., 0], tdist[..., -1]).astype(F32)
    r["depth"] = np.clip(nan_to_num((weights * t_mids).sum(-1, dtype=F32)), tdist[..., 0], tdist[..., -1]).astype(F32)
    t_aug = np.concatenate([tdist, t_far.reshape(tdist.shape[:-1] + (1,))], axis=-1)
    w_aug = np.concatenate([weights, bg_w], axis=-1)
    pct = weighted_percentile(t_aug, w_aug, [5, 50, 95])
    for i, p in enumerate([5, 50, 95]):
        r["distance_median" if p == 50 else "distance_percentile_%d" % p] = pct[..., i]
    return r


def depth_loss_kl(weights, tdist, termination_depth, sigma, dirs):
    """depth_loss.py:66-97 -> ds_nerf_depth_loss :5-26 as the trainer reaches it (train_utils.py:122-129): the
    batch carries size-1 patch axes (datasets.py:458-472), so ``.sum(-2)`` collapses a size-1 axis and the result is
    the mean over rays x samples of the per-sample term with invalid rays (prior <= 0) zeroed but still counted
    (SURVEY.md row A14).  Quirks kept: 1e-7, division by 2*sigma (not sigma^2), no far mask."""
    weights, tdist, td, dirs = (np.asarray(a, F32) for a in (weights, tdist, termination_depth, dirs))
    steps = (0.5 * (tdist[..., :-1] + tdist[..., 1:])).astype(F32)
    lengths = (tdist[..., 1:] - tdist[..., :-1]) * np.linalg.norm(dirs[..., None, :], axis=-1).astype(F32)
    mask = (td > 0).astype(F32)
    with np.errstate(divide="ignore"):
        loss = -np.log(weights + F32(1e-7)) * np.exp(-((steps - td[:, None]) ** 2) / F32(2 * sigma)) * lengths
    return F32((loss.astype(F32) * mask[:, None]).mean(dtype=np.float64))


def depth_loss_mse(distance_mean, disps_sup):
    """train_utils.py:109-117."""
    dm, gt = np.asarray(distance_mean, F32), np.asarray(disps_sup, F32)
    m = (gt > 0).astype(F32)
    return F32(((m * dm - m * gt) ** 2).mean(dtype=np.float64))


def depth_loss_l1(distance_mean, disps_sup):
    """train_utils.py:118-121."""
    dm, gt = np.asarray(distance_mean, F32), np.asarray(disps_sup, F32)
    m = (gt > 0).astype(F32)
    return F32(np.abs(m * dm - m * gt).mean(dtype=np.float64))


def synthetic_level(batch, n_bins, seed=0):
    """Seeded inputs of config 3's shapes: sorted sdist in [0,1], logits, densities, colours, directions, priors."""
    g = np.random.default_rng(seed)
    t = np.sort(g.random((batch, n_bins + 1), dtype=F32), axis=-1)
    t[:, 0], t[:, -1] = 0, 1
    logits = (3 * g.standard_normal((batch, n_bins))).astype(F32)
    density = np.exp(g.standard_normal((batch, n_bins)) + 1).astype(F32)
    rgbs = g.random((batch, n_bins, 3), dtype=F32)
    dirs = g.standard_normal((batch, 3)).astype(F32)
    prior = (g.random(batch, dtype=F32) * 0.8 + 0.1).astype(F32)
    prior[::5] = 0
    return dict(t=t, logits=logits, density=density, rgbs=rgbs, dirs=dirs, prior=prior)


# ----------------------------------------------------------------------------------------------
# N4: the two regularisers of the mipnerf360 trainer (train_utils.py:160-180)
# pinned by the reference's own tests, ported in tests/test_mip360_oracle.py:
#   stepfun_test.py:588-622 (lossfun_outer: same / ablated point sets), :624-655 (inner <= w <= outer),
#   :657-681 (invariance to monotonic maps of t), :683-697 (self loss ~ 0), :699-735 (brute-force outer / inner),
#   :227-250 (interval_distortion vs brute force), :252-275 (lossfun_distortion == sum of interval distortions)
# ----------------------------------------------------------------------------------------------
def searchsorted(a, v):
    """stepfun.py:30-53: (idx_lo, idx_hi) with a[idx_lo] <= v < a[idx_hi]; both clamp to the first / last index outside."""
    a, v = np.asarray(a, F32), np.asarray(v, F32)
    i = np.arange(a.shape[-1])
    v_ge_a = v[..., None, :] >= a[..., :, None]
    idx_lo = np.max(np.where(v_ge_a, i[:, None], i[:1, None]), -2)
    idx_hi = np.min(np.where(~v_ge_a, i[:, None], i[-1:, None]), -2)
    return idx_lo, idx_hi


def inner_outer(t0, t1, y1):
    """stepfun.py:64-79: inner and outer measures of the step function (t1, y1) on the intervals of t0."""
    y1 = np.asarray(y1, F32)
    cy1 = np.concatenate([np.zeros_like(y1[..., :1]), np.cumsum(y1, axis=-1, dtype=F32)], axis=-1)
    idx_lo, idx_hi = searchsorted(t1, t0)
    cy1_lo = np.take_along_axis(cy1, idx_lo, axis=-1)
    cy1_hi = np.take_along_axis(cy1, idx_hi, axis=-1)
    y0_outer = cy1_hi[..., 1:] - cy1_lo[..., :-1]
    y0_inner = np.where(idx_hi[..., :-1] <= idx_lo[..., 1:], cy1_lo[..., 1:] - cy1_hi[..., :-1], F32(0))
    return y0_inner.astype(F32), y0_outer.astype(F32)


def lossfun_outer(t, w, t_env, w_env, eps=EPS):
    """stepfun.py:82-89: max(0, w - w_outer)^2 / (w + eps) per interval of t."""
    w = np.asarray(w, F32)
    _, w_outer = inner_outer(t, t_env, w_env)
    return (np.maximum(F32(0), w - w_outer) ** 2 / (w + F32(eps))).astype(F32)


def lossfun_distortion(t, w):
    """stepfun.py:266-276: iint w_i w_j |t_i - t_j| over the step function, per ray."""
    t, w = np.asarray(t, F32), np.asarray(w, F32)
    ut = (t[..., 1:] + t[..., :-1]) / F32(2)
    dut = np.abs(ut[..., :, None] - ut[..., None, :])
    loss_inter = np.sum(w * np.sum(w[..., None, :] * dut, axis=-1, dtype=F32), axis=-1, dtype=F32)
    loss_intra = np.sum(w ** 2 * (t[..., 1:] - t[..., :-1]), axis=-1, dtype=F32) / F32(3)
    return (loss_inter + loss_intra).astype(F32)


def interval_distortion(t0_lo, t0_hi, t1_lo, t1_hi):
    """stepfun.py:279-297: mean |x - y| for x in [t0_lo, t0_hi], y in [t1_lo, t1_hi] (float64: a test reference)."""
    t0_lo, t0_hi, t1_lo, t1_hi = (np.asarray(x, np.float64) for x in (t0_lo, t0_hi, t1_lo, t1_hi))
    d_disjoint = np.abs((t1_lo + t1_hi) / 2 - (t0_lo + t0_hi) / 2)
    d_overlap = (2 * (np.minimum(t0_hi, t1_hi) ** 3 - np.maximum(t0_lo, t1_lo) ** 3)
                 + 3 * (t1_hi * t0_hi * np.abs(t1_hi - t0_hi) + t1_lo * t0_lo * np.abs(t1_lo - t0_lo)
                        + t1_hi * t0_lo * (t0_lo - t1_hi) + t1_lo * t0_hi * (t1_lo - t0_hi))) / (6 * (t0_hi - t0_lo) * (t1_hi - t1_lo))
    return np.where((t0_lo > t1_hi) | (t1_lo > t0_hi), d_disjoint, d_overlap)


def interlevel_loss(c, w, proposals, mult=1.0):
    """train_utils.py:160-171: sum over proposal levels of mean(lossfun_outer(c, w, cp, wp))."""
    return F32(mult) * sum(F32(np.mean(lossfun_outer(c, w, cp, wp))) for cp, wp in proposals)


def distortion_loss(c, w, mult=0.01):
    """train_utils.py:174-180."""
    return F32(mult) * F32(np.mean(lossfun_distortion(c, w)))


# ---- proposal dilation (models.py:150-169 calls max_dilate_weights between levels) -- pinned by stepfun_test.py:277-300
def query(tq, t, y, outside_value=0):
    """stepfun.py:56-61: value of the step function (t, y) at tq."""
    idx_lo, idx_hi = searchsorted(t, tq)
    y = np.asarray(y, F32)       # (jnp clamps the out-of-range gather index of a query at/after the last edge; that value is masked)
    return np.where(idx_lo == idx_hi, F32(outside_value), np.take_along_axis(y, np.minimum(idx_lo, y.shape[-1] - 1), axis=-1))


def max_dilate(t, w, dilation, domain=(-np.inf, np.inf)):
    """stepfun.py:99-113: max-pool a non-negative step function over +-dilation.  t [..., M+1], w [..., M] ->
    t_dilate [..., 3M+1], w_dilate [..., 3M]."""
    t, w = np.asarray(t, F32), np.asarray(w, F32)
    t0 = t[..., :-1] - F32(dilation)
    t1 = t[..., 1:] + F32(dilation)
    t_dilate = np.sort(np.concatenate([t, t0, t1], axis=-1), axis=-1)
    t_dilate = np.clip(t_dilate, F32(domain[0]), F32(domain[1]))
    inside = (t0[..., None, :] <= t_dilate[..., None]) & (t1[..., None, :] > t_dilate[..., None])
    w_dilate = np.max(np.where(inside, w[..., None, :], F32(0)), axis=-1)[..., :-1]
    return t_dilate.astype(F32), w_dilate.astype(F32)


def max_dilate_weights(t, w, dilation, domain=(-np.inf, np.inf), renormalize=False, eps=EPS ** 2):
    """stepfun.py:116-128: the same on weights (through the pdf), optionally renormalised to sum 1."""
    t, w = np.asarray(t, F32), np.asarray(w, F32)
    p = w / np.maximum(F32(eps), t[..., 1:] - t[..., :-1])                     # weight_to_pdf :89-91
    t_dilate, p_dilate = max_dilate(t, p, dilation, domain)
    w_dilate = p_dilate * (t_dilate[..., 1:] - t_dilate[..., :-1])               # pdf_to_weight :94-96
    if renormalize:
        w_dilate = w_dilate / np.maximum(F32(eps), np.sum(w_dilate, axis=-1, keepdims=True, dtype=F32))
    return t_dilate, w_dilate.astype(F32)
