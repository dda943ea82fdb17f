"""CPU oracle for config 3's field and model loop (mipnerf360's Model / MLP) -- TEST INFRASTRUCTURE.

numpy fp32 restatement of what ``Model.__call__`` (nerf-methods/mipnerf360/internal/models.py:75-330) does per
sampling level under ``configs/360.gin``: proposal dilation + annealed resampling (functions of
``oracle/mip360_oracle.py``), the reciprocal ray warp (coord.py:63-99), ``render.cast_rays`` (render.py:21-127,
cone, diag=False), and ``MLP.__call__`` (models.py:398-611): ``coord.contract`` carried through
``coord.track_linearize`` (coord.py:22-60), ``lift_and_diagonalize`` on the icosahedron basis (geopoly.py:46-126),
``integrated_pos_enc`` (coord.py:101-126), the Dense stack with its skip concatenation, softplus density, and the
bottleneck / view-direction colour branch.  Citations are relative to ``nerf-methods/mipnerf360/internal/``.

JAX / flax are not installed in the build image, so the reference cannot run here.  The restatement is pinned by the
reference's OWN known-answer and property tests, ported in ``tests/test_mip360_model_oracle.py``:
  geopoly_test.py:76-99    golden table of generate_basis('icosahedron', 2) (21 x 3),
  coord_test.py:61-69      contract o s_to_t(reciprocal) spaces samples uniformly (Figure 2 of the paper),
  coord_test.py:71-91      contract is bounded by 2 and the identity inside the unit ball,
  coord_test.py:129-140    integrated_pos_enc with zero variance == pos_enc,
  coord_test.py:142-177    track_linearize of an affine map is exact  (here: contract's analytic Jacobian against
                           central differences, and J cov J^T for it),
  coord_test.py:199-221    reciprocal ray warp against its closed form,
  coord_test.py:223-228    expected_sin against a Monte-Carlo mean,
  render_test.py:279-302   conical_frustum_to_gaussian against sampled frusta (stable form == the closed form).
The Dense stack itself (matmul + bias + ReLU) has no reference test; it is restated from models.py:442-466,524-609 and
checked against an independent float64 evaluation.  flax's ``Dense`` computes in float32 with
``precision=None`` -> on GPU/TPU the default matmul precision may be bf16-like; the oracle is plain fp32 (what
``jax_default_matmul_precision=highest`` / CPU gives) -- "parity unpinned" for that choice.

Only ``tests/``, ``smoke()`` and ``bench.py``'s CPU legs may import this module.
"""
import numpy as np

try:
    from . import mip360_oracle as mo
except ImportError:      # imported by path (tests put oracle/ on sys.path)
    import mip360_oracle as mo

F32 = np.float32
EPS = np.finfo(np.float32).eps


# ---------------------------------------------------------------------------------------------------------------------
# geopoly.py
# ---------------------------------------------------------------------------------------------------------------------
def compute_sq_dist(mat0, mat1=None):
    """geopoly.py:21-31."""
    if mat1 is None:
        mat1 = mat0
    sq0, sq1 = np.sum(mat0 ** 2, 0), np.sum(mat1 ** 2, 0)
    return np.maximum(0, sq0[:, None] + sq1[None, :] - 2 * mat0.T @ mat1)


def tesselate_geodesic(base_verts, base_faces, v, eps=1e-4):
    """geopoly.py:34-77: barycentric subdivision of every face, projected to the sphere, duplicates merged in order of
    first appearance."""
    w = np.array([(i, j, v - (i + j)) for i in range(v + 1) for j in range(v + 1 - i)]) / v
    verts = []
    for face in base_faces:
        nv = w @ base_verts[face, :]
        verts.append(nv / np.sqrt(np.sum(nv ** 2, 1, keepdims=True)))
    verts = np.concatenate(verts, 0)
    sq = compute_sq_dist(verts.T)
    assignment = np.array([np.min(np.argwhere(d <= eps)) for d in sq])
    return verts[np.unique(assignment), :]


def generate_basis_icosahedron(angular_tesselation=2, eps=1e-4):
    """geopoly.generate_basis('icosahedron', v, remove_symmetries=True) (geopoly.py:80-126) -> [n, 3] (float64)."""
    a = (np.sqrt(5) + 1) / 2
    verts = np.array([(-1, 0, a), (1, 0, a), (-1, 0, -a), (1, 0, -a), (0, a, 1), (0, a, -1), (0, -a, 1), (0, -a, -1),
                      (a, 1, 0), (-a, 1, 0), (a, -1, 0), (-a, -1, 0)]) / np.sqrt(a + 2)
    faces = np.array([(0, 4, 1), (0, 9, 4), (9, 5, 4), (4, 5, 8), (4, 8, 1), (8, 10, 1), (8, 3, 10), (5, 3, 8), (5, 2, 3),
                      (2, 7, 3), (7, 10, 3), (7, 6, 10), (7, 11, 6), (11, 0, 6), (0, 1, 6), (6, 1, 10), (9, 0, 11),
                      (9, 11, 2), (9, 2, 5), (7, 2, 11)])
    verts = tesselate_geodesic(verts, faces, angular_tesselation)
    match = compute_sq_dist(verts.T, -verts.T) < eps
    verts = verts[np.any(np.triu(match), 1), :]
    return verts[:, ::-1]


_BASIS = None


def pos_basis_t():
    """MLP.setup (models.py:391-392): the transposed basis [3, 21] as float32."""
    global _BASIS
    if _BASIS is None:
        _BASIS = np.ascontiguousarray(generate_basis_icosahedron(2).T).astype(F32)
    return _BASIS


# ---------------------------------------------------------------------------------------------------------------------
# coord.py
# ---------------------------------------------------------------------------------------------------------------------
def contract(x):
    """coord.py:21-27."""
    x = np.asarray(x, F32)
    mag_sq = np.maximum(EPS, np.sum(x ** 2, axis=-1, keepdims=True, dtype=F32))
    with np.errstate(invalid="ignore", divide="ignore"):
        z = np.where(mag_sq <= 1, x, ((2 * np.sqrt(mag_sq) - 1) / mag_sq) * x)
    return z.astype(F32)


def contract_jacobian(x):
    """d contract / d x at x [..., 3] -> [..., 3, 3]: identity inside the unit ball, else
    scale I + 2 (1 - sqrt(s)) / s^2 x x^T with s = |x|^2, scale = (2 sqrt(s) - 1) / s -- what jax.linearize(contract)
    evaluates (coord.py:57)."""
    x = np.asarray(x, F32)
    s = np.maximum(EPS, np.sum(x ** 2, axis=-1, keepdims=True, dtype=F32))[..., None]
    rs = np.sqrt(s)
    scale = (2 * rs - 1) / s
    k = 2 * (1 - rs) / (s * s)
    eye = np.eye(x.shape[-1], dtype=F32)
    J = scale * eye + k * (x[..., :, None] * x[..., None, :])
    return np.where(s <= 1, eye, J).astype(F32)


def track_linearize_contract(mean, cov):
    """coord.track_linearize(coord.contract, mean, cov) (coord.py:38-60): fn_cov = J cov J^T."""
    J = contract_jacobian(mean)
    cov = np.asarray(cov, F32)
    fn_cov = np.matmul(np.matmul(J, cov), np.swapaxes(J, -1, -2))
    return contract(mean), fn_cov.astype(F32)


def s_to_t_reciprocal(s, t_near, t_far):
    """coord.construct_ray_warps(jnp.reciprocal, near, far)[1] (coord.py:63-99)."""
    s, t_near, t_far = (np.asarray(a, F32) for a in (s, t_near, t_far))
    s_near, s_far = F32(1) / t_near, F32(1) / t_far
    return (F32(1) / (s * s_far + (F32(1) - s) * s_near)).astype(F32)


def t_to_s_reciprocal(t, t_near, t_far):
    t, t_near, t_far = (np.asarray(a, F32) for a in (t, t_near, t_far))
    s_near, s_far = F32(1) / t_near, F32(1) / t_far
    return ((F32(1) / t - s_near) / (s_far - s_near)).astype(F32)


def safe_sin(x):
    """math.py:27-40: sin(where(|x| < 100 pi, x, x % (100 pi)))."""
    x = np.asarray(x, F32)
    t = F32(100 * np.pi)
    with np.errstate(invalid="ignore"):
        return np.sin(np.where(np.abs(x) < t, x, np.mod(x, t))).astype(F32)


def expected_sin(mean, var):
    """coord.py:101-104."""
    return (np.exp(F32(-0.5) * np.asarray(var, F32)) * safe_sin(mean)).astype(F32)


def integrated_pos_enc(mean, var, min_deg, max_deg):
    """coord.py:107-126."""
    mean, var = np.asarray(mean, F32), np.asarray(var, F32)
    scales = (2 ** np.arange(min_deg, max_deg)).astype(F32)
    shape = mean.shape[:-1] + (-1,)
    scaled_mean = np.reshape(mean[..., None, :] * scales[:, None], shape)
    scaled_var = np.reshape(var[..., None, :] * scales[:, None] ** 2, shape)
    return expected_sin(np.concatenate([scaled_mean, scaled_mean + F32(0.5 * np.pi)], axis=-1),
                        np.concatenate([scaled_var] * 2, axis=-1))


def lift_and_diagonalize(mean, cov, basis):
    """coord.py:129-133."""
    mean, cov, basis = (np.asarray(a, F32) for a in (mean, cov, basis))
    fn_mean = np.matmul(mean, basis)
    fn_cov_diag = np.sum(basis * np.matmul(cov, basis), axis=-2, dtype=F32)
    return fn_mean.astype(F32), fn_cov_diag.astype(F32)


def pos_enc(x, min_deg, max_deg, append_identity=True):
    """coord.py:136-148 (plain sin, not safe_sin)."""
    x = np.asarray(x, F32)
    scales = (2 ** np.arange(min_deg, max_deg)).astype(F32)
    shape = x.shape[:-1] + (-1,)
    scaled_x = np.reshape(x[..., None, :] * scales[:, None], shape)
    four_feat = np.sin(np.concatenate([scaled_x, scaled_x + F32(0.5 * np.pi)], axis=-1)).astype(F32)
    return np.concatenate([x, four_feat], axis=-1) if append_identity else four_feat


# ---------------------------------------------------------------------------------------------------------------------
# render.py: cast_rays
# ---------------------------------------------------------------------------------------------------------------------
def lift_gaussian(d, t_mean, t_var, r_var):
    """render.py:21-39, diag=False."""
    d = np.asarray(d, F32)
    mean = d[..., None, :] * t_mean[..., None]
    d_mag_sq = np.maximum(F32(1e-10), np.sum(d ** 2, axis=-1, keepdims=True, dtype=F32))
    d_outer = d[..., :, None] * d[..., None, :]
    eye = np.eye(d.shape[-1], dtype=F32)
    null_outer = eye - d[..., :, None] * (d / d_mag_sq)[..., None, :]
    t_cov = t_var[..., None, None] * d_outer[..., None, :, :]
    xy_cov = r_var[..., None, None] * null_outer[..., None, :, :]
    return mean.astype(F32), (t_cov + xy_cov).astype(F32)


def conical_frustum_to_gaussian(d, t0, t1, base_radius):
    """render.py:42-73, stable=True, diag=False."""
    t0, t1, base_radius = (np.asarray(a, F32) for a in (t0, t1, base_radius))
    mu = (t0 + t1) / F32(2)
    hw = (t1 - t0) / F32(2)
    denom = np.maximum(EPS, F32(3) * mu ** 2 + hw ** 2)
    t_mean = mu + (F32(2) * mu * hw ** 2) / denom
    t_var = (hw ** 2) / F32(3) - F32(4 / 15) * hw ** 4 * (F32(12) * mu ** 2 - hw ** 2) / denom ** 2
    r_var = (mu ** 2) / F32(4) + F32(5 / 12) * hw ** 2 - F32(4 / 15) * (hw ** 4) / denom
    r_var = r_var * base_radius ** 2
    return lift_gaussian(d, t_mean.astype(F32), t_var.astype(F32), r_var.astype(F32))


def cast_rays(tdist, origins, directions, radii):
    """render.py:101-127 with ray_shape='cone', diag=False.  radii [..., 1]."""
    tdist = np.asarray(tdist, F32)
    means, covs = conical_frustum_to_gaussian(directions, tdist[..., :-1], tdist[..., 1:], np.asarray(radii, F32))
    return (means + np.asarray(origins, F32)[..., None, :]).astype(F32), covs


# ---------------------------------------------------------------------------------------------------------------------
# models.py: MLP
# ---------------------------------------------------------------------------------------------------------------------
NUM_DEG = 12            # max_deg_point (models.py:351)
DEG_VIEW = 4            # models.py:356
BOTTLENECK = 256        # models.py:346
VIEW_WIDTH = 128        # models.py:348
RGB_PADDING = 0.001     # models.py:373
DENSITY_BIAS = -1.0     # models.py:367
SKIP = 4                # models.py:353


def dense_shapes(net_depth, net_width, has_rgb):
    """(in, out) of Dense_0.. in flax's construction order (models.py:442-466,508-597)."""
    feat = 2 * NUM_DEG * 21
    shapes, cur = [], feat
    for i in range(net_depth):
        shapes.append((cur, net_width))
        cur = net_width + (feat if (i % SKIP == 0 and i > 0) else 0)
    shapes.append((cur, 1))
    if has_rgb:
        shapes.append((cur, BOTTLENECK))
        shapes.append((BOTTLENECK + 3 + 3 * 2 * DEG_VIEW, VIEW_WIDTH))
        shapes.append((VIEW_WIDTH, 3))
    return shapes


def init_mlp_params(net_depth, net_width, has_rgb, seed):
    """flax Dense defaults under the gin config: kernel he_uniform = U(+-sqrt(6 / fan_in)) (models.py:353,428-429), bias
    zeros.  (numpy's generator, not jax.random: the values, not the draws, are what the path consumes.)"""
    rng = np.random.default_rng(seed)
    params = []
    for fan_in, fan_out in dense_shapes(net_depth, net_width, has_rgb):
        lim = np.sqrt(6.0 / fan_in)
        params.append((rng.uniform(-lim, lim, size=(fan_in, fan_out)).astype(F32), np.zeros(fan_out, F32)))
    return params


def softplus(x):
    x = np.asarray(x, F32)
    return (np.maximum(x, 0) + np.log1p(np.exp(-np.abs(x)))).astype(F32)


def encode_gaussians(means, covs):
    """predict_density's featurisation (models.py:432-441): contract -> lift -> IPE.  [..., 3], [..., 3, 3] -> [..., 504]."""
    m, c = track_linearize_contract(means, covs)
    lm, lv = lift_and_diagonalize(m, c, pos_basis_t())
    return integrated_pos_enc(lm, lv, 0, NUM_DEG)


def lifted_gaussians_ordered(tdist, origins, directions, radii, full=False):
    """cast_rays -> track_linearize(contract) -> lift_and_diagonalize (render.py:21-127, coord.py:21-60,129-133) written out
    element by element with ONE fixed fp32 evaluation order -- sums left to right, every product rounded before it is
    added (no FMA), hw**4 = (hw*hw)*(hw*hw), the Jacobian in its closed form J = scale I + (k x) x^T.  The literal
    restatement above leaves these orders to numpy's matmul / power; the path is badly conditioned (the degree-11 features
    turn one ulp of a lifted mean into 5e-4, J cov J^T cancels ten digits for distant samples), so the CUDA kernel mirrors
    THIS order operation for operation and is compared against it; tests/test_mip360_model_oracle.py holds it to the
    literal restatement within that conditioning.  -> lifted means, variances [n, S, 21]."""
    f = F32
    tdist = np.asarray(tdist, F32)
    t0, t1 = tdist[..., :-1], tdist[..., 1:]
    o = [np.asarray(origins, F32)[..., i, None] for i in range(3)]
    d = [np.asarray(directions, F32)[..., i, None] for i in range(3)]
    radius = np.asarray(radii, F32).reshape(tdist.shape[:-1] + (1,))
    with np.errstate(all="ignore"):
        mu, hw = (t0 + t1) / f(2), (t1 - t0) / f(2)
        mu2, hw2 = mu * mu, hw * hw
        hw4 = hw2 * hw2
        denom = np.maximum(EPS, f(3) * mu2 + hw2)
        t_mean = mu + ((f(2) * mu) * hw2) / denom
        t_var = hw2 / f(3) - ((f(4 / 15) * hw4) * (f(12) * mu2 - hw2)) / (denom * denom)
        r_var = (mu2 / f(4) + f(5 / 12) * hw2) - (f(4 / 15) * hw4) / denom
        r_var = r_var * (radius * radius)
        dms = np.maximum(f(1e-10), (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2])
        x = [d[i] * t_mean + o[i] for i in range(3)]
        cov = [[t_var * (d[i] * d[j]) + r_var * (f(1 if i == j else 0) - d[i] * (d[j] / dms)) for j in range(3)] for i in range(3)]
        s = np.maximum(EPS, (x[0] * x[0] + x[1] * x[1]) + x[2] * x[2])
        inside = s <= 1
        rs = np.sqrt(s)
        scale = (f(2) * rs - f(1)) / s
        k = (f(2) * (f(1) - rs)) / (s * s)
        z = [np.where(inside, x[i], scale * x[i]) for i in range(3)]
        J = [[(scale if i == j else f(0)) + (k * x[i]) * x[j] for j in range(3)] for i in range(3)]
        T = [[(J[i][0] * cov[0][j] + J[i][1] * cov[1][j]) + J[i][2] * cov[2][j] for j in range(3)] for i in range(3)]
        C = [[np.where(inside, cov[i][j], (T[i][0] * J[j][0] + T[i][1] * J[j][1]) + T[i][2] * J[j][2]) for j in range(3)] for i in range(3)]
        B = pos_basis_t()
        lm, lv = [], []
        for b in range(B.shape[1]):
            b0, b1, b2 = B[0, b], B[1, b], B[2, b]
            lm.append((z[0] * b0 + z[1] * b1) + z[2] * b2)
            c = [(C[i][0] * b0 + C[i][1] * b1) + C[i][2] * b2 for i in range(3)]
            lv.append((b0 * c[0] + b1 * c[1]) + b2 * c[2])
    lm, lv = np.stack(lm, -1).astype(F32), np.stack(lv, -1).astype(F32)
    if full:        # + the contracted means [n, S, 3] and covariances [n, S, 3, 3]
        return lm, lv, np.stack(z, -1).astype(F32), np.stack([np.stack(r, -1) for r in C], -2).astype(F32)
    return lm, lv


def encode_ordered(tdist, origins, directions, radii):
    """The field's featurisation with the fixed evaluation order: [n, S+1] fenceposts -> [n, S, 504]."""
    lm, lv = lifted_gaussians_ordered(tdist, origins, directions, radii)
    return integrated_pos_enc(lm, lv, 0, NUM_DEG)


def mlp_forward(params, net_depth, has_rgb, means, covs, viewdirs=None):
    """MLP.__call__ (models.py:398-611) with disable_density_normals=True, rng=None: -> density [...], rgb [..., 3] or None."""
    return mlp_forward_features(params, net_depth, has_rgb, encode_gaussians(means, covs), viewdirs)


def mlp_forward_features(params, net_depth, has_rgb, x, viewdirs=None):
    """The part of MLP.__call__ behind the featurisation (models.py:442-466, 507-609): x [..., S, 504] -> density, rgb."""
    x = np.asarray(x, F32)
    inputs = x
    for i in range(net_depth):
        k, b = params[i]
        x = np.maximum(np.matmul(x, k, dtype=F32) + b, 0).astype(F32)
        if i % SKIP == 0 and i > 0:
            x = np.concatenate([x, inputs], axis=-1)
    k, b = params[net_depth]
    raw_density = (np.matmul(x, k, dtype=F32) + b)[..., 0]
    density = softplus(raw_density + F32(DENSITY_BIAS))
    if not has_rgb:
        return density, None
    k, b = params[net_depth + 1]
    bottleneck = (np.matmul(x, k, dtype=F32) + b).astype(F32)
    dir_enc = pos_enc(viewdirs, 0, DEG_VIEW, append_identity=True)
    dir_enc = np.broadcast_to(dir_enc[..., None, :], bottleneck.shape[:-1] + (dir_enc.shape[-1],))
    h = np.concatenate([bottleneck, dir_enc], axis=-1)
    k, b = params[net_depth + 2]
    h = np.maximum(np.matmul(h, k, dtype=F32) + b, 0).astype(F32)
    k, b = params[net_depth + 3]
    raw = (np.matmul(h, k, dtype=F32) + b).astype(F32)
    rgb = F32(1) / (F32(1) + np.exp(-raw))
    rgb = (rgb * F32(1 + 2 * RGB_PADDING) - F32(RGB_PADDING)).astype(F32)
    return density, rgb


# ---------------------------------------------------------------------------------------------------------------------
# models.py: Model.__call__
# ---------------------------------------------------------------------------------------------------------------------
def field_level(params, net_depth, has_rgb, sdist, near, far, origins, directions, viewdirs, radii):
    """models.py:203-231: s_to_t -> cast_rays -> MLP.  near / far / radii [n, 1]."""
    tdist = s_to_t_reciprocal(sdist, near, far)
    density, rgb = mlp_forward_features(params, net_depth, has_rgb, encode_ordered(tdist, origins, directions, radii), viewdirs)
    return tdist, density, rgb


def model_forward(prop_params, nerf_params, rays, train_frac=1.0, u_levels=None, num_prop_samples=64, num_nerf_samples=32,
                  anneal_slope=10.0, dilation_multiplier=0.5, dilation_bias=0.0025, resample_padding=0.0, bg_rgb=1.0,
                  opaque_background=True, prop_shape=(4, 256), nerf_shape=(8, 1024)):
    """Model.__call__ (models.py:75-330) under configs/360.gin: three levels (PropMLP, PropMLP, NerfMLP), one shared
    PropMLP.  ``rays``: dict origins/directions/viewdirs [n,3], radii/near/far [n,1].  ``u_levels``: per level None (the
    rng=None centres) or the inverse-CDF ordinates [n, num_samples].  Returns (renderings, ray_history)."""
    n = rays["origins"].shape[0]
    near, far = np.asarray(rays["near"], F32), np.asarray(rays["far"], F32)
    sdist = np.concatenate([np.zeros_like(near), np.ones_like(far)], axis=-1)
    weights = np.ones_like(near)
    prod_num_samples = 1
    renderings, history = [], []
    for i_level in range(3):
        is_prop = i_level < 2
        num_samples = num_prop_samples if is_prop else num_nerf_samples
        dilation = dilation_bias + dilation_multiplier * (1.0 - 0.0) / prod_num_samples
        prod_num_samples *= num_samples
        if i_level > 0:
            sdist, weights = mo.max_dilate_weights(sdist, weights, dilation, domain=(0.0, 1.0), renormalize=True)
            sdist, weights = sdist[..., 1:-1], weights[..., 1:-1]
        anneal = (anneal_slope * train_frac) / ((anneal_slope - 1) * train_frac + 1) if anneal_slope > 0 else 1.0
        with np.errstate(divide="ignore"):
            logits = np.where(sdist[..., 1:] > sdist[..., :-1], F32(anneal) * np.log(weights + F32(resample_padding)), -np.inf).astype(F32)
        u = None if u_levels is None else u_levels[i_level]
        sdist = mo.sample_intervals(u, sdist, logits, num_samples, domain=(0.0, 1.0))
        params, shape = (prop_params, prop_shape) if is_prop else (nerf_params, nerf_shape)
        tdist, density, rgb = field_level(params, shape[0], not is_prop, sdist, near, far, rays["origins"], rays["directions"],
                                          rays["viewdirs"], rays["radii"])
        if rgb is None:
            rgb = np.zeros(density.shape + (3,), F32)         # disable_rgb: zeros_like(means) (models.py:511-512)
        weights = mo.compute_alpha_weights(density, tdist, rays["directions"], opaque_background=opaque_background)[0]
        rendering = mo.volumetric_rendering(rgb, weights, tdist, F32(bg_rgb), far)
        renderings.append(rendering)
        history.append(dict(sdist=sdist, tdist=tdist, weights=weights, density=density, rgb=rgb))
    return renderings, history


def synthetic_rays(n, seed=0):
    """A seeded 360-style ray batch: origins near the unit ball, unnormalised directions, pixel-footprint radii, near 0.2,
    far 1e6 (configs/360.gin:2-3)."""
    rng = np.random.default_rng(seed)
    o = rng.normal(size=(n, 3))
    o = (o / np.linalg.norm(o, axis=-1, keepdims=True) * rng.uniform(0.2, 1.2, size=(n, 1))).astype(F32)
    d = rng.normal(size=(n, 3))
    d = (d / np.linalg.norm(d, axis=-1, keepdims=True) * rng.uniform(0.9, 1.3, size=(n, 1))).astype(F32)
    v = (d / np.linalg.norm(d, axis=-1, keepdims=True)).astype(F32)
    radii = rng.uniform(5e-4, 2e-3, size=(n, 1)).astype(F32)
    return dict(origins=o, directions=d, viewdirs=v, radii=radii, near=np.full((n, 1), 0.2, F32), far=np.full((n, 1), 1e6, F32))
