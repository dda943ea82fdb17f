"""Pins oracle/mip360_model_oracle.py (config 3's field: ray warp, conical frusta, contraction, integrated positional
encoding, Dense stack) with the reference's own known-answer / property tests, ported from absltest + JAX to numpy
(nerf-methods/mipnerf360/tests/{geopoly,coord,render}_test.py)."""
import numpy as np
import pytest

import mip360_model_oracle as MM

F32 = np.float32

# geopoly_test.py:79-99 -- the reference's golden table for generate_basis('icosahedron', 2)
BASIS_GOLDEN = np.array([
    [0.85065081, 0.00000000, 0.52573111], [0.80901699, 0.50000000, 0.30901699], [0.52573111, 0.85065081, 0.00000000],
    [1.00000000, 0.00000000, 0.00000000], [0.80901699, 0.50000000, -0.30901699], [0.85065081, 0.00000000, -0.52573111],
    [0.30901699, 0.80901699, -0.50000000], [0.00000000, 0.52573111, -0.85065081], [0.50000000, 0.30901699, -0.80901699],
    [0.00000000, 1.00000000, 0.00000000], [-0.52573111, 0.85065081, 0.00000000], [-0.30901699, 0.80901699, -0.50000000],
    [0.00000000, 0.52573111, 0.85065081], [-0.30901699, 0.80901699, 0.50000000], [0.30901699, 0.80901699, 0.50000000],
    [0.50000000, 0.30901699, 0.80901699], [0.50000000, -0.30901699, 0.80901699], [0.00000000, 0.00000000, 1.00000000],
    [-0.50000000, 0.30901699, 0.80901699], [-0.80901699, 0.50000000, 0.30901699], [-0.80901699, 0.50000000, -0.30901699]])


def stable_pos_enc(x, n):
    """coord_test.py:35-43."""
    sin_x, cos_x = np.sin(x), np.cos(x)
    output = []
    rotmat = np.array([[cos_x, -sin_x], [sin_x, cos_x]], dtype="double")
    for _ in range(n):
        output.append(rotmat[::-1, 0, :])
        rotmat = np.einsum("ijn,jkn->ikn", rotmat, rotmat)
    return np.reshape(np.transpose(np.stack(output, 0), [2, 1, 0]), [-1, 2 * n])


def test_generate_basis_golden():
    """geopoly_test.py:76-99: same vectors, and (stronger) in the same order -- the order fixes the feature layout."""
    basis = MM.generate_basis_icosahedron(2)
    assert basis.shape == (21, 3)
    np.testing.assert_allclose(basis, BASIS_GOLDEN, atol=1e-7)
    assert MM.pos_basis_t().shape == (3, 21) and MM.pos_basis_t().dtype == F32


def test_contract_matches_special_case():
    """coord_test.py:61-69 -- Figure 2 of arxiv 2111.12077: contract(s_to_t(s)) is uniform in s."""
    n = 10
    s = np.linspace(0, 1 - MM.EPS, n + 1).astype(F32)
    with np.errstate(divide="ignore"):
        t = MM.s_to_t_reciprocal(s, F32(1), F32(np.inf))
    tc = MM.contract(t[:, None])[:, 0]
    delta = tc[1:] - tc[:-1]
    np.testing.assert_allclose(delta, np.full_like(delta, 1 / n), atol=1e-5, rtol=1e-5)


def test_contract_is_bounded():
    """coord_test.py:71-78."""
    g = np.random.default_rng(0)
    x = np.where(g.random((10000, 3)) < 0.5, 1, -1) * np.exp(g.uniform(-10, 10, (10000, 3)))
    assert np.max(MM.contract(x.astype(F32))) <= 2


def test_contract_is_noop_when_norm_is_leq_one():
    """coord_test.py:80-91."""
    g = np.random.default_rng(0)
    x = g.standard_normal((10000, 3)).astype(F32)
    xc = x / np.maximum(1, np.linalg.norm(x, axis=-1, keepdims=True))
    np.testing.assert_allclose(xc, MM.contract(xc), atol=1e-5, rtol=1e-5)


def test_contract_jacobian_against_central_differences():
    """The piece track_linearize gets from jax.linearize (coord.py:57), restated analytically: checked against float64
    central differences of contract itself, inside and outside the unit ball (finite at the origin, coord_test.py:93-97)."""
    g = np.random.default_rng(1)
    x = np.concatenate([g.standard_normal((200, 3)) * 0.4, g.standard_normal((200, 3)) * 3, np.zeros((1, 3))]).astype(F32)

    def contract64(v):
        s = np.maximum(np.finfo(np.float32).eps, np.sum(v ** 2, -1, keepdims=True))
        return np.where(s <= 1, v, (2 * np.sqrt(s) - 1) / s * v)

    J = MM.contract_jacobian(x)
    assert np.all(np.isfinite(J))
    h = 1e-6
    keep = np.abs(np.sum(x.astype(np.float64) ** 2, -1) - 1) > 1e-3          # away from the kink at |x| = 1
    for j in range(3):
        e = np.zeros(3); e[j] = h
        col = (contract64(x.astype(np.float64) + e) - contract64(x.astype(np.float64) - e)) / (2 * h)
        np.testing.assert_allclose(J[keep, :, j], col[keep], atol=2e-5, rtol=1e-4)


def test_track_linearize_is_exact_for_the_linear_part():
    """coord_test.py:142-177 in the form that applies to contract: where contract is linear (inside the ball, J = I) the
    covariance passes through unchanged; outside, fn_cov = J cov J^T stays symmetric PSD."""
    g = np.random.default_rng(2)
    half = g.standard_normal((50, 3, 3)).astype(F32)
    cov = np.matmul(half, np.swapaxes(half, -1, -2))
    inside = (g.standard_normal((50, 3)) * 0.2).astype(F32)
    m, c = MM.track_linearize_contract(inside, cov)
    np.testing.assert_allclose(m, inside, atol=1e-6)
    np.testing.assert_allclose(c, cov, atol=1e-5, rtol=1e-5)
    outside = (g.standard_normal((50, 3)) * 4 + 3).astype(F32)
    m, c = MM.track_linearize_contract(outside, cov)
    np.testing.assert_allclose(c, np.swapaxes(c, -1, -2), atol=1e-5)
    assert np.all(np.linalg.eigvalsh(c.astype(np.float64)) > -1e-5)
    assert np.all(np.linalg.norm(m, axis=-1) <= 2)


def test_pos_enc_matches_integrated():
    """coord_test.py:129-140."""
    x = np.linspace(-np.pi, np.pi, 10000).astype(F32)
    z_ipe = MM.integrated_pos_enc(x, np.zeros_like(x), 0, 10)
    z_pe = MM.pos_enc(x, 0, 10, append_identity=False)
    np.testing.assert_allclose(z_pe, z_ipe, atol=1e-4)


@pytest.mark.parametrize("n,tol", [(5, 1e-5), (10, 1e-4), (15, 0.005)])
def test_pos_enc_against_stable(n, tol):
    """coord_test.py:112-127."""
    x = np.linspace(-np.pi, np.pi, 10001)
    z = MM.pos_enc(x[:, None].astype(F32), 0, n, append_identity=False)
    assert np.max(np.abs(z - stable_pos_enc(x, n))) < tol


def test_construct_ray_warps_special_reciprocal():
    """coord_test.py:199-221."""
    g = np.random.default_rng(0)
    n = 100
    t_near = np.exp(g.standard_normal(n)).astype(F32)
    t_far = (t_near + np.exp(g.standard_normal(n))).astype(F32)
    u = g.random(n).astype(F32)
    t = t_near * (1 - u) + t_far * u
    s = g.random(n).astype(F32)
    np.testing.assert_allclose(MM.t_to_s_reciprocal(t, t_near, t_far), (t_far * (t - t_near)) / (t * (t_far - t_near)), atol=1e-5, rtol=1e-5)
    np.testing.assert_allclose(MM.s_to_t_reciprocal(s, t_near, t_far), 1 / (s / t_far + (1 - s) / t_near), atol=1e-5, rtol=1e-5)
    # coord_test.py:180-197: extents
    np.testing.assert_allclose(MM.s_to_t_reciprocal(np.zeros(n, F32), t_near, t_far), t_near, atol=1e-5, rtol=1e-5)
    np.testing.assert_allclose(MM.s_to_t_reciprocal(np.ones(n, F32), t_near, t_far), t_far, atol=1e-5, rtol=1e-5)


def test_expected_sin():
    """coord_test.py:223-228."""
    samples = np.random.default_rng(0).standard_normal(10000)
    for mu, var in [(0, 1), (1, 3), (-2, .2), (10, 10)]:
        np.testing.assert_allclose(MM.expected_sin(F32(mu), F32(var)), np.mean(np.sin(np.sqrt(var) * samples + mu)), atol=2e-2)


def test_integrated_pos_enc_against_samples():
    """coord_test.py:230-263: the IPE of a Gaussian is the mean encoding of its samples."""
    g = np.random.default_rng(0)
    max_deg = 4
    for _ in range(5):
        mean = g.standard_normal(2)
        half = g.standard_normal((2, 2))
        cov = half @ half.T
        enc = MM.integrated_pos_enc(mean.astype(F32), np.diag(cov).astype(F32), 0, max_deg)
        samples = g.multivariate_normal(mean, cov, 100000)
        enc_samples = np.concatenate([stable_pos_enc(x, max_deg) for x in tuple(samples.T)], axis=-1)
        enc_gt = np.mean(enc_samples, 0).reshape([2, max_deg * 2]).T.reshape([-1])
        np.testing.assert_allclose(enc, enc_gt, rtol=1e-2, atol=1e-2)


def _sample_conical_frustum(g, num, d, t0, t1, base_radius):
    """render_test.py:66-94."""
    u = g.random(num)
    t = (t0 ** 3 * (1 - u) + t1 ** 3 * u) ** (1 / 3)
    theta = g.uniform(0, 2 * np.pi, num)
    r = base_radius * t * np.sqrt(g.random(num))
    dn = d / np.linalg.norm(d)
    basis = np.linalg.svd(np.eye(3) - dn[:, None] * dn[None, :])[0][:, :2]
    return ((basis[:, 0:1] * r * np.cos(theta)) + (basis[:, 1:2] * r * np.sin(theta)) + d[:, None] * t).T


def test_conical_frustum_against_samples():
    """render_test.py:279-302 (stable form, full covariance)."""
    g = np.random.default_rng(0)
    for _ in range(10):
        z_mean, z_delta = g.uniform(1.5, 3, 4), g.uniform(0.1, 0.3, 4)
        t0, t1 = z_mean - z_delta, z_mean + z_delta
        r = g.uniform(0.01, 0.05)
        d = g.standard_normal(3)
        d = d / np.linalg.norm(d) * g.uniform(0.8, 1.2)
        mean, cov = MM.conical_frustum_to_gaussian(d.astype(F32), t0.astype(F32), t1.astype(F32), F32(r))
        for i in range(4):
            s = _sample_conical_frustum(g, 100000, d, t0[i], t1[i], r)
            np.testing.assert_allclose(mean[i], s.mean(0), atol=0.002)
            np.testing.assert_allclose(cov[i], np.cov(s.T), atol=0.0004)


def test_ordered_featurisation_agrees_with_the_literal_restatement():
    """lifted_gaussians_ordered (the fixed-evaluation-order form the CUDA kernel mirrors) against the literal restatement
    cast_rays -> track_linearize_contract -> lift_and_diagonalize: lifted means to 2 ulp of their range, variances to 1e-4 of
    the Gaussian's largest one (J cov J^T cancels), features within the conditioning 2^deg * 5e-7 + the variances' share."""
    rays = MM.synthetic_rays(64, seed=7)
    g = np.random.default_rng(8)
    sd = np.sort(g.random((64, 33)), -1).astype(F32)
    sd[:, 0], sd[:, -1] = 0, 1
    tdist = MM.s_to_t_reciprocal(sd, rays["near"], rays["far"])
    means, covs = MM.cast_rays(tdist, rays["origins"], rays["directions"], rays["radii"])
    m, c = MM.track_linearize_contract(means, covs)
    lm_ref, lv_ref = MM.lift_and_diagonalize(m, c, MM.pos_basis_t())
    lm, lv = MM.lifted_gaussians_ordered(tdist, rays["origins"], rays["directions"], rays["radii"])
    assert np.max(np.abs(lm - lm_ref)) <= 1e-6
    m64, c64 = means.astype(np.float64), covs.astype(np.float64)       # float64 evaluation of the same Gaussians
    s = np.sum(m64 ** 2, -1)[..., None, None]
    J = np.where(s <= 1, np.eye(3), (2 * np.sqrt(s) - 1) / s * np.eye(3) + 2 * (1 - np.sqrt(s)) / s ** 2 * m64[..., :, None] * m64[..., None, :])
    B = MM.pos_basis_t().astype(np.float64)
    lv64 = np.sum(B * (J @ c64 @ J @ B), -2)
    ref_scale = lv64.max(-1, keepdims=True)
    assert np.max(np.abs(lv - lv64) / ref_scale) <= np.max(np.abs(lv_ref - lv64) / ref_scale) * 2 + 1e-5      # no worse than the literal form
    enc, enc_ref = MM.encode_ordered(tdist, rays["origins"], rays["directions"], rays["radii"]), MM.encode_gaussians(means, covs)
    deg = np.tile(np.repeat(np.arange(12), 21), 2)
    # one ulp of a lifted mean is 2^deg * 2.4e-7 in the argument; the variances of the two orders differ by up to 1.5 % for
    # distant samples, which moves exp(-var 4^deg / 2) by a few 1e-3 where that factor is O(1) (degrees 10-11)
    assert np.all(np.abs(enc - enc_ref).max(axis=(0, 1)) <= (2.0 ** deg) * 3e-6 + 1e-6)


def test_dense_shapes_of_the_gin_configuration():
    """models.py:442-466,524-597 under configs/360.gin:12-19."""
    assert MM.dense_shapes(4, 256, False) == [(504, 256), (256, 256), (256, 256), (256, 256), (256, 1)]
    nerf = MM.dense_shapes(8, 1024, True)
    assert nerf[:8] == [(504, 1024)] + [(1024, 1024)] * 4 + [(1528, 1024)] + [(1024, 1024)] * 2
    assert nerf[8:] == [(1024, 1), (1024, 256), (283, 128), (128, 3)]
    assert sum(i * o for i, o in nerf) == 8672000 and sum(i * o for i, o in MM.dense_shapes(4, 256, False)) == 325888


def test_mlp_forward_against_float64():
    """The Dense stack restated in float64 from models.py:442-466 (skip after layer 4, density from the last hidden
    layer, bottleneck + view encoding) agrees with the fp32 oracle."""
    rays = MM.synthetic_rays(6, seed=3)
    sdist = np.broadcast_to(np.linspace(0, 1, 9, dtype=F32), (6, 9))
    params = MM.init_mlp_params(8, 256, True, seed=5)
    tdist, density, rgb = MM.field_level(params, 8, True, sdist, rays["near"], rays["far"], rays["origins"], rays["directions"],
                                         rays["viewdirs"], rays["radii"])
    assert tdist.shape == (6, 9) and density.shape == (6, 8) and rgb.shape == (6, 8, 3)
    x = MM.encode_ordered(tdist, rays["origins"], rays["directions"], rays["radii"]).astype(np.float64)
    inputs, p = x, [(k.astype(np.float64), b.astype(np.float64)) for k, b in params]
    for i in range(8):
        x = np.maximum(x @ p[i][0] + p[i][1], 0)
        if i == 4:
            x = np.concatenate([x, inputs], -1)
    raw = (x @ p[8][0] + p[8][1])[..., 0] - 1
    np.testing.assert_allclose(density, np.logaddexp(raw, 0), rtol=2e-5, atol=1e-6)
    bott = x @ p[9][0] + p[9][1]
    de = np.broadcast_to(MM.pos_enc(rays["viewdirs"], 0, 4).astype(np.float64)[:, None, :], bott.shape[:-1] + (27,))
    h = np.maximum(np.concatenate([bott, de], -1) @ p[10][0] + p[10][1], 0)
    col = 1 / (1 + np.exp(-(h @ p[11][0] + p[11][1]))) * 1.002 - 0.001
    np.testing.assert_allclose(rgb, col, atol=2e-6)
    assert rgb.min() >= -0.001 and rgb.max() <= 1.001


def test_model_forward_three_levels():
    """Model.__call__ (models.py:138-310) as configs/360.gin runs it: 64 / 64 / 32 intervals, weights sum to <= 1
    (opaque background: exactly 1), fenceposts sorted inside the ray warp's range."""
    rays = MM.synthetic_rays(16, seed=1)
    prop = MM.init_mlp_params(4, 256, False, seed=11)
    nerf = MM.init_mlp_params(8, 256, True, seed=12)
    rend, hist = MM.model_forward(prop, nerf, rays, nerf_shape=(8, 256))
    assert [h["sdist"].shape[-1] for h in hist] == [65, 65, 33]
    for h in hist:
        assert np.all(np.diff(h["sdist"], axis=-1) >= 0) and h["sdist"].min() >= 0 and h["sdist"].max() <= 1
        assert np.all(np.diff(h["tdist"], axis=-1) >= 0)
        np.testing.assert_allclose(h["weights"].sum(-1), 1, atol=1e-5)
    assert rend[-1]["rgb"].shape == (16, 3) and np.all(np.isfinite(rend[-1]["rgb"]))
    assert np.all(rend[0]["rgb"] < 1e-6)                   # the proposal levels have no colour (disable_rgb)
    assert np.all(rend[-1]["depth"] >= hist[-1]["tdist"][:, 0]) and np.all(rend[-1]["depth"] <= hist[-1]["tdist"][:, -1])


def test_oracle_reproduces_its_committed_fixture():
    """tests/golden/mip360_model.npz (oracle/gen_golden_mip360_model.py): the restatement has not drifted.  Fenceposts to 1e-6,
    densities / colours to 1e-5 relative (BLAS summation order may differ between machines)."""
    import os
    import gen_golden_mip360_model as G
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mip360_model.npz")
    gold = np.load(path)
    out, _ = G.build()
    assert set(gold.files) == set(out)
    for k in gold.files:
        if k.startswith(("ray_", "u_")):
            assert np.array_equal(gold[k], out[k]), k
        elif k.startswith(("sdist", "tdist")):
            np.testing.assert_allclose(out[k], gold[k], rtol=1e-5, atol=1e-6, err_msg=k)
        else:
            np.testing.assert_allclose(out[k], gold[k], rtol=2e-4, atol=2e-5, err_msg=k)
