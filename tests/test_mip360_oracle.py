"""Pins oracle/mip360_oracle.py with the reference's own known-answer / property tests for the mipnerf360 twins
(nerf-methods/mipnerf360/tests/, ported from absltest+JAX to numpy)."""
import numpy as np
import pytest

import mip360_oracle as M


def test_sample_single_interval_is_linspace():
    """stepfun_test.py:579-586 -- known answer."""
    t = np.array([1, 2, 3, 4, 5, 6], np.float32)
    logits = np.array([0, 0, 100, 0, 0], np.float32)
    out = M.sample_intervals(None, t, logits, 10)
    np.testing.assert_allclose(out, np.linspace(3, 4, 11), atol=1e-5, rtol=1e-5)


def test_sorted_interp_equals_interp():
    """math_test.py:157-180."""
    g = np.random.default_rng(0)
    n, d0, d1 = 100, 10, 20
    x = g.standard_normal((n, d0)).astype(np.float32)
    xp = np.sort(g.standard_normal((n, d1)).astype(np.float32), -1)
    fp = np.sort(g.standard_normal((n, d1)).astype(np.float32), -1)
    z = M.sorted_interp(x, xp, fp)
    z_true = np.stack([np.interp(x[i], xp[i], fp[i]) for i in range(n)])
    np.testing.assert_allclose(z, z_true, atol=1e-5, rtol=1e-5)


def test_alpha_weights_delta_correct():
    """render_test.py:443-463 -- a single interval with a huge density gives one-hot weights/alpha."""
    g = np.random.default_rng(0)
    n, d = 100, 128
    r = g.standard_normal((n, d))
    mask = r == r.max(-1, keepdims=True)
    density = (1e10 * mask).astype(np.float32)
    tvals = np.sort(2 * g.random((n, d + 1)) - 1, -1).astype(np.float32)
    dirs = g.standard_normal((n, 3)).astype(np.float32)
    w, a, _ = M.compute_alpha_weights(density, tvals, dirs)
    np.testing.assert_allclose(mask.astype(np.float32), w, atol=1e-5, rtol=1e-5)
    np.testing.assert_allclose(mask.astype(np.float32), a, atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize("ldm,tlm", [(-100, -100), (-100, 10), (-10, 0), (0, 0), (0, 10), (10, -10), (10, 10), (10, -100)])
def test_alpha_weights_finite(ldm, tlm):
    """render_test.py:408-441 (values; gradients are a JAX-only check)."""
    g = np.random.default_rng(0)
    n, d = 100, 128
    density = np.exp(ldm + g.standard_normal((n, d))).astype(np.float32)
    tvals = (np.exp(tlm) * np.sort(2 * g.random((n, d + 1)) - 1, -1)).astype(np.float32)
    dirs = g.standard_normal((n, 3)).astype(np.float32)
    for out in M.compute_alpha_weights(density, tvals, dirs):
        assert np.all(np.isfinite(out))


def test_integrate_weights_ends_and_percentile():
    """stepfun.py:131-150 contract (exact 0 / 1 ends) and weighted_percentile == interp (stepfun_test.py:739-790)."""
    g = np.random.default_rng(1)
    w = M.softmax(g.standard_normal((50, 17)).astype(np.float32))
    cw = M.integrate_weights(w)
    assert np.all(cw[:, 0] == 0) and np.all(cw[:, -1] == 1) and np.all(np.diff(cw, axis=-1) >= 0)
    t = np.sort(g.random((50, 18)).astype(np.float32), -1)
    ps = [5, 50, 95]
    got = M.weighted_percentile(t, w, ps)
    ref = np.stack([np.interp(np.array(ps) / 100, np.concatenate([[0], np.minimum(1, np.cumsum(w[i][:-1])), [1]]), t[i]) for i in range(50)])
    np.testing.assert_allclose(got, ref, atol=1e-6)


def test_sample_intervals_flat_and_sparse():
    """stepfun_test.py:385-495 in spirit: a flat PDF resamples to (nearly) uniform fenceposts; a PDF with zero-weight
    bins never places a centre inside them; outputs are sorted and stay within the domain."""
    t = np.linspace(0, 1, 65, dtype=np.float32)[None]
    out = M.sample_intervals(None, t, np.zeros((1, 64), np.float32), 32, domain=(0, 1))
    np.testing.assert_allclose(np.diff(out[0][1:-1]), 1 / 32, atol=1e-4)
    logits = np.full((1, 64), -1e9, np.float32)
    logits[0, 10:20] = 0
    out = M.sample_intervals(None, t, logits, 32, domain=(0, 1))
    assert np.all(np.diff(out) >= 0) and out.min() >= t[0, 10] - 1e-6 and out.max() <= t[0, 20] + 1e-6


def test_volumetric_rendering_and_losses_shapes():
    lv = M.synthetic_level(64, 32, seed=3)
    w, _, _ = M.compute_alpha_weights(lv["density"], lv["t"], lv["dirs"])
    r = M.volumetric_rendering(lv["rgbs"], w, lv["t"], np.ones(3, np.float32), np.ones((64, 1), np.float32))
    assert r["rgb"].shape == (64, 3) and np.all(r["depth"] >= lv["t"][:, 0]) and np.all(r["depth"] <= lv["t"][:, -1])
    assert np.all(r["distance_percentile_5"] <= r["distance_median"] + 1e-6) and np.all(r["distance_median"] <= r["distance_percentile_95"] + 1e-6)
    np.testing.assert_allclose(r["acc"] + np.maximum(0, 1 - r["acc"]), np.maximum(1, r["acc"]), rtol=1e-6)
    assert np.isfinite(M.depth_loss_kl(w, lv["t"], lv["prior"], 0.01 * 0.05, lv["dirs"]))
    assert M.depth_loss_mse(r["distance_mean"], lv["prior"]) >= 0 and M.depth_loss_l1(r["distance_mean"], lv["prior"]) >= 0


# ---- N4: the interlevel / distortion regularisers, pinned by the reference's own property tests ----------------------
def _brute_outer(t0, t1, w1):        # stepfun_test.py:40-50
    return np.array([sum(w1[j] for j in range(len(t1) - 1) if t1[j + 1] >= t0[i] and t1[j] <= t0[i + 1]) for i in range(len(t0) - 1)])


def _brute_inner(t0, t1, w1):        # stepfun_test.py:27-37
    return np.array([sum(w1[j] for j in range(len(t1) - 1) if t1[j] >= t0[i] and t1[j + 1] < t0[i + 1]) for i in range(len(t0) - 1)])


@pytest.mark.parametrize("num_ablate,is_all_zero", [(0, True), (2, False)])
def test_lossfun_outer_same_and_ablated_point_sets(num_ablate, is_all_zero):
    """stepfun_test.py:588-622: histograms of the same points bound each other (loss 0); drop points from one and it does not."""
    rng = np.random.default_rng(0)
    all_zero = True
    for _ in range(10):
        num_pts, d0, d1 = rng.integers(10, 20, 3)
        t0, t1 = np.sort(rng.random(d0 + 1)).astype(np.float32), np.sort(rng.random(d1 + 1)).astype(np.float32)
        lo, hi = max(t0.min(), t1.min()) + 0.1, min(t0.max(), t1.max()) - 0.1
        pts = rng.uniform(lo, hi, num_pts)
        pts_ablate = pts[:-num_ablate] if num_ablate > 0 else pts
        w0 = np.array([np.mean((pts_ablate >= t0[i]) & (pts_ablate < t0[i + 1])) for i in range(d0)], np.float32)
        w1 = np.array([np.mean((pts >= t1[i]) & (pts < t1[i + 1])) for i in range(d1)], np.float32)
        all_zero &= bool(np.all(M.lossfun_outer(t0, w0, t1, w1) < 1e-12))
    assert all_zero == is_all_zero


def test_inner_outer_bound_and_match_brute_force():
    """stepfun_test.py:624-655 (inner <= w <= outer for histograms of the same points) and :699-735 (brute-force measures)."""
    rng = np.random.default_rng(4)
    for _ in range(10):
        d0, d1, num_pts = rng.integers(10, 20, 3)
        t0, t1 = np.sort(rng.random(d0 + 1)).astype(np.float32), np.sort(rng.random(d1 + 1)).astype(np.float32)
        lo, hi = max(t0.min(), t1.min()) + 0.1, min(t0.max(), t1.max()) - 0.1
        pts = rng.uniform(lo, hi, num_pts)
        w0 = np.array([np.sum((pts >= t0[i]) & (pts < t0[i + 1])) for i in range(d0)], np.float32)
        w1 = np.array([np.sum((pts >= t1[i]) & (pts < t1[i + 1])) for i in range(d1)], np.float32)
        w0_in, w0_out = M.inner_outer(t0, t1, w1)
        w1_in, w1_out = M.inner_outer(t1, t0, w0)
        assert np.all(w0_in <= w0) and np.all(w0 <= w0_out) and np.all(w1_in <= w1) and np.all(w1 <= w1_out)
        wr = np.exp(rng.normal(size=d0)).astype(np.float32)
        inn, out = M.inner_outer(t1, t0, wr)
        np.testing.assert_allclose(out, _brute_outer(t1, t0, wr), atol=1e-5, rtol=1e-5)
        np.testing.assert_allclose(inn, _brute_inner(t1, t0, wr), atol=1e-5, rtol=1e-5)


def test_lossfun_outer_monotonic_invariance_and_self_zero():
    """stepfun_test.py:657-697."""
    rng = np.random.default_rng(0)
    for _ in range(10):
        d0, d1 = rng.integers(10, 20, 2)
        t0, t1 = np.sort(rng.random(d0 + 1)).astype(np.float32), np.sort(rng.random(d1 + 1)).astype(np.float32)
        w0, w1 = np.exp(rng.normal(size=d0)).astype(np.float32), np.exp(rng.normal(size=d1)).astype(np.float32)
        curve = lambda x: (1 + x ** 3).astype(np.float32)
        assert np.array_equal(M.lossfun_outer(t0, w0, t1, w1), M.lossfun_outer(curve(t0), w0, curve(t1), w1))
        assert np.all(M.lossfun_outer(t0, w0, t0, w0) < 1e-10)


def test_distortion_loss_matches_interval_distortion():
    """stepfun_test.py:227-275: interval_distortion vs brute force, and lossfun_distortion == sum_ij w_i w_j d_ij."""
    rng = np.random.default_rng(0)
    n, d = 3, 7
    t0 = np.sort(rng.uniform(-3, 3, (n, d + 1)), -1)
    t1 = np.sort(rng.uniform(-3, 3, (n, d + 1)), -1)
    dist = M.interval_distortion(t0[..., :-1], t0[..., 1:], t1[..., :-1], t1[..., 1:])
    brute = np.zeros_like(dist)
    for i in range(n):
        for j in range(d):
            brute[i, j] = np.mean(np.abs(np.linspace(t0[i, j], t0[i, j + 1], 2001)[:, None] - np.linspace(t1[i, j], t1[i, j + 1], 2001)[None, :]))
    np.testing.assert_allclose(dist, brute, atol=1e-6, rtol=1e-3)
    n, d = 3, 8
    t = np.sort(rng.uniform(-3, 3, (n, d + 1)), -1).astype(np.float32)
    w = M.softmax((2 * rng.normal(size=(n, d))).astype(np.float32))
    losses = M.lossfun_distortion(t, w)
    dd = M.interval_distortion(t[..., :-1, None], t[..., 1:, None], t[..., None, :-1], t[..., None, 1:])
    alt = np.sum(w[:, None, :] * w[:, :, None] * dd, axis=(-1, -2))
    np.testing.assert_allclose(losses, alt, atol=1e-6, rtol=1e-4)


def test_max_dilate_against_brute_force_queries():
    """stepfun_test.py:277-300: a query of the dilated step function is the max of the original over +-dilation."""
    rng = np.random.default_rng(0)
    n, d, dilation = 20, 8, 0.53
    t = (np.cumsum(rng.integers(1, 10, (n, d + 1)), -1) / 10).astype(np.float32)
    w = M.softmax(rng.normal(size=(n, d)).astype(np.float32))
    td, wd = M.max_dilate(t, w, dilation)
    assert td.shape == (n, 3 * d + 1) and wd.shape == (n, 3 * d)
    tq = ((np.arange((d + 4) * 10) - 2.5) / 10).astype(np.float32)
    wq = M.query(np.broadcast_to(tq, (n, tq.size)), t, w)
    wdq = M.query(np.broadcast_to(tq, (n, tq.size)), td, wd)
    mask = np.abs(tq[None, :] - tq[:, None]) <= dilation
    for i in range(n):
        np.testing.assert_array_equal(wdq[i], np.max(mask * wq[i], axis=-1))
    # weights variant: renormalised output sums to one, the domain clips the fenceposts
    td, wd = M.max_dilate_weights(t, w, 0.2, domain=(0.5, 4.0), renormalize=True)
    np.testing.assert_allclose(wd.sum(-1), 1.0, rtol=1e-5)
    assert td.min() >= 0.5 and td.max() <= 4.0 and np.all(np.diff(td, axis=-1) >= 0) and np.all(wd >= 0)


def test_searchsorted_contract():
    """stepfun_test.py:55-146: a[lo] <= v < a[hi] inside the range, both indices clamp to the first / last edge outside
    it, and idx_hi equals numpy's searchsorted(side='right') for in-range queries."""
    rng = np.random.default_rng(0)
    for _ in range(10):
        n, m = rng.integers(10, 100, 2)
        v = rng.uniform(1e-7, 1 - 1e-7, n).astype(np.float32)
        a = np.sort(np.concatenate([[0.0, 1.0], rng.random(m)])).astype(np.float32)
        lo, hi = M.searchsorted(a, v)
        assert np.all(a[lo] <= v) and np.all(v < a[hi])
        assert np.array_equal(hi, np.searchsorted(a, v, side="right"))
        below, above = (a[0] - 0.1 - rng.random(5)).astype(np.float32), (a[-1] + 0.1 + rng.random(5)).astype(np.float32)
        lo, hi = M.searchsorted(a, below)
        assert np.all(lo == 0) and np.all(hi == 0)
        lo, hi = M.searchsorted(a, above)
        assert np.all(lo == len(a) - 1) and np.all(hi == len(a) - 1)
