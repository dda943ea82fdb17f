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
