"""GPU parity of the mipnerf360 twins against oracle/mip360_oracle.py, through the C ABI."""
import numpy as np
import pytest
import torch

import mip360_oracle as M

pytestmark = pytest.mark.gpu


def G(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def N(t):
    return t.detach().cpu().numpy()


def test_known_answer_single_interval():
    """stepfun_test.py:579-586 through the CUDA path."""
    from nerfpp_b200 import mip360
    t = G(np.array([[1, 2, 3, 4, 5, 6]], np.float32))
    logits = G(np.array([[0, 0, 100, 0, 0]], np.float32))
    out = N(mip360.sample_intervals(None, t, logits, 10))
    np.testing.assert_allclose(out[0], np.linspace(3, 4, 11), atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize("bins,ns", [(64, 64), (64, 32), (32, 64), (7, 5)])
@pytest.mark.parametrize("jitter", [False, True])
def test_sample_intervals_matches_oracle(bins, ns, jitter):
    from nerfpp_b200 import mip360
    lv = M.synthetic_level(4096 if bins >= 32 else 37, bins, seed=bins + ns)
    u = None
    if jitter:
        g = torch.Generator(device="cuda").manual_seed(5)
        u = mip360.jittered_u((lv["t"].shape[0],), ns, False, "cuda", g)
    ref = M.sample_intervals(None if u is None else N(u), lv["t"], lv["logits"], ns, domain=(0.0, 1.0))
    got = N(mip360.sample_intervals(u, G(lv["t"]), G(lv["logits"]), ns, domain=(0.0, 1.0)))
    assert got.shape == ref.shape
    # Conditioning: a centre is t0 + (u - cw0) / (cw1 - cw0) * (t1 - t0).  With softmax weights spanning e^+-9 many bins
    # have cw1 - cw0 ~ 1e-6, where the one-ulp differences between the two fp32 cumsums (warp-shuffle scan here,
    # sequential in numpy) move the centre anywhere inside that bin -- never outside it.  So: all but a <1e-3 fraction
    # agree to 2e-5, and no fencepost is off by more than the widest bin of its ray.
    err = np.abs(got - ref)
    assert (err > 2e-5).mean() < 1e-3, (err > 2e-5).mean()
    assert np.all(err.max(-1) <= np.diff(lv["t"], axis=-1).max(-1) + 1e-6)
    assert np.all(np.diff(got, axis=-1) >= -1e-7) and got.min() >= 0 and got.max() <= 1


def test_alpha_weights_render_and_losses_match_oracle():
    from nerfpp_b200 import mip360
    for S, opaque in ((64, False), (32, True), (128, False)):
        lv = M.synthetic_level(4096, S, seed=S)
        wr, ar, tr = M.compute_alpha_weights(lv["density"], lv["t"], lv["dirs"], opaque)
        w, a, t = mip360.compute_alpha_weights(G(lv["density"]), G(lv["t"]), G(lv["dirs"]), opaque)
        # trans = exp(-cumsum): an absolute error of ~1e-6 in the fp32 sum is a relative error of ~1e-6 * sum in trans,
        # so compare with an absolute floor (weights live in [0, 1])
        np.testing.assert_allclose(N(w), wr, rtol=2e-5, atol=5e-7)
        np.testing.assert_allclose(N(a), ar, rtol=2e-5, atol=5e-7)
        np.testing.assert_allclose(N(t), tr, rtol=2e-5, atol=5e-7)
        t_far = np.full((4096, 1), 1.5, np.float32)
        rr = M.volumetric_rendering(lv["rgbs"], wr, lv["t"], np.ones(3, np.float32), t_far)
        r = mip360.volumetric_rendering(G(lv["rgbs"]), G(wr), G(lv["t"]), torch.ones(3), G(t_far))
        for k in rr:
            np.testing.assert_allclose(N(r[k]), rr[k], rtol=3e-5, atol=2e-6, err_msg=k)
        sigma = 0.01 * 0.05
        kl = float(mip360.depth_loss(G(wr), G(lv["t"]), G(lv["prior"]), r["distance_mean"], sigma, G(lv["dirs"]), "kl"))
        assert abs(kl - float(M.depth_loss_kl(wr, lv["t"], lv["prior"], sigma, lv["dirs"]))) <= 1e-5 * abs(kl) + 1e-9
        for typ, fn in (("mse", M.depth_loss_mse), ("l1", M.depth_loss_l1)):
            got = float(mip360.depth_point_loss(G(rr["distance_mean"]), G(lv["prior"]), typ))
            assert abs(got - float(fn(rr["distance_mean"], lv["prior"]))) <= 1e-5 * abs(got) + 1e-9


def test_alpha_weights_delta_and_finite():
    """render_test.py:408-463 through the CUDA path."""
    from nerfpp_b200 import mip360
    g = np.random.default_rng(0)
    n, d = 100, 128
    r = g.standard_normal((n, d))
    mask = (r == r.max(-1, keepdims=True)).astype(np.float32)
    tv = np.sort(2 * g.random((n, d + 1)) - 1, -1).astype(np.float32)
    dirs = g.standard_normal((n, 3)).astype(np.float32)
    w, a, _ = mip360.compute_alpha_weights(G(1e10 * mask), G(tv), G(dirs))
    np.testing.assert_allclose(N(w), mask, atol=1e-5, rtol=1e-5)
    np.testing.assert_allclose(N(a), mask, atol=1e-5, rtol=1e-5)
    for ldm, tlm in ((-100, -100), (-100, 10), (0, 0), (10, 10), (10, -100)):
        dens = np.exp(ldm + g.standard_normal((n, d))).astype(np.float32)
        tvs = (np.exp(tlm) * tv).astype(np.float32)
        for out in mip360.compute_alpha_weights(G(dens), G(tvs), G(dirs)):
            assert torch.isfinite(out).all()


def test_mip360_errors():
    from nerfpp_b200 import mip360, NerfppError
    with pytest.raises(ValueError):
        mip360.sample_intervals(None, torch.zeros(2, 5).cuda(), torch.zeros(2, 4).cuda(), 1)
    with pytest.raises(NerfppError):
        mip360.compute_alpha_weights(torch.zeros(2, 4), torch.zeros(2, 5), torch.zeros(2, 3))
    with pytest.raises(NotImplementedError):
        mip360.depth_loss(None, None, None, None, 1.0, None, "urf")


# ---- N4 (partial): interlevel / distortion regularisers -------------------------------------------------------------
def _torch_lossfun_outer(t, w, t_env, w_env, eps):
    """stepfun.py:30-89 restated with torch ops (autograd reference for the CUDA backward)."""
    i = torch.arange(t_env.shape[-1], device=t.device)
    v_ge_a = t[..., None, :] >= t_env[..., :, None]
    idx_lo = torch.where(v_ge_a, i[:, None], i[:1, None]).amax(-2)
    idx_hi = torch.where(~v_ge_a, i[:, None], i[-1:, None]).amin(-2)
    cy = torch.cat([torch.zeros_like(w_env[..., :1]), torch.cumsum(w_env, -1)], -1)
    w_outer = torch.gather(cy, -1, idx_hi)[..., 1:] - torch.gather(cy, -1, idx_lo)[..., :-1]
    return torch.clamp(w - w_outer, min=0) ** 2 / (w + eps)


def _torch_lossfun_distortion(t, w):
    ut = (t[..., 1:] + t[..., :-1]) / 2
    dut = (ut[..., :, None] - ut[..., None, :]).abs()
    return (w * (w[..., None, :] * dut).sum(-1)).sum(-1) + (w ** 2 * (t[..., 1:] - t[..., :-1])).sum(-1) / 3


def _histograms(n, S, Pn, seed, temp=2.0):
    rng = np.random.default_rng(seed)
    t = np.sort(rng.random((n, S + 1)), -1).astype(np.float32)
    t[:, 0], t[:, -1] = 0.0, 1.0
    te = np.sort(rng.random((n, Pn + 1)), -1).astype(np.float32)
    te[:, 0], te[:, -1] = 0.0, 1.0
    te[::7, 3] = t[::7, 2]                        # ties between a fencepost and an envelope edge
    te = np.sort(te, -1)
    w = M.softmax((temp * rng.normal(size=(n, S))).astype(np.float32))
    we = M.softmax((temp * rng.normal(size=(n, Pn))).astype(np.float32))
    return t, w, te, we


@pytest.mark.parametrize("S,Pn", [(32, 64), (64, 64), (5, 9), (200, 33)])
def test_lossfun_outer_matches_oracle_and_autograd(S, Pn):
    from nerfpp_b200 import mip360
    # Conditioning: the excess w - w_outer carries the few-ulp (of 1) differences between two fp32 cumsums, and the gradient
    # divides it by w: mildly peaked weights (temperature 0.5, w >~ 1e-3) keep that at ~1e-4 while any indexing error
    # (a range off by one bin) is an O(1) difference.
    t, w, te, we = _histograms(1000 if S <= 64 else 50, S, Pn, seed=S + Pn, temp=0.5)
    ref = M.lossfun_outer(t, w, te, we)
    we_g = G(we).requires_grad_(True)
    got = mip360.lossfun_outer(G(t), G(w), G(te), we_g)
    # w_outer is a difference of two entries of a cumsum (<= 1): absolute error ~ a few ulp of 1, squared excess / w
    np.testing.assert_allclose(N(got), ref, rtol=1e-4, atol=2e-6)
    g = torch.rand_like(got)
    (got * g).sum().backward()
    we_r = G(we).requires_grad_(True)
    (_torch_lossfun_outer(G(t), G(w), G(te), we_r, M.EPS) * g).sum().backward()
    np.testing.assert_allclose(N(we_g.grad), N(we_r.grad), rtol=1e-3, atol=5e-4)
    # self loss ~ 0 (stepfun_test.py:683-697: excess = cumsum rounding, squared) and the trainer-level wrapper
    assert float(mip360.lossfun_outer(G(t), G(w), G(t), G(w)).max()) < 1e-8
    hist = [dict(sdist=G(te), weights=we_g), dict(sdist=G(t), weights=G(w))]
    want = M.interlevel_loss(t, w, [(te, we)])
    assert abs(float(mip360.interlevel_loss(hist)) - float(want)) <= 1e-4 * abs(float(want)) + 1e-8


@pytest.mark.parametrize("S", [32, 64, 7, 130])
def test_lossfun_distortion_matches_oracle_and_autograd(S):
    from nerfpp_b200 import mip360
    t, w, _, _ = _histograms(1000 if S <= 64 else 40, S, 8, seed=S)
    ref = M.lossfun_distortion(t, w)
    tg, wg = G(t).requires_grad_(True), G(w).requires_grad_(True)
    got = mip360.lossfun_distortion(tg, wg)
    np.testing.assert_allclose(N(got), ref, rtol=1e-4, atol=1e-7)
    g = torch.rand_like(got)
    (got * g).sum().backward()
    tr, wr = G(t).requires_grad_(True), G(w).requires_grad_(True)
    (_torch_lossfun_distortion(tr, wr) * g).sum().backward()
    np.testing.assert_allclose(N(wg.grad), N(wr.grad), rtol=2e-4, atol=2e-6)
    np.testing.assert_allclose(N(tg.grad), N(tr.grad), rtol=2e-4, atol=2e-6)
    want = M.distortion_loss(t, w)
    assert abs(float(mip360.distortion_loss([dict(sdist=tg, weights=wg)])) - float(want)) <= 1e-4 * abs(float(want))


@pytest.mark.parametrize("M_bins", [64, 8, 85, 1])
def test_max_dilate_matches_oracle(M_bins):
    """stepfun.max_dilate / max_dilate_weights (stepfun.py:99-128): fenceposts bit-exact (sort + clip), values bit-exact
    without renormalisation (max, one division, one product), 1e-5 with it (one fp32 sum)."""
    from nerfpp_b200 import mip360
    rng = np.random.default_rng(M_bins)
    n = 500
    t = np.sort(rng.random((n, M_bins + 1)), -1).astype(np.float32)
    t[::5, 1:2] = t[::5, 0:1]                                   # an empty interval now and then
    w = M.softmax((2 * rng.normal(size=(n, M_bins))).astype(np.float32))
    for dil, dom in ((0.0123, (-np.inf, np.inf)), (0.05, (0.0, 1.0)), (0.5 / 64, (0.1, 0.9))):
        rt, rw = M.max_dilate(t, w, dil, dom)
        gt, gw = mip360.max_dilate(G(t), G(w), dil, dom)
        assert np.array_equal(N(gt), rt) and np.array_equal(N(gw), rw), (M_bins, dil)
        rt, rw = M.max_dilate_weights(t, w, dil, dom, renormalize=False)
        gt, gw = mip360.max_dilate_weights(G(t), G(w), dil, dom, renormalize=False)
        assert np.array_equal(N(gt), rt)
        np.testing.assert_allclose(N(gw), rw, rtol=1e-6, atol=0)
        rt, rw = M.max_dilate_weights(t, w, dil, dom, renormalize=True)
        gt, gw = mip360.max_dilate_weights(G(t), G(w), dil, dom, renormalize=True)
        np.testing.assert_allclose(N(gw), rw, rtol=1e-5, atol=1e-12)
        # sums to one wherever the dilated histogram has mass (a zero-width histogram clipped to nothing stays all zero)
        np.testing.assert_allclose(N(gw).sum(-1), rw.sum(-1), rtol=1e-5, atol=1e-12)
        assert np.all((np.abs(rw.sum(-1) - 1) < 1e-4) | (rw.sum(-1) == 0))
    with pytest.raises(Exception):
        mip360.max_dilate(G(np.zeros((2, 100), np.float32)), G(np.zeros((2, 99), np.float32)), 0.1)
