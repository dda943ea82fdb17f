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
