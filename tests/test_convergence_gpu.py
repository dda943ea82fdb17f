"""Convergence / PSNR parity of the CUDA training path against the fp32 oracle, and forward / backward parity on TRAINED
weights (the second half of BASELINE.json's metric: "PSNR parity").

One session-scoped fixture trains both arms for 1000 steps of 1024 rays on a synthetic on-disk scene
(oracle/synth_scene.py) from identical initial weights, pixels and uniform draws (oracle/train_harness.py) -- about 70 s
for the fp32 torch oracle and 7 s for the CUDA path on a B200 -- and the tests read from it.  Numbers measured in round 2
are recorded in profiles/r2_convergence.json."""
import os
import tempfile

import pytest
import torch

import nerfpp_oracle as O

pytestmark = pytest.mark.gpu

STEPS, RAYS = 1000, 1024


@pytest.fixture(scope="module")
def trained():
    import synth_scene
    import train_harness as TH
    dev = torch.device("cuda:0")
    base = tempfile.mkdtemp()
    synth_scene.write_scene(base)
    train = TH.load_views(base, "synth_learnable", "train")
    test = TH.load_views(base, "synth_learnable", "test")
    # the reference's initialisation with the density bias raised by 5 (the non-degenerate variant of the
    # synthetic workload).  With the plain initialisation the untrained background puts its expected depth at ~1/eps
    # (ddp_model.py:44), the first losses are ~1e8 and Adam's second moments stay poisoned for thousands of steps: both
    # arms still reach the same PSNR (measured -0.03 dB after 1000 steps, profiles/r2_convergence.json) but the
    # trajectory in between is chaotic, which makes a poor regression test.  That regime is covered by
    # test_backward_parity_along_training[init] and tests/test_trainer_gpu.py instead.
    init = [O.densify(p, 5.0) for p in O.make_params_levels(2)]
    w_ref, h_ref = TH.train_oracle(init, train, STEPS, dev, n_rand=RAYS)
    nets, h_ours = TH.train_ours(init, train, STEPS, dev, n_rand=RAYS)
    return dict(TH=TH, dev=dev, train=train, test=test, init=init, w_ref=w_ref, nets=nets, h_ref=h_ref, h_ours=h_ours)


def test_first_step_loss_matches(trained):
    """Identical weights, pixels and draws: the very first losses of the two arms agree to 1e-4."""
    for m in range(2):
        a, b = trained["h_ours"][0][m], trained["h_ref"][0][m]
        assert abs(a - b) <= 1e-4 * abs(b), (m, a, b)


def test_psnr_parity_after_training(trained):
    """Held-out views rendered with each arm's own weights and renderer after 1000 steps, where PSNR still climbs by
    ~0.7 dB per 50 steps.  Measured (profiles/r2_convergence_dense.json and two more runs): fp32 oracle 22.74 dB, CUDA
    path 22.51 / 22.31 dB (its own run-to-run spread: the fp32 atomics of the weight-gradient accumulation are not ordered),
    a second fp32 run from weights perturbed by 1e-6: 22.67 dB; depth RMSE 2.48 m (oracle), 2.09 m (CUDA), 1.87 m (twin).
    The oracle's weights rendered by the CUDA renderer give the oracle's PSNR to 0.002 dB -- the difference is the
    training trajectory, not the renderer.  From the reference's plain initialisation (chaotic start, see the fixture)
    the same comparison gave -0.03 dB at 14.4 dB.  Bars: |dPSNR| <= 0.8 dB (a lag of ~50 steps; twice the worst seen),
    depth RMSE within 50 % (the two fp32 runs differ by 25 %)."""
    TH, dev = trained["TH"], trained["dev"]
    ps, rm = {"ref": [], "ours": []}, {"ref": [], "ours": []}
    for v in trained["test"]:
        for arm, render in (("ref", lambda: TH.render_oracle(trained["w_ref"], v, dev)), ("ours", lambda: TH.render_ours(trained["nets"], v, dev))):
            rgb, depth = render()
            p, r = TH.psnr_rmse(rgb, depth, v)
            ps[arm].append(p)
            rm[arm].append(r)
    mean = lambda x: sum(x) / len(x)
    d_psnr = mean(ps["ours"]) - mean(ps["ref"])
    print("PSNR oracle %.3f dB, CUDA path %.3f dB, delta %+.3f dB; depth RMSE %.2f vs %.2f" % (
        mean(ps["ref"]), mean(ps["ours"]), d_psnr, mean(rm["ref"]), mean(rm["ours"])))
    assert mean(ps["ref"]) > 20.0 and mean(ps["ours"]) > 20.0
    assert abs(d_psnr) <= 0.8, (ps, d_psnr)
    assert abs(mean(rm["ours"]) / mean(rm["ref"]) - 1.0) <= 0.50, rm
    # the renderer alone: the oracle's weights through the CUDA path render the oracle's image
    nets_ref = TH.make_ours(trained["w_ref"], dev)
    v = trained["test"][0]
    p_cuda, _ = TH.psnr_rmse(*TH.render_ours(nets_ref, v, dev), v)
    assert abs(p_cuda - ps["ref"][0]) <= 0.02, (p_cuda, ps["ref"][0])


def _forward_pair(TH, levels, view, dev):
    nets = TH.make_ours(levels, dev)
    b = TH.view_batch(view, dev)
    with torch.no_grad():
        ref, far = TH.oracle_cascade([{k: t.to(dev) for k, t in p.items()} for p in levels], b, None, dev)
        ours = nets[-1](b["ray_o"], b["ray_d"], far, ref[-1][1].contiguous(), ref[-1][2].contiguous())
    return ours, ref[-1][0]


def test_forward_parity_per_ray_initial_weights(trained):
    """Per-ray relative error |a - b| / max(|b|, floor) of the fine level on every pixel of a view, reference's default
    initialisation: within BASELINE.json's 1e-4 for rgb (floor 1e-2) and depth (floor 1e-3) at the MAXIMUM over rays."""
    TH = trained["TH"]
    for levels in (O.make_params_levels(2), trained["init"]):       # plain and density-raised initialisation
        ours, ref = _forward_pair(TH, levels, trained["train"][0], trained["dev"])
        for k, floor in (("rgb", 1e-2), ("depth", 1e-3)):
            e = TH.per_ray_rel(ours[k], ref[k], floor)
            print(k, e)
            assert e["max"] <= 1e-4, (k, e)


def test_forward_parity_per_ray_trained_weights(trained):
    """The same on weights trained by the fp32 oracle.  Single-pass fp16 operands do NOT hold 1e-4 here at the tail:
    measured rgb p50 5e-5..9e-5 / p99 3e-4..8e-4 / max 5e-4..2e-3 depending on the run (the sharper the trained field,
    the larger) -- tests/diag_precision.py attributes it in equal parts to the rounding of inputs, weights and
    activations; only splitting all three into hi + lo halves (3x the tensor work) removes it, which is what
    impl=FIELD_TC_SPLIT does (next tests).  The bars below are what the FAST kernel is held to; nothing is non-finite."""
    TH = trained["TH"]
    ours, ref = _forward_pair(TH, trained["w_ref"], trained["train"][0], trained["dev"])
    assert all(bool(torch.isfinite(ours[k]).all()) for k in ("rgb", "depth"))
    bars = {"rgb": (1e-2, 2e-4, 1.5e-3, 4e-3), "depth": (1e-3, 4e-4, 1.5e-3, 4e-3), "fg_depth": (1e-3, 2e-4, 1.5e-3, 4e-3)}
    for k, (floor, p50, p99, mx) in bars.items():
        e = TH.per_ray_rel(ours[k], ref[k], floor)
        print(k, e)
        assert e["p50"] <= p50 and e["p99"] <= p99 and e["max"] <= mx, (k, e)


def test_forward_fp32_evaluator_on_trained_weights(trained):
    """impl=FIELD_SIMT, the plain-fp32 evaluator, meets 1e-4 per ray on the trained weights (the full-precision mode)."""
    from nerfpp_b200 import FIELD_SIMT
    TH, dev = trained["TH"], trained["dev"]
    nets = TH.make_ours(trained["w_ref"], dev)
    v = trained["train"][0]
    b = {k: t[:1024] for k, t in TH.view_batch(v, dev).items()}
    with torch.no_grad():
        ref, far = TH.oracle_cascade([{k: t.to(dev) for k, t in p.items()} for p in trained["w_ref"]], b, None, dev)
        ours = nets[-1](b["ray_o"], b["ray_d"], far, ref[-1][1].contiguous(), ref[-1][2].contiguous(), impl=FIELD_SIMT)
    for k, floor in (("rgb", 1e-2), ("depth", 1e-3)):
        e = TH.per_ray_rel(ours[k], ref[-1][0][k], floor)
        print(k, e)
        assert e["max"] <= 1e-4, (k, e)


def test_forward_split_precision_on_trained_weights(trained):
    """impl=FIELD_TC_SPLIT: the tensor-core kernel with every operand carried as a hi + lo fp16 pair (A_hi W_hi + A_lo W_hi +
    A_hi W_lo, fp32 accumulate) meets BASELINE.json's 1e-4 per ray on the trained weights -- the full-precision inference
    mode at about a third of the fast kernel's rate."""
    from nerfpp_b200 import FIELD_TC_SPLIT
    TH, dev = trained["TH"], trained["dev"]
    for levels in (trained["w_ref"], O.make_params_levels(2)):
        nets = TH.make_ours(levels, dev)
        b = TH.view_batch(trained["train"][0], dev)
        with torch.no_grad():
            ref, far = TH.oracle_cascade([{k: t.to(dev) for k, t in p.items()} for p in levels], b, None, dev)
            ours = nets[-1](b["ray_o"], b["ray_d"], far, ref[-1][1].contiguous(), ref[-1][2].contiguous(), impl=FIELD_TC_SPLIT)
        for k, floor in (("rgb", 1e-2), ("depth", 1e-3), ("fg_depth", 1e-3)):
            e = TH.per_ray_rel(ours[k], ref[-1][0][k], floor)
            print("split", k, e)
            assert e["max"] <= 1e-4, (k, e)


def test_forward_scaled_first_layer(trained):
    """Trained weights with the first layer (the one that meets the 2^9 frequency bands) scaled x4: finite, and the error
    stays within the trained-weights bars x2."""
    TH = trained["TH"]
    lv = []
    for p in trained["w_ref"]:
        q = type(p)((k, (v * 4.0 if (k.endswith("base_layers.0.0.weight")) else v.clone())) for k, v in p.items())
        lv.append(q)
    ours, ref = _forward_pair(TH, lv, trained["train"][0], trained["dev"])
    assert all(bool(torch.isfinite(ours[k]).all()) for k in ("rgb", "depth"))
    e = TH.per_ray_rel(ours["rgb"], ref["rgb"], 1e-2)
    print(e)
    assert e["max"] <= 3e-3 and e["p99"] <= 1.4e-3, e


def test_overflowing_activations_saturate_instead_of_nan(trained):
    """All weights x4: activations leave fp16's range (4^9 gain).  The epilogue's cvt.rn.satfinite clamps to +-65504 where an
    unguarded convert would produce inf and then NaN in the next layer: every output stays finite."""
    TH = trained["TH"]
    lv = [type(p)((k, (v * 4.0 if k.endswith(".weight") else v.clone())) for k, v in p.items()) for p in trained["w_ref"]]
    ours, _ = _forward_pair(TH, lv, trained["train"][0], trained["dev"])
    for k in ("rgb", "depth", "fg_weights", "bg_weights"):
        assert bool(torch.isfinite(ours[k]).all()), k


@pytest.mark.parametrize("which", ["init", "trained"])
def test_backward_parity_along_training(trained, which):
    """Gradients of the trainer's loss from the CUDA backward vs torch autograd of the fp32 oracle, same batch and draws, at
    the initial weights (loss ~1e8, the depth term 1e9 x the colour term: the case the two loss scales of backward.cu
    exist for -- with one scale the colour-path layers got all-zero gradients) and at the trained weights.  Per tensor:
    cosine >= 0.999, norm within 2 %, single entries within 5 % of the tensor's largest."""
    import depth_loss as DL
    TH, dev = trained["TH"], trained["dev"]
    levels = O.make_params_levels(2) if which == "init" else trained["w_ref"]       # "init": the plain initialisation (loss ~1e8)
    views = trained["train"]
    img, sel, rand = TH.step_draws(123, 0, len(views), views[0]["H"] * views[0]["W"], 1024)
    b = TH.batch_of(views[img], sel, dev)
    rand = {k: v.to(dev) for k, v in rand.items()}
    p_ref = [{k: t.detach().clone().to(dev).requires_grad_(True) for k, t in p.items()} for p in levels]
    out, far = TH.oracle_cascade(p_ref, b, rand, dev)
    nets = TH.make_ours(levels, dev)
    for m in range(2):
        ret, fg_z, bg_z = out[m]
        loss = O.img2mse(ret["rgb"], b["rgb"]) + 0.1 * O.depth_mse(b["depth_sup"], ret["depth"])
        g_ref = dict(zip(p_ref[m].keys(), torch.autograd.grad(loss, list(p_ref[m].values()))))
        net = nets[m]
        net.zero_grad()
        o = net(b["ray_o"], b["ray_d"], far, fg_z.detach().contiguous(), bg_z.detach().contiguous())
        l2 = torch.mean((o["rgb"] - b["rgb"]) ** 2) + 0.1 * DL.depth_mse(b["depth_sup"], o["depth"])
        l2.backward()
        for name, prm in net.named_parameters():
            a, w = prm.grad.double().reshape(-1), g_ref[name].double().reshape(-1)
            if float(w.abs().max()) < 1e-14:          # e.g. the background net behind an opaque foreground: nothing to compare
                continue
            cos = float((a * w).sum() / (a.norm() * w.norm() + 1e-300))
            assert cos >= 0.999, (which, m, name, cos)
            assert abs(float(a.norm() / w.norm()) - 1.0) <= 2e-2, (which, m, name)
            assert float((a - w).abs().max() / w.abs().max()) <= 5e-2, (which, m, name)
