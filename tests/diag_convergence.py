"""Diagnostic (GPU box): convergence / PSNR parity of the CUDA training path against the fp32 oracle, and forward / backward
parity on TRAINED weights.  Prints one JSON document (also written to gpurun_out/).

    python tests/diag_convergence.py [--steps 1000] [--rays 1024] [--loss mse] [--twin]
"""
import argparse
import json
import os
import sys
import tempfile
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa: F401  (puts the package and oracle/ on sys.path)
import nerfpp_oracle as O
import synth_scene
import train_harness as TH


def scaled(levels, factor, which="all"):
    out = []
    for p in levels:
        q = type(p)()
        for k, v in p.items():
            hit = k.endswith(".weight") and (which == "all" or (which == "first" and "base_layers.0.0" in k))
            q[k] = v * factor if hit else v.clone()
        out.append(q)
    return out


def forward_parity(levels, views, dev, n_views=2):
    """Per-ray relative error of the CUDA forward against the fp32 oracle on every pixel of ``n_views`` images
    (deterministic cascade), finest level.  Floors: rgb 1e-2 (an 8-bit step is 4e-3), depths 1e-3 scene units."""
    nets = TH.make_ours(levels, dev)
    from nerfpp_b200 import cascade_forward
    res = {}
    acc = {}
    for v in views[:n_views]:
        b = TH.view_batch(v, dev)
        with torch.no_grad():
            ref, far = TH.oracle_cascade([{k: t.to(dev) for k, t in p.items()} for p in levels], b, None, dev)
            # same depths for both arms at the fine level: feed the oracle's z to our forward (isolates the field + composite)
            ours = nets[-1](b["ray_o"], b["ray_d"], far, ref[-1][1].contiguous(), ref[-1][2].contiguous())
            casc, _ = cascade_forward(nets, b["ray_o"], b["ray_d"], b["min_depth"], TH.CASCADE, train=False)
        for k in ("rgb", "depth", "fg_depth", "fg_rgb", "bg_lambda"):
            acc.setdefault(k, []).append((ours[k], ref[-1][0][k]))
            acc.setdefault("cascade_" + k, []).append((casc[-1][0][k], ref[-1][0][k]))
        nonfinite = sum(int((~torch.isfinite(ours[k])).sum()) for k in ("rgb", "depth"))
        res["nonfinite"] = res.get("nonfinite", 0) + nonfinite
    for k, pairs in acc.items():
        a = torch.cat([p[0] for p in pairs])
        w = torch.cat([p[1] for p in pairs])
        floor = 1e-2 if "rgb" in k else 1e-3
        res[k] = TH.per_ray_rel(a, w, floor)
        res[k]["whole_tensor_rel"] = float((a.double() - w.double()).abs().max() / w.double().abs().max())
        res[k]["ref_absmax"] = float(w.abs().max())
    return res


def backward_parity(levels, views, dev, n_rand=1024):
    import depth_loss as DL
    img, sel, rand = TH.step_draws(123, 0, len(views), views[0]["H"] * views[0]["W"], n_rand)
    b = TH.batch_of(views[img], sel, dev)
    rand = {k: v.to(dev) for k, v in rand.items()}
    p_ref = [{k: t.detach().clone().to(dev).requires_grad_(True) for k, t in p.items()} for p in levels]
    out, far = TH.oracle_cascade(p_ref, b, rand, dev)
    nets = TH.make_ours(levels, dev)
    rep = {}
    for m in range(2):
        ret, fg_z, bg_z = out[m]
        loss = O.img2mse(ret["rgb"], b["rgb"]) + 0.1 * O.depth_mse(b["depth_sup"], ret["depth"])
        g_ref = dict(zip(p_ref[m].keys(), torch.autograd.grad(loss, list(p_ref[m].values()))))
        net = nets[m]
        net.zero_grad()
        o = net(b["ray_o"], b["ray_d"], far, fg_z.detach().contiguous(), bg_z.detach().contiguous())
        l2 = torch.mean((o["rgb"] - b["rgb"]) ** 2) + 0.1 * DL.depth_mse(b["depth_sup"], o["depth"])
        l2.backward()
        worst_cos, worst_rel, worst_norm = 1.0, 0.0, 0.0
        for name, prm in net.named_parameters():
            a, w = prm.grad.double().reshape(-1), g_ref[name].double().reshape(-1)
            if float(w.norm()) == 0.0:
                continue
            cos = float((a * w).sum() / (a.norm() * w.norm() + 1e-300))
            rel = float((a - w).abs().max() / w.abs().max())
            nr = abs(float(a.norm() / w.norm()) - 1.0)
            worst_cos, worst_rel, worst_norm = min(worst_cos, cos), max(worst_rel, rel), max(worst_norm, nr)
        rep["level%d" % m] = {"min_cos": worst_cos, "max_entry_err_over_absmax": worst_rel, "max_norm_dev": worst_norm,
                              "loss_ref": float(loss), "loss_ours": float(l2)}
    return rep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--rays", type=int, default=1024)
    ap.add_argument("--loss", default="mse")
    ap.add_argument("--dense", action="store_true", help="raise the density bias of the initial weights by 5 (well-conditioned start)")
    ap.add_argument("--quick", action="store_true", help="training + PSNR only (skip the parity tables)")
    ap.add_argument("--twin", action="store_true", help="also train a second fp32 arm from weights perturbed by 1e-6 (noise floor)")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out", "r2_convergence.json"))
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    base = tempfile.mkdtemp()
    synth_scene.write_scene(base)
    train = TH.load_views(base, "synth_learnable", "train")
    test = TH.load_views(base, "synth_learnable", "test")
    init = O.make_params_levels(2)
    if args.dense:
        init = [O.densify(p, 5.0) for p in init]
    rep = {"steps": args.steps, "rays": args.rays, "loss": args.loss, "lambda_depth": 0.1, "dense_init": bool(args.dense)}

    if not args.quick:
        rep["init_forward_parity"] = forward_parity(init, train, dev)

    t0 = time.time()
    w_ref, h_ref = TH.train_oracle(init, train, args.steps, dev, n_rand=args.rays, loss_type=args.loss, log=100)
    torch.cuda.synchronize()
    rep["oracle_train_s"] = time.time() - t0
    t0 = time.time()
    nets, h_ours = TH.train_ours(init, train, args.steps, dev, n_rand=args.rays, loss_type=args.loss, log=100)
    torch.cuda.synchronize()
    rep["ours_train_s"] = time.time() - t0
    rep["loss_first"] = {"oracle": h_ref[0], "ours": h_ours[0]}
    k = max(1, args.steps // 10)
    mean = lambda h: [sum(r[m] for r in h[-k:]) / k for m in range(2)]
    rep["loss_last_mean"] = {"oracle": mean(h_ref), "ours": mean(h_ours)}

    def score(render):
        ps, rm = [], []
        for v in test:
            rgb, depth = render(v)
            p, r = TH.psnr_rmse(rgb, depth, v)
            ps.append(p)
            rm.append(r)
        return {"psnr": ps, "rmse": rm, "psnr_mean": sum(ps) / len(ps), "rmse_mean": sum(rm) / len(rm)}

    rep["test_oracle"] = score(lambda v: TH.render_oracle(w_ref, v, dev))
    rep["test_ours"] = score(lambda v: TH.render_ours(nets, v, dev))
    rep["delta_psnr_db"] = rep["test_ours"]["psnr_mean"] - rep["test_oracle"]["psnr_mean"]
    # the oracle's weights rendered by OUR forward and vice versa: separates "training differs" from "rendering differs"
    w_ours = [{k2: v2.detach().clone() for k2, v2 in n.state_dict().items()} for n in nets]
    nets_ref = TH.make_ours(w_ref, dev)
    rep["test_oracle_weights_our_renderer"] = score(lambda v: TH.render_ours(nets_ref, v, dev))
    rep["test_our_weights_oracle_renderer"] = score(lambda v: TH.render_oracle(w_ours, v, dev))
    if args.twin:
        g = torch.Generator().manual_seed(5)
        pert = [type(p)((k2, v2 * (1.0 + 1e-6 * torch.randn(v2.shape, generator=g))) for k2, v2 in p.items()) for p in init]
        w_twin, _ = TH.train_oracle(pert, train, args.steps, dev, n_rand=args.rays, loss_type=args.loss)
        rep["test_oracle_twin"] = score(lambda v: TH.render_oracle(w_twin, v, dev))
        rep["noise_floor_db"] = rep["test_oracle_twin"]["psnr_mean"] - rep["test_oracle"]["psnr_mean"]

    rep["loss_curve_every_50"] = {"oracle": [h_ref[i] for i in range(0, len(h_ref), 50)], "ours": [h_ours[i] for i in range(0, len(h_ours), 50)]}
    if args.quick:
        s = json.dumps(rep, indent=1)
        print(s)
        open(args.out, "w").write(s)
        return
    rep["trained_forward_parity"] = forward_parity(w_ref, train, dev)
    rep["trained_backward_parity"] = backward_parity(w_ref, train, dev)
    rep["init_backward_parity"] = backward_parity(init, train, dev)
    for tag, fac, which in (("x4_first_layer", 4.0, "first"), ("x1.5_all", 1.5, "all"), ("x2_all", 2.0, "all"), ("x4_all", 4.0, "all")):
        try:
            rep["trained_%s_forward_parity" % tag] = forward_parity(scaled(w_ref, fac, which), train, dev, n_views=1)
        except Exception as e:   # noqa: BLE001
            rep["trained_%s_forward_parity" % tag] = {"error": repr(e)[:300]}
    wmax = max(float(v.abs().max()) for p in w_ref for k2, v in p.items() if k2.endswith(".weight"))
    rep["trained_weight_absmax"] = wmax
    s = json.dumps(rep, indent=1)
    print(s)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    open(args.out, "w").write(s)


if __name__ == "__main__":
    main()
