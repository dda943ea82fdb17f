"""Diagnostic (GPU box): where along a training run does the CUDA backward disagree with fp32 autograd?  Trains the fp32
oracle, snapshots its weights at a few steps, and at every snapshot compares the gradients of the CUDA path with torch
autograd of the oracle on the SAME batch and draws (per parameter tensor: cosine, norm ratio)."""
import os
import sys
import tempfile

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa: F401
import nerfpp_oracle as O
import synth_scene
import train_harness as TH
import depth_loss as DL


def grads_at(levels, views, dev, step_id):
    img, sel, rand = TH.step_draws(0, step_id, len(views), views[0]["H"] * views[0]["W"], 1024)
    b = TH.batch_of(views[img], sel, dev)
    rand = {k: v.to(dev) for k, v in rand.items()}
    p_ref = [{k: t.detach().clone().to(dev).requires_grad_(True) for k, t in p.items()} for p in levels]
    out, far = TH.oracle_cascade(p_ref, b, rand, dev)
    nets = TH.make_ours(levels, dev)
    rows = []
    for m in range(2):
        ret, fg_z, bg_z = out[m]
        l_rgb = O.img2mse(ret["rgb"], b["rgb"])
        l_d = O.depth_mse(b["depth_sup"], ret["depth"])
        loss = l_rgb + 0.1 * l_d
        g_ref = dict(zip(p_ref[m].keys(), torch.autograd.grad(loss, list(p_ref[m].values()))))
        net = nets[m]
        net.zero_grad()
        o = net(b["ray_o"], b["ray_d"], far, fg_z.detach().contiguous(), bg_z.detach().contiguous())
        l2 = torch.mean((o["rgb"] - b["rgb"]) ** 2) + 0.1 * DL.depth_mse(b["depth_sup"], o["depth"])
        l2.backward()
        worst = []
        for name, prm in net.named_parameters():
            a, w = prm.grad.double().reshape(-1), g_ref[name].double().reshape(-1)
            if float(w.norm()) == 0.0:
                continue
            cos = float((a * w).sum() / (a.norm() * w.norm() + 1e-300))
            worst.append((cos, float(a.norm() / w.norm()), name.replace("nerf_net.", "")))
        worst.sort()
        rows.append((m, float(l_rgb), float(l_d), float(l2.detach()), worst[:4], sum(c for c, _, _ in worst) / len(worst)))
    return rows


def main():
    dev = torch.device("cuda:0")
    base = tempfile.mkdtemp()
    synth_scene.write_scene(base)
    train = TH.load_views(base, "synth_learnable", "train")
    snaps = {}
    levels = O.make_params_levels(2)
    marks = [0, 10, 30, 60, 100, 150, 200, 300, 450]
    cur = levels
    done = 0
    snaps[0] = cur
    for mk in marks[1:]:
        # continue training from `cur` for (mk - done) steps with the same per-step draws the full run would see
        cur = train_segment(cur, train, done, mk, dev)
        snaps[mk] = cur
        done = mk
    for mk in marks:
        for (m, l_rgb, l_d, l2, worst, mean_cos) in grads_at(snaps[mk], train, dev, 10000 + mk):
            print("step %4d level %d  rgb_loss %.4g depth_loss %.4g ours_total %.4g  mean cos %.5f  worst: %s" % (
                mk, m, l_rgb, l_d, l2, mean_cos, "; ".join("%s cos %.4f norm x%.3f" % (n, c, r) for c, r, n in worst)), flush=True)


_OPT = {}


def train_segment(levels, views, s0, s1, dev):
    """train_oracle's loop for steps [s0, s1) keeping the Adam state across segments."""
    import collections
    if "p" not in _OPT:
        _OPT["p"] = [collections.OrderedDict((k, v.detach().clone().to(dev).requires_grad_(True)) for k, v in p.items()) for p in levels]
        _OPT["o"] = [torch.optim.Adam(list(p.values()), lr=5e-4) for p in _OPT["p"]]
    P, opts = _OPT["p"], _OPT["o"]
    npix = views[0]["H"] * views[0]["W"]
    scale = views[0]["depth_scale"]
    for step in range(s0, s1):
        img, sel, rand = TH.step_draws(0, step, len(views), npix, 1024)
        b = TH.batch_of(views[img], sel, dev)
        rand = {k: v.to(dev) for k, v in rand.items()}
        fg_far = O.intersect_sphere(b["ray_o"], b["ray_d"])
        fg_z = bg_z = ret = None
        for m, S in enumerate(TH.CASCADE):
            if m == 0:
                fg_z = O.perturb_samples(O.coarse_fg_depths(b["min_depth"], fg_far, S), rand["t_fg"])
                bg_z = O.perturb_samples(torch.linspace(0.0, 1.0, S).to(dev).view(1, S).expand(1024, S), rand["t_bg"])
            else:
                fg_z = O.resample_level(fg_z, ret["fg_weights"].detach(), rand["u_fg_%d" % m])
                bg_z = O.resample_level(bg_z, ret["bg_weights"].detach(), rand["u_bg_%d" % m])
            opts[m].zero_grad()
            ret = O.nerfpp_forward(P[m], b["ray_o"], b["ray_d"], fg_far, fg_z, bg_z)
            loss, _, _ = O.level_loss(ret, b["rgb"], b["depth_sup"], fg_z, fg_far, True, "mse", 0.1, 0.01 * scale)
            loss.backward()
            opts[m].step()
    return [collections.OrderedDict((k, v.detach().clone()) for k, v in p.items()) for p in P]


if __name__ == "__main__":
    main()
