"""Diagnostic (GPU box): WHICH fp16 roundings of the tensor-core field cost accuracy on trained weights?  Emulates the
kernel's arithmetic in torch (fp32 accumulate) with selectable roundings -- inputs (E operand), weights, activations --
and reports the per-ray relative error of the rendered rgb / depth against the unrounded fp32 oracle."""
import json
import os
import sys
import tempfile

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa: F401
import nerfpp_oracle as O
import synth_scene
import train_harness as TH


def h16(x):
    return x.half().float()


def split2(x):   # hi + lo fp16 pair (22 bits)
    hi = h16(x)
    return hi + h16(x - hi)


def mlp_q(params, net, pos_emb, view_emb, qe, qw, qa):
    """nerf_network.py:120-142 with rounding hooks: qe on the encodings, qw on weights, qa on hidden activations."""
    def lin(name, x, wq=True):
        p = "nerf_net.%s.%s" % (net, name)
        w = params[p + ".weight"]
        return torch.nn.functional.linear(x, qw(w) if wq else w, params[p + ".bias"])
    pe, ve = qe(pos_emb), qe(view_emb)
    h = qa(torch.relu(lin("base_layers.0.0", pe)))
    for i in range(7):
        x = torch.cat((pe, h), -1) if i == 4 else h
        h = torch.relu(lin("base_layers.%d.0" % (i + 1), x))
        hq = qa(h)
        if i == 6:
            sigma = torch.abs(lin("sigma_layers.0", h, wq=False)).squeeze(-1)      # fp32 dot product on the fp32 accumulators
        h = hq
    remap = qa(lin("base_remap_layers.0", h))
    c = torch.relu(lin("rgb_layers.0", torch.cat((remap, ve), -1)))
    rgb = torch.sigmoid(lin("rgb_layers.2", c, wq=False))
    return rgb, sigma


def forward_q(params, b, far, fg_z, bg_z, qe, qw, qa):
    ray_o, ray_d = b["ray_o"], b["ray_d"]
    d_norm = torch.norm(ray_d, dim=-1, keepdim=True)
    viewdir = ray_d / d_norm
    N, S = fg_z.shape
    o, d, v = (t[:, None, :].expand(N, S, 3) for t in (ray_o, ray_d, viewdir))
    pts = o + fg_z[..., None] * d
    fg_rgb, fg_sigma = mlp_q(params, "fg_net", O.posenc(pts, 10), O.posenc(v, 4), qe, qw, qa)
    bg_pts, bg_dr = O.inverted_sphere_points(o, d, bg_z)
    pe = torch.flip(O.posenc(bg_pts, 10), dims=[-2])
    ve = torch.flip(O.posenc(v, 4), dims=[-2])
    bg_rgb, bg_sigma = mlp_q(params, "bg_net", pe, ve, qe, qw, qa)
    return O.composite(fg_sigma, fg_rgb, bg_sigma, bg_rgb, torch.flip(bg_dr, dims=[-1]), ray_d, far, fg_z, bg_z)


def main():
    dev = torch.device("cuda:0")
    steps = int(os.environ.get("STEPS", 600))
    base = tempfile.mkdtemp()
    synth_scene.write_scene(base)
    train = TH.load_views(base, "synth_learnable", "train")
    w_ref, _ = TH.train_oracle(O.make_params_levels(2), train, steps, dev, n_rand=1024, log=200)
    ident = lambda x: x
    variants = {"all_fp16 (the kernel)": (h16, h16, h16), "inputs only": (h16, ident, ident), "weights only": (ident, h16, ident),
                "activations only": (ident, ident, h16), "weights+acts fp16, inputs hi+lo": (split2, h16, h16),
                "inputs+acts fp16, weights hi+lo": (h16, split2, h16), "inputs+weights fp16, acts hi+lo": (h16, h16, split2),
                "acts fp16, inputs+weights hi+lo": (split2, split2, h16), "all hi+lo": (split2, split2, split2)}
    rep = {"steps": steps}
    with torch.no_grad():
        b = TH.view_batch(train[0], dev)
        levels = [{k: t.to(dev) for k, t in p.items()} for p in w_ref]
        ref, far = TH.oracle_cascade(levels, b, None, dev)
        fg_z, bg_z = ref[-1][1], ref[-1][2]
        want = ref[-1][0]
        for name, (qe, qw, qa) in variants.items():
            got = forward_q(levels[-1], b, far, fg_z, bg_z, qe, qw, qa)
            rep[name] = {k: TH.per_ray_rel(got[k], want[k], 1e-2 if "rgb" in k else 1e-3) for k in ("rgb", "depth", "fg_depth")}
        # and the kernel itself
        nets = TH.make_ours(w_ref, dev)
        ours = nets[-1](b["ray_o"], b["ray_d"], far, fg_z.contiguous(), bg_z.contiguous())
        rep["KERNEL"] = {k: TH.per_ray_rel(ours[k], want[k], 1e-2 if "rgb" in k else 1e-3) for k in ("rgb", "depth", "fg_depth")}
    for k, v in rep.items():
        if isinstance(v, dict):
            print("%-40s rgb max %.2e p99 %.2e | depth max %.2e p99 %.2e | fg_depth max %.2e" % (
                k, v["rgb"]["max"], v["rgb"]["p99"], v["depth"]["max"], v["depth"]["p99"], v["fg_depth"]["max"]))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out", "r2_precision.json")
    open(out, "w").write(json.dumps(rep, indent=1))


if __name__ == "__main__":
    main()
