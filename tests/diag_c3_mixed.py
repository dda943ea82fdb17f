"""Diagnostic (not collected): which network's operand precision the rendered error comes from -- the four combinations of
fp16 / split-precision PropMLP and NerfMLP against the oracle on 256 rays (same ordinates)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa
import mip360_model_oracle as MM, mip360_oracle as mo
from nerfpp_b200.mip360_model import Model, NerfMLP, PropMLP, Rays
dev = torch.device("cuda:0")
n = 256
rays = MM.synthetic_rays(n, seed=4)
prop, nerf = MM.init_mlp_params(4, 256, False, seed=21), MM.init_mlp_params(8, 1024, True, seed=22)
g = np.random.default_rng(9)
u_levels = []
for ns in (64, 64, 32):
    base, mj = mo.jitter_base_u(ns)
    u_levels.append((base[None, :] + g.random((n, 1)).astype(np.float32) * mj).astype(np.float32))
rend_ref, hist_ref = MM.model_forward(prop, nerf, rays, train_frac=0.5, u_levels=u_levels)
R = Rays(*(torch.from_numpy(rays[k]).to(dev) for k in ("origins", "directions", "viewdirs", "radii", "near", "far")))
for pp, pn in ((0, 0), (1, 0), (0, 1), (1, 1)):
    model = Model(dev, nerf_mlp=NerfMLP(dev, prec=bool(pn)).load(nerf), prop_mlp=PropMLP(dev, prec=bool(pp)).load(prop))
    rend, hist = model(None, R, train_frac=0.5, u_levels=[torch.from_numpy(u).to(dev) for u in u_levels])
    torch.cuda.synchronize()
    out = []
    for key, floor in (("rgb", 1e-2), ("depth", 1e-3)):
        a, r = rend[-1][key].cpu().numpy(), rend_ref[-1][key]
        err = np.abs(a - r) / np.maximum(np.abs(r), floor)
        rel = err.max(-1) if err.ndim > 1 else err
        out.append("%s max %.2e p99 %.2e p50 %.2e" % (key, rel.max(), np.percentile(rel, 99), np.median(rel)))
    sd = float(np.abs(hist[-1]["sdist"].cpu().numpy() - hist_ref[-1]["sdist"]).max())
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(5): model(None, R, train_frac=0.5, u_levels=[torch.from_numpy(u).to(dev) for u in u_levels])
    e1.record(); torch.cuda.synchronize()
    print("prop split %d nerf split %d: %s | %s | final sdist max diff %.2e | %.2f ms (256 rays)" % (pp, pn, out[0], out[1], sd, e0.elapsed_time(e1) / 5), flush=True)
