"""Generate tests/golden/nerfpp_*.npz by running the UNMODIFIED reference code in this container.

    python oracle/gen_golden.py            # needs /root/reference (read-only), CPU only

Test infrastructure.  The reference's RNG draws (``torch.rand_like`` ddp_train_nerf.py:75,
``torch.rand`` :107) are intercepted at the torch API boundary so that the recorded ``u``/``t_rand``
are exactly what the reference consumed; ``torch.gather`` is wrapped to record the reference's own
inverse-CDF indices (``inds_g`` :115-121) and ``cdf``.  No reference source is modified or copied.
"""
import hashlib
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import nerfpp_oracle as O   # only for the synthetic-ray generator (inputs, not outputs)
from _refload import load_reference

OUT = os.path.join(HERE, "..", "tests", "golden")


class Intercept:
    """Feeds prescribed random tensors to the reference and records gather() calls."""

    def __init__(self):
        self.queue, self.gathers = [], []
        self._rand, self._rand_like, self._gather = torch.rand, torch.rand_like, torch.gather

    def __enter__(self):
        def rand(*shape, **kw):
            t = self.queue.pop(0)
            assert tuple(t.shape) == tuple(shape), (t.shape, shape)
            return t.clone()

        def rand_like(x, **kw):
            t = self.queue.pop(0)
            assert t.shape == x.shape
            return t.clone()

        def gather(input, dim, index, **kw):
            self.gathers.append((input.detach().clone(), index.detach().clone()))
            return self._gather(input=input, dim=dim, index=index, **kw)

        torch.rand, torch.rand_like, torch.gather = rand, rand_like, gather
        return self

    def __exit__(self, *a):
        torch.rand, torch.rand_like, torch.gather = self._rand, self._rand_like, self._gather


def np_(t):
    return t.detach().cpu().numpy()


def param_digest(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(np_(sd[k]).astype(np.float32).tobytes())
    return h.hexdigest()


def build_reference_nets(ref_model, n_levels, sigma_bias):
    args = SimpleNamespace(max_freq_log2=10, max_freq_log2_viewdirs=4, netdepth=8, netwidth=256,
                           use_viewdirs=True)
    torch.manual_seed(777)                       # ddp_train_nerf.py:308
    nets = [ref_model.NerfNetWithAutoExpo(args) for _ in range(n_levels)]   # :315-322
    digests = [param_digest(n.state_dict()) for n in nets]
    if sigma_bias:
        with torch.no_grad():
            for n in nets:
                n.nerf_net.fg_net.sigma_layers[0].bias += sigma_bias
                n.nerf_net.bg_net.sigma_layers[0].bias += sigma_bias
    return nets, digests


def run_case(name, n_rays, cascade, train, sigma_bias, ray_seed, rand_seed, with_grads=False):
    T, M, D, U = load_reference()
    nets, digests = build_reference_nets(M, len(cascade), sigma_bias)
    rays = O.synthetic_rays(n_rays, seed=ray_seed)
    rnd = O.synthetic_rand(n_rays, cascade, seed=rand_seed) if train else None
    ray_o, ray_d, near = rays["ray_o"], rays["ray_d"], rays["min_depth"]
    depth_sup, rgb_gt = rays["depth_sup"], rays["rgb"]
    g = {"ray_o": ray_o, "ray_d": ray_d, "min_depth": near, "rgb_gt": rgb_gt, "depth_sup": depth_sup}
    meta = dict(cascade=list(cascade), train=int(train), sigma_bias=float(sigma_bias),
                ray_seed=ray_seed, rand_seed=rand_seed, depth_scale=rays["depth_scale"], depth_sigma=0.01)
    if rnd:
        for k, v in rnd.items():
            g["rand_" + k] = v

    # ---- the cascade loop, calling the reference's functions exactly as ddp_train_nerf.py:432-468 /
    # :162-208 wires them
    ret = None
    for m, S in enumerate(cascade):
        if m == 0:
            fg_far = T.intersect_sphere(ray_o, ray_d)
            step = (fg_far - near) / (S - 1)
            fg_depth = torch.stack([near + i * step for i in range(S)], dim=-1)
            bg_depth = torch.linspace(0., 1., S).view(1, S).expand(n_rays, S)
            g["fg_far"] = fg_far
            g["fg_z_grid"] = fg_depth
            if train:
                with Intercept() as ic:
                    ic.queue = [rnd["t_fg"], rnd["t_bg"]]
                    fg_depth = T.perturb_samples(fg_depth)
                    bg_depth = T.perturb_samples(bg_depth)
        else:
            fg_w = ret["fg_weights"].clone().detach()[..., 1:-1]
            fg_mid = .5 * (fg_depth[..., 1:] + fg_depth[..., :-1])
            bg_w = ret["bg_weights"].clone().detach()[..., 1:-1]
            bg_mid = .5 * (bg_depth[..., 1:] + bg_depth[..., :-1])
            with Intercept() as ic:
                if train:
                    ic.queue = [rnd["u_fg_%d" % m], rnd["u_bg_%d" % m]]
                fg_new = T.sample_pdf(bins=fg_mid, weights=fg_w, N_samples=S, det=not train)
                n_fg = len(ic.gathers)
                bg_new = T.sample_pdf(bins=bg_mid, weights=bg_w, N_samples=S, det=not train)
            # gathers: [cdf_g, bins_g] per call; cdf input is the expanded [N,Ns,M+1] view
            (cdf_e, inds), _ = ic.gathers[0], ic.gathers[1]
            g["fg_cdf_%d" % m] = cdf_e[:, 0, :]
            g["fg_inds_%d" % m] = inds[..., 1].to(torch.int32)         # above_inds
            (cdf_e, inds) = ic.gathers[n_fg]
            g["bg_cdf_%d" % m] = cdf_e[:, 0, :]
            g["bg_inds_%d" % m] = inds[..., 1].to(torch.int32)
            g["fg_new_%d" % m], g["bg_new_%d" % m] = fg_new, bg_new
            fg_depth, _ = torch.sort(torch.cat((fg_depth, fg_new), dim=-1))
            bg_depth, _ = torch.sort(torch.cat((bg_depth, bg_new), dim=-1))
        g["fg_z_%d" % m], g["bg_z_%d" % m] = fg_depth, bg_depth
        net = nets[m]
        if with_grads:
            net.zero_grad()
            ret = net(ray_o, ray_d, fg_far, fg_depth, bg_depth, img_name=None)
        else:
            with torch.no_grad():
                ret = net(ray_o, ray_d, fg_far, fg_depth, bg_depth)
        for k, v in ret.items():
            g["ret%d_%s" % (m, k)] = v
        # losses (ddp_train_nerf.py:481-491)
        g["loss%d_rgb" % m] = U.img2mse(ret["rgb"], rgb_gt)
        g["loss%d_mse" % m] = D.depth_mse(depth_sup, ret["depth"])
        g["loss%d_l1" % m] = D.depth_l1(depth_sup, ret["depth"])
        sig = meta["depth_sigma"] * meta["depth_scale"]
        g["loss%d_kl" % m] = D.depth_kl(ret["fg_weights"], depth_sup, fg_depth, ret["fg_dists"], sig, fg_far)
        if with_grads:
            # three independent backward passes: rgb + 0.1*{mse,l1,kl}
            for lt in ("mse", "l1", "kl"):
                net.zero_grad()
                loss = U.img2mse(ret["rgb"], rgb_gt) + 0.1 * g["loss%d_%s" % (m, lt)]
                loss.backward(retain_graph=True)
                for pname, p in net.named_parameters():
                    gr = p.grad.detach()
                    short = pname.replace("nerf_net.", "").replace("_layers", "").replace(".weight", ".w").replace(".bias", ".b")
                    g["grad%d_%s_norm/%s" % (m, lt, short)] = gr.norm()
                    g["grad%d_%s_head/%s" % (m, lt, short)] = gr.reshape(-1)[:16].clone()
    out = {k: (np_(v) if torch.is_tensor(v) else np.asarray(v)) for k, v in g.items()}
    for k, v in meta.items():
        out["meta_" + k] = np.asarray(v)
    out["meta_param_digest"] = np.asarray(digests)
    path = os.path.join(OUT, "nerfpp_%s.npz" % name)
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024), "keys", len(out))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    # C1: plumbing shape -- one coarse level of 64, test-time (det, no perturb), untrained density
    run_case("c1_coarse_det", n_rays=32, cascade=(64,), train=False, sigma_bias=0.0, ray_seed=11, rand_seed=0)
    # C2: headline shape -- (64, +128), training path with recorded random draws, dense variant + grads
    run_case("c2_train_dense", n_rays=12, cascade=(64, 128), train=True, sigma_bias=5.0, ray_seed=12, rand_seed=21,
             with_grads=True)
    # C2 at init density (weights ~ 1e-3: stresses the +1e-6 terms and tiny-denominator branch)
    run_case("c2_train_init", n_rays=12, cascade=(64, 128), train=True, sigma_bias=0.0, ray_seed=13, rand_seed=22)
    # test-time cascade (det=True, linspace u): the render_single_image path
    run_case("c2_det_dense", n_rays=12, cascade=(64, 128), train=False, sigma_bias=5.0, ray_seed=14, rand_seed=0)
