"""Writes tests/golden/mip360_model.npz: a small seeded case of config 3's model loop evaluated by the numpy restatement
oracle/mip360_model_oracle.py (16 rays, PropMLP 4 x 256 and NerfMLP 8 x 1024 with seeded he_uniform weights, jittered ordinates).

The JAX reference cannot run in the build image (no jax / flax / gin), so -- unlike the NeRF++ fixtures, which were recorded
from the unmodified reference -- this fixture records the ORACLE: it pins the restatement against drift (a CPU test re-derives
it) and gives the GPU tests a committed target; what pins the restatement to the reference is the ported reference tests
(tests/test_mip360_model_oracle.py).

    python oracle/gen_golden_mip360_model.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import mip360_model_oracle as MM   # noqa: E402
import mip360_oracle as mo         # noqa: E402

N_RAYS, SEEDS = 16, dict(rays=4, prop=21, nerf=22, u=9)


def build():
    rays = MM.synthetic_rays(N_RAYS, seed=SEEDS["rays"])
    prop = MM.init_mlp_params(4, 256, False, seed=SEEDS["prop"])
    nerf = MM.init_mlp_params(8, 1024, True, seed=SEEDS["nerf"])
    g = np.random.default_rng(SEEDS["u"])
    u_levels = []
    for ns in (64, 64, 32):
        base, mj = mo.jitter_base_u(ns)
        u_levels.append((base[None, :] + g.random((N_RAYS, 1)).astype(np.float32) * mj).astype(np.float32))
    rend, hist = MM.model_forward(prop, nerf, rays, train_frac=0.5, u_levels=u_levels)
    out = {"ray_" + k: v for k, v in rays.items()}
    for i, u in enumerate(u_levels):
        out["u_%d" % i] = u
    for i, (r, h) in enumerate(zip(rend, hist)):
        out["sdist_%d" % i], out["tdist_%d" % i] = h["sdist"], h["tdist"]
        out["density_%d" % i], out["weights_%d" % i] = h["density"], h["weights"]
        out["rgb_%d" % i], out["depth_%d" % i], out["acc_%d" % i] = r["rgb"], r["depth"], r["acc"]
    out["rgb_samples_2"] = hist[2]["rgb"]
    return out, (prop, nerf)


if __name__ == "__main__":
    out, _ = build()
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "mip360_model.npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in out.items()})
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1e3))
