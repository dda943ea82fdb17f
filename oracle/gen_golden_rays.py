"""Generate tests/golden/rays_tat_truck.npz with the UNMODIFIED reference ray sampler (nerf_sample_ray_split.py) on two of
the Tanks&Temples-Truck cameras the reference ships (camera_visualizer/train/cam_dict_norm.json).  CPU only, needs
/root/reference.  Test infrastructure."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _refload import load_reference, REF_ROOT

load_reference()                       # installs the stub modules (imageio ...) the sampler module imports
import nerf_sample_ray_split as RS     # the reference's module, unmodified

cams = json.load(open(os.path.join(REF_ROOT, "camera_visualizer", "train", "cam_dict_norm.json")))
out = {}
rng = np.random.RandomState(0)
for ci, name in enumerate(sorted(cams)[:2]):
    c = cams[name]
    W, H = c["img_size"]
    K = np.array(c["K"], dtype=np.float32).reshape(4, 4)
    c2w = np.linalg.inv(np.array(c["W2C"], dtype=np.float32).reshape(4, 4))
    ro, rd, dp = RS.get_rays_single_image(H, W, K, c2w)
    ids = rng.choice(H * W, size=(2048,), replace=False)
    ids[:4] = [0, W - 1, H * W - W, H * W - 1]          # the four corners
    out.update({"K%d" % ci: K, "c2w%d" % ci: c2w, "hw%d" % ci: np.array([H, W]), "ids%d" % ci: ids.astype(np.int64),
                "ray_o%d" % ci: ro[ids].astype(np.float32), "ray_d%d" % ci: rd[ids].astype(np.float32), "depth%d" % ci: dp[ids].astype(np.float32)})
np.savez_compressed(os.path.join(HERE, "..", "tests", "golden", "rays_tat_truck.npz"), **out)
print("wrote rays_tat_truck.npz", {k: v.shape for k, v in out.items()})
