"""Generate tests/golden/loader_scene.npz: the UNMODIFIED reference loader (data_loader_split.load_data_split +
RaySamplerSingleImage.get_all / random_sample) run on the synthetic on-disk scene of oracle/loader_scene.py.
CPU only, needs /root/reference.  Test infrastructure.

imageio is not installed here; the reference's ``imageio.imread`` is given a PNG reader built on cv2 (lossless format:
same pixels whatever the decoder)."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _refload import load_reference
import loader_scene
import data_loader_oracle

load_reference()
import imageio                                   # the stub module installed by load_reference (or the real one)
if not hasattr(imageio, "imread") or getattr(imageio, "__file__", None) is None:
    imageio.imread = data_loader_oracle.imread
import nerf_sample_ray_split as RS
RS.imageio = imageio
import data_loader_split as DL                   # the reference's module, unmodified
DL.imageio = imageio
assert DL.__file__.startswith("/root/reference"), DL.__file__

out = {}
with tempfile.TemporaryDirectory() as tmp:
    loader_scene.write_scene(tmp, "synth", seed=0)
    for tag, split, skip, typ in (("a", "train", 1, "gt"), ("b", "train", 2, "mono"), ("c", "test", 1, "mono")):
        samplers = DL.load_data_split(tmp, "synth", split, skip=skip, try_load_min_depth=True, depth_sup_type=typ)
        out["%s_n" % tag] = np.array(len(samplers))
        for i, s in enumerate(samplers):
            ret = s.get_all()
            for k in ("ray_o", "ray_d", "depth", "rgb", "min_depth", "depth_gt", "depth_sup"):
                out["%s_%d_%s" % (tag, i, k)] = ret[k].numpy().astype(np.float32)
            out["%s_%d_name" % (tag, i)] = np.array(os.path.basename(s.img_path))
            out["%s_%d_scale" % (tag, i)] = np.array(s.get_depth_scale(), np.float64)
        if tag == "a":
            np.random.seed(3)
            r = samplers[1].random_sample(64, center_crop=False)
            for k in ("ray_o", "ray_d", "depth", "rgb", "min_depth", "depth_gt", "depth_sup"):
                out["a_rs_%s" % k] = r[k].numpy().astype(np.float32)
np.savez_compressed(os.path.join(HERE, "..", "tests", "golden", "loader_scene.npz"), **out)
print("wrote loader_scene.npz", len(out), "arrays")
