"""CPU restatement (numpy) of the reference's on-disk loader for SURVEY.md section 8(f) N2 -- test infrastructure only.

Follows data_loader_split.py:27-129 (file discovery, sorted order, ``skip``, the ``scale`` file, ``max_depth.txt``) and
RaySamplerSingleImage.set_resolution_level (nerf_sample_ray_split.py:65-106) at resolution_level 1.  Pinned against the
unmodified reference by tests/golden/loader_scene.npz (oracle/gen_golden_loader.py)."""
import glob
import os

import cv2
import numpy as np

from nerfpp_oracle import get_rays_single_image


def imread(path):
    """imageio.imread's result for the PNGs of the format: HxWx3 uint8 RGB, HxW uint8 / uint16 grey."""
    a = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if a is None:
        raise IOError(path)
    if a.ndim == 3:
        a = a[:, :, ::-1] if a.shape[2] == 3 else a[:, :, [2, 1, 0, 3]]
    return np.ascontiguousarray(a)


def find_files(d, exts):                                     # data_loader_split.py:14-24
    if not os.path.isdir(d):
        return []
    out = []
    for e in exts:
        out.extend(glob.glob(os.path.join(d, e)))
    return sorted(out)


def parse_txt(path):                                         # :29-32
    return np.array([float(x) for x in open(path).read().split()]).reshape(4, 4).astype(np.float32)


def load_data_split(basedir, scene, split, skip=1, try_load_min_depth=True, depth_sup_type="gt"):
    """Returns a list of dicts (one per camera) with what RaySamplerSingleImage.get_all() returns, as numpy arrays."""
    basedir = basedir.rstrip("/")
    sd = "%s/%s/%s" % (basedir, scene, split)
    img_ext = ["*.png", "*.jpg"]
    intr = find_files(sd + "/intrinsics", ["*.txt"])[::skip]
    pose = find_files(sd + "/pose", ["*.txt"])[::skip]
    n = len(pose)
    pick = lambda files: files[::skip] if len(files) > 0 else [None] * n
    img = pick(find_files(sd + "/rgb", img_ext))
    mask = pick(find_files(sd + "/mask", img_ext))
    mind = pick(find_files(sd + "/min_depth", img_ext)) if try_load_min_depth else [None] * n
    dgt = pick(find_files(sd + "/depth", img_ext))
    scale = float(open(os.path.join(basedir, scene, "scale")).readlines()[0].strip()) if dgt[0] is not None else None   # :86
    suffix = "_" + depth_sup_type if depth_sup_type != "gt" else ""                                                      # :91
    dsup = pick(find_files(sd + "/depth" + suffix, img_ext))
    H, W = imread(find_files("%s/%s/train/rgb" % (basedir, scene), img_ext)[0]).shape[:2]                               # :102-104
    try:
        max_depth = float(open(sd + "/max_depth.txt").readline().strip())                                                # :113-116
    except Exception:
        max_depth = None
    out = []
    for i in range(n):
        K, c2w = parse_txt(intr[i]), parse_txt(pose[i])
        ro, rd, dp = get_rays_single_image(H, W, K, c2w)
        d = dict(H=H, W=W, intrinsics=K, c2w=c2w, ray_o=ro, ray_d=rd, depth=dp, depth_scale=scale, img_path=img[i])
        if img[i] is not None:
            d["rgb"] = (imread(img[i]).astype(np.float32) / 255.).reshape(-1, 3)                                        # :73-75
        if mask[i] is not None:
            d["mask"] = (imread(mask[i]).astype(np.float32) / 255.).reshape(-1)                                          # :80-82
        if mind[i] is not None:
            d["min_depth"] = (imread(mind[i]).astype(np.float32) / 255. * max_depth + 1e-4).reshape(-1)                  # :87-89
        else:
            d["min_depth"] = 1e-4 * np.ones_like(rd[..., 0])                                                             # :135
        if dgt[i] is not None:
            d["depth_gt"] = scale * (np.array(imread(dgt[i]).astype(np.float32)) / 256.0).reshape(-1)                   # :94-96
        if dsup[i] is not None:
            d["depth_sup"] = scale * (np.array(imread(dsup[i]).astype(np.float32)) / 256.0).reshape(-1)                 # :99-101
        out.append(d)
    return out
