"""Deterministic synthetic scene in the reference's on-disk format (data_loader_split.py:27-129; README "Data format"):

    <base>/<scene>/scale                      one float (metres -> unit-sphere units)
    <base>/<scene>/<split>/intrinsics/*.txt   16 floats, row-major 4x4
    <base>/<scene>/<split>/pose/*.txt         16 floats, row-major 4x4 camera-to-world
    <base>/<scene>/<split>/rgb/*.png          8-bit RGB
    <base>/<scene>/<split>/depth/*.png        uint16, metres x 256 (0 = no measurement)
    <base>/<scene>/<split>/depth_<type>/*.png same, the supervision prior of that type
    <base>/<scene>/<split>/min_depth/*.png    8-bit, fraction of max_depth           (optional)
    <base>/<scene>/<split>/max_depth.txt      one float                               (optional)

Test infrastructure: used by oracle/gen_golden_loader.py (which runs the unmodified reference loader on it) and by the
tests (which run the oracle restatement and the product loader on the same files).  PNG is lossless, so every reader
sees the same pixels."""
import os

import cv2
import numpy as np

H, W = 24, 40
N_TRAIN, N_TEST = 5, 2
SCALE = 0.0125
MAX_DEPTH = 3.5


def _write_txt(path, mat):
    with open(path, "w") as f:
        f.write(" ".join("%.9g" % float(x) for x in np.asarray(mat, np.float64).reshape(-1)) + "\n")


def write_scene(base, scene="synth", seed=0, with_min_depth=True):
    rng = np.random.RandomState(seed)
    root = os.path.join(base, scene)
    os.makedirs(root, exist_ok=True)
    with open(os.path.join(root, "scale"), "w") as f:
        f.write("%r\n" % SCALE)
    for split, cnt in (("train", N_TRAIN), ("test", N_TEST)):
        sd = os.path.join(root, split)
        for sub in ("intrinsics", "pose", "rgb", "depth", "depth_mono") + (("min_depth",) if with_min_depth else ()):
            os.makedirs(os.path.join(sd, sub), exist_ok=True)
        if with_min_depth:
            with open(os.path.join(sd, "max_depth.txt"), "w") as f:
                f.write("%r\n" % MAX_DEPTH)
        for i in range(cnt):
            name = "%06d" % (3 * i + (0 if split == "train" else 1))
            K = np.eye(4)
            K[0, 0], K[1, 1], K[0, 2], K[1, 2] = 35.0 + rng.rand(), 34.0 + rng.rand(), W / 2 + rng.rand(), H / 2 + rng.rand()
            ang = rng.randn(3) * 0.3
            Rx = np.array([[1, 0, 0], [0, np.cos(ang[0]), -np.sin(ang[0])], [0, np.sin(ang[0]), np.cos(ang[0])]])
            Ry = np.array([[np.cos(ang[1]), 0, np.sin(ang[1])], [0, 1, 0], [-np.sin(ang[1]), 0, np.cos(ang[1])]])
            Rz = np.array([[np.cos(ang[2]), -np.sin(ang[2]), 0], [np.sin(ang[2]), np.cos(ang[2]), 0], [0, 0, 1]])
            c2w = np.eye(4)
            c2w[:3, :3] = Rz @ Ry @ Rx
            c2w[:3, 3] = rng.randn(3) * 0.2            # inside the unit sphere
            _write_txt(os.path.join(sd, "intrinsics", name + ".txt"), K)
            _write_txt(os.path.join(sd, "pose", name + ".txt"), c2w)
            rgb = rng.randint(0, 256, size=(H, W, 3)).astype(np.uint8)
            cv2.imwrite(os.path.join(sd, "rgb", name + ".png"), rgb[:, :, ::-1])          # cv2 writes BGR
            depth = rng.randint(0, 80 * 256, size=(H, W)).astype(np.uint16)
            depth[rng.rand(H, W) < 0.6] = 0                                                 # sparse LiDAR
            cv2.imwrite(os.path.join(sd, "depth", name + ".png"), depth)
            mono = rng.randint(1, 80 * 256, size=(H, W)).astype(np.uint16)                  # dense prior
            cv2.imwrite(os.path.join(sd, "depth_mono", name + ".png"), mono)
            if with_min_depth:
                cv2.imwrite(os.path.join(sd, "min_depth", name + ".png"), rng.randint(0, 256, size=(H, W)).astype(np.uint8))
    return root
