"""A LEARNABLE synthetic scene in the reference's on-disk format  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

``loader_scene.py`` writes noise images (enough to pin the loader); training / convergence tests need something a NeRF
can fit: this module ray-casts a small analytic scene (a checkered ground plane, three striped spheres and a distant
"sky" shell outside the unit sphere) from a fan of cameras near the origin and writes it in the layout of
data_loader_split.py:27-129:

    <base>/<scene>/scale, <split>/{intrinsics,pose}/*.txt, <split>/rgb/*.png (8-bit), <split>/depth/*.png (uint16, metres x 256;
    0 = no return, as a LiDAR map has for the sky), <split>/depth_mono/*.png (a dense, slightly biased prior)

Depth convention (nerf_sample_ray_split.py:94-102, ddp_model.py:105): the maps hold the camera-z depth -- the rays are
``K^-1 [u+.5, v+.5, 1]`` rotated into the world, NOT normalised, so the ray parameter ``t`` of a hit IS its z-depth --
in metres; ``png / 256 * scale`` brings it to scene units.  The images have >= 1024 pixels because the unmodified
trainer forces ``N_rand = 1024`` on any GPU with more than 14 GB (ddp_train_nerf.py:364-369) and draws the pixels
without replacement (nerf_sample_ray_split.py:178).
"""
import os

import cv2
import numpy as np

H, W = 48, 64                      # 3072 pixels per image
N_TRAIN, N_TEST = 12, 3
SCALE = 0.02                       # metres -> unit-sphere units: the sphere radius is 50 m
SPHERES = (                        # centre, radius, base colour
    ((0.10, -0.02, 0.45), 0.13, (0.9, 0.2, 0.2)),
    ((-0.22, 0.03, 0.60), 0.18, (0.2, 0.8, 0.3)),
    ((0.30, 0.05, 0.75), 0.20, (0.2, 0.3, 0.9)),
)
GROUND_Y = 0.15                    # image y points down: the ground is at +y
SKY_R = 4.0                        # colour of the shell the background net has to learn


def _write_txt(path, mat):
    with open(path, "w") as f:
        f.write(" ".join("%.9g" % float(x) for x in np.asarray(mat, np.float64).reshape(-1)) + "\n")


def cast(rays_o, rays_d):
    """Nearest hit of every ray: returns rgb [n,3] in [0,1] and the ray parameter t [n] (inf for the sky)."""
    n = rays_o.shape[0]
    t_best = np.full(n, np.inf)
    rgb = np.zeros((n, 3))
    # ground plane y = GROUND_Y, only inside the unit sphere
    with np.errstate(divide="ignore", invalid="ignore"):
        t = (GROUND_Y - rays_o[:, 1]) / rays_d[:, 1]
    p = rays_o + t[:, None] * rays_d
    ok = (t > 1e-3) & np.isfinite(t) & ((p * p).sum(-1) < 0.95 ** 2)
    chk = ((np.floor(p[:, 0] * 8.0) + np.floor(p[:, 2] * 8.0)) % 2 == 0)
    col = np.where(chk[:, None], np.array([0.85, 0.85, 0.80]), np.array([0.25, 0.22, 0.20]))
    t_best = np.where(ok, t, t_best)
    rgb = np.where(ok[:, None], col, rgb)
    for c, r, base in SPHERES:
        c = np.asarray(c)
        oc = rays_o - c
        a = (rays_d * rays_d).sum(-1)
        b = 2.0 * (oc * rays_d).sum(-1)
        cc = (oc * oc).sum(-1) - r * r
        disc = b * b - 4 * a * cc
        t = (-b - np.sqrt(np.maximum(disc, 0.0))) / (2 * a)
        ok = (disc > 0) & (t > 1e-3) & (t < t_best)
        p = rays_o + t[:, None] * rays_d
        nrm = (p - c) / r
        stripes = 0.5 + 0.5 * np.sin(24.0 * nrm[:, 1] + 6.0 * nrm[:, 0])
        shade = 0.55 + 0.45 * np.clip(nrm @ np.array([0.4, -0.8, -0.45]), 0.0, 1.0)
        col = (np.asarray(base)[None, :] * (0.55 + 0.45 * stripes[:, None])) * shade[:, None]
        t_best = np.where(ok, t, t_best)
        rgb = np.where(ok[:, None], col, rgb)
    # sky: smooth function of the direction (what the inverted-sphere background has to represent)
    sky = ~np.isfinite(t_best)
    d = rays_d / np.linalg.norm(rays_d, axis=-1, keepdims=True)
    skycol = np.stack((0.45 + 0.25 * np.sin(3.0 * d[:, 0] + 1.0), 0.60 + 0.20 * np.sin(2.0 * d[:, 1] - 0.5),
                       0.80 + 0.15 * np.cos(4.0 * d[:, 0] * d[:, 2])), -1)
    rgb = np.where(sky[:, None], skycol, rgb)
    return np.clip(rgb, 0.0, 1.0), t_best


def camera(i, n, rng):
    """Camera i of n: positions on a short arc around the origin, all looking roughly along +z (a driving sequence)."""
    K = np.eye(4)
    K[0, 0] = K[1, 1] = 58.0
    K[0, 2], K[1, 2] = W / 2.0, H / 2.0
    s = (i + 0.5) / n - 0.5
    pos = np.array([0.35 * s, -0.02 + 0.04 * rng.rand(), -0.25 + 0.12 * rng.rand()])
    yaw, pitch = -0.5 * s + 0.05 * rng.randn(), 0.10 + 0.03 * rng.randn()
    Ry = np.array([[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]])
    Rx = np.array([[1, 0, 0], [0, np.cos(pitch), -np.sin(pitch)], [0, np.sin(pitch), np.cos(pitch)]])
    c2w = np.eye(4)
    c2w[:3, :3] = Ry @ Rx
    c2w[:3, 3] = pos
    return K, c2w


def rays_of(K, c2w):
    """nerf_sample_ray_split.py:10-34, float64 (only to render the ground truth)."""
    u, v = np.meshgrid(np.arange(W) + 0.5, np.arange(H) + 0.5)
    pix = np.stack((u.reshape(-1), v.reshape(-1), np.ones(H * W)), 0)
    d = (c2w[:3, :3] @ (np.linalg.inv(K[:3, :3]) @ pix)).T
    o = np.tile(c2w[:3, 3][None, :], (H * W, 1))
    return o, d


def write_scene(base, scene="synth_learnable", seed=0):
    rng = np.random.RandomState(seed)
    root = os.path.join(base, scene)
    os.makedirs(root, exist_ok=True)
    with open(os.path.join(root, "scale"), "w") as f:
        f.write("%r\n" % SCALE)
    idx = 0
    for split, cnt in (("train", N_TRAIN), ("test", N_TEST)):
        sd = os.path.join(root, split)
        for sub in ("intrinsics", "pose", "rgb", "depth", "depth_mono"):
            os.makedirs(os.path.join(sd, sub), exist_ok=True)
        for i in range(cnt):
            name = "%06d" % idx
            idx += 1
            K, c2w = camera(i if split == "train" else i * 4 + 1.5, N_TRAIN, rng)
            o, d = rays_of(K, c2w)
            rgb, t = cast(o, d)
            _write_txt(os.path.join(sd, "intrinsics", name + ".txt"), K)
            _write_txt(os.path.join(sd, "pose", name + ".txt"), c2w)
            img = np.round(rgb.reshape(H, W, 3) * 255.0).astype(np.uint8)
            cv2.imwrite(os.path.join(sd, "rgb", name + ".png"), img[:, :, ::-1])                 # cv2 writes BGR
            metres = np.where(np.isfinite(t), t / SCALE, 0.0).reshape(H, W)
            gt = np.clip(np.round(metres * 256.0), 0, 65535).astype(np.uint16)
            lidar = gt.copy()
            lidar[rng.rand(H, W) < 0.5] = 0                                                      # sparse returns
            cv2.imwrite(os.path.join(sd, "depth", name + ".png"), lidar)
            mono = np.clip(np.round(metres * (1.0 + 0.03 * rng.randn(H, W)) * 256.0), 0, 65535).astype(np.uint16)
            mono[gt == 0] = 0
            cv2.imwrite(os.path.join(sd, "depth_mono", name + ".png"), mono)
    return root
