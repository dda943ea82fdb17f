"""Device-resident drop-in for ``RaySamplerSingleImage`` (nerf-methods/nerfplusplus/nerf_sample_ray_split.py:37-221),
SURVEY.md section 8(f) N1.

The reference precomputes every ray of the image on the host (:104-106), and per training step gathers ``N_rand`` rows of
seven numpy arrays and copies them to the GPU (ddp_train_nerf.py:423-427).  Here the image, depth prior and min-depth map
are uploaded once; a step is one kernel that turns pixel indices into rays (K^-1, c2w) and gathers the pixels' rgb / prior.
Only the arithmetic moved: the pixel indices are drawn by the same ``np.random.choice`` call as the reference by default,
so a seeded run selects the same pixels (``device_rng=True`` draws them with torch on the device instead)."""
import ctypes
from collections import OrderedDict

import numpy as np
import torch

from . import _lib
from .ops import _p, _stream, check


class DeviceRaySampler(object):
    def __init__(self, H, W, intrinsics, c2w, img=None, depth_sup=None, min_depth=None, img_path=None, depth_scale=None,
                 device="cuda"):
        self.H, self.W = int(H), int(W)
        self.img_path, self.depth_scale = img_path, depth_scale
        self.device = torch.device(device)
        intrinsics, c2w = np.asarray(intrinsics, np.float32), np.asarray(c2w, np.float32)
        self._kinv = np.ascontiguousarray(np.linalg.inv(intrinsics[:3, :3]), np.float32)       # :23
        self._c2w = np.ascontiguousarray(c2w, np.float32)
        self._cam_depth = float(np.linalg.inv(c2w)[2, 3])                                     # :31
        up = lambda a, shape: None if a is None else torch.as_tensor(np.ascontiguousarray(a, np.float32).reshape(shape)).to(self.device)
        self.img = up(img, (self.H * self.W, 3))
        self.depth_sup = up(depth_sup, (self.H * self.W,))
        self.min_depth = up(min_depth, (self.H * self.W,))

    def _rays(self, ids, n):
        dev = self.device
        out = OrderedDict(ray_o=torch.empty(n, 3, device=dev), ray_d=torch.empty(n, 3, device=dev), depth=torch.empty(n, device=dev),
                          rgb=torch.empty(n, 3, device=dev) if self.img is not None else None, mask=None,
                          min_depth=torch.empty(n, device=dev))
        ds = torch.empty(n, device=dev) if self.depth_sup is not None else None
        f = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        with torch.cuda.device(dev):
            check(_lib.lib().nerfpp_gen_rays(f(self._kinv), f(self._c2w), self._cam_depth, self.W, _p(ids), n, _p(self.img),
                                             _p(self.depth_sup), _p(self.min_depth), _p(out["ray_o"]), _p(out["ray_d"]), _p(out["depth"]),
                                             _p(out["rgb"]), _p(ds), _p(out["min_depth"]), _stream()), "gen_rays")
        if ds is not None:
            out["depth_sup"] = ds
        return out

    def get_depth_scale(self):
        return self.depth_scale if self.depth_sup is not None else None

    def get_all(self):
        """:131-153 -- every pixel in row-major order, as device tensors."""
        return self._rays(None, self.H * self.W)

    def random_sample(self, N_rand, center_crop=False, select_inds=None, device_rng=False):
        """:155-221.  ``select_inds`` overrides the draw (tests)."""
        if select_inds is None:
            if center_crop:
                half_H, half_W = self.H // 2, self.W // 2
                quad_H, quad_W = half_H // 2, half_W // 2
                u, v = np.meshgrid(np.arange(half_W - quad_W, half_W + quad_W), np.arange(half_H - quad_H, half_H + quad_H))
                u, v = u.reshape(-1), v.reshape(-1)
                sel = np.random.choice(u.shape[0], size=(N_rand,), replace=False)
                select_inds = v[sel] * self.W + u[sel]
            elif device_rng:
                select_inds = torch.randperm(self.H * self.W, device=self.device)[:N_rand]
            else:
                select_inds = np.random.choice(self.H * self.W, size=(N_rand,), replace=False)
        ids = torch.as_tensor(select_inds, dtype=torch.int64).to(self.device)
        ret = self._rays(ids, int(ids.numel()))
        ret["img_name"] = self.img_path
        return ret
