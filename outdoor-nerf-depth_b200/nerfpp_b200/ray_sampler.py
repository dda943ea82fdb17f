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


def decode_pixels(raw, div, mul=1.0, add=0.0, device="cuda"):
    """N2: uploads a uint8 / uint16 pixel array as it is and decodes it on the device to fp32
    ``((x / div) * mul) + add`` (nerf_sample_ray_split.py:73-102).  Returns a flat float32 device tensor."""
    raw = np.ascontiguousarray(raw)
    if raw.dtype == np.uint8:
        bits, t = 8, torch.from_numpy(raw.reshape(-1))
    elif raw.dtype == np.uint16:
        bits, t = 16, torch.from_numpy(raw.reshape(-1).view(np.int16))      # same bytes; torch has no uint16 arithmetic
    else:
        raise ValueError("decode_pixels: uint8 or uint16 pixels expected, got %s" % raw.dtype)
    dev = torch.device(device)
    if dev.type != "cuda":
        raise _lib.NerfppError("decode_pixels: CUDA device required (there is no CPU path)")
    src = t.to(dev, non_blocking=True)
    out = torch.empty(src.numel(), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        check(_lib.lib().nerfpp_decode_pixels(_p(src), bits, src.numel(), float(div), float(np.float32(mul)), float(np.float32(add)),
                                              _p(out), _stream()), "decode_pixels")
    return out


class DeviceRaySampler(object):
    def __init__(self, H, W, intrinsics, c2w, img=None, depth_sup=None, min_depth=None, img_path=None, depth_scale=None,
                 device="cuda", depth_gt=None, mask=None):
        self.H, self.W = int(H), int(W)
        self.img_path, self.depth_scale = img_path, depth_scale
        self.device = torch.device(device)
        intrinsics, c2w = np.asarray(intrinsics, np.float32), np.asarray(c2w, np.float32)
        self.intrinsics, self.c2w_mat = intrinsics, c2w
        self._kinv = np.ascontiguousarray(np.linalg.inv(intrinsics[:3, :3]), np.float32)       # :23
        self._c2w = np.ascontiguousarray(c2w, np.float32)
        self._cam_depth = float(np.linalg.inv(c2w)[2, 3])                                     # :31

        def up(a, shape):
            if a is None:
                return None
            if torch.is_tensor(a):
                return a.to(self.device).float().reshape(shape)
            return torch.as_tensor(np.ascontiguousarray(a, np.float32).reshape(shape)).to(self.device)
        self.img = up(img, (self.H * self.W, 3))
        self.depth_sup = up(depth_sup, (self.H * self.W,))
        self.min_depth = up(min_depth, (self.H * self.W,))
        self.depth_gt = up(depth_gt, (self.H * self.W,))
        self.mask = up(mask, (self.H * self.W,))

    @classmethod
    def from_files(cls, H, W, intrinsics, c2w, img_path=None, mask_path=None, min_depth_path=None, max_depth=None,
                   depth_gt_path=None, depth_sup_path=None, depth_scale=None, read_image=None, device="cuda"):
        """N2: the constructor arguments of the reference's RaySamplerSingleImage (:37-46).  Every image is read as raw
        pixels, uploaded and decoded on the device; no per-pixel work happens on the host."""
        def dec(path, div, mul=1.0, add=0.0, channels=1):
            if path is None:
                return None
            raw = read_image(path)
            want = (int(H), int(W), 3) if channels == 3 else (int(H), int(W))
            if raw.shape != want:
                raise ValueError("%s: shape %s, expected %s (only resolution_level 1 is supported)" % (path, raw.shape, want))
            return decode_pixels(raw, div, mul, add, device)
        if min_depth_path is not None and max_depth is None:
            raise ValueError("min_depth images need max_depth.txt (nerf_sample_ray_split.py:87)")
        if (depth_gt_path is not None or depth_sup_path is not None) and depth_scale is None:
            raise ValueError("depth images need the scene's scale file (data_loader_split.py:86)")
        s = cls(H, W, intrinsics, c2w, img=dec(img_path, 255.0, channels=3), img_path=img_path, depth_scale=depth_scale, device=device,
                mask=dec(mask_path, 255.0), min_depth=dec(min_depth_path, 255.0, max_depth, 1e-4) if min_depth_path is not None else None,
                depth_gt=dec(depth_gt_path, 256.0, depth_scale), depth_sup=dec(depth_sup_path, 256.0, depth_scale))
        s.mask_path, s.min_depth_path, s.max_depth = mask_path, min_depth_path, max_depth
        s.depth_gt_path, s.depth_sup_path = depth_gt_path, depth_sup_path
        return s

    # accessors of the reference class (:108-129).  The unmodified trainer does numpy arithmetic on what they return
    # (ddp_train_nerf.py:557-570, ddp_test_nerf.py:82-91), so they hand out HOST arrays like the reference; the
    # device-resident copies are ``self.img`` / ``self.depth_gt`` / ``self.depth_sup`` (flat, row-major).
    resolution_level = 1                      # ddp_train_nerf.py:421 logs it; the trainer never changes it

    def _host(self, name, shape):
        t = getattr(self, name)
        if t is None:
            return None
        cache = self.__dict__.setdefault("_host_cache", {})
        if name not in cache:
            cache[name] = t.reshape(shape).cpu().numpy()
        return cache[name]

    def get_img(self):
        return self._host("img", (self.H, self.W, 3))

    def get_gt_depth_img(self):
        return self._host("depth_gt", (self.H, self.W))

    def get_sup_depth_img(self):
        return self._host("depth_sup", (self.H, self.W))

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
        if ids is None:
            out["mask"] = self.mask
        elif self.mask is not None:
            out["mask"] = self.mask[ids]
        if self.depth_gt is not None and (ids is None or ds is not None):       # random_sample returns it with depth_sup only (:209-211)
            out["depth_gt"] = self.depth_gt if ids is None else self.depth_gt[ids]
        if ds is not None:
            out["depth_sup"] = ds
        return out

    def get_depth_scale(self):
        return self.depth_scale if (self.depth_sup is not None or self.depth_gt is not None) else None

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
