"""Drop-in for the reference's ``ddp_model`` module (nerf-methods/nerfplusplus/ddp_model.py).

Same import surface -- ``NerfNetWithAutoExpo``, ``NerfNet``, ``depth2pts_outside``, ``remap_name`` --
same constructor arguments, ``forward`` signature, 10-key ``OrderedDict`` return and, crucially,
the same ``state_dict`` names and shapes (``nerf_net.{fg,bg}_net.base_layers.{i}.0.weight`` ...), so
reference checkpoints load (ddp_train_nerf.py:349-352) and ``ddp_train_nerf.py`` drives it
unchanged when this directory precedes the reference on PYTHONPATH.  The modules below only hold
parameters; ``forward`` hands them to the CUDA library (nerfpp_b200) -- there is no torch
implementation of the maths here and no CPU fallback.
"""
import logging
from collections import OrderedDict

import torch
import torch.nn as nn

from nerfpp_b200 import ops

logger = logging.getLogger(__package__)
TINY_NUMBER = 1e-6   # utils.py:8 (kept local so this module imports without the reference's utils)


class Embedder(nn.Module):
    """Parameter-free stand-in for nerf_network.Embedder (nerf_network.py:11-60): records the
    encoding shape; the encoding itself is computed inside the field kernels."""

    def __init__(self, input_dim, max_freq_log2, N_freqs):
        super().__init__()
        self.input_dim, self.N_freqs, self.max_freq_log2 = input_dim, N_freqs, max_freq_log2
        self.out_dim = input_dim * (1 + 2 * N_freqs)


class MLPNet(nn.Module):
    """Parameter container with the exact layer names/shapes/creation order of
    nerf_network.MLPNet (nerf_network.py:70-118) -- default nn.Linear init, so one
    torch.manual_seed reproduces the reference's weights."""

    def __init__(self, D=8, W=256, input_ch=3, input_ch_viewdirs=3, skips=(4,), use_viewdirs=False):
        super().__init__()
        self.input_ch, self.input_ch_viewdirs, self.use_viewdirs, self.skips = input_ch, input_ch_viewdirs, use_viewdirs, list(skips)
        layers, dim = [], input_ch
        for i in range(D):
            layers.append(nn.Sequential(nn.Linear(dim, W), nn.ReLU()))
            dim = W + input_ch if (i in self.skips and i != D - 1) else W
        self.base_layers = nn.ModuleList(layers)
        self.sigma_layers = nn.Sequential(nn.Linear(dim, 1))
        self.base_remap_layers = nn.Sequential(nn.Linear(dim, 256))
        self.rgb_layers = nn.Sequential(nn.Linear(256 + input_ch_viewdirs, W // 2), nn.ReLU(), nn.Linear(W // 2, 3), nn.Sigmoid())

    def tensors(self):
        """The 24 parameter tensors in the C ABI's layer order (NerfppNetParams)."""
        mods = [seq[0] for seq in self.base_layers] + [self.sigma_layers[0], self.base_remap_layers[0],
                                                       self.rgb_layers[0], self.rgb_layers[2]]
        out = []
        for m in mods:
            out += [m.weight, m.bias]
        return tuple(out)


class NerfNet(nn.Module):
    """ddp_model.NerfNet (ddp_model.py:48-147)."""

    def __init__(self, args):
        super().__init__()
        cfg = (args.netdepth, args.netwidth, args.max_freq_log2, args.max_freq_log2_viewdirs)
        if cfg != (8, 256, 10, 4):
            raise ValueError("nerfpp_b200 builds the reference's only scripted shape netdepth=8 netwidth=256 "
                             "max_freq_log2=10 max_freq_log2_viewdirs=4 (configs/kitti.txt:36-40); got %r" % (cfg,))
        for side, dim in (("fg", 3), ("bg", 4)):
            pos = Embedder(dim, args.max_freq_log2 - 1, args.max_freq_log2)
            view = Embedder(3, args.max_freq_log2_viewdirs - 1, args.max_freq_log2_viewdirs)
            setattr(self, side + "_embedder_position", pos)
            setattr(self, side + "_embedder_viewdir", view)
            setattr(self, side + "_net", MLPNet(D=args.netdepth, W=args.netwidth, input_ch=pos.out_dim,
                                                input_ch_viewdirs=view.out_dim, use_viewdirs=args.use_viewdirs))
        self._packed = (ops.PackedNet(False), ops.PackedNet(True))

    def invalidate_packed(self):
        """Drops the cached fp16 weight tiles.  Needed only after parameter writes that bypass torch's version counter
        (``p.data.copy_(...)``, EMA through ``.data``); optimizer steps and ``load_state_dict`` are seen automatically."""
        for c in self._packed:
            c.invalidate()

    def forward(self, ray_o, ray_d, fg_z_max, fg_z_vals, bg_z_vals, impl=None):
        lead = tuple(ray_d.shape[:-1])
        flat = len(lead) != 1
        if flat:   # the reference accepts arbitrary leading dims ([..., 3]); kernels take [n, .]
            ray_o, ray_d = ray_o.reshape(-1, 3), ray_d.reshape(-1, 3)
            fg_z_max = fg_z_max.reshape(-1)
            fg_z_vals, bg_z_vals = fg_z_vals.reshape(-1, fg_z_vals.shape[-1]), bg_z_vals.reshape(-1, bg_z_vals.shape[-1])
        ret = ops.nerfpp_forward(self._packed, self.fg_net.tensors(), self.bg_net.tensors(), ray_o, ray_d, fg_z_max,
                                 fg_z_vals, bg_z_vals, impl)
        if flat:
            ret = OrderedDict((k, v.reshape(lead + tuple(v.shape[1:]))) for k, v in ret.items())
        return ret


def depth2pts_outside(ray_o, ray_d, depth):
    """ddp_model.depth2pts_outside (ddp_model.py:16-45) evaluated by the background field kernel's
    geometry stage. ray_o, ray_d [..., 3], depth [...] -> pts [..., 4], depth_real [...]."""
    return ops_depth2pts(ray_o, ray_d, depth)


def ops_depth2pts(ray_o, ray_d, depth):
    from nerfpp_b200 import geometry
    return geometry.depth2pts_outside(ray_o, ray_d, depth)


def remap_name(name):
    """ddp_model.remap_name (ddp_model.py:150-158): last three path components, dots -> dashes."""
    name = name.replace(".", "-")
    if name.endswith("/"):
        name = name[:-1]
    parts = name.split("/")
    return "/".join(parts[-3:])


class NerfNetWithAutoExpo(nn.Module):
    """ddp_model.NerfNetWithAutoExpo (ddp_model.py:161-192)."""

    def __init__(self, args, optim_autoexpo=False, img_names=None):
        super().__init__()
        self.nerf_net = NerfNet(args)
        self.optim_autoexpo = optim_autoexpo
        if optim_autoexpo:
            assert img_names is not None
            logger.info("Optimizing autoexposure!")
            self.img_names = [remap_name(x) for x in img_names]
            logger.info("\n".join(self.img_names))
            self.autoexpo_params = nn.ParameterDict(OrderedDict((x, nn.Parameter(torch.Tensor([0.5, 0.]))) for x in self.img_names))

    def forward(self, ray_o, ray_d, fg_z_max, fg_z_vals, bg_z_vals, img_name=None, impl=None):
        ret = self.nerf_net(ray_o, ray_d, fg_z_max, fg_z_vals, bg_z_vals, impl=impl)
        if img_name is not None:
            img_name = remap_name(img_name)
        if self.optim_autoexpo and img_name in self.autoexpo_params:
            autoexpo = self.autoexpo_params[img_name]
            ret["autoexpo"] = (torch.abs(autoexpo[0]) + 0.5, autoexpo[1])   # scale kept positive, ddp_model.py:188
        return ret
