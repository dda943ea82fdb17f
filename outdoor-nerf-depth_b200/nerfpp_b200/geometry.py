"""Standalone geometry ops of the path (drop-ins for helpers the reference exports)."""
import torch

from . import _lib
from .ops import _c, _p, _stream
from ._lib import check


def depth2pts_outside(ray_o, ray_d, depth):
    """ddp_model.depth2pts_outside (ddp_model.py:16-45). ray_o, ray_d [..., 3], depth [...]."""
    lead = tuple(depth.shape)
    o = _c(ray_o, "ray_o").expand(lead + (3,)).reshape(-1, 3).contiguous()
    d = _c(ray_d, "ray_d").expand(lead + (3,)).reshape(-1, 3).contiguous()
    z = _c(depth, "depth").reshape(-1).contiguous()
    n = z.numel()
    pts = torch.empty(n, 4, device=z.device, dtype=torch.float32)
    real = torch.empty(n, device=z.device, dtype=torch.float32)
    with torch.cuda.device(z.device):
        check(_lib.lib().nerfpp_depth2pts_outside(_p(o), _p(d), _p(z), n, _p(pts), _p(real), _stream()), "depth2pts_outside")
    return pts.reshape(lead + (4,)), real.reshape(lead)
