"""Backward of NerfNet.forward through the CUDA library (nerfpp_backward)."""
import ctypes

import torch

from . import _lib
from ._lib import NerfppError, check
from .ops import RET_KEYS, _p, _stream, net_params_struct


def render_backward(impl, params, inputs, outs, ws, grads):
    """grads: d(loss)/d(each of the 10 outputs) or None. Returns the 48 parameter gradients
    (fg 24 then bg 24, weight/bias interleaved in C-ABI layer order)."""
    L = _lib.lib()
    if not hasattr(L, "nerfpp_backward"):
        raise NerfppError("libnerfpp_b200.so was built without nerfpp_backward; training needs it "
                          "(no PyTorch fallback exists for this path)")
    o, d, zmax, fz, bz = inputs
    n, sf = fz.shape
    sb = bz.shape[1]
    dev = fz.device
    gst = _lib.RenderGrads()
    keep = []
    for k, g in zip(RET_KEYS, grads):
        if g is None or k == "fg_dists":
            setattr(gst, k, None)
        else:
            g = g.contiguous().float()
            keep.append(g)
            setattr(gst, k, g.data_ptr())
    ost = _lib.RenderOut()
    for k in RET_KEYS:
        setattr(ost, k, outs[k].data_ptr())
    pgrads = [torch.zeros_like(p) for p in params]
    gfg, gbg = _lib.NetGrads(), _lib.NetGrads()
    for l in range(_lib.NLAYERS):
        gfg.w[l], gfg.b[l] = pgrads[2 * l].data_ptr(), pgrads[2 * l + 1].data_ptr()
        gbg.w[l], gbg.b[l] = pgrads[24 + 2 * l].data_ptr(), pgrads[24 + 2 * l + 1].data_ptr()
    pfg, pbg = net_params_struct(params[:24]), net_params_struct(params[24:])
    bws = torch.empty(max(int(L.nerfpp_backward_workspace_bytes(n, sf, sb)), 1), device=dev, dtype=torch.uint8)
    with torch.cuda.device(dev):
        check(L.nerfpp_backward(ctypes.byref(pfg), ctypes.byref(pbg), _p(o), _p(d), _p(zmax), _p(fz), _p(bz), n, sf, sb,
                                ctypes.byref(ost), ctypes.byref(gst), _p(ws), ctypes.byref(gfg), ctypes.byref(gbg),
                                _p(bws), _stream()), "backward")
    return pgrads
