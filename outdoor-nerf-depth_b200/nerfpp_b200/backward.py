"""Backward of NerfNet.forward through the CUDA library (nerfpp_backward): composite backward, then per net the
tcgen05 data-gradient chain and the split-K weight-gradient GEMMs.  No PyTorch fallback."""
import ctypes
import os

import torch

from . import _lib
from .ops import RET_KEYS, _p, _stream, check, net_params_struct


LAST_FLAT_GRADS = [None]      # the flat gradient buffer of the most recent backward (GraphedTrainStep all-reduces it in one call)


def render_backward(params, inputs, outs, ws, tws, grads):
    """grads: d(loss)/d(each of the 10 outputs) or None. Returns the 48 parameter gradients
    (fg 24 then bg 24, weight/bias interleaved in C-ABI layer order)."""
    L = _lib.lib()
    o, d, zmax, fz, bz = inputs
    n, sf = fz.shape
    sb = bz.shape[1]
    dev = fz.device
    gst = _lib.RenderOut()
    keep = []
    for k, g in zip(RET_KEYS, grads):
        if g is not None and k != "fg_dists":
            g = g.contiguous().float()
            keep.append(g)
            setattr(gst, k, g.data_ptr())
    ost = _lib.RenderOut()
    for k in RET_KEYS:
        setattr(ost, k, outs[k].data_ptr())
    # one zero-filled buffer for all 48 gradients (the wgrad kernels accumulate with red.global.add): one fill kernel
    # instead of 48; every tensor starts on a 16-byte boundary
    sizes = [(p.numel() + 3) // 4 * 4 for p in params]
    flat = torch.zeros(sum(sizes), device=dev, dtype=torch.float32)
    LAST_FLAT_GRADS[0] = flat
    pgrads, off = [], 0
    for p, sz in zip(params, sizes):
        pgrads.append(flat[off:off + p.numel()].view(p.shape))
        off += sz
    gfg, gbg = _lib.NetGrads(), _lib.NetGrads()
    for l in range(_lib.NLAYERS):
        gfg.w[l], gfg.b[l] = pgrads[2 * l].data_ptr(), pgrads[2 * l + 1].data_ptr()
        gbg.w[l], gbg.b[l] = pgrads[24 + 2 * l].data_ptr(), pgrads[24 + 2 * l + 1].data_ptr()
    pfg, pbg = net_params_struct(params[:24]), net_params_struct(params[24:])
    bws = torch.empty(int(L.nerfpp_backward_workspace_bytes(n, sf, sb)), device=dev, dtype=torch.uint8)
    with torch.cuda.device(dev):
        check(L.nerfpp_backward(ctypes.byref(pfg), ctypes.byref(pbg), _p(d), _p(zmax), _p(fz), _p(bz), n, sf, sb,
                                ctypes.byref(ost), ctypes.byref(gst), _p(ws), _p(tws), ctypes.byref(gfg), ctypes.byref(gbg),
                                _p(bws), _stream()), "backward")
    return pgrads
