"""Autograd nodes for the depth-prior losses (CUDA forward + backward, no torch maths)."""

import torch

from . import _lib
from ._lib import DEPTH_KL, DEPTH_L1, DEPTH_MSE, NerfppError, check  # noqa: F401  (the constants are re-exported: depth_loss.py uses them)
from .ops import _c, _p, _stream


def _loss_call(typ, depth, depth_sup, w, z, dists, far, n, S, kl_sigma, dev):
    L = _lib.lib()
    out = torch.empty(4, device=dev, dtype=torch.float32)
    ws = torch.empty(int(L.nerfpp_loss_workspace_bytes()), device=dev, dtype=torch.uint8)
    with torch.cuda.device(dev):
        check(L.nerfpp_depth_loss(_p(depth), _p(depth_sup), _p(w), _p(z), _p(dists), _p(far), n, S, typ, float(kl_sigma),
                                  _p(out), _p(ws), _stream()), "depth_loss")
    return out


class DepthPointLoss(torch.autograd.Function):
    """depth_mse / depth_l1 (depth_loss.py:4-18): masked mean over rays with gt > 0."""

    @staticmethod
    def forward(ctx, depth_gt, depth_pred, typ):
        gt, pred = _c(depth_gt, "depth_gt").reshape(-1), _c(depth_pred, "depth_pred").reshape(-1)
        if gt.shape != pred.shape:
            raise ValueError("depth_gt and depth_pred must have the same number of rays")
        out = _loss_call(typ, pred, gt, None, None, None, None, gt.numel(), 1, 1.0, gt.device)
        ctx.typ, ctx.shape = typ, depth_pred.shape
        ctx.save_for_backward(gt, pred, out)
        return out[1]

    @staticmethod
    def backward(ctx, g):
        gt, pred, out = ctx.saved_tensors
        grad = torch.empty_like(pred)
        with torch.cuda.device(gt.device):
            check(_lib.lib().nerfpp_depth_loss_backward(_p(pred), _p(gt), None, None, None, None, gt.numel(), 1, ctx.typ, 1.0,
                                                        _p(out), _p(g.contiguous().float()), _p(grad), _stream()),
                  "depth_loss_backward")
        return None, grad.reshape(ctx.shape), None


class DepthKLLoss(torch.autograd.Function):
    """depth_kl (depth_loss.py:20-44): gradient flows to ``weights`` only (steps/lengths carry none
    in the trainer: fg_depth is a detached input, ddp_train_nerf.py:452-457,489)."""

    @staticmethod
    def forward(ctx, weights, termination_depth, steps, lengths, sigma, fg_far_depth):
        w, z, dl = _c(weights, "weights", 2), _c(steps, "steps", 2), _c(lengths, "lengths", 2)
        t = _c(termination_depth, "termination_depth").reshape(-1)
        n, S = w.shape
        if fg_far_depth is None:
            far = torch.full((n,), float("inf"), device=w.device)
        else:
            far = _c(fg_far_depth, "fg_far_depth").reshape(-1)
        out = _loss_call(DEPTH_KL, None, t, w, z, dl, far, n, S, sigma, w.device)
        ctx.sigma = sigma
        ctx.save_for_backward(w, t, z, dl, far, out)
        return out[1]

    @staticmethod
    def backward(ctx, g):
        w, t, z, dl, far, out = ctx.saved_tensors
        grad = torch.empty_like(w)
        n, S = w.shape
        with torch.cuda.device(w.device):
            check(_lib.lib().nerfpp_depth_loss_backward(None, _p(t), _p(w), _p(z), _p(dl), _p(far), n, S, DEPTH_KL, ctx.sigma,
                                                        _p(out), _p(g.contiguous().float()), _p(grad), _stream()),
                  "depth_loss_backward")
        return grad, None, None, None, None, None
