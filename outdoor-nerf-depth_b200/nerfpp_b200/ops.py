"""Torch-facing wrappers over the C ABI (include/nerfpp_b200.h).

PyTorch is plumbing here: it owns device memory, the current stream and autograd bookkeeping.
All arithmetic of the hot path happens in libnerfpp_b200.so; nothing in this module computes a
result with torch ops (the only torch maths is the RNG draw and the linspace constants whose
rounding has to be the framework's own).  No CPU fallback: CPU tensors raise.
"""
import ctypes
import os
from collections import OrderedDict

import torch

from . import _lib
from ._lib import DEPTH_KL, DEPTH_L1, DEPTH_MSE, DEPTH_NONE, FIELD_SIMT, FIELD_TC, NerfppError, check

LAYER_NAMES = (["base_layers.%d.0" % i for i in range(8)]
               + ["sigma_layers.0", "base_remap_layers.0", "rgb_layers.0", "rgb_layers.2"])
RET_KEYS = ("rgb", "fg_weights", "bg_weights", "fg_dists", "fg_rgb", "fg_depth", "bg_rgb", "bg_depth",
            "bg_lambda", "depth")
LAUNCHES = [0]   # kernels of libnerfpp_b200.so launched through this module (bench.py reads it)
_KERNELS_PER_CALL = {"backward": 16, "intersect_sphere": 1, "coarse_depths": 1, "perturb_samples": 1, "sample_pdf": 1, "sample_cdf": 1,
                     "resample_merge": 1, "pack_weights": 1, "field_forward": 1, "forward": 3, "loss": 2, "mip360_depth_loss": 2}


def check(rc, what):   # noqa: F811  (wraps _lib.check to count launches)
    _lib.check(rc, what)
    LAUNCHES[0] += _KERNELS_PER_CALL.get(what, 1)


UNBOUNDED_MSG = ("Not all your cameras are bounded by the unit sphere; please make sure the cameras are "
                 "normalized properly!")   # ddp_train_nerf.py:63


def default_field_impl():
    """The field runs on the tcgen05 tensor-core kernel.  The plain-fp32 SIMT evaluator (FIELD_SIMT) is the cross-check /
    full-precision reference: callers ask for it explicitly (``impl=FIELD_SIMT``); no environment switch selects it."""
    return FIELD_TC


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk(t, name, ndim=None):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise NerfppError("%s is on %s: this path only exists as CUDA kernels (no CPU fallback)" % (name, t.device))
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32, got %s" % (name, t.dtype))
    if ndim is not None and t.dim() != ndim:
        raise ValueError("%s must have %d dims, got shape %s" % (name, ndim, tuple(t.shape)))
    return t


def _c(t, name, ndim=None):
    return _chk(t, name, ndim).detach().contiguous()


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _rows(t):
    """Flattens leading dims: [..., S] -> ([n, S] contiguous, leading shape)."""
    lead = tuple(t.shape[:-1])
    return t.reshape(-1, t.shape[-1]).contiguous(), lead


# ------------------------------------------------------------------------------------------------
# A1-A5 sampling
# ------------------------------------------------------------------------------------------------
class UnboundedFlag(object):
    """Deferred form of the reference's out-of-sphere check (ddp_train_nerf.py:62-63): the kernel sets a device flag;
    ``raise_if_set()`` reads it (one host sync) and raises the reference's exception."""

    _pinned, _next = [], [0]

    def __init__(self, flag):
        self.flag = flag
        if torch.cuda.is_current_stream_capturing():
            # inside a CUDA-graph capture (nerfpp_b200.graph.GraphedRenderStep) the owner of the graph copies the
            # flag out with the step's other results and checks it after each replay
            self.host = self.event = None
            return
        # the flag travels to a pinned host word right behind the kernel that sets it, so reading it later waits for
        # that kernel only, not for whatever has been enqueued since
        if len(self._pinned) < 16:
            self._pinned.append(torch.empty(1, dtype=torch.int32, pin_memory=True))
        self.host = self._pinned[self._next[0] % len(self._pinned)]
        self._next[0] += 1
        self.host.copy_(flag, non_blocking=True)
        self.event = torch.cuda.Event()
        self.event.record()

    def raise_if_set(self):
        if self.event is None:
            raise NerfppError("this flag belongs to a captured CUDA graph; GraphedRenderStep checks it after each replay")
        self.event.synchronize()
        if int(self.host[0]) != 0:
            raise Exception(UNBOUNDED_MSG)


def intersect_sphere(ray_o, ray_d, deferred=False):
    """ddp_train_nerf.py:51-66. ray_o, ray_d [..., 3] -> [...]. Raises like the reference when a
    camera is outside the unit sphere (same data-dependent host sync as its ``.any()``).  ``deferred=True`` returns
    (far, UnboundedFlag) without synchronising, for callers that check once at the end of a step."""
    o, lead = _rows(_c(ray_o, "ray_o"))
    d, _ = _rows(_c(ray_d, "ray_d"))
    if o.shape[-1] != 3 or d.shape != o.shape:
        raise ValueError("ray_o/ray_d must be [..., 3] of equal shape")
    n = o.shape[0]
    far = torch.empty(n, device=o.device, dtype=torch.float32)
    flag = torch.zeros(1, device=o.device, dtype=torch.int32)
    with torch.cuda.device(o.device):
        check(_lib.lib().nerfpp_intersect_sphere(_p(o), _p(d), n, _p(far), _p(flag), _stream()), "intersect_sphere")
    if deferred:
        return far.reshape(lead), UnboundedFlag(flag)
    UnboundedFlag(flag).raise_if_set()
    return far.reshape(lead)


_LINSPACE = {}


def linspace01(n, device):
    """torch.linspace(0,1,n) computed on the CPU exactly as the reference does
    (ddp_train_nerf.py:447 builds it on the host and moves it), cached per device."""
    key = (n, str(device))
    if key not in _LINSPACE:
        _LINSPACE[key] = torch.linspace(0., 1., n).to(device)
    return _LINSPACE[key]


def coarse_depths(near, far, n_samples, t_rand_fg=None, t_rand_bg=None):
    """ddp_train_nerf.py:441-449 / 168-175 fused with perturb_samples. near, far [n] -> fg_z, bg_z [n,S]."""
    near, far = _c(near, "near", 1), _c(far, "far", 1)
    n = near.shape[0]
    fg = torch.empty(n, n_samples, device=near.device, dtype=torch.float32)
    bg = torch.empty_like(fg)
    tf = _c(t_rand_fg, "t_rand_fg", 2) if t_rand_fg is not None else None
    tb = _c(t_rand_bg, "t_rand_bg", 2) if t_rand_bg is not None else None
    for t in (tf, tb):
        if t is not None and tuple(t.shape) != (n, n_samples):
            raise ValueError("t_rand must be [n, S]")
    with torch.cuda.device(near.device):
        check(_lib.lib().nerfpp_coarse_depths(_p(near), _p(far), _p(linspace01(n_samples, near.device)), n, n_samples,
                                              _p(tf), _p(tb), _p(fg), _p(bg), _stream()), "coarse_depths")
    return fg, bg


def perturb_samples(z_vals, t_rand=None):
    """ddp_train_nerf.py:69-78; the torch.rand_like draw (:75) stays in torch (H3) unless supplied."""
    z, lead = _rows(_c(z_vals, "z_vals"))
    if t_rand is None:
        t_rand = torch.rand_like(z)
    t, _ = _rows(_c(t_rand, "t_rand"))
    out = torch.empty_like(z)
    with torch.cuda.device(z.device):
        check(_lib.lib().nerfpp_perturb_samples(_p(z), _p(t), z.shape[0], z.shape[1], _p(out), _stream()), "perturb_samples")
    return out.reshape(lead + (z.shape[1],))


def _u_rows(u, n, n_new, device):
    if u.dim() == 1:
        u = _c(u, "u")
        return u, 0
    u, _ = _rows(_c(u, "u"))
    if u.shape != (n, n_new):
        raise ValueError("u must be [n, N_samples]")
    return u, n_new


def sample_pdf(bins, weights, N_samples, det=False, u=None, return_aux=False):
    """ddp_train_nerf.py:81-130. bins [..., M+1], weights [..., M] -> [..., N_samples].
    ``u`` overrides the draw (tests); otherwise linspace (det) or torch.rand (:102-107)."""
    b, lead = _rows(_c(bins, "bins"))
    w = _chk(weights, "weights").detach()
    M = w.shape[-1]
    w = w.reshape(-1, M)
    if w.stride(-1) != 1:
        w = w.contiguous()
    n = b.shape[0]
    if b.shape[1] != M + 1 or w.shape[0] != n:
        raise ValueError("bins must be [..., M+1] for weights [..., M]")
    if u is None:
        u = linspace01(N_samples, b.device) if det else torch.rand(n, N_samples, device=b.device)
    uu, u_ld = _u_rows(u, n, N_samples, b.device)
    out = torch.empty(n, N_samples, device=b.device, dtype=torch.float32)
    above = torch.empty(n, N_samples, device=b.device, dtype=torch.int32) if return_aux else None
    cdf = torch.empty(n, M + 1, device=b.device, dtype=torch.float32) if return_aux else None
    with torch.cuda.device(b.device):
        check(_lib.lib().nerfpp_sample_pdf(_p(b), M + 1, _p(w), w.stride(0) if n > 1 else M, _p(uu), u_ld, n, M,
                                           N_samples, _p(out), _p(above), _p(cdf), _stream()), "sample_pdf")
    out = out.reshape(lead + (N_samples,))
    return (out, cdf, above) if return_aux else out


def sample_cdf(bins, cdf, u):
    """Search + interpolation on a supplied cdf [n, M+1] (bit-exact contract). Returns (samples, above)."""
    b, c = _c(bins, "bins", 2), _c(cdf, "cdf", 2)
    n, M = c.shape[0], c.shape[1] - 1
    n_new = u.shape[-1]
    uu, u_ld = _u_rows(u, n, n_new, b.device)
    out = torch.empty(n, n_new, device=b.device, dtype=torch.float32)
    above = torch.empty(n, n_new, device=b.device, dtype=torch.int32)
    with torch.cuda.device(b.device):
        check(_lib.lib().nerfpp_sample_cdf(_p(b), M + 1, _p(c), M + 1, _p(uu), u_ld, n, M, n_new, _p(out), _p(above),
                                           _stream()), "sample_cdf")
    return out, above


def resample_merge(z_prev, w_prev, N_samples, det=False, u=None):
    """One cascade refinement, ddp_train_nerf.py:452-457: mids -> sample_pdf(w[1:-1]) -> sort(cat)."""
    z, w = _c(z_prev, "z_prev", 2), _c(w_prev, "w_prev", 2)
    n, sp = z.shape
    if w.shape != z.shape:
        raise ValueError("weights must match depths")
    if u is None:
        u = linspace01(N_samples, z.device) if det else torch.rand(n, N_samples, device=z.device)
    uu, u_ld = _u_rows(u, n, N_samples, z.device)
    out = torch.empty(n, sp + N_samples, device=z.device, dtype=torch.float32)
    with torch.cuda.device(z.device):
        check(_lib.lib().nerfpp_resample_merge(_p(z), _p(w), _p(uu), u_ld, n, sp, N_samples, _p(out), _stream()),
              "resample_merge")
    return out


def resample_merge_pair(fg_z_prev, fg_w_prev, bg_z_prev, bg_w_prev, N_samples, det=False, u_fg=None, u_bg=None):
    """Both refinements of one cascade level (ddp_train_nerf.py:452-457 foreground, :460-465 background) in one launch.
    Random draws, when not supplied, come from ONE torch.rand call ([2, n, N_samples]: row 0 foreground, row 1 background)."""
    fz, fw = _c(fg_z_prev, "fg_z_prev", 2), _c(fg_w_prev, "fg_w_prev", 2)
    bz, bw = _c(bg_z_prev, "bg_z_prev", 2), _c(bg_w_prev, "bg_w_prev", 2)
    n, sp = fz.shape
    if fw.shape != fz.shape or bz.shape != fz.shape or bw.shape != fz.shape:
        raise ValueError("foreground and background depths / weights must all be [n, S_prev]")
    if (u_fg is None) != (u_bg is None):
        raise ValueError("give both u_fg and u_bg or neither")
    if u_fg is None:
        if det:
            u_fg = u_bg = linspace01(N_samples, fz.device)
        else:
            u_fg, u_bg = torch.rand(2, n, N_samples, device=fz.device).unbind(0)
    uf, ld_f = _u_rows(u_fg, n, N_samples, fz.device)
    ub, ld_b = _u_rows(u_bg, n, N_samples, fz.device)
    if ld_f != ld_b:
        raise ValueError("u_fg and u_bg must have the same layout")
    out_f = torch.empty(n, sp + N_samples, device=fz.device, dtype=torch.float32)
    out_b = torch.empty_like(out_f)
    with torch.cuda.device(fz.device):
        check(_lib.lib().nerfpp_resample_merge_pair(_p(fz), _p(fw), _p(uf), _p(out_f), _p(bz), _p(bw), _p(ub), _p(out_b), ld_f, n, sp,
                                                    N_samples, _stream()), "resample_merge")
    return out_f, out_b


# ------------------------------------------------------------------------------------------------
# packed weights
# ------------------------------------------------------------------------------------------------
class PackedNet:
    """Cache of one MLPNet's repacked parameters (fp16 MMA tiles in issue order).

    Re-packed when any parameter's storage or ``_version`` changes -- optimizer steps and every in-place torch op bump
    ``_version``.  Writes that bypass the version counter (``param.data.copy_(...)``, EMA through ``.data``, external
    kernels) do NOT: call ``invalidate()`` after those (``NerfNet.invalidate_packed()`` does it for both nets).
    The packed copy is always written into the SAME allocation per (device, impl), so captured CUDA graphs that baked its
    address in stay valid.  Inside a stream capture nothing is packed by default (GraphedRenderStep refreshes the copy
    outside its graph before each replay); a graph that CONTAINS an optimizer step (GraphedTrainStep) captures under
    ``PackedNet.pack_in_capture`` so that the pack kernels are graph nodes and run on every replay."""
    pack_in_capture = False

    def __init__(self, is_bg):
        self.is_bg = int(is_bg)
        self._buf = {}
        self._key = {}

    def invalidate(self):
        self._key.clear()

    def get(self, tensors, impl):
        """tensors: 24 parameter tensors in LAYER_NAMES order (weight, bias interleaved per layer)."""
        for t in tensors:
            _chk(t, "parameter")                 # CPU parameters: NerfppError (there is no CPU path), before any CUDA query
        key = tuple((t.data_ptr(), t._version) for t in tensors)
        dev = tensors[0].device
        slot = (impl, str(dev))
        capturing = torch.cuda.is_current_stream_capturing()
        if (capturing and PackedNet.pack_in_capture) or (not capturing and self._key.get(slot) != key) or slot not in self._buf:
            buf = self._buf.get(slot)
            if buf is None:
                nbytes = _lib.lib().nerfpp_packed_bytes(self.is_bg, impl)
                if capturing:
                    raise NerfppError("PackedNet: first use inside a CUDA-graph capture; run one warm-up pass before capturing")
                buf = self._buf[slot] = torch.empty(nbytes, device=dev, dtype=torch.uint8)
            ps = net_params_struct(tensors)
            with torch.cuda.device(dev):
                check(_lib.lib().nerfpp_pack_weights(ctypes.byref(ps), self.is_bg, impl, _p(buf), _stream()), "pack_weights")
            self._key[slot] = None if capturing else key
        return self._buf[slot]


def net_params_struct(tensors):
    ps = _lib.NetParams()
    for l in range(_lib.NLAYERS):
        w, b = tensors[2 * l], tensors[2 * l + 1]
        for t in (w, b):
            _chk(t, "parameter")
            if not t.is_contiguous():
                raise ValueError("parameters must be contiguous")
        ps.w[l], ps.b[l] = w.data_ptr(), b.data_ptr()
    return ps


def check_param_shapes(tensors, is_bg):
    emb = 84 if is_bg else 63
    want = [(256, emb)] + [(256, 256)] * 4 + [(256, 256 + emb)] + [(256, 256)] * 2 + [(1, 256), (256, 256), (128, 283), (3, 128)]
    for l, shp in enumerate(want):
        if tuple(tensors[2 * l].shape) != shp or tuple(tensors[2 * l + 1].shape) != (shp[0],):
            raise ValueError("layer %s: expected weight %s, got %s -- only the reference's netdepth=8, netwidth=256, "
                             "max_freq_log2=10, max_freq_log2_viewdirs=4 shape is built (configs/kitti.txt:36-40)"
                             % (LAYER_NAMES[l], shp, tuple(tensors[2 * l].shape)))


# ------------------------------------------------------------------------------------------------
# A6-A11 forward (+ backward)
# ------------------------------------------------------------------------------------------------
def field_forward(packed, is_bg, ray_o, ray_d, z, impl=None):
    """Per-sample sigma/rgb of one net (embed + MLP). Returns sigma [n,S], rgb [n,S,3], depth_real [n,S] or None."""
    impl = default_field_impl() if impl is None else impl
    o, d, zz = _c(ray_o, "ray_o", 2), _c(ray_d, "ray_d", 2), _c(z, "z", 2)
    n, S = zz.shape
    sigma = torch.empty(n, S, device=zz.device, dtype=torch.float32)
    rgb = torch.empty(n, S, 3, device=zz.device, dtype=torch.float32)
    dr = torch.empty(n, S, device=zz.device, dtype=torch.float32) if is_bg else None
    with torch.cuda.device(zz.device):
        check(_lib.lib().nerfpp_field_forward(_p(packed), int(is_bg), impl, _p(o), _p(d), _p(zz), n, S, _p(sigma), _p(rgb),
                                              _p(dr), _stream()), "field_forward")
    return sigma, rgb, dr


def field_forward_train(packed, is_bg, ray_o, ray_d, z):
    """Training-mode field evaluation (tensor-core path): returns (sigma, rgb, depth_real, train_workspace)."""
    o, d, zz = _c(ray_o, "ray_o", 2), _c(ray_d, "ray_d", 2), _c(z, "z", 2)
    n, S = zz.shape
    L = _lib.lib()
    sigma = torch.empty(n, S, device=zz.device, dtype=torch.float32)
    rgb = torch.empty(n, S, 3, device=zz.device, dtype=torch.float32)
    dr = torch.empty(n, S, device=zz.device, dtype=torch.float32) if is_bg else None
    ws = torch.empty(int(L.nerfpp_field_train_workspace_bytes(n, S)), device=zz.device, dtype=torch.uint8)
    with torch.cuda.device(zz.device):
        check(L.nerfpp_field_forward_train(_p(packed), int(is_bg), _p(o), _p(d), _p(zz), n, S, _p(sigma), _p(rgb), _p(dr), _p(ws),
                                           _stream()), "field_forward")
    return sigma, rgb, dr, ws


def unpack_chunks(buf, n_tiles, n_chunks):
    """Decodes [n_tiles][n_chunks] SWIZZLE_128B operand chunks (uint8 tensor) into a [n_tiles*128, n_chunks*64] fp16 matrix
    (diagnostics / tests: the layout is documented in csrc/tc_common.cuh)."""
    x = buf.view(torch.float16).reshape(n_tiles, n_chunks, 16, 8, 8, 8)     # [tile, chunk, row>>3, row&7, unit, 8 halves]
    r7 = torch.arange(8, device=buf.device)
    src_unit = (torch.arange(8, device=buf.device)[None, :] ^ r7[:, None])  # logical unit u of row r lives at u ^ (r & 7)
    idx = src_unit[None, None, None, :, :, None].expand(n_tiles, n_chunks, 16, 8, 8, 8)
    y = torch.gather(x, 4, idx)
    return y.reshape(n_tiles, n_chunks, 128, 64).permute(0, 2, 1, 3).reshape(n_tiles * 128, n_chunks * 64)


def _alloc_outputs(n, sf, sb, device):
    shapes = dict(rgb=(n, 3), fg_weights=(n, sf), bg_weights=(n, sb), fg_dists=(n, sf), fg_rgb=(n, 3), fg_depth=(n,),
                  bg_rgb=(n, 3), bg_depth=(n,), bg_lambda=(n,), depth=(n,))
    outs = OrderedDict((k, torch.empty(shapes[k], device=device, dtype=torch.float32)) for k in RET_KEYS)
    st = _lib.RenderOut()
    for k in RET_KEYS:
        setattr(st, k, outs[k].data_ptr())
    return outs, st


def render_forward(packed_fg, packed_bg, ray_o, ray_d, fg_z_max, fg_z, bg_z, impl=None, keep_workspace=False, train=False):
    """NerfNet.forward (ddp_model.py:74-147) without autograd. Returns OrderedDict of the 10 keys
    (+ the per-sample workspace when keep_workspace; with ``train`` also the training workspace the backward reads)."""
    impl = default_field_impl() if impl is None else impl
    o, d = _c(ray_o, "ray_o", 2), _c(ray_d, "ray_d", 2)
    zmax, fz, bz = _c(fg_z_max, "fg_z_max", 1), _c(fg_z, "fg_z_vals", 2), _c(bg_z, "bg_z_vals", 2)
    n, sf = fz.shape
    sb = bz.shape[1]
    if o.shape != (n, 3) or d.shape != (n, 3) or zmax.shape != (n,) or bz.shape[0] != n:
        raise ValueError("inconsistent ray batch shapes")
    L = _lib.lib()
    ws = torch.empty(max(int(L.nerfpp_forward_workspace_bytes(n, sf, sb)), 1), device=fz.device, dtype=torch.uint8)
    outs, st = _alloc_outputs(n, sf, sb, fz.device)
    tws = None
    with torch.cuda.device(fz.device):
        if train:
            if impl != FIELD_TC:
                raise NerfppError("training (autograd) runs on the tensor-core field only (impl=FIELD_SIMT is inference-only)")
            tws = torch.empty(int(L.nerfpp_forward_train_workspace_bytes(n, sf, sb)) + 1024, device=fz.device, dtype=torch.uint8)
            pad = (-tws.data_ptr()) % 1024
            tws = tws[pad:]
            check(L.nerfpp_forward_train(_p(packed_fg), _p(packed_bg), _p(o), _p(d), _p(zmax), _p(fz), _p(bz), n, sf, sb,
                                         ctypes.byref(st), _p(ws), _p(tws), _stream()), "forward")
        else:
            check(L.nerfpp_forward(_p(packed_fg), _p(packed_bg), impl, _p(o), _p(d), _p(zmax), _p(fz), _p(bz), n, sf, sb,
                                   ctypes.byref(st), _p(ws), _stream()), "forward")
    if keep_workspace:
        return outs, ws, (o, d, zmax, fz, bz), tws
    return outs


def composite_backward(ray_d, fg_z_max, fg_z, bg_z, fg_sigma, fg_rgb, bg_sigma, bg_rgb, bg_depth_real, bg_lambda, grads):
    """Backward of the composite (autograd of ddp_model.py:95-134).  ``grads``: dict key -> upstream gradient for any of
    the ten output keys (missing/None = zero).  Returns d_fg_sigma [n,S], d_fg_rgb [n,S,3], d_bg_sigma, d_bg_rgb
    (background in the forward's flipped order)."""
    d, zmax, fz, bz = _c(ray_d, "ray_d", 2), _c(fg_z_max, "fg_z_max", 1), _c(fg_z, "fg_z", 2), _c(bg_z, "bg_z", 2)
    n, sf = fz.shape
    sb = bz.shape[1]
    ins = [_c(t, nm) for t, nm in ((fg_sigma, "fg_sigma"), (fg_rgb, "fg_rgb"), (bg_sigma, "bg_sigma"), (bg_rgb, "bg_rgb"),
                                   (bg_depth_real, "bg_depth_real"))]
    lam = _c(bg_lambda, "bg_lambda", 1)
    fwd = _lib.RenderOut()
    fwd.bg_lambda = lam.data_ptr()
    gst = _lib.RenderOut()
    keep = []
    for k in RET_KEYS:
        g = grads.get(k)
        if g is not None and k != "fg_dists":
            g = _c(g.float(), "grad_" + k)
            keep.append(g)
            setattr(gst, k, g.data_ptr())
    outs = [torch.empty(n, sf, device=fz.device), torch.empty(n, sf, 3, device=fz.device),
            torch.empty(n, sb, device=fz.device), torch.empty(n, sb, 3, device=fz.device)]
    with torch.cuda.device(fz.device):
        check(_lib.lib().nerfpp_composite_backward(_p(d), _p(zmax), _p(fz), _p(bz), *[_p(t) for t in ins], n, sf, sb,
                                                   ctypes.byref(fwd), ctypes.byref(gst), *[_p(t) for t in outs], _stream()),
              "composite_backward")
    return tuple(outs)


class _NerfppFunction(torch.autograd.Function):
    """Autograd node for NerfNet.forward: forward = nerfpp_forward, backward = nerfpp_backward."""

    @staticmethod
    def forward(ctx, model_cache, impl, ray_o, ray_d, fg_z_max, fg_z, bg_z, *params):
        if any(ctx.needs_input_grad[2:7]):
            # the reference's NerfNet.forward is differentiable w.r.t. the rays and depths too (pose / depth refinement
            # would use that); this path only produces PARAMETER gradients -- say so instead of returning zeros
            raise NerfppError("nerfpp_forward: gradients w.r.t. ray_o / ray_d / fg_z_max / fg_z_vals / bg_z_vals are not "
                              "implemented (parameter gradients only); detach those inputs")
        fg_t, bg_t = params[:24], params[24:]
        packed_fg = model_cache[0].get(fg_t, impl)
        packed_bg = model_cache[1].get(bg_t, impl)
        outs, ws, inputs, tws = render_forward(packed_fg, packed_bg, ray_o, ray_d, fg_z_max, fg_z, bg_z, impl, keep_workspace=True,
                                               train=True)
        ctx.impl, ctx.ws, ctx.tws = impl, ws, tws
        vals = tuple(outs.values())
        # the outputs go through save_for_backward, never a ctx attribute: an attribute would close the reference cycle
        # node -> output -> grad_fn -> node, and the workspaces (2.5-7.5 GB per call) would then wait for Python's cycle
        # collector instead of being returned to the allocator when the graph is freed
        ctx.save_for_backward(*params, *inputs, *vals)
        ctx.set_materialize_grads(False)      # outputs the loss does not use arrive as None, not as zero-filled tensors
        ctx.mark_non_differentiable(outs["fg_dists"])
        return vals

    @staticmethod
    def backward(ctx, *grads):
        from . import backward as B   # CUDA backward (nerfpp_backward); raises if the library lacks it
        saved = ctx.saved_tensors
        if ctx.tws is None:
            raise NerfppError("backward through this NerfNet.forward ran already; its training workspace is released after the "
                              "first backward (retain_graph=True is not supported, the reference trainer never uses it)")
        params, inputs, outs = saved[:48], saved[48:53], OrderedDict(zip(RET_KEYS, saved[53:]))
        pg = B.render_backward(params, inputs, outs, ctx.ws, ctx.tws, grads)
        ctx.ws = ctx.tws = None
        return (None, None, None, None, None, None, None) + tuple(pg)


def nerfpp_forward(model_cache, fg_tensors, bg_tensors, ray_o, ray_d, fg_z_max, fg_z, bg_z, impl=None):
    """Autograd-connected NerfNet.forward. Returns the reference's 10-key OrderedDict."""
    impl = default_field_impl() if impl is None else impl
    params = tuple(fg_tensors) + tuple(bg_tensors)
    if torch.is_grad_enabled() and any(p.requires_grad for p in params):
        vals = _NerfppFunction.apply(model_cache, impl, ray_o, ray_d, fg_z_max, fg_z, bg_z, *params)
        return OrderedDict(zip(RET_KEYS, vals))
    packed_fg = model_cache[0].get(tuple(fg_tensors), impl)
    packed_bg = model_cache[1].get(tuple(bg_tensors), impl)
    return render_forward(packed_fg, packed_bg, ray_o, ray_d, fg_z_max, fg_z, bg_z, impl)


# ------------------------------------------------------------------------------------------------
# A12-A14 losses
# ------------------------------------------------------------------------------------------------
_LOSS_TYPES = {None: DEPTH_NONE, "none": DEPTH_NONE, "mse": DEPTH_MSE, "l1": DEPTH_L1, "kl": DEPTH_KL}


def fused_loss(rgb, rgb_gt, depth=None, depth_sup=None, depth_loss_type=None, lambda_depth=0.0, fg_weights=None,
               fg_z=None, fg_dists=None, fg_z_max=None, kl_sigma=1.0):
    """img2mse + depth prior loss + total in one pass (ddp_train_nerf.py:481-493), no autograd.
    Returns a [4] device tensor: rgb_loss, depth_loss, rgb + lambda*depth, #valid rays."""
    typ = _LOSS_TYPES[depth_loss_type]
    r, g = _c(rgb, "rgb", 2), _c(rgb_gt, "rgb_gt", 2)
    n = r.shape[0]
    S = fg_z.shape[-1] if fg_z is not None else 1
    args = [(_c(t, nm) if t is not None else None) for t, nm in
            ((depth, "depth"), (depth_sup, "depth_sup"), (fg_weights, "fg_weights"), (fg_z, "fg_z"),
             (fg_dists, "fg_dists"), (fg_z_max, "fg_z_max"))]
    L = _lib.lib()
    out = torch.empty(4, device=r.device, dtype=torch.float32)
    ws = torch.empty(int(L.nerfpp_loss_workspace_bytes()), device=r.device, dtype=torch.uint8)
    with torch.cuda.device(r.device):
        check(L.nerfpp_loss(_p(r), _p(g), _p(args[0]), _p(args[1]), _p(args[2]), _p(args[3]), _p(args[4]), _p(args[5]),
                            n, S, typ, float(lambda_depth), float(kl_sigma), _p(out), _p(ws), _stream()), "loss")
    return out


METRIC_KEYS = ("mse", "psnr", "n_valid", "rmse", "rmse_log", "abs_diff", "abs_rel", "sq_rel")


def image_metrics(rgb, rgb_gt=None, depth=None, depth_gt=None, depth_scale=1.0, cap=80.0):
    """The per-image metrics of the reference's test loop (ddp_train_nerf.py:556-600) on device tensors of any shape
    ([H,W,3] / [H,W]); returns a dict of python floats (one device->host read)."""
    r = _c(rgb, "rgb").reshape(-1, 3) if rgb is not None else None
    n = r.shape[0] if r is not None else depth.numel()
    g = _c(rgb_gt, "rgb_gt").reshape(-1, 3) if rgb_gt is not None else None
    d = _c(depth, "depth").reshape(-1) if depth is not None else None
    dg = _c(depth_gt, "depth_gt").reshape(-1) if depth_gt is not None else None
    dev = (r if r is not None else d).device
    L = _lib.lib()
    out = torch.empty(8, device=dev, dtype=torch.float32)
    ws = torch.empty(2 * int(L.nerfpp_loss_workspace_bytes()), device=dev, dtype=torch.uint8)
    with torch.cuda.device(dev):
        check(L.nerfpp_image_metrics(_p(r), _p(g), _p(d), _p(dg), n, float(depth_scale), float(cap), _p(out), _p(ws), _stream()), "loss")
    return dict(zip(METRIC_KEYS, out.cpu().tolist()))
