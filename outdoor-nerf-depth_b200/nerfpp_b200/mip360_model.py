"""Config 3 (BASELINE.json configs[2]): the mipnerf360 model loop -- host mirror of ``Model.__call__`` / ``MLP.__call__``
(nerf-methods/mipnerf360/internal/models.py:75-330, 398-611) under ``configs/360.gin`` over the C ABI
(``mip360_field_forward`` and the A16 entry points).  Names follow the reference: ``Model``, ``NerfMLP``, ``PropMLP``,
``Rays`` (internal/utils.py), per-level ``renderings`` / ``ray_history`` dictionaries with the reference's keys.

The reference is JAX/flax: its parameters are a tree ``params[<module>]['Dense_i']['kernel' | 'bias']`` with kernels
stored [in, out].  ``MLP.load_flax`` takes exactly that tree (numpy arrays) so that a checkpoint of the reference can be
dropped in; ``MLP.init`` draws flax's ``he_uniform`` default.  Inference only: the 1024-wide network has a forward here
(render / evaluate / the proposal resampling cascade), no backward yet.

All arithmetic is in libnerfpp_b200.so; torch owns memory, the stream and the RNG draw of the jitter."""
import ctypes
import math
from collections import namedtuple

import torch

from . import _lib, mip360
from . import ops
from ._lib import Mip360MlpParams, NerfppError
from .ops import _c, _p, _stream, check      # ops.check counts the library's kernel launches

Rays = namedtuple("Rays", ("origins", "directions", "viewdirs", "radii", "near", "far"))   # internal/utils.py:60-77 (the fields the path reads)

NUM_FEATURES = 504          # 21 basis vectors x 12 degrees x (sin, cos): models.py:350-351, 379-380
DIR_FEATURES = 27           # pos_enc(viewdirs, 0, 4, append_identity=True)
BOTTLENECK, VIEW_WIDTH, SKIP = 256, 128, 4


def dense_shapes(net_depth, net_width, has_rgb):
    """(in, out) of Dense_0.. in flax's construction order (models.py:442-466, 508-597)."""
    shapes, cur = [], NUM_FEATURES
    for i in range(net_depth):
        shapes.append((cur, net_width))
        cur = net_width + (NUM_FEATURES if (i % SKIP == 0 and i > 0) else 0)
    shapes.append((cur, 1))
    if has_rgb:
        shapes += [(cur, BOTTLENECK), (BOTTLENECK + DIR_FEATURES, VIEW_WIDTH), (VIEW_WIDTH, 3)]
    return shapes


class MLP(object):
    """models.MLP as ``configs/360.gin`` configures it (warp_fn = contract, disable_density_normals = True, IPE degrees
    0..12, icosahedron basis).  ``disable_rgb`` = PropMLP."""

    def __init__(self, net_depth, net_width, disable_rgb, device, prec=False):
        self.net_depth, self.net_width, self.disable_rgb = int(net_depth), int(net_width), bool(disable_rgb)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise NerfppError("the mipnerf360 field only exists as CUDA kernels (no CPU fallback)")
        self.prec = bool(prec)
        self.shapes = dense_shapes(self.net_depth, self.net_width, not self.disable_rgb)
        self.params = None          # list of (kernel [in,out], bias [out]) CUDA fp32 tensors
        self._packed = None         # the fp16 operand images; ONE allocation for the object's lifetime (captured graphs hold its address)
        self._dirty = True
        self._ws = None

    # -- parameters ------------------------------------------------------------------------------------------------
    def init(self, seed=0):
        """flax Dense under the config: kernel he_uniform = U(+-sqrt(6 / fan_in)), bias zeros (models.py:353, 428-429)."""
        g = torch.Generator(device="cpu").manual_seed(int(seed))
        ps = []
        for fan_in, fan_out in self.shapes:
            lim = math.sqrt(6.0 / fan_in)
            ps.append(((torch.rand(fan_in, fan_out, generator=g) * 2 - 1) * lim, torch.zeros(fan_out)))
        return self.load(ps)

    def load(self, params):
        """params: sequence of (kernel [in, out], bias [out]) in Dense_0.. order (numpy arrays or tensors)."""
        if len(params) != len(self.shapes):
            raise ValueError("expected %d Dense layers, got %d" % (len(self.shapes), len(params)))
        out = []
        for (k, b), (fi, fo) in zip(params, self.shapes):
            k = torch.as_tensor(k, dtype=torch.float32).to(self.device).contiguous()
            b = torch.as_tensor(b, dtype=torch.float32).to(self.device).contiguous()
            if tuple(k.shape) != (fi, fo) or tuple(b.shape) != (fo,):
                raise ValueError("Dense kernel %s / bias %s do not match (%d, %d)" % (tuple(k.shape), tuple(b.shape), fi, fo))
            out.append((k, b))
        self.params = out
        self._dirty = True          # re-packed into the same buffer on the next use
        return self

    def load_flax(self, tree):
        """tree: {'Dense_0': {'kernel': [in,out], 'bias': [out]}, ...} -- one module of the reference's checkpoint."""
        return self.load([(tree["Dense_%d" % i]["kernel"], tree["Dense_%d" % i]["bias"]) for i in range(len(self.shapes))])

    def packed(self):
        if self.params is None:
            raise NerfppError("MLP has no parameters: call init() or load()")
        if self._dirty:
            L = _lib.lib()
            has_rgb = int(not self.disable_rgb)
            if self._packed is None:
                nbytes = L.mip360_mlp_packed_bytes(self.net_depth, self.net_width, has_rgb, int(self.prec))
                if nbytes < 0:
                    check(-1, "mip360_mlp_packed_bytes")
                self._packed = torch.zeros(int(nbytes), dtype=torch.uint8, device=self.device)
            buf = self._packed
            st = Mip360MlpParams()
            for i, (k, b) in enumerate(self.params):
                st.kernel[i], st.bias[i] = k.data_ptr(), b.data_ptr()
            with torch.cuda.device(self.device):
                check(L.mip360_mlp_pack(ctypes.byref(st), self.net_depth, self.net_width, has_rgb, int(self.prec), _p(buf), _stream()),
                      "mip360_mlp_pack")
            self._dirty = False
        return self._packed

    def _workspace(self, n_samples):
        L = _lib.lib()
        need = int(L.mip360_field_workspace_bytes(int(n_samples), self.net_depth, self.net_width, int(not self.disable_rgb), int(self.prec)))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    # -- one level: s_to_t -> cast_rays -> MLP (models.py:203-231) ----------------------------------------------------------
    def level(self, sdist, rays):
        """sdist [n, S+1] -> tdist [n, S+1], density [n, S], rgb [n, S, 3] (None for a PropMLP)."""
        sd = _c(sdist, "sdist", 2)
        n, S = sd.shape[0], sd.shape[1] - 1
        o, d, v = (_c(x, nm, 2) for x, nm in ((rays.origins, "origins"), (rays.directions, "directions"), (rays.viewdirs, "viewdirs")))
        rad, near, far = (_c(x, nm).reshape(-1) for x, nm in ((rays.radii, "radii"), (rays.near, "near"), (rays.far, "far")))
        if not (o.shape == d.shape == v.shape == (n, 3)) or not (rad.numel() == near.numel() == far.numel() == n):
            raise ValueError("rays do not match sdist's %d rows" % n)
        tdist = torch.empty(n, S + 1, device=sd.device, dtype=torch.float32)
        density = torch.empty(n, S, device=sd.device, dtype=torch.float32)
        rgb = None if self.disable_rgb else torch.empty(n, S, 3, device=sd.device, dtype=torch.float32)
        if n == 0 or S == 0:        # an empty batch is a no-op (the reference returns empty arrays)
            return tdist, density, rgb
        ws = self._workspace(n * S)
        with torch.cuda.device(sd.device):
            check(_lib.lib().mip360_field_forward(_p(self.packed()), self.net_depth, self.net_width, int(not self.disable_rgb), int(self.prec),
                                                  _p(sd), _p(near), _p(far), _p(o), _p(d), _p(v), _p(rad), n, S, _p(tdist), _p(density),
                                                  _p(rgb), _p(ws), _stream()), "mip360_field_forward")
        # encode + one GEMM per Dense layer + density head (+ bottleneck, view layer, rgb head); check() counted one
        if self.net_depth == 4 and self.net_width == 256 and self.disable_rgb:
            ops.LAUNCHES[0] += 1          # encode + the fused PropMLP chain kernel
        else:
            ops.LAUNCHES[0] += self.net_depth + 1 + (0 if self.disable_rgb else 3)
        return tdist, density, rgb


def NerfMLP(device, net_depth=8, net_width=1024, prec=False):      # configs/360.gin:16-19
    return MLP(net_depth, net_width, False, device, prec)


def PropMLP(device, net_depth=4, net_width=256, prec=False):       # configs/360.gin:10-14
    return MLP(net_depth, net_width, True, device, prec)


def resample_logits(sdist, weights, anneal, resample_padding=0.0):
    """models.py:171-185: where(sdist[1:] > sdist[:-1], anneal * log(weights + padding), -inf)."""
    sd, w = _c(sdist, "sdist", 2), _c(weights, "weights", 2)
    n, M = w.shape
    if sd.shape != (n, M + 1):
        raise ValueError("sdist must be [n, M+1] for weights [n, M]")
    out = torch.empty_like(w)
    with torch.cuda.device(w.device):
        check(_lib.lib().mip360_resample_logits(_p(sd), _p(w), n, M, float(anneal), float(resample_padding), _p(out), _stream()),
              "mip360_resample_logits")
    return out


def resample_level(t, weights, anneal, resample_padding, num_samples, u=None, jitter=None, domain=(0.0, 1.0)):
    """models.py:171-200 in one launch: logits from the weights, then stepfun.sample_intervals.  ``t`` [n, M+1] / ``weights``
    [n, M] may be row-strided VIEWS (the dilated histogram's [..., 1:-1] slices).  ``u`` [n, Ns]: explicit ordinates; else
    ``jitter`` [n] in [0,1) (single_jitter) or None (the deterministic centres)."""
    for x, nm in ((t, "t"), (weights, "weights")):
        if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1):
            raise NerfppError("%s must be a float32 CUDA matrix with unit column stride" % nm)
    n, M = weights.shape
    if t.shape != (n, M + 1):
        raise ValueError("t must be [n, M+1] for weights [n, M]")
    dev = t.device
    jit, mj = None, 0.0
    if u is not None:
        uu, u_ld = _c(u, "u", 2), num_samples
    elif jitter is not None:
        uu, u_ld = mip360.jitter_base(num_samples, dev), 0
        jit = _c(jitter, "jitter").reshape(-1)
        mj = mip360.max_jitter(num_samples)
    else:
        uu, u_ld = mip360.centers_u(num_samples, dev), 0
    out = torch.empty(n, num_samples + 1, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        check(_lib.lib().mip360_resample_level(_p(t), t.stride(0), _p(weights), weights.stride(0), n, M, float(anneal), float(resample_padding),
                                               _p(uu), u_ld, _p(jit), float(mj), num_samples, float(domain[0]), float(domain[1]), _p(out),
                                               _stream()), "mip360_resample_level")
    return out


class Model(object):
    """models.Model (models.py:47-330) with the gin bindings of configs/360.gin: raydist_fn = reciprocal,
    opaque_background = True, three levels (64, 64 proposal intervals by one shared PropMLP, 32 by the NerfMLP), dilation
    0.0025 + 0.5 / prod(samples so far), annealing slope 10, single_jitter, bg_intensity_range (1, 1)."""

    num_prop_samples, num_nerf_samples, num_levels = 64, 32, 3
    anneal_slope, dilation_multiplier, dilation_bias, resample_padding = 10.0, 0.5, 0.0025, 0.0
    single_jitter, opaque_background, bg_intensity = True, True, 1.0

    def __init__(self, device, nerf_mlp=None, prop_mlp=None, prec=False):
        self.device = torch.device(device)
        self.nerf_mlp = nerf_mlp if nerf_mlp is not None else NerfMLP(device, prec=prec)
        self.prop_mlp = prop_mlp if prop_mlp is not None else PropMLP(device, prec=prec)

    def init(self, seed=0):
        self.nerf_mlp.init(seed)
        self.prop_mlp.init(seed + 1)
        return self

    def load_flax(self, params):
        """params: the reference checkpoint's ``params`` tree (keys 'NerfMLP_0', 'PropMLP_0')."""
        self.nerf_mlp.load_flax(params["NerfMLP_0"])
        self.prop_mlp.load_flax(params["PropMLP_0"])
        return self

    def __call__(self, rng, rays, train_frac=1.0, compute_extras=True, u_levels=None, on_level=None):
        """-> (renderings, ray_history), one entry per level.  ``rng``: None (deterministic interval centres) or a
        torch.Generator / True (torch's default CUDA generator) for the jitter; ``on_level(i, rendering, history_entry)`` is
        called as soon as a level is rendered (GraphedModelStep forks that level's loss kernels from it); ``u_levels`` overrides the per-level inverse-CDF ordinates (tests)."""
        near, far = _c(rays.near, "near").reshape(-1, 1), _c(rays.far, "far").reshape(-1, 1)
        n = near.shape[0]
        # the initial interval [0, 1] with weight 1 (models.py:129-136) and the proposal levels' all-zero colours are constants
        # of the batch size: built once, not once per call (five fills and a cat fewer per step)
        key = (n, str(near.device))
        if getattr(self, "_const_key", None) != key:
            self._const = (torch.tensor([0.0, 1.0], device=near.device).repeat(n, 1).contiguous(), torch.ones(n, 1, device=near.device),
                           torch.zeros(n, self.num_prop_samples, 3, device=near.device))
            self._const_key = key
        sdist, weights, zero_rgb = self._const
        prod_num_samples = 1
        renderings, ray_history = [], []
        for i_level in range(self.num_levels):
            is_prop = i_level < self.num_levels - 1
            num_samples = self.num_prop_samples if is_prop else self.num_nerf_samples
            dilation = self.dilation_bias + self.dilation_multiplier * 1.0 / prod_num_samples
            prod_num_samples *= num_samples
            if i_level > 0:
                sdist, weights = mip360.max_dilate_weights(sdist, weights, dilation, domain=(0.0, 1.0), renormalize=True)
                sdist, weights = sdist[..., 1:-1], weights[..., 1:-1]          # views: the resampling kernel reads them strided
            s = self.anneal_slope
            anneal = (s * train_frac) / ((s - 1) * train_frac + 1) if s > 0 else 1.0
            # logits + ordinates + sample_intervals in one launch (models.py:171-200)
            if u_levels is not None and u_levels[i_level] is not None:
                sdist = resample_level(sdist, weights, anneal, self.resample_padding, num_samples, u=u_levels[i_level])
            elif rng is None:
                sdist = resample_level(sdist, weights, anneal, self.resample_padding, num_samples)
            else:
                d = 1 if self.single_jitter else num_samples
                if not self.single_jitter:
                    raise NotImplementedError("per-sample jitter: pass u_levels (configs/360.gin uses single_jitter)")
                jit = torch.rand(n, d, device=self.device, generator=rng if isinstance(rng, torch.Generator) else None)
                sdist = resample_level(sdist, weights, anneal, self.resample_padding, num_samples, jitter=jit)
            mlp = self.prop_mlp if is_prop else self.nerf_mlp
            tdist, density, rgb = mlp.level(sdist, rays)
            weights = mip360.compute_alpha_weights(density, tdist, rays.directions, opaque_background=self.opaque_background, weights_only=True)[0]
            if rgb is None:
                rgb = zero_rgb                                                 # disable_rgb: zeros (models.py:511-512)
            rendering = mip360.volumetric_rendering(rgb, weights, tdist, self.bg_intensity, far.reshape(-1), compute_extras)
            renderings.append(rendering)
            ray_history.append(dict(density=density, rgb=rgb, sdist=sdist, tdist=tdist, weights=weights))
            if on_level is not None:
                on_level(i_level, rendering, ray_history[-1])
        return renderings, ray_history


RENDER_KEYS = ("rgb", "acc", "distance_mean", "depth", "distance_percentile_5", "distance_median", "distance_percentile_95")
RENDER_WIDTHS = (3, 1, 1, 1, 1, 1, 1)


def band(num_rays, world, rank):
    """Rank ``rank``'s contiguous band [lo, hi) of ``num_rays`` rays and the common band length (the last may be short)."""
    per = (num_rays + world - 1) // world
    return min(rank * per, num_rays), min((rank + 1) * per, num_rays), per


def render_image(model, rays, render_chunk_size=16384, train_frac=1.0, process_group=None, render_chunk=None, device=None):
    """models.render_image (models.py:625-703) in test mode: every pixel of an image through ``model`` in chunks of
    ``render_chunk_size`` rays (configs.py: render_chunk_size = 16384), the LAST level's 2-D buffers reshaped to
    [height, width, ...].  ``rays``: a ``Rays`` of [height, width, C] CUDA tensors.  With a ``process_group`` (one process per
    GPU) every rank renders its contiguous band of the image -- the reference shards its chunks over its devices
    (models.py:665-669) -- and ONE all-gather of the packed keys per image collects the bands on every rank.
    ``render_chunk(chunk_rays) -> last level's rendering dict`` replaces the model call (tests of the host logic on CPU)."""
    height, width = rays.origins.shape[:2]
    num_rays = height * width
    if render_chunk is None:
        flat = Rays(*(_c(r, "rays").reshape(num_rays, -1) for r in rays))
        device = model.device

        def render_chunk(chunk):
            renderings, _ = model(None, chunk, train_frac=train_frac, compute_extras=True)
            return renderings[-1]
    else:
        flat = Rays(*(r.reshape(num_rays, -1) for r in rays))
    world, rank = 1, 0
    if process_group is not None:
        import torch.distributed as dist
        world, rank = dist.get_world_size(process_group), dist.get_rank(process_group)
    lo, hi, per = band(num_rays, world, rank)
    packed = torch.zeros(per, sum(RENDER_WIDTHS), device=device)
    with torch.no_grad():
        for idx0 in range(lo, hi, render_chunk_size):
            idx1 = min(idx0 + render_chunk_size, hi)
            last = render_chunk(Rays(*(r[idx0:idx1] for r in flat)))
            packed[idx0 - lo:idx1 - lo] = torch.cat([last[k].reshape(idx1 - idx0, -1) for k in RENDER_KEYS], dim=-1)
    if world > 1:
        import torch.distributed as dist
        gathered = torch.empty(world * per, sum(RENDER_WIDTHS), device=device)
        dist.all_gather_into_tensor(gathered, packed, group=process_group)
        packed = gathered
    packed = packed[:num_rays]
    out, off = {}, 0
    for k, w in zip(RENDER_KEYS, RENDER_WIDTHS):
        out[k] = packed[:, off:off + w].reshape((height, width) + ((w,) if w > 1 else ()))
        off += w
    return out


IN_KEYS = (("origins", 3), ("directions", 3), ("viewdirs", 3), ("radii", 1), ("near", 1), ("far", 1), ("rgb", 3), ("disps_sup", 1))
LOSS_KEYS = ("mse_0", "mse_1", "mse_2", "depth_0", "depth_1", "depth_2", "interlevel", "distortion")


class GraphedModelStep(object):
    """One evaluation of the mipnerf360 trainer's forward (train_utils.py:284-336: ``model.apply`` -> compute_data_loss with
    the depth prior -> interlevel_loss -> distortion_loss) captured ONCE into a CUDA graph: ~45 library kernels, torch's
    rand / cat nodes, and with ``host_io`` one H2D memcpy node in front (the batch: 16 floats per ray from a pinned staging
    buffer) and one D2H node behind (rgb | depth | the eight loss terms).  ``step(batch)`` -> dict(rgb [n,3], depth [n],
    losses [8] in LOSS_KEYS order); host tensors (views of a pinned buffer) when host_io, else device tensors."""

    def __init__(self, model, n_rays, train_frac=1.0, jitter=True, depth_loss_type="kl", depth_sigma=0.01, depth_scale=1.0,
                 host_io=True, warmup=2):
        self.model, self.n, self.host_io = model, int(n_rays), bool(host_io)
        dev, n = model.device, self.n
        total = sum(w for _, w in IN_KEYS) * n
        self._in_dev = torch.zeros(total, device=dev)
        self._in_host = torch.zeros(total).pin_memory() if host_io else None

        def views(flat):
            out, off = {}, 0
            for k, w in IN_KEYS:
                out[k] = flat[off:off + n * w].view(n, w)
                off += n * w
            return out
        self.dev_in = views(self._in_dev)
        self.host_in = views(self._in_host) if host_io else None
        self.dev_in["directions"][:, 2] = 1.0          # a valid batch for the warm-up passes
        self.dev_in["viewdirs"][:, 2] = 1.0
        self.dev_in["radii"].fill_(1e-3)
        self.dev_in["near"].fill_(0.2)
        self.dev_in["far"].fill_(1e6)
        if host_io:
            self._in_host.copy_(self._in_dev)
        self._n_out = 4 * n + len(LOSS_KEYS)
        self._side = torch.cuda.Stream(device=dev)
        self._out_host = torch.zeros(self._n_out).pin_memory() if host_io else None
        sigma = float(depth_sigma) * float(depth_scale)                      # train_utils.py:125

        def body():
            if host_io:
                self._in_dev.copy_(self._in_host, non_blocking=True)
            d = self.dev_in
            rays = Rays(d["origins"], d["directions"], d["viewdirs"], d["radii"], d["near"], d["far"])
            # each level's data terms are forked onto a side stream as soon as the level is rendered: they run beside the
            # next level's field kernels (and the last level's beside the interlevel / distortion terms) instead of in a
            # serial tail of ~20 small launches
            main = torch.cuda.current_stream(dev)
            mses, dls = [None] * 3, [None] * 3

            def level_losses(i, r, h):
                ev = torch.cuda.Event()
                ev.record(main)
                with torch.cuda.stream(self._side):
                    self._side.wait_event(ev)
                    mses[i] = ops.fused_loss(r["rgb"], d["rgb"], depth_loss_type=None)[0:1]          # lossmult = 1: mean squared residual
                    if depth_loss_type == "kl":
                        dls[i] = mip360.depth_loss(h["weights"], h["tdist"], d["disps_sup"].reshape(-1), r["distance_mean"], sigma,
                                                   d["directions"], "kl").reshape(1)
                    else:
                        dls[i] = mip360.depth_point_loss(r["distance_mean"], d["disps_sup"].reshape(-1), depth_loss_type).reshape(1)

            renderings, history = model(True if jitter else None, rays, train_frac=train_frac, compute_extras=True, on_level=level_losses)
            inter = mip360.interlevel_loss(history).reshape(1)
            dist = mip360.distortion_loss(history).reshape(1)
            main.wait_stream(self._side)
            packed = torch.cat([renderings[-1]["rgb"].reshape(-1), renderings[-1]["depth"].reshape(-1)] + mses + dls + [inter, dist])
            if host_io:
                self._out_host.copy_(packed, non_blocking=True)
            return packed

        with torch.no_grad():
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):          # warm-up off the capture: library init, weight packing, allocator
                for _ in range(max(int(warmup), 1)):
                    body()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            before = ops.LAUNCHES[0]
            with torch.cuda.graph(self.graph):
                self._packed = body()
            self.kernels_per_replay = ops.LAUNCHES[0] - before
        self._done = torch.cuda.Event() if host_io else None

    def _split(self, flat):
        n = self.n
        return dict(rgb=flat[:3 * n].view(n, 3), depth=flat[3 * n:4 * n], losses=flat[4 * n:4 * n + len(LOSS_KEYS)])

    def launch(self, batch=None):
        if batch is not None:
            dst = self.host_in if self.host_io else self.dev_in
            for k in dst:
                if batch[k] is not dst[k]:
                    dst[k].copy_(batch[k].reshape(dst[k].shape), non_blocking=True)
        self.model.nerf_mlp.packed()
        self.model.prop_mlp.packed()
        self.graph.replay()
        if self.host_io:
            self._done.record()

    def fetch(self):
        if not self.host_io:
            return self._split(self._packed)
        self._done.synchronize()
        return self._split(self._out_host)

    def __call__(self, batch=None):
        self.launch(batch)
        return self.fetch()
