"""One pass of the hot path (render_rays: sampling -> field -> composite -> losses, every cascade level) captured once
into a CUDA graph and replayed per step.

The path is ~14 kernels of 3-460 us behind a Python caller: replaying one graph removes the per-launch host cost and the
inter-kernel gaps, and with ``host_io=True`` the step's host->device and device->host transfers are graph nodes as well
(ONE copy each way between pinned staging buffers and packed device buffers), so a step that starts and ends in host
memory -- what the reference trainer's ``ray_batch[key].to(rank)`` / ``.item()`` lines amount to
(ddp_train_nerf.py:423-427,474-496) -- costs one ``cudaGraphLaunch`` and one stream synchronisation.

Nothing here adds arithmetic: the captured calls are exactly render.render_rays'.  Weights are read from the packed
buffers of each model's PackedNet cache, whose addresses are stable; ``__call__`` re-packs (outside the graph) when an
optimizer step has bumped a parameter's version."""
from collections import OrderedDict

import torch

from . import ops
from ._lib import NerfppError
from .render import render_rays

IN_KEYS = (("ray_o", 3), ("ray_d", 3), ("min_depth", 1), ("rgb", 3), ("depth_sup", 1))


class GraphedRenderStep(object):
    """step = GraphedRenderStep(models, n_rays, ...);  out = step(batch)  or  step.host_in[...] = ...; out = step()

    ``host_io=True``:  ``batch`` (or ``step.host_in``) holds HOST tensors; returns HOST tensors (views of one pinned
                       buffer, valid until the next call): rgb [n,3], depth [n], losses [levels,4].
    ``host_io=False``: ``batch`` holds DEVICE tensors (copied into the graph's static inputs unless they already are
                       ``step.dev_in``); returns device tensors, no synchronisation."""

    def __init__(self, models, n_rays, cascade_samples=(64, 128), train=True, depth_loss_type="mse", lambda_depth=0.1,
                 depth_sigma=0.01, depth_scale=1.0, with_rgb=True, with_depth_sup=True, host_io=True, device=None, warmup=2):
        self.models, self.n, self.host_io = list(models), int(n_rays), bool(host_io)
        dev = torch.device(device) if device is not None else next(models[0].parameters()).device
        if dev.type != "cuda":
            raise NerfppError("GraphedRenderStep: CUDA device required (there is no CPU path)")
        self.device = dev
        keys = [(k, w) for k, w in IN_KEYS if (k not in ("rgb",) or with_rgb) and (k != "depth_sup" or (with_rgb and with_depth_sup))]
        n = self.n
        total = sum(w for _, w in keys) * n
        self._in_dev = torch.zeros(total, device=dev)
        self._in_host = torch.zeros(total).pin_memory() if host_io else None

        def views(flat):
            out, off = OrderedDict(), 0
            for k, w in keys:
                out[k] = flat[off:off + n * w].view((n, w) if w > 1 else (n,))
                off += n * w
            return out
        self.dev_in = views(self._in_dev)
        self.host_in = views(self._in_host) if host_io else None
        self.dev_in["ray_d"][:, 2] = 1.0            # a valid ray for the warm-up passes (origin 0, direction +z)
        self.dev_in["min_depth"].fill_(1e-4)
        if host_io:
            self._in_host.copy_(self._in_dev)
        nl = len(cascade_samples)
        self._n_out = 4 * n + 4 * nl + 1            # rgb | depth | losses | out-of-sphere flag (as float)
        self._out_host = torch.zeros(self._n_out).pin_memory() if host_io else None
        kw = dict(cascade_samples=tuple(cascade_samples), train=train, depth_loss_type=depth_loss_type if with_depth_sup else None,
                  lambda_depth=lambda_depth, depth_sigma=depth_sigma, defer_unbounded_check=True)
        self._scale = float(depth_scale)

        def body():
            if host_io:
                self._in_dev.copy_(self._in_host, non_blocking=True)
            b = dict(self.dev_in)
            b["depth_scale"] = self._scale
            res = render_rays(self.models, b, **kw)
            ret = res["levels"][-1][0]
            flag = res["unbounded"].flag
            parts = [ret["rgb"].reshape(-1), ret["depth"].reshape(-1)]
            parts += [l.reshape(-1) for l in res.get("losses", [])] or [torch.zeros(4 * nl, device=dev)]
            parts.append(flag.view(torch.float32))       # the int32 flag's bits ride along (read back as int32): no convert kernel
            packed = torch.cat(parts)
            if host_io:
                self._out_host.copy_(packed, non_blocking=True)
            return res, packed

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
                self.res, self._packed = body()
            self.kernels_per_replay = ops.LAUNCHES[0] - before      # kernels of libnerfpp_b200.so in the graph
        self.replays = 0
        self._done = torch.cuda.Event() if host_io else None

    def _refresh_weights(self):
        impl = ops.default_field_impl()
        for m in self.models:
            net = m.nerf_net if hasattr(m, "nerf_net") else m
            net._packed[0].get(net.fg_net.tensors(), impl)
            net._packed[1].get(net.bg_net.tensors(), impl)

    def _split(self, flat):
        n, nl = self.n, (self._n_out - 4 * self.n - 1) // 4
        return OrderedDict(rgb=flat[:3 * n].view(n, 3), depth=flat[3 * n:4 * n], losses=flat[4 * n:4 * n + 4 * nl].view(nl, 4))

    def launch(self, batch=None):
        """Enqueues one step on the CURRENT stream without waiting for it: stage the inputs (``batch`` or whatever the
        caller wrote into ``host_in`` / ``dev_in``), re-pack weights if an optimizer touched them, replay the graph."""
        if batch is not None:
            dst = self.host_in if self.host_io else self.dev_in
            for k in dst:
                if batch[k] is not dst[k]:
                    dst[k].copy_(batch[k], non_blocking=True)
        self._refresh_weights()
        self.graph.replay()
        self.replays += 1
        ops.LAUNCHES[0] += self.kernels_per_replay
        if self.host_io:
            self._done.record()

    def fetch(self):
        """The results of the last ``launch``.  host_io: waits for that replay (its D2H node is its last node), raises
        the reference's out-of-sphere exception if the flag is set, returns views of the pinned result buffer."""
        if not self.host_io:
            return self._split(self._packed)
        self._done.synchronize()
        if int(self._out_host[-1:].view(torch.int32)[0]) != 0:
            raise Exception(ops.UNBOUNDED_MSG)
        return self._split(self._out_host)

    def __call__(self, batch=None):
        self.launch(batch)
        return self.fetch()

    def check_unbounded(self):
        """host_io=False only: reads the out-of-sphere flag of the last replay (one host sync)."""
        if int(self._packed[-1:].view(torch.int32)[0]) != 0:
            raise Exception(ops.UNBOUNDED_MSG)


class PipelinedRenderStep(object):
    """``depth`` host-to-host GraphedRenderSteps on their own streams, used round-robin: while step i computes, step
    i+1's inputs are staged and copied (copy engine) and step i-1's results are read on the host, so the GPU never waits
    for the host between steps -- the prefetching a trainer's data loader does.

        pipe.submit(batch_0)
        for b in batches[1:]:
            pipe.submit(b); out = pipe.result()      # result of the OLDEST outstanding step (host views, valid until
        out = pipe.result()                          # that slot is submitted to again)

    ``after_launch(slot_index, step)`` (optional) runs on the slot's stream right behind the replay -- bench.py issues the
    N > 1 all-gather of the rendered tile there."""

    def __init__(self, models, n_rays, depth=2, device=None, after_launch=None, **kw):
        kw["host_io"] = True
        self.slots = [GraphedRenderStep(models, n_rays, device=device, **kw) for _ in range(int(depth))]
        self.device = self.slots[0].device
        self.streams = [torch.cuda.Stream(device=self.device) for _ in self.slots]
        self.after_launch = after_launch
        self._next, self._pending = 0, []
        self._params = [p for m in models for p in m.parameters()]
        self._wkey = None

    @property
    def kernels_per_replay(self):
        return self.slots[0].kernels_per_replay

    def submit(self, batch=None):
        if len(self._pending) == len(self.slots):
            raise NerfppError("PipelinedRenderStep: %d steps outstanding; fetch a result() first" % len(self.slots))
        # the slots share the packed weight buffers, which launch() re-packs IN PLACE after an optimizer step: steps
        # still in flight must have finished reading them first
        wkey = tuple(p._version for p in self._params)
        if wkey != self._wkey:
            for j in self._pending:
                self.slots[j]._done.synchronize()
            self._wkey = wkey
        i = self._next
        self._next = (i + 1) % len(self.slots)
        st = self.streams[i]
        st.wait_stream(torch.cuda.current_stream(self.device))     # behind whatever the caller has enqueued (weight updates)
        with torch.cuda.stream(st):
            self.slots[i].launch(batch)
            if self.after_launch is not None:
                self.after_launch(i, self.slots[i])
                self.slots[i]._done.record()
        self._pending.append(i)
        return i

    def result(self):
        if not self._pending:
            raise NerfppError("PipelinedRenderStep: nothing submitted")
        return self.slots[self._pending.pop(0)].fetch()


class GraphedTrainStep(object):
    """One optimisation step of the reference trainer (ddp_train_nerf.py:432-498: for every cascade level sample ->
    NerfNet.forward -> rgb MSE + lambda * depth loss -> backward -> Adam) captured ONCE into a CUDA graph.

        step = GraphedTrainStep(models, n_rays, depth_loss_type="mse", lambda_depth=0.1, lr=5e-4)
        step.dev_in["ray_o"].copy_(...) ...           # or step(batch) with device tensors
        losses = step()                               # device tensor [levels]: the total loss of each level, no sync

    Eagerly the step is ~60 kernel launches plus autograd bookkeeping per level from Python, and the GPU idles between
    them (measured 0.9 ms of a 10.8 ms step); the graph also contains the weight re-packing that follows each optimizer
    step (``PackedNet.pack_in_capture``).  ``models[m]`` are ddp_model.NerfNetWithAutoExpo; the optimizers are created here
    (``torch.optim.Adam(capturable=True)``, as the trainer's ``Adam(lr)`` :325) and exposed as ``step.optimizers``.

    ``process_group`` (N > 1): the parameter gradients of a level -- ONE flat fp32 buffer per backward (backward.py) -- are
    averaged across ranks with a single NCCL all-reduce per level inside the graph, replacing DDP-over-gloo (:298,323)."""

    def __init__(self, models, n_rays, cascade_samples=(64, 128), depth_loss_type="mse", lambda_depth=0.1, depth_sigma=0.01,
                 depth_scale=1.0, lr=5e-4, device=None, process_group=None, warmup=3, batch=None):
        from . import backward as B
        from . import losses as LS
        self.models, self.n = list(models), int(n_rays)
        dev = torch.device(device) if device is not None else next(models[0].parameters()).device
        if dev.type != "cuda":
            raise NerfppError("GraphedTrainStep: CUDA device required (there is no CPU path)")
        self.device = dev
        n = self.n
        keys = list(IN_KEYS)
        self._in_dev = torch.zeros(sum(w for _, w in keys) * n, device=dev)
        self.dev_in, off = OrderedDict(), 0
        for k, w in keys:
            self.dev_in[k] = self._in_dev[off:off + n * w].view((n, w) if w > 1 else (n,))
            off += n * w
        self.dev_in["ray_d"][:, 2] = 1.0              # a valid batch for the warm-up passes: origin 0, direction +z,
        self.dev_in["min_depth"].fill_(1e-4)          # grey pixels, a depth prior inside the sphere (no empty-mask NaN)
        self.dev_in["rgb"].fill_(0.5)
        self.dev_in["depth_sup"].fill_(0.5)
        if batch is not None:
            for k in self.dev_in:
                self.dev_in[k].copy_(batch[k])
        # the trainer's Adam(lr) (:325) in its single-kernel form: same update rule, one launch per level instead of ~10
        self.optimizers = [torch.optim.Adam(m.parameters(), lr=lr, capturable=True, fused=True) for m in self.models]
        self.process_group = process_group
        world = 1
        if process_group is not None:
            import torch.distributed as dist
            world = dist.get_world_size(process_group)
        sig = float(depth_sigma) * float(depth_scale)
        cascade = tuple(cascade_samples)

        def body():
            b = self.dev_in
            far, flag = ops.intersect_sphere(b["ray_o"], b["ray_d"], deferred=True)
            fg_z = bg_z = ret = None
            losses = []
            for m, S in enumerate(cascade):
                if m == 0:
                    t = torch.rand(2, n, S, device=dev)
                    fg_z, bg_z = ops.coarse_depths(b["min_depth"], far, S, t[0], t[1])
                else:
                    fg_z, bg_z = ops.resample_merge_pair(fg_z, ret["fg_weights"].detach(), bg_z, ret["bg_weights"].detach(), S)
                self.optimizers[m].zero_grad(set_to_none=True)
                ret = self.models[m](b["ray_o"], b["ray_d"], far, fg_z, bg_z)
                loss = torch.mean((ret["rgb"] - b["rgb"]) * (ret["rgb"] - b["rgb"]))        # img2mse, utils.py:12-14
                if depth_loss_type == "kl":           # depth_loss.py:20-44 / :4-18 (the drop-in module's autograd nodes)
                    loss = loss + lambda_depth * LS.DepthKLLoss.apply(ret["fg_weights"], b["depth_sup"], fg_z, ret["fg_dists"], sig, far)
                elif depth_loss_type in ("mse", "l1"):
                    typ = LS.DEPTH_MSE if depth_loss_type == "mse" else LS.DEPTH_L1
                    loss = loss + lambda_depth * LS.DepthPointLoss.apply(b["depth_sup"], ret["depth"], typ)
                loss.backward()
                if world > 1:
                    import torch.distributed as dist
                    flat = B.LAST_FLAT_GRADS[0]            # all 48 gradients of this level's net, one buffer
                    params = list(self.models[m].parameters())
                    if all(p.grad is not None and p.grad.untyped_storage().data_ptr() == flat.untyped_storage().data_ptr() for p in params):
                        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=process_group)      # ONE collective: p.grad are views of it
                    else:                                  # autograd copied the gradients (it normally adopts the views)
                        for p in params:
                            if p.grad is not None:
                                dist.all_reduce(p.grad, op=dist.ReduceOp.AVG, group=process_group)
                self.optimizers[m].step()
                losses.append(loss.detach())
            return torch.stack(losses), flag.flag

        # the warm-up passes below are real optimisation steps: remember the weights and undo them afterwards (in place --
        # the graph has the parameter and optimizer-state addresses baked in)
        saved = [[p.detach().clone() for p in m.parameters()] for m in self.models]
        ops.PackedNet.pack_in_capture = True
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(max(int(warmup), 3)):          # torch asks for >= 3 eager steps before capturing an optimizer
                    body()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            before = ops.LAUNCHES[0]
            with torch.cuda.graph(self.graph):
                self.losses, self._flag = body()
            self.kernels_per_replay = ops.LAUNCHES[0] - before
        finally:
            ops.PackedNet.pack_in_capture = False
        with torch.no_grad():
            for m, keep, opt in zip(self.models, saved, self.optimizers):
                for p, k in zip(m.parameters(), keep):
                    p.copy_(k)
                    p.grad = None if p.grad is None else p.grad.zero_()
                for st in opt.state.values():          # exp_avg, exp_avg_sq, step: back to a fresh optimizer, same tensors
                    for v in st.values():
                        if torch.is_tensor(v):
                            v.zero_()
        self.replays = 0

    def __call__(self, batch=None):
        if batch is not None:
            for k in self.dev_in:
                if batch[k] is not self.dev_in[k]:
                    self.dev_in[k].copy_(batch[k], non_blocking=True)
        self.graph.replay()
        self.replays += 1
        ops.LAUNCHES[0] += self.kernels_per_replay
        return self.losses

    def check_unbounded(self):
        """Reads the out-of-sphere flag of the last replay (one host sync; ddp_train_nerf.py:62-63 raises there)."""
        if int(self._flag[0]) != 0:
            raise Exception(ops.UNBOUNDED_MSG)

    def invalidate_inference_caches(self):
        """After training through the graph, the models' cached weight tiles are current on the device, but the host-side
        cache keys are not: eager / GraphedRenderStep use afterwards must re-pack once."""
        for m in self.models:
            (m.nerf_net if hasattr(m, "nerf_net") else m).invalidate_packed()
