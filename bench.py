#!/usr/bin/env python
"""Benchmark of the NeRF++ ray-marching hot path on B200 (BASELINE.json metric: rays/sec, 4096 rays x
(64 coarse + 128 importance) samples, 8x256 MLP, depth_loss=mse -- configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step = one pass of the whole path over one batch of 4096 rays per GPU: intersect_sphere ->
stratified coarse depths -> NerfNet.forward (level 0, 64+64 samples) -> sample_pdf + merge (fg, bg)
-> NerfNet.forward (level 1, 192+192 samples) -> rgb MSE + depth loss per level; for N > 1 each rank
renders its contiguous band of the 4096*N-ray batch and one NCCL all-gather collects the rendered
(rgb, depth) tiles.  `value` times that with inputs resident in HBM; `e2e` times the same through
the public API from pinned host buffers with the H2D/D2H copies inside the timed region.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "outdoor-nerf-depth_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

N_RAYS = 4096
CASCADE = (64, 128)
# BASELINE.json:configs restated: name -> (global rays, GPUs the config is quoted on, depth losses)
CONFIGS = {
    "c2": dict(global_rays=4096, quoted_gpus=1, losses=("mse",), what="NeRF++ configs[1]: KITTI Seq00 sample_every=8, depth_sup_type=gt, depth_loss=mse"),
    "c4": dict(global_rays=8192, quoted_gpus=4, losses=("l1",), what="NeRF++ configs[3]: Argoverse 2c07fcda, depth_sup_type=stereo_crop, depth_loss=l1, 8192 rays tile-sharded over 4 GPUs"),
    "c5": dict(global_rays=16384, quoted_gpus=8, losses=("mse", "l1", "kl"), what="NeRF++ configs[4]: KITTI sweep, {mse,l1,kl}, 16384 rays tile-sharded over 8 GPUs"),
    # configs[2] is a different network: its own leg, run_c3()
    "c3": dict(global_rays=4096, quoted_gpus=1, losses=("kl",), what="MipNeRF-360 configs[2]: KITTI Seq06, depth_sup_type=mono_crop, depth_loss=kl, 4096 rays, "
               "PropMLP 4x256 (64 + 64 intervals) + NerfMLP 8x1024 (32 intervals), contract + IPE"),
}
LAMBDA_DEPTH = 0.1          # scripts/train.sh:5
DEPTH_SIGMA, DEPTH_SCALE = 0.01, 0.05
MACS_FG, MACS_BG = 593408, 604160      # per sample, SURVEY.md section 8(d)
METRIC = "rays/sec (4096 rays x 128 samples, 8x256 MLP)"
# roofline.traffic (dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel) cannot be measured
# inside a timed run (ncu replays kernels); it is read from the committed summary of the latest `ncu --set full` capture,
# and the line carries that file's name and hash so a stale figure is visible
TRAFFIC_FILE = os.path.join("profiles", "traffic.json")
WORKLOAD = ("NeRF++ configs[1]: 4096 rays/GPU, cascade 64 -> +128 (192 fine) fg and bg, depth_loss=mse lambda=0.1, "
            "forward of both levels + sampling + composite + losses")


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return p.get("bf16_tflops", 1590.0), p.get("bf16_tflops_sustained", 1400.0), p.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


NVML_CHILD = r"""
import sys, time
import pynvml
pynvml.nvmlInit()
key = sys.argv[1]
try:
    h = pynvml.nvmlDeviceGetHandleByUUID(key) if key.startswith("GPU-") else pynvml.nvmlDeviceGetHandleByIndex(int(key))
except Exception:
    h = pynvml.nvmlDeviceGetHandleByIndex(int(sys.argv[2]))
reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
out = sys.stdout
sys.stdin.readline()            # the parent says when the timed region starts: no polling before that
while True:
    out.write("%.6f,%d,%d,%d" % (time.time(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), mx, int(reasons(h))) + chr(10))
    out.flush()
    time.sleep(0.004)
"""


class ClockSampler:
    """SM clock and throttle reasons of this rank's GPU, sampled WHILE the timed region runs (B200_PROFILING.md recipe).

    Default: an `nvidia-smi -lms 100` child, started when the process starts (eight ranks starting one each need more
    than the 40 ms region just to print their first row) and read until 150 ms after the region; rows are stamped on
    arrival and only those from the region on count.  NERFPP_BENCH_CLOCKS=nvml selects the denser sampler below.

    The region is 20 steps of ~1.9 ms: a `nvidia-smi -lms 100` child delivers zero to two rows in 40 ms (and none at all
    when eight ranks start one each).  So the numbers nvidia-smi prints are read from NVML (nvidia_ml_py, the library
    nvidia-smi itself is built on) every ~4 ms by a CHILD process, like nvidia-smi: started early (NVML init takes a
    moment), told on its stdin when the region starts, killed when it ends.  Polling is confined to the region because
    seconds of 500 Hz NVML queries before it left the GPU measurably slower afterwards (the e2e leg lost 5-8 %), from a
    thread of this process as well as from a child.  Rows carry the child's wall-clock time; only rows inside the
    region count.  Without pynvml: the nvidia-smi child."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    BITS = [0x8, 0x40, 0x20, 0x4]

    def __init__(self, index):
        self.index, self.rows, self.proc, self.source, self.t0, self.t1 = index, [], None, None, None, None
        try:
            if os.environ.get("NERFPP_BENCH_CLOCKS") != "nvml":
                raise RuntimeError("default: the nvidia-smi child")
            import pynvml  # noqa: F401  (only to know the child can import it)
            try:
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                key = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
            except Exception:
                key = str(index)
            self.proc = subprocess.Popen([sys.executable, "-c", NVML_CHILD, key, str(index)], stdin=subprocess.PIPE,
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "NVML (nvidia_ml_py) polled by a child process every ~4 ms, inside the timed region only"
            self.nvml = True
        except Exception:
            self.nvml = False
            try:
                self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                              "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
                self.source = "nvidia-smi -lms 100"
            except Exception:
                self.proc = None
        if self.proc is not None:
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if self.nvml:
                try:
                    r = int(f[3])
                    self.rows.append((float(f[0]), [f[1], f[2]] + ["Active" if r & b else "Not Active" for b in self.BITS]))
                except Exception:
                    pass
            else:
                self.rows.append((time.time(), f))

    def __enter__(self):          # the timed region starts
        if self.nvml and self.proc is not None:
            try:
                self.proc.stdin.write("go\n")
                self.proc.stdin.flush()
            except Exception:
                pass
        self.t0 = time.time()
        return self

    def __exit__(self, *a):       # ... and ends
        self.t1 = time.time()
        if self.proc is not None:
            if not self.nvml:
                time.sleep(0.15)
                self.t1 = time.time()
            self.proc.terminate()
            self.t.join(timeout=2)

    def close(self):
        if self.proc is not None and self.proc.poll() is None:
            self.proc.terminate()

    def summary(self):
        rows = [f for (t, f) in self.rows if self.t0 is not None and self.t0 <= t <= self.t1]
        sm = sorted(int(r[0]) for r in rows if r and r[0].isdigit())
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "source": self.source}
        mx = max(int(r[1]) for r in rows if len(r) > 1 and r[1].isdigit())
        reasons = [nm for i, nm in enumerate(self.NAMES) if any(len(r) > 2 + i and r[2 + i].lower() == "active" for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": mx, "reasons": reasons, "samples": len(sm), "source": self.source}


def ncu_traffic(kernel):
    """{"bytes_per_launch": ..., "source": file, "sha16": ...} from profiles/traffic.json, or None."""
    import hashlib
    path = os.path.join(ROOT, TRAFFIC_FILE)
    try:
        raw = open(path, "rb").read()
        d = json.loads(raw)[kernel]
        return {"bytes_per_launch": d["dram_bytes_per_launch"], "source": TRAFFIC_FILE, "capture": d.get("capture"),
                "sha16": hashlib.sha256(raw).hexdigest()[:16]}
    except Exception:
        return None


def make_rays(n, seed):
    """SURVEY.md section 8(d) synthetic batch (host tensors): origins inside the ball of radius 0.5, un-normalised
    directions, rgb ~ U[0,1], depth prior = far * U with every 5th ray invalid (0).  Same draws as the oracle's
    ``synthetic_rays`` (the cpu_baseline leg uses that one), restated here so this arm never imports oracle/."""
    g = torch.Generator().manual_seed(seed)
    dirs = torch.randn(n, 3, generator=g)
    dirs = dirs / dirs.norm(dim=-1, keepdim=True)
    ray_o = dirs * (0.5 * torch.rand(n, 1, generator=g) ** (1.0 / 3.0))
    d = torch.randn(n, 3, generator=g)
    ray_d = d / d.norm(dim=-1, keepdim=True) * (1.0 + 0.2 * torch.rand(n, 1, generator=g))
    rgb = torch.rand(n, 3, generator=g)
    d1 = -(ray_d * ray_o).sum(-1) / (ray_d * ray_d).sum(-1)              # ddp_train_nerf.py:57-64
    pm = ray_o + d1[:, None] * ray_d
    far = d1 + torch.sqrt(1.0 - (pm * pm).sum(-1)) / ray_d.norm(dim=-1)
    depth_sup = far * torch.rand(n, generator=g)
    depth_sup[::5] = 0.0
    return dict(ray_o=ray_o.contiguous(), ray_d=ray_d.contiguous(), min_depth=torch.full((n,), 1e-4), rgb=rgb,
                depth_sup=depth_sup, depth_scale=DEPTH_SCALE)


def make_models(device):
    from types import SimpleNamespace
    import ddp_model
    args = SimpleNamespace(max_freq_log2=10, max_freq_log2_viewdirs=4, netdepth=8, netwidth=256, use_viewdirs=True)
    torch.manual_seed(777)   # ddp_train_nerf.py:308
    nets = [ddp_model.NerfNetWithAutoExpo(args) for _ in CASCADE]
    with torch.no_grad():
        for n in nets:        # non-degenerate density
            n.nerf_net.fg_net.sigma_layers[0].bias += 5.0
            n.nerf_net.bg_net.sigma_layers[0].bias += 5.0
    return [n.to(device) for n in nets]


def plan(args, world):
    """rays per rank, global rays, loss types and the scaling label for the asked config / scaling mode."""
    cfg = CONFIGS[args.config]
    if args.config == "c2" and args.scaling == "weak":
        n = N_RAYS                                   # the headline: 4096 rays per GPU, whatever N
    else:
        g = cfg["global_rays"]                       # strong scaling / tile-sharded configs: the global batch is fixed
        if g % world:
            raise SystemExit("config %s: %d rays do not divide over %d ranks (ddp_train_nerf.py:137-139)" % (args.config, g, world))
        n = g // world
    if args.rays_per_gpu:
        n = args.rays_per_gpu
    scaling = "weak" if (args.config == "c2" and args.scaling == "weak") else "strong"
    return n, n * world, cfg["losses"], scaling, cfg["what"]


def run_ours(args):
    from nerfpp_b200 import GraphedRenderStep, PipelinedRenderStep, _lib, ops, render_rays
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    n_rays, global_rays, loss_types, scaling, what = plan(args, world)
    clk = ClockSampler(local)          # its child process needs a moment to come up: started before the warm-up
    L = _lib.lib()
    if args.overlap is not None:       # A/B diagnostic: fg / bg nets on two streams -- 0 never, 1 default, 2 always
        import ctypes
        L.nerfpp_debug_set_overlap.argtypes = [ctypes.c_int]
        L.nerfpp_debug_set_overlap(args.overlap)
    models = make_models(dev)
    host = make_rays(n_rays, seed=rank)                 # this rank's contiguous band of the global batch
    host = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in host.items()}
    batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in host.items()}
    gathered = torch.empty(world * n_rays * 4, device=dev) if world > 1 else None   # per rank: rgb [n,3] | depth [n]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    pending_flag = [None]
    kws = [dict(cascade_samples=CASCADE, train=True, depth_loss_type=lt, lambda_depth=LAMBDA_DEPTH, depth_sigma=DEPTH_SIGMA) for lt in loss_types]
    graphed = not args.no_graph
    g_devs = g_hosts = None
    if graphed:
        # the step captured once into a CUDA graph (nerfpp_b200.graph): one launch per step instead of ~14 through Python;
        # one graph per depth-loss type of the config (a sweep config alternates them step by step)
        g_devs = [GraphedRenderStep(models, n_rays, depth_scale=DEPTH_SCALE, host_io=False, device=dev, **kw) for kw in kws]
        for g in g_devs:
            for k, v in g.dev_in.items():
                v.copy_(batch[k])
        # host -> host: two such graphs on their own streams, used alternately, so that step i+1's H2D and step i-1's host
        # read-back overlap step i's kernels (nerfpp_b200.graph.PipelinedRenderStep); for N > 1 the all-gather of the
        # rendered tile and the D2H of the gathered image ride on the slot's stream right behind the replay
        gath_dev = [torch.empty(world * n_rays * 4, device=dev) for _ in range(args.e2e_depth)] if world > 1 else None
        gath_host = [torch.empty(world * n_rays * 4).pin_memory() for _ in range(args.e2e_depth)] if world > 1 else None

        def gather_tiles(i, st):
            dist.all_gather_into_tensor(gath_dev[i], st._packed[:4 * n_rays])
            gath_host[i].copy_(gath_dev[i], non_blocking=True)
        g_hosts = [PipelinedRenderStep(models, n_rays, depth=args.e2e_depth, depth_scale=DEPTH_SCALE, device=dev,
                                       after_launch=gather_tiles if world > 1 else None, **kw) for kw in kws]

    # N > 1: the step's rendered tile (rgb | depth) is copied aside and all-gathered on a SIDE stream, so the collective of
    # step i runs under the kernels of step i+1 (it needs no SM time to speak of; it cannot share an SM with a field
    # CTA, which owns all of the shared memory, so it slips into the tail waves).  Two tile / image buffers alternate; what
    # is left of the last gather when the loop ends is timed separately and added (below).
    main_stream = torch.cuda.current_stream(dev)
    side = torch.cuda.Stream(device=dev) if world > 1 else None
    tile_bufs = [torch.empty(4 * n_rays, device=dev) for _ in range(2)] if world > 1 else None
    image_bufs = [torch.empty(world * n_rays * 4, device=dev) for _ in range(2)] if world > 1 else None
    gather_done = [None, None]
    step_no = [0]

    def step(b):
        """One pass over this rank's rays with the inputs resident in HBM."""
        which = step_no[0] % len(kws)
        with torch.no_grad():
            if graphed:
                g_dev = g_devs[which]
                out = g_dev()                            # inputs: g_dev.dev_in (filled once above)
                i = step_no[0] & 1
                step_no[0] += 1
                if world > 1:
                    if gather_done[i] is not None:
                        main_stream.wait_event(gather_done[i])        # the gather two steps back has read this tile buffer
                    tile_bufs[i].copy_(g_dev._packed[:4 * n_rays])
                    side.wait_stream(main_stream)
                    with torch.cuda.stream(side):
                        dist.all_gather_into_tensor(image_bufs[i], tile_bufs[i])
                        gather_done[i] = torch.cuda.Event()
                        gather_done[i].record()
                return out
            step_no[0] += 1
            if pending_flag[0] is not None:          # the previous step's out-of-sphere flag (its work is long done)
                pending_flag[0].raise_if_set()
            res = render_rays(models, b, defer_unbounded_check=True, **kws[which])
            pending_flag[0] = res["unbounded"]
            ret = res["levels"][-1][0]
            if world > 1:
                tile = torch.cat((ret["rgb"].reshape(-1), ret["depth"]))
                dist.all_gather_into_tensor(gathered, tile)
        return res, ret

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step(batch)
    barrier()
    # ---- timed: K steps, device time per step by CUDA events, L2 flushed between steps ----
    ops.LAUNCHES[0] = 0
    step_no[0] = 0
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with clk:
        barrier()
        t_wall = time.perf_counter()
        for a, b in evs:
            flush.zero_()
            a.record()
            step(batch)
            b.record()
        tail = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        tail[0].record()
        if side is not None:
            main_stream.wait_stream(side)            # what is left of the last step's all-gather
        tail[1].record()
        barrier()
        t_wall = time.perf_counter() - t_wall
    launches = ops.LAUNCHES[0]
    if graphed:
        for g in g_devs:
            g.check_unbounded()
    ms = sum(a.elapsed_time(b) for a, b in evs) + tail[0].elapsed_time(tail[1])
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = global_rays / (ms_per_step * 1e-3)

    # ---- e2e: pinned host buffers -> H2D -> path -> D2H of rgb/depth/loss, wall clock ----
    def e2e_step():
        """Eager variant (--no-graph): the same pass from HOST buffers to HOST results, one step at a time."""
        b = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in host.items()}
        res, ret = step(b)
        out = torch.cat((ret["rgb"], ret["depth"][:, None]), -1) if world == 1 else gathered
        h = out.cpu()
        loss = [l.cpu() for l in res["losses"]]
        return h, loss

    def e2e_run(k):
        """k steps from HOST buffers to HOST results; returns the wall-clock seconds.  Every step copies its batch
        host -> device and its rgb / depth / losses (and, N > 1, the gathered image) device -> host; the L2 flush of
        the measurement protocol is enqueued before every step and is INSIDE the timed region.  A sweep config runs its
        loss types one after the other, k steps in total."""
        t0 = time.perf_counter()
        if graphed:
            shares = [k // len(g_hosts) + (1 if i < k % len(g_hosts) else 0) for i in range(len(g_hosts))]
            for g_host, kk in zip(g_hosts, shares):
                ahead = min(args.e2e_depth - 1, kk)
                for _ in range(ahead):               # fill the pipeline
                    flush.zero_()
                    g_host.submit(host)
                for _ in range(kk - ahead):
                    flush.zero_()
                    g_host.submit(host)              # step i+depth-1 is staged and enqueued ...
                    g_host.result()                  # ... before step i's results are read on the host
                for _ in range(ahead):
                    g_host.result()
        else:
            for _ in range(k):
                flush.zero_()
                e2e_step()
        return time.perf_counter() - t0

    e2e_run(3)
    barrier()
    t_sum = e2e_run(args.steps)
    barrier()
    t_e2e = torch.tensor([t_sum], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = global_rays * args.steps / float(t_e2e.item())
    h2d = sum(v.numel() * v.element_size() for v in host.values() if torch.is_tensor(v))
    d2h = (world * n_rays * 4 * 4 if world > 1 else n_rays * 4 * 4) + 2 * 4 * 4
    if graphed:   # the graph's one D2H (rgb, depth, 2 x 4 losses, flag), plus the gathered tiles when N > 1
        d2h = g_hosts[0].slots[0]._n_out * 4 + (world * n_rays * 4 * 4 if world > 1 else 0)

    # ---- the trainer's step (ddp_train_nerf.py:432-498: per level forward, loss, backward, Adam), every N ----
    # (auxiliary legs never take the headline line down with them: a failure is reported in place of the numbers)
    train = None
    if not args.no_train:
        try:
            train = train_step_rate(models, batch, dev, n_rays, loss_types[0], dist if world > 1 else None, world)
        except Exception as e:   # noqa: BLE001
            if world > 1:
                raise            # a rank that drops out of a collective would hang the others: fail loudly instead
            train = {"error": repr(e)[:500]}

    # ---- roofline of the dominant kernel (field_tc_kernel), timed alone with CUDA events ----
    roof = field_roofline(models, batch, dev)
    cpu = parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu, parity = cpu_baseline(dev=dev)
        except Exception as e:   # noqa: BLE001
            cpu = {"error": repr(e)[:500]}
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f32 (fp16 tensor-core operands, fp32 accumulate)", "data": "synthetic",
            "config": {"workload": WORKLOAD if args.config == "c2" else what + "; cascade 64 -> +128 fg and bg, forward of both levels + sampling + composite + losses",
                       "name": args.config, "rays_per_gpu": n_rays, "global_rays": global_rays, "depth_losses": list(loss_types),
                       "cascade_samples": list(CASCADE), "unit_of_work": "U2 forward (SURVEY 8(d)): 2.512 TFLOP algorithmic per 4096 rays",
                       "l2": "flushed between timed steps (256 MiB write)", "field": "tcgen05",
                       "parallelism": ("rays sharded in contiguous bands, %d ranks, 1 NCCL all-gather/step on a side stream (overlaps the next step)" % world)
                       if world > 1 else "single GPU",
                       "wall_s_timed_region": t_wall,
                       "launch": ("one CUDA graph per step (%d library kernels + torch rand/cat nodes)" % g_devs[0].kernels_per_replay) if graphed
                       else "eager: one Python call per kernel"},
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps_in_flight": args.e2e_depth if graphed else 1,
                    "how": ("PipelinedRenderStep: host batch -> pinned staging -> graph (H2D, kernels, D2H) -> host results, one "
                            "graph and stream per step in flight so step i+1's copies overlap step i's kernels; wall clock over all steps, L2 "
                            "flush included") if graphed else "eager, one step at a time, wall clock, L2 flush included"},
            "gpu_launches": launches,
            "roofline": roof,
        }
        if train is not None:
            line["train"] = train
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if parity is not None:
            line["parity"] = parity
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def train_step_rate(models, batch, dev, n_rays, loss_type, dist, world, steps=20):
    """rays/s of the reference trainer's step on this rank's rays: for each cascade level forward (training mode), rgb MSE +
    0.1 * depth loss, loss.backward() through nerfpp_backward, Adam step -- captured into ONE CUDA graph
    (nerfpp_b200.GraphedTrainStep).  N > 1: the flat gradient buffer of each level is averaged with one NCCL all-reduce
    inside the graph (what DDP-over-gloo does in the reference, ddp_train_nerf.py:298,323).  Device time by CUDA events,
    max over ranks; reported beside the headline metric, with its own roofline."""
    import copy
    from nerfpp_b200 import GraphedTrainStep, ops
    burst, sustained, hbm, how = peaks()
    nets = [copy.deepcopy(m) for m in models]
    group = dist.group.WORLD if dist is not None else None
    ts = GraphedTrainStep(nets, n_rays, CASCADE, depth_loss_type=loss_type, lambda_depth=LAMBDA_DEPTH, depth_sigma=DEPTH_SIGMA,
                          depth_scale=DEPTH_SCALE, device=dev, process_group=group, batch=batch)
    for _ in range(3):
        ts()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        ts()
    b.record()
    torch.cuda.synchronize()
    ts.check_unbounded()
    t = torch.tensor([a.elapsed_time(b) / steps], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    loss = [float(x) for x in ts.losses.tolist()]
    pairs_per_ray = sum(sum(CASCADE[:i + 1]) for i in range(len(CASCADE)))        # level m evaluates all depths so far: 64 + 192
    flops = 3.0 * 2.0 * (MACS_FG + MACS_BG) * n_rays * pairs_per_ray               # fwd + dgrad + wgrad, SURVEY 8(d) unit U2
    achieved = flops / (ms * 1e-3) / 1e12
    tr = ncu_traffic("train_step")
    out = {"value": n_rays * world / (ms * 1e-3), "unit": "rays/s", "ms_per_step": ms, "steps": steps, "rays_per_gpu": n_rays, "n_gpus": world,
           "depth_loss": loss_type, "final_losses": loss, "kernels_per_step": ts.kernels_per_replay,
           "what": "forward + loss + backward (tcgen05 dgrad / wgrad kernels) + Adam, both cascade levels, ONE CUDA graph per step"
                   + ("; gradients averaged by one NCCL all-reduce of the flat buffer per level" if world > 1 else ""),
           "roofline": {"bound": "hbm + tensor (see DESIGN.md section 5: the step moves ~44 GB of saved operands)", "achieved": achieved,
                        "peak": sustained, "unit": "TFLOP/s", "frac": achieved / sustained,
                        "peak_source": "%s bf16 sustained (MEASURED_PEAKS.json): timed inside a long step" % how,
                        "algorithmic_flops_per_step": flops, "traffic": tr["bytes_per_launch"] if tr else None, "traffic_source": tr}}
    del ts, nets
    torch.cuda.empty_cache()
    return out


def field_roofline(models, batch, dev, reps=10):
    """Average duration of the field kernel's four launches in a step (coarse/fine x fg/bg), each
    timed alone with CUDA events on the launching stream; achieved = algorithmic FLOPs / time."""
    from nerfpp_b200 import cascade_forward, ops
    burst, sustained, hbm, how = peaks()
    with torch.no_grad():
        out, far = cascade_forward(models, batch["ray_o"], batch["ray_d"], batch["min_depth"], CASCADE, train=True)
    impl = 0      # FIELD_TC
    flops = tot_ms = 0.0
    n_launch = 0
    with torch.no_grad():
        for m, (ret, fg_z, bg_z) in enumerate(out):
            net = models[m].nerf_net
            for is_bg, z, tensors, macs in ((0, fg_z, net.fg_net.tensors(), MACS_FG), (1, bg_z, net.bg_net.tensors(), MACS_BG)):
                packed = net._packed[is_bg].get(tensors, impl)
                for _ in range(3):
                    ops.field_forward(packed, is_bg, batch["ray_o"], batch["ray_d"], z, impl)
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(reps):
                    ops.field_forward(packed, is_bg, batch["ray_o"], batch["ray_d"], z, impl)
                b.record()
                torch.cuda.synchronize()
                tot_ms += a.elapsed_time(b) / reps
                flops += 2.0 * macs * z.numel()
                n_launch += 1
    achieved = flops / (tot_ms * 1e-3) / 1e12
    return {"bound": "tensor", "kernel": "field_tc_kernel" if impl == 0 else "field_simt_kernel", "achieved": achieved, "peak": burst,
            "unit": "TFLOP/s", "frac": achieved / burst, "traffic": (ncu_traffic("field_tc_kernel") or {}).get("bytes_per_launch"),
            "traffic_source": ncu_traffic("field_tc_kernel"), "peak_source": "%s bf16 burst (MEASURED_PEAKS.json)" % how,
            "launches_per_step": n_launch, "avg_launch_ms": tot_ms / n_launch, "algorithmic_flops_per_step": flops}


def reference_step(levels, rays, O):
    """The reference's cascade loop (ddp_train_nerf.py:432-493) restated by the oracle port, no_grad forward + losses."""
    with torch.no_grad():
        out, far = O.cascade_forward(levels, rays["ray_o"], rays["ray_d"], rays["min_depth"], CASCADE,
                                     O.synthetic_rand(rays["ray_o"].shape[0], CASCADE, seed=1))
        losses = [O.level_loss(ret, rays["rgb"], rays["depth_sup"], fg_z, far, True, "mse", LAMBDA_DEPTH, DEPTH_SIGMA * DEPTH_SCALE)
                  for ret, fg_z, _ in out]
    return out, losses


def parity_vs_oracle(levels, rays, ref_out, ref_losses, O, dev):
    """The checker half of the cpu_baseline leg: the CUDA path on the very rays, weights and random draws the oracle
    has just been timed on.  Reports max relative error (normalised by the largest reference magnitude, as the tests
    do) of the finest level's rgb / depth, of each level's total loss, and the PSNR of our rgb against the oracle's."""
    from types import SimpleNamespace
    import math
    import ddp_model
    from nerfpp_b200 import render_rays
    args = SimpleNamespace(max_freq_log2=10, max_freq_log2_viewdirs=4, netdepth=8, netwidth=256, use_viewdirs=True)
    nets = []
    for p in levels:
        net = ddp_model.NerfNetWithAutoExpo(args)
        net.load_state_dict(p)
        nets.append(net.to(dev))
    n = rays["ray_o"].shape[0]
    rand = {k: v.to(dev) for k, v in O.synthetic_rand(n, CASCADE, seed=1).items()}
    b = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in rays.items()}
    with torch.no_grad():
        res = render_rays(nets, b, CASCADE, train=True, depth_loss_type="mse", lambda_depth=LAMBDA_DEPTH, depth_sigma=DEPTH_SIGMA, rand=rand)
    got, want = res["levels"][-1][0], ref_out[-1][0]
    rel = lambda a, w: float((a.cpu().double() - w.double()).abs().max() / w.double().abs().max())
    mse = float(((got["rgb"].cpu().double() - want["rgb"].double()) ** 2).mean())
    loss_rel = max(abs(float(res["losses"][m][2]) - float(ref_losses[m][0])) / abs(float(ref_losses[m][0])) for m in range(len(CASCADE)))
    coarse = {}
    for name, idx in (("fg", 1), ("bg", 2)):      # stratified + perturbed depths of level 0: integer-exact work, expected 0 / 0
        a, w = res["levels"][0][idx].cpu(), ref_out[0][idx]
        coarse[name] = {"mismatched": int((a != w).sum()), "of": a.numel(), "max_abs": float((a - w).abs().max())}
    def per_ray(k, floor):
        """|a - b| / max(|b|, floor) per ray (a ray's error = its worst channel): max / p99 / p50 over the batch."""
        a, w = got[k].cpu().double(), want[k].double()
        e = (a - w).abs() / w.abs().clamp_min(floor)
        e = e.reshape(e.shape[0], -1).max(dim=1).values
        q = torch.quantile(e, torch.tensor([0.5, 0.99], dtype=torch.float64))
        return {"max": float(e.max()), "p99": float(q[1]), "p50": float(q[0]), "floor": floor}
    return {"rays": n, "rgb_max_rel": rel(got["rgb"], want["rgb"]), "depth_max_rel": rel(got["depth"], want["depth"]),
            "per_ray_rel": {"rgb": per_ray("rgb", 1e-2), "depth": per_ray("depth", 1e-3), "fg_depth": per_ray("fg_depth", 1e-3),
                            "note": "default-init weights (+5 density bias); on weights TRAINED by the fp32 oracle the single-pass fp16 "
                                    "operands give rgb p50 9e-5 / p99 3e-4 / max 5e-4 and depth p50 2e-4 / max 6e-4 "
                                    "(tests/test_convergence_gpu.py, profiles/r2_convergence.json): above 1e-4 at the tail"},
            "loss_max_rel": loss_rel, "psnr_vs_oracle_db": (-10.0 * math.log10(mse) if mse > 0 else float("inf")),
            "coarse_depths": coarse,
            "coarse_depths_note": "fg depths inherit intersect_sphere's far bound: the LIVE oracle's torch-CPU sum/norm round the "
                                  "scalar tail elements of each thread's chunk without FMA, so a few rays per host thread differ by "
                                  "1 ulp depending on the host's thread count; against the goldens recorded from the unmodified "
                                  "reference the same kernels are bit-exact (tests/test_parity_gpu.py)",
            "tolerance": 1e-4,
            "against": "oracle/nerfpp_oracle.py on the same rays, weights and random draws (train path, both levels)"}


def reference_cpu_arm():
    """The CPU arm's step function and what it is.  When the reference's own files are present (the git-ignored copy
    ``baseline/_ref/nerfplusplus`` made by oracle/install_reference.py, or /root/reference in the build container) the step
    calls the UNMODIFIED reference functions wired as the trainer's cascade loop (oracle/ref_harness.py): kind
    "reference".  Otherwise the oracle port (kind "port")."""
    try:
        import ref_harness as RH
        if RH.available():
            nets = RH.build_nets(len(CASCADE), sigma_bias=5.0)
            step = lambda rays: RH.reference_step(nets, rays, CASCADE, "mse", LAMBDA_DEPTH, DEPTH_SIGMA)
            return step, "reference", ("the reference's own functions, unmodified (intersect_sphere / perturb_samples / sample_pdf, "
                                       "NerfNetWithAutoExpo.forward, img2mse, depth_mse) wired as ddp_train_nerf.py:432-493, torch-CPU fp32")
    except Exception as e:   # noqa: BLE001  (fall back to the port, say why)
        sys.stderr.write("reference files unusable (%r): timing the oracle port instead\n" % (e,))
    import nerfpp_oracle as O
    levels = [O.densify(p, 5.0) for p in O.make_params_levels(2)]
    return (lambda rays: reference_step(levels, rays, O)), "port", "oracle/nerfpp_oracle.py: torch-CPU fp32 restatement of the reference"


def cpu_baseline(sample=N_RAYS, reps=4, dev=None):
    """cpu_baseline leg of the product arm: the reference's CPU implementation on all host cores on ``reps`` batches of
    the SAME 4096-ray workload, and -- the same leg, as the checker -- the parity of the CUDA path against the oracle."""
    import nerfpp_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, kind, what = reference_cpu_arm()
    rays = make_rays(sample, 6)
    step(make_rays(64, 5))     # warm-up
    t0 = time.perf_counter()
    for _ in range(reps):
        step(rays)
    dt = time.perf_counter() - t0
    base = {"value": reps * sample / dt, "unit": "rays/s", "cores": cores, "kind": kind,
            "sample": "%d x %d rays of the same workload (%s, all host threads), %.1f s" % (reps, sample, what, dt)}
    parity = None
    if dev is not None:
        levels = [O.densify(p, 5.0) for p in O.make_params_levels(2)]
        ref_out, ref_losses = reference_step(levels, rays, O)
        parity = parity_vs_oracle(levels, rays, ref_out, ref_losses, O, dev)
    return base, parity


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores, on this arm's config
    (4096 rays per step, cascade 64 -> +128, mse): baseline/_ref when the reference files travelled, else the oracle port."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, kind, what = reference_cpu_arm()
    rays = make_rays(N_RAYS, 6)
    warm, steps = max(1, args.warmup), max(1, args.steps)
    for _ in range(warm):
        step(rays)
    t0 = time.perf_counter()
    for _ in range(steps):
        step(rays)
    dt = (time.perf_counter() - t0) / steps
    v = N_RAYS / dt
    world = int(os.environ.get("WORLD_SIZE", 1))
    emit(({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "rays/s", "n_gpus": world, "steps": steps, "warmup": warm,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rays_per_gpu": N_RAYS, "global_rays": N_RAYS, "cascade_samples": list(CASCADE),
                   "unit_of_work": "U2 forward (SURVEY 8(d)): 2.512 TFLOP algorithmic per 4096 rays", "where": "host cores of the GPU box"},
        "cpu_baseline": {"value": v, "unit": "rays/s", "cores": cores, "kind": kind,
                         "sample": "%d rays/step of the same workload (%s), %d threads" % (N_RAYS, what, cores)},
        "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ---------------------------------------------------------------------------------------------------------------------
# config 3: mipnerf360 (models.py:75-330 under configs/360.gin) -- forward of the three levels + the trainer's loss terms
# ---------------------------------------------------------------------------------------------------------------------
C3_PROP_MACS, C3_NERF_MACS = 325888, 8672000          # per sample (Dense stacks; tests/test_mip360_model_oracle.py)
C3_SAMPLES = (64, 64, 32)
C3_FLOPS_PER_RAY = 2.0 * (128 * C3_PROP_MACS + 32 * C3_NERF_MACS)      # 638.4 MFLOP


def c3_batch(n, seed):
    """A seeded 360-style batch (origins around the unit ball, unnormalised directions, pixel-footprint radii, near 0.2,
    far 1e6 as configs/360.gin; rgb ~ U[0,1]^3, depth prior with every 5th ray invalid)."""
    g = torch.Generator().manual_seed(seed)
    o = torch.randn(n, 3, generator=g)
    o = o / o.norm(dim=-1, keepdim=True) * (0.2 + torch.rand(n, 1, generator=g))
    d = torch.randn(n, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    dirs = d * (0.9 + 0.4 * torch.rand(n, 1, generator=g))
    sup = 0.5 + 20 * torch.rand(n, 1, generator=g)
    sup[::5] = 0
    return dict(origins=o, directions=dirs, viewdirs=dirs / dirs.norm(dim=-1, keepdim=True), radii=5e-4 + 1.5e-3 * torch.rand(n, 1, generator=g),
                near=torch.full((n, 1), 0.2), far=torch.full((n, 1), 1e6), rgb=torch.rand(n, 3, generator=g), disps_sup=sup)


def c3_gemm_roofline(dev, n_rays, prec, model, reps=10):
    """The dominant kernel of the step, gemm_tc_kernel: the NerfMLP's ten Dense-layer launches (8 trunk layers, bottleneck,
    view layer; M = rays x 32), each shape timed alone with CUDA events through the C ABI's one-layer entry point;
    achieved = algorithmic FLOPs (true in-features: 504, 1024, 1528, 283 -- padding excluded) / time.  Plus, as context, each
    level's whole field call (encode + Dense stack + heads) timed the same way."""
    import ctypes
    from nerfpp_b200 import _lib
    from nerfpp_b200.mip360_model import Rays, dense_shapes
    burst, sustained, hbm, how = peaks()
    L = _lib.lib()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    M = n_rays * C3_SAMPLES[2]
    true_shapes = dense_shapes(8, 1024, True)
    gemm_layers = [true_shapes[i] for i in range(8)] + [true_shapes[9], true_shapes[10]]       # (in, out); the 1-wide heads are not GEMMs
    tot_ms = flops = 0.0
    cache = {}
    for fin, fout in gemm_layers:
        K = (fin + 63) // 64 * 64 if fin != 1528 else 1536
        if (fout, K) not in cache:
            a = torch.randn(M, K, device=dev).half()
            w = (torch.randn(fout, K, device=dev) / K ** 0.5).half()
            b = torch.zeros(fout, device=dev)
            out = torch.empty(M, fout, device=dev, dtype=torch.float16)
            call = lambda: _lib.check(L.mip360_dense_f16(a.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), M, fout, K, 1, st), "dense")
            for _ in range(3):
                call()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                call()
            e1.record()
            torch.cuda.synchronize()
            cache[(fout, K)] = e0.elapsed_time(e1) / reps
            del a, w, b, out
        tot_ms += cache[(fout, K)]
        flops += 2.0 * M * fin * fout
    torch.cuda.empty_cache()
    passes = 1
    # per level: the whole field call
    host = c3_batch(n_rays, 3)
    R = Rays(*(host[k].to(dev) for k in ("origins", "directions", "viewdirs", "radii", "near", "far")))
    levels = {}
    for name, mlp, S, macs in (("prop", model.prop_mlp, C3_SAMPLES[0], C3_PROP_MACS), ("nerf", model.nerf_mlp, C3_SAMPLES[2], C3_NERF_MACS)):
        sd = torch.linspace(0, 1, S + 1, device=dev).repeat(n_rays, 1).contiguous()
        for _ in range(3):
            mlp.level(sd, R)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            mlp.level(sd, R)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        levels[name] = {"ms": ms, "samples": n_rays * S, "tflops": 2.0 * macs * n_rays * S / (ms * 1e-3) / 1e12,
                        "what": "cast + contract + IPE encode, " + ("PropMLP chain kernel (4 layers + density head, one launch)" if name == "prop" else
                                                                    "10 Dense-layer GEMM launches (CTA pairs), density / rgb heads")}
    if prec:
        # the one-layer entry point runs the fp16 kernel; the split-precision launches (hi + lo operands, three MMA groups per
        # K-chunk) are timed through the NerfMLP level's field call instead, whose encode and heads (~6 %) are then inside
        tot_ms = levels["nerf"]["ms"]
    achieved = flops / (tot_ms * 1e-3) / 1e12
    t = ncu_traffic("gemm_tc_kernel")
    return {"bound": "tensor", "kernel": "gemm_tc_kernel", "achieved": achieved, "peak": burst, "unit": "TFLOP/s", "frac": achieved / burst,
            "traffic": (t or {}).get("bytes_per_launch"), "traffic_source": t, "peak_source": "%s bf16 burst (MEASURED_PEAKS.json)" % how,
            "launches_per_step": len(gemm_layers), "avg_launch_ms": tot_ms * passes / len(gemm_layers), "algorithmic_flops_per_step": flops,
            "peak_sustained": sustained, "frac_sustained": achieved / sustained,      # the leg runs after the timed steps, on a box already at its power cap
            "levels": levels,
            "note": ("the NerfMLP's Dense layers (87 %% of the step's FLOPs); K is padded to 64 (504 -> 512, 283 -> 320), padded FLOPs are not in the numerator"
                     + ("; split precision: timed through the NerfMLP level's whole field call (the one-layer entry point is fp16 only); its three MMA "
                        "groups per K-chunk do not add to the numerator" if prec else ""))}


def c3_oracle_step(prop, nerf, rays, u_levels, MM, mo):
    rend, hist = MM.model_forward(prop, nerf, rays, train_frac=0.5, u_levels=u_levels)
    losses = []
    for r, h in zip(rend, hist):
        losses.append(float(((r["rgb"] - rays["rgb"]) ** 2).mean()))
        losses.append(float(mo.depth_loss_kl(h["weights"], h["tdist"], rays["disps_sup"].reshape(-1), DEPTH_SIGMA * 1.0, rays["directions"])))
    return rend, hist, losses


def c3_cpu_baseline(dev, n=256, reps=2):
    """The CPU restatement of config 3's step (oracle/mip360_model_oracle.py: numpy fp32, BLAS threads = all host cores) on a
    bounded sample, and -- same leg, as the checker -- the parity of the CUDA model against it in both precision modes."""
    import numpy as np
    import mip360_model_oracle as MM
    import mip360_oracle as mo
    from nerfpp_b200.mip360_model import Model, Rays
    cores = os.cpu_count() or 1
    b = c3_batch(n, 6)
    rays = {k: v.numpy() for k, v in b.items()}
    prop, nerf = MM.init_mlp_params(4, 256, False, seed=21), MM.init_mlp_params(8, 1024, True, seed=22)
    g = np.random.default_rng(9)
    u_levels = []
    for ns in C3_SAMPLES:
        base, mj = mo.jitter_base_u(ns)
        u_levels.append((base[None, :] + g.random((n, 1)).astype(np.float32) * mj).astype(np.float32))
    c3_oracle_step(prop, nerf, {k: v[:16] for k, v in rays.items()}, [u[:16] for u in u_levels], MM, mo)      # warm-up
    t0 = time.perf_counter()
    for _ in range(reps):
        rend_ref, hist_ref, _ = c3_oracle_step(prop, nerf, rays, u_levels, MM, mo)
    dt = time.perf_counter() - t0
    base = {"value": reps * n / dt, "unit": "rays/s", "cores": cores, "kind": "port",
            "sample": "%d x %d rays of the same workload (oracle/mip360_model_oracle.py: numpy fp32 restatement of the JAX reference, which "
                      "cannot run here -- no jax/flax in the image; BLAS on all host threads), %.1f s" % (reps, n, dt)}
    parity = {"rays": n, "floors": {"rgb": 1e-2, "depth": 1e-3}, "norm": "per ray |a-b| / max(|b|, floor): max / p99 / p50"}
    R = Rays(*(b[k].to(dev) for k in ("origins", "directions", "viewdirs", "radii", "near", "far")))
    for prec in (False, True):
        model = Model(dev, prec=prec)
        model.nerf_mlp.load(nerf)
        model.prop_mlp.load(prop)
        with torch.no_grad():
            rend, hist = model(None, R, train_frac=0.5, u_levels=[torch.from_numpy(u).to(dev) for u in u_levels])
        out = {}
        for key, floor in (("rgb", 1e-2), ("depth", 1e-3)):
            a, r = rend[-1][key].cpu().numpy(), rend_ref[-1][key]
            err = np.abs(a - r)
            den = np.maximum(np.abs(r), floor)
            rel = (err / den).max(-1) if err.ndim > 1 else err / den
            out[key] = [float(rel.max()), float(np.percentile(rel, 99)), float(np.median(rel))]
        parity["split_precision" if prec else "fp16_operands"] = out
        del model
    torch.cuda.empty_cache()
    return base, parity


def run_c3(args):
    """BASELINE.json configs[2]: one pass = Model.__call__ over 4096 rays per GPU (PropMLP 64 + 64 intervals, NerfMLP 32) +
    the trainer's loss terms (per-level mse and kl depth prior, interlevel, distortion), forward only -- the 1024-wide network
    has no backward in this repository yet.  Rays are independent: N ranks run N shards, no collective."""
    from nerfpp_b200 import _lib, ops
    from nerfpp_b200.mip360_model import GraphedModelStep, Model
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    n_rays = args.rays_per_gpu or N_RAYS
    clk = ClockSampler(local)
    _lib.lib()
    prec = bool(args.c3_split)
    model = Model(dev, prec=prec).init(rank)
    host = c3_batch(n_rays, seed=rank)
    dev_step = GraphedModelStep(model, n_rays, train_frac=0.5, depth_sigma=DEPTH_SIGMA, host_io=False)
    for k, v in dev_step.dev_in.items():
        v.copy_(host[k].to(dev))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        dev_step()
    barrier()
    ops.LAUNCHES[0] = 0
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with clk:
        barrier()
        t_wall = time.perf_counter()
        for a, b in evs:
            flush.zero_()
            a.record()
            dev_step.launch()
            b.record()
        barrier()
        t_wall = time.perf_counter() - t_wall
    launches = args.steps * dev_step.kernels_per_replay
    t = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = n_rays * world / (ms_per_step * 1e-3)

    # ---- e2e: host batch -> pinned staging -> graph (H2D, kernels, D2H) -> host rgb / depth / losses; two steps in flight on
    # one stream (step i+1 is staged and enqueued before step i's results are read) ----
    slots = [GraphedModelStep(model, n_rays, train_frac=0.5, depth_sigma=DEPTH_SIGMA, host_io=True) for _ in range(2)]

    def e2e_run(k):
        t0 = time.perf_counter()
        flush.zero_()
        slots[0].launch(host)
        for i in range(1, k):
            flush.zero_()
            slots[i & 1].launch(host)
            slots[(i - 1) & 1].fetch()
        slots[(k - 1) & 1].fetch()
        return time.perf_counter() - t0

    e2e_run(3)
    barrier()
    t_e2e = torch.tensor([e2e_run(args.steps)], device=dev, dtype=torch.float64)
    barrier()
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = n_rays * world * args.steps / float(t_e2e.item())
    h2d = sum(v.numel() * 4 for v in host.values())
    d2h = slots[0]._n_out * 4
    last = slots[(args.steps - 1) & 1].fetch()
    losses = dict(zip(("mse_0", "mse_1", "mse_2", "depth_0", "depth_1", "depth_2", "interlevel", "distortion"), [float(x) for x in last["losses"]]))
    del slots
    # ---- the same step with split-precision operands (the mode that holds 1e-4 on the rendered colour):
    # measured beside the default so that both rates are on the record ----
    split = None
    if not prec:
        try:
            m2 = Model(dev, prec=True).init(rank)
            s2 = GraphedModelStep(m2, n_rays, train_frac=0.5, depth_sigma=DEPTH_SIGMA, host_io=False)
            for k, v in s2.dev_in.items():
                v.copy_(host[k].to(dev))
            for _ in range(3):
                s2()
            barrier()
            k2 = max(10, min(40, args.steps // 5))
            ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k2)]
            for a, b in ev2:
                flush.zero_()
                a.record()
                s2.launch()
                b.record()
            barrier()
            t2 = torch.tensor([sum(a.elapsed_time(b) for a, b in ev2)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t2, op=dist.ReduceOp.MAX)
            ms2 = float(t2.item()) / k2
            split = {"value": n_rays * world / (ms2 * 1e-3), "unit": "rays/s", "ms_per_step": ms2, "steps": k2,
                     "what": "the same graphed step with hi + lo fp16 operands (three MMA passes per Dense layer, libm sin / exp in the encoding): "
                             "per-ray colour error <= 7e-5 instead of 3e-3 (parity object); device-resident, L2 flushed between steps"}
            del s2, m2
            torch.cuda.empty_cache()
        except Exception as e:   # noqa: BLE001
            if world > 1:
                raise
            split = {"error": repr(e)[:300]}
    roof = c3_gemm_roofline(dev, n_rays, prec, model)
    cpu = parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu, parity = c3_cpu_baseline(dev)
        except Exception as e:   # noqa: BLE001
            cpu = {"error": repr(e)[:500]}
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (hi + lo fp16 tensor-core operands, three MMA passes, fp32 accumulate)" if prec else "f32 (fp16 tensor-core operands, fp32 accumulate)",
            "data": "synthetic",
            "config": {"workload": CONFIGS["c3"]["what"] + "; forward of the three levels (resampling, ray warp, cast, contraction, IPE, Dense stacks, "
                                                           "compositing) + per-level mse / kl depth prior + interlevel + distortion terms; no backward",
                       "name": "c3", "rays_per_gpu": n_rays, "global_rays": n_rays * world, "depth_losses": ["kl"], "samples_per_level": list(C3_SAMPLES),
                       "unit_of_work": "%.1f MFLOP per ray = %.3f TFLOP algorithmic per %d rays (Dense stacks only)" % (C3_FLOPS_PER_RAY / 1e6, C3_FLOPS_PER_RAY * n_rays / 1e12, n_rays),
                       "l2": "flushed between timed steps (256 MiB write)", "field": "tcgen05 GEMM per Dense layer (gemm_tc_kernel)",
                       "parallelism": ("rays sharded over %d ranks, no collective" % world) if world > 1 else "single GPU",
                       "wall_s_timed_region": t_wall, "launch": "one CUDA graph per step (%d library kernels + torch rand/cat/fill nodes)" % dev_step.kernels_per_replay},
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps_in_flight": 2,
                    "how": "GraphedModelStep(host_io=True): host batch -> pinned staging -> graph (H2D node, kernels, D2H node) -> host rgb / depth / 8 loss "
                           "terms; step i+1 staged and enqueued before step i is read; wall clock over all steps, L2 flush included"},
            "gpu_launches": launches, "roofline": roof, "losses_last_step": losses,
            "algorithmic_tflops_per_s_whole_step": C3_FLOPS_PER_RAY * n_rays / (ms_per_step * 1e-3) / 1e12,
        }
        if split is not None:
            line["split_precision"] = split
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if parity is not None:
            line["parity"] = parity
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_reference_c3(args):
    """--impl reference --config c3: the JAX reference cannot run in this image (no jax / flax / gin); the CPU arm is the numpy
    restatement oracle/mip360_model_oracle.py (kind "port"), a bounded sample of the same workload per step."""
    if int(os.environ.get("RANK", 0)) != 0:
        return
    import numpy as np
    import mip360_model_oracle as MM
    import mip360_oracle as mo
    n = 256
    b = c3_batch(n, 6)
    rays = {k: v.numpy() for k, v in b.items()}
    prop, nerf = MM.init_mlp_params(4, 256, False, seed=21), MM.init_mlp_params(8, 1024, True, seed=22)
    g = np.random.default_rng(9)
    u_levels = []
    for ns in C3_SAMPLES:
        base, mj = mo.jitter_base_u(ns)
        u_levels.append((base[None, :] + g.random((n, 1)).astype(np.float32) * mj).astype(np.float32))
    warm, steps = max(1, min(args.warmup, 2)), max(1, min(args.steps, 10))
    for _ in range(warm):
        c3_oracle_step(prop, nerf, rays, u_levels, MM, mo)
    t0 = time.perf_counter()
    for _ in range(steps):
        c3_oracle_step(prop, nerf, rays, u_levels, MM, mo)
    dt = (time.perf_counter() - t0) / steps
    v = n / dt
    cores = os.cpu_count() or 1
    emit({"impl": "reference", "metric": METRIC, "value": v, "unit": "rays/s", "n_gpus": int(os.environ.get("WORLD_SIZE", 1)), "steps": steps, "warmup": warm,
          "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
          "config": {"workload": CONFIGS["c3"]["what"], "name": "c3", "rays_per_step": n, "where": "host cores of the GPU box"},
          "cpu_baseline": {"value": v, "unit": "rays/s", "cores": cores, "kind": "port",
                           "sample": "%d rays/step (oracle/mip360_model_oracle.py, numpy fp32, BLAS threads = host cores; the JAX reference cannot run here)" % n},
          "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})


RESULT_OUT = [None]


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner through C
    stdio when a communicator is created): from here on file descriptor 1 points at stderr, and the JSON line goes to
    a private duplicate of the real stdout."""
    sys.stdout.flush()
    RESULT_OUT[0] = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = RESULT_OUT[0] or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200, help="timed steps (200 x ~1.8 ms: long enough for the clock sampler to see the region)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying the captured step")
    ap.add_argument("--e2e-depth", type=int, default=2, help="host-to-host steps in flight in the e2e leg (PipelinedRenderStep)")
    ap.add_argument("--no-train", action="store_true", help="skip the trainer-step measurement")
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS), help="BASELINE.json config: c2 = configs[1] (headline), c3 = configs[2] (mipnerf360), c4 = configs[3], c5 = configs[4]")
    ap.add_argument("--c3-split", type=int, default=0, help="config c3: 1 = split-precision operands (hi + lo fp16, three MMA passes per layer)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="c2 only: weak = 4096 rays per GPU (default, the driver's curve); strong = 4096 rays in total, sharded over the ranks")
    ap.add_argument("--rays-per-gpu", type=int, default=0, help="override the rays each rank processes")
    ap.add_argument("--overlap", type=int, default=None, help="diagnostic: fg / bg nets on two streams (0 never, 1 library default, 2 always)")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference_c3(args) if args.config == "c3" else run_reference(args)
    elif args.config == "c3":
        run_c3(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
