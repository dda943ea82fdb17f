"""A15 over NCCL with more than one rank (VERDICT r1 weak point 9): render_single_image (ddp_train_nerf.py:133-249) with the
real CUDA renderer on 2 GPUs, bands merged by ONE NCCL all-gather, against the single-GPU render.  Skipped below 2 GPUs."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_path):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import conftest  # noqa: F401
    import torch.distributed as dist
    import nerfpp_oracle as O
    from test_parity_gpu import make_models
    from nerfpp_b200 import render_single_image
    from nerfpp_b200.ray_sampler import DeviceRaySampler
    import synth_scene
    import numpy as np
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        levels = [O.densify(p, 5.0) for p in O.make_params_levels(2)]
        nets = [m.to("cuda:%d" % rank) for m in make_models(levels)]
        K, c2w = synth_scene.camera(3, 12, np.random.RandomState(0))
        sampler = DeviceRaySampler(synth_scene.H, synth_scene.W, K.astype(np.float32), c2w.astype(np.float32), device="cuda:%d" % rank)
        models = {"cascade_level": 2, "cascade_samples": [64, 128], "net_0": nets[0], "net_1": nets[1]}
        ret = render_single_image(rank, world, models, sampler, 1024)
        if rank == 0:
            torch.save({k: ret[-1][k] for k in ("rgb", "depth", "fg_rgb", "bg_lambda")}, out_path)
        else:
            assert ret is None
    finally:
        dist.destroy_process_group()


def test_render_single_image_two_ranks_nccl(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    two, one = str(tmp_path / "two.pt"), str(tmp_path / "one.pt")
    mp.spawn(_worker, args=(2, port, two), nprocs=2, join=True)
    mp.spawn(_worker, args=(1, port + 1, one), nprocs=1, join=True)
    a, b = torch.load(two), torch.load(one)
    for k in a:
        assert a[k].shape == b[k].shape, k
        assert torch.equal(a[k], b[k]), k            # same kernels on the same rays: the banding must not change a bit


def _worker_c3(rank, world, port, out_path):
    """Config 3: mip360_model.render_image with every rank rendering its band of the image, bands merged by ONE all-gather."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import conftest  # noqa: F401
    import torch.distributed as dist
    import mip360_model_oracle as MM
    from nerfpp_b200.mip360_model import Model, Rays, render_image
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        dev = torch.device("cuda", rank)
        h, w = 13, 17                                                    # 221 rays: the bands are ragged
        rays = MM.synthetic_rays(h * w, seed=12)
        R = Rays(*(torch.from_numpy(rays[k]).to(dev).reshape(h, w, -1) for k in ("origins", "directions", "viewdirs", "radii", "near", "far")))
        model = Model(dev)
        model.nerf_mlp.load(MM.init_mlp_params(8, 1024, True, seed=5))
        model.prop_mlp.load(MM.init_mlp_params(4, 256, False, seed=6))
        out = render_image(model, R, render_chunk_size=64, process_group=dist.group.WORLD if world > 1 else None)
        if rank == 0:
            torch.save({k: v.cpu() for k, v in out.items()}, out_path)
    finally:
        dist.destroy_process_group()


def test_mip360_render_image_two_ranks_nccl(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    two, one = str(tmp_path / "two_c3.pt"), str(tmp_path / "one_c3.pt")
    mp.spawn(_worker_c3, args=(2, port, two), nprocs=2, join=True)
    mp.spawn(_worker_c3, args=(1, port + 1, one), nprocs=1, join=True)
    a, b = torch.load(two), torch.load(one)
    assert a["rgb"].shape == (13, 17, 3)
    for k in a:
        assert a[k].shape == b[k].shape, k
        assert torch.equal(a[k], b[k]), k
