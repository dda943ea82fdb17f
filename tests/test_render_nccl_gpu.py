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
