"""A15 host logic on CPU: band sharding, chunking, packing and the single all-gather of render_single_image
(ddp_train_nerf.py:133-249), exercised with world_size 2 over gloo and a stub per-chunk renderer (the CUDA cascade has
no CPU path; only the plumbing around it is tested here)."""
import os
from collections import OrderedDict

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nerfpp_b200 import render as R

H, W = 6, 8
MODELS = {"cascade_level": 2, "cascade_samples": [4, 6]}


class Sampler:
    """The slice of RaySamplerSingleImage the driver uses (nerf_sample_ray_split.py:131-153)."""
    def __init__(self, h=H, w=W):
        self.H, self.W = h, w

    def get_all(self):
        n = self.H * self.W
        idx = torch.arange(n, dtype=torch.float32)
        return OrderedDict(ray_o=torch.stack([idx, idx + 0.25, idx + 0.5], -1), ray_d=torch.ones(n, 3),
                           min_depth=torch.full((n,), 1e-4), depth=None, rgb=None, mask=None, img_name="x.png")


def stub_render(models, chunk):
    """Deterministic per-ray outputs with the forward dict's keys and shapes."""
    o = chunk["ray_o"]
    rets, tot = [], 0
    for m in range(models["cascade_level"]):
        tot += models["cascade_samples"][m]
        base = o[:, :1] + 1000.0 * m
        ret = OrderedDict(rgb=base + torch.arange(3.), fg_weights=None, bg_weights=None,
                          fg_dists=base + 0.01 * torch.arange(float(tot)), fg_rgb=base + 10 + torch.arange(3.),
                          fg_depth=base[:, 0] + 20, bg_rgb=base + 30 + torch.arange(3.), bg_depth=base[:, 0] + 40,
                          bg_lambda=base[:, 0] + 50, depth=base[:, 0] + 60)
        rets.append(ret)
    return rets


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out = R.render_single_image(rank, world, MODELS, Sampler(), chunk_size=5, render_chunk=stub_render, device=torch.device("cpu"))
        if rank == 0:
            q.put([{k: v.numpy().copy() for k, v in lvl.items()} for lvl in out])   # plain pickles, not shared-memory handles
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


def test_band_sizes_and_divisibility_error():
    assert R.band_sizes(48, 4) == [12, 12, 12, 12]
    with pytest.raises(Exception, match="not divisible"):
        R.band_sizes(50, 4)


def test_single_process_matches_direct_render():
    out = R.render_single_image(0, 1, MODELS, Sampler(), chunk_size=7, render_chunk=stub_render, device=torch.device("cpu"))
    full = stub_render(MODELS, {"ray_o": Sampler().get_all()["ray_o"]})
    assert len(out) == 2
    for m in range(2):
        assert list(out[m].keys()) == list(R.RENDER_KEYS)
        for k in R.RENDER_KEYS:
            want = full[m][k].reshape(H, W, -1).squeeze()
            assert out[m][k].shape == want.shape and torch.equal(out[m][k], want), (m, k)


def test_two_ranks_gloo_match_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    ref = R.render_single_image(0, 1, MODELS, Sampler(), chunk_size=48, render_chunk=stub_render, device=torch.device("cpu"))
    for m in range(2):
        for k in R.RENDER_KEYS:
            assert torch.equal(torch.from_numpy(got[m][k]), ref[m][k]), (m, k)


def test_lazy_render_dict_behaves_like_the_plain_dict():
    """The per-level result of render_single_image keeps the reference's keys and order; a deferred entry is fetched by
    any access that could see its value, exactly once."""
    from collections import OrderedDict
    import torch
    from nerfpp_b200.render import _LazyRenderDict
    calls = []

    def fetch():
        calls.append(1)
        return torch.arange(6.).reshape(2, 3)
    d = _LazyRenderDict([("rgb", torch.zeros(2)), ("fg_dists", None), ("depth", torch.ones(2))], {"fg_dists": fetch})
    assert list(d.keys()) == ["rgb", "fg_dists", "depth"] and len(d) == 3 and "fg_dists" in d and not calls
    assert torch.equal(d["rgb"], torch.zeros(2)) and not calls
    assert d["fg_dists"].shape == (2, 3) and len(calls) == 1
    assert d.get("fg_dists").shape == (2, 3) and len(calls) == 1
    d2 = _LazyRenderDict([("a", torch.zeros(1)), ("fg_dists", None)], {"fg_dists": fetch})
    assert [tuple(v.shape) for v in d2.values()] == [(1,), (2, 3)] and len(calls) == 2
    d3 = _LazyRenderDict([("fg_dists", None)], {"fg_dists": fetch})
    assert dict(d3.items())["fg_dists"].shape == (2, 3)
    d4 = _LazyRenderDict([("fg_dists", None), ("b", torch.zeros(1))], {"fg_dists": fetch})
    c = d4.copy()
    assert isinstance(c, OrderedDict) and c["fg_dists"].shape == (2, 3) and d4.pop("fg_dists").shape == (2, 3)


# ---- config 3: mip360_model.render_image's banding / chunking / single all-gather, world size 2 over gloo -----------------
def _stub_chunk(chunk):
    """Per-ray outputs with the last level's keys (models.py:270-305), deterministic functions of the ray origin."""
    o = chunk.origins[:, :1]
    from nerfpp_b200.mip360_model import RENDER_KEYS
    out = {k: o[:, 0] + 10.0 * i for i, k in enumerate(RENDER_KEYS)}
    out["rgb"] = o + torch.arange(3.)
    return out


def _rays_c3(h, w):
    from nerfpp_b200.mip360_model import Rays
    idx = torch.arange(h * w, dtype=torch.float32).reshape(h, w, 1)
    return Rays(idx.repeat(1, 1, 3), torch.ones(h, w, 3), torch.ones(h, w, 3), torch.ones(h, w, 1), torch.ones(h, w, 1), torch.ones(h, w, 1))


def _worker_c3(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from nerfpp_b200.mip360_model import render_image
        out = render_image(None, _rays_c3(5, 7), render_chunk_size=4, process_group=dist.group.WORLD, render_chunk=_stub_chunk,
                           device=torch.device("cpu"))
        q.put((rank, {k: v.numpy().copy() for k, v in out.items()}))
    finally:
        dist.destroy_process_group()


def test_mip360_render_image_bands_over_gloo():
    from nerfpp_b200.mip360_model import band, render_image
    assert band(35, 2, 0) == (0, 18, 18) and band(35, 2, 1) == (18, 35, 18) and band(3, 8, 7) == (3, 3, 1)
    single = render_image(None, _rays_c3(5, 7), render_chunk_size=4, render_chunk=_stub_chunk, device=torch.device("cpu"))
    assert single["rgb"].shape == (5, 7, 3) and single["depth"].shape == (5, 7)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = [ctx.Process(target=_worker_c3, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for r in (0, 1):                      # every rank ends up with the whole image (35 rays: ragged bands of 18 + 17)
        for k, v in single.items():
            assert (got[r][k] == v.numpy()).all(), (r, k)
