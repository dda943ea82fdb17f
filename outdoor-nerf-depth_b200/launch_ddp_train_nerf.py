"""Launcher that runs the reference's UNMODIFIED ``ddp_train_nerf.py`` / ``ddp_test_nerf.py`` on the B200 kernels.

    python launch_ddp_train_nerf.py --reference /path/to/outdoor-nerf-depth/nerf-methods/nerfplusplus [--test] [--nccl] \
           --config configs/kitti.txt ...            (every other flag is the trainer's own, ddp_train_nerf.py:657-727)

How the drop-in works:
  * ``ddp_model`` and ``depth_loss`` are resolved by name through ``sys.path``: this directory is put AHEAD of the
    reference directory, so the trainer's ``from ddp_model import NerfNetWithAutoExpo`` / ``from depth_loss import *``
    (ddp_train_nerf.py:9,18) bind the B200 modules, and ``from data_loader_split import load_data_split`` (:11) binds
    the device-resident loader (set NERFPP_REFERENCE_LOADER=1 to keep the reference's host
    loader); its own ``utils``, ``nerf_sample_ray_split`` ... still come from the reference.
  * ``intersect_sphere``, ``perturb_samples`` and ``sample_pdf`` are defined INSIDE the trainer module
    (ddp_train_nerf.py:51-130) and children are started with the ``spawn`` method (fresh import per child,
    :742-745), so they are patched inside each child before ``ddp_train_nerf.ddp_train_nerf(rank, args)`` runs.
    ``ddp_test_nerf.py:16`` imports ``render_single_image`` from the trainer module, so the same patch covers it.
  * the trainer imports four non-numeric packages at module scope (configargparse, tensorboardX, imageio, matplotlib);
    where they are not installed the minimal stand-ins in ``compat/`` (appended to the END of sys.path) let it start.

tests/test_trainer_gpu.py runs the unmodified trainer this way (world size 1 and 2) on a synthetic on-disk scene.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


COMPAT = os.path.join(HERE, "compat")


def _paths(reference_dir):
    for p in (reference_dir, HERE):          # HERE ends up first
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    # stand-ins for configargparse / tensorboardX / imageio / matplotlib go LAST: an installed package always wins
    if COMPAT not in sys.path:
        sys.path.append(COMPAT)


def patch_trainer_module(mod):
    """Rebinds the three sampling functions of an imported ``ddp_train_nerf`` module to the CUDA ones."""
    from nerfpp_b200 import ops
    mod.intersect_sphere = ops.intersect_sphere
    mod.perturb_samples = ops.perturb_samples
    mod.sample_pdf = ops.sample_pdf
    return mod


def _use_reference_loader(reference_dir):
    """NERFPP_REFERENCE_LOADER=1: bind ``data_loader_split`` to the reference's own file although this directory, which
    holds a module of the same name, comes first on sys.path."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("data_loader_split", os.path.join(reference_dir, "data_loader_split.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["data_loader_split"] = mod
    spec.loader.exec_module(mod)
    return mod


def patch_process_group(mod):
    """--nccl: the trainer's ``setup`` (ddp_train_nerf.py:292-298) creates a gloo group, so DDP's gradient all-reduce
    stages every bucket through host memory and TCP.  This replacement keeps its rendezvous (localhost, --port) but asks
    for ``cpu:gloo,cuda:nccl``: CUDA tensors (the gradients) go over NCCL / NVLink, CPU tensors (render_single_image's
    gather of host images, :229-243) still find gloo."""
    import torch.distributed as dist

    def setup(rank, world_size, port):
        os.environ['MASTER_ADDR'] = '127.0.0.1'
        os.environ['MASTER_PORT'] = str(port)
        dist.init_process_group("cpu:gloo,cuda:nccl", rank=rank, world_size=world_size)
    mod.setup = setup
    return mod


def _child(rank, args, reference_dir, entry, nccl=False):
    _paths(reference_dir)
    import importlib
    import torch
    # the trainer loads the data (ddp_train_nerf.py:389) before create_nerf() selects the device (:310); the device-
    # resident loader must already see this rank's GPU as the current one
    if torch.cuda.is_available():
        torch.cuda.set_device(rank)
    if os.environ.get("NERFPP_REFERENCE_LOADER") == "1":
        _use_reference_loader(reference_dir)
    trainer = patch_trainer_module(importlib.import_module("ddp_train_nerf"))
    if nccl:
        patch_process_group(trainer)
    if entry == "train":
        trainer.ddp_train_nerf(rank, args)
    else:
        tester = importlib.import_module("ddp_test_nerf")
        tester.render_single_image = trainer.render_single_image
        tester.ddp_test_nerf(rank, args)


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if "--reference" not in argv:
        raise SystemExit("usage: launch_ddp_train_nerf.py --reference <.../nerf-methods/nerfplusplus> [--test] <trainer flags>")
    i = argv.index("--reference")
    reference_dir = os.path.abspath(argv[i + 1])
    del argv[i:i + 2]
    entry = "train"
    if "--test" in argv:
        argv.remove("--test")
        entry = "test"
    nccl = "--nccl" in argv
    if nccl:
        argv.remove("--nccl")
    _paths(reference_dir)
    import torch
    if os.environ.get("NERFPP_REFERENCE_LOADER") == "1":
        _use_reference_loader(reference_dir)
    import ddp_train_nerf as trainer          # the reference's module, unmodified
    parser = trainer.config_parser()
    args = parser.parse_args(argv)
    if args.world_size == -1:                 # ddp_train_nerf.py:737-739
        args.world_size = torch.cuda.device_count()
    torch.multiprocessing.spawn(_child, args=(args, reference_dir, entry, nccl), nprocs=args.world_size, join=True)


if __name__ == "__main__":
    main()
