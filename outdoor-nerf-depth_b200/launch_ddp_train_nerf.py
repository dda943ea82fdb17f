"""Launcher that runs the reference's UNMODIFIED ``ddp_train_nerf.py`` / ``ddp_test_nerf.py`` on the B200 kernels.

    python launch_ddp_train_nerf.py --reference /path/to/outdoor-nerf-depth/nerf-methods/nerfplusplus \
           --config configs/kitti.txt ...            (every other flag is the trainer's own, ddp_train_nerf.py:657-727)

How the drop-in works:
  * ``ddp_model`` and ``depth_loss`` are resolved by name through ``sys.path``: this directory is put AHEAD of the
    reference directory, so the trainer's ``from ddp_model import NerfNetWithAutoExpo`` / ``from depth_loss import *``
    (ddp_train_nerf.py:9,18) bind the B200 modules; its own ``utils``, ``data_loader_split`` ... still come from the
    reference.
  * ``intersect_sphere``, ``perturb_samples`` and ``sample_pdf`` are defined INSIDE the trainer module
    (ddp_train_nerf.py:51-130) and children are started with the ``spawn`` method (fresh import per child,
    :742-745), so they are patched inside each child before ``ddp_train_nerf.ddp_train_nerf(rank, args)`` runs.
    ``ddp_test_nerf.py:16`` imports ``render_single_image`` from the trainer module, so the same patch covers it.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def _paths(reference_dir):
    for p in (reference_dir, HERE):          # HERE ends up first
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)


def patch_trainer_module(mod):
    """Rebinds the three sampling functions of an imported ``ddp_train_nerf`` module to the CUDA ones."""
    from nerfpp_b200 import ops
    mod.intersect_sphere = ops.intersect_sphere
    mod.perturb_samples = ops.perturb_samples
    mod.sample_pdf = ops.sample_pdf
    return mod


def _child(rank, args, reference_dir, entry):
    _paths(reference_dir)
    import importlib
    trainer = patch_trainer_module(importlib.import_module("ddp_train_nerf"))
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
    _paths(reference_dir)
    import torch
    import ddp_train_nerf as trainer          # the reference's module, unmodified
    parser = trainer.config_parser()
    args = parser.parse_args(argv)
    if args.world_size == -1:                 # ddp_train_nerf.py:737-739
        args.world_size = torch.cuda.device_count()
    torch.multiprocessing.spawn(_child, args=(args, reference_dir, entry), nprocs=args.world_size, join=True)


if __name__ == "__main__":
    main()
