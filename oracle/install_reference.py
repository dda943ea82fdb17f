"""Copies the reference's NeRF++ Python files into the git-ignored ``baseline/_ref/nerfplusplus`` (TEST INFRASTRUCTURE).

    python oracle/install_reference.py [--src /root/reference/nerf-methods/nerfplusplus]

``/root/reference`` exists only in the build container; ``baseline/_ref/`` is listed in .gitignore (so no reference source
ever enters the history) but NOT in .gpurunignore, so the copy travels to the GPU box with the built ``.so``.  It is what
``bench.py --impl reference`` times (the reference's own functions, unmodified, on the host cores) and what
``launch_ddp_train_nerf.py --reference baseline/_ref/nerfplusplus`` drives in tests/test_trainer_gpu.py.  The reference is a
directory of scripts without setup.py / pyproject.toml: there is nothing to ``pip install``, a file copy IS the install.
"""
import argparse
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/nerf-methods/nerfplusplus"
DST = os.path.join(ROOT, "baseline", "_ref", "nerfplusplus")
FILES = ("ddp_train_nerf.py", "ddp_test_nerf.py", "ddp_model.py", "nerf_network.py", "depth_loss.py", "utils.py",
         "nerf_sample_ray_split.py", "data_loader_split.py")


def install(src=SRC, dst=DST):
    """Returns dst, or None when the reference tree is not present (the GPU box: the copy made here is used)."""
    if not os.path.isdir(src):
        return dst if os.path.isdir(dst) else None
    os.makedirs(os.path.join(dst, "configs"), exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))
    for f in os.listdir(os.path.join(src, "configs")):
        if os.path.isfile(os.path.join(src, "configs", f)):
            shutil.copyfile(os.path.join(src, "configs", f), os.path.join(dst, "configs", f))
    return dst


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default=SRC)
    print(install(ap.parse_args().src))
