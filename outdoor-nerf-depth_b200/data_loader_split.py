"""Drop-in for the reference's ``data_loader_split`` module (nerf-methods/nerfplusplus/data_loader_split.py:27-129),
SURVEY.md section 8(f) N2: same ``load_data_split`` signature, same directory format, same file order / ``skip`` /
``scale`` / ``max_depth.txt`` handling -- but it returns device-resident samplers (nerfpp_b200.DeviceRaySampler): each
image's raw PNG pixels (1-2 bytes each) are uploaded once and decoded on the GPU (``nerfpp_decode_pixels``), and rays are
generated per batch by ``nerfpp_gen_rays`` instead of being precomputed for every pixel on the host.

Differences from the reference, on purpose:
  * the PNGs are read with OpenCV (imageio is not a dependency); PNG is lossless, the pixels are the same;
  * a split without a ``depth/`` directory works (the reference raises ``IndexError`` there: it assigns the placeholder
    list to ``depth_files`` but reads ``depth_gt_files``, data_loader_split.py:83-89);
  * only resolution_level 1 (the only level the trainer uses, ddp_train_nerf.py never calls set_resolution_level).
"""
import fnmatch
import logging
import os

import cv2
import numpy as np

from nerfpp_b200.ray_sampler import DeviceRaySampler

logger = logging.getLogger(__package__)


def find_files(dir, exts):                        # noqa: A002  (the reference's parameter names: callers pass exts=...)
    """Sorted paths of the directory entries matching any of the shell patterns ([] when the directory is missing) --
    what data_loader_split.py:14-24 returns."""
    try:
        names = os.listdir(dir)
    except (FileNotFoundError, NotADirectoryError):
        return []
    # (like glob, a leading-dot name is not matched by '*')
    return sorted(os.path.join(dir, nm) for nm in names if not nm.startswith('.') and any(fnmatch.fnmatch(nm, pat) for pat in exts))


def read_image(path):
    """Raw pixels in the layout imageio.imread gives the reference: HxWx3 uint8 RGB, HxW uint8 or uint16."""
    px = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if px is None:
        raise IOError("cannot read image %s" % path)
    if px.ndim == 3:                              # OpenCV decodes to BGR(A)
        px = px[:, :, ::-1] if px.shape[2] == 3 else px[:, :, [2, 1, 0, 3]]
    return np.ascontiguousarray(px)


def parse_txt(filename):
    """A 4x4 matrix stored as 16 whitespace-separated numbers, row-major (data_loader_split.py:29-32) -> float32."""
    with open(filename) as f:
        vals = np.array(f.read().split(), dtype=np.float64)
    if vals.size != 16:
        raise ValueError("%s: expected 16 numbers, found %d" % (filename, vals.size))
    return vals.reshape(4, 4).astype(np.float32)


def list_split(basedir, scene, split, skip=1, try_load_min_depth=True, depth_sup_type='gt'):
    """The host half of load_data_split: which files make up the split (no pixel is read except the first training image,
    for H and W, as in the reference).  Returns a dict of per-camera lists plus H, W, depth_scale, max_depth."""
    if basedir[-1] == '/':
        basedir = basedir[:-1]
    split_dir = '{}/{}/{}'.format(basedir, scene, split)
    img_exts = ['*.png', '*.jpg']
    intrinsics_files = find_files('{}/intrinsics'.format(split_dir), exts=['*.txt'])[::skip]
    pose_files = find_files('{}/pose'.format(split_dir), exts=['*.txt'])[::skip]
    cam_cnt = len(pose_files)

    def optional(sub, enabled=True):
        files = find_files('{}/{}'.format(split_dir, sub), exts=img_exts) if enabled else []
        if len(files) > 0:
            files = files[::skip]
            assert len(files) == cam_cnt, '{}: {} files for {} cameras'.format(sub, len(files), cam_cnt)
            return files
        return [None, ] * cam_cnt

    img_files = optional('rgb')
    mask_files = optional('mask')
    mindepth_files = optional('min_depth', try_load_min_depth)
    depth_gt_files = optional('depth')
    depth_scale = None
    if depth_gt_files and depth_gt_files[0] is not None:
        depth_scale = float(open(os.path.join(basedir, scene, 'scale')).readlines()[0].strip())
    suffix = '_' + depth_sup_type if depth_sup_type != 'gt' else ''
    depth_sup_files = optional('depth{}'.format(suffix))
    if depth_sup_files and depth_sup_files[0] is not None and depth_scale is None:
        depth_scale = float(open(os.path.join(basedir, scene, 'scale')).readlines()[0].strip())
    # assume all images have the same size as training image
    train_imgfile = find_files('{}/{}/train/rgb'.format(basedir, scene), exts=img_exts)[0]
    H, W = read_image(train_imgfile).shape[:2]
    try:
        max_depth = float(open('{}/max_depth.txt'.format(split_dir)).readline().strip())
    except Exception:
        max_depth = None
    return dict(intrinsics=intrinsics_files, pose=pose_files, rgb=img_files, mask=mask_files, min_depth=mindepth_files,
                depth_gt=depth_gt_files, depth_sup=depth_sup_files, H=H, W=W, depth_scale=depth_scale, max_depth=max_depth,
                cam_cnt=cam_cnt)


def load_data_split(basedir, scene, split, skip=1, try_load_min_depth=True, only_img_files=False, depth_sup_type='gt',
                    device='cuda'):
    if only_img_files:
        b = basedir[:-1] if basedir[-1] == '/' else basedir
        return find_files('{}/{}/{}/rgb'.format(b, scene, split), exts=['*.png', '*.jpg'])
    ls = list_split(basedir, scene, split, skip, try_load_min_depth, depth_sup_type)
    ray_samplers = []
    for i in range(ls['cam_cnt']):
        ray_samplers.append(DeviceRaySampler.from_files(
            H=ls['H'], W=ls['W'], intrinsics=parse_txt(ls['intrinsics'][i]), c2w=parse_txt(ls['pose'][i]),
            img_path=ls['rgb'][i], mask_path=ls['mask'][i], min_depth_path=ls['min_depth'][i], max_depth=ls['max_depth'],
            depth_gt_path=ls['depth_gt'][i], depth_sup_path=ls['depth_sup'][i], depth_scale=ls['depth_scale'],
            read_image=read_image, device=device))
    logger.info('Split {}, # views: {}'.format(split, ls['cam_cnt']))
    return ray_samplers
