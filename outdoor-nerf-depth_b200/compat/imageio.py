"""Stand-in for the two ``imageio`` calls of the reference (imread: data_loader_split.py:103, nerf_sample_ray_split.py:74-100;
imwrite: ddp_train_nerf.py:564-621) on top of OpenCV: RGB channel order, uint8 / uint16 kept.  Used only when the real
package is not installed."""
import cv2
import numpy as np


def imread(path, *a, **k):
    px = cv2.imread(str(path), cv2.IMREAD_UNCHANGED)
    if px is None:
        raise IOError("cannot read image %s" % path)
    if px.ndim == 3:
        px = px[:, :, ::-1] if px.shape[2] == 3 else px[:, :, [2, 1, 0, 3]]
    return np.ascontiguousarray(px)


def imwrite(path, im, *a, **k):
    im = np.asarray(im)
    if im.ndim == 3:
        im = im[:, :, ::-1] if im.shape[2] == 3 else im[:, :, [2, 1, 0, 3]]
    if not cv2.imwrite(str(path), np.ascontiguousarray(im)):
        raise IOError("cannot write image %s" % path)


imsave = imwrite
