"""Test helper: the reference post-processing written as direct cv2 / scipy calls.

Same library calls in the same order as reference `models/hovernet/post_proc.py:36-88` and
`misc/utils.py:169-180`, used ONLY to pin `oracle/postproc_oracle.c` stage by stage.  The one call
that cannot be made here is `skimage.segmentation.watershed` (scikit-image is not installed):
the oracle's restatement is substituted, so the last stage is unpinned.
"""
import cv2
import numpy as np
from scipy import ndimage

cv2.setNumThreads(1)


def remove_small(lab, min_size=10):
    sizes = np.bincount(lab.ravel())
    out = lab.copy()
    out[(sizes < min_size)[lab]] = 0
    return out


def minmax01(a):
    return cv2.normalize(a, None, alpha=0, beta=1, norm_type=cv2.NORM_MINMAX, dtype=cv2.CV_32F)


def stages(pred, watershed_fn):
    pred = np.array(pred, dtype=np.float32)
    fg = np.array(pred[..., 0] >= 0.5, dtype=np.int32)
    fg = remove_small(ndimage.label(fg)[0], 10)
    fg[fg > 0] = 1
    hn, vn = minmax01(pred[..., 1]), minmax01(pred[..., 2])
    sh = cv2.Sobel(hn, cv2.CV_64F, 1, 0, ksize=21)
    sv = cv2.Sobel(vn, cv2.CV_64F, 0, 1, ksize=21)
    eh, ev = 1 - minmax01(sh), 1 - minmax01(sv)
    edge32 = np.maximum(eh, ev)
    edge = edge32 - (1 - fg)
    edge[edge < 0] = 0
    dist = -cv2.GaussianBlur((1.0 - edge) * fg, (3, 3), 0)
    strong = np.array(edge >= 0.4, dtype=np.int32)
    mk = fg - strong
    mk[mk < 0] = 0
    mk = ndimage.binary_fill_holes(mk).astype("uint8")
    mk = cv2.morphologyEx(mk, cv2.MORPH_OPEN, cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (5, 5)))
    mk = remove_small(ndimage.label(mk)[0], 10)
    inst = watershed_fn(dist, mk, fg)
    return dict(blb=fg, sobelh=sh, sobelv=sv, overall32=edge32, dist=dist, marker=mk, inst=inst)
