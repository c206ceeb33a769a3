"""Drop-in for reference `models/hovernet/post_proc.py:94-186` (`process`).

Module-level and picklable like the reference's (it is shipped to pool workers under `spawn`,
infer/tile.py:353-363); each process lazily opens its own post-processing context on
`HVN_DEVICE` / `LOCAL_RANK` / device 0.  Contours come from cv2.findContours on the bbox crop of
the device-produced inst_map exactly as the reference does (:133-143) -- host work per instance,
SURVEY.md row f3."""
import os

import numpy as np

from ... import _lib

_CTX = {}


def _ctx():
    dev = int(os.environ.get("HVN_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    key = (os.getpid(), dev)
    if key not in _CTX:
        _CTX[key] = _lib.Context(dev)
    return _CTX[key]


def table_to_dict(inst, table, nr_types):
    """Instance table rows -> the reference's inst_info_dict (post_proc.py:120-181)."""
    import cv2

    info = {}
    for r in table:
        iid, rmin, cmin, rmax, cmax, area, sx, sy, tp, tc = (int(v) for v in r)
        crop = (inst[rmin:rmax, cmin:cmax] == iid).astype(np.uint8)
        cnt = cv2.findContours(crop, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
        cnt = np.squeeze(cnt[0][0].astype("int32"))
        if cnt.shape[0] < 3:  # < 3 points dont make a contour (:140-141)
            continue
        if len(cnt.shape) != 2:
            continue
        m = cv2.moments(crop)  # the reference's exact float path for the centroid (== sum/area)
        cen = np.array([m["m10"] / m["m00"], m["m01"] / m["m00"]])
        cnt[:, 0] += cmin
        cnt[:, 1] += rmin
        cen[0] += cmin
        cen[1] += rmin
        info[iid] = {"bbox": np.array([[rmin, cmin], [rmax, cmax]]), "centroid": cen, "contour": cnt,
                     "type_prob": None, "type": None}
        if nr_types is not None:
            info[iid]["type"] = int(tp)
            info[iid]["type_prob"] = float(tc / (area + 1.0e-6))
    return info


def process(pred_map, nr_types=None, return_centroids=False):
    """pred_map [H,W,C] -> (pred_inst int32 [H,W], inst_info_dict or None)."""
    pm = np.asarray(pred_map)
    if nr_types is not None:
        # reference: pred_type = pred_map[..., :1].astype(int32); pred_inst = pred_map[..., 1:]
        if pm.shape[-1] != 4:
            raise ValueError("typed post-processing expects [tp, np, hv_x, hv_y] channels")
    elif pm.shape[-1] != 3:
        raise ValueError("seg-only post-processing expects [np, hv_x, hv_y] channels")
    inst, table, nrows = _ctx().postproc(pm.astype(np.float32, copy=False), nr_types)
    inst = inst[0]
    info = None
    if return_centroids or nr_types is not None:
        info = table_to_dict(inst, table[0, : int(nrows[0])], nr_types)
    return inst, info
