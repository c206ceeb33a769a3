"""ORACLE (test infrastructure, not product): ctypes binding + Python tail of the CPU restatement.

`process()` below mirrors reference `models/hovernet/post_proc.py:94-186` on top of
`oracle/postproc_oracle.c`; the per-instance contour / <3-point drop uses cv2.findContours exactly
as the reference does (post_proc.py:133-143).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
ROW = 10


class Stages(ctypes.Structure):
    _fields_ = [("blb", ctypes.c_void_p), ("sobelh", ctypes.c_void_p), ("sobelv", ctypes.c_void_p),
                ("overall32", ctypes.c_void_p), ("dist", ctypes.c_void_p), ("marker", ctypes.c_void_p)]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "libhvo.so"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libhvo.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.hvo_label4.restype = ctypes.c_int
        _LIB.hvo_process.restype = ctypes.c_int
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def label4(binary):
    b = np.ascontiguousarray(binary, dtype=np.int32)
    out = np.empty_like(b)
    n = lib().hvo_label4(_p(b), b.shape[0], b.shape[1], _p(out))
    return out, n


def remove_small(lab, min_size=10):
    out = np.ascontiguousarray(lab, dtype=np.int32).copy()
    lib().hvo_remove_small(_p(out), out.size, int(out.max()), int(min_size))
    return out


def normalize(src):
    s = np.ascontiguousarray(src)
    out = np.empty(s.shape, dtype=np.float32)
    if s.dtype == np.float32:
        lib().hvo_normalize_f32(_p(s), s.size, _p(out))
    else:
        s = s.astype(np.float64)
        lib().hvo_normalize_f64(_p(s), s.size, _p(out))
    return out


def sobel21(src, dx):
    s = np.ascontiguousarray(src, dtype=np.float32)
    out = np.empty(s.shape, dtype=np.float64)
    lib().hvo_sobel21(_p(s), s.shape[0], s.shape[1], int(dx), _p(out))
    return out


def gauss3(src):
    s = np.ascontiguousarray(src, dtype=np.float64)
    out = np.empty_like(s)
    lib().hvo_gauss3_f64(_p(s), s.shape[0], s.shape[1], _p(out))
    return out


def fill_holes(binary):
    b = np.ascontiguousarray(binary, dtype=np.int32)
    out = np.empty(b.shape, dtype=np.uint8)
    lib().hvo_fill_holes(_p(b), b.shape[0], b.shape[1], _p(out))
    return out


def open_ellipse5(src):
    s = np.ascontiguousarray(src, dtype=np.uint8)
    out = np.empty_like(s)
    lib().hvo_open_ellipse5(_p(s), s.shape[0], s.shape[1], _p(out))
    return out


def watershed(image, markers, mask):
    im = np.ascontiguousarray(image, dtype=np.float64)
    mk = np.ascontiguousarray(markers, dtype=np.int32)
    ms = np.ascontiguousarray(mask, dtype=np.int32)
    out = np.empty_like(mk)
    lib().hvo_watershed(_p(im), _p(mk), _p(ms), im.shape[0], im.shape[1], _p(out))
    return out


def proc_np_hv(pred, want_stages=False):
    """pred float32 [H,W,3].  Returns inst int32 [H,W] (and the stage dict)."""
    pred = np.ascontiguousarray(pred, dtype=np.float32)
    H, W, C = pred.shape
    inst = np.empty((H, W), dtype=np.int32)
    if not want_stages:
        lib().hvo_proc_np_hv(_p(pred), C, H, W, _p(inst), None)
        return inst
    st = dict(blb=np.empty((H, W), np.int32), sobelh=np.empty((H, W), np.float64),
              sobelv=np.empty((H, W), np.float64), overall32=np.empty((H, W), np.float32),
              dist=np.empty((H, W), np.float64), marker=np.empty((H, W), np.int32))
    s = Stages(*[_p(st[k]) for k in ("blb", "sobelh", "sobelv", "overall32", "dist", "marker")])
    lib().hvo_proc_np_hv(_p(pred), C, H, W, _p(inst), ctypes.byref(s))
    return inst, st


def process_table(pred_map, nr_types=None):
    """inst int32 [H,W], table int64 [n,10] (id,rmin,cmin,rmax,cmax,area,sum_x,sum_y,type,type_cnt)."""
    pm = np.ascontiguousarray(pred_map, dtype=np.float32)
    H, W, C = pm.shape
    inst = np.empty((H, W), dtype=np.int32)
    cap = H * W // 10 + 16
    table = np.zeros((cap, ROW), dtype=np.int64)
    n = lib().hvo_process(_p(pm), H, W, C, int(nr_types or 0), _p(inst), _p(table), cap)
    assert n >= 0
    return inst, table[:n].copy()


def contour(inst, iid, rmin, cmin, rmax, cmax, cap=4096):
    """[n,2] int32 (x, y) in map coordinates: the oracle's restatement of
    cv2.findContours(crop, RETR_TREE, CHAIN_APPROX_SIMPLE)[0][0] + (cmin, rmin)."""
    a = np.ascontiguousarray(inst, dtype=np.int32)
    pts = np.empty((cap, 2), dtype=np.int32)
    lib().hvo_contour.restype = ctypes.c_int
    n = lib().hvo_contour(_p(a), a.shape[0], a.shape[1], int(iid), int(rmin), int(cmin), int(rmax), int(cmax), _p(pts), cap)
    assert 0 <= n <= cap
    return pts[:n].copy()


def process(pred_map, nr_types=None, return_centroids=False):
    """Reference-shaped result: (pred_inst, inst_info_dict or None)  -- post_proc.py:94-186."""
    import cv2

    inst, table = process_table(pred_map, nr_types)
    info = None
    if return_centroids or nr_types is not None:
        info = {}
        for r in table:
            iid, rmin, cmin, rmax, cmax, area, sx, sy, tp, tc = (int(v) for v in r)
            crop = (inst[rmin:rmax, cmin:cmax] == iid).astype(np.uint8)
            cnt = cv2.findContours(crop, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
            cnt = np.squeeze(cnt[0][0].astype("int32"))
            if cnt.shape[0] < 3 or len(cnt.shape) != 2:
                continue
            m = cv2.moments(crop)
            cen = np.array([m["m10"] / m["m00"] + cmin, m["m01"] / m["m00"] + rmin])
            cnt[:, 0] += cmin
            cnt[:, 1] += rmin
            info[iid] = {"bbox": np.array([[rmin, cmin], [rmax, cmax]]), "centroid": cen,
                         "contour": cnt, "type_prob": None, "type": None}
            if nr_types is not None:
                info[iid]["type"] = int(tp)
                info[iid]["type_prob"] = float(tc / (area + 1.0e-6))
    return inst, info
