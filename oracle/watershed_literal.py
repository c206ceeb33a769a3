"""ORACLE (test infrastructure, not product): a second, independent restatement of
`skimage.segmentation.watershed(image, markers, mask=mask)` as scikit-image 0.17.2 runs it for the one call the
reference makes (models/hovernet/post_proc.py:88: connectivity=1, compactness=0, watershed_line=False).

Written literally from SURVEY.md Appendix B in pure Python -- padded raveled arrays, the neighbour offsets
(-W', -1, +1, +W'), an array-based binary heap with heap_general.pxi's sift rules -- and deliberately sharing no
code with oracle/postproc_oracle.c (`hvo_watershed`), so that the two can be cross-checked
(tests/test_oracle_watershed.py).  PARITY UNPINNED: scikit-image is not installed in this image; when it is,
`python -m oracle.regen_with_skimage` checks both restatements against the real function and regenerates the
post-processing goldens with it.

`tie_break`: how two entries with equal (value, age) -- only possible between age-0 marker pixels -- are ordered:
  "heap"   skimage: no tie rule, the outcome is whatever the binary heap's layout yields (Appendix B);
  "raster" by pixel index: the strict total order (value, age, index) that the device floods use
           (hover_net_b200/csrc/postproc.cu).  Comparing the two measures the declared corner case.
"""
import numpy as np


class _Heap(object):
    """heap_general.pxi: push = append + sift-up while smaller(child, parent), parent = (child + 1) // 2 - 1;
    pop = root out, last to root, sift-down taking the left child if smaller than the current smallest, then the
    right one if smaller than that; stop when nothing moved."""

    def __init__(self, smaller):
        self.a = []
        self.smaller = smaller

    def push(self, e):
        a = self.a
        a.append(e)
        child = len(a) - 1
        while child > 0:
            parent = (child + 1) // 2 - 1
            if self.smaller(a[child], a[parent]):
                a[child], a[parent] = a[parent], a[child]
                child = parent
            else:
                break

    def pop(self):
        a = self.a
        top = a[0]
        last = a.pop()
        if a:
            a[0] = last
            n = len(a)
            i = 0
            while True:
                l, r, s = 2 * i + 1, 2 * i + 2, i
                if l < n and self.smaller(a[l], a[s]):
                    s = l
                if r < n and self.smaller(a[r], a[s]):
                    s = r
                if s == i:
                    break
                a[s], a[i] = a[i], a[s]
                i = s
        return top


def watershed(image, markers, mask, tie_break="heap", count_ties=None):
    """image [H,W] (any float; used as float64), markers [H,W] int, mask [H,W] -> int32 [H,W].
    count_ties: optional dict; ["marker_ties"] receives the number of heap comparisons that found two entries
    with equal (value, age)."""
    image = np.asarray(image, dtype=np.float64)
    H, W = image.shape
    Wp = W + 2
    # _validate_inputs + pad by one pixel of (0, no marker, outside the mask)
    img = np.zeros((H + 2) * Wp, dtype=np.float64)
    msk = np.zeros((H + 2) * Wp, dtype=bool)
    out = np.zeros((H + 2) * Wp, dtype=np.int32)
    inner = (np.arange(1, H + 1)[:, None] * Wp + np.arange(1, W + 1)[None, :]).ravel()
    img[inner] = image.ravel()
    msk[inner] = np.asarray(mask).astype(bool).ravel()
    out[inner] = (np.asarray(markers).astype(np.int32) * np.asarray(mask).astype(bool)).ravel()
    offsets = (-Wp, -1, 1, Wp)  # connectivity 1, centre removed, raveled order: up, left, right, down
    ties = [0]

    if tie_break == "heap":
        def smaller(a, b):
            if a[0] != b[0]:
                return a[0] < b[0]
            if a[1] == b[1]:
                ties[0] += 1
            return a[1] < b[1]
    else:
        def smaller(a, b):
            return a < b  # (value, age, index) tuples: strict total order

    heap = _Heap(smaller)
    img_l, msk_l, out_l = img.tolist(), msk.tolist(), out.tolist()
    for idx in np.flatnonzero(out).tolist():  # marker pixels, raveled order, age 0
        heap.push((img_l[idx], 0, idx))
    age = 1
    while heap.a:
        value, _, index = heap.pop()
        for off in offsets:
            n = index + off
            if not msk_l[n]:
                continue
            if out_l[n]:
                continue
            age += 1
            out_l[n] = out_l[index]  # label at push time
            heap.push((img_l[n], age, n))
    if count_ties is not None:
        count_ties["marker_ties"] = ties[0]
    return np.asarray(out_l, dtype=np.int32).reshape(H + 2, Wp)[1:-1, 1:-1].copy()
