"""Known-answer grids and invariants for the UNPINNED step: the restated skimage-0.17.2 watershed
(reference call site models/hovernet/post_proc.py:88; scikit-image is not installed here)."""
import numpy as np
from scipy import ndimage

from hover_net_b200 import synth


def test_two_basins_split_at_ridge(oracle_pp):
    img = np.array([[0, 1, 2, 3, 2, 1, 0]], dtype=np.float64).repeat(3, axis=0)
    mk = np.zeros_like(img, dtype=np.int32)
    mk[1, 0] = 1
    mk[1, 6] = 2
    out = oracle_pp.watershed(img, mk, np.ones_like(mk))
    # value 3 column: reached first from the side whose value-2 pixel pops first. Both value-2
    # columns tie on value; the left one is pushed earlier (lower age) => ridge goes to label 1.
    assert np.array_equal(out, np.array([[1, 1, 1, 1, 2, 2, 2]] * 3))


def test_label_at_push_time_and_neighbour_order(oracle_pp):
    # plateau: all values equal => pure age (FIFO) order => BFS from the markers in raster push
    # order with neighbour order up, left, right, down.
    img = np.zeros((5, 5))
    mk = np.zeros((5, 5), np.int32)
    mk[0, 0] = 7
    mk[4, 4] = 9
    out = oracle_pp.watershed(img, mk, np.ones((5, 5), np.int32))
    yy, xx = np.mgrid[0:5, 0:5]
    d7, d9 = yy + xx, (4 - yy) + (4 - xx)
    expect = np.where(d7 <= d9, 7, 9)  # ties (anti-diagonal) go to the earlier-pushed marker 7
    assert np.array_equal(out, expect)


def test_mask_confines_flood_and_unreached_stay_zero(oracle_pp):
    img = np.zeros((4, 9))
    mask = np.ones((4, 9), np.int32)
    mask[:, 4] = 0
    mk = np.zeros((4, 9), np.int32)
    mk[0, 0] = 3
    mk[0, 4] = 5  # marker outside the mask is dropped (markers * mask)
    out = oracle_pp.watershed(img, mk, mask)
    assert (out[:, :4] == 3).all() and (out[:, 4:] == 0).all()


def test_invariants_on_synthetic_nuclei(oracle_pp):
    for seed in range(4):
        pm = synth.synth_pred_map(120, 140, None, seed)
        inst, st = oracle_pp.proc_np_hv(pm, True)
        mk, blb = st["marker"], st["blb"]
        keep = (mk > 0) & (blb > 0)
        assert np.array_equal(inst[keep], mk[keep])            # markers keep their labels
        assert set(np.unique(inst)) <= set(np.unique(mk)) | {0}  # output ids are marker ids
        assert (inst[blb == 0] == 0).all()                      # never leaves the mask
        comp, n = ndimage.label(blb)
        for c in range(1, n + 1):
            has_marker = (mk[comp == c] > 0).any()
            assert ((inst[comp == c] > 0).all()) == bool(has_marker)
