"""Pins the oracle's contour restatement (oracle/postproc_oracle.c hvo_contour) against cv2 4.13's
`findContours(crop, RETR_TREE, CHAIN_APPROX_SIMPLE)[0][0]` -- the call of reference post_proc.py:133-137 --
on connected masks: random 4-connected growths (thin limbs, diagonals, holes), known shapes, and every
instance of the synthetic nuclei maps."""
import cv2
import numpy as np
import pytest


def _ref(mask):
    c = cv2.findContours(mask.astype(np.uint8), cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
    return c[0][0].reshape(-1, 2).astype(np.int32)


def _grow(rng, h, w, n):
    """random 4-connected region of ~n pixels"""
    m = np.zeros((h, w), bool)
    y, x = rng.integers(0, h), rng.integers(0, w)
    m[y, x] = True
    cells = [(y, x)]
    while len(cells) < n:
        y, x = cells[rng.integers(0, len(cells))]
        dy, dx = ((0, 1), (1, 0), (0, -1), (-1, 0))[rng.integers(0, 4)]
        yy, xx = y + dy, x + dx
        if 0 <= yy < h and 0 <= xx < w and not m[yy, xx]:
            m[yy, xx] = True
            cells.append((yy, xx))
    return m


def _check(P, mask, iid=7):
    ys, xs = np.nonzero(mask)
    rmin, rmax, cmin, cmax = ys.min(), ys.max() + 1, xs.min(), xs.max() + 1
    inst = np.where(mask, iid, 0).astype(np.int32)
    inst[~mask] = np.where(np.random.default_rng(0).random(mask.shape) < 0.3, 3, 0)[~mask]  # other ids around it
    got = P.contour(inst, iid, rmin, cmin, rmax, cmax)
    ref = _ref(mask[rmin:rmax, cmin:cmax]) + np.array([cmin, rmin], np.int32)
    assert np.array_equal(got, ref), (got.tolist(), ref.tolist())


def test_known_shapes(oracle_pp):
    P = oracle_pp
    shapes = []
    m = np.zeros((9, 9), bool); m[4, 4] = True; shapes.append(m)                       # single pixel
    m = np.zeros((9, 9), bool); m[4, 2:7] = True; shapes.append(m)                     # horizontal line
    m = np.zeros((9, 9), bool); m[1:8, 3] = True; shapes.append(m)                     # vertical line
    m = np.zeros((9, 9), bool); m[2:7, 2:7] = True; shapes.append(m)                   # square
    m = np.zeros((9, 9), bool); m[2:7, 2:7] = True; m[4, 4] = False; shapes.append(m)  # square with a hole
    m = np.zeros((9, 9), bool); m[1, 3] = True; m[2, 2:5] = True; m[3, 1:6] = True; m[4, 2:5] = True; m[5, 3] = True
    shapes.append(m)                                                                   # diamond
    m = np.zeros((9, 9), bool); m[2:7, 2:7] = True; m[2, 2] = False; shapes.append(m)  # notch at the start corner
    m = np.zeros((9, 9), bool); m[3, 3:5] = True; shapes.append(m)                     # two pixels
    m = np.zeros((9, 9), bool); m[2, 2:6] = True; m[3, 2] = True; m[4, 2:6] = True; shapes.append(m)  # C shape
    for m in shapes:
        _check(P, m)


def test_random_connected_masks(oracle_pp):
    rng = np.random.default_rng(42)
    for _ in range(400):
        h, w = int(rng.integers(3, 30)), int(rng.integers(3, 30))
        n = int(rng.integers(1, max(2, h * w // 2)))
        _check(oracle_pp, _grow(rng, h, w, n))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_every_instance_of_a_synthetic_map(oracle_pp, seed):
    from hover_net_b200 import synth
    pm = synth.synth_pred_map(164, 164, 6, seed)
    inst, table = oracle_pp.process_table(pm, 6)
    assert len(table) > 10
    for r in table:
        iid, rmin, cmin, rmax, cmax = (int(v) for v in r[:5])
        got = oracle_pp.contour(inst, iid, rmin, cmin, rmax, cmax)
        ref = _ref(inst[rmin:rmax, cmin:cmax] == iid) + np.array([cmin, rmin], np.int32)
        assert np.array_equal(got, ref)
