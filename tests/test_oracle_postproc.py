"""Pins oracle/postproc_oracle.c: per stage against cv2/scipy, end to end against the goldens that
the reference's own `process` produced (tests/golden/pp_*.npz, made by oracle/gen_golden.py)."""
import glob
import os

import cv2
import numpy as np
import pytest
from scipy import ndimage

import ref_stages
from hover_net_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("seed", range(8))
def test_stage_primitives_bit_exact(oracle_pp, seed):
    P = oracle_pp
    rng = np.random.default_rng(seed)
    H, W = (int(v) for v in rng.integers(12, 140, 2))
    b = (rng.uniform(size=(H, W)) < rng.uniform(0.3, 0.7)).astype(np.int32)
    lab, n = P.label4(b)
    ref, rn = ndimage.label(b)
    assert n == rn and np.array_equal(lab, ref)
    assert np.array_equal(P.remove_small(lab, 10), ref_stages.remove_small(ref, 10))
    f = rng.standard_normal((H, W)).astype(np.float32)
    assert np.array_equal(P.normalize(f), ref_stages.minmax01(f))
    d = rng.standard_normal((H, W)) * 1e6
    assert np.array_equal(P.normalize(d), ref_stages.minmax01(d))
    assert np.array_equal(P.normalize(np.full((H, W), 3.0, np.float32)), np.zeros((H, W), np.float32))
    fn = P.normalize(f)
    for dx in (1, 0):
        assert np.array_equal(P.sobel21(fn, dx), cv2.Sobel(fn, cv2.CV_64F, dx, 1 - dx, ksize=21))
    g = rng.uniform(size=(H, W))
    assert np.array_equal(P.gauss3(g), cv2.GaussianBlur(g, (3, 3), 0))
    assert np.array_equal(P.fill_holes(b), ndimage.binary_fill_holes(b).astype(np.uint8))
    u = (rng.uniform(size=(H, W)) < 0.8).astype(np.uint8)
    k = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (5, 5))
    assert np.array_equal(P.open_ellipse5(u), cv2.morphologyEx(u, cv2.MORPH_OPEN, k))


@pytest.mark.parametrize("shape_seed", [(80, 80, 0), (164, 164, 1), (270, 270, 2), (97, 133, 3), (33, 41, 4)])
def test_pipeline_stages_vs_cv2_scipy(oracle_pp, shape_seed):
    h, w, seed = shape_seed
    pm = synth.synth_pred_map(h, w, None, seed)
    inst, st = oracle_pp.proc_np_hv(pm, True)
    ref = ref_stages.stages(pm, oracle_pp.watershed)
    for k in st:
        assert np.array_equal(st[k], ref[k]), k
    assert np.array_equal(inst, ref["inst"])


def test_degenerate_maps(oracle_pp):
    z = np.zeros((64, 64, 3), np.float32)
    inst = oracle_pp.proc_np_hv(z)
    assert inst.max() == 0
    o = np.ones((64, 64, 3), np.float32)  # all foreground, constant HV -> no marker survives
    ref = ref_stages.stages(o, oracle_pp.watershed)
    assert np.array_equal(oracle_pp.proc_np_hv(o), ref["inst"])


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "pp_*.npz"))))
def test_process_matches_reference_golden(oracle_pp, path):
    g = np.load(path)
    nt = None if int(g["nr_types"]) < 0 else int(g["nr_types"])
    pm = synth.synth_pred_map(int(g["h"]), int(g["w"]), nt, int(g["seed"]))
    assert float(pm.astype(np.float64).sum()) == float(g["pm_sum"]), "synthetic input drifted"
    inst, info = oracle_pp.process(pm, nr_types=nt, return_centroids=True)
    assert inst.dtype == np.int32 and np.array_equal(inst, g["inst"])
    ids = np.array(sorted(info.keys()), dtype=np.int32)
    assert np.array_equal(ids, g["ids"])
    for j, i in enumerate(ids):
        assert np.array_equal(info[i]["bbox"], g["bbox"][j])
        assert np.array_equal(info[i]["centroid"], g["centroid"][j])
        assert len(info[i]["contour"]) == g["contour_len"][j]
        assert np.array_equal(info[i]["contour"].sum(0), g["contour_sum"][j])
        if nt is not None:
            assert info[i]["type"] == g["type"][j]
            assert info[i]["type_prob"] == g["type_prob"][j]
