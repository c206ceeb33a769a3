"""GPU parity of the post-processing path: libhvn (through the C ABI) vs the CPU oracle, bit-exact."""
import glob
import os

import numpy as np
import pytest

from hover_net_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def ctx():
    from hover_net_b200 import _lib
    c = _lib.Context(0)
    yield c
    c.close()


def _check(ctx, P, maps, nr_types):
    maps = np.stack(maps)
    inst, table, nrows = ctx.postproc(maps, nr_types)
    for i in range(maps.shape[0]):
        oi, ot = P.process_table(maps[i], nr_types)
        assert np.array_equal(inst[i], oi), "inst_map differs on map %d (%d px)" % (i, int((inst[i] != oi).sum()))
        assert int(nrows[i]) == ot.shape[0]
        assert np.array_equal(table[i, : ot.shape[0]], ot)


@pytest.mark.parametrize("hw,nt", [((80, 80), None), ((80, 80), 5), ((164, 164), 6), ((270, 270), None),
                                   ((97, 133), 5), ((33, 47), None)])
def test_synthetic_nuclei_batches(ctx, oracle_pp, hw, nt):
    _check(ctx, oracle_pp, [synth.synth_pred_map(hw[0], hw[1], nt, s) for s in range(6)], nt)


def test_noise_and_degenerate_maps(ctx, oracle_pp):
    rng = np.random.default_rng(11)
    maps = [np.zeros((64, 64, 3), np.float32), np.ones((64, 64, 3), np.float32)]
    for _ in range(6):  # salt-and-pepper foreground, noisy HV: many tiny components, holes, ties in np
        m = rng.uniform(-1, 1, (64, 64, 3)).astype(np.float32)
        m[..., 0] = rng.uniform(0, 1, (64, 64)) ** rng.uniform(0.3, 2.0)
        maps.append(m)
    _check(ctx, oracle_pp, maps, None)


def test_smooth_random_fields(ctx, oracle_pp):
    import cv2
    rng = np.random.default_rng(5)
    maps = []
    for _ in range(4):  # CNN-like smooth fields: large irregular blobs with several markers each
        m = np.stack([cv2.GaussianBlur(rng.standard_normal((164, 164)), (0, 0), s) for s in (6, 4, 4)], -1)
        m = m / np.abs(m).max((0, 1))
        m[..., 0] = 0.5 + 0.5 * m[..., 0]
        maps.append(m.astype(np.float32))
    _check(ctx, oracle_pp, maps, None)


def test_large_tile(ctx, oracle_pp):
    _check(ctx, oracle_pp, [synth.synth_pred_map(1024, 1024, 6, 3)], 6)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "pp_*.npz"))))
def test_process_matches_reference_golden(path):
    from hover_net_b200.models.hovernet.post_proc import process
    g = np.load(path)
    nt = None if int(g["nr_types"]) < 0 else int(g["nr_types"])
    pm = synth.synth_pred_map(int(g["h"]), int(g["w"]), nt, int(g["seed"]))
    inst, info = process(pm, nr_types=nt, return_centroids=True)
    assert inst.dtype == np.int32 and np.array_equal(inst, g["inst"])
    ids = np.array(sorted(info.keys()), dtype=np.int32)
    assert np.array_equal(ids, g["ids"])
    for j, i in enumerate(ids):
        assert np.array_equal(info[i]["bbox"], g["bbox"][j])
        assert np.array_equal(info[i]["centroid"], g["centroid"][j])
        assert len(info[i]["contour"]) == g["contour_len"][j]
        if nt is not None:
            assert info[i]["type"] == g["type"][j] and info[i]["type_prob"] == g["type_prob"][j]


def test_idempotent_and_batch_invariant(ctx):
    maps = np.stack([synth.synth_pred_map(164, 164, 6, s) for s in range(8)])
    a = ctx.postproc(maps, 6)
    b = ctx.postproc(maps, 6)
    one = ctx.postproc(maps[3], 6)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert np.array_equal(a[0][3], one[0][0]) and int(a[2][3]) == int(one[2][0])


def test_three_flood_implementations_agree(ctx, oracle_pp):
    """calendar queue (default) == shared-memory heap == global-memory heap == oracle."""
    import cv2
    rng = np.random.default_rng(9)
    maps = [synth.synth_pred_map(164, 164, None, s, density=1.0 / 350.0) for s in range(3)]  # crowded: touching nuclei
    for _ in range(3):
        m = np.stack([cv2.GaussianBlur(rng.standard_normal((164, 164)), (0, 0), s) for s in (5, 3, 3)], -1)
        m = m / np.abs(m).max((0, 1))
        m[..., 0] = 0.55 + 0.5 * m[..., 0]
        maps.append(m.astype(np.float32))
    try:
        for impl in (0, 1, 2):
            ctx.set_option("flood_impl", impl)
            _check(ctx, oracle_pp, maps, None)
    finally:
        ctx.set_option("flood_impl", 0)


def _cv2_contour(inst, row):
    import cv2
    iid, rmin, cmin, rmax, cmax = (int(v) for v in row[:5])
    c = cv2.findContours((inst[rmin:rmax, cmin:cmax] == iid).astype(np.uint8), cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
    return c[0][0].reshape(-1, 2).astype(np.int32) + np.array([cmin, rmin], np.int32)


@pytest.mark.parametrize("hw,nt,n", [((164, 164), 6, 6), ((80, 80), None, 4), ((97, 133), 5, 3), ((512, 512), None, 1)])
def test_device_contours_equal_cv2_findcontours(ctx, hw, nt, n):
    """Row f3: every instance's device-traced contour == the reference's cv2.findContours call
    (post_proc.py:133-137), point for point, on nuclei maps and on irregular smooth-field blobs."""
    import cv2
    rng = np.random.default_rng(21)
    maps = [synth.synth_pred_map(hw[0], hw[1], nt, s) for s in range(n)]
    if nt is None:  # irregular blobs with holes and thin necks
        m = np.stack([cv2.GaussianBlur(rng.standard_normal(hw), (0, 0), s) for s in (5, 3, 3)], -1)
        m = m / np.abs(m).max((0, 1))
        m[..., 0] = 0.55 + 0.5 * m[..., 0]
        maps.append(m.astype(np.float32))
    maps = np.stack(maps)
    inst, table, nrows, offs, pts = ctx.postproc_contours(maps, nt)
    i2, t2, n2 = ctx.postproc(maps, nt)
    assert np.array_equal(inst, i2) and np.array_equal(table, t2) and np.array_equal(nrows, n2)
    max_rows = table.shape[1]
    assert offs[0] == 0 and offs[-1] == len(pts) and np.all(np.diff(offs) >= 0)
    total = 0
    for m in range(maps.shape[0]):
        for r in range(int(nrows[m])):
            k = m * max_rows + r
            got = pts[offs[k] : offs[k + 1]]
            assert np.array_equal(got, _cv2_contour(inst[m], table[m, r])), (m, r)
            total += 1
        assert offs[m * max_rows + int(nrows[m])] == offs[(m + 1) * max_rows]  # rows past n_rows are empty
    assert total > 10


def test_device_contours_small_buffer_reports_capacity(ctx):
    from hover_net_b200 import _lib
    maps = np.stack([synth.synth_pred_map(164, 164, 6, 0)])
    inst, table, nrows, offs, pts = ctx.postproc_contours(maps, 6, pts_cap=8)  # retried internally with the exact size
    assert len(pts) == offs[-1] > 8
    full = ctx.postproc_contours(maps, 6)
    assert np.array_equal(pts, full[4]) and np.array_equal(offs, full[3])
    _ = _lib


def test_process_contours_match_oracle_process(oracle_pp):
    """`process` (device contours + table centroids) == the oracle's `process` (cv2.findContours / cv2.moments)."""
    from hover_net_b200.models.hovernet.post_proc import process
    for nt, seed in ((6, 3), (None, 4)):
        pm = synth.synth_pred_map(164, 164, nt, seed)
        inst, info = process(pm, nr_types=nt, return_centroids=True)
        oi, oinfo = oracle_pp.process(pm, nr_types=nt, return_centroids=True)
        assert np.array_equal(inst, oi) and sorted(info.keys()) == sorted(oinfo.keys())
        for k in info:
            assert np.array_equal(info[k]["contour"], oinfo[k]["contour"]) and info[k]["contour"].dtype == np.int32
            assert np.array_equal(info[k]["centroid"], oinfo[k]["centroid"])
            assert np.array_equal(info[k]["bbox"], oinfo[k]["bbox"])
            assert info[k]["type"] == oinfo[k]["type"] and info[k]["type_prob"] == oinfo[k]["type_prob"]


def test_pack_tables_kernel_matches_padded_tables():
    """hvn_pack_tables_dev (the send buffer of the end-of-batch gather): packed rows == the padded tables' valid rows in
    map order, offs == exclusive prefix of n_rows; a too-small capacity drops rows but keeps offs exact."""
    import torch
    from hover_net_b200 import _lib
    from hover_net_b200.dist import PackedGather
    nt, n = 6, 5
    maps = np.stack([synth.synth_pred_map(96, 112, nt, 40 + s) for s in range(n)])
    maps[3, ..., 1] = 0.0                                    # a map without instances
    c = _lib.Context(0)
    inst, table, nrows = c.postproc(maps, nt)
    max_rows = table.shape[1]
    d_tab = torch.from_numpy(table).cuda()
    d_nr = torch.from_numpy(nrows).cuda()
    pg = PackedGather(c, n, max_rows, int(nrows.sum()) + 3, 1, torch.device("cuda", 0))
    k = pg.launch(d_tab, d_nr)
    c.sync()
    (offs, packed), = pg.rows(k)
    assert np.array_equal(offs, np.concatenate([[0], np.cumsum(nrows)]))
    assert np.array_equal(packed, np.concatenate([table[i, : nrows[i]] for i in range(n)]))
    small = PackedGather(c, n, max_rows, int(nrows[0]) + 1, 1, torch.device("cuda", 0))
    k = small.launch(d_tab, d_nr)
    c.sync()
    (offs2, packed2), = small.rows(k)
    assert np.array_equal(offs2, offs) and np.array_equal(packed2, packed[: int(nrows[0]) + 1])
    c.close()
