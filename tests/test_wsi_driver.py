"""Row f2 (WSI driver): patch / chunk / tile geometry, raw-prediction assembly and the three-phase tile
merge pinned against goldens produced by the reference's own `infer/wsi.py` (oracle/gen_golden_wsi.py);
row f1's command line.  CPU tests post-process with the oracle; the `gpu` test with the device path."""
import json
import os
import socket

import numpy as np
import pytest

from hover_net_b200.infer import wsi

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_geometry_matches_reference():
    g = np.load(os.path.join(GOLD, "wsi_geom.npz"))
    for i in range(int(g["n"])):
        h, w, chunk, tile, amb, pin, pout = (int(v) for v in g["c%d_args" % i])
        shp = np.array([h, w])
        itl, otl = wsi._get_patch_top_left_info(shp, np.array([pin, pin]), np.array([pout, pout]))
        assert np.array_equal(itl, g["c%d_in_tl" % i]) and np.array_equal(otl, g["c%d_out_tl" % i])
        grid, bnd, cross = wsi._get_tile_info(shp, np.array([tile, tile]).astype(np.int64), amb)
        assert np.array_equal(grid, g["c%d_grid" % i]), (h, w)
        assert np.array_equal(bnd, g["c%d_boundary" % i]) and np.array_equal(cross, g["c%d_cross" % i])
        ci, pi = wsi._get_chunk_patch_info(shp, np.array([chunk, chunk]), np.array([pin, pin]), np.array([pout, pout]))
        assert np.array_equal(ci, g["c%d_chunk" % i]) and np.array_equal(pi, g["c%d_patch" % i])
    # BASELINE configs[4]: 40000^2 slide, fast mode -> 244 x 244 patches, 20 x 20 grid tiles of 2048
    assert g["c3_patch"].shape[0] == 244 * 244 and g["c3_grid"].shape[0] == 400


def _manager(h, w, tile, amb, chunk, pin, pout, nt, mask):
    mgr = wsi.InferManager.__new__(wsi.InferManager)
    mgr.method = {"model_args": {"nr_types": nt, "mode": "fast"}}
    mgr.nr_types = nt
    mgr.cache_path = "/tmp/hvn_wsi_cache_unused"
    mgr.ambiguous_size = amb
    mgr.tile_shape = [tile, tile]
    mgr.chunk_shape = [chunk, chunk]
    mgr.patch_input_shape = [pin, pin]
    mgr.patch_output_shape = [pout, pout]
    mgr.proc_mag = 40
    mgr.save_mask = mgr.save_thumb = False
    mgr.batch_size = 8
    mgr.wsi_mask = mask
    mgr.wsi_proc_shape = np.array([h, w])
    return mgr


def _fake_run_step(batch):
    x = np.asarray(batch)
    o = 164 if x.shape[1] == 256 else 80
    m = (x.shape[1] - o) // 2
    c = x[:, m : m + o, m : m + o, :].astype(np.float32)
    return np.concatenate([c, c.sum(-1, keepdims=True)], axis=-1)


class _ArrayHandler(wsi.ArrayHandler):
    def __init__(self, arr):
        self.array = arr
        self.metadata = {"available_mag": [40.0], "base_mag": 40.0, "base_shape": np.array([arr.shape[1], arr.shape[0]])}
        self.image_ptr = arr


@pytest.mark.parametrize("name", ["fast", "orig"])
def test_raw_prediction_assembly_matches_reference(name):
    g = np.load(os.path.join(GOLD, "wsi_raw_%s.npz" % name))
    h, w, chunk, pin, pout, seed = (int(v) for v in g["args"])
    img = np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)
    mgr = _manager(h, w, 256, 32, chunk, pin, pout, 5, g["mask"])
    mgr.wsi_handler = _ArrayHandler(img)
    mgr.run_step = _fake_run_step
    mgr.wsi_pred_map = np.zeros((h, w, 4), np.float32)
    ci, pi = wsi._get_chunk_patch_info(np.array([h, w]), np.array([chunk, chunk]), np.array([pin, pin]), np.array([pout, pout]))
    mgr._get_raw_prediction(ci, pi)
    assert np.array_equal(mgr.wsi_pred_map, g["pred"].astype(np.float32))
    assert (mgr.wsi_pred_map[..., 3] > 0).any() and (mgr.wsi_pred_map[..., 3] == 0).any()


def _check_merge(mgr, g):
    info = mgr.wsi_inst_info
    ids = np.array(sorted(info.keys()), dtype=np.int64)
    assert np.array_equal(ids, g["ids"])
    assert np.array_equal(np.asarray(mgr.wsi_inst_map), g["inst_map"])
    for k, i in enumerate(ids):
        v = info[i]
        assert np.array_equal(v["bbox"], g["bbox"][k]) and np.array_equal(v["centroid"], g["centroid"][k])
        assert len(v["contour"]) == g["contour_len"][k] and np.array_equal(np.asarray(v["contour"]).sum(0), g["contour_sum"][k])
        assert (-1 if v["type"] is None else v["type"]) == g["type"][k]
        assert (-1.0 if v["type_prob"] is None else v["type_prob"]) == g["type_prob"][k]


def _run_merge(name, post_proc_func, tmp_path):
    import cv2
    from hover_net_b200 import synth

    g = np.load(os.path.join(GOLD, "wsi_merge_%s.npz" % name))
    h, w, tile, amb, nt, seed = (int(v) for v in g["args"])
    nt = None if nt < 0 else nt
    pm = synth.synth_pred_map(h, w, nt, seed)
    np.save(str(tmp_path / "slide.npy"), np.zeros((h, w, 3), np.uint8))
    cv2.imwrite(str(tmp_path / "mask.png"), g["mask"] * 255)
    mgr = _manager(h, w, tile, amb, 600, 256, 164, nt, None)
    mgr.post_proc_func = post_proc_func
    mgr.cache_path = str(tmp_path / "cache")

    def writer(ci, pi):  # stands in for the network: the synthetic nuclei map
        mgr.wsi_pred_map[:] = pm

    mgr._get_raw_prediction = writer
    mgr.process_single_file(str(tmp_path / "slide.npy"), str(tmp_path / "mask.png"), str(tmp_path))
    return mgr, g


@pytest.mark.parametrize("name", ["typed", "seg"])
def test_three_phase_merge_matches_reference_cpu(name, tmp_path, oracle_pp):
    mgr, g = _run_merge(name, oracle_pp.process, tmp_path)
    _check_merge(mgr, g)
    js = json.load(open(str(tmp_path / "slide.json")))
    assert js["mag"] == 40 and sorted(int(k) for k in js["nuc"]) == [int(i) for i in g["ids"]]
    k0 = int(g["ids"][0])
    assert js["nuc"][str(k0)]["bbox"] == g["bbox"][0].tolist() and js["nuc"][str(k0)]["centroid"] == g["centroid"][0].tolist()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["typed", "seg"])
def test_three_phase_merge_matches_reference_device(name, tmp_path):
    from hover_net_b200.models.hovernet import post_proc
    mgr, g = _run_merge(name, post_proc.process, tmp_path)
    _check_merge(mgr, g)


def _merge_worker(rank, world, port, tmp, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import pathlib
    from oracle import postproc_oracle as P
    mgr, g = _run_merge("typed", P.process, pathlib.Path(tmp))
    if rank == 0:
        info = mgr.wsi_inst_info
        ids = sorted(info.keys())
        out.put((ids, np.asarray(mgr.wsi_inst_map), [info[i]["centroid"].tolist() for i in ids]))
    dist.barrier()
    dist.destroy_process_group()


def test_three_phase_merge_gloo_world2_equals_reference(tmp_path, oracle_pp):
    """configs[4] plumbing: tiles of every phase sharded over 2 ranks, results gathered to rank 0 and
    merged in tile order == the reference's single-process result."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_merge_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    ids, inst_map, cents = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = np.load(os.path.join(GOLD, "wsi_merge_typed.npz"))
    assert np.array_equal(np.array(ids), g["ids"]) and np.array_equal(inst_map, g["inst_map"])
    assert np.array_equal(np.array(cents), g["centroid"])


def test_process_wsi_list_on_npy_slide(tmp_path, oracle_pp):
    """`process_wsi_list` end to end on a `.npy` slide with a fake network: directory handling, the
    auto-generated tissue mask, the JSON file; an existing output is skipped (reference wsi.py:736-742)."""
    from hover_net_b200 import synth
    h, w, nt = 420, 380, None
    pm = synth.synth_pred_map(h, w, nt, 5)
    img = np.full((h, w, 3), 235, np.uint8)
    img[pm[..., 0] > 0.5] = (90, 60, 140)  # dark nuclei on a bright background -> Otsu tissue mask
    os.makedirs(tmp_path / "in")
    np.save(str(tmp_path / "in" / "s1.npy"), img)
    mgr = wsi.InferManager.__new__(wsi.InferManager)
    mgr.method = {"model_args": {"nr_types": nt, "mode": "fast"}}
    mgr.nr_types = nt
    mgr.post_proc_func = oracle_pp.process
    # the network stage is replaced by a writer of the synthetic nuclei map
    mgr._get_raw_prediction = lambda ci, pi: mgr.wsi_pred_map.__setitem__(slice(None), pm)
    args = {"batch_size": 4, "nr_inference_workers": 0, "nr_post_proc_workers": 0, "patch_input_shape": 256,
            "patch_output_shape": 164, "input_dir": str(tmp_path / "in"), "output_dir": str(tmp_path / "out"),
            "input_mask_dir": None, "cache_path": str(tmp_path / "cache"), "proc_mag": 40, "ambiguous_size": 24,
            "chunk_shape": 400, "tile_shape": 200, "save_thumb": False, "save_mask": True}
    mgr.process_wsi_list(args)
    js = json.load(open(str(tmp_path / "out" / "json" / "s1.json")))
    assert js["mag"] == 40 and len(js["nuc"]) > 20
    assert os.path.exists(str(tmp_path / "out" / "mask" / "s1.png"))
    first = next(iter(js["nuc"].values()))
    assert set(first.keys()) == {"bbox", "centroid", "contour", "type_prob", "type"}
    t0 = os.path.getmtime(str(tmp_path / "out" / "json" / "s1.json"))
    for k in ("chunk_shape", "tile_shape", "patch_input_shape", "patch_output_shape"):
        setattr(mgr, k, args[k])
    mgr.process_wsi_list(args)  # output exists -> skipped
    assert os.path.getmtime(str(tmp_path / "out" / "json" / "s1.json")) == t0


def test_cli_builds_reference_run_args():
    from hover_net_b200 import run_infer
    cmd, margs, rargs, gpu = run_infer.parse(
        ["--gpu=0,1", "--nr_types=6", "--type_info_path=type_info.json", "--batch_size=64", "--model_mode=fast",
         "--model_path=ck.tar", "--nr_inference_workers=8", "--nr_post_proc_workers=16", "tile",
         "--input_dir=in/", "--output_dir=out/", "--mem_usage=0.1", "--draw_dot", "--save_qupath"], nr_gpus=2)
    assert cmd == "tile" and gpu == "0,1"
    assert margs == {"method": {"model_args": {"nr_types": 6, "mode": "fast"}, "model_path": "ck.tar"},
                     "type_info_path": "type_info.json"}
    assert rargs == {"batch_size": 128, "nr_inference_workers": 8, "nr_post_proc_workers": 16, "patch_input_shape": 256,
                     "patch_output_shape": 164, "input_dir": "in/", "output_dir": "out/", "mem_usage": 0.1,
                     "draw_dot": True, "save_qupath": True, "save_raw_map": False}
    cmd, margs, rargs, _ = run_infer.parse(["--model_mode", "original", "--model_path", "ck.tar", "wsi", "--input_dir", "a",
                                            "--output_dir", "b", "--tile_shape=1024", "--save_mask"])
    assert cmd == "wsi" and margs["method"]["model_args"] == {"nr_types": None, "mode": "original"}
    assert margs["type_info_path"] is None
    assert rargs["patch_input_shape"] == 270 and rargs["patch_output_shape"] == 80 and rargs["batch_size"] == 32
    assert rargs["tile_shape"] == 1024 and rargs["chunk_shape"] == 10000 and rargs["ambiguous_size"] == 128
    assert rargs["proc_mag"] == 40 and rargs["cache_path"] == "cache" and rargs["save_mask"] and not rargs["save_thumb"]
    with pytest.raises(Exception, match="model path"):
        run_infer.parse(["tile", "--input_dir=a", "--output_dir=b"])
    assert run_infer.parse(["--help"])[0] is None


@pytest.mark.gpu
def test_cli_wsi_end_to_end_on_device(tmp_path, oracle_pp):
    """`run_infer.py ... wsi` on a synthetic `.npy` slide with the real network: sampled patches of the
    slide-sized prediction map agree with the CPU oracle network (<= 1e-4), and the JSON equals the
    three-phase merge re-run with the oracle's `process` on the same prediction map (bit-exact)."""
    import torch
    from hover_net_b200 import run_infer, synth
    from oracle import hovernet_torch as O

    mode, nt = "fast", 6
    h, w = 600, 700
    base = synth.make_patches(12, 256, seed=60)
    img = np.concatenate([np.concatenate(list(base[r * 3:(r + 1) * 3]), 1) for r in range(3)], 0)[:h, :w].copy()
    os.makedirs(tmp_path / "in")
    np.save(str(tmp_path / "in" / "slide.npy"), img)
    sd = synth.make_state_dict(mode, nt, seed=0)
    np.savez(str(tmp_path / "ckpt.npz"), **sd)
    import cv2
    os.makedirs(tmp_path / "msk")
    cv2.imwrite(str(tmp_path / "msk" / "slide.png"), np.full((h // 8, w // 8), 255, np.uint8))
    argv = ["--nr_types=6", "--model_mode=fast", "--model_path=%s" % (tmp_path / "ckpt.npz"), "--batch_size=8", "wsi",
            "--input_dir=%s" % (tmp_path / "in"), "--output_dir=%s" % (tmp_path / "out"), "--cache_path=%s" % (tmp_path / "cache"),
            "--input_mask_dir=%s" % (tmp_path / "msk"), "--chunk_shape=500", "--tile_shape=256", "--ambiguous_size=32"]
    cmd, margs, rargs, _ = run_infer.parse(argv)
    assert cmd == "wsi"
    margs["method"]["model_args"]["device"] = 0
    mgr = wsi.InferManager(**margs)
    mgr.process_wsi_list(rargs)
    js = json.load(open(str(tmp_path / "out" / "slide.json")))
    assert js["mag"] == 40 and len(js["nuc"]) > 0
    pred = np.array(mgr.wsi_pred_map)
    assert pred.shape == (h, w, 4)
    # (1) network: three patches of the grid against the CPU oracle network
    _, pinfo = wsi._get_chunk_patch_info(np.array([h, w]), np.array([500, 500]), np.array([256, 256]), np.array([164, 164]))
    assert pinfo.shape[0] == 16
    tsd = O.to_torch_state_dict(sd)
    torch.set_num_threads(min(32, os.cpu_count() or 1))
    for k in (0, 5, 10):  # (the last grid row / column sticks out of the slide and belongs to no chunk, as in the reference)
        y, x = (int(v) for v in pinfo[k, 0, 0])
        ref = O.infer_step(img[None, y:y + 256, x:x + 256], tsd, mode, nt)[0]
        got = pred[y + 46:y + 46 + 164, x + 46:x + 46 + 164]
        assert np.abs(got[..., 1:] - ref[..., 1:]).max() <= 1e-4
        assert (got[..., 0] != ref[..., 0]).mean() < 2e-3
    assert np.all(pred[:46] == 0) and np.all(pred[:, :46] == 0)  # the border the patch grid never covers stays empty
    # (2) merge: same phases with the oracle's process on the same map
    ref_mgr = wsi.InferManager.__new__(wsi.InferManager)
    ref_mgr.__dict__.update({k: v for k, v in mgr.__dict__.items() if k not in ("net", "run_step", "wsi_inst_info", "wsi_inst_map")})
    ref_mgr.post_proc_func = oracle_pp.process
    ref_mgr._get_raw_prediction = lambda ci, pi: ref_mgr.wsi_pred_map.__setitem__(slice(None), pred)
    os.makedirs(tmp_path / "ref")
    ref_mgr.process_single_file(str(tmp_path / "in" / "slide.npy"), str(tmp_path / "msk" / "slide.png"), str(tmp_path / "ref"))
    assert sorted(ref_mgr.wsi_inst_info.keys()) == sorted(mgr.wsi_inst_info.keys()) == sorted(int(k) for k in js["nuc"])
    assert np.array_equal(ref_mgr.wsi_inst_map, mgr.wsi_inst_map)
    for k, v in ref_mgr.wsi_inst_info.items():
        j = js["nuc"][str(k)]
        assert j["bbox"] == v["bbox"].tolist() and j["centroid"] == v["centroid"].tolist()
        assert j["contour"] == v["contour"].tolist() and j["type"] == v["type"] and j["type_prob"] == v["type_prob"]
    mgr.net.ctx.close()


def _raw_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = np.load(os.path.join(GOLD, "wsi_raw_fast.npz"))
    h, w, chunk, pin, pout, seed = (int(v) for v in g["args"])
    img = np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)
    mgr = _manager(h, w, 256, 32, chunk, pin, pout, 5, g["mask"])
    mgr.wsi_handler = _ArrayHandler(img)
    mgr.run_step = _fake_run_step
    mgr.batch_size = 3
    mgr.wsi_pred_map = np.zeros((h, w, 4), np.float32)
    ci, pi = wsi._get_chunk_patch_info(np.array([h, w]), np.array([chunk, chunk]), np.array([pin, pin]), np.array([pout, pout]))
    mgr._get_raw_prediction(ci, pi)
    ok = bool(np.array_equal(mgr.wsi_pred_map, g["pred"].astype(np.float32)))
    flags = [None] * world
    dist.all_gather_object(flags, ok)
    if rank == 0:
        out.put(flags)
    dist.barrier()
    dist.destroy_process_group()


def test_raw_prediction_sharded_gloo_world2_equals_reference():
    """configs[4] plumbing: the patches of every chunk sharded over 2 ranks and exchanged with one all_gather
    per chunk -> every rank holds the reference's slide-sized prediction map."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_raw_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    flags = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert flags == [True, True]


def test_cli_help_and_version(capsys):
    from hover_net_b200 import run_infer
    assert run_infer.main(["--version"]) == 0
    assert run_infer.VERSION in capsys.readouterr().out
    assert run_infer.main(["--help"]) == 0
    out = capsys.readouterr().out
    assert "--model_path" in out and "--tile_shape" in out and "--save_qupath" in out
