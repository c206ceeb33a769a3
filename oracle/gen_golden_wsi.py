"""Generates tests/golden/wsi_*.npz from the UNMODIFIED reference `infer/wsi.py`, imported from
/root/reference with stubs for the packages that are not installed (docopt, openslide, matplotlib,
skimage -- whose `segmentation.watershed` is provided by oracle/postproc_oracle.c, so that one step
stays UNPINNED exactly as in gen_golden.py).  Build-container only:  python -m oracle.gen_golden_wsi

* wsi_geom.npz   : `_get_patch_top_left_info`, `_get_tile_info`, `_get_chunk_patch_info` (wsi.py:64-221)
                   on several slide shapes incl. BASELINE configs[4] (40000^2, chunk 10000, tile 2048, fast).
* wsi_raw_*.npz  : `__get_raw_prediction` (wsi.py:329-383) with a position-coding fake `run_step`:
                   the assembled slide-sized prediction map.
* wsi_merge_*.npz: `process_single_file` (wsi.py:449-708) phases 1-3 on a synthetic nuclei prediction map
                   (raw prediction replaced by a writer of that map): final instance dict and map.
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")

from hover_net_b200 import synth  # noqa: E402
from oracle import postproc_oracle as P  # noqa: E402

GEOM_CASES = (  # (h, w, chunk, tile, amb, patch_in, patch_out)
    (700, 900, 600, 256, 32, 256, 164),
    (1000, 777, 500, 300, 40, 270, 80),
    (5000, 4100, 2000, 2048, 128, 256, 164),
    (40000, 40000, 10000, 2048, 128, 256, 164),
)
MERGE_CASES = (  # (name, h, w, tile, amb, nr_types, seed, mask)
    ("typed", 700, 900, 256, 32, 6, 21, "full"),
    ("seg", 600, 520, 200, 24, None, 22, "partial"),
)
RAW_CASES = (("fast", 700, 900, 600, 256, 164, 31, "partial"), ("orig", 520, 610, 400, 270, 80, 32, "full"))


def _stubs():
    for name in ("matplotlib", "matplotlib.pyplot", "skimage", "skimage.color", "skimage.segmentation",
                 "skimage.morphology", "imgaug", "imgaug.imgaug", "termcolor", "tensorboardX", "docopt", "openslide"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["imgaug"].imgaug = sys.modules["imgaug.imgaug"]
    sys.modules["termcolor"].colored = lambda s, *a, **k: s
    sys.modules["docopt"].docopt = lambda *a, **k: {}
    mpl, plt = sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"]
    mpl.cm = None
    mpl.pyplot = plt
    plt.get_cmap = lambda *a, **k: None
    sk = sys.modules["skimage"]
    sk.color = sys.modules["skimage.color"]
    sk.img_as_ubyte = lambda x: x
    sk.segmentation = sys.modules["skimage.segmentation"]
    sk.segmentation.watershed = lambda image, markers=None, mask=None: P.watershed(image, markers, mask)
    sys.path.insert(0, REF)


class FakeHandler(object):
    """FileHandler protocol (misc/wsi_handler.py:12-92) over an in-memory RGB array."""

    def __init__(self, arr):
        self.arr = arr

    def get_dimensions(self, mag):
        return np.array([self.arr.shape[1], self.arr.shape[0]]).astype(np.int32)

    def prepare_reading(self, read_mag=None, cache_path=None):
        pass

    def read_region(self, coords, size):
        return np.array(self.arr[coords[1] : coords[1] + size[1], coords[0] : coords[0] + size[0]])[..., :3]

    def get_full_img(self, read_mag=None):
        return self.arr


def make_mask(kind, h, w, seed):
    m = np.ones((h // 8, w // 8), np.uint8)
    if kind == "partial":
        rng = np.random.default_rng(seed)
        m[:] = 0
        for _ in range(3):
            y, x = rng.integers(0, m.shape[0] - 8), rng.integers(0, m.shape[1] - 8)
            m[y : y + rng.integers(8, 40), x : x + rng.integers(8, 40)] = 1
    return m


def fake_run_step(batch):
    """Position-coding stand-in for infer_step: the centre region of the patch itself."""
    x = np.asarray(batch.numpy() if hasattr(batch, "numpy") else batch)
    b, h, w, _ = x.shape
    o = 164 if h == 256 else 80
    m = (h - o) // 2
    c = x[:, m : m + o, m : m + o, :].astype(np.float32)
    return np.concatenate([c, c.sum(-1, keepdims=True)], axis=-1)


def _manager(wsi, tmp, h, w, tile, amb, chunk, pin, pout, nt, mask):
    mgr = wsi.InferManager.__new__(wsi.InferManager)
    mgr.method = {"model_args": {"nr_types": nt, "mode": "fast"}}
    mgr.nr_types = nt
    mgr.cache_path = tmp
    mgr.ambiguous_size = amb
    mgr.tile_shape = [tile, tile]
    mgr.chunk_shape = [chunk, chunk]
    mgr.patch_input_shape = [pin, pin]
    mgr.patch_output_shape = [pout, pout]
    mgr.proc_mag = 40
    mgr.save_mask = mgr.save_thumb = False
    mgr.nr_post_proc_workers = 0
    mgr.nr_inference_workers = 0
    mgr.batch_size = 8
    mgr.wsi_mask = mask
    mgr.wsi_proc_shape = np.array([h, w])
    return mgr


def gen_geom(wsi):
    out = {}
    for i, (h, w, chunk, tile, amb, pin, pout) in enumerate(GEOM_CASES):
        shp = np.array([h, w])
        itl, otl = wsi._get_patch_top_left_info(shp, np.array([pin, pin]), np.array([pout, pout]))
        g, b, c = wsi._get_tile_info(shp, np.array([tile, tile]).astype(np.int64), amb)
        ci, pi = wsi._get_chunk_patch_info(shp, np.array([chunk, chunk]), np.array([pin, pin]), np.array([pout, pout]))
        out.update({"c%d_args" % i: np.array([h, w, chunk, tile, amb, pin, pout]), "c%d_in_tl" % i: itl.astype(np.int32),
                    "c%d_out_tl" % i: otl.astype(np.int32), "c%d_grid" % i: g.astype(np.int64),
                    "c%d_boundary" % i: b.astype(np.int64), "c%d_cross" % i: c.astype(np.int64),
                    "c%d_chunk" % i: ci.astype(np.int64), "c%d_patch" % i: pi.astype(np.int32)})
        print("geom", (h, w), "patches", pi.shape[0], "chunks", ci.shape[0], "tiles", g.shape[0], b.shape[0], c.shape[0])
    np.savez_compressed(os.path.join(OUT, "wsi_geom.npz"), n=len(GEOM_CASES), **out)


def gen_raw(wsi, tmp):
    import cv2
    for name, h, w, chunk, pin, pout, seed, mk in RAW_CASES:
        rng = np.random.default_rng(seed)
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        mask = make_mask(mk, h, w, seed)
        mgr = _manager(wsi, tmp, h, w, 256, 32, chunk, pin, pout, 5, mask)
        mgr.wsi_handler = FakeHandler(img)
        mgr.run_step = fake_run_step
        pm = np.lib.format.open_memmap("%s/pred_map.npy" % tmp, mode="w+", shape=(h, w, 4), dtype=np.float32)
        del pm
        ci, pi = wsi._get_chunk_patch_info(np.array([h, w]), np.array([chunk, chunk]), np.array([pin, pin]), np.array([pout, pout]))
        mgr._InferManager__get_raw_prediction(ci, pi)
        pred = np.array(np.load("%s/pred_map.npy" % tmp, mmap_mode="r"))
        np.savez_compressed(os.path.join(OUT, "wsi_raw_%s.npz" % name), args=np.array([h, w, chunk, pin, pout, seed]),
                            mask=mask, pred=pred.astype(np.uint16 if pred.max() < 65536 else np.float32))
        print("raw", name, pred.shape, "covered px", int((pred[..., 3] > 0).sum()))
        _ = cv2


def gen_merge(wsi, tmp):
    import cv2
    from models.hovernet.post_proc import process

    for name, h, w, tile, amb, nt, seed, mk in MERGE_CASES:
        pm = synth.synth_pred_map(h, w, nt, seed)
        mask = make_mask(mk, h, w, seed)
        cv2.imwrite("%s/mask.png" % tmp, mask * 255)
        mgr = _manager(wsi, tmp, h, w, tile, amb, 600, 256, 164, nt, mask)
        mgr.post_proc_func = process
        wsi.get_file_handler = lambda path, backend: FakeHandler(np.zeros((h, w, 3), np.uint8))

        def writer(self_, ci, pi, _pm=pm):
            ptr = np.load("%s/pred_map.npy" % tmp, mmap_mode="r+")
            ptr[:] = _pm
            ptr.flush()

        mgr._InferManager__get_raw_prediction = types.MethodType(writer, mgr)
        os.makedirs("%s/out" % tmp, exist_ok=True)
        mgr.process_single_file("%s/slide.npy" % tmp, "%s/mask.png" % tmp, "%s/out" % tmp)
        info = mgr.wsi_inst_info
        ids = np.array(sorted(info.keys()), dtype=np.int64)
        inst_map = np.array(mgr.wsi_inst_map)
        np.savez_compressed(
            os.path.join(OUT, "wsi_merge_%s.npz" % name), args=np.array([h, w, tile, amb, -1 if nt is None else nt, seed]),
            mask=mask, ids=ids, inst_map=inst_map.astype(np.int32),
            bbox=np.array([info[i]["bbox"] for i in ids], dtype=np.int64).reshape(-1, 2, 2),
            centroid=np.array([info[i]["centroid"] for i in ids], dtype=np.float64).reshape(-1, 2),
            type=np.array([-1 if info[i]["type"] is None else info[i]["type"] for i in ids], dtype=np.int32),
            type_prob=np.array([-1.0 if info[i]["type_prob"] is None else info[i]["type_prob"] for i in ids]),
            contour_len=np.array([len(info[i]["contour"]) for i in ids], dtype=np.int32),
            contour_sum=np.array([np.asarray(info[i]["contour"]).sum(0) for i in ids], dtype=np.int64).reshape(-1, 2))
        print("merge", name, (h, w), len(ids), "instances; map ids", len(np.unique(inst_map)) - 1)


def main():
    import importlib
    import tempfile
    import warnings
    warnings.simplefilter("ignore")
    _stubs()
    P.build()
    wsi = importlib.import_module("infer.wsi")
    wsi.log_info = lambda *a, **k: None

    class InlinePool(object):
        """The reference hands `_assemble_and_flush` to a 1-process spawn pool (wsi.py:338); a spawned child
        could not import the stubbed modules, so the same calls run inline here."""

        def __init__(self, processes=None):
            pass

        def apply_async(self, func, args=()):
            func(*args)

        def close(self):
            pass

        def join(self):
            pass

    wsi.Pool = InlinePool
    # numpy >= 2 writes np.int64(...) reprs into the .npy header when the shape tuple holds numpy scalars
    # (wsi.py:518-534 passes tuple(np.array)); numpy 1.19, the reference's pin, wrote plain ints
    _open = np.lib.format.open_memmap
    np.lib.format.open_memmap = lambda f, mode="r+", dtype=None, shape=None, **k: _open(
        f, mode=mode, dtype=dtype, shape=None if shape is None else tuple(int(v) for v in shape), **k)
    gen_geom(wsi)
    with tempfile.TemporaryDirectory() as tmp:
        gen_raw(wsi, tmp)
        gen_merge(wsi, tmp)


if __name__ == "__main__":
    main()
