"""Generates tests/golden/tile_*.npz from the UNMODIFIED reference `infer/tile.py`
(`_prepare_patching` :46-94 and the stitching inside `_post_process_patches` :98-143), imported from
/root/reference with stubs for the packages that are not installed (matplotlib, skimage.color).
Build-container only:  python -m oracle.gen_golden_tile"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
CASES = (("fast", 256, 164, 270, 270, 11), ("orig", 270, 80, 270, 270, 12), ("fast_rect", 256, 164, 500, 333, 13),
         ("orig_rect", 270, 80, 161, 402, 14))


def _stubs():
    for name in ("matplotlib", "matplotlib.pyplot", "skimage", "skimage.color", "imgaug", "imgaug.imgaug", "termcolor",
                 "tensorboardX"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["imgaug"].imgaug = sys.modules["imgaug.imgaug"]
    sys.modules["termcolor"].colored = lambda s, *a, **k: s
    mpl, plt = sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"]
    mpl.cm = None
    mpl.pyplot = plt
    plt.get_cmap = lambda *a, **k: None
    sys.modules["skimage"].color = sys.modules["skimage.color"]
    sys.path.insert(0, "/root/reference")


def main():
    _stubs()
    import importlib
    import warnings
    warnings.simplefilter("ignore")
    if not hasattr(np.lib, "pad"):
        np.lib.pad = np.pad  # numpy >= 2 dropped the alias the reference (numpy 1.19) calls at tile.py:76
    tile = importlib.import_module("infer.tile")
    for name, win, msk, h, w, seed in CASES:
        rng = np.random.default_rng(seed)
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        padded, pinfo, corner = tile._prepare_patching(img, win, msk, True)
        # per-patch "network output": a deterministic function of the patch position, 4 channels
        data = []
        for y, x, _, _ in pinfo:
            yy, xx = np.mgrid[0:msk, 0:msk]
            data.append(np.stack([yy + y, xx + x, (yy + y) * 1000 + (xx + x), np.full_like(yy, 7)], -1).astype(np.float32)[None])
        captured = {}

        def fake_post(pred_map, **kw):
            captured["map"] = np.array(pred_map)
            return np.zeros(pred_map.shape[:2], np.int32), {}

        tile.visualize_instances_dict = lambda im, d, **k: im
        items = [[np.concatenate([p, [0]]), d] for p, d in zip(pinfo, data)]
        perm = rng.permutation(len(items))  # arrival order must not matter
        tile._post_process_patches(fake_post, {}, [items[i] for i in perm],
                                   {"src_shape": img.shape, "src_image": img, "name": name}, {})
        np.savez_compressed(os.path.join(OUT, "tile_%s.npz" % name), win=win, msk=msk, h=h, w=w,
                            seed=seed, patch_info=pinfo.astype(np.int32), corner=np.array(corner),
                            padded_shape=np.array(padded.shape), padded_sum=np.int64(padded.astype(np.int64).sum()),
                            padded_rowsum=padded.astype(np.int64).sum((1, 2)), stitched=captured["map"].astype(np.float32))
        print(name, padded.shape, pinfo.shape, captured["map"].shape)
    gen_writers()


def gen_writers():
    """writers_tile.npz: the reference's QuPath TSV writer (convert_format.py:19-50) and typed overlay
    (misc/viz_utils.py:94-125; typed colours are deterministic, per-instance random colours are not) on a
    seeded instance dict."""
    import importlib
    import tempfile
    conv = importlib.import_module("convert_format")
    viz = importlib.import_module("misc.viz_utils")
    rng = np.random.default_rng(99)
    img = rng.integers(0, 256, (120, 150, 3), dtype=np.uint8)
    type_info = {0: ("nolabe", (0, 0, 0)), 1: ("neopla", (255, 0, 0)), 2: ("inflam", (0, 255, 0)), 3: ("connec", (0, 0, 255))}
    info = {}
    for k in range(1, 9):
        cy, cx, r = int(rng.integers(15, 105)), int(rng.integers(15, 135)), int(rng.integers(4, 12))
        ang = np.linspace(0, 2 * np.pi, 12, endpoint=False)
        cnt = np.stack([cx + r * np.cos(ang), cy + r * np.sin(ang)], -1).round().astype(np.int32)
        info[k] = {"bbox": np.array([[cy - r, cx - r], [cy + r + 1, cx + r + 1]]), "centroid": np.array([cx + 0.25, cy + 0.5]),
                   "contour": cnt, "type_prob": 0.5 + 0.05 * k, "type": int(k % 4)}
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "a.tsv")
        conv.to_qupath(path, [v["centroid"] for v in info.values()], [v["type"] for v in info.values()], type_info)
        tsv = open(path).read()
    over = viz.visualize_instances_dict(img, info, draw_dot=True, type_colour=type_info, line_thickness=2)
    np.savez_compressed(os.path.join(OUT, "writers_tile.npz"), tsv=np.array(tsv), overlay=over.astype(np.uint8),
                        ids=np.array(sorted(info.keys())), contour=np.stack([info[k]["contour"] for k in sorted(info)]),
                        centroid=np.stack([info[k]["centroid"] for k in sorted(info)]),
                        type=np.array([info[k]["type"] for k in sorted(info)]))
    print("writers", over.shape, len(tsv), "bytes of tsv")


if __name__ == "__main__":
    main()
