"""Generates tests/golden/*.npz by running the UNMODIFIED reference from /root/reference (CPU).

Run in the build container only (the GPU box has no /root/reference):  python -m oracle.gen_golden

* cnn_<mode>_<nr_types>.npz : output of reference `infer_step` (models/hovernet/run_desc.py:171-197)
  on seeded patches with the seeded synthetic checkpoint (hover_net_b200.synth).  The reference
  hard-codes `.to("cuda")`; there is no GPU here, so Tensor.to is wrapped to map "cuda" -> "cpu"
  for the duration of the call -- nothing else is touched.  `matplotlib` (unused import at
  models/hovernet/utils.py:7) is stubbed.
* pp_<name>.npz : output of reference `process` (models/hovernet/post_proc.py:94-186) on seeded
  synthetic nuclei maps, with `skimage.segmentation.watershed` -- the only un-installed dependency
  on the path -- provided by oracle/postproc_oracle.c (so that one step stays UNPINNED).
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")

from hover_net_b200 import arch, synth  # noqa: E402
from oracle import postproc_oracle as P  # noqa: E402

CNN_CASES = (("original", None, 2), ("original", 5, 1), ("fast", 6, 1))
PP_CASES = (("p80_seg", 80, 80, None, 0), ("p80_typed", 80, 80, 5, 1), ("p164_typed", 164, 164, 6, 2),
            ("t270_seg", 270, 270, None, 3), ("t270_typed", 270, 270, 6, 4), ("r97x133_typed", 97, 133, 5, 5))


def _stub_imports():
    mpl = types.ModuleType("matplotlib")
    mpl.cm = None
    mpl.pyplot = types.ModuleType("matplotlib.pyplot")  # imported (unused on this path) at run_desc.py:2
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", mpl.pyplot)
    try:  # the real function, when scikit-image is installed (oracle/regen_with_skimage.py) -- no stub then
        import skimage.segmentation  # noqa: F401
    except ImportError:
        sk = types.ModuleType("skimage")
        seg = types.ModuleType("skimage.segmentation")

        def watershed(image, markers=None, mask=None):
            return P.watershed(image, markers, mask)

        seg.watershed = watershed
        sk.segmentation = seg
        sys.modules["skimage"] = sk
        sys.modules["skimage.segmentation"] = seg
    sys.path.insert(0, REF)


def gen_cnn():
    from models.hovernet.net_desc import create_model
    from models.hovernet.run_desc import infer_step

    orig_to = torch.Tensor.to

    def to_cpu(self, *a, **k):
        a = tuple("cpu" if (isinstance(x, str) and x == "cuda") else x for x in a)
        return orig_to(self, *a, **k)

    for mode, nt, batch in CNN_CASES:
        net = create_model(mode=mode, input_ch=3, nr_types=nt, freeze=False)
        sd = {k: torch.from_numpy(v) for k, v in synth.make_state_dict(mode, nt, seed=0).items()}
        net.load_state_dict(sd, strict=True)
        x = synth.make_patches(batch, arch.PATCH_GEOMETRY[mode][0], seed=7)
        torch.Tensor.to = to_cpu
        try:
            out = infer_step(torch.from_numpy(x), net)
        finally:
            torch.Tensor.to = orig_to
        name = "cnn_%s_%s.npz" % (mode, nt)
        np.savez_compressed(os.path.join(OUT, name), out=out.astype(np.float32),
                            in_sum=np.int64(x.astype(np.int64).sum()), batch=batch, patch_seed=7, ckpt_seed=0)
        print(name, out.shape, float(np.abs(out).max()))


def gen_pp():
    from models.hovernet.post_proc import process

    for name, h, w, nt, seed in PP_CASES:
        pm = synth.synth_pred_map(h, w, nt, seed)
        inst, info = process(pm, nr_types=nt, return_centroids=True)
        ids = np.array(sorted(info.keys()), dtype=np.int32)
        np.savez_compressed(
            os.path.join(OUT, "pp_%s.npz" % name), inst=inst.astype(np.int32), ids=ids,
            bbox=np.array([info[i]["bbox"] for i in ids], dtype=np.int64).reshape(-1, 2, 2),
            centroid=np.array([info[i]["centroid"] for i in ids], dtype=np.float64).reshape(-1, 2),
            type=np.array([-1 if info[i]["type"] is None else info[i]["type"] for i in ids], dtype=np.int32),
            type_prob=np.array([-1.0 if info[i]["type_prob"] is None else info[i]["type_prob"] for i in ids]),
            contour_len=np.array([len(info[i]["contour"]) for i in ids], dtype=np.int32),
            contour_sum=np.array([info[i]["contour"].sum(0) for i in ids], dtype=np.int64).reshape(-1, 2),
            pm_sum=np.float64(pm.astype(np.float64).sum()), h=h, w=w, nr_types=-1 if nt is None else nt, seed=seed)
        print("pp_" + name, inst.shape, len(ids), "instances")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    P.build()
    _stub_imports()
    gen_pp()
    gen_cnn()
