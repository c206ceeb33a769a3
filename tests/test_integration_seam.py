"""INTEGRATION.md section A, executed: the reference resolves its model, step function and post-processing by
MODULE PATH + NAME (reference infer/base.py:56-78).  These tests alias `models.hovernet.{net_desc,run_desc,
post_proc}` to this repo's modules in sys.modules and run the reference's `__load_model` logic -- restated line for
line below, because /root/reference does not exist on the GPU box -- against a reference-format checkpoint file
(`torch.save({"desc": state_dict})`, DataParallel `module.` prefix included), then the calls the reference's
drivers make: `run_step(batch)` on the default-collate uint8 tensor and `post_proc_func` shipped to a
ProcessPoolExecutor under the *spawn* start method (infer/tile.py:5,232-234,353-363)."""
import importlib
import inspect
import os
import pickle
import sys
import types

import numpy as np
import pytest

from hover_net_b200 import arch, synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _alias_reference_module_paths():
    """What a maintainer's PYTHONPATH shim does: `models.hovernet.*` resolves to hover_net_b200.models.hovernet.*"""
    import hover_net_b200.models as m
    import hover_net_b200.models.hovernet as mh
    from hover_net_b200.models.hovernet import net_desc, post_proc, run_desc
    sys.modules["models"] = m
    sys.modules["models.hovernet"] = mh
    sys.modules["models.hovernet.net_desc"] = net_desc
    sys.modules["models.hovernet.run_desc"] = run_desc
    sys.modules["models.hovernet.post_proc"] = post_proc


def _convert_pytorch_checkpoint(net_state_dict):
    """reference run_utils/utils.py:15-29: strip `module.` only when EVERY key carries it."""
    names = list(net_state_dict.keys())
    if all(v.split(".")[0] == "module" for v in names):
        net_state_dict = {".".join(k.split(".")[1:]): v for k, v in net_state_dict.items()}
    return net_state_dict


class _RefInferManager(object):
    """reference infer/base.py:22-78 with the three-line patch of INTEGRATION.md A applied: the module roots are
    untouched here (they are aliased instead) and the DataParallel / .to("cuda") lines are dropped."""

    def __init__(self, **kwargs):
        self.run_step = None
        for variable, value in kwargs.items():
            self.__setattr__(variable, value)
        self.__load_model()
        self.nr_types = self.method["model_args"]["nr_types"]

    def __load_model(self):
        import torch
        from importlib import import_module
        model_desc = import_module("models.hovernet.net_desc")
        model_creator = getattr(model_desc, "create_model")
        net = model_creator(**self.method["model_args"])
        saved_state_dict = torch.load(self.method["model_path"])["desc"]
        saved_state_dict = _convert_pytorch_checkpoint(saved_state_dict)
        net.load_state_dict(saved_state_dict, strict=True)
        module_lib = import_module("models.hovernet.run_desc")
        run_step = getattr(module_lib, "infer_step")
        self.run_step = lambda input_batch: run_step(input_batch, net)
        module_lib = import_module("models.hovernet.post_proc")
        self.post_proc_func = getattr(module_lib, "process")
        self.net = net


def _spawn_probe(fn_bytes):
    fn = pickle.loads(fn_bytes)
    return fn.__module__, fn.__name__


def test_seam_names_signatures_and_spawn_pickling():
    """CPU: the three callables resolve under the reference's module paths with the reference's parameter names, and
    `process` survives pickling into a spawn worker (no device is touched by importing it)."""
    _alias_reference_module_paths()
    nd = importlib.import_module("models.hovernet.net_desc")
    rd = importlib.import_module("models.hovernet.run_desc")
    pp = importlib.import_module("models.hovernet.post_proc")
    assert list(inspect.signature(nd.create_model).parameters)[0] == "mode"
    assert list(inspect.signature(rd.infer_step).parameters) == ["batch_data", "model"]
    assert list(inspect.signature(pp.process).parameters) == ["pred_map", "nr_types", "return_centroids"]
    assert inspect.signature(pp.process).parameters["nr_types"].default is None
    assert inspect.signature(pp.process).parameters["return_centroids"].default is False
    import multiprocessing as mp
    with mp.get_context("spawn").Pool(1) as pool:
        mod, name = pool.apply(_spawn_probe, (pickle.dumps(pp.process),))
    assert (mod, name) == ("hover_net_b200.models.hovernet.post_proc", "process")


def test_checkpoint_prefix_rule_matches_reference():
    """`module.` is stripped only when all keys carry it (run_utils/utils.py:15-29); a mixed dict is a strict-load
    error, not a silent partial strip."""
    from hover_net_b200.models.hovernet.net_desc import convert_pytorch_checkpoint
    sd = {"module.a": 1, "module.b": 2}
    assert convert_pytorch_checkpoint(sd) == {"a": 1, "b": 2}
    mixed = {"module.a": 1, "b": 2}
    assert convert_pytorch_checkpoint(mixed) == mixed


@pytest.mark.gpu
def test_reference_load_model_logic_runs_on_the_device(tmp_path):
    import torch
    from concurrent.futures import ProcessPoolExecutor
    import multiprocessing as mp
    from oracle import postproc_oracle as P
    P.build()
    mode, nt = "fast", 6
    g = np.load(os.path.join(GOLD, "cnn_%s_%s.npz" % (mode, nt)))
    sd = synth.make_state_dict(mode, nt, seed=int(g["ckpt_seed"]))
    ckpt = str(tmp_path / "hovernet_fast_synth.tar")
    torch.save({"desc": {"module." + k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}}, ckpt)  # DataParallel-style
    _alias_reference_module_paths()
    mgr = _RefInferManager(method={"model_args": {"nr_types": nt, "mode": mode}, "model_path": ckpt}, type_info_path=None)
    x = synth.make_patches(int(g["batch"]), arch.PATCH_GEOMETRY[mode][0], seed=int(g["patch_seed"]))
    out = mgr.run_step(torch.from_numpy(x))                       # default-collate tensor, uint8 NHWC
    assert isinstance(out, np.ndarray) and out.dtype == np.float32 and out.shape == g["out"].shape
    assert np.abs(out[..., -3:] - g["out"][..., -3:]).max() <= 1e-4
    # the reference ships post_proc_func to pool workers under spawn; each worker opens its own device context
    pm = synth.synth_pred_map(164, 164, nt, 2)
    with ProcessPoolExecutor(1, mp_context=mp.get_context("spawn")) as pool:
        inst, info = pool.submit(mgr.post_proc_func, pm, nr_types=nt, return_centroids=True).result(timeout=300)
    oi, ot = P.process_table(pm, nt)
    assert inst.dtype == np.int32 and np.array_equal(inst, oi)
    assert set(info) <= set(int(i) for i in ot[:, 0]) and len(info) > 0
    k = next(iter(info))
    assert set(info[k]) == {"bbox", "centroid", "contour", "type_prob", "type"}
    # strict load errors like torch's: an unexpected key under the all-keys prefix rule
    bad = {"desc": {**{k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, "module.extra": torch.zeros(1)}}
    torch.save(bad, ckpt)
    with pytest.raises(RuntimeError):
        _RefInferManager(method={"model_args": {"nr_types": nt, "mode": mode}, "model_path": ckpt}, type_info_path=None)
    mgr.net.ctx.close()
