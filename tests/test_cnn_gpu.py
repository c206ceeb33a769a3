"""GPU parity of the CNN path: libhvn `infer_step` vs the reference goldens and the fp64 oracle.
Tolerance (BASELINE.json north_star): 1e-4 absolute on the NP / HV float maps."""
import os

import numpy as np
import pytest
import torch

from hover_net_b200 import arch, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-4


def _model(mode, nt, seed=0, conv_path=0):
    from hover_net_b200.models.hovernet.net_desc import create_model
    net = create_model(mode=mode, input_ch=3, nr_types=nt, freeze=False)
    net.load_state_dict(synth.make_state_dict(mode, nt, seed=seed), strict=True)
    net.ctx.set_option("conv_path", conv_path)
    return net


@pytest.mark.parametrize("mode,nt", [("original", None), ("original", 5), ("fast", 6)])
@pytest.mark.parametrize("conv_path", [1, 0])
def test_infer_step_matches_reference_golden(mode, nt, conv_path):
    from hover_net_b200.models.hovernet.run_desc import infer_step
    g = np.load(os.path.join(GOLD, "cnn_%s_%s.npz" % (mode, nt)))
    x = synth.make_patches(int(g["batch"]), arch.PATCH_GEOMETRY[mode][0], seed=int(g["patch_seed"]))
    net = _model(mode, nt, int(g["ckpt_seed"]), conv_path)
    out = infer_step(torch.from_numpy(x), net)
    assert out.shape == g["out"].shape and out.dtype == np.float32
    err = np.abs(out[..., -3:] - g["out"][..., -3:]).max()
    assert err <= TOL, "max abs err %.3e" % err
    if nt is not None:
        assert (out[..., 0] != g["out"][..., 0]).mean() < 2e-3
    net.ctx.close()


def test_fp64_referee_and_batch_chunking():
    from oracle import hovernet_torch as O
    from hover_net_b200.models.hovernet.run_desc import infer_step
    mode, nt = "fast", 6
    x = synth.make_patches(3, 256, seed=21)
    net = _model(mode, nt, 3)
    net.ctx.set_option("chunk", 2)  # 3 patches in chunks of 2 + 1
    out = infer_step(x, net)
    sd = O.to_torch_state_dict(synth.make_state_dict(mode, nt, seed=3))
    ref = O.infer_step(x[:1], sd, mode, nt, dtype=torch.float64)
    assert np.abs(out[:1, ..., 1:] - ref[..., 1:]).max() <= TOL
    net.ctx.set_option("chunk", 3)
    out2 = infer_step(x, net)
    assert np.array_equal(out, out2), "results must not depend on the sub-batch split"
    net.ctx.close()


def test_strict_checkpoint_errors():
    from hover_net_b200.models.hovernet.net_desc import create_model
    net = create_model(mode="fast", nr_types=6)
    sd = synth.make_state_dict("fast", 6, seed=0)
    bad = dict(sd)
    bad.pop("d1.units.2.conv2.weight")
    with pytest.raises(RuntimeError):
        net.load_state_dict(bad, strict=True)
    bad = dict(sd)
    bad["bogus.weight"] = np.zeros((1,), np.float32)
    with pytest.raises(RuntimeError):
        net.load_state_dict(bad, strict=True)
    with pytest.raises(Exception):
        net.ctx.forward(np.zeros((1, 256, 256, 3), np.uint8))  # weights never finalised
    sd2 = {"module." + k: v for k, v in sd.items()}  # DataParallel-saved checkpoint
    net.load_state_dict(sd2, strict=True)
    with pytest.raises(Exception):
        net.ctx.forward(np.zeros((1, 270, 270, 3), np.uint8))  # 270 is not a legal `fast` input
    net.ctx.close()


def test_fused_forward_postproc_equals_two_step():
    from hover_net_b200.models.hovernet.run_desc import infer_step, infer_step_fused
    from hover_net_b200.models.hovernet.post_proc import process
    net = _model("fast", 6, 0)
    x = synth.make_patches(2, 256, seed=5)
    pred = infer_step(x, net)
    fpred, finst, ftab, fn = infer_step_fused(x, net)
    assert np.array_equal(pred, fpred)
    for i in range(2):
        inst, _ = process(pred[i], nr_types=6, return_centroids=True)
        assert np.array_equal(inst, finst[i])
    net.ctx.close()
