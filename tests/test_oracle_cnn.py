"""Pins oracle/hovernet_torch.py against tests/golden/cnn_*.npz (output of the unmodified reference
`infer_step` on CPU, see oracle/gen_golden.py) and checks the checkpoint key inventory."""
import os

import numpy as np
import pytest
import torch

from hover_net_b200 import arch, synth
from oracle import hovernet_torch as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("mode,nt", [("original", None), ("original", 5), ("fast", 6)])
def test_oracle_cnn_matches_reference_golden(mode, nt):
    g = np.load(os.path.join(GOLD, "cnn_%s_%s.npz" % (mode, nt)))
    x = synth.make_patches(int(g["batch"]), arch.PATCH_GEOMETRY[mode][0], seed=int(g["patch_seed"]))
    assert int(x.astype(np.int64).sum()) == int(g["in_sum"]), "synthetic patches drifted"
    sd = O.to_torch_state_dict(synth.make_state_dict(mode, nt, seed=int(g["ckpt_seed"])))
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    out = O.infer_step(x, sd, mode, nt)
    assert out.shape == g["out"].shape and out.dtype == np.float32
    # same graph, same library: only thread-count dependent reduction order may differ
    assert np.abs(out[..., -3:] - g["out"][..., -3:]).max() <= 1e-5
    if nt is not None:
        assert (out[..., 0] != g["out"][..., 0]).mean() < 1e-3


def test_state_dict_spec_counts():
    assert len(arch.state_dict_spec("fast", 6)) == 798
    s = arch.state_dict_spec("original", None)
    assert s["decoder.np.u3.conva.weight"] == (256, 1024, 5, 5)
    assert s["decoder.hv.u2.dense.units.3.conv1.weight"] == (128, 224, 1, 1)
    assert "decoder.tp.u0.conv.bias" not in s
    assert arch.out_size("original", 270) == 80 and arch.out_size("fast", 256) == 164
