"""Minimal forward(+post-proc) run for ncu captures: one warm-up pass, one profiled pass
(`python tools/ncu_target.py [B]`, default B=8; fast mode, nr_types=6, one chunk, branch streams off)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hover_net_b200 import synth
from hover_net_b200.models.hovernet.net_desc import create_model

net = create_model(mode="fast", nr_types=6)
net.load_state_dict(synth.make_state_dict("fast", 6, 0))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
if os.environ.get("HVN_TC_HALO"):  # development knob: 0 | 1 | 2
    net.ctx.set_option("tc_halo", int(os.environ["HVN_TC_HALO"]))
net.ctx.set_option("chunk", B)
net.ctx.set_option("branch_streams", 0)
x = np.concatenate([synth.make_patches(8, 256, seed=1)] * ((B + 7) // 8))[:B]
for _ in range(2):
    net.ctx.forward_postproc(x, want_pred=False)
net.ctx.close()
