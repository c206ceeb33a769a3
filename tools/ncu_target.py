"""Minimal forward(+post-proc) run for ncu captures: one warm-up pass, one profiled pass (B=8, fast, nr_types=6)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hover_net_b200 import synth
from hover_net_b200.models.hovernet.net_desc import create_model

net = create_model(mode="fast", nr_types=6)
net.load_state_dict(synth.make_state_dict("fast", 6, 0))
net.ctx.set_option("chunk", 8)
net.ctx.set_option("branch_streams", 0)
x = synth.make_patches(8, 256, seed=1)
for _ in range(2):
    net.ctx.forward_postproc(x, want_pred=False)
net.ctx.close()
