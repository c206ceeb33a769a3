"""Minimal forward(+post-proc) run for ncu captures: one warm-up pass, one profiled pass
(`python tools/ncu_target.py [B] [mode]`, default B=8 fast; one chunk, branch streams off so launches are serial)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hover_net_b200 import synth
from hover_net_b200.models.hovernet.net_desc import create_model

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
mode = sys.argv[2] if len(sys.argv) > 2 else "fast"
nt = {"fast": 6, "original": 5}[mode]
net = create_model(mode=mode, nr_types=nt)
net.load_state_dict(synth.make_state_dict(mode, nt, 0))
for kv in os.environ.get("HVN_OPTS", "").split(","):  # development knobs: key=value,...
    if "=" in kv:
        k, v = kv.split("=")
        net.ctx.set_option(k, int(v))
net.ctx.set_option("chunk", B)
net.ctx.set_option("branch_streams", 0)
x = np.concatenate([synth.make_patches(8, 256 if mode == "fast" else 270, seed=1)] * ((B + 7) // 8))[:B]
for _ in range(2):
    net.ctx.forward_postproc(x, want_pred=False)
net.ctx.close()
