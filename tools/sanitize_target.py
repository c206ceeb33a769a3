"""Small end-to-end run for compute-sanitizer (memcheck): 1 fast-mode patch through CNN + post-processing,
the CUDA-core referee path, and a 300x200 synthetic map through the generic (large-map) flood."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hover_net_b200 import synth
from hover_net_b200.models.hovernet.net_desc import create_model

net = create_model(mode="fast", nr_types=6)
net.load_state_dict(synth.make_state_dict("fast", 6, 0))
x = synth.make_patches(1, 256, seed=1)
pred, inst, tab, n = net.ctx.forward_postproc(x)
print("tc path ok", int(n[0]))
net.ctx.set_option("conv_path", 1)
pred2 = net.ctx.forward(x)
print("referee path ok", float(np.abs(pred2[..., 1:] - pred[..., 1:]).max()))
pm = synth.synth_pred_map(300, 200, 6, 0)
net.ctx.set_option("flood_impl", 2)
gi, gt, gn = net.ctx.postproc(pm, 6)
print("generic flood ok", int(gn[0]))
net.ctx.close()
