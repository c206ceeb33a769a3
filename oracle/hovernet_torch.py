"""ORACLE (test infrastructure, not product): CPU restatement of the reference CNN path.

A functional PyTorch fp32/fp64 restatement of `infer_step` -> `HoVerNet.forward`
(reference `models/hovernet/run_desc.py:171-197`, `models/hovernet/net_desc.py:101-145`,
`models/hovernet/net_utils.py:39-294`, `models/hovernet/utils.py:11-50`) written over a flat
checkpoint dict instead of the reference's module tree.  It is pinned against the
*unmodified* reference imported from /root/reference by `oracle/gen_golden.py`
(fixtures in `tests/golden/cnn_*.npz`; see tests/test_oracle_cnn.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this file.  The product (hover_net_b200/) never does.
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

_GROUPS = (("d0", 3, 1), ("d1", 4, 2), ("d2", 6, 2), ("d3", 3, 2))  # net_desc.py:36-39


def _bn_relu(x, sd, prefix):
    # eval-mode BatchNorm2d(eps=1e-5) + ReLU  (net_utils.py:98-99,176-177)
    y = F.batch_norm(
        x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"],
        sd[prefix + ".weight"], sd[prefix + ".bias"], training=False, eps=1e-5)
    return F.relu(y)


def _tf_same_pad(x, ksize, stride):
    # net_utils.py:51-67 : uses H for both dims; odd pad -> extra at bottom/right
    if x.shape[2] % stride == 0:
        pad = max(ksize - stride, 0)
    else:
        pad = max(ksize - (x.shape[2] % stride), 0)
    lo = pad // 2
    hi = pad - lo
    return F.pad(x, (lo, hi, lo, hi), "constant", 0)


def _crop(x, ch, cw):
    # utils.py:11-28 (crop_op): crop_t = c//2, crop_b = c - crop_t
    t = ch // 2
    l = cw // 2
    return x[:, :, t:x.shape[2] - (ch - t), l:x.shape[3] - (cw - l)]


def _up2(x):
    # net_utils.py:284-294 : nearest-neighbour x2 (tensordot with ones(2,2))
    return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)


def _residual_group(x, sd, name, units, stride):
    # net_utils.py:250-266
    shortcut = F.conv2d(x, sd[name + ".shortcut.weight"], stride=stride)
    prev = x
    for u in range(units):
        p = "%s.units.%d." % (name, u)
        f = prev
        if u != 0:
            f = _bn_relu(f, sd, p + "preact/bn")
        f = F.conv2d(f, sd[p + "conv1.weight"])
        f = _bn_relu(f, sd, p + "conv1/bn")
        s = stride if u == 0 else 1
        f = _tf_same_pad(f, 3, s)
        f = F.conv2d(f, sd[p + "conv2.weight"], stride=s)
        f = _bn_relu(f, sd, p + "conv2/bn")
        f = F.conv2d(f, sd[p + "conv3.weight"])
        prev = f + shortcut
        shortcut = prev
    return _bn_relu(prev, sd, name + ".blk_bna.bn")


def _dense(x, sd, prefix, units):
    # net_utils.py:144-151
    for u in range(units):
        q = prefix + "units.%d." % u
        f = _bn_relu(x, sd, q + "preact_bna/bn")
        f = F.conv2d(f, sd[q + "conv1.weight"])
        f = _bn_relu(f, sd, q + "conv1/bn")
        f = F.conv2d(f, sd[q + "conv2.weight"], groups=4)
        x = _crop(x, x.shape[2] - f.shape[2], x.shape[3] - f.shape[3])
        x = torch.cat([x, f], dim=1)
    return _bn_relu(x, sd, prefix + "blk_bna.bn")


def forward_logits(imgs_nchw, sd, mode, nr_types):
    """HoVerNet.forward in eval mode.  imgs_nchw: float tensor 0..255.  Returns OrderedDict."""
    k = 5 if mode == "original" else 3
    x = imgs_nchw / 255.0
    if mode == "fast":
        x = _tf_same_pad(x, 7, 1)
    x = F.conv2d(x, sd["conv0./.weight"])
    x = _bn_relu(x, sd, "conv0.bn")
    d = []
    for name, units, stride in _GROUPS:
        x = _residual_group(x, sd, name, units, stride)
        d.append(x)
    d[3] = F.conv2d(d[3], sd["conv_bot.weight"])
    if mode == "original":
        d[0] = _crop(d[0], 184, 184)
        d[1] = _crop(d[1], 72, 72)
    else:
        d[0] = _crop(d[0], 92, 92)
        d[1] = _crop(d[1], 36, 36)
    out = OrderedDict()
    branches = ("np", "hv") if nr_types is None else ("tp", "np", "hv")
    for b in branches:
        p = "decoder.%s." % b
        u3 = _up2(d[3]) + d[2]
        u3 = F.conv2d(u3, sd[p + "u3.conva.weight"])
        u3 = _dense(u3, sd, p + "u3.dense.", 8)
        u3 = F.conv2d(u3, sd[p + "u3.convf.weight"])
        u2 = _up2(u3) + d[1]
        u2 = F.conv2d(u2, sd[p + "u2.conva.weight"])
        u2 = _dense(u2, sd, p + "u2.dense.", 4)
        u2 = F.conv2d(u2, sd[p + "u2.convf.weight"])
        u1 = _up2(u2) + d[0]
        u1 = F.conv2d(_tf_same_pad(u1, k, 1), sd[p + "u1.conva.weight"])
        u0 = _bn_relu(u1, sd, p + "u0.bn")
        u0 = F.conv2d(u0, sd[p + "u0.conv.weight"], sd[p + "u0.conv.bias"])
        out[b] = u0
    return out


def infer_step(batch_u8_nhwc, sd, mode, nr_types, dtype=torch.float32, device="cpu"):
    """run_desc.py:171-197 restated.  batch: uint8 [B,H,W,3].  Returns np.float32/64 [B,h,w,C]."""
    sd = {k: v.to(device=device, dtype=dtype) for k, v in sd.items() if v.dim() > 0}
    x = torch.as_tensor(batch_u8_nhwc).to(device).type(dtype).permute(0, 3, 1, 2).contiguous()
    with torch.no_grad():
        pred = forward_logits(x, sd, mode, nr_types)
        pred = OrderedDict((k, v.permute(0, 2, 3, 1).contiguous()) for k, v in pred.items())
        pred["np"] = F.softmax(pred["np"], dim=-1)[..., 1:]
        if "tp" in pred:
            t = F.softmax(pred["tp"], dim=-1)
            t = torch.argmax(t, dim=-1, keepdim=True)
            pred["tp"] = t.type(dtype)
        out = torch.cat(list(pred.values()), -1)
    return out.cpu().numpy()


def to_torch_state_dict(np_sd):
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in np_sd.items()}
