"""HoVer-Net architecture facts (host side): checkpoint key inventory and patch geometry.

This is *data about* the reference graph, not the graph: the state-dict key list the
strict checkpoint load accepts (reference `infer/base.py:64-68`; keys follow
`models/hovernet/net_desc.py:17-99` and `models/hovernet/net_utils.py:71-266`), and the
patch in/out sizes the CLI fixes per mode (reference `run_infer.py:145-150`).
"""
from collections import OrderedDict

# (in_ch, [mid, mid, out], units, stride) -- reference net_desc.py:36-39
ENCODER_GROUPS = (
    ("d0", 64, (64, 64, 256), 3, 1),
    ("d1", 256, (128, 128, 512), 4, 2),
    ("d2", 512, (256, 256, 1024), 6, 2),
    ("d3", 1024, (512, 512, 2048), 3, 2),
)

PATCH_GEOMETRY = {  # mode -> (patch_input, patch_output) ; reference run_infer.py:145-150
    "original": (270, 80),
    "fast": (256, 164),
}


def decoder_ksize(mode):
    return 5 if mode == "original" else 3  # reference net_desc.py:76


def branch_names(nr_types):
    # reference net_desc.py:77-95 : order tp, np, hv when typed ; np, hv otherwise
    return ("np", "hv") if nr_types is None else ("tp", "np", "hv")


def _bn(keys, prefix, ch):
    keys[prefix + ".weight"] = (ch,)
    keys[prefix + ".bias"] = (ch,)
    keys[prefix + ".running_mean"] = (ch,)
    keys[prefix + ".running_var"] = (ch,)
    keys[prefix + ".num_batches_tracked"] = ()


def state_dict_spec(mode="original", nr_types=None, input_ch=3):
    """OrderedDict name -> shape, in the reference module-registration order."""
    if mode not in ("original", "fast"):
        raise ValueError("Unknown mode `%s`. Only support `original` or `fast`." % mode)
    k = decoder_ksize(mode)
    keys = OrderedDict()
    keys["conv0./.weight"] = (64, input_ch, 7, 7)
    _bn(keys, "conv0.bn", 64)
    for name, in_ch, (c1, c2, c3), units, _stride in ENCODER_GROUPS:
        unit_in = in_ch
        for u in range(units):
            p = "%s.units.%d." % (name, u)
            if u != 0:
                _bn(keys, p + "preact/bn", unit_in)
            keys[p + "conv1.weight"] = (c1, unit_in, 1, 1)
            _bn(keys, p + "conv1/bn", c1)
            keys[p + "conv2.weight"] = (c2, c1, 3, 3)
            _bn(keys, p + "conv2/bn", c2)
            keys[p + "conv3.weight"] = (c3, c2, 1, 1)
            unit_in = c3
        # every encoder group changes channels or stride => has a shortcut conv
        keys[name + ".shortcut.weight"] = (c3, in_ch, 1, 1)
        _bn(keys, name + ".blk_bna.bn", c3)
    keys["conv_bot.weight"] = (1024, 2048, 1, 1)
    for b in branch_names(nr_types):
        out_ch = nr_types if b == "tp" else 2
        for uname, cin, ca, units in (("u3", 1024, 256, 8), ("u2", 512, 128, 4)):
            p = "decoder.%s.%s." % (b, uname)
            keys[p + "conva.weight"] = (ca, cin, k, k)
            c = ca
            for u in range(units):
                q = p + "dense.units.%d." % u
                _bn(keys, q + "preact_bna/bn", c)
                keys[q + "conv1.weight"] = (128, c, 1, 1)
                _bn(keys, q + "conv1/bn", 128)
                keys[q + "conv2.weight"] = (32, 32, k, k)  # groups=4: 128/4 in, 32 out
                c += 32
            _bn(keys, p + "dense.blk_bna.bn", c)
            keys[p + "convf.weight"] = (c, c, 1, 1)
        keys["decoder.%s.u1.conva.weight" % b] = (64, 256, k, k)
        _bn(keys, "decoder.%s.u0.bn" % b, 64)
        keys["decoder.%s.u0.conv.weight" % b] = (out_ch, 64, 1, 1)
        keys["decoder.%s.u0.conv.bias" % b] = (out_ch,)
    keys["upsample2x.unpool_mat"] = (2, 2)
    return keys


def output_channels(nr_types):
    return 3 if nr_types is None else 4  # reference net_desc.py:22


def out_size(mode, in_size):
    """Spatial size of the network output for a square input (valid-conv arithmetic)."""
    k = decoder_ksize(mode)
    s = in_size if mode == "fast" else in_size - 6  # conv0: same-pad in fast, valid otherwise
    d0 = s
    d1 = (d0 + 1) // 2
    d2 = (d1 + 1) // 2
    d3 = (d2 + 1) // 2
    u3 = 2 * d3 - (k - 1) * 9  # conva + 8 dense units, all valid
    u2 = 2 * u3 - (k - 1) * 5
    return 2 * u2
