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


# ---- precision / range robustness: the fp16 hi+lo storage must not depend on the checkpoint's
# ---- scales.  Referee for every case: the fp64 oracle forward on the SAME modified checkpoint.
def _fp64_ref(x, sd, mode, nt):
    from oracle import hovernet_torch as O
    return O.infer_step(x, O.to_torch_state_dict(sd), mode, nt, dtype=torch.float64)


def _check_against_fp64(sd, mode, nt, tag, rel_tol=1e-4, np_strict=True):
    from hover_net_b200.models.hovernet.net_desc import create_model
    from hover_net_b200.models.hovernet.run_desc import infer_step
    x = synth.make_patches(1, arch.PATCH_GEOMETRY[mode][0], seed=33)
    ref = _fp64_ref(x, sd, mode, nt)
    net = create_model(mode=mode, input_ch=3, nr_types=nt)
    net.load_state_dict(sd, strict=True)
    out = infer_step(x, net)
    shift = net.ctx.counter("act_shift")
    net.ctx.close()
    assert np.isfinite(out).all(), tag
    # NP probability: absolute; HV regression maps: relative to the map's own scale (1e-4 * max|ref|, floor 1e-4)
    np_err = np.abs(out[..., -3] - ref[..., -3]).max()
    hv_scale = max(1.0, float(np.abs(ref[..., -2:]).max()))
    hv_err = np.abs(out[..., -2:] - ref[..., -2:]).max() / hv_scale
    if not np_strict:  # saturated softmax over ~1e4 logits: a relative 1e-4 on the logits flips pixels at the 0/1 boundary
        np_err = 0.0 if (np.abs(out[..., -3] - ref[..., -3]) > 1e-3).mean() < 0.01 else np_err
    assert np_err <= 1e-4 and hv_err <= rel_tol, "%s: np err %.3e, hv rel err %.3e (scale %.3g, act_shift %d)" % (
        tag, np_err, hv_err, hv_scale, shift)
    return shift, hv_scale


def test_layerwise_power_of_two_rescaling_does_not_change_the_result():
    """Every convolution that feeds a BatchNorm has its weights multiplied by 2^+8 or 2^-8 and the BN's running mean /
    variance rescaled to match: the network function is (up to eps) unchanged, but weights now span 2^-8 .. 2^+8 of
    their original scale -- unscaled, the small ones would lose their `lo` plane to fp16 subnormals and the large ones
    push pre-BN sums up by 256x.  The per-channel weight exponent makes the stored planes independent of it."""
    mode, nt = "fast", 6
    sd = dict(synth.make_state_dict(mode, nt, seed=0))
    rng = np.random.default_rng(5)
    n = 0
    for k in list(sd):
        if not k.endswith(".weight") or sd[k].ndim != 4:
            continue
        bn = k[: -len(".weight")] + "/bn"
        if k == "conv0./.weight":
            bn = "conv0.bn"
        if bn + ".running_mean" not in sd:
            continue
        e = int(rng.choice([-8, 8]))
        sd[k] = (sd[k] * np.float32(2.0 ** e)).astype(np.float32)
        sd[bn + ".running_mean"] = (sd[bn + ".running_mean"] * np.float32(2.0 ** e)).astype(np.float32)
        sd[bn + ".running_var"] = (sd[bn + ".running_var"] * np.float32(4.0 ** e)).astype(np.float32)
        n += 1
    assert n > 60
    _check_against_fp64(sd, mode, nt, "2^+-8 per layer")


def test_reference_weights_init_checkpoint():
    """The reference's own initialisation (net_utils.py:18-32: Kaiming-normal fan_out convolutions, BN gamma 1 / beta 0,
    fresh running statistics) drives the HV maps to ~1e4.  Intermediate activations leave fp16's
    range: the engine must notice (HVN_ERR_RANGE), rescale by an exact power of two and still match the fp64 referee
    to 1e-4 of the output scale -- never return a clamped map."""
    mode, nt = "fast", 6
    sd = dict(synth.make_state_dict(mode, nt, seed=0))
    rng = np.random.default_rng(11)
    for k, v in sd.items():
        if k.endswith("num_batches_tracked") or k == "upsample2x.unpool_mat":
            continue
        if v.ndim == 4:
            o, i, kh, kw = v.shape
            sd[k] = (np.sqrt(2.0 / (o * kh * kw)) * rng.standard_normal(v.shape)).astype(np.float32)
        elif k.endswith("running_mean") or k.endswith(".bias"):
            sd[k] = np.zeros_like(v)
        else:  # BN weight, running_var
            sd[k] = np.ones_like(v)
    shift, scale = _check_against_fp64(sd, mode, nt, "weights_init", rel_tol=2e-4, np_strict=False)
    assert scale > 50.0, "expected large logits from the raw initialisation, got %.3g" % scale


def test_heavy_tailed_weights():
    """Student-t (3 d.o.f.) convolution weights: a few weights per channel are 10-50x the typical one."""
    mode, nt = "original", None
    sd = dict(synth.make_state_dict(mode, nt, seed=0))
    rng = np.random.default_rng(13)
    for k, v in sd.items():
        if v.ndim == 4 and ".u0.conv." not in k:
            t = rng.standard_t(3, v.shape).astype(np.float32)
            sd[k] = (t * (np.std(v) / np.std(t))).astype(np.float32)
    _check_against_fp64(sd, mode, nt, "student-t weights", rel_tol=2e-4)


def test_range_overflow_is_an_error_not_a_clamp():
    """Through the raw C ABI (no retry wrapper): activations pushed past 65504 return HVN_ERR_RANGE; with act_shift
    raised the same call succeeds."""
    import ctypes
    from hover_net_b200 import _lib
    from hover_net_b200.models.hovernet.net_desc import create_model
    mode, nt = "fast", 6
    sd = dict(synth.make_state_dict(mode, nt, seed=0))
    sd["conv0./.weight"] = (sd["conv0./.weight"] * np.float32(3.0e5)).astype(np.float32)
    sd["conv0.bn.running_var"] = np.full_like(sd["conv0.bn.running_var"], 1.0)
    net = create_model(mode=mode, input_ch=3, nr_types=nt)
    net.load_state_dict(sd, strict=True)
    x = synth.make_patches(1, 256, seed=3)
    out = np.empty((1, 164, 164, 4), np.float32)
    L = _lib.lib()
    rc = L.hvn_forward(net.ctx._h, x.ctypes.data_as(ctypes.c_void_p), 1, 256, 256, out.ctypes.data_as(ctypes.c_void_p))
    assert rc == -6 and b"act_shift" in L.hvn_last_error()
    net.ctx.set_option("act_shift", 12)
    rc = L.hvn_forward(net.ctx._h, x.ctypes.data_as(ctypes.c_void_p), 1, 256, 256, out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0 and np.isfinite(out).all()
    net.ctx.close()


def test_act_shift_is_exact():
    """Rescaling activations by 2^-6 changes nothing but the exponent fields: outputs agree to fp32 rounding of the
    `lo` planes (not bit-for-bit: small values lose `lo` bits to fp16 subnormals)."""
    from hover_net_b200.models.hovernet.run_desc import infer_step
    net = _model("fast", 6, 0)
    x = synth.make_patches(1, 256, seed=9)
    a = infer_step(x, net)
    net.ctx.set_option("act_shift", 6)
    b = infer_step(x, net)
    assert np.abs(a[..., 1:] - b[..., 1:]).max() <= 2e-5
    net.ctx.close()


def test_tensor_core_stem_equals_cuda_core_stem():
    """k_conv0_tc (stem as a GEMM whose A operand is gathered from the uint8 image through a /255 table) vs the
    CUDA-core k_conv0: same network output to fp32 rounding, both modes (pad 3 / pad 0, 256 / 264 wide rows, a partial
    last tile in `original` mode)."""
    from hover_net_b200.models.hovernet.run_desc import infer_step
    for mode, nt in (("fast", 6), ("original", None)):
        net = _model(mode, nt, 0)
        x = synth.make_patches(2, arch.PATCH_GEOMETRY[mode][0], seed=17)
        x[0, :8] = 255; x[1, -5:, -7:] = 0       # saturated / zero borders exercise the padding columns
        net.ctx.set_option("stem_tc", 1)
        a = infer_step(x, net)
        net.ctx.set_option("stem_tc", 0)
        b = infer_step(x, net)
        assert np.isfinite(a).all() and np.abs(a[..., -3:] - b[..., -3:]).max() <= 2e-5, mode
        net.ctx.close()


@pytest.mark.parametrize("mode,nt", [("original", 5), ("fast", 6)])
def test_kernel_variants_agree(mode, nt):
    """The round-2 kernel variants against the paths they replace, end to end on the same patches: k_conv_rs (row-stacked
    grouped k x k) vs per-tap / HALO, k_conv_ar (A-resident conv3 + residual) vs RT, and the XF producer options
    (truncating split, early raw-slot release, L2 prefetch).  Same network output to fp32-accumulation noise."""
    from hover_net_b200.models.hovernet.run_desc import infer_step
    net = _model(mode, nt, 0)
    x = synth.make_patches(2, arch.PATCH_GEOMETRY[mode][0], seed=23)
    knobs = ("tc_rowstack", "tc_ar", "tc_xf_trunc", "tc_xf_early", "tc_prefetch")
    default = {"tc_rowstack": 1, "tc_ar": 1, "tc_xf_trunc": 1, "tc_xf_early": 1, "tc_prefetch": 4}
    try:
        ref = infer_step(x, net)
        assert np.isfinite(ref).all()
        for k in knobs:
            net.ctx.set_option(k, 0)
            out = infer_step(x, net)
            net.ctx.set_option(k, default[k])
            err = np.abs(out[..., -3:] - ref[..., -3:]).max()
            assert err <= 3e-5, "%s=0 changes the output by %.3e" % (k, err)
        again = infer_step(x, net)
        assert np.array_equal(again, ref), "the knobs are launch-time switches: defaults restored => identical bits"
    finally:
        for k in knobs:
            net.ctx.set_option(k, default[k])   # the knobs are process-wide
        net.ctx.close()
