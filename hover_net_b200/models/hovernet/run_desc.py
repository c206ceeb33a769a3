"""Drop-in for reference `models/hovernet/run_desc.py:171-197` (`infer_step`)."""
import numpy as np


def _as_u8(batch_data):
    x = batch_data
    if hasattr(x, "detach"):  # torch tensor from the default collate (infer_loader.py:65-72)
        x = x.detach().cpu().numpy()
    x = np.asarray(x)
    if x.dtype != np.uint8:
        # the reference casts to float32 and divides by 255; image patches are integral 0..255
        if not np.array_equal(x, np.round(x)) or x.min() < 0 or x.max() > 255:
            raise ValueError("infer_step expects integral RGB values in 0..255")
        x = x.astype(np.uint8)
    return np.ascontiguousarray(x)


def infer_step(batch_data, model):
    """uint8 NHWC [B,H,W,3] -> np.float32 [B,h,w,C] ([tp?, np, hv_x, hv_y])."""
    net = getattr(model, "module", model)
    return net.ctx.forward(_as_u8(batch_data))


def infer_step_fused(batch_data, model, return_pred=True):
    """infer_step + post_proc.process on every patch without the maps leaving the device.
    Returns (pred or None, inst [B,h,w] int32, table [B,max_rows,10] int64, n_rows [B])."""
    net = getattr(model, "module", model)
    return net.ctx.forward_postproc(_as_u8(batch_data), want_pred=return_pred)
