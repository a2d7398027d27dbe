"""PSNR / RangeInvariantPsnr on the GPU (reference: core/psnr.py:46-82).

Same names and argument meaning as the reference: ``PSNR(gt, pred, range_=None)`` and ``RangeInvariantPsnr(gt, pred)``
take (batch, H, W) images and return one value per image.  Here both are ONE pass over the two tensors
(``ds_psnr``: fp64 moments, closed-form metrics), the inputs stay in HBM (numpy inputs are uploaded, as the reference's
``allow_numpy`` converts them) and the result is a CUDA fp32 tensor.  ``psnr_frames`` evaluates stitched
(F, H, W, C) predictions in place, with the caller's un-normalisation and uint16 cast (split.py:198-203) folded
into the same pass.  There is no CPU fallback.
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib


def _to_cuda(x, device):
    if isinstance(x, np.ndarray):
        x = torch.Tensor(x)                       # float32, like the reference's allow_numpy
    if not isinstance(x, torch.Tensor):
        raise TypeError("expected a numpy array or a torch tensor")
    if not x.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError("diffsplit_b200: PSNR needs a CUDA device (no CPU fallback)")
        x = x.to(device or "cuda")
    return x.float()


def _run(gt, pred, n_frames, Cn, npix, gt_strides, pred_strides, scale=None, offset=None, quantize=False):
    dev = gt.device
    out = torch.empty((n_frames * Cn, 4), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        lib = _lib.lib()
        ws_bytes = lib.ds_psnr_workspace_bytes(n_frames, Cn, npix)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        a = _lib.PsnrArgs()
        a.d_gt, a.d_pred = gt.data_ptr(), pred.data_ptr()
        a.gt_frame_stride, a.gt_channel_stride, a.gt_pixel_stride = gt_strides
        a.pred_frame_stride, a.pred_channel_stride, a.pred_pixel_stride = pred_strides
        a.n_frames, a.C, a.npix = n_frames, Cn, npix
        a.unnormalize = int(scale is not None)
        a.quantize_u16 = int(bool(quantize))
        if scale is not None:
            sc = (C.c_double * Cn)(*[float(v) for v in np.asarray(scale, dtype=np.float64).reshape(-1)])
            of = (C.c_double * Cn)(*[float(v) for v in np.asarray(offset, dtype=np.float64).reshape(-1)])
            a.scale, a.offset = sc, of
        a.d_out, a.d_workspace, a.workspace_bytes = out.data_ptr(), ws.data_ptr(), ws_bytes
        _lib.check(lib.ds_psnr(C.byref(a), _lib.stream_ptr()))
    return out


def _batch(gt, pred):
    assert len(gt.shape) == 3, 'Images must be in shape: (batch,H,W)'
    gt = _to_cuda(gt, None).contiguous()
    pred = _to_cuda(pred, gt.device).contiguous()
    if pred.numel() != gt.numel():
        raise ValueError("gt and pred differ in size")
    B = gt.shape[0]
    npix = gt.numel() // B
    return _run(gt, pred, B, 1, npix, (npix, 0, 1), (npix, 0, 1))


def PSNR(gt, pred, range_=None):
    """core/psnr.py:46-62.  ``range_`` None: max - min of each ground-truth image."""
    out = _batch(gt, pred)
    if range_ is None:
        return out[:, 0].clone()
    range_ = torch.as_tensor(range_, dtype=torch.float32, device=out.device)
    return 20 * torch.log10(range_ / torch.sqrt(out[:, 2]))


def RangeInvariantPsnr(gt, pred):
    """core/psnr.py:65-82 (grayscale images)."""
    return _batch(gt, pred)[:, 1].clone()


def psnr_frames(target, prediction, mean_target=None, std_target=None, quantize_u16=True, channels_last=True):
    """Per-(frame, channel) PSNR and RangeInvariantPsnr of whole frames, read in place.

    ``target`` / ``prediction``: CUDA fp32, (F, H, W, C) as ``stitch_predictions`` returns (``channels_last``) or
    (F, C, H, W).  With ``mean_target`` / ``std_target`` (the dataset's normalisation dict entries, length C) both are
    un-normalised as in split.py:198-203 (``v * std + mean`` in float64; the prediction clamped to [0, 65535]; both cast to
    uint16 when ``quantize_u16``).  Returns two (F, C) CUDA tensors.
    """
    for t, what in ((target, "target"), (prediction, "prediction")):
        _lib.require_cuda(t, what)
    if target.shape != prediction.shape or target.dim() != 4:
        raise ValueError("target and prediction must both be (F,H,W,C) or (F,C,H,W)")
    target, prediction = target.float().contiguous(), prediction.float().contiguous()
    if channels_last:
        F, H, W, Cn = target.shape
        strides = (H * W * Cn, 1, Cn)
    else:
        F, Cn, H, W = target.shape
        strides = (Cn * H * W, H * W, 1)
    scale = offset = None
    if mean_target is not None:
        scale, offset = std_target, mean_target
    out = _run(target, prediction, F, Cn, H * W, strides, strides, scale, offset, quantize_u16 and scale is not None)
    out = out.view(F, Cn, 4)
    return out[..., 0].clone(), out[..., 1].clone()
