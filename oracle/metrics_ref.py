"""CPU restatement of the evaluation metrics next to the sampling path (oracle; test infrastructure only).

  * PSNR, RangeInvariantPsnr      core/psnr.py:31-82
  * the caller's un-normalisation split.py:198-203 (v * std + mean in float64, prediction clamped to [0, 65535],
                                  both cast to uint16 by truncation)

Pinned by ``oracle/make_golden.py`` against the reference's own ``core.psnr`` (torch, float32) on seeded images
(``tests/golden/metrics.npz``).  The restatement follows the reference step by step but in float64 numpy, so it agrees
with the float32 reference to float32 rounding (the reference itself notes that its numpy and torch versions differ
slightly, core/psnr.py:3).
"""
import numpy as np


def _flat(x):
    x = np.asarray(x, dtype=np.float64)
    assert x.ndim == 3, "Images must be in shape: (batch,H,W)"
    return x.reshape(len(x), -1)


def psnr(gt, pred, range_=None):
    """core/psnr.py:39-62."""
    gt, pred = _flat(gt), _flat(pred)
    if range_ is None:
        range_ = gt.max(axis=1) - gt.min(axis=1)                                   # :40-41
    mse = np.mean((gt - pred) ** 2, axis=1)                                        # :43
    with np.errstate(divide="ignore"):
        return 20 * np.log10(range_ / np.sqrt(mse))                                # :44


def range_invariant_psnr(gt, pred):
    """core/psnr.py:65-82 with zero_mean :31-32, fix_range :35-37, fix :40-42; torch.std is the unbiased estimator."""
    gt, pred = _flat(gt), _flat(pred)
    std = gt.std(axis=1, ddof=1, keepdims=True)
    ra = (gt.max(axis=1) - gt.min(axis=1)) / std[:, 0]                             # :79
    gt_ = (gt - gt.mean(axis=1, keepdims=True)) / std                              # :80
    gt0 = gt_ - gt_.mean(axis=1, keepdims=True)
    x0 = pred - pred.mean(axis=1, keepdims=True)
    a = np.sum(gt0 * x0, axis=1, keepdims=True) / np.sum(x0 * x0, axis=1, keepdims=True)
    return psnr(gt0.reshape(len(gt0), 1, -1), (x0 * a).reshape(len(gt0), 1, -1), ra)   # :81


def unnormalize_u16(x, mean, std, clamp):
    """split.py:198-203: x (C,H,W) float32 normalised -> the uint16 image the metrics are computed on (as float64)."""
    mean = np.asarray(mean, dtype=np.float64).reshape(-1, 1, 1)
    std = np.asarray(std, dtype=np.float64).reshape(-1, 1, 1)
    v = np.asarray(x, dtype=np.float32) * std + mean
    if clamp:
        v = np.clip(v, 0, 65535)
    return v.astype(np.uint16) * 1.0
