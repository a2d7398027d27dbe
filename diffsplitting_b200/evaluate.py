"""MMSE evaluation of the joint InDI split over tiled frames, on the GPU end to end.

Reference: notebooks/EvaluateJointIndi.ipynb cells 55-62 - for each of ``mmse_count`` passes and every tile, the two
network inputs are mixed from the normalised target channels (``t0*(1-w) + t1*w`` and the mirrored one), each is pushed
through ``JointIndi.inference`` (channel 0 of the first result and channel 1 of the second are kept), the passes are
averaged (``mmse_pred += pred / mmse_count``), the tiles stitched and ``RangeInvariantPsnr`` taken per frame and channel.
Here the tiles come from frames resident in HBM (``TiledFrames.batch``), tiles are batched (``chunk``), the stitch and the
metric run on the device (``ds_stitch_tiles``, ``ds_psnr``), and nothing returns to the host but the final numbers.
"""
import torch

from .core.psnr import psnr_frames
from .data.tile_stitcher import stitch_predictions
from .parallel import chunk_ranges


def mixed_inputs(target, mixing_t):
    """(n,2,P,P) normalised targets -> the two (n,1,P,P) inputs of cell 55."""
    t0, t1 = target[:, 0:1], target[:, 1:2]
    return t0 * (1 - mixing_t) + t1 * mixing_t, t1 * (1 - mixing_t) + t0 * mixing_t


def _clamp_t(t, eps=1e-6):
    t = float(t)
    return 0.0 if t < eps else (1.0 if t > 1 else t)


def _predict(sampler, inp, t, num_timesteps, channel):
    """Channel ``channel`` of ``sampler.inference`` for every tile of the batch.  t <= 0 is the clean end of the InDI path
    (the reference computes 0 / 0 there): the input already is the prediction."""
    if hasattr(sampler, "indi1") and (t <= 0.0 or t >= 1.0):
        # JointIndi at an end point: one of its two loops would start at t = 0; only the loop whose channel is kept runs
        return _predict(sampler.indi1, inp, t, num_timesteps, 0) if channel == 0 else _predict(sampler.indi2, inp, 1 - t, num_timesteps, 0)
    if t <= 0.0:
        return inp[:, 0]
    out = sampler.inference(inp, continuous=False, t_float_start=t, num_timesteps=num_timesteps, all_samples=True)
    return out[:, channel]


def mmse_predict_tiles(joint_model, tiled, mixing_t=0.5, num_timesteps=5, mmse_count=5, chunk=1, t_float_start=None,
                       replay_reference_rng=False):
    """Returns (mmse prediction, targets), both (N,2,P,P) CUDA fp32 for all N tiles of ``tiled``.

    ``t_float_start``: None (= ``mixing_t`` for both inputs, as the notebook), a pair of floats, or a pair of callables
    ``f(inp) -> float`` (the time predictors).  ``replay_reference_rng``: run the full ``JointIndi.inference`` for both inputs
    as the notebook does (half of each result is discarded there) so that, with ``chunk=1``, the torch CUDA generator is
    consumed in the notebook's order; otherwise only the two loops whose output is kept are run (half the work).
    """
    total = len(tiled)
    P = tiled.patch_size
    dev = tiled.frames.device
    pred = torch.zeros((total, 2, P, P), dtype=torch.float32, device=dev)
    targets = torch.empty((total, 2, P, P), dtype=torch.float32, device=dev)
    for m in range(mmse_count):
        for first, n in chunk_ranges(total, chunk):
            _, target = tiled.batch(first, n)
            if m == 0:
                targets[first:first + n] = target
            inp0, inp1 = mixed_inputs(target, mixing_t)
            if t_float_start is None:
                ts = (mixing_t, mixing_t)
            else:
                ts = tuple(t(i) if callable(t) else t for t, i in zip(t_float_start, (inp0, inp1)))
            t0, t1 = _clamp_t(ts[0]), _clamp_t(ts[1])
            # continuous=False returns only the LAST batch element (the reference's `ret_img[-1:]`, indi.py:95): the batched
            # path asks for every element's final state (`all_samples`), so chunk > 1 predicts every tile of the chunk
            if replay_reference_rng:
                p0 = _predict(joint_model, inp0, t0, num_timesteps, 0)
                p1 = _predict(joint_model, inp1, t1, num_timesteps, 1)
            else:       # JointIndi.inference :131-135: channel 0 = indi1 at t, channel 1 = indi2 at 1 - t
                p0 = _predict(joint_model.indi1, inp0, t0, num_timesteps, 0)
                p1 = _predict(joint_model.indi2, inp1, 1 - t1, num_timesteps, 0)
            pred[first:first + n, 0].add_(p0 / mmse_count)
            pred[first:first + n, 1].add_(p1 / mmse_count)
    return pred, targets


def evaluate_mmse(joint_model, tiled, **kw):
    """Cells 55-62 in one call: dict with the stitched (F,H,W,2) prediction / target and the per-frame
    ``RangeInvariantPsnr`` of both channels (``(F,2)`` CUDA tensor)."""
    pred, targets = mmse_predict_tiles(joint_model, tiled, **kw)
    pred_st = stitch_predictions(pred, tiled.tile_manager)
    tar_st = stitch_predictions(targets, tiled.tile_manager)
    _, ri = psnr_frames(tar_st, pred_st)
    return {"prediction": pred_st, "target": tar_st, "range_invariant_psnr": ri}
