"""CPU suite part 1: the oracle against the committed golden vectors (generated from the real reference by
oracle/make_golden.py) and against size-independent properties."""
import os

import numpy as np
import pytest
import torch

from oracle import metrics_ref as M
from oracle import philox_ref as PH
from oracle import samplers_ref as S
from oracle import tiling_ref as TR
from oracle import unet_ref as U
from oracle.make_golden import TIMEPRED_CASES, UNET_CASES, Replay


def _load(gold_dir, name):
    return np.load(os.path.join(gold_dir, name))


@pytest.mark.parametrize("case", list(UNET_CASES))
def test_unet_oracle_matches_reference_golden(gold_dir, case):
    cfg, B, H, W = UNET_CASES[case]
    g = _load(gold_dir, f"unet_{case}.npz")
    sd = U.random_state_dict(cfg, seed=int(g["seed"]))
    assert abs(float(sum(v.double().sum() for v in sd.values())) - float(g["weight_checksum"])) < 1e-6
    y = U.unet_forward(sd, cfg, torch.from_numpy(g["x"]), torch.from_numpy(g["t"]))
    ref = torch.from_numpy(g["y"])
    assert y.shape == ref.shape
    assert (y - ref).abs().max().item() <= 4e-6 * max(1.0, ref.abs().max().item())


def test_flop_count_matches_baseline_md():
    cfgD = U.make_cfg("ddpm", 1, 1, 16, 16, (1, 2, 4, 8), (), 1, 32)
    assert abs(U.count_flops(cfgD, 64, 64) / 1e9 - 0.973) < 1e-3
    assert abs(U.count_flops(cfgD, 512, 512) / 1e9 - 70.72) < 1e-2
    cfgB = U.make_cfg("sr3", 6, 3, 64, 32, (1, 2, 4, 8, 8), (16,), 2, 128)
    assert abs(U.count_flops(cfgB, 128, 128) / 1e9 - 92.35) < 1e-2


def test_schedule_tables_match_reference(gold_dir):
    g = _load(gold_dir, "samplers.npz")
    for nm, so in (("tiny", dict(schedule="linear", n_timestep=6, linear_start=1e-4, linear_end=0.3)),
                   ("prod", dict(schedule="linear", n_timestep=2000, linear_start=1e-6, linear_end=1e-2))):
        tab = S.schedule_tables(so)
        for k in ("betas", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_mean_coef1",
                  "posterior_mean_coef2", "posterior_log_variance_clipped", "alphas_cumprod_prev"):
            assert np.array_equal(g[f"sched_{nm}_{k}"], tab[k].astype(np.float32)), (nm, k)
        assert np.array_equal(g[f"sched_{nm}_sqrt_alphas_cumprod_prev"], tab["sqrt_alphas_cumprod_prev"])
    # hard part 4 of the survey: amplification ~151 at t = T-1 for the production schedule
    assert 150 < S.schedule_tables(so)["sqrt_recip_alphas_cumprod"][-1] < 152


def _sampler_nets():
    cfg = U.make_cfg("sr3", 3, 2, 16, 8, (1, 2), (), 1, 16)
    cfgd = U.make_cfg("ddpm", 3, 2, 16, 8, (1, 2), (), 1, 16)
    cfgi = U.make_cfg("ddpm", 1, 1, 16, 8, (1, 2), (), 1, 16)
    return (cfg, U.random_state_dict(cfg, seed=3)), (cfgd, U.random_state_dict(cfgd, seed=4)), \
        (cfgi, U.random_state_dict(cfgi, seed=7), U.random_state_dict(cfgi, seed=8))


def test_sampler_oracles_match_reference_golden(gold_dir):
    g = _load(gold_dir, "samplers.npz")
    (cfg, sd), (cfgd, sdd), (cfgi, sd1, sd2) = _sampler_nets()
    tab = S.schedule_tables(dict(schedule="linear", n_timestep=6, linear_start=1e-4, linear_end=0.3))
    cond = torch.from_numpy(g["sr3_cond"])
    y = S.sr3_sample_loop(tab, lambda x, t: U.unet_forward(sd, cfg, x, t), cond, 2, True,
                          Replay(list(torch.from_numpy(g["sr3_noise"]))), continous=True)
    assert (y - torch.from_numpy(g["sr3_out"])).abs().max() < 1e-5
    y = S.ddpm_sample_loop(tab, lambda x, t: U.unet_forward(sdd, cfgd, x, t), cond, 2, True,
                           Replay(list(torch.from_numpy(g["ddpm_noise"]))), continous=True)
    assert (y - torch.from_numpy(g["ddpm_out"])).abs().max() < 1e-5
    x_in = torch.from_numpy(g["indi_x"])
    for T in (1, 4):
        y = S.indi_inference(lambda x, t: U.unet_forward(sd1, cfgi, x, t), x_in, T,
                             Replay(list(torch.from_numpy(g[f"indi_T{T}_noise"]))), continuous=True)
        assert y.shape[0] == 2 * (T + 1)          # the intent of the reference's tests/test_joint_indi.py
        assert (y - torch.from_numpy(g[f"indi_T{T}_out"])).abs().max() < 1e-5
    y = S.joint_indi_inference(lambda x, t: U.unet_forward(sd1, cfgi, x, t), lambda x, t: U.unet_forward(sd2, cfgi, x, t),
                               x_in, 3, Replay(list(torch.from_numpy(g["joint_noise"]))), t_float_start=0.5, continuous=True)
    assert (y - torch.from_numpy(g["joint_out"])).abs().max() < 1e-5


def test_psnr_oracles(gold_dir):
    g = _load(gold_dir, "samplers.npz")
    gt, pr = torch.from_numpy(g["psnr_gt"]), torch.from_numpy(g["psnr_pred"])
    assert np.allclose(S.psnr(gt, pr).numpy(), g["psnr"], atol=1e-4)
    assert np.allclose(S.range_invariant_psnr(gt, pr).numpy(), g["ripsnr"], atol=1e-4)


def test_metrics_oracle_matches_reference_golden(gold_dir):
    """float64 restatement of core/psnr.py + the un-normalisation of split.py:198-203 vs values recorded from the reference."""
    g = _load(gold_dir, "metrics.npz")
    for tag in ("unit", "u16", "offset"):
        assert np.allclose(M.psnr(g[f"{tag}_gt"], g[f"{tag}_pred"]), g[f"{tag}_psnr"], rtol=0, atol=2e-4)
        assert np.allclose(M.range_invariant_psnr(g[f"{tag}_gt"], g[f"{tag}_pred"]), g[f"{tag}_ripsnr"], rtol=0, atol=2e-4)
    assert np.allclose(M.psnr(g["unit_gt"], g["unit_pred"], 2.0), g["unit_psnr_range2"], atol=2e-4)
    t = M.unnormalize_u16(g["val_target"], g["val_mean"], g["val_std"], False)
    p = M.unnormalize_u16(g["val_prediction"], g["val_mean"], g["val_std"], True)
    assert p.min() == 0.0                                   # the clamped rows
    vals = [M.psnr(t[c:c + 1], p[c:c + 1])[0] for c in range(2)]
    assert np.allclose(vals, g["val_psnr"], atol=2e-4)
    same = M.psnr(g["unit_gt"], g["unit_gt"])
    assert np.all(np.isinf(same))                           # identical images: mse 0, as the reference (log10 of inf)


@pytest.mark.parametrize("case", list(TIMEPRED_CASES))
def test_time_predictor_oracle_matches_reference_golden(gold_dir, case):
    cfg, B, H, W = TIMEPRED_CASES[case]
    g = _load(gold_dir, "time_predictor.npz")
    sd = U.time_predictor_state_dict(cfg, seed=41)
    y = U.time_predictor_forward(sd, cfg, torch.from_numpy(g[f"{case}_x"]))
    assert y.shape == (B,) and np.allclose(y.numpy(), g[f"{case}_y"], rtol=0, atol=2e-6)


def test_tile_batch_oracle_matches_reference_dataset_items(gold_dir):
    """crop + normalise + mix vs items of the reference's SplitDatasetTiledPred (both input modes), bit-exact."""
    g = _load(gold_dir, "metrics.npz")
    fr, idx = g["tiles_frames"], [int(i) for i in g["tiles_idx"]]
    tg = TR.TileGrid((3, 96, 128), (1, 16, 16), (1, 32, 32), TR.SHIFT)
    nd = dict(mean_t=[600.3, 610.7], std_t=[598.9, 611.1], mean_in=1210.4, std_in=1207.3)
    raw = TR.crop_tiles(fr, tg, idx)
    for tag, w, fn in (("mix", (0.7, 0.4), False), ("normtar", (0.5, 0.5), True)):
        inp, tar = TR.normalise_and_mix(raw, nd["mean_t"], nd["std_t"], nd["mean_in"], nd["std_in"], w, fn)
        assert inp.dtype == np.float32 and tar.dtype == np.float32
        assert np.array_equal(inp, g[f"tiles_{tag}_input"]) and np.array_equal(tar, g[f"tiles_{tag}_target"])


def test_tiling_oracle_matches_reference_tables(gold_dir):
    g = _load(gold_dir, "tiling.npz")
    n = 0
    for key in g.files:
        if not key.startswith("shape_"):
            continue
        data, grid, patch = (tuple(int(v) for v in row) for row in g[key])
        mode = int(key.split("_")[-1])
        tg = TR.TileGrid(data, grid, patch, mode)
        assert np.array_equal(tg.patch_table(), g[key.replace("shape_", "tab_")]), key
        n += 1
    assert n == 18


def test_reference_unit_test_construction():
    """tests/test_tiling_setup.py:35-55 of the reference, restated on the oracle: frames whose pixel value is
    their own flat index, every tile cropped, stitched back, must equal the frames exactly."""
    tg = TR.TileGrid((5, 512, 512), (1, 128, 128), (1, 256, 256), TR.SHIFT)
    data = np.arange(5 * 512 * 512 * 2).reshape(5, 512, 512, 2)
    frames = np.stack([data[..., 0], data[..., 1]]).astype(np.float32)
    assert tg.total == 5 * 3 * 3
    out = TR.stitch(TR.crop_tiles(frames, tg), tg)
    assert out.shape == (5, 512, 512, 2)
    assert np.array_equal(out, frames.transpose(1, 2, 3, 0))


def test_stitch_ragged_golden(gold_dir):
    g = _load(gold_dir, "tiling.npz")
    tg = TR.TileGrid((3, 100, 130), (1, 16, 16), (1, 32, 32), TR.SHIFT)
    rng = np.random.default_rng(0)
    # same generator order as make_golden: the unit case draws first
    unit = TR.TileGrid((5, 512, 512), (1, 128, 128), (1, 256, 256), TR.SHIFT)
    rng.standard_normal((unit.total, 2, 256, 256)).astype(np.float32)
    tiles = rng.standard_normal((tg.total, 2, 32, 32)).astype(np.float32)
    assert np.array_equal(tiles[:, :, ::4, ::4], g["stitch_ragged_tiles"])
    assert np.array_equal(TR.stitch(tiles, tg), g["stitch_ragged_out"])


def test_production_tile_grid():
    tg = TR.TileGrid((10, 2048, 2048), (1, 256, 256), (1, 512, 512), TR.SHIFT)
    assert tg.counts == (10, 7, 7) and tg.total == 490          # EvaluateJointIndi.ipynb: (490, 2, 512, 512)
    assert tg.patch_location(0) == (0, 0, 0) and tg.patch_location(6) == (0, 0, 1536) and tg.patch_location(7) == (0, 256, 0)
    tg = TR.TileGrid((10, 2048, 2048), (1, 32, 32), (1, 64, 64), TR.SHIFT)
    assert tg.total == 39690


def test_philox_known_answer_and_moments():
    # Random123 known-answer vectors for Philox4x32-10
    z = PH.philox4x32_10(np.zeros((1, 4), np.uint32), np.zeros(2, np.uint32))[0]
    assert [hex(int(v)) for v in z] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    f = np.full((1, 4), 0xFFFFFFFF, np.uint32)
    z = PH.philox4x32_10(f, np.full(2, 0xFFFFFFFF, np.uint32))[0]
    assert [hex(int(v)) for v in z] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    x = PH.randn_like_cuda(1 << 16, seed=1234, offset=0, sm_count=148)
    assert abs(x.mean()) < 0.02 and abs(x.std() - 1) < 0.02
    assert PH.offset_increment(16 * 64 * 64, 148) == 4 and PH.launch_grid(16 * 64 * 64, 148) == 256
    assert PH.offset_increment(8 * 3 * 512 * 512, 148) == 24
