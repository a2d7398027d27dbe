"""Pin the oracle against the real reference and write ``tests/golden/*.npz``.

Runs ONLY in the build container (needs ``/root/reference``); the fixtures it
writes are committed and re-checked by ``pytest -m "not gpu"`` and, against the
CUDA library, by ``pytest -m gpu``.  Usage:  python -m oracle.make_golden
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("DIFFSPLIT_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")

sys.path.insert(0, os.path.dirname(HERE))
from oracle import samplers_ref as S          # noqa: E402
from oracle import tiling_ref as TR           # noqa: E402
from oracle import unet_ref as U              # noqa: E402

UNET_CASES = {
    # name: (cfg, B, H, W)
    "sr3_attn": (U.make_cfg("sr3", 3, 2, 16, 8, (1, 2, 2), (8,), 1, 16), 2, 16, 16),
    "sr3_wide": (U.make_cfg("sr3", 6, 3, 32, 16, (1, 2), (), 2, 16), 1, 16, 24),
    "ddpm_cifar": (U.make_cfg("ddpm", 9, 6, 16, 16, (1, 2, 4, 8), (), 1, 32), 1, 32, 32),
    "ddpm_hagen": (U.make_cfg("ddpm", 1, 1, 16, 16, (1, 2, 4, 8), (), 1, 32), 2, 64, 64),
}


def _import_reference():
    if not os.path.isdir(REF):
        raise SystemExit(f"reference not found at {REF}")
    sys.path.insert(0, REF)
    for name in ("albumentations", "skimage", "skimage.io"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["skimage.io"].imread = None
    sys.modules["albumentations"].Compose = None
    import model.sr3_modules.unet as sr3_unet
    import model.sr3_modules.diffusion as sr3_diff
    import model.ddpm_modules.unet as ddpm_unet
    import model.ddpm_modules.diffusion as ddpm_diff
    import model.ddpm_modules.indi as indi
    import model.ddpm_modules.joint_indi as joint
    import model.ddpm_modules.time_predictor as timepred
    import data.tiling_manager as tm
    import data.tile_stitcher as ts
    import data.split_dataset_tiledpred as tp
    import core.psnr as psnr
    return types.SimpleNamespace(sr3_unet=sr3_unet, sr3_diff=sr3_diff, ddpm_unet=ddpm_unet,
                                 ddpm_diff=ddpm_diff, indi=indi, joint=joint, tm=tm, ts=ts, tp=tp, psnr=psnr, timepred=timepred)


def build_ref_unet(R, cfg):
    mod = R.sr3_unet if cfg["variant"] == "sr3" else R.ddpm_unet
    return mod.UNet(in_channel=cfg["in_channel"], out_channel=cfg["out_channel"],
                    inner_channel=cfg["inner_channel"], norm_groups=cfg["norm_groups"],
                    channel_mults=cfg["channel_mults"], attn_res=cfg["attn_res"],
                    res_blocks=cfg["res_blocks"], dropout=0, image_size=cfg["image_size"]).eval()


def time_input(cfg, B, g):
    if cfg["variant"] == "sr3":
        return torch.rand((B, 1), generator=g) * 0.98 + 0.01
    return torch.rand((B,), generator=g)


def gen_keys(R):
    """state_dict key lists (name -> shape) of the reference modules: the checkpoint side of the boundary."""
    import json
    cases = dict(UNET_CASES)
    cases["sr3_16_128"] = (U.make_cfg("sr3", 6, 3, 64, 32, (1, 2, 4, 8, 8), (16,), 2, 128), 1, 128, 128)
    cases["sr3_64_512"] = (U.make_cfg("sr3", 6, 3, 64, 16, (1, 2, 4, 8, 16), (), 1, 512), 1, 512, 512)
    out = {}
    for name, (cfg, _, _, _) in cases.items():
        ref = build_ref_unet(R, cfg)
        out[name] = dict(cfg={k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg.items()},
                         keys={k: list(v.shape) for k, v in ref.state_dict().items()})
        assert set(out[name]["keys"]) == set(U.random_state_dict(cfg).keys()), name
    g = R.sr3_diff.GaussianDiffusion(build_ref_unet(R, UNET_CASES["sr3_attn"][0]), 16, channels=2)
    g.set_new_noise_schedule(dict(schedule="linear", n_timestep=6, linear_start=1e-4, linear_end=0.3), "cpu")
    out["sr3_sampler_keys"] = [k for k in g.state_dict().keys() if not k.startswith("denoise_fn.")]
    with open(os.path.join(GOLD, "state_dict_keys.json"), "w") as fh:
        json.dump(out, fh, indent=0, sort_keys=True)
    print(f"[keys] recorded state_dict keys for {len(cases)} architectures")


def gen_unet(R):
    for name, (cfg, B, H, W) in UNET_CASES.items():
        sd = U.random_state_dict(cfg, seed=11)
        ref = build_ref_unet(R, cfg)
        missing = ref.load_state_dict(sd, strict=True)
        g = torch.Generator().manual_seed(5)
        x = torch.randn((B, cfg["in_channel"], H, W), generator=g)
        t = time_input(cfg, B, g)
        with torch.no_grad():
            y_ref = ref(x, t)
        y_or = U.unet_forward(sd, cfg, x, t)
        err = (y_ref - y_or).abs().max().item()
        print(f"[unet {name}] out {tuple(y_ref.shape)} absmax {y_ref.abs().max():.4f} oracle-vs-reference max err {err:.3e}")
        assert err <= 2e-6 * max(1.0, y_ref.abs().max().item())
        wsum = float(sum(v.double().sum() for v in sd.values()))
        np.savez_compressed(os.path.join(GOLD, f"unet_{name}.npz"), x=x.numpy(), t=t.numpy(),
                            y=y_ref.numpy(), weight_checksum=np.float64(wsum), seed=np.int64(11))
        # flop count cross-check with forward hooks on the reference
    # flops: compare against the published per-sample figures
    cfgD = U.make_cfg("ddpm", 1, 1, 16, 16, (1, 2, 4, 8), (), 1, 32)
    print("[flops] hagen 64^2 per sample GF:", U.count_flops(cfgD, 64, 64) / 1e9, "(BASELINE.md: 0.973)")
    print("[flops] hagen 512^2 per sample GF:", U.count_flops(cfgD, 512, 512) / 1e9, "(BASELINE.md: 70.72)")
    cfgB = U.make_cfg("sr3", 6, 3, 64, 32, (1, 2, 4, 8, 8), (16,), 2, 128)
    print("[flops] sr3_16_128 per sample GF:", U.count_flops(cfgB, 128, 128) / 1e9, "(BASELINE.md: 92.35)")


class Recorder:
    """Replaces torch.randn / randn_like inside the reference so that the noise it
    consumed can be replayed into the oracle (and, on the GPU box, the CUDA path)."""

    _randn = torch.randn

    def __init__(self, seed):
        self.g = torch.Generator().manual_seed(seed)
        self.log = []

    def draw(self, shape):
        z = Recorder._randn(tuple(shape), generator=self.g)
        self.log.append(z)
        return z


def patched_randn(rec):
    class Ctx:
        def __enter__(self_):
            self_.a, self_.b = torch.randn, torch.randn_like
            torch.randn = lambda *s, **kw: rec.draw(s[0] if len(s) == 1 and not isinstance(s[0], int) else s)
            torch.randn_like = lambda x, **kw: rec.draw(x.shape)

        def __exit__(self_, *a):
            torch.randn, torch.randn_like = self_.a, self_.b
    return Ctx()


class Replay:
    def __init__(self, log):
        self.log, self.i = log, 0

    def __call__(self, shape):
        z = self.log[self.i]
        self.i += 1
        assert tuple(z.shape) == tuple(shape), (z.shape, shape)
        return z


def gen_samplers(R):
    sched = dict(schedule="linear", n_timestep=6, linear_start=1e-4, linear_end=0.3)
    out = {}
    # ---- schedule tables (T=2000 production schedule + the tiny one) -----------------
    prod = dict(schedule="linear", n_timestep=2000, linear_start=1e-6, linear_end=1e-2)
    for nm, so in (("tiny", sched), ("prod", prod)):
        ref = R.sr3_diff.GaussianDiffusion(None, 8)
        ref.set_new_noise_schedule(so, "cpu")
        tab = S.schedule_tables(so)
        for k in ("betas", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_mean_coef1",
                  "posterior_mean_coef2", "posterior_log_variance_clipped", "alphas_cumprod_prev"):
            assert np.array_equal(getattr(ref, k).numpy(), tab[k].astype(np.float32)), k
            out[f"sched_{nm}_{k}"] = getattr(ref, k).numpy()
        assert np.array_equal(ref.sqrt_alphas_cumprod_prev, tab["sqrt_alphas_cumprod_prev"])
        out[f"sched_{nm}_sqrt_alphas_cumprod_prev"] = ref.sqrt_alphas_cumprod_prev
    print("[sched] tables identical (tiny, prod)")

    # ---- SR3 conditional loop with a real (tiny) UNet ---------------------------------
    cfg = U.make_cfg("sr3", 3, 2, 16, 8, (1, 2), (), 1, 16)
    sd = U.random_state_dict(cfg, seed=3)
    net = build_ref_unet(R, cfg)
    net.load_state_dict(sd)
    ref = R.sr3_diff.GaussianDiffusion(net, 16, channels=2, conditional=True)
    ref.set_new_noise_schedule(sched, "cpu")
    g = torch.Generator().manual_seed(1)
    cond = torch.rand((2, 1, 16, 16), generator=g) * 2 - 1
    rec = Recorder(2)
    import tqdm as _tq
    R.sr3_diff.tqdm = lambda it, **kw: it
    with patched_randn(rec):
        y_ref = ref.p_sample_loop(cond, continous=True)
    tab = S.schedule_tables(sched)
    y_or = S.sr3_sample_loop(tab, lambda x, t: U.unet_forward(sd, cfg, x, t), cond, 2, True, Replay(rec.log), continous=True)
    err = (y_ref - y_or).abs().max().item()
    print(f"[sr3 loop] {tuple(y_ref.shape)} oracle-vs-reference max err {err:.3e}")
    assert err < 1e-5
    out.update(sr3_cond=cond.numpy(), sr3_out=y_ref.numpy(), sr3_noise=np.stack([z.numpy() for z in rec.log]))

    # ---- DDPM conditional loop ---------------------------------------------------------
    cfgd = U.make_cfg("ddpm", 3, 2, 16, 8, (1, 2), (), 1, 16)
    sdd = U.random_state_dict(cfgd, seed=4)
    netd = build_ref_unet(R, cfgd)
    netd.load_state_dict(sdd)
    refd = R.ddpm_diff.GaussianDiffusion(netd, 16, channels=2, conditional=True)
    refd.set_new_noise_schedule(sched, "cpu")
    R.ddpm_diff.tqdm = lambda it, **kw: it
    rec = Recorder(6)
    with patched_randn(rec):
        y_ref = refd.p_sample_loop(cond, continous=True)
    y_or = S.ddpm_sample_loop(tab, lambda x, t: U.unet_forward(sdd, cfgd, x, t), cond, 2, True, Replay(rec.log), continous=True)
    err = (y_ref - y_or).abs().max().item()
    print(f"[ddpm loop] {tuple(y_ref.shape)} oracle-vs-reference max err {err:.3e}")
    assert err < 1e-5
    out.update(ddpm_out=y_ref.numpy(), ddpm_noise=np.stack([z.numpy() for z in rec.log]))

    # ---- InDI / JointIndi ----------------------------------------------------------------
    cfgi = U.make_cfg("ddpm", 1, 1, 16, 8, (1, 2), (), 1, 16)
    sd1, sd2 = U.random_state_dict(cfgi, seed=7), U.random_state_dict(cfgi, seed=8)
    n1, n2 = build_ref_unet(R, cfgi), build_ref_unet(R, cfgi)
    n1.load_state_dict(sd1)
    n2.load_state_dict(sd2)
    R.indi.tqdm = lambda it, **kw: it
    x_in = torch.rand((2, 1, 16, 16), generator=g) * 2 - 1
    for T in (1, 4):
        indi = R.indi.InDI(n1, 16, channels=1, out_channel=1, conditional=False, val_schedule_opt={"n_timestep": T})
        indi.set_new_noise_schedule({"n_timestep": T}, "cpu")
        rec = Recorder(20 + T)
        with patched_randn(rec):
            y_ref = indi.inference(x_in, continuous=True)
        y_or = S.indi_inference(lambda x, t: U.unet_forward(sd1, cfgi, x, t), x_in, T, Replay(rec.log), continuous=True)
        err = (y_ref - y_or).abs().max().item()
        print(f"[indi T={T}] {tuple(y_ref.shape)} oracle-vs-reference max err {err:.3e}")
        assert err < 1e-5 and y_ref.shape[0] == 2 * (T + 1)
        out.update({f"indi_T{T}_out": y_ref.numpy(), f"indi_T{T}_noise": np.stack([z.numpy() for z in rec.log])})
    joint = R.joint.JointIndi(None, 16, channels=1, out_channel=1, conditional=False, denoise_fn_ch1=n1,
                              denoise_fn_ch2=n2, val_schedule_opt={"n_timestep": 3})
    joint.set_new_noise_schedule({"n_timestep": 3}, "cpu")
    rec = Recorder(31)
    with patched_randn(rec):
        y_ref = joint.inference(x_in, continuous=True, t_float_start=0.5)
    y_or = S.joint_indi_inference(lambda x, t: U.unet_forward(sd1, cfgi, x, t),
                                  lambda x, t: U.unet_forward(sd2, cfgi, x, t), x_in, 3, Replay(rec.log),
                                  t_float_start=0.5, continuous=True)
    err = (y_ref - y_or).abs().max().item()
    print(f"[joint indi] {tuple(y_ref.shape)} oracle-vs-reference max err {err:.3e}")
    assert err < 1e-5
    out.update(indi_x=x_in.numpy(), joint_out=y_ref.numpy(), joint_noise=np.stack([z.numpy() for z in rec.log]))

    # ---- PSNR metrics (core/psnr.py) ---------------------------------------------------------
    gt = torch.rand((3, 32, 32), generator=g) * 900
    pr = gt + torch.randn((3, 32, 32), generator=g) * 20
    assert torch.allclose(R.psnr.PSNR(gt, pr), S.psnr(gt, pr), atol=1e-4)
    assert torch.allclose(R.psnr.RangeInvariantPsnr(gt, pr), S.range_invariant_psnr(gt, pr), atol=1e-4)
    out.update(psnr_gt=gt.numpy(), psnr_pred=pr.numpy(), psnr=R.psnr.PSNR(gt, pr).numpy(),
               ripsnr=R.psnr.RangeInvariantPsnr(gt, pr).numpy())
    np.savez_compressed(os.path.join(GOLD, "samplers.npz"), **out)


def gen_tiling(R):
    shapes = [
        ((5, 512, 512), (1, 128, 128), (1, 256, 256)),       # the reference unit test
        ((10, 2048, 2048), (1, 256, 256), (1, 512, 512)),    # split.py:60-61 production
        ((3, 100, 130), (1, 16, 16), (1, 32, 32)),           # ragged: shifted last row/col
        ((2, 64, 64), (1, 64, 64), (1, 64, 64)),             # patch == grid, single tile
        ((1, 33, 47), (1, 8, 4), (1, 16, 12)),
        ((4, 70, 70), (1, 32, 32), (1, 64, 64)),
    ]
    out = {}
    n = 0
    for data, grid, patch in shapes:
        for mode in (TR.TRIM, TR.PAD, TR.SHIFT):
            ref = R.tm.TileIndexManager(data, grid, patch, mode)
            tg = TR.TileGrid(data, grid, patch, mode)
            assert ref.total_grid_count() == tg.total, (data, grid, patch, mode)
            for d in range(3):
                assert ref.get_individual_dim_grid_count(d) == tg.counts[d]
                assert ref.grid_count(d) == tg.strides[d]
            tab = tg.patch_table()
            step = max(1, tg.total // 400)
            for i in list(range(0, tg.total, step)) + [tg.total - 1]:
                assert tuple(int(v) for v in ref.get_patch_location_from_dataset_idx(i)) == tuple(tab[i])
                assert tuple(int(v) for v in ref.get_location_from_dataset_idx(i)) == tg.grid_location(i)
            out[f"tab_{n}_{mode}"] = tab.astype(np.int32)
            out[f"shape_{n}_{mode}"] = np.array([data, grid, patch], dtype=np.int32)
        n += 1
    # stitch: the reference test's construction + a ragged random case
    import contextlib
    import io
    rng = np.random.default_rng(0)
    for tag, (data, grid, patch) in (("unit", shapes[0]), ("ragged", shapes[2]), ("odd", shapes[4])):
        tg = TR.TileGrid(data, grid, patch, TR.SHIFT)
        ref = R.tm.TileIndexManager(data, grid, patch, R.tm.TilingMode.ShiftBoundary)
        frames = np.arange(2 * np.prod(data)).reshape(data + (2,)).transpose(3, 0, 1, 2).astype(np.float32)   # (C,F,H,W)
        if patch[1] != patch[2]:
            continue
        tiles = TR.crop_tiles(frames, tg)
        with contextlib.redirect_stdout(io.StringIO()):
            s_ref = R.ts.stitch_predictions(tiles, ref)
        s_or = TR.stitch(tiles, tg)
        assert np.array_equal(s_ref, s_or)
        assert np.array_equal(s_or, frames.transpose(1, 2, 3, 0)), "stitch(crop(frames)) must be the identity"
        noise_tiles = rng.standard_normal(tiles.shape).astype(np.float32)
        with contextlib.redirect_stdout(io.StringIO()):
            s_ref = R.ts.stitch_predictions(noise_tiles, ref)
        assert np.array_equal(s_ref, TR.stitch(noise_tiles, tg))
        if tag == "ragged":
            out["stitch_ragged_tiles"] = noise_tiles[:, :, ::4, ::4].copy()     # thumbnails only, keep the file small
            out["stitch_ragged_out_sum"] = np.float64(s_ref.astype(np.float64).sum())
            out["stitch_ragged_out"] = s_ref
    # the reference's own dataset-level test (tests/test_tiling_setup.py:35-55) through its dataset class
    def get_data(*a, **k):
        d = np.arange(5 * 512 * 512 * 2).reshape(5, 512, 512, 2)
        return {i: d[..., i] for i in range(2)}
    import data.split_dataset as sdmod
    sdmod.load_data = get_data
    ident = {'mean_input': 0, 'mean_target': np.array([0, 0]), 'std_input': 1, 'std_target': np.array([1, 1]),
             'target0_max': 1, 'target1_max': 1, 'input_max': 1}
    with contextlib.redirect_stdout(io.StringIO()):
        dset = R.tp.SplitDatasetTiledPred('Hagen', None, 256, grid_size=128, upper_clip=False, normalization_dict=ident,
                                          enable_transforms=False, uncorrelated_channels=False, random_patching=False)
    tg = TR.TileGrid((5, 512, 512), (1, 128, 128), (1, 256, 256), TR.SHIFT)
    assert len(dset) == tg.total
    frames = np.stack([np.stack(get_data()[c]) for c in range(2)])
    for i in (0, 1, 8, 9, 17, tg.total - 1):
        item = dset[i]
        assert tuple(int(v) for v in dset.patch_location(i)) == tg.patch_location(i)
        assert np.array_equal(item['target'], TR.crop_tiles(frames, tg, [i])[0])
    print(f"[tiling] {n} shapes x 3 modes identical; stitch identical (unit, ragged); dataset crop identical")
    np.savez_compressed(os.path.join(GOLD, "tiling.npz"), **out)


def gen_metrics(R):
    """PSNR / RangeInvariantPsnr vectors from the reference's core/psnr.py, the un-normalisation of split.py:198-203,
    and normalised tile batches through the reference's SplitDatasetTiledPred with a non-trivial normalisation dict."""
    from oracle import metrics_ref as M
    import contextlib
    import io
    rng = np.random.default_rng(7)
    out = {}
    cases = {
        "unit": (rng.standard_normal((3, 24, 40)).astype(np.float32), 0.1),
        "u16": (rng.integers(0, 4000, size=(2, 32, 32)).astype(np.float32), 25.0),
        "offset": ((rng.standard_normal((2, 16, 48)) * 3 + 1000).astype(np.float32), 0.5),
    }
    for tag, (gt, noise) in cases.items():
        pred = (gt * 0.8 + 0.3 + noise * rng.standard_normal(gt.shape)).astype(np.float32)
        p_ref = R.psnr.PSNR(torch.from_numpy(gt), torch.from_numpy(pred)).numpy()
        r_ref = R.psnr.RangeInvariantPsnr(torch.from_numpy(gt), torch.from_numpy(pred)).numpy()
        assert np.allclose(M.psnr(gt, pred), p_ref, rtol=0, atol=2e-4), (tag, M.psnr(gt, pred), p_ref)
        assert np.allclose(M.range_invariant_psnr(gt, pred), r_ref, rtol=0, atol=2e-4), (tag, M.range_invariant_psnr(gt, pred), r_ref)
        out[f"{tag}_gt"], out[f"{tag}_pred"], out[f"{tag}_psnr"], out[f"{tag}_ripsnr"] = gt, pred, p_ref, r_ref
    p_fixed = R.psnr.PSNR(torch.from_numpy(cases["unit"][0]), torch.from_numpy(out["unit_pred"]), range_=torch.tensor(2.0)).numpy()
    assert np.allclose(M.psnr(cases["unit"][0], out["unit_pred"], 2.0), p_fixed, atol=2e-4)
    out["unit_psnr_range2"] = p_fixed
    # the validation loop's un-normalise + uint16 cast + PSNR, split.py:189-208, on one (2,H,W) normalised pair
    mean_t, std_t = np.array([600.25, 610.75]), np.array([600.25, 610.75])
    target = rng.uniform(-1, 1, size=(2, 40, 56)).astype(np.float32)
    prediction = (target + 0.05 * rng.standard_normal(target.shape)).astype(np.float32)
    prediction[0, :2] = -1.2                       # below 0 after un-normalisation: clamped
    target_img = (target * std_t.reshape(-1, 1, 1) + mean_t.reshape(-1, 1, 1)).astype(np.uint16)
    pred_img = prediction * std_t.reshape(-1, 1, 1) + mean_t.reshape(-1, 1, 1)
    pred_img[pred_img < 0] = 0
    pred_img[pred_img > 65535] = 65535
    pred_img = pred_img.astype(np.uint16)
    vals = np.array([R.psnr.PSNR(target_img[c:c + 1] * 1.0, pred_img[c:c + 1] * 1.0).mean().item() for c in range(2)])
    assert np.array_equal(M.unnormalize_u16(target, mean_t, std_t, False), target_img * 1.0)
    assert np.array_equal(M.unnormalize_u16(prediction, mean_t, std_t, True), pred_img * 1.0)
    mine = np.array([M.psnr(M.unnormalize_u16(target, mean_t, std_t, False)[c:c + 1],
                            M.unnormalize_u16(prediction, mean_t, std_t, True)[c:c + 1])[0] for c in range(2)])
    assert np.allclose(mine, vals, atol=2e-4), (mine, vals)
    out["val_target"], out["val_prediction"], out["val_mean"], out["val_std"], out["val_psnr"] = target, prediction, mean_t, std_t, vals
    # normalised tile batches through the reference dataset class (split_dataset.py:237-278), both input modes
    fr = rng.integers(0, 1994, size=(2, 3, 96, 128)).astype(np.uint16)

    def get_data(*a, **k):
        return {i: fr[i] for i in range(2)}
    import data.split_dataset as sdmod
    sdmod.load_data = get_data
    nd = {'mean_input': np.float64(1210.4), 'std_input': np.float64(1207.3), 'mean_target': np.array([600.3, 610.7]),
          'std_target': np.array([598.9, 611.1]), 'target0_max': 1, 'target1_max': 1, 'input_max': 1}
    tg = TR.TileGrid((3, 96, 128), (1, 16, 16), (1, 32, 32), TR.SHIFT)
    idx = [0, 5, 17, tg.total - 1]
    for tag, kw in (("mix", dict(channel_weights=[0.7, 0.4])), ("normtar", dict(channel_weights=[0.5, 0.5], input_from_normalized_target=True))):
        with contextlib.redirect_stdout(io.StringIO()):
            dset = R.tp.SplitDatasetTiledPred('Hagen', None, 32, grid_size=16, upper_clip=False, normalization_dict=nd,
                                              enable_transforms=False, uncorrelated_channels=False, random_patching=False, **kw)
        assert len(dset) == tg.total
        items = [dset[i] for i in idx]
        inp_ref = np.stack([it['input'] for it in items])
        tar_ref = np.stack([it['target'] for it in items])
        inp, tar = TR.normalise_and_mix(TR.crop_tiles(fr, tg, idx), nd['mean_target'], nd['std_target'], nd['mean_input'],
                                        nd['std_input'], kw['channel_weights'], kw.get('input_from_normalized_target', False))
        assert inp_ref.dtype == np.float32 and np.array_equal(inp, inp_ref) and np.array_equal(tar, tar_ref), tag
        out[f"tiles_{tag}_input"], out[f"tiles_{tag}_target"] = inp_ref, tar_ref
    out["tiles_frames"], out["tiles_idx"] = fr, np.array(idx)
    print("[metrics] PSNR / RangeInvariantPsnr / un-normalise agree with core/psnr.py; tile batches identical to the dataset class")
    np.savez_compressed(os.path.join(GOLD, "metrics.npz"), **out)


TIMEPRED_CASES = {
    "hagen": (U.make_cfg("ddpm", 1, 1, 16, 16, (1, 2, 4, 8), (), 1, 64, with_time_emb=False), 3, 64, 64),   # the shipped config
    "multi": (U.make_cfg("ddpm", 3, 2, 16, 8, (1, 2), (8,), 2, 16, with_time_emb=False), 2, 16, 24),
}


def gen_time_predictor(R):
    """TimePredictor (time_predictor.py:13-45) with seeded weights loaded through its own load_state_dict (strict: pins the
    key names), inputs and outputs recorded."""
    out = {}
    for tag, (cfg, B, H, W) in TIMEPRED_CASES.items():
        sd = U.time_predictor_state_dict(cfg, seed=41)
        ref = R.timepred.TimePredictor(in_channel=cfg["in_channel"], out_channel=cfg["out_channel"],
                                       inner_channel=cfg["inner_channel"], norm_groups=cfg["norm_groups"],
                                       channel_mults=cfg["channel_mults"], attn_res=cfg["attn_res"],
                                       res_blocks=cfg["res_blocks"], dropout=0, image_size=cfg["image_size"]).eval()
        ref.load_state_dict(sd, strict=True)
        x = torch.randn((B, cfg["in_channel"], H, W), generator=torch.Generator().manual_seed(5))
        with torch.no_grad():
            y_ref = ref(x)
        y = U.time_predictor_forward(sd, cfg, x)
        err = (y - y_ref).abs().max().item()
        assert err < 2e-6, (tag, err)
        out[f"{tag}_x"], out[f"{tag}_y"] = x.numpy(), y_ref.numpy()
        print(f"[time predictor] {tag}: oracle vs reference {err:.2e}; values {y_ref.numpy()}")
    np.savez_compressed(os.path.join(GOLD, "time_predictor.npz"), **out)


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    R = _import_reference()
    gen_tiling(R)
    gen_metrics(R)
    gen_time_predictor(R)
    gen_keys(R)
    gen_unet(R)
    gen_samplers(R)
    tot = sum(os.path.getsize(os.path.join(GOLD, f)) for f in os.listdir(GOLD))
    print(f"golden fixtures written to {GOLD}: {tot / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
