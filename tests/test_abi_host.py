"""CPU suite part 2: the C-ABI library loads and exports every declared symbol; host-side logic (architecture
walk, state_dict keys, tile index arithmetic, factory) matches the reference.  No kernel is launched."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest
import torch

from diffsplitting_b200 import _lib
from diffsplitting_b200.data import TileIndexManager, TilingMode, get_tile_manager, get_tiling_dataset
from diffsplitting_b200.model import networks
from diffsplitting_b200.model.samplers import (GaussianDiffusionDdpm, GaussianDiffusionSr3, InDI, JointIndi,
                                                 _schedule_tables)
from diffsplitting_b200.model.unet import UNet
from oracle import samplers_ref as S
from oracle import tiling_ref as TR
from oracle import unet_ref as U
from tests.configs import MODELS, make_opt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "diffsplit_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ds_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    lib = C.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/diffsplit_b200.h but not exported"
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert _lib.lib().ds_version() >= 100


def test_struct_layouts_match_header():
    assert C.sizeof(_lib.SamplerState) == 24
    assert C.sizeof(_lib.OpProfile) == 48
    assert C.sizeof(_lib.TensorView) == 8 + 8 + 8 + 32


def test_struct_sizes_match_a_c_compiler(tmp_path):
    """sizeof of every struct of include/diffsplit_b200.h as gcc sees it == the ctypes mirrors in _lib.py."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    pairs = [("ds_unet_desc", _lib.UNetDesc), ("ds_tensor_view", _lib.TensorView), ("ds_op_profile", _lib.OpProfile),
             ("ds_sampler_state", _lib.SamplerState), ("ds_step_args", _lib.StepArgs), ("ds_tile_norm", _lib.TileNorm),
             ("ds_psnr_args", _lib.PsnrArgs)]
    src = tmp_path / "sizes.c"
    body = "".join(f'  printf("%zu\\n", sizeof({c}));\n' for c, _ in pairs)
    src.write_text('#include <stdio.h>\n#include "diffsplit_b200.h"\nint main(void) {\n' + body + "  return 0;\n}\n")
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-I", inc, str(src), "-o", str(exe)], check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    for (cname, ct), size in zip(pairs, out):
        assert C.sizeof(ct) == int(size), (cname, C.sizeof(ct), size)


def test_time_predictor_keys_and_no_cpu_fallback():
    """TimePredictor has exactly the reference's state_dict keys (model/ddpm_modules/time_predictor.py:13-36; the oracle's
    seeded dict was loaded strictly into the reference class by make_golden) and refuses CPU tensors."""
    from diffsplitting_b200.model.time_predictor import TimePredictor
    cfg = U.make_cfg("ddpm", 1, 1, 16, 16, (1, 2, 4, 8), (), 1, 64, with_time_emb=False)
    m = TimePredictor(in_channel=1, out_channel=1, inner_channel=16, norm_groups=16, channel_mults=(1, 2, 4, 8), attn_res=(),
                      res_blocks=1, dropout=0.2, image_size=64)
    sd = U.time_predictor_state_dict(cfg, seed=1)
    assert set(m.state_dict().keys()) == set(sd.keys())
    m.load_state_dict(sd, strict=True)
    assert not m.unet.with_time_emb
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 1, 64, 64))


def test_evaluate_host_logic():
    from diffsplitting_b200.evaluate import _clamp_t, mixed_inputs
    t = torch.arange(2 * 2 * 3 * 3, dtype=torch.float32).reshape(2, 2, 3, 3)
    a, b = mixed_inputs(t, 0.25)
    assert torch.equal(a, t[:, :1] * 0.75 + t[:, 1:2] * 0.25) and torch.equal(b, t[:, 1:2] * 0.75 + t[:, :1] * 0.25)
    assert _clamp_t(1e-9) == 0.0 and _clamp_t(1.5) == 1.0 and _clamp_t(0.3) == 0.3


def test_error_reporting_without_device():
    d = _lib.UNetDesc()
    d.variant = 7
    h = C.c_void_p()
    rc = _lib.lib().ds_unet_create(C.byref(d), C.byref(h))
    assert rc == -1 and b"variant" in _lib.lib().ds_last_error()
    with pytest.raises(RuntimeError, match="variant"):
        _lib.check(rc)


def _build(cfg):
    return UNet(in_channel=cfg["in_channel"], out_channel=cfg["out_channel"], inner_channel=cfg["inner_channel"],
                norm_groups=cfg["norm_groups"], channel_mults=cfg["channel_mults"], attn_res=cfg["attn_res"],
                res_blocks=cfg["res_blocks"], image_size=cfg["image_size"], variant=cfg["variant"])


def test_state_dict_keys_match_reference(gold_dir):
    rec = json.load(open(os.path.join(gold_dir, "state_dict_keys.json")))
    n = 0
    for name, ent in rec.items():
        if name == "sr3_sampler_keys":
            continue
        cfg = ent["cfg"]
        net = _build(cfg)
        mine = {k: list(v.shape) for k, v in net.state_dict().items()}
        assert list(mine.keys()) == list(ent["keys"].keys()) or set(mine) == set(ent["keys"]), name
        assert mine == ent["keys"], name
        # reference-format checkpoints load strictly
        sd = U.random_state_dict(U.make_cfg(cfg["variant"], cfg["in_channel"], cfg["out_channel"], cfg["inner_channel"],
                                            cfg["norm_groups"], cfg["channel_mults"], cfg["attn_res"], cfg["res_blocks"],
                                            cfg["image_size"]))
        net.load_state_dict(sd, strict=True)
        n += 1
    assert n >= 6
    g = GaussianDiffusionSr3(_build(rec["sr3_attn"]["cfg"]), 16, channels=2)
    g.set_new_noise_schedule(dict(schedule="linear", n_timestep=6, linear_start=1e-4, linear_end=0.3), "cpu")
    assert [k for k in g.state_dict() if not k.startswith("denoise_fn.")] == rec["sr3_sampler_keys"]


def test_flops_and_launch_count_from_library():
    cfgD = U.make_cfg("ddpm", 1, 1, 16, 16, (1, 2, 4, 8), (), 1, 32)
    net = _build(cfgD)
    assert abs(net.flops(64, 64) - U.count_flops(cfgD, 64, 64)) < 1
    assert abs(net.flops(512, 512) / 1e9 - 70.72) < 1e-2
    assert net.launches(16, 64, 64, "fp32") > 50
    # the bf16 plan of the benchmark workload (host-side planning, no GPU needed): fused GN->conv kernels, two per-sample
    # chains, entry kernel, attention, time MLP + the statistics memset = 40; the sampler update makes it 41 per reverse step
    assert net.launches(16, 64, 64, "bf16") == 40
    assert _lib.lib().ds_unet_workspace_bytes(net._handle, 16, 64, 64, 0) > 0
    assert _lib.lib().ds_unet_workspace_bytes(net._handle, 16, 60, 64, 0) == 0      # not divisible by 8
    assert b"divisible" in _lib.lib().ds_last_error()


def test_no_cpu_fallback():
    cfg = U.make_cfg("ddpm", 1, 1, 16, 8, (1, 2), (), 1, 16)
    net = _build(cfg)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.zeros(1, 1, 16, 16), torch.zeros(1))
    indi = InDI(net, 16, channels=1, out_channel=1, conditional=False, val_schedule_opt={"n_timestep": 2})
    indi.set_new_noise_schedule({"n_timestep": 2}, "cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        indi.inference(torch.zeros(1, 1, 16, 16))
    with pytest.raises(NotImplementedError):
        indi(dict(target=torch.zeros(1, 1, 16, 16)))


@pytest.mark.parametrize("name", list(MODELS))
def test_define_G_builds_every_baseline_config(name):
    """The reference's own factory raises TypeError for sr3/ddpm; ours must build all five."""
    opt = make_opt(name)
    netG = networks.define_G(opt)
    which = MODELS[name]["which_model_G"]
    assert isinstance(netG, {"sr3": GaussianDiffusionSr3, "ddpm": GaussianDiffusionDdpm, "indi": InDI,
                             "joint_indi": JointIndi}[which])
    for meth in ("set_loss", "set_new_noise_schedule", "inference", "get_current_log", "state_dict", "load_state_dict"):
        assert hasattr(netG, meth), meth
    netG.set_new_noise_schedule(opt["model"]["beta_schedule"]["val"], "cpu")
    if which == "joint_indi":
        assert netG.indi1.num_timesteps == 3 and netG.indi1.denoise_fn is not netG.indi2.denoise_fn
        assert {"alpha_param", "offset_param", "scale_param"} <= set(dict(netG.named_parameters()))
    elif which == "indi":
        assert netG.num_timesteps == 3 and netG.e == 0.01
    else:
        assert netG.num_timesteps == opt["model"]["beta_schedule"]["val"]["n_timestep"]
        assert netG.betas.shape == (netG.num_timesteps,)
    if name == "sr_sr3_16_128":
        assert netG.denoise_fn.norm_groups == 32          # networks.py:95-96 default


def test_define_G_orthogonal_init_in_train_phase():
    netG = networks.define_G(make_opt("splitting_hagen_indi_single_ch", phase="train"))
    w = netG.denoise_fn.downs._modules["1"].res_block.block1.block._modules["3"].weight
    m = w.reshape(w.shape[0], -1)
    assert torch.allclose(m @ m.t(), torch.eye(m.shape[0]), atol=1e-4)
    assert float(netG.denoise_fn.downs._modules["0"].bias.detach().abs().max()) == 0.0


def test_create_model_requires_gpu_ids():
    from diffsplitting_b200.model import create_model
    with pytest.raises(RuntimeError, match="no CPU path"):
        create_model(make_opt("splitting_hagen_indi_single_ch", gpu_ids=None))


def test_sampler_tables_match_oracle():
    so = dict(schedule="linear", n_timestep=50, linear_start=1e-6, linear_end=1e-2)
    tabs, sacp = _schedule_tables(so)
    ref = S.schedule_tables(so)
    for k, v in tabs.items():
        assert np.array_equal(v, ref[k]), k
    assert np.array_equal(sacp, ref["sqrt_alphas_cumprod_prev"])
    g = GaussianDiffusionSr3(None, 8, channels=2)
    g.set_new_noise_schedule(so, "cpu")
    coef, ttab = g._coef()
    assert coef.shape == (50, 5) and ttab.shape == (51,)
    assert float(coef[0, 0]) == float(np.float32(ref["sqrt_recip_alphas_cumprod"][49]))
    assert float(coef[-1, 4]) == 0.0 and float(coef[0, 4]) > 0
    assert float(ttab[0]) == float(np.float32(ref["sqrt_alphas_cumprod_prev"][50]))
    assert float(ttab[49]) == float(np.float32(ref["sqrt_alphas_cumprod_prev"][1]))
    # InDI rows: same fp32 expressions as indi.py:62-69
    indi = InDI(None, 8, channels=1, out_channel=1, conditional=False, val_schedule_opt={"n_timestep": 4})
    coef, ttab = indi._tables(4, 0.5)
    cur, delta = 0.5, 0.5 / 4
    for k in range(4):
        t32 = torch.Tensor([cur])
        assert float(coef[k, 2]) == float((delta / t32)[0]) and float(coef[k, 3]) == float((1 - delta / t32)[0])
        assert float(coef[k, 4]) == float((0.01 * (t32 - delta))[0]) and float(ttab[k]) == float(t32[0])
        cur -= delta


SHAPES = [((5, 512, 512), (1, 128, 128), (1, 256, 256)), ((10, 2048, 2048), (1, 256, 256), (1, 512, 512)),
          ((3, 100, 130), (1, 16, 16), (1, 32, 32)), ((2, 64, 64), (1, 64, 64), (1, 64, 64)),
          ((1, 33, 47), (1, 8, 4), (1, 16, 12)), ((4, 70, 70), (1, 32, 32), (1, 64, 64)),
          ((10, 2048, 2048), (1, 32, 32), (1, 64, 64))]


@pytest.mark.parametrize("mode", [TilingMode.TrimBoundary, TilingMode.PadBoundary, TilingMode.ShiftBoundary])
def test_tile_manager_and_native_tables_match_oracle(mode, gold_dir):
    for data, grid, patch in SHAPES:
        tg = TR.TileGrid(data, grid, patch, mode)
        m = TileIndexManager(data, grid, patch, mode)
        assert m.total_grid_count() == tg.total
        counts, total = m.native_counts()
        assert counts == tg.counts and total == tg.total
        tab = m.patch_locations()
        assert np.array_equal(tab, tg.patch_table())                      # C ABI, full index range
        step = max(1, tg.total // 300)
        for i in list(range(0, tg.total, step)) + [tg.total - 1]:
            assert tuple(int(v) for v in m.get_patch_location_from_dataset_idx(i)) == tg.patch_location(i)
            loc = m.get_location_from_dataset_idx(i)
            assert loc == tg.grid_location(i)
            if mode != TilingMode.PadBoundary:
                assert m.get_dataset_idx_from_grid_location(loc) == i     # the reference's __main__ round trip
        assert np.array_equal(m.patch_locations(3, 2), tab[3:5]) if tg.total >= 5 else True
    g = np.load(os.path.join(gold_dir, "tiling.npz"))
    for key in g.files:
        if key.startswith("shape_") and int(key.split("_")[-1]) == mode:
            data, grid, patch = (tuple(int(v) for v in row) for row in g[key])
            assert np.array_equal(TileIndexManager(data, grid, patch, mode).patch_locations(), g[key.replace("shape_", "tab_")])


def test_tile_manager_5d_round_trip_and_validation():
    m = TileIndexManager((5, 5, 64, 64, 2), (1, 1, 8, 8, 2), (1, 3, 16, 16, 2), TilingMode.ShiftBoundary)
    for i in range(m.total_grid_count()):
        assert m.get_dataset_idx_from_grid_location(m.get_location_from_dataset_idx(i)) == i
    assert m.on_boundary(40, 0) in (True, False)
    with pytest.raises(ValueError):
        TileIndexManager((4, 32, 32), (1, 16, 16), (1, 8, 8), TilingMode.ShiftBoundary)
    with pytest.raises(ValueError):
        TileIndexManager((4, 32, 32), (1, 16, 16), (1, 17, 16), TilingMode.ShiftBoundary)
    rc = _lib.lib().ds_tile_counts(_lib.i3((4, 8, 8)), _lib.i3((1, 4, 4)), _lib.i3((1, 16, 16)), 2, (C.c_int32 * 3)(), None)
    assert rc == -1


def test_predtiler_shim():
    class Base:
        def __init__(self, n):
            self.n = n

        def __len__(self):
            return self.n

    mgr = get_tile_manager((10, 2048, 2048), (1, 256, 256), (1, 512, 512))
    ds = get_tiling_dataset(Base, mgr)(3)
    assert len(ds) == 490 and ds.tile_manager is mgr
    assert tuple(int(v) for v in ds.patch_location(8)) == (0, 256, 256)


def test_install_aliases():
    import sys

    import diffsplitting_b200 as dsb
    saved = {k: sys.modules.get(k) for k in ("model", "model.networks", "data.tile_stitcher", "predtiler", "predtiler.dataset",
                                             "model.model", "model.base_model", "data.tiling_manager", "model.ddpm_modules",
                                             "model.ddpm_modules.time_predictor")}
    try:
        dsb.install()
        import model as M
        from predtiler.dataset import get_tile_manager as gtm
        assert M.create_model is dsb.model.create_model and gtm is get_tile_manager
        from model.ddpm_modules.time_predictor import TimePredictor as TP
        assert TP is dsb.model.time_predictor.TimePredictor
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_indi_last_step_rounding_is_tolerated():
    """The reference's `assert delta_t <= t_cur` (indi.py:64) fires on the last step for T=1000, t_start=1.0 because
    cur_t is a python float decremented T times; the port tolerates exactly that rounding and nothing more."""
    indi = InDI(None, 8, channels=1, out_channel=1, conditional=False, val_schedule_opt={"n_timestep": 4})
    for T, ts in ((1000, 1.0), (20, 0.5), (50, 1.0), (3, 0.5), (16, 0.5)):
        coef, ttab = indi._tables(T, ts)
        assert coef.shape == (T, 5) and torch.isfinite(coef).all()
    d, cur = 1.0 / 1000, 1.0
    for _ in range(999):
        cur -= d
    assert not d <= cur                      # the reference raises here
    with pytest.raises(AssertionError):
        S.indi_one_step(lambda x, t: x, torch.zeros(1, 1, 2, 2), d, cur, 0.01, torch.zeros(1, 1, 2, 2))
    S.indi_one_step(lambda x, t: x, torch.zeros(1, 1, 2, 2), d, cur, 0.01, torch.zeros(1, 1, 2, 2), strict=False)
    with pytest.raises(AssertionError):
        indi.inference_one_step(torch.zeros(1, 1, 2, 2), 0.2, 0.1)       # a genuinely too large delta still asserts


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (no GPU needed): one JSON line with the contract's keys, the unmodified reference from
    baseline/_ref when it is installed, and the same `config` dict our own arm prints for that workload."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "cifar10_ddpm_32_b1_T50",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "unet_denoising_steps_per_sec" and d["unit"] == "steps/s" and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    from baseline import refshim
    assert d["cpu_baseline"]["kind"] == ("reference" if refshim.available() else "port")
    sys.path.insert(0, root)
    import bench
    assert d["config"] == bench.config_dict("cifar10_ddpm_32_b1_T50", bench.WORKLOADS["cifar10_ddpm_32_b1_T50"])
