"""Drives the UNMODIFIED reference front end (baseline/_ref) against this package's drop-in modules, in a process of its
own (the reference's top-level packages are called ``model`` / ``data`` / ``core``).

    python tests/boundary_driver.py host          # no GPU needed
    python tests/boundary_driver.py gpu <config>  # one of the reference's config/*.json

What it proves ("split.py / infer.py / eval.py run unchanged"):
  * the reference's own ``core/logger.parse`` + ``dict_to_nonedict`` on its own JSON files produce the ``opt`` that
    ``Model.create_model`` consumes here - after ``diffsplitting_b200.install()`` aliased ``model`` to this package;
  * ``import split`` (the reference's script, unmodified) resolves ``predtiler.dataset``, ``model`` and ``data.tile_stitcher``
    to this package; ``split.get_datasets(opt, tiled_pred=True)`` with ``load_data`` stubbed (as the reference's own
    tests/test_tiling_setup.py does) yields the 490-tile dataset of the notebook;
  * gpu: ``create_model(opt)`` -> ``feed_data`` -> ``test`` -> ``get_current_visuals`` for the config, and the reference's
    ``tests/test_tiling_setup.py::test_stich_prediction`` construction through this package's stitcher.
Prints one JSON object.
"""
import argparse
import json
import os
import sys
import tempfile
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def parse_opt(cfg_name, phase="val"):
    import core.logger as Logger                      # the reference's, from baseline/_ref
    from baseline import refshim
    root = tempfile.mkdtemp(prefix="dsb_boundary_")
    args = types.SimpleNamespace(phase=phase, config=os.path.join(refshim.REF, "config", cfg_name), gpu_ids="0",
                                 enable_wandb=False, debug=False, rootdir=root, log_wandb_ckpt=False, log_eval=False)
    visible = os.environ.get("CUDA_VISIBLE_DEVICES")
    try:
        opt = Logger.parse(args)
    except KeyError as e:
        # the reference's get_model_name() reads model.loss_type, which its own sr_sr3_* / sample_* JSONs do not have: its
        # parse() cannot load them (the splitting fork broke the original SR3 configs).  Same JSON through its load_json.
        opt = Logger.load_json(args.config)
        opt["phase"], opt["gpu_ids"], opt["distributed"], opt["enable_wandb"] = phase, [0], False, False
        opt["path"]["checkpoint"] = os.path.join(root, "checkpoint")
        opt["reference_parse_error"] = repr(e)
    opt = Logger.dict_to_nonedict(opt)
    if visible is None:                                # parse() exports CUDA_VISIBLE_DEVICES=0: harmless on the 1-GPU box
        os.environ.pop("CUDA_VISIBLE_DEVICES", None)
    else:
        os.environ["CUDA_VISIBLE_DEVICES"] = visible
    return opt


def fake_load_data(*a, **k):
    import numpy as np
    n, H, W = 10, 2048, 2048
    base = np.arange(n * H * W, dtype=np.float32).reshape(n, H, W)
    return {0: base, 1: base + 0.5}


def host_checks():
    import numpy as np
    from baseline import refshim
    refshim.activate()
    import diffsplitting_b200 as dsb
    dsb.install()
    import split                                        # the reference's split.py, unmodified
    import model as Model
    import predtiler.dataset as PT
    import data.tile_stitcher as TS
    out = {"model_is_ours": Model.__name__.startswith("diffsplitting_b200"),
           "predtiler_is_ours": PT.get_tile_manager.__module__.startswith("diffsplitting_b200"),
           "stitcher_is_ours": TS.stitch_predictions.__module__.startswith("diffsplitting_b200"),
           "split_file": os.path.relpath(split.__file__, ROOT)}
    import data.split_dataset as SD
    SD.load_data = fake_load_data
    opt = parse_opt("splitting_hagen_indi_joint.json")
    train_set, val_set = split.get_datasets(opt, tiled_pred=True)
    out["tiled_len"] = len(val_set)
    out["patch_location_0"] = [int(v) for v in val_set.patch_location(0)]
    out["patch_location_last"] = [int(v) for v in val_set.patch_location(len(val_set) - 1)]
    item = val_set[7]
    out["item_shapes"] = {k: list(np.asarray(v).shape) for k, v in item.items() if hasattr(v, "shape")}
    out["configs_parsed"] = {}
    for cfg in ("splitting_cifar10.json", "splitting_hagen_indi_single_ch.json", "splitting_hagen_indi_joint.json",
                "sr_sr3_16_128.json", "sr_sr3_64_512.json", "splitting.json"):
        o = parse_opt(cfg)
        out["configs_parsed"][cfg] = [o["model"]["which_model_G"], o["reference_parse_error"]]
    return out, val_set


def gpu_checks(cfg_name):
    import numpy as np
    import torch
    full = cfg_name == "splitting_hagen_indi_joint.json"          # the dataset / tiling half runs once, with the joint config
    if full:
        out, val_set = host_checks()
    else:
        from baseline import refshim
        refshim.activate()
        import diffsplitting_b200 as dsb
        dsb.install()
        out = {}
    import model as Model
    opt = parse_opt(cfg_name)
    m = opt["model"]
    out["config"] = cfg_name
    diffusion = Model.create_model(opt)
    diffusion.set_new_noise_schedule(dict(opt["model"]["beta_schedule"]["val"], n_timestep=3), schedule_phase="val")
    size = 64
    g = torch.Generator().manual_seed(0)
    cin = m["unet"]["in_channel"] - (m["diffusion"]["channels"] if m["diffusion"]["conditional"] else 0)
    data = {"input": torch.rand((1, cin, size, size), generator=g) * 2 - 1,
            "target": torch.rand((1, max(2, m["diffusion"]["channels"]), size, size), generator=g) * 2 - 1}
    diffusion.feed_data(data)
    diffusion.test(continuous=False)
    vis = diffusion.get_current_visuals()
    out["visuals"] = {k: [list(v.shape), str(v.dtype), str(v.device), bool(torch.isfinite(v).all())] for k, v in vis.items()}
    out["netG"] = type(diffusion.netG).__module__ + "." + type(diffusion.netG).__name__
    if not full:
        return out
    # the reference's own tiling test, through this package's stitcher (data.tile_stitcher is aliased)
    import data.split_dataset as SD
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref", "tests"))
    import test_tiling_setup as T

    class MP:
        def setattr(self, name, value):
            mod, attr = name.rsplit(".", 1)
            setattr(sys.modules[mod], attr, value)
    T.test_stich_prediction(MP())
    out["reference_tiling_test"] = "passed"
    # 490-tile stitch identity through the dataset split.get_datasets built
    SD.load_data = fake_load_data
    from data.tile_stitcher import stitch_predictions
    preds = np.stack([val_set[i]["target"] for i in range(0, len(val_set))]).astype(np.float32)
    st = stitch_predictions(preds, val_set.tile_manager)
    nd = val_set.get_normalization_dict()
    frames = fake_load_data()
    ref = np.stack([(frames[c] - nd["mean_target"].reshape(-1)[c]) / nd["std_target"].reshape(-1)[c] for c in range(2)], axis=-1).astype(np.float32)
    out["stitch_490_max_abs_err"] = float(np.abs(st - ref).max())
    out["stitch_490_shape"] = list(st.shape)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["host", "gpu"])
    ap.add_argument("config", nargs="?", default="splitting_hagen_indi_joint.json")
    a = ap.parse_args()
    saved = os.dup(1)
    os.dup2(2, 1)
    res = host_checks()[0] if a.mode == "host" else gpu_checks(a.config)
    sys.stdout.flush()
    os.dup2(saved, 1)
    print(json.dumps(res))
