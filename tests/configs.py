"""`model` sections of the reference's shipped configs (config/*.json in the reference repo), transcribed as
data so the tests do not need the reference checkout.  Only the keys define_G reads (networks.py:93-170)."""

_SCHED = lambda T: {"schedule": "linear", "n_timestep": T, "linear_start": 1e-6, "linear_end": 1e-2}


def _model(which, in_ch, out_ch, inner, groups, mults, attn_res, res_blocks, dropout, image_size, channels,
           conditional, train_T, val_T, **extra):
    m = {"which_model_G": which, "finetune_norm": False, "loss_type": "l1", "lr_reduction": None,
         "unet": {"in_channel": in_ch, "out_channel": out_ch, "inner_channel": inner, "norm_groups": groups,
                  "channel_multiplier": list(mults), "attn_res": list(attn_res), "res_blocks": res_blocks,
                  "dropout": dropout},
         "beta_schedule": {"train": _SCHED(train_T), "val": _SCHED(val_T)},
         "diffusion": {"image_size": image_size, "channels": channels, "conditional": conditional}}
    m.update(extra)
    return m


MODELS = {
    "splitting_cifar10": _model("ddpm", 9, 6, 16, 16, (1, 2, 4, 8), (), 1, 0, 32, 6, True, 3, 3),
    "splitting_hagen_indi_single_ch": _model("indi", 1, 1, 16, 16, (1, 2, 4, 8), (), 1, 0, 32, 1, False, 20, 3),
    "splitting_hagen_indi_joint": _model("joint_indi", 1, 1, 16, 16, (1, 2, 4, 8), (), 1, 0, 32, 1, False, 2000, 3,
                                         w_input_loss=0.0),
    "sr_sr3_16_128": _model("sr3", 6, 3, 64, None, (1, 2, 4, 8, 8), (16,), 2, 0.2, 128, 3, True, 2000, 2000),
    "sr_sr3_64_512": _model("sr3", 6, 3, 64, 16, (1, 2, 4, 8, 16), (), 1, 0, 512, 3, True, 2000, 2000),
    "splitting": _model("sr3", 3, 2, 16, 16, (1, 2, 4, 8), (), 1, 0, 512, 2, True, 2000, 2000),
}


class NoneDict(dict):
    def __missing__(self, key):
        return None


def to_nonedict(o):
    if isinstance(o, dict):
        return NoneDict(**{k: to_nonedict(v) for k, v in o.items()})
    if isinstance(o, list):
        return [to_nonedict(v) for v in o]
    return o


def make_opt(name, phase="val", gpu_ids=(0,)):
    import copy
    return to_nonedict({"name": name, "phase": phase, "gpu_ids": list(gpu_ids) if gpu_ids is not None else None,
                        "distributed": False, "model": copy.deepcopy(MODELS[name]),
                        "path": {"resume_state": None, "checkpoint": "/tmp"},
                        "train": {"optimizer": {"type": "adam", "lr": 1e-4}}})
