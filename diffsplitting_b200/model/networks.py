"""Generator factory with the reference's signature: ``define_G(opt) -> netG`` (model/networks.py:91-180).

Differences from the reference, both required for "split.py / infer.py / eval.py run unchanged":
  * every sampler accepts the full kwarg set the factory passes (the reference sr3/ddpm classes raise
    ``TypeError`` on ``out_channel`` / ``lr_reduction`` / ``val_schedule_opt``);
  * ``nn.DataParallel`` is never applied: inference in the reference already bypasses it
    (model/model.py:66-70) and multi-GPU here is one process per GPU (diffsplitting_b200/parallel.py).
"""
import logging

import torch
from torch.nn import init

from .samplers import GaussianDiffusionDdpm, GaussianDiffusionSr3, InDI, JointIndi
from .unet import UNet

logger = logging.getLogger("base")


def init_weights(net, init_type="orthogonal", scale=1, std=0.02):
    """Re-initialise conv / linear weights the way the reference does in train phase (networks.py:21-82)."""
    logger.info("Initialization method [{:s}]".format(init_type))
    with torch.no_grad():
        for name, p in net.named_parameters():
            is_norm = ".block.0." in name or ".norm." in name
            if p.dim() < 2 or is_norm:
                if name.endswith("bias") and not is_norm and p.dim() == 1:
                    p.zero_()
                continue
            if init_type == "orthogonal":
                init.orthogonal_(p, gain=1)
            elif init_type == "kaiming":
                init.kaiming_normal_(p, a=0, mode="fan_in")
                p.mul_(scale)
            elif init_type == "normal":
                init.normal_(p, 0.0, std)
            else:
                raise NotImplementedError("initialization method [{:s}] not implemented".format(init_type))


def _unet(model_opt, variant):
    u = model_opt["unet"]
    net = UNet(in_channel=u["in_channel"], out_channel=u["out_channel"], norm_groups=u["norm_groups"],
               inner_channel=u["inner_channel"], channel_mults=u["channel_multiplier"], attn_res=u["attn_res"],
               res_blocks=u["res_blocks"], dropout=u["dropout"], image_size=model_opt["diffusion"]["image_size"],
               variant=variant)
    net.reset_parameters()
    return net


def define_G(opt):
    model_opt = opt["model"]
    model_kwargs = {}
    if ("norm_groups" not in model_opt["unet"]) or model_opt["unet"]["norm_groups"] is None:
        model_opt["unet"]["norm_groups"] = 32
    which = model_opt["which_model_G"]
    table = {"ddpm": (GaussianDiffusionDdpm, "ddpm"), "sr3": (GaussianDiffusionSr3, "sr3"),
             "indi": (InDI, "ddpm"), "joint_indi": (JointIndi, "ddpm")}
    if which not in table:
        raise NotImplementedError("Generator model [{:s}] not recognized".format(str(which)))
    netG_class, variant = table[which]
    if which != "joint_indi":
        model = _unet(model_opt, variant)
    else:
        model_kwargs["allow_full_translation"] = model_opt.get("allow_full_translation", False) or False
        model_kwargs["w_input_loss"] = model_opt["w_input_loss"]
        model_kwargs["denoise_fn_ch1"] = _unet(model_opt, variant)
        model_kwargs["denoise_fn_ch2"] = _unet(model_opt, variant)
        model = None
    netG = netG_class(model, image_size=model_opt["diffusion"]["image_size"], channels=model_opt["diffusion"]["channels"],
                      loss_type=model_opt["loss_type"], out_channel=model_opt["unet"]["out_channel"],
                      lr_reduction=model_opt["lr_reduction"], conditional=model_opt["diffusion"]["conditional"],
                      schedule_opt=model_opt["beta_schedule"]["train"], val_schedule_opt=model_opt["beta_schedule"]["val"],
                      **model_kwargs)
    if opt["phase"] == "train":
        init_weights(netG, init_type="orthogonal")
    return netG
