"""``DDPM`` wrapper with the reference's public surface (model/model.py:12-173): owns ``netG``, feeds data,
runs ``test`` / ``sample`` through the native sampling path, returns CPU fp32 visuals, saves / loads
reference-format checkpoints.  Optimisation (``optimize_parameters``) is outside the sampling hot path."""
import logging
import os
from collections import OrderedDict

import torch

from . import networks
from .base_model import BaseModel

logger = logging.getLogger("base")


class DDPM(BaseModel):
    def __init__(self, opt):
        super().__init__(opt)
        self.netG = self.set_device(networks.define_G(opt))
        self.schedule_phase = None
        self.set_loss()
        self.set_new_noise_schedule(opt["model"]["beta_schedule"]["train"], schedule_phase="train")
        if self.opt["phase"] == "train":
            self.netG.train()
            if opt["model"]["finetune_norm"]:
                optim_params = []
                for k, v in self.netG.named_parameters():
                    v.requires_grad = False
                    if k.find("transformer") >= 0:
                        v.requires_grad = True
                        v.data.zero_()
                        optim_params.append(v)
            else:
                optim_params = list(self.netG.parameters())
            if optim_params:
                self.optG = torch.optim.Adam(optim_params, lr=opt["train"]["optimizer"]["lr"])
            self.log_dict = OrderedDict()
        self.load_network()

    def feed_data(self, data):
        self.data = self.set_device(data)

    def optimize_parameters(self):
        raise NotImplementedError("diffsplit_b200 covers the sampling hot path; train with the reference and load the "
                                  "checkpoint here (path.resume_state)")

    def test(self, continuous=False, clip_denoised=True):
        self.netG.eval()
        with torch.no_grad():
            self.prediction = self.netG.inference(self.data["input"], continuous=continuous)
        self.netG.train()

    def sample(self, batch_size=1, continous=False):
        self.netG.eval()
        with torch.no_grad():
            self.prediction = self.netG.sample(batch_size, continous)
        self.netG.train()

    def set_loss(self):
        self.netG.set_loss(self.device)

    def set_new_noise_schedule(self, schedule_opt, schedule_phase="train"):
        if self.schedule_phase is None or self.schedule_phase != schedule_phase:
            self.schedule_phase = schedule_phase
            self.netG.set_new_noise_schedule(schedule_opt, self.device)

    def get_current_log(self):
        return self.log_dict

    def get_current_visuals(self, need_LR=True, sample=False):
        out = OrderedDict()
        if sample:
            out["SAM"] = self.prediction.detach().float().cpu()
        else:
            out["prediction"] = self.prediction.detach().float().cpu()
            out["input"] = self.data["input"].detach().float().cpu()
            out["target"] = self.data["target"].detach().float().cpu()
        return out

    def print_network(self):
        s, n = self.get_network_description(self.netG)
        logger.info("Network G structure: {}, with parameters: {:,d}".format(self.netG.__class__.__name__, n))
        logger.info(s)

    def save_network(self, epoch, iter_step):
        gen_path = os.path.join(self.opt["path"]["checkpoint"], "I{}_E{}_gen.pth".format(iter_step, epoch))
        opt_path = os.path.join(self.opt["path"]["checkpoint"], "I{}_E{}_opt.pth".format(iter_step, epoch))
        state_dict = {k: v.cpu() for k, v in self.netG.state_dict().items()}
        torch.save(state_dict, gen_path)
        opt_state = {"epoch": epoch, "iter": iter_step, "scheduler": None, "optimizer": None}
        if hasattr(self, "optG"):
            opt_state["optimizer"] = self.optG.state_dict()
        torch.save(opt_state, opt_path)
        logger.info("Saved model in [{:s}] ...".format(gen_path))

    def load_network(self):
        load_path = self.opt["path"]["resume_state"]
        if load_path is None:
            return
        logger.info("Loading pretrained model for G [{:s}] ...".format(load_path))
        gen_path = "{}_gen.pth".format(load_path)
        opt_path = "{}_opt.pth".format(load_path)
        self.netG.load_state_dict(torch.load(gen_path, map_location=self.device),
                                  strict=(not self.opt["model"]["finetune_norm"]))
        if self.opt["phase"] == "train" and os.path.exists(opt_path):
            opt = torch.load(opt_path, map_location="cpu")
            if hasattr(self, "optG") and opt.get("optimizer") is not None:
                self.optG.load_state_dict(opt["optimizer"])
            self.begin_step = opt["iter"]
            self.begin_epoch = opt["epoch"]
