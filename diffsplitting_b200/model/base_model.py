"""BaseModel of the reference wrapper layer (model/base_model.py:6-48)."""
import torch


class BaseModel:
    def __init__(self, opt):
        self.opt = opt
        if opt["gpu_ids"] is None:
            raise RuntimeError("diffsplit_b200 has no CPU path: run with -gpu <id> (opt['gpu_ids'] is None)")
        self.device = torch.device("cuda")
        self.begin_step = 0
        self.begin_epoch = 0

    def feed_data(self, data):
        pass

    def optimize_parameters(self):
        pass

    def get_current_visuals(self):
        pass

    def get_current_losses(self):
        pass

    def print_network(self):
        pass

    def set_device(self, x):
        if isinstance(x, dict):
            for key, item in x.items():
                if item is not None:
                    x[key] = item.to(self.device)
        elif isinstance(x, list):
            x = [item.to(self.device) if item is not None else None for item in x]
        else:
            x = x.to(self.device)
        return x

    def get_network_description(self, network):
        s = str(network)
        n = sum(p.numel() for p in network.parameters())
        return s, n
