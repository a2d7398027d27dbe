"""Time predictor (reference: model/ddpm_modules/time_predictor.py:5-45): a ``ddpm_modules.UNet`` without time
embedding, ReLU, a 7x7-conv sigmoid foreground mask and a masked mean, giving one scalar per image (the ``t_float_start``
the InDI loop is started from).  Same constructor arguments, attribute names (``unet``, ``foreground_mask.layer``) and
``state_dict`` keys as the reference; the UNet runs through ``ds_unet_forward`` and the tail (mask conv + sigmoid + ReLU +
product + both sums + division) is one fused pass (``ds_time_head_f32``).  Inference only; no CPU fallback.
"""
import torch
import torch.nn as nn

from .. import _lib
from .unet import UNet


class ForegroundMask(nn.Module):
    def __init__(self, in_channel, out_channel):
        super().__init__()
        self.layer = nn.Conv2d(in_channel, out_channel, 7, padding=3)        # parameter holder (PyTorch default init)

    def forward(self, x):
        raise RuntimeError("the foreground mask is evaluated inside TimePredictor.forward (fused kernel)")


class TimePredictor(nn.Module):
    def __init__(self, in_channel=6, out_channel=3, inner_channel=32, norm_groups=32, channel_mults=(1, 2, 4, 8, 8),
                 attn_res=(8,), res_blocks=3, dropout=0, image_size=128, precision=None):
        super().__init__()
        self.unet = UNet(in_channel=in_channel, out_channel=out_channel, inner_channel=inner_channel,
                         norm_groups=norm_groups, channel_mults=channel_mults, attn_res=attn_res, res_blocks=res_blocks,
                         dropout=dropout, image_size=image_size, with_time_emb=False, variant="ddpm", precision=precision)
        self.foreground_mask = ForegroundMask(in_channel, out_channel)

    def forward(self, x):
        _lib.require_cuda(x, "TimePredictor input")
        x = x.float().contiguous()
        out = self.unet(x, None)
        w, b = self.foreground_mask.layer.weight, self.foreground_mask.layer.bias
        _lib.require_cuda(w, "TimePredictor parameter")
        B, cin, H, W = x.shape
        res = torch.empty((B,), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            lib = _lib.lib()
            nbytes = lib.ds_time_head_workspace_bytes(B, H, W)
            ws = torch.empty((nbytes,), dtype=torch.uint8, device=x.device)
            _lib.check(lib.ds_time_head_f32(x.data_ptr(), out.data_ptr(), w.detach().float().contiguous().data_ptr(),
                                            b.detach().float().contiguous().data_ptr(), B, cin, out.shape[1], H, W,
                                            res.data_ptr(), ws.data_ptr(), nbytes, _lib.stream_ptr()))
        return res
