"""Functional CPU restatement of the reference UNets (oracle; test infrastructure).

Follows, without sharing code with:
  * ``model/sr3_modules/unet.py:161-259``  (noise-level conditioned UNet, FiLM add)
  * ``model/ddpm_modules/unet.py:147-243`` (integer/float time conditioned UNet)

The network is described by a plain dict ``cfg`` and evaluated straight from a
``state_dict`` (reference key names), so the same weights can be pushed through
the reference module, this oracle and the CUDA library.
"""
import math
from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def make_cfg(variant: str, in_channel: int, out_channel: int, inner_channel: int,
             norm_groups: int, channel_mults: Sequence[int], attn_res: Sequence[int],
             res_blocks: int, image_size: int, with_time_emb: bool = True) -> dict:
    assert variant in ("sr3", "ddpm")
    return dict(variant=variant, in_channel=in_channel, out_channel=out_channel,
                inner_channel=inner_channel, norm_groups=norm_groups,
                channel_mults=tuple(channel_mults), attn_res=tuple(attn_res or ()),
                res_blocks=res_blocks, image_size=image_size,
                with_time_emb=with_time_emb)


def layer_table(cfg: dict) -> Dict[str, List[tuple]]:
    """Walk the constructor logic (sr3 unet.py:187-233 / ddpm unet.py:172-219)
    and emit (kind, state-dict prefix, cin, cout, has_attn) rows."""
    inner = cfg["inner_channel"]
    mults = cfg["channel_mults"]
    res_now = cfg["image_size"]
    attn_res = cfg["attn_res"]
    downs = [("conv", "downs.0", cfg["in_channel"], inner, False)]
    skip_ch = [inner]
    ch = inner
    for lvl, m in enumerate(mults):
        cout = inner * m
        for _ in range(cfg["res_blocks"]):
            downs.append(("res", f"downs.{len(downs)}", ch, cout, res_now in attn_res))
            ch = cout
            skip_ch.append(ch)
        if lvl != len(mults) - 1:
            downs.append(("down", f"downs.{len(downs)}", ch, ch, False))
            skip_ch.append(ch)
            res_now //= 2
    mid = [("res", "mid.0", ch, ch, True), ("res", "mid.1", ch, ch, False)]
    ups = []
    for lvl in reversed(range(len(mults))):
        cout = inner * mults[lvl]
        for _ in range(cfg["res_blocks"] + 1):
            ups.append(("res", f"ups.{len(ups)}", ch + skip_ch.pop(), cout, res_now in attn_res))
            ch = cout
        if lvl >= 1:
            ups.append(("up", f"ups.{len(ups)}", ch, ch, False))
            res_now *= 2
    assert not skip_ch
    return dict(downs=downs, mid=mid, ups=ups, final=("final", "final_conv", ch, cfg["out_channel"], False))


def swish(x: Tensor) -> Tensor:
    return x * torch.sigmoid(x)


def noise_level_encoding(level: Tensor, dim: int) -> Tensor:
    """sr3 unet.py:23-31.  level: (B,1) -> (B,1,dim)."""
    half = dim // 2
    step = torch.arange(half, dtype=level.dtype, device=level.device) / half
    enc = level.unsqueeze(1) * torch.exp(-math.log(1e4) * step.unsqueeze(0))
    return torch.cat([enc.sin(), enc.cos()], dim=-1)


def time_encoding(t: Tensor, dim: int) -> Tensor:
    """ddpm unet.py:19-34.  t: any shape -> (*shape, dim)."""
    inv_freq = torch.exp(torch.arange(0, dim, 2, dtype=torch.float32, device=t.device) * (-math.log(10000) / dim))
    s = torch.outer(t.reshape(-1).float(), inv_freq)
    return torch.cat([s.sin(), s.cos()], dim=-1).reshape(*t.shape, dim)


def _embedding(sd, cfg, time: Tensor, pfx: str):
    inner = cfg["inner_channel"]
    if cfg["variant"] == "sr3":
        e = noise_level_encoding(time, inner)                       # (B,1,dim)
        key = "noise_level_mlp"
    else:
        e = time_encoding(time, inner)                              # (B,dim) or (1,dim)
        key = "time_mlp"
    e = F.linear(e, sd[f"{pfx}{key}.1.weight"], sd[f"{pfx}{key}.1.bias"])
    e = swish(e)
    return F.linear(e, sd[f"{pfx}{key}.3.weight"], sd[f"{pfx}{key}.3.bias"])


def _gn_swish_conv(sd, p, x, groups):
    """Block: GroupNorm -> Swish -> (Dropout: identity in eval) -> Conv3x3 (unet.py:80-91)."""
    h = F.group_norm(x, groups, sd[p + ".block.0.weight"], sd[p + ".block.0.bias"], eps=1e-5)
    return F.conv2d(swish(h), sd[p + ".block.3.weight"], sd[p + ".block.3.bias"], padding=1)


def _film_bias(sd, cfg, p, emb):
    """Per-block conditioning vector (B or 1, Cout).
    sr3: FeatureWiseAffine without affine level (unet.py:42-50);
    ddpm: Swish -> Linear (ddpm unet.py:81-84,93-94)."""
    if emb is None:
        return None
    if cfg["variant"] == "sr3":
        v = F.linear(emb, sd[p + ".noise_func.noise_func.0.weight"], sd[p + ".noise_func.noise_func.0.bias"])
        return v.reshape(v.shape[0], -1)
    v = F.linear(swish(emb), sd[p + ".mlp.1.weight"], sd[p + ".mlp.1.bias"])
    return v.reshape(-1, v.shape[-1])


def _attention(sd, p, x, groups):
    """Single-head self-attention over H*W tokens, head_dim = C (unet.py:113-142)."""
    b, c, hh, ww = x.shape
    n = F.group_norm(x, groups, sd[p + ".norm.weight"], sd[p + ".norm.bias"], eps=1e-5)
    qkv = F.conv2d(n, sd[p + ".qkv.weight"])
    q, k, v = qkv.reshape(b, 3, c, hh * ww).unbind(1)              # each (B,C,N)
    s = torch.einsum("bcq,bck->bqk", q, k) / math.sqrt(c)
    a = torch.softmax(s, dim=-1)
    o = torch.einsum("bqk,bck->bcq", a, v).reshape(b, c, hh, ww)
    return F.conv2d(o, sd[p + ".out.weight"], sd[p + ".out.bias"]) + x


def _res_block(sd, cfg, row, x, emb):
    _, p, cin, cout, attn = row
    g = cfg["norm_groups"]
    rb = p + ".res_block"
    h = _gn_swish_conv(sd, rb + ".block1", x, g)
    fb = _film_bias(sd, cfg, rb, emb)
    if fb is not None:
        h = h + fb[:, :, None, None]
    h = _gn_swish_conv(sd, rb + ".block2", h, g)
    if cin != cout:
        x = F.conv2d(x, sd[rb + ".res_conv.weight"], sd[rb + ".res_conv.bias"])
    h = h + x
    if attn:
        h = _attention(sd, p + ".attn", h, g)
    return h


@torch.no_grad()
def unet_forward(sd: Dict[str, Tensor], cfg: dict, x: Tensor, time, prefix: str = "",
                 taps: dict = None) -> Tensor:
    """Evaluate the UNet.  ``x`` (B,Cin,H,W); ``time`` (B,1) [sr3] / (B,) or (1,) [ddpm] / None.
    ``taps`` (optional dict) receives named intermediate activations for layer-wise debugging."""
    sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)} if prefix else sd
    tab = layer_table(cfg)
    emb = _embedding(sd, cfg, time, "") if (cfg["with_time_emb"] and time is not None) else None
    skips = []
    for row in tab["downs"]:
        kind, p = row[0], row[1]
        if kind == "conv":
            x = F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], padding=1)
        elif kind == "res":
            x = _res_block(sd, cfg, row, x, emb)
        else:  # down: 3x3 stride 2 (unet.py:68-74)
            x = F.conv2d(x, sd[p + ".conv.weight"], sd[p + ".conv.bias"], stride=2, padding=1)
        skips.append(x)
        if taps is not None:
            taps[p] = x
    for row in tab["mid"]:
        x = _res_block(sd, cfg, row, x, emb)
        if taps is not None:
            taps[row[1]] = x
    for row in tab["ups"]:
        kind, p = row[0], row[1]
        if kind == "res":
            x = _res_block(sd, cfg, row, torch.cat([x, skips.pop()], dim=1), emb)
        else:  # up: nearest x2 then 3x3 (unet.py:58-65)
            x = F.interpolate(x, scale_factor=2, mode="nearest")
            x = F.conv2d(x, sd[p + ".conv.weight"], sd[p + ".conv.bias"], padding=1)
        if taps is not None:
            taps[p] = x
    return _gn_swish_conv(sd, "final_conv", x, cfg["norm_groups"])


def time_predictor_forward(sd: Dict[str, Tensor], cfg: dict, x: Tensor) -> Tensor:
    """TimePredictor.forward (model/ddpm_modules/time_predictor.py:38-45): ``cfg`` describes the inner ddpm UNet
    (with_time_emb False), ``sd`` has the reference keys ``unet.*`` and ``foreground_mask.layer.{weight,bias}``."""
    out = F.relu(unet_forward(sd, cfg, x, None, prefix="unet."))                                    # :39-40
    attention = torch.sigmoid(F.conv2d(x, sd["foreground_mask.layer.weight"], sd["foreground_mask.layer.bias"],
                                       padding=3))                                                   # :5-12, :41
    out = (out * attention).reshape(out.shape[0], -1)                                               # :42-43
    return out.sum(dim=1) / attention.reshape(out.shape).sum(dim=1)                                 # :44


def time_predictor_state_dict(cfg: dict, seed: int = 0) -> Dict[str, Tensor]:
    sd = {"unet." + k: v for k, v in random_state_dict(cfg, seed=seed).items()}
    g = torch.Generator().manual_seed(seed + 1000)
    cin, cout = cfg["in_channel"], cfg["out_channel"]
    bound = 1.0 / math.sqrt(cin * 49)
    sd["foreground_mask.layer.weight"] = (torch.rand((cout, cin, 7, 7), generator=g) * 2 - 1) * bound
    sd["foreground_mask.layer.bias"] = (torch.rand((cout,), generator=g) * 2 - 1) * bound
    return sd


def random_state_dict(cfg: dict, seed: int = 0, scale: float = 1.0) -> Dict[str, Tensor]:
    """Seeded random weights with PyTorch-default-like fan-in scaling, generated
    independently of any nn.Module so they travel to the GPU box.  GroupNorm
    affine parameters are perturbed away from (1,0) so they are actually tested."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}

    def uni(shape, bound):
        return (torch.rand(shape, generator=g) * 2 - 1) * bound

    def conv(p, cin, cout, k, bias=True):
        bound = scale / math.sqrt(cin * k * k)
        sd[p + ".weight"] = uni((cout, cin, k, k), bound)
        if bias:
            sd[p + ".bias"] = uni((cout,), bound)

    def lin(p, cin, cout):
        bound = scale / math.sqrt(cin)
        sd[p + ".weight"] = uni((cout, cin), bound)
        sd[p + ".bias"] = uni((cout,), bound)

    def gn(p, c):
        sd[p + ".weight"] = 1.0 + 0.2 * uni((c,), 1.0)
        sd[p + ".bias"] = 0.2 * uni((c,), 1.0)

    inner = cfg["inner_channel"]
    if cfg["with_time_emb"]:
        key = "noise_level_mlp" if cfg["variant"] == "sr3" else "time_mlp"
        if cfg["variant"] == "ddpm":
            sd["time_mlp.0.inv_freq"] = torch.exp(
                torch.arange(0, inner, 2, dtype=torch.float32) * (-math.log(10000) / inner))
        lin(key + ".1", inner, inner * 4)
        lin(key + ".3", inner * 4, inner)
    tab = layer_table(cfg)
    for row in tab["downs"] + tab["mid"] + tab["ups"]:
        kind, p, cin, cout, attn = row
        if kind == "conv":
            conv(p, cin, cout, 3)
        elif kind in ("down", "up"):
            conv(p + ".conv", cin, cout, 3)
        else:
            rb = p + ".res_block"
            if cfg["with_time_emb"]:
                if cfg["variant"] == "sr3":
                    lin(rb + ".noise_func.noise_func.0", inner, cout)
                else:
                    lin(rb + ".mlp.1", inner, cout)
            gn(rb + ".block1.block.0", cin)
            conv(rb + ".block1.block.3", cin, cout, 3)
            gn(rb + ".block2.block.0", cout)
            conv(rb + ".block2.block.3", cout, cout, 3)
            if cin != cout:
                conv(rb + ".res_conv", cin, cout, 1)
            if attn:
                gn(p + ".attn.norm", cout)
                conv(p + ".attn.qkv", cout, 3 * cout, 1, bias=False)
                conv(p + ".attn.out", cout, cout, 1)
    _, p, cin, cout, _ = tab["final"]
    gn(p + ".block.0", cin)
    conv(p + ".block.3", cin, cout, 3)
    return sd


def count_flops(cfg: dict, H: int, W: int) -> float:
    """Algorithmic FLOPs (2*MAC) of one forward for one sample: every conv (upsample
    convs at the upsampled size, concat convs at full K) + 4*N^2*C per attention
    (BASELINE.md section 3 convention)."""
    tab = layer_table(cfg)
    fl = 0.0
    h, w = H, W

    def res(cin, cout, attn):
        f = 2.0 * h * w * 9 * (cin * cout + cout * cout)
        if cin != cout:
            f += 2.0 * h * w * cin * cout
        if attn:
            n = h * w
            f += 2.0 * n * cout * 3 * cout + 2.0 * n * cout * cout + 4.0 * n * n * cout
        return f

    for kind, p, cin, cout, attn in tab["downs"]:
        if kind == "conv":
            fl += 2.0 * h * w * 9 * cin * cout
        elif kind == "res":
            fl += res(cin, cout, attn)
        else:
            h //= 2
            w //= 2
            fl += 2.0 * h * w * 9 * cin * cout
    for kind, p, cin, cout, attn in tab["mid"]:
        fl += res(cin, cout, attn)
    for kind, p, cin, cout, attn in tab["ups"]:
        if kind == "res":
            fl += res(cin, cout, attn)
        else:
            h *= 2
            w *= 2
            fl += 2.0 * h * w * 9 * cin * cout
    _, p, cin, cout, _ = tab["final"]
    fl += 2.0 * h * w * 9 * cin * cout
    return fl
