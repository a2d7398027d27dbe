"""Host-side mirror of the reference UNets: same constructor, same ``forward(x, time)``, same ``state_dict``
keys - but ``forward`` is a single call into the C ABI (``ds_unet_forward``).

Reference: ``model/sr3_modules/unet.py:161-259`` (``variant='sr3'``, ``time`` = noise level ``(B,1)``) and
``model/ddpm_modules/unet.py:147-243`` (``variant='ddpm'``, ``time`` = ``(B,)`` or ``(1,)``).

The parameter tree is generated from the native library's own weight registry
(``ds_unet_weight_name/shape``), so the names the kernels look up and the names a reference checkpoint
provides cannot drift apart; ``tests/test_abi_host.py::test_state_dict_keys_match_reference`` pins them against
key lists recorded from the reference modules (``tests/golden/state_dict_keys.json``).
"""
import ctypes as C
import math
import os

import torch
from torch import nn

from .. import _lib

_PRECISIONS = {"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16, "tf32": _lib.PREC_TF32}
TF32_MAX_INNER = 32


def default_precision() -> str:
    return os.environ.get("DIFFSPLIT_B200_PRECISION", "auto")


def resolve_precision(precision, inner_channel) -> str:
    """``auto`` (the default): TF32 tensor-core operands for the narrow splitting nets (inner_channel <= 32) - they are
    latency / bandwidth bound, so the halved MMA rate costs nothing and the per-step error drops 4x below bf16's (it is what
    the reference's own cuDNN convolutions use on a GPU) -, bf16 operands for the wide SR3 nets, which are tensor bound."""
    precision = precision or default_precision()
    if precision == "auto":
        precision = "tf32" if inner_channel <= TF32_MAX_INNER else "bf16"
    if precision not in _PRECISIONS:
        raise ValueError(f"precision must be one of auto, fp32, bf16, tf32 - got {precision!r}")
    return precision


class _Node(nn.Module):
    """Anonymous container; exists only to reproduce the reference's dotted parameter paths."""

    def forward(self, *a, **k):
        raise RuntimeError("parameter holder: call the owning UNet")


class UNet(nn.Module):
    def __init__(self, in_channel=6, out_channel=3, inner_channel=32, norm_groups=32,
                 channel_mults=(1, 2, 4, 8, 8), attn_res=(8), res_blocks=3, dropout=0,
                 with_noise_level_emb=True, with_time_emb=None, image_size=128, variant="sr3",
                 precision=None):
        super().__init__()
        if variant not in ("sr3", "ddpm"):
            raise ValueError(f"variant must be 'sr3' or 'ddpm', got {variant!r}")
        if with_time_emb is None:
            with_time_emb = with_noise_level_emb
        if isinstance(attn_res, int):          # the reference default `(8)` is an int, `x in 8` would raise there
            attn_res = (attn_res,)
        attn_res = tuple(attn_res or ())
        channel_mults = tuple(channel_mults)
        if out_channel is None:
            out_channel = in_channel
        self.variant = variant
        self.in_channel, self.out_channel = int(in_channel), int(out_channel)
        self.inner_channel, self.norm_groups = int(inner_channel), int(norm_groups)
        self.channel_mults, self.attn_res, self.res_blocks = channel_mults, attn_res, int(res_blocks)
        self.dropout = dropout           # identity at inference (Block, unet.py:86); kept for repr/config parity
        self.image_size = int(image_size)
        self.with_time_emb = bool(with_time_emb)
        self.precision = resolve_precision(precision, self.inner_channel)

        d = _lib.UNetDesc()
        d.variant = _lib.UNET_SR3 if variant == "sr3" else _lib.UNET_DDPM
        d.in_channel, d.out_channel = self.in_channel, self.out_channel
        d.inner_channel, d.norm_groups = self.inner_channel, self.norm_groups
        if len(channel_mults) > _lib.MAX_LEVELS or len(attn_res) > _lib.MAX_LEVELS:
            raise ValueError("too many levels")
        d.n_mults = len(channel_mults)
        for i, m in enumerate(channel_mults):
            d.channel_mults[i] = int(m)
        d.n_attn_res = len(attn_res)
        for i, r in enumerate(attn_res):
            d.attn_res[i] = int(r)
        d.res_blocks, d.image_size, d.with_time_emb = self.res_blocks, self.image_size, int(self.with_time_emb)
        d.tf32_weights = int(self.precision == "tf32")
        self._desc = d
        handle = C.c_void_p()
        _lib.check(_lib.lib().ds_unet_create(C.byref(d), C.byref(handle)))
        self._handle = handle
        self._handle_device = None            # CUDA device the library-side weight arenas live on (set by the first commit)
        self._names = []
        self._build_tree()
        self._sig = None
        self._ws = {}
        self._time_cache = None

    # ------------------------------------------------------------------ parameter tree
    def _build_tree(self):
        L = _lib.lib()
        n = L.ds_unet_num_weights(self._handle)
        ndim, shape = C.c_int32(), (C.c_int64 * 4)()
        for i in range(n):
            name = L.ds_unet_weight_name(self._handle, i).decode()
            _lib.check(L.ds_unet_weight_shape(self._handle, i, C.byref(ndim), shape))
            shp = tuple(shape[j] for j in range(ndim.value))
            parts = name.split(".")
            node = self
            for p in parts[:-1]:
                if p not in node._modules:
                    node.add_module(p, _Node())
                node = node._modules[p]
            leaf = parts[-1]
            if leaf == "inv_freq":
                dim = self.inner_channel
                node.register_buffer(leaf, torch.exp(torch.arange(0, dim, 2, dtype=torch.float32) * (-math.log(10000) / dim)))
            else:
                node.register_parameter(leaf, nn.Parameter(self._init(name, shp)))
            self._names.append(name)

    @staticmethod
    def _init(name, shp):
        # PyTorch default initialisation of the layer types the reference uses
        if len(shp) == 1 and (".block.0." in name or ".norm." in name):       # GroupNorm affine
            return torch.ones(shp) if name.endswith("weight") else torch.zeros(shp)
        if name.endswith("weight"):
            w = torch.empty(shp)
            nn.init.kaiming_uniform_(w, a=math.sqrt(5))
            return w
        return torch.empty(shp)      # biases: filled by _init_biases once the matching weight exists

    def reset_parameters(self):
        sd = dict(self.named_parameters())
        for name, p in sd.items():
            if name.endswith("bias") and not (".block.0." in name or ".norm." in name):
                w = sd[name[:-4] + "weight"]
                fan_in = w[0].numel()
                bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
                with torch.no_grad():
                    p.uniform_(-bound, bound)

    def _apply(self, fn, *a, **k):
        self._sig = None
        return super()._apply(fn, *a, **k)

    def __del__(self):
        try:
            if getattr(self, "_handle", None):
                _lib.lib().ds_unet_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    # ------------------------------------------------------------------ weights -> device pack
    def _tensors(self):
        out = []
        for name in self._names:
            obj = self
            for p in name.split("."):
                obj = getattr(obj, p)
            out.append(obj)
        return out

    def commit(self):
        """(Re)pack the weights into the library's layouts if any parameter changed since the last call."""
        ts = self._tensors()
        sig = tuple((t.data_ptr(), t._version) for t in ts)
        if sig == self._sig:
            return
        for t in ts:
            _lib.require_cuda(t, "UNet parameter")
        dev = ts[0].device
        if self._handle_device is not None and dev != self._handle_device:
            # .to(another GPU): the handle's arenas, streams and events are bound to the device of the first commit -> new handle
            _lib.lib().ds_unet_destroy(self._handle)
            handle = C.c_void_p()
            _lib.check(_lib.lib().ds_unet_create(C.byref(self._desc), C.byref(handle)))
            self._handle = handle
            self._ws = {}
        self._handle_device = dev
        views = (_lib.TensorView * len(ts))()
        keep = []
        for i, (name, t) in enumerate(zip(self._names, ts)):
            td = t.detach()
            if td.dtype != torch.float32 or not td.is_contiguous():
                td = td.float().contiguous()
            keep.append(td)
            views[i].name = name.encode()
            views[i].d_data = td.data_ptr()
            views[i].ndim = td.dim()
            for j, s in enumerate(td.shape):
                views[i].shape[j] = s
        with torch.cuda.device(ts[0].device):
            _lib.check(_lib.lib().ds_unet_load_weights(self._handle, views, len(ts), _lib.stream_ptr()))
        self._sig = sig

    # ------------------------------------------------------------------ forward
    def workspace(self, B, H, W, prec, device):
        """Scratch tensor for one (shape, precision).  The cache only avoids re-allocation for eager callers: anything that
        bakes the address into a CUDA graph (the sampler engines) keeps its own reference, so evicting an entry here never
        frees memory a captured graph still uses."""
        if prec is None:
            prec = _PRECISIONS[self.precision]
        key = (B, H, W, prec, torch.device(device))
        ws = self._ws.get(key)
        if ws is None:
            nbytes = _lib.lib().ds_unet_workspace_bytes(self._handle, B, H, W, prec)
            if nbytes == 0:
                _lib.check(-1)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
            while len(self._ws) >= 8:
                self._ws.pop(next(iter(self._ws)))          # oldest entry only
            self._ws[key] = ws
        return ws

    def forward_into(self, out, xa, xb, time, precision=None, ws=None):
        """Raw entry: all tensors fp32 contiguous CUDA; ``xb`` may be None; ``time`` fp32 with 1 or B entries."""
        B, ca, H, W = xa.shape
        cb = 0 if xb is None else xb.shape[1]
        prec = _PRECISIONS[precision or self.precision]
        if ws is None:
            ws = self.workspace(B, H, W, prec, xa.device)
        _lib.check(_lib.lib().ds_unet_forward(
            self._handle, xa.data_ptr(), ca, None if xb is None else xb.data_ptr(), cb,
            None if time is None else time.data_ptr(), 0 if time is None else time.numel(),
            out.data_ptr(), B, H, W, prec, ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
        return out

    def forward(self, x, time=None, cond=None):
        """``UNet(x, time)`` as in the reference.  ``cond`` (optional) is concatenated IN FRONT of ``x`` along
        channels without materialising the concat (p_mean_variance, sr3 diffusion.py:157-158)."""
        _lib.require_cuda(x, "UNet input")
        self.commit()
        xa, xb = (x, None) if cond is None else (cond, x)
        xa = xa.float().contiguous()
        xb = None if xb is None else xb.float().contiguous()
        B = xa.shape[0]
        t = None
        if self.with_time_emb:
            if time is None:
                raise ValueError("this UNet was built with a time embedding; pass `time`")
            t = time.to(device=x.device, dtype=torch.float32).reshape(-1).contiguous()
            if t.numel() not in (1, B):
                raise ValueError(f"time must have 1 or {B} entries, got {tuple(time.shape)}")
        out = torch.empty((B, self.out_channel, xa.shape[2], xa.shape[3]), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            self.forward_into(out, xa, xb, t)
        return out

    def profile(self, xa, xb, time, precision=None, max_ops=1024):
        """One forward with per-operator CUDA-event timing (``ds_unet_forward_profiled``); list of dicts."""
        self.commit()
        B, ca, H, W = xa.shape
        cb = 0 if xb is None else xb.shape[1]
        prec = _PRECISIONS[precision or self.precision]
        ws = self.workspace(B, H, W, prec, xa.device)
        out = torch.empty((B, self.out_channel, H, W), dtype=torch.float32, device=xa.device)
        ops = (_lib.OpProfile * max_ops)()
        n = C.c_int()
        _lib.check(_lib.lib().ds_unet_forward_profiled(
            self._handle, xa.data_ptr(), ca, None if xb is None else xb.data_ptr(), cb,
            None if time is None else time.data_ptr(), 0 if time is None else time.numel(), out.data_ptr(),
            B, H, W, prec, ws.data_ptr(), ws.numel(), _lib.stream_ptr(), ops, max_ops, C.byref(n)))
        names = {0: "temb", 1: "conv_f32", 2: "groupnorm_swish", 3: "attention", 4: "conv_tc", 5: "gn_stats",
                 6: "gn_swish_conv_tc", 7: "conv_chain_tc"}
        return [dict(kind=names[o.kind], cin=o.cin, cout=o.cout, ksize=o.ksize, h=o.h, w=o.w, launches=o.launches,
                     ms=o.ms, flops=o.flops, bytes=o.bytes) for o in ops[:n.value]]

    def flops(self, H, W):
        return _lib.lib().ds_unet_flops(self._handle, H, W)

    def launches(self, B, H, W, precision=None):
        return _lib.lib().ds_unet_launches(self._handle, B, H, W, _PRECISIONS[precision or self.precision])

    def read_tap(self, name, B, C_, H, W, precision=None):
        prec = _PRECISIONS[precision or self.precision]
        ws = next(v for k, v in self._ws.items() if k[3] == prec)
        out = torch.empty((B, C_, H, W), dtype=torch.float32, device=ws.device)
        _lib.check(_lib.lib().ds_unet_read_tap(self._handle, name.encode(), out.data_ptr(), out.numel(), ws.data_ptr(),
                                               _lib.stream_ptr()))
        return out

    def extra_repr(self):
        return (f"variant={self.variant}, in={self.in_channel}, out={self.out_channel}, inner={self.inner_channel}, "
                f"groups={self.norm_groups}, mults={self.channel_mults}, attn_res={self.attn_res}, "
                f"res_blocks={self.res_blocks}, precision={self.precision}")


def UNetSr3(**kw):
    kw.setdefault("variant", "sr3")
    net = UNet(**kw)
    net.reset_parameters()
    return net


def UNetDdpm(**kw):
    kw.setdefault("variant", "ddpm")
    net = UNet(**kw)
    net.reset_parameters()
    return net
