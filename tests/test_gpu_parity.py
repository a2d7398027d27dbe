"""GPU suite: the CUDA path (through the C ABI / the reference-shaped Python boundary) against the oracle, the
golden vectors recorded from the real reference, and size-independent properties at BASELINE sizes.

Tolerances: per-step UNet output, rel-err = max|y - ref| / max|ref|: <= 1e-5 in fp32 mode and <= 1e-2 in the tensor-core
mode a network runs in by default (north_star's two gates): "auto" = TF32 operands for the narrow splitting nets (inner
channel <= 32; measured ~3e-3), bf16 operands for the wide SR3 nets (measured 0.5-0.8e-2).  bf16 operands on a NARROW
random-weight net (not the default; kept as a selectable mode) sit at the rounding floor of bf16 itself - rounding ONLY the
conv operands to bf16 inside the fp32 CPU oracle gives max-rel 1.3e-2 on the hagen net - and are held to 2e-2.
Final PSNR within 0.1 dB at a 34 dB operating point;
sampler updates with injected noise: <= 2e-5 abs on O(1) images; tile indexing / stitching: bit-exact; Philox replay vs
torch.randn on the same device: bit-exact.
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from diffsplitting_b200 import _lib
from diffsplitting_b200.data import TiledFrames, TileIndexManager, TilingMode, stitch_predictions
from diffsplitting_b200.model import create_model
from diffsplitting_b200.model.samplers import GaussianDiffusionDdpm, GaussianDiffusionSr3, InDI, JointIndi
from diffsplitting_b200.model.unet import UNet
from oracle import metrics_ref as M
from oracle import philox_ref as PH
from oracle import samplers_ref as S
from oracle import tiling_ref as TR
from oracle import unet_ref as U
from oracle.make_golden import TIMEPRED_CASES, UNET_CASES, Replay
from tests.configs import make_opt

DEV = "cuda"
PRECISIONS = ["fp32", "auto", "bf16"]


def tol(net, requested):
    """(max-rel, rel-RMS) gate for `net` built with precision `requested`."""
    if net.precision == "fp32":
        return 1e-5, 5e-6
    if requested == "bf16" and net.inner_channel <= 32:
        return 2e-2, 1.5e-2            # non-default mode: bf16 operands on a narrow net (see the module docstring)
    return 1e-2, 7.5e-3                # north_star: <= 1e-2 in the tensor-core mode




def relerr(y, ref):
    return float((y.double().cpu() - ref.double().cpu()).abs().max() / ref.double().abs().max().clamp_min(1e-12))


def relrms(y, ref):
    d = y.double().cpu() - ref.double().cpu()
    return float(d.pow(2).mean().sqrt() / ref.double().pow(2).mean().sqrt().clamp_min(1e-12))


def build(cfg, sd=None, precision="fp32"):
    net = UNet(in_channel=cfg["in_channel"], out_channel=cfg["out_channel"], inner_channel=cfg["inner_channel"],
               norm_groups=cfg["norm_groups"], channel_mults=cfg["channel_mults"], attn_res=cfg["attn_res"],
               res_blocks=cfg["res_blocks"], image_size=cfg["image_size"], variant=cfg["variant"], precision=precision)
    if sd is not None:
        net.load_state_dict(sd, strict=True)
    return net.to(DEV).eval()


def sptr():
    return _lib.stream_ptr()


# ------------------------------------------------------------------------------------------------ RNG
@pytest.mark.parametrize("numel", [100, 12345, 16 * 64 * 64, 8 * 3 * 512 * 512 + 7])
def test_philox_replays_torch_randn_bit_exact(numel):
    sms, mt, major, _ = _lib.device_info(0)
    assert major == 10, "sm_100 device expected"
    gen = torch.cuda.default_generators[0]
    torch.manual_seed(1234)
    gen.set_offset(40)
    grid = min(sms * (mt // 256), (numel + 255) // 256)
    threads, inc = 256 * grid, ((numel - 1) // (256 * grid * 4) + 1) * 4
    mine = torch.empty(numel, device=DEV)
    _lib.check(_lib.lib().ds_randn_axpy(None, 1.0, mine.data_ptr(), numel, gen.initial_seed(), gen.get_offset(), None,
                                        inc, threads, sptr()))
    ref = torch.randn(numel, device=DEV)
    assert gen.get_offset() == 40 + inc, "offset bookkeeping differs from torch"
    assert torch.equal(mine, ref)
    if numel <= 1 << 16:
        cpu = PH.randn_like_cuda(numel, 1234, 40, sms, mt)
        assert np.abs(cpu - ref.cpu().numpy()).max() < 1e-5
        assert PH.offset_increment(numel, sms, mt) == inc


# ------------------------------------------------------------------------------------------------ operators
@pytest.mark.parametrize("cin,cout,ks,stride,up,B,H,W", [
    (16, 16, 3, 1, 0, 2, 16, 16), (48, 16, 3, 1, 0, 1, 24, 20), (32, 32, 3, 2, 0, 2, 16, 16), (64, 64, 3, 1, 1, 1, 8, 12),
    (128, 384, 1, 1, 0, 2, 8, 8), (3, 16, 3, 1, 0, 1, 9, 11), (16, 2, 3, 1, 0, 2, 16, 16), (192, 64, 1, 1, 0, 1, 16, 16),
    (256, 128, 3, 1, 0, 1, 8, 8)])
def test_conv_operator(cin, cout, ks, stride, up, B, H, W):
    g = torch.Generator().manual_seed(cin * 7 + cout)
    x = torch.randn((B, cin, H, W), generator=g)
    w = torch.randn((cout, cin, ks, ks), generator=g) / (cin * ks * ks) ** 0.5
    b = torch.randn((cout,), generator=g)
    xr = F.interpolate(x, scale_factor=2, mode="nearest") if up else x
    ref = F.conv2d(xr.double(), w.double(), b.double(), stride=stride, padding=ks // 2).float()
    xd = x.permute(0, 2, 3, 1).contiguous().to(DEV)
    out = torch.empty((B, ref.shape[2], ref.shape[3], cout), device=DEV)
    nb = _lib.lib().ds_conv2d_scratch_bytes(cin, cout, ks)
    scratch = torch.empty(nb, dtype=torch.uint8, device=DEV)
    wd, bd = w.to(DEV), b.to(DEV)
    _lib.check(_lib.lib().ds_conv2d_f32(xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), out.data_ptr(), B, H, W, cin, cout, ks,
                                        stride, up, scratch.data_ptr(), nb, sptr()))
    assert relerr(out.permute(0, 3, 1, 2), ref) <= 1e-5


TC_CASES = [
    # ca, cb, cout, ks, stride, up, B, H, W, residual, out_nchw
    (16, 0, 16, 3, 1, 0, 2, 16, 16, 0, 0), (32, 0, 32, 3, 1, 0, 1, 32, 32, 0, 0), (64, 0, 64, 3, 1, 0, 2, 8, 8, 1, 0),
    (128, 0, 128, 3, 1, 0, 16, 8, 8, 1, 0), (32, 16, 16, 3, 1, 0, 2, 16, 16, 0, 0), (128, 64, 128, 1, 1, 0, 1, 16, 16, 0, 0),
    (16, 0, 16, 3, 2, 0, 2, 16, 16, 0, 0), (64, 0, 64, 3, 2, 0, 1, 32, 32, 0, 0), (32, 0, 32, 3, 1, 1, 2, 8, 8, 0, 0),
    (128, 0, 128, 3, 1, 1, 1, 16, 16, 0, 0), (16, 0, 1, 3, 1, 0, 2, 16, 16, 0, 1), (16, 0, 6, 3, 1, 0, 1, 32, 32, 0, 1),
    (128, 0, 384, 1, 1, 0, 2, 8, 8, 0, 0), (16, 0, 16, 3, 1, 0, 1, 24, 20, 1, 0), (256, 0, 256, 3, 1, 0, 1, 4, 4, 0, 0),
    (16, 0, 16, 3, 1, 0, 16, 64, 64, 1, 0), (48, 0, 16, 3, 1, 0, 1, 128, 128, 0, 0), (192, 0, 64, 3, 1, 0, 2, 16, 16, 0, 0),
    (512, 0, 512, 3, 1, 0, 1, 16, 16, 1, 0),
    # enough tiles for the tall-patch variant (>= 296 CTAs, >= 64 channels): one TMA patch per chunk, taps = row shifts
    (64, 0, 64, 3, 1, 0, 8, 128, 128, 1, 0), (64, 64, 32, 3, 1, 0, 10, 100, 90, 0, 0),
    # upsample x2 + 3x3 in the persistent kernel (parity class = part of the work item), aligned and ragged
    (128, 0, 128, 3, 1, 1, 8, 64, 64, 0, 0), (64, 0, 96, 3, 1, 1, 6, 40, 36, 0, 0), (256, 0, 256, 3, 1, 1, 3, 32, 32, 0, 0)]


@pytest.mark.parametrize("ca,cb,cout,ks,stride,up,B,H,W,residual,out_nchw", TC_CASES)
def test_conv_tc_operator(ca, cb, cout, ks, stride, up, B, H, W, residual, out_nchw):
    """tcgen05 implicit-GEMM conv vs an fp64 conv of the same bf16-rounded operands (tolerance: bf16 output rounding)."""
    g = torch.Generator().manual_seed(ca * 3 + cout + ks + H)
    cin = ca + cb
    x = torch.randn((B, cin, H, W), generator=g).bfloat16()
    w = (torch.randn((cout, cin, ks, ks), generator=g) / (cin * ks * ks) ** 0.5)
    b = torch.randn((cout,), generator=g)
    xr = F.interpolate(x.float(), scale_factor=2, mode="nearest") if up else x.float()
    ref = F.conv2d(xr.double(), w.bfloat16().double(), b.double(), stride=stride, padding=ks // 2)
    Ho, Wo = ref.shape[2], ref.shape[3]
    res = None
    if residual:
        res = torch.randn((B, cout, Ho, Wo), generator=g)
        ref = ref + res.double()
    xa = x[:, :ca].permute(0, 2, 3, 1).contiguous().to(DEV)
    xb = x[:, ca:].permute(0, 2, 3, 1).contiguous().to(DEV) if cb else None
    rd = res.permute(0, 2, 3, 1).contiguous().to(DEV) if residual else None
    out = torch.full((B, cout, Ho, Wo), float("nan"), device=DEV) if out_nchw else \
        torch.full((B, Ho, Wo, cout), float("nan"), dtype=torch.bfloat16, device=DEV)
    out32 = None if out_nchw else torch.full((B, Ho, Wo, cout), float("nan"), device=DEV)
    nb = _lib.lib().ds_conv2d_bf16_scratch_bytes(cin, cout, ks)
    scratch = torch.zeros(nb, dtype=torch.uint8, device=DEV)
    wd, bd = w.to(DEV), b.to(DEV)
    _lib.check(_lib.lib().ds_conv2d_bf16(xa.data_ptr(), ca, None if xb is None else xb.data_ptr(), cb, wd.data_ptr(), bd.data_ptr(),
                                         None if rd is None else rd.data_ptr(), out.data_ptr(),
                                         None if out32 is None else out32.data_ptr(), out_nchw, B, H, W, cout, ks, stride,
                                         up, scratch.data_ptr(), nb, sptr()))
    torch.cuda.synchronize()
    y = out if out_nchw else out.float().permute(0, 3, 1, 2)
    e = relerr(y, ref.float())
    e32 = relerr(out32.permute(0, 3, 1, 2), ref.float()) if out32 is not None else 0.0
    print(f"[conv_tc {ca}+{cb}->{cout} k{ks} s{stride} up{up} {B}x{H}x{W}] rel err bf16 out {e:.3e}, fp32 out {e32:.3e}")
    assert e <= (1e-2 if up else 5e-3)
    assert e32 <= (1e-2 if up else 1e-5)       # fp32 accumulation of exact bf16 products: only summation order differs


@pytest.mark.parametrize("ca,cb,cout,ks,G,swish,B,H,W,residual", [
    (16, 0, 16, 3, 16, 1, 2, 16, 16, 0), (16, 0, 16, 3, 16, 1, 16, 64, 64, 1), (32, 16, 16, 3, 16, 1, 2, 64, 64, 0),
    (32, 0, 32, 3, 8, 1, 3, 32, 32, 1), (64, 32, 32, 3, 16, 1, 1, 32, 32, 0), (64, 0, 64, 3, 16, 1, 2, 16, 16, 1),
    (128, 0, 128, 3, 16, 1, 4, 8, 8, 1), (128, 0, 384, 1, 16, 0, 2, 8, 8, 0), (16, 0, 1, 3, 16, 1, 2, 32, 32, 0),
    (16, 0, 16, 3, 0, 0, 1, 24, 20, 0), (64, 64, 128, 3, 32, 1, 5, 4, 4, 0), (48, 0, 16, 3, 16, 1, 1, 128, 96, 0),
    (128, 64, 64, 3, 16, 1, 3, 16, 16, 1), (128, 96, 48, 3, 16, 1, 2, 12, 12, 0),
    # wide images: 2-D tiles of 7 x 16 outputs (ragged in both directions), 3x3 and 1x1, concat
    (16, 0, 16, 3, 16, 1, 2, 40, 150, 1), (32, 16, 16, 3, 16, 1, 1, 33, 141, 0), (32, 0, 48, 1, 8, 0, 1, 20, 160, 0),
    # >= 4 waves of tiles: the persistent pipelined variant (conv_stream_kernel), ragged tiles, concat, N split, 3 outputs
    (64, 0, 64, 3, 16, 1, 8, 128, 128, 1), (32, 32, 64, 3, 16, 1, 5, 100, 150, 0), (64, 0, 3, 3, 16, 1, 6, 133, 121, 0),
    (64, 0, 128, 3, 16, 1, 4, 96, 160, 1), (16, 0, 16, 3, 16, 1, 8, 200, 168, 1)])
def test_fused_gn_swish_conv_operator(ca, cb, cout, ks, G, swish, B, H, W, residual):
    """conv_halo_kernel: GroupNorm statistics pass + ONE tensor-core kernel (normalise + Swish in the operand staging)
    vs fp64 conv of the bf16-rounded normalised activations and weights."""
    g = torch.Generator().manual_seed(ca + cb + cout + H)
    cin = ca + cb
    x = torch.randn((B, cin, H, W), generator=g) * 1.5 + 0.3
    gamma, beta = 1 + 0.3 * torch.randn(cin, generator=g), 0.3 * torch.randn(cin, generator=g)
    w = torch.randn((cout, cin, ks, ks), generator=g) / (cin * ks * ks) ** 0.5
    b = torch.randn((cout,), generator=g)
    a = x.double()
    if G:
        a = F.group_norm(a, G, gamma.double(), beta.double(), eps=1e-5)
    if swish:
        a = a * torch.sigmoid(a)
    ref = F.conv2d(a.float().bfloat16().double(), w.bfloat16().double(), b.double(), padding=ks // 2)
    res = None
    if residual:
        res = torch.randn((B, cout, H, W), generator=g)
        ref = ref + res.double()
    xa = x[:, :ca].permute(0, 2, 3, 1).contiguous().to(DEV)
    xb = x[:, ca:].permute(0, 2, 3, 1).contiguous().to(DEV) if cb else None
    rd = res.permute(0, 2, 3, 1).contiguous().to(DEV) if residual else None
    out16 = torch.full((B, H, W, cout), float("nan"), dtype=torch.bfloat16, device=DEV)
    out32 = torch.full((B, H, W, cout), float("nan"), device=DEV)
    nb = _lib.lib().ds_gnconv_bf16_scratch_bytes(B, max(G, 1), cin, cout, ks)
    scratch = torch.zeros(nb, dtype=torch.uint8, device=DEV)
    gd, bd, wd, biasd = gamma.to(DEV), beta.to(DEV), w.to(DEV), b.to(DEV)
    _lib.check(_lib.lib().ds_gnconv_bf16(xa.data_ptr(), ca, None if xb is None else xb.data_ptr(), cb, gd.data_ptr(), bd.data_ptr(),
                                         G, swish, wd.data_ptr(), biasd.data_ptr(), None if rd is None else rd.data_ptr(),
                                         out16.data_ptr(), out32.data_ptr(), B, H, W, cout, ks, scratch.data_ptr(), nb, sptr()))
    torch.cuda.synchronize()
    e32 = relerr(out32.permute(0, 3, 1, 2), ref.float())
    e16 = relerr(out16.float().permute(0, 3, 1, 2), ref.float())
    print(f"[gnconv {ca}+{cb}->{cout} k{ks} G{G} {B}x{H}x{W}] rel err fp32 out {e32:.3e} bf16 out {e16:.3e}")
    # the fp32 output differs from the reference only by bf16 re-rounding of activations whose fp32 value differs in the
    # last bits (fast exp, fp32 statistics): a handful of 1-ulp bf16 flips
    assert e32 <= 2e-3 and e16 <= 6e-3


TF32_CONV_CASES = [
    # ca, cb, cout, ks, stride, up, B, H, W, residual, out_nchw
    (16, 0, 16, 3, 1, 0, 2, 16, 16, 0, 0), (32, 0, 32, 3, 1, 0, 1, 32, 32, 1, 0), (32, 16, 16, 3, 1, 0, 2, 16, 16, 0, 0),
    (128, 64, 128, 1, 1, 0, 1, 16, 16, 0, 0), (16, 0, 16, 3, 2, 0, 2, 16, 16, 0, 0), (64, 0, 64, 3, 2, 0, 1, 32, 32, 0, 0),
    (32, 0, 32, 3, 1, 1, 2, 8, 8, 0, 0), (128, 0, 128, 3, 1, 1, 1, 16, 16, 0, 0), (16, 0, 1, 3, 1, 0, 2, 16, 16, 0, 1),
    (8, 0, 24, 1, 1, 0, 1, 12, 20, 0, 0), (192, 0, 64, 3, 1, 0, 2, 16, 16, 1, 0), (256, 0, 128, 3, 1, 0, 1, 8, 8, 0, 0),
    # tall-patch variant (many tiles, >= 32 channels)
    (32, 0, 32, 3, 1, 0, 8, 128, 128, 1, 0), (64, 32, 32, 3, 1, 0, 10, 100, 90, 0, 0), (32, 0, 16, 3, 1, 0, 8, 128, 128, 0, 0),
    (32, 0, 128, 3, 1, 0, 8, 32, 32, 0, 0),
    # 8 input channels (K = 8 per tap): the network's first conv on the 8-channel repack of the input
    (8, 0, 64, 3, 1, 0, 2, 40, 36, 0, 0)]


@pytest.mark.parametrize("ca,cb,cout,ks,stride,up,B,H,W,residual,out_nchw", TF32_CONV_CASES)
def test_conv_tf32_operator(ca, cb, cout, ks, stride, up, B, H, W, residual, out_nchw):
    """The TMA-fed tcgen05 conv with fp32 tensors read as TF32 operands (kind::tf32) vs an fp64 conv of the fp32 operands:
    the difference is the TF32 operand rounding (2^-11 per operand), far below bf16's."""
    g = torch.Generator().manual_seed(ca * 3 + cout + ks + H)
    cin = ca + cb
    x = torch.randn((B, cin, H, W), generator=g)
    w = (torch.randn((cout, cin, ks, ks), generator=g) / (cin * ks * ks) ** 0.5)
    b = torch.randn((cout,), generator=g)
    xr = F.interpolate(x, scale_factor=2, mode="nearest") if up else x
    ref = F.conv2d(xr.double(), w.double(), b.double(), stride=stride, padding=ks // 2)
    Ho, Wo = ref.shape[2], ref.shape[3]
    res = None
    if residual:
        res = torch.randn((B, cout, Ho, Wo), generator=g)
        ref = ref + res.double()
    xa = x[:, :ca].permute(0, 2, 3, 1).contiguous().to(DEV)
    xb = x[:, ca:].permute(0, 2, 3, 1).contiguous().to(DEV) if cb else None
    rd = res.permute(0, 2, 3, 1).contiguous().to(DEV) if residual else None
    out = torch.full((B, cout, Ho, Wo) if out_nchw else (B, Ho, Wo, cout), float("nan"), device=DEV)
    nb = _lib.lib().ds_conv2d_tf32_scratch_bytes(cin, cout, ks)
    scratch = torch.zeros(nb, dtype=torch.uint8, device=DEV)
    wd, bd = w.to(DEV), b.to(DEV)
    _lib.check(_lib.lib().ds_conv2d_tf32(xa.data_ptr(), ca, None if xb is None else xb.data_ptr(), cb, wd.data_ptr(), bd.data_ptr(),
                                         None if rd is None else rd.data_ptr(), out.data_ptr(), out_nchw, B, H, W, cout, ks, stride,
                                         up, scratch.data_ptr(), nb, sptr()))
    torch.cuda.synchronize()
    y = out if out_nchw else out.permute(0, 3, 1, 2)
    e = relerr(y, ref.float())
    print(f"[conv_tf32 {ca}+{cb}->{cout} k{ks} s{stride} up{up} {B}x{H}x{W}] rel err {e:.3e}")
    assert e <= 1.5e-3


@pytest.mark.parametrize("ca,cb,cout,ks,G,swish,B,H,W,residual", [
    (16, 0, 16, 3, 16, 1, 2, 16, 16, 0), (16, 0, 16, 3, 16, 1, 16, 64, 64, 1), (32, 16, 16, 3, 16, 1, 2, 64, 64, 0),
    (32, 0, 32, 3, 8, 1, 3, 32, 32, 1), (64, 32, 32, 3, 16, 1, 1, 32, 32, 0), (64, 0, 64, 3, 16, 1, 2, 16, 16, 1),
    (128, 0, 128, 3, 16, 1, 4, 8, 8, 1), (128, 0, 384, 1, 16, 0, 2, 8, 8, 0), (16, 0, 1, 3, 16, 1, 2, 32, 32, 0),
    (16, 0, 16, 3, 0, 0, 1, 24, 20, 0), (48, 0, 16, 3, 16, 1, 1, 128, 96, 0), (64, 64, 64, 3, 16, 1, 3, 16, 16, 1),
    (16, 0, 16, 3, 16, 1, 2, 40, 150, 1), (32, 16, 16, 3, 16, 1, 1, 33, 141, 0), (32, 0, 48, 1, 8, 0, 1, 20, 160, 0),
    # the persistent pipelined variant (>= 4 waves of tiles)
    (16, 0, 16, 3, 16, 1, 8, 200, 168, 1), (32, 16, 16, 3, 16, 1, 5, 150, 140, 0), (32, 0, 32, 3, 8, 1, 8, 128, 128, 1),
    (48, 0, 16, 3, 16, 1, 6, 133, 121, 0), (64, 0, 64, 3, 16, 1, 8, 128, 128, 1)])
def test_fused_gn_swish_conv_tf32_operator(ca, cb, cout, ks, G, swish, B, H, W, residual):
    """conv_halo_kernel with TF32 operands: GroupNorm + Swish in the operand staging, fp32 -> tf32 (round to nearest),
    vs the fp64 reference of the fp32 operands."""
    g = torch.Generator().manual_seed(ca + cb + cout + H)
    cin = ca + cb
    x = torch.randn((B, cin, H, W), generator=g) * 1.5 + 0.3
    gamma, beta = 1 + 0.3 * torch.randn(cin, generator=g), 0.3 * torch.randn(cin, generator=g)
    w = torch.randn((cout, cin, ks, ks), generator=g) / (cin * ks * ks) ** 0.5
    b = torch.randn((cout,), generator=g)
    a = x.double()
    if G:
        a = F.group_norm(a, G, gamma.double(), beta.double(), eps=1e-5)
    if swish:
        a = a * torch.sigmoid(a)
    ref = F.conv2d(a, w.double(), b.double(), padding=ks // 2)
    res = None
    if residual:
        res = torch.randn((B, cout, H, W), generator=g)
        ref = ref + res.double()
    xa = x[:, :ca].permute(0, 2, 3, 1).contiguous().to(DEV)
    xb = x[:, ca:].permute(0, 2, 3, 1).contiguous().to(DEV) if cb else None
    rd = res.permute(0, 2, 3, 1).contiguous().to(DEV) if residual else None
    out32 = torch.full((B, H, W, cout), float("nan"), device=DEV)
    nb = _lib.lib().ds_gnconv_tf32_scratch_bytes(B, max(G, 1), cin, cout, ks)
    scratch = torch.zeros(nb, dtype=torch.uint8, device=DEV)
    gd, bd, wd, biasd = gamma.to(DEV), beta.to(DEV), w.to(DEV), b.to(DEV)
    _lib.check(_lib.lib().ds_gnconv_tf32(xa.data_ptr(), ca, None if xb is None else xb.data_ptr(), cb, gd.data_ptr(), bd.data_ptr(),
                                         G, swish, wd.data_ptr(), biasd.data_ptr(), None if rd is None else rd.data_ptr(),
                                         out32.data_ptr(), B, H, W, cout, ks, scratch.data_ptr(), nb, sptr()))
    torch.cuda.synchronize()
    e32 = relerr(out32.permute(0, 3, 1, 2), ref.float())
    print(f"[gnconv tf32 {ca}+{cb}->{cout} k{ks} G{G} {B}x{H}x{W}] rel err {e32:.3e}")
    assert e32 <= 1.5e-3


@pytest.mark.parametrize("ca,cb,cmid,cout,ks,G,B,H,W,residual", [
    (128, 0, 128, 128, 3, 16, 16, 8, 8, 1), (64, 0, 128, 128, 3, 16, 2, 8, 8, 0), (128, 128, 128, 128, 3, 16, 3, 8, 8, 0),
    (128, 64, 128, 128, 3, 16, 2, 8, 8, 1), (64, 32, 64, 64, 3, 16, 2, 16, 16, 1), (32, 0, 64, 64, 3, 16, 1, 16, 16, 0),
    (128, 0, 256, 256, 1, 16, 2, 8, 8, 0), (16, 16, 32, 32, 3, 8, 5, 4, 4, 0), (64, 0, 64, 64, 3, 16, 2, 12, 10, 0),
    (32, 0, 48, 40, 3, 8, 3, 8, 8, 1), (16, 0, 16, 16, 3, 16, 2, 16, 16, 0)])
def test_persistent_chain_operator(ca, cb, cmid, cout, ks, G, B, H, W, residual):
    """conv_chain_kernel: two fused GN+Swish->conv ops of one sample in ONE persistent CTA (the second GroupNorm reads the
    statistics the first op's epilogue produced inside the same launch) vs the same math in float64 with bf16-rounded
    conv operands."""
    g = torch.Generator().manual_seed(ca + cb + cout + H)
    cin = ca + cb
    x = torch.randn((B, cin, H, W), generator=g) * 1.5 + 0.3
    g1, b1 = 1 + 0.3 * torch.randn(cin, generator=g), 0.3 * torch.randn(cin, generator=g)
    g2, b2 = 1 + 0.3 * torch.randn(cmid, generator=g), 0.3 * torch.randn(cmid, generator=g)
    w1 = torch.randn((cmid, cin, ks, ks), generator=g) / (cin * ks * ks) ** 0.5
    w2 = torch.randn((cout, cmid, ks, ks), generator=g) / (cmid * ks * ks) ** 0.5
    c1, c2 = torch.randn((cmid,), generator=g), torch.randn((cout,), generator=g)

    def block(a, gam, bet, w, bias):
        a = F.group_norm(a, G, gam.double(), bet.double(), eps=1e-5)
        a = a * torch.sigmoid(a)
        return F.conv2d(a.float().bfloat16().double(), w.bfloat16().double(), bias.double(), padding=ks // 2)

    h = block(x.double(), g1, b1, w1, c1)
    ref = block(h.float().double(), g2, b2, w2, c2)       # h is stored as fp32
    res = None
    if residual:
        res = torch.randn((B, cout, H, W), generator=g)
        ref = ref + res.double()
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous().to(DEV)
    xa, xb = nhwc(x[:, :ca]), (nhwc(x[:, ca:]) if cb else None)
    rd = nhwc(res) if residual else None
    out32 = torch.full((B, H, W, cout), float("nan"), device=DEV)
    out16 = torch.full((B, H, W, cout), float("nan"), dtype=torch.bfloat16, device=DEV)
    L = _lib.lib()
    nb = L.ds_chain2_bf16_scratch_bytes(B, H, W, ca, cb, cmid, cout, ks)
    scratch = torch.zeros(nb, dtype=torch.uint8, device=DEV)
    dv = [t.to(DEV) for t in (g1, b1, w1, c1, g2, b2, w2, c2)]
    _lib.check(L.ds_chain2_bf16(xa.data_ptr(), ca, None if xb is None else xb.data_ptr(), cb, dv[0].data_ptr(), dv[1].data_ptr(),
                                dv[2].data_ptr(), dv[3].data_ptr(), dv[4].data_ptr(), dv[5].data_ptr(), dv[6].data_ptr(),
                                dv[7].data_ptr(), None if rd is None else rd.data_ptr(), G, out32.data_ptr(), out16.data_ptr(),
                                B, H, W, cmid, cout, ks, scratch.data_ptr(), nb, sptr()))
    torch.cuda.synchronize()
    e32 = relerr(out32.permute(0, 3, 1, 2), ref.float())
    e16 = relerr(out16.float().permute(0, 3, 1, 2), ref.float())
    print(f"[chain {ca}+{cb}->{cmid}->{cout} k{ks} G{G} {B}x{H}x{W}] rel err fp32 out {e32:.3e} bf16 out {e16:.3e}")
    assert e32 <= 4e-3 and e16 <= 8e-3      # two ops: twice the 1-ulp bf16 re-rounding budget of the single fused op


@pytest.mark.parametrize("ca,cb,G,B,HW,swish", [(16, 0, 16, 2, (16, 16), 1), (16, 32, 16, 1, (12, 20), 1), (128, 0, 16, 3, (8, 8), 0),
                                                 (512, 256, 32, 1, (4, 4), 1), (32, 0, 8, 16, (64, 64), 1), (2048, 0, 16, 1, (8, 8), 1)])
def test_groupnorm_swish_operator(ca, cb, G, B, HW, swish):
    H, W = HW
    g = torch.Generator().manual_seed(ca + cb)
    x = torch.randn((B, ca + cb, H, W), generator=g) * 2 + 0.5
    gamma, beta = torch.randn(ca + cb, generator=g), torch.randn(ca + cb, generator=g)
    ref = F.group_norm(x.double(), G, gamma.double(), beta.double(), eps=1e-5)
    if swish:
        ref = ref * torch.sigmoid(ref)
    xa = x[:, :ca].permute(0, 2, 3, 1).contiguous().to(DEV)
    xb = x[:, ca:].permute(0, 2, 3, 1).contiguous().to(DEV) if cb else None
    out = torch.empty((B, H, W, ca + cb), device=DEV)
    nb = _lib.lib().ds_groupnorm_scratch_bytes(B, G)
    scratch = torch.empty(nb, dtype=torch.uint8, device=DEV)
    gd, bd = gamma.to(DEV), beta.to(DEV)
    _lib.check(_lib.lib().ds_groupnorm_swish_f32(xa.data_ptr(), ca, None if xb is None else xb.data_ptr(), cb, gd.data_ptr(),
                                                 bd.data_ptr(), out.data_ptr(), B, H, W, G, swish, scratch.data_ptr(), nb, sptr()))
    assert relerr(out.permute(0, 3, 1, 2), ref.float()) <= 1e-5


@pytest.mark.parametrize("B,N,Cc", [(2, 16, 128), (1, 64, 32), (2, 100, 64), (1, 256, 512), (1, 1024, 128), (1, 40, 1024)])
def test_attention_operator(B, N, Cc):
    g = torch.Generator().manual_seed(N + Cc)
    qkv = torch.randn((B, N, 3 * Cc), generator=g)
    q, k, v = qkv.double().split(Cc, dim=2)
    a = torch.softmax(q @ k.transpose(1, 2) / Cc ** 0.5, dim=-1)
    ref = (a @ v).float()
    qd = qkv.to(DEV)
    out = torch.empty((B, N, Cc), device=DEV)
    _lib.check(_lib.lib().ds_attention_f32(qd.data_ptr(), out.data_ptr(), B, N, Cc, sptr()))
    assert relerr(out, ref) <= 1e-5


@pytest.mark.parametrize("B,N,Cc,scale", [(16, 64, 128, 1.0), (2, 16, 128, 1.0), (2, 256, 512, 1.0), (1, 300, 128, 3.0), (1, 1024, 256, 1.0),
                                          (2, 100, 64, 4.0), (1, 4096, 128, 2.0), (1, 1024, 1024, 1.0)])
def test_attention_tensor_core_operator(B, N, Cc, scale):
    """tcgen05 attention (bf16 operands, fp32 scores / accumulator) vs float64 softmax attention of the SAME bf16-rounded
    q, k, v.  Error budget: P and the output are rounded to bf16 (2^-9 relative each).  `scale` > 1 sharpens the softmax so
    that the max-subtraction and the two-pass denominator matter."""
    g = torch.Generator().manual_seed(N + Cc)
    qkv = (torch.randn((B, N, 3 * Cc), generator=g) * scale).to(torch.bfloat16)
    q, k, v = qkv.double().split(Cc, dim=2)
    ref = torch.softmax(q @ k.transpose(1, 2) / Cc ** 0.5, dim=-1) @ v
    qd = qkv.to(DEV)
    out = torch.full((B, N, Cc), float("nan"), device=DEV, dtype=torch.bfloat16)
    _lib.check(_lib.lib().ds_attention_bf16(qd.data_ptr(), out.data_ptr(), B, N, Cc, sptr()))
    torch.cuda.synchronize()
    err = (out.double().cpu() - ref).abs().max().item() / ref.abs().max().item()
    rms = ((out.double().cpu() - ref).pow(2).mean() / ref.pow(2).mean()).sqrt().item()
    print(f"[attention tc B{B} N{N} C{Cc}] max-rel {err:.3e} rel-rms {rms:.3e}")
    assert err <= 1e-2 and rms <= 5e-3


# ------------------------------------------------------------------------------------------------ UNet
@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("case", list(UNET_CASES))
def test_unet_matches_reference_golden(gold_dir, case, precision):
    cfg, B, H, W = UNET_CASES[case]
    gd = np.load(os.path.join(gold_dir, f"unet_{case}.npz"))
    sd = U.random_state_dict(cfg, seed=int(gd["seed"]))
    net = build(cfg, sd, precision)
    y = net(torch.from_numpy(gd["x"]).to(DEV), torch.from_numpy(gd["t"]).to(DEV))
    e, r = relerr(y, torch.from_numpy(gd["y"])), relrms(y, torch.from_numpy(gd["y"]))
    print(f"[unet {case} {precision}->{net.precision}] vs reference golden: max-rel {e:.3e} rel-rms {r:.3e}")
    assert e <= tol(net, precision)[0] and r <= tol(net, precision)[1]


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("name,cfg,B,H,W,cond", [
    ("hagen64_b16", U.make_cfg("ddpm", 1, 1, 16, 16, (1, 2, 4, 8), (), 1, 32), 16, 64, 64, 0),
    ("cifar_sr3_cond", U.make_cfg("sr3", 9, 6, 16, 16, (1, 2, 4, 8), (), 1, 32), 1, 32, 32, 3),
    ("sr3_attn_all_levels", U.make_cfg("sr3", 6, 3, 32, 8, (1, 2, 2), (32, 16, 8), 2, 32), 2, 32, 32, 3),
    ("splitting_128", U.make_cfg("sr3", 3, 2, 16, 16, (1, 2, 4, 8), (), 1, 512), 1, 128, 96, 1),
    ("bcast_time", U.make_cfg("ddpm", 1, 1, 16, 16, (1, 2, 4), (), 1, 32), 3, 16, 16, 0)])
def test_unet_matches_oracle(name, cfg, B, H, W, cond, precision):
    sd = U.random_state_dict(cfg, seed=21)
    net = build(cfg, sd, precision)
    g = torch.Generator().manual_seed(9)
    x = torch.randn((B, cfg["in_channel"], H, W), generator=g)
    if cfg["variant"] == "sr3":
        t = torch.rand((B, 1), generator=g) * 0.9 + 0.05
    else:
        t = torch.rand((1,) if name == "bcast_time" else (B,), generator=g)
    ref = U.unet_forward(sd, cfg, x, t)
    if cond:      # split input: the condition / state concat is never materialised
        y = net(x[:, cond:].to(DEV), t.to(DEV), cond=x[:, :cond].to(DEV))
    else:
        y = net(x.to(DEV), t.to(DEV))
    e, r = relerr(y, ref), relrms(y, ref)
    print(f"[unet {name} {precision}->{net.precision}] vs oracle: max-rel {e:.3e} rel-rms {r:.3e}")
    assert e <= tol(net, precision)[0] and r <= tol(net, precision)[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_unet_follows_its_parameters_to_another_gpu():
    """The library-side weight arenas are bound to the device of the first commit: after ``.to('cuda:1')`` the
    wrapper creates a fresh handle there, and both placements give the same result."""
    cfg = U.make_cfg("ddpm", 1, 1, 16, 16, (1, 2, 4), (), 1, 32)
    sd = U.random_state_dict(cfg, seed=5)
    net = build(cfg, sd, "fp32")
    g = torch.Generator().manual_seed(2)
    x, t = torch.randn((2, 1, 32, 32), generator=g), torch.rand((2,), generator=g)
    y0 = net(x.to("cuda:0"), t.to("cuda:0")).cpu()
    net = net.to("cuda:1")
    y1 = net(x.to("cuda:1"), t.to("cuda:1")).cpu()
    assert torch.equal(y0, y1)
    net = net.to("cuda:0")
    assert torch.equal(net(x.to("cuda:0"), t.to("cuda:0")).cpu(), y0)


@pytest.mark.parametrize("precision", ["fp32", "auto"])
@pytest.mark.parametrize("name,cfg,H,W,cond", [
    ("hagen_512", U.make_cfg("ddpm", 1, 1, 16, 16, (1, 2, 4, 8), (), 1, 32), 512, 512, 0),
    ("splitting_512", U.make_cfg("sr3", 3, 2, 16, 16, (1, 2, 4, 8), (), 1, 512), 512, 512, 1),
    ("sr3_16_128", U.make_cfg("sr3", 6, 3, 64, 32, (1, 2, 4, 8, 8), (16,), 2, 128), 128, 128, 3),
    ("sr3_64_512", U.make_cfg("sr3", 6, 3, 64, 16, (1, 2, 4, 8, 16), (), 1, 512), 512, 512, 3)])
def test_unet_baseline_configs_at_full_resolution(name, cfg, H, W, cond, precision):
    """The BASELINE.json networks at their full resolution (one sample; the CPU oracle needs a few seconds each): the
    large-shape kernel variants (2-D tiles of the fused conv, tall-patch TMA conv, 1024 / 2048-channel layers, attention at
    N = 1024 x C = 1024 and N = 256 x C = 512) inside the whole network, same gates as the small cases."""
    sd = U.random_state_dict(cfg, seed=33)
    net = build(cfg, sd, precision)
    g = torch.Generator().manual_seed(13)
    x = torch.randn((1, cfg["in_channel"], H, W), generator=g)
    t = torch.rand((1, 1), generator=g) * 0.9 + 0.05 if cfg["variant"] == "sr3" else torch.rand((1,), generator=g)
    ref = U.unet_forward(sd, cfg, x, t)
    y = net(x[:, cond:].to(DEV), t.to(DEV), cond=x[:, :cond].to(DEV)) if cond else net(x.to(DEV), t.to(DEV))
    e, r = relerr(y, ref), relrms(y, ref)
    print(f"[unet {name} {precision}->{net.precision} full resolution] vs oracle: max-rel {e:.3e} rel-rms {r:.3e}")
    assert e <= tol(net, precision)[0] and r <= tol(net, precision)[1]


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("case", list(TIMEPRED_CASES))
def test_time_predictor_matches_reference_golden(gold_dir, case, precision):
    """TimePredictor (UNet without time embedding + fused mask / masked-mean tail) vs outputs recorded from the reference's
    class; its state_dict has exactly the reference's keys.  rel-err of the scalar: <= 1e-5 fp32, <= 1e-2 bf16."""
    from diffsplitting_b200.model.time_predictor import TimePredictor
    cfg, B, H, W = TIMEPRED_CASES[case]
    g = np.load(os.path.join(gold_dir, "time_predictor.npz"))
    sd = U.time_predictor_state_dict(cfg, seed=41)
    m = TimePredictor(in_channel=cfg["in_channel"], out_channel=cfg["out_channel"], inner_channel=cfg["inner_channel"],
                      norm_groups=cfg["norm_groups"], channel_mults=cfg["channel_mults"], attn_res=cfg["attn_res"],
                      res_blocks=cfg["res_blocks"], dropout=0.2, image_size=cfg["image_size"], precision=precision)
    assert set(m.state_dict().keys()) == set(sd.keys())
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV)
    y = m(torch.from_numpy(g[f"{case}_x"]).to(DEV)).cpu().numpy()
    ref = g[f"{case}_y"]
    err = float(np.abs(y - ref).max() / np.abs(ref).max())
    print(f"[time predictor {case} {precision}] rel-err {err:.3e}")
    assert y.shape == (B,) and err <= (1e-5 if precision == "fp32" else 1e-2)
    with pytest.raises(RuntimeError):
        m(torch.from_numpy(g[f"{case}_x"]))                       # CPU tensor: no fallback


# ------------------------------------------------------------------------------------------------ samplers
SCHED = dict(schedule="linear", n_timestep=6, linear_start=1e-4, linear_end=0.3)


def _nets():
    cfg = U.make_cfg("sr3", 3, 2, 16, 8, (1, 2), (), 1, 16)
    cfgd = U.make_cfg("ddpm", 3, 2, 16, 8, (1, 2), (), 1, 16)
    cfgi = U.make_cfg("ddpm", 1, 1, 16, 8, (1, 2), (), 1, 16)
    return (cfg, U.random_state_dict(cfg, seed=3)), (cfgd, U.random_state_dict(cfgd, seed=4)), \
        (cfgi, U.random_state_dict(cfgi, seed=7), U.random_state_dict(cfgi, seed=8))


def test_sr3_and_ddpm_steps_match_reference_golden_with_injected_noise(gold_dir):
    gd = np.load(os.path.join(gold_dir, "samplers.npz"))
    (cfg, sd), (cfgd, sdd), _ = _nets()
    cond = torch.from_numpy(gd["sr3_cond"]).to(DEV)
    for cls, c, s, key, always in ((GaussianDiffusionSr3, cfg, sd, "sr3", False), (GaussianDiffusionDdpm, cfgd, sdd, "ddpm", True)):
        netG = cls(build(c, s), 16, channels=2, conditional=True).to(DEV)
        netG.set_new_noise_schedule(SCHED, DEV)
        noise = torch.from_numpy(gd[f"{key}_noise"]).to(DEV)
        out = torch.from_numpy(gd[f"{key}_out"])        # [cond.repeat, x_5, ..., x_0], B = 2
        img = noise[0]
        ni = 1
        for step, t in enumerate(reversed(range(6))):
            z = None
            if always or t > 0:
                z = noise[ni]
                ni += 1
            tt = torch.full((2,), t, device=DEV, dtype=torch.long) if always else t
            img = netG.p_sample(img, tt, condition_x=cond, noise=z if z is not None else torch.zeros_like(img))
            e = float((img.cpu() - out[2 * (step + 1):2 * (step + 2)]).abs().max())
            assert e < 2e-5, (key, t, e)


def test_indi_steps_match_reference_golden_with_injected_noise(gold_dir):
    gd = np.load(os.path.join(gold_dir, "samplers.npz"))
    _, _, (cfgi, sd1, sd2) = _nets()
    x_in = torch.from_numpy(gd["indi_x"]).to(DEV)
    for T in (1, 4):
        indi = InDI(build(cfgi, sd1), 16, channels=1, out_channel=1, conditional=False, val_schedule_opt={"n_timestep": T})
        noise = torch.from_numpy(gd[f"indi_T{T}_noise"]).to(DEV)
        out = torch.from_numpy(gd[f"indi_T{T}_out"])
        x_t = x_in + noise[0] * (0.01 * torch.Tensor([1.0])).to(DEV)
        delta, cur = 1.0 / T, 1.0
        for i in range(T):
            x_t = indi.inference_one_step(x_t, delta, cur, noise=noise[i + 1])
            cur -= delta
            assert float((x_t.cpu() - out[2 * (i + 1):2 * (i + 2)]).abs().max()) < 2e-5


def _cuda_draws(seed, shapes):
    torch.manual_seed(seed)
    return [torch.randn(s, device=DEV).cpu() for s in shapes]


@pytest.mark.parametrize("graph", ["1", "0"])
def test_seeded_loops_replay_the_reference_rng_stream(graph, monkeypatch):
    """Seeded run through the public API (CUDA-graph loop, in-kernel Philox) == oracle fed with the draws
    torch.randn would have produced on this device, and the torch generator ends where the reference leaves it."""
    monkeypatch.setenv("DIFFSPLIT_B200_GRAPH", graph)
    (cfg, sd), (cfgd, sdd), (cfgi, sd1, sd2) = _nets()
    g = torch.Generator().manual_seed(1)
    cond = torch.rand((2, 1, 16, 16), generator=g) * 2 - 1
    torch.cuda.init()                                  # default_generators is empty until CUDA is initialised (test run on its own)
    gen = torch.cuda.default_generators[0]
    tab = S.schedule_tables(SCHED)
    # SR3: initial draw + one per step with t > 0
    netG = GaussianDiffusionSr3(build(cfg, sd), 16, channels=2, conditional=True).to(DEV)
    netG.set_new_noise_schedule(SCHED, DEV)
    draws = _cuda_draws(77, [(2, 2, 16, 16)] * 6)
    end_offset = gen.get_offset()
    torch.manual_seed(77)
    y = netG.super_resolution(cond.to(DEV), continous=True)
    assert gen.get_offset() == end_offset
    ref = S.sr3_sample_loop(tab, lambda x, t: U.unet_forward(sd, cfg, x, t), cond, 2, True, Replay(draws), continous=True)
    assert y.shape == ref.shape and float((y.cpu() - ref).abs().max()) < 5e-5
    last = netG.super_resolution(cond.to(DEV), continous=False)
    assert last.shape == (2, 16, 16)                       # ret_img[-1]: last batch element only (reference quirk)
    # DDPM: draws on every step
    netD = GaussianDiffusionDdpm(build(cfgd, sdd), 16, channels=2, conditional=True).to(DEV)
    netD.set_new_noise_schedule(SCHED, DEV)
    draws = _cuda_draws(78, [(2, 2, 16, 16)] * 7)
    end_offset = gen.get_offset()
    torch.manual_seed(78)
    y = netD.p_sample_loop(cond.to(DEV), continous=True)
    assert gen.get_offset() == end_offset
    ref = S.ddpm_sample_loop(tab, lambda x, t: U.unet_forward(sdd, cfgd, x, t), cond, 2, True, Replay(draws), continous=True)
    assert float((y.cpu() - ref).abs().max()) < 5e-5
    # JointIndi: channel 1 fully, then channel 2
    joint = JointIndi(None, 16, channels=1, out_channel=1, conditional=False, denoise_fn_ch1=build(cfgi, sd1),
                      denoise_fn_ch2=build(cfgi, sd2), val_schedule_opt={"n_timestep": 3}).to(DEV)
    joint.set_new_noise_schedule({"n_timestep": 3}, DEV)
    x_in = torch.rand((2, 1, 16, 16), generator=g) * 2 - 1
    draws = _cuda_draws(79, [(2, 1, 16, 16)] * 8)
    end_offset = gen.get_offset()
    torch.manual_seed(79)
    y = joint.inference(x_in.to(DEV), continuous=True, t_float_start=0.5)
    assert gen.get_offset() == end_offset
    ref = S.joint_indi_inference(lambda x, t: U.unet_forward(sd1, cfgi, x, t), lambda x, t: U.unet_forward(sd2, cfgi, x, t),
                                 x_in, 3, Replay(draws), t_float_start=0.5, continuous=True)
    assert y.shape == ref.shape == (8, 2, 16, 16) and float((y.cpu() - ref).abs().max()) < 5e-5
    for T in (1, 2, 10):                                   # tests/test_joint_indi.py of the reference: T+1 snapshots
        assert joint.inference(x_in[:1].to(DEV), continuous=True, num_timesteps=T).shape[0] == T + 1
    assert joint.inference(x_in.to(DEV)).shape == (1, 2, 16, 16)          # ret_img[-1:]
    # repeated call with the same seed through the cached graph gives the same bits
    torch.manual_seed(79)
    y2 = joint.inference(x_in.to(DEV), continuous=True, t_float_start=0.5)
    assert torch.equal(y, y2)


def test_ddpm_p_sample_with_a_different_timestep_per_sample():
    """ddpm_modules/diffusion.py:64-67,195-203: `extract` gathers every buffer with the (B,) tensor t, so the samples of a
    batch may sit at different timesteps (incl. t == 0, whose noise is masked)."""
    _, (cfgd, sdd), _ = _nets()
    netD = GaussianDiffusionDdpm(build(cfgd, sdd), 16, channels=2, conditional=True).to(DEV)
    netD.set_new_noise_schedule(SCHED, DEV)
    g = torch.Generator().manual_seed(21)
    cond = torch.rand((3, 1, 16, 16), generator=g) * 2 - 1
    x = torch.randn((3, 2, 16, 16), generator=g)
    z = torch.randn((3, 2, 16, 16), generator=g)
    t = torch.tensor([4, 0, 2])
    tab = S.schedule_tables(SCHED)
    ref = S.ddpm_p_sample(tab, lambda a, tt: U.unet_forward(sdd, cfgd, a, tt), x, t, cond, True, z)
    y = netD.p_sample(x.to(DEV), t.to(DEV), condition_x=cond.to(DEV), noise=z)
    assert float((y.cpu() - ref).abs().max()) < 2e-5
    # generated noise: the same draw torch.randn would produce, one generator increment consumed
    gen = torch.cuda.default_generators[0]
    zd = _cuda_draws(31, [(3, 2, 16, 16)])[0]
    end = gen.get_offset()
    torch.manual_seed(31)
    y2 = netD.p_sample(x.to(DEV), t.to(DEV), condition_x=cond.to(DEV))
    assert gen.get_offset() == end
    ref2 = S.ddpm_p_sample(tab, lambda a, tt: U.unet_forward(sdd, cfgd, a, tt), x, t, cond, True, zd)
    assert float((y2.cpu() - ref2).abs().max()) < 2e-5


@pytest.mark.parametrize("numel,offset_elems", [(1003, 0), (4096 + 5, 1), (3 * 512 * 512, 0), (1 << 20, 3)])
def test_sampler_update_vectorised_and_scalar_paths_agree_bitwise(numel, offset_elems):
    """ds_sampler_step: the 16-byte path (four torch threads per CUDA thread) and the scalar path (taken for unaligned
    tensors) produce the same bits as the reference expression evaluated by torch on torch's own draw."""
    gen = torch.cuda.default_generators[0]
    sms, max_thr, _, _ = _lib.device_info(0)
    grid = min(sms * (max_thr // 256), (numel + 255) // 256)
    threads, inc = 256 * grid, ((numel - 1) // (256 * grid * 4) + 1) * 4
    torch.manual_seed(99)
    z = torch.randn(numel, device=DEV)
    g = torch.Generator().manual_seed(5)
    buf_x = torch.randn(numel + 8, generator=g).to(DEV)
    buf_n = torch.randn(numel + 8, generator=g).to(DEV)
    x, n = buf_x[offset_elems:offset_elems + numel], buf_n[offset_elems:offset_elems + numel]
    coef = torch.tensor([[1.25, 0.75, 0.3, 0.7, 0.11]], device=DEV)
    for mode, clip in ((0, 1), (1, 0)):
        x0 = (coef[0, 0] * x - coef[0, 1] * n).clamp(-1, 1) if mode == 0 else n
        ref = (coef[0, 2] * x0 + coef[0, 3] * x) + z * coef[0, 4]
        out = torch.empty(numel + 8, device=DEV)[offset_elems:offset_elems + numel]
        a = _lib.StepArgs()
        a.d_x = x.data_ptr(); a.d_net = n.data_ptr(); a.d_out = out.data_ptr(); a.numel = numel
        a.mode = mode; a.clip = clip; a.d_coef = coef.data_ptr(); a.n_steps = 1; a.step = 0; a.d_state = None
        torch.manual_seed(99)
        a.seed = gen.initial_seed(); a.offset = gen.get_offset(); a.offset_inc = inc; a.rng_threads = threads
        _lib.check(_lib.lib().ds_sampler_step(C.byref(a), sptr()))
        torch.cuda.synchronize()
        assert torch.equal(out, ref), (numel, offset_elems, mode)
        # injected noise, same paths
        a.d_noise = z.data_ptr()
        out.zero_()
        _lib.check(_lib.lib().ds_sampler_step(C.byref(a), sptr()))
        torch.cuda.synchronize()
        assert torch.equal(out, ref)


def test_joint_indi_concurrent_branches_equal_the_serial_loops(monkeypatch):
    """joint_indi.py:132-135 runs its two InDI loops one after the other; here they are the two branches of one CUDA graph.
    Same seed -> same result as the serial order (DIFFSPLIT_B200_JOINT_SERIAL=1), same generator end state, and
    `all_samples` returns every batch element where the reference's `ret_img[-1:]` keeps the last one."""
    cfg = U.make_cfg("ddpm", 1, 1, 16, 16, (1, 2, 4, 8), (), 1, 32)
    nets = [build(cfg, U.random_state_dict(cfg, seed=s), "fp32") for s in (3, 4)]
    joint = JointIndi(None, 32, channels=1, out_channel=1, conditional=False, denoise_fn_ch1=nets[0], denoise_fn_ch2=nets[1],
                      val_schedule_opt={"n_timestep": 5}).to(DEV)
    joint.set_new_noise_schedule({"n_timestep": 5}, DEV)
    x = (torch.rand((3, 1, 32, 32), generator=torch.Generator().manual_seed(0)) * 2 - 1).to(DEV)
    gen = torch.cuda.default_generators[0]
    torch.manual_seed(5)
    y = joint.inference(x, continuous=True, t_float_start=0.3)
    end = gen.get_offset()
    torch.manual_seed(5)
    ya = joint.inference(x, continuous=False, t_float_start=0.3, all_samples=True)
    torch.manual_seed(5)
    yl = joint.inference(x, continuous=False, t_float_start=0.3)
    monkeypatch.setenv("DIFFSPLIT_B200_JOINT_SERIAL", "1")
    torch.manual_seed(5)
    ys = joint.inference(x, continuous=True, t_float_start=0.3)
    assert gen.get_offset() == end
    assert y.shape == ys.shape and float((y - ys).abs().max()) < 1e-5
    assert ya.shape == (3, 2, 32, 32) and float((ya - ys[-3:]).abs().max()) < 1e-5
    assert yl.shape == (1, 2, 32, 32) and torch.equal(yl, ya[-1:])
    with pytest.raises(ValueError, match="t_float_start"):
        joint.indi1.inference(x, t_float_start=0.0)


@pytest.mark.parametrize("T", [1, 2, 10, 21])
def test_joint_indi_snapshot_count_and_multi_step_graphs(T):
    """The intent of the reference's tests/test_joint_indi.py:9-25 (continuous=True returns n_timestep + 1 snapshots per
    batch element for small T), run through JointIndi with two real UNets; T = 10 / 21 also cross the 8-step CUDA graphs
    of `_Engine.run`, whose result must equal stepping one graph replay at a time."""
    from diffsplitting_b200.model.samplers import JointIndi
    cfg = U.make_cfg("ddpm", 1, 1, 16, 16, (1, 2, 4, 8), (), 1, 32)
    nets = [build(cfg, U.random_state_dict(cfg, seed=s), "bf16") for s in (3, 4)]
    joint = JointIndi(None, 32, channels=1, out_channel=1, conditional=False, denoise_fn_ch1=nets[0], denoise_fn_ch2=nets[1],
                      val_schedule_opt={"n_timestep": T}).to(DEV)
    joint.set_new_noise_schedule({"n_timestep": T}, DEV)
    x = (torch.rand((2, 1, 32, 32), generator=torch.Generator().manual_seed(0)) * 2 - 1).to(DEV)
    inter = 1 | (T // 20)
    n_snap = len([i for i in range(T) if i % inter == 0 or i == T - 1])
    torch.manual_seed(5)
    y = joint.inference(x, continuous=True)
    assert tuple(y.shape) == ((n_snap + 1) * 2, 2, 32, 32) and torch.isfinite(y).all()
    os.environ["DIFFSPLIT_B200_GRAPH_STEPS"] = "1"
    try:
        torch.manual_seed(5)
        y1 = joint.inference(x, continuous=True)
    finally:
        del os.environ["DIFFSPLIT_B200_GRAPH_STEPS"]
    assert torch.equal(y, y1)


def test_mmse_tiled_evaluation_follows_the_notebook_loop():
    """evaluate_mmse == the loop of notebooks/EvaluateJointIndi.ipynb cells 55-62 written out literally (batch 1, both
    full JointIndi.inference calls per tile, `mmse_pred += pred / mmse_count`, stitch, RangeInvariantPsnr): bit-identical
    under the same seed with chunk=1 / replay_reference_rng, and the device metric agrees with the float64 oracle on the
    stitched frames.  The batched fast path (chunk 4, only the two kept loops) gives the same shapes and finite values."""
    from diffsplitting_b200.evaluate import evaluate_mmse
    cfg = U.make_cfg("ddpm", 1, 1, 16, 16, (1, 2), (), 1, 32)
    nets = [build(cfg, U.random_state_dict(cfg, seed=s), "bf16") for s in (3, 4)]
    joint = JointIndi(None, 32, channels=1, out_channel=1, conditional=False, denoise_fn_ch1=nets[0], denoise_fn_ch2=nets[1],
                      val_schedule_opt={"n_timestep": 2}).to(DEV)
    joint.set_new_noise_schedule({"n_timestep": 2}, DEV)
    rng = np.random.default_rng(3)
    frames = rng.integers(0, 2000, size=(2, 2, 64, 64)).astype(np.uint16)
    nd = {"mean_input": np.float64(1000.0), "std_input": np.float64(1000.0), "mean_target": np.array([500.0, 500.0]),
          "std_target": np.array([500.0, 500.0])}
    tf = TiledFrames(frames, 32, 16, normalization_dict=nd)
    W_, T_, M_ = 0.5, 2, 2
    torch.manual_seed(11)
    res = evaluate_mmse(joint, tf, mixing_t=W_, num_timesteps=T_, mmse_count=M_, chunk=1, replay_reference_rng=True)
    # the notebook loop
    torch.manual_seed(11)
    runs, targets = [], []
    for m in range(M_):
        preds = []
        for i in range(len(tf)):
            data = tf[i]
            inp0 = data["target"][:1] * (1 - W_) + data["target"][1:2] * W_
            inp1 = data["target"][1:2] * (1 - W_) + data["target"][:1] * W_
            p0 = joint.inference(torch.Tensor(inp0[None]).to(DEV), continuous=False, t_float_start=W_, num_timesteps=T_)[:, 0]
            p1 = joint.inference(torch.Tensor(inp1[None]).to(DEV), continuous=False, t_float_start=W_, num_timesteps=T_)[:, 1]
            preds.append(torch.stack([p0, p1], dim=1).cpu().numpy())
            if m == 0:
                targets.append(data["target"][None])
        runs.append(np.concatenate(preds, axis=0))
    mmse = 0
    for m in range(M_):
        mmse += runs[m] / M_
    pred_st = stitch_predictions(mmse, tf.tile_manager)
    tar_st = stitch_predictions(np.concatenate(targets, axis=0), tf.tile_manager)
    assert np.array_equal(res["prediction"].cpu().numpy(), pred_st) and np.array_equal(res["target"].cpu().numpy(), tar_st)
    for c in range(2):
        ref = M.range_invariant_psnr(tar_st[..., c], pred_st[..., c])
        assert np.allclose(res["range_invariant_psnr"][:, c].cpu().numpy(), ref, rtol=0, atol=1e-4)
    fast = evaluate_mmse(joint, tf, mixing_t=W_, num_timesteps=T_, mmse_count=M_, chunk=4)
    assert fast["prediction"].shape == res["prediction"].shape and torch.isfinite(fast["range_invariant_psnr"]).all()
    assert torch.equal(fast["target"], res["target"])
    # batched chunks predict EVERY tile of the chunk (continuous=False alone keeps only the last batch element): with the
    # noise switched off (e = 0) the tile-by-tile result cannot depend on the chunk size
    for ind in (joint.indi1, joint.indi2):
        ind.e = 0.0
    from diffsplitting_b200.evaluate import mmse_predict_tiles
    p1, _ = mmse_predict_tiles(joint, tf, mixing_t=W_, num_timesteps=T_, mmse_count=1, chunk=1)
    p4, _ = mmse_predict_tiles(joint, tf, mixing_t=W_, num_timesteps=T_, mmse_count=1, chunk=4)
    p3, _ = mmse_predict_tiles(joint, tf, mixing_t=W_, num_timesteps=T_, mmse_count=1, chunk=3, replay_reference_rng=True)
    scale = float(p1.abs().max())
    assert float((p1 - p4).abs().max()) < 2e-2 * scale and float((p1 - p3).abs().max()) < 2e-2 * scale
    assert float((p1[0] - p1[1]).abs().max()) > 0.05 * scale          # tiles do differ


def test_step_rate_against_torch_eager_on_the_same_gpu():
    """'the existing Blackwell path' = the reference's PyTorch ops run eagerly on the B200 (cuDNN convs with
    TF32 allowed, native_group_norm, ~225 launches per step).  The oracle IS those ops, so it is moved to the GPU and timed
    beside the engine on the bench workload (hagen InDI 64x64, batch 16).  Printed with -s; the engine must be >= 3x faster
    (measured ~2200 vs ~270 steps/s)."""
    cfg = U.make_cfg("ddpm", 1, 1, 16, 16, (1, 2, 4, 8), (), 1, 32)
    sd = U.random_state_dict(cfg, seed=1)
    sd_dev = {k: v.to(DEV) for k, v in sd.items()}
    B, T = 16, 1000
    x0 = (torch.rand((B, 1, 64, 64), generator=torch.Generator().manual_seed(0)) * 2 - 1).to(DEV)
    den = lambda xx, tt: U.unet_forward(sd_dev, cfg, xx, tt)
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        def eager_steps(n):
            x, t = x0.clone(), 1.0
            for _ in range(n):
                t32 = torch.Tensor([t]).to(DEV)                       # indi.py:62-69 with the H2D of :65
                d = 1.0 / T
                x = d / t32 * den(x, t32) + (1 - d / t32) * x + torch.randn_like(x) * (0.01 * (t32 - d))
                t -= d
            return x
        with torch.no_grad():
            eager_steps(5)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eager_steps(40)
            e1.record()
            torch.cuda.synchronize()
        eager_rate = 40 / (e0.elapsed_time(e1) * 1e-3)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    net = build(cfg, sd, "bf16")
    indi = InDI(net, 32, channels=1, out_channel=1, conditional=False, val_schedule_opt={"n_timestep": T}).to(DEV)
    indi.set_new_noise_schedule({"n_timestep": T}, DEV)
    indi.inference(x0, num_timesteps=400)                     # builds the step graphs
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    indi.inference(x0, num_timesteps=400)
    e1.record()
    torch.cuda.synchronize()
    rate = 400 / (e0.elapsed_time(e1) * 1e-3)
    print(f"[same-GPU baseline] torch eager (TF32) {eager_rate:.0f} steps/s, engine {rate:.0f} steps/s, x{rate / eager_rate:.1f}")
    assert rate >= 3 * eager_rate


@pytest.mark.parametrize("T", [1, 5, 16])
def test_final_psnr_within_point1_db_of_oracle(T):
    """north_star: final split channels within 0.1 dB PSNR of the reference for the same seeds, in both precision modes.
    JointIndi chains of T = 1 and T = 5 steps - the operating points the reference publishes (notebooks/EvaluateJointIndi.ipynb
    :1868, :1982) - plus T = 16 (T = 20 would trip the reference's own `delta_t <= t_cur` assertion through float accumulation).
    The gate is taken where it matters: the target is the ORACLE's prediction plus noise sized so that the oracle scores 34 dB
    against it (the published 33.8 / 36.0 dB) - against an unrelated random target both scores would be noise-vs-noise and
    insensitive to the kernels.  Also asserted directly: PSNR(ours, oracle).  A random-weight UNet is expansive (a per-step
    rel-RMS of ~1.3e-3 grows ~20x over 16 chained steps), so T = 16 is held to 0.5 dB; the published T pass the 0.1 dB gate."""
    cfgi = U.make_cfg("ddpm", 1, 1, 16, 16, (1, 2, 4, 8), (), 1, 32)
    sd1, sd2 = U.random_state_dict(cfgi, seed=7), U.random_state_dict(cfgi, seed=8)
    g = torch.Generator().manual_seed(4)
    x_in = torch.rand((2, 1, 64, 64), generator=g) * 2 - 1
    draws = _cuda_draws(5, [(2, 1, 64, 64)] * (2 * (T + 1)))
    ref = S.joint_indi_inference(lambda x, t: U.unet_forward(sd1, cfgi, x, t), lambda x, t: U.unet_forward(sd2, cfgi, x, t),
                                 x_in, T, Replay(draws), continuous=True)[-2:]
    rng = (ref.reshape(2, 2, -1).max(dim=2).values - ref.reshape(2, 2, -1).min(dim=2).values).reshape(2, 2, 1, 1)
    target = ref + torch.randn(ref.shape, generator=g) * rng / 10 ** (34.0 / 20.0)
    gate = 0.1 if T <= 5 else 0.5
    for precision in ("fp32", "auto"):
        joint = JointIndi(None, 32, channels=1, out_channel=1, conditional=False, denoise_fn_ch1=build(cfgi, sd1, precision),
                          denoise_fn_ch2=build(cfgi, sd2, precision), val_schedule_opt={"n_timestep": T}).to(DEV)
        joint.set_new_noise_schedule({"n_timestep": T}, DEV)
        torch.manual_seed(5)
        y = joint.inference(x_in.to(DEV), continuous=True)[-2:].cpu()
        for c in range(2):
            p_ref, p_our = S.psnr(target[:, c], ref[:, c]), S.psnr(target[:, c], y[:, c])
            direct = S.psnr(ref[:, c], y[:, c])
            d = (p_our - p_ref).abs().max()
            print(f"[psnr T={T} {precision}->{joint.indi1.denoise_fn.precision}] ch{c}: oracle {[round(v, 3) for v in p_ref.tolist()]} dB, "
                  f"ours {[round(v, 3) for v in p_our.tolist()]} dB, delta {float(d):.4f} dB; PSNR(ours, oracle) {[round(v, 1) for v in direct.tolist()]} dB")
            assert 33.5 < float(p_ref.min()) and float(p_ref.max()) < 34.5
            assert float(d) < gate
            assert float(direct.min()) > (80.0 if precision == "fp32" else (50.0 if T <= 5 else 42.0))


def test_sr3_step_at_the_noisiest_timestep():
    """sr3_modules/diffusion.py:141-168 at t = T-1 of the shipped schedule (linear 1e-6..1e-2, T = 2000): x0_hat = c1 x - c2 eps
    with c1 ~ c2 ~ 151, i.e. the UNet's error is amplified 151x before the clamp.  The state and the update stay fp32, so the
    step must still match the oracle: tightly with clip_denoised (the reference's default), and within the amplified
    network tolerance without it."""
    cfg = U.make_cfg("sr3", 3, 2, 16, 8, (1, 2), (), 1, 16)
    sd = U.random_state_dict(cfg, seed=3)
    sched = dict(schedule="linear", n_timestep=2000, linear_start=1e-6, linear_end=1e-2)
    tab = S.schedule_tables(sched)
    assert 140 < float(np.sqrt(1.0 / np.cumprod(1 - np.linspace(1e-6, 1e-2, 2000))[-1])) < 160
    g = torch.Generator().manual_seed(2)
    cond = torch.rand((2, 1, 16, 16), generator=g) * 2 - 1
    x = torch.randn((2, 2, 16, 16), generator=g)
    z = torch.randn((2, 2, 16, 16), generator=g)
    for precision in ("fp32", "auto"):
        netG = GaussianDiffusionSr3(build(cfg, sd, precision), 16, channels=2, conditional=True).to(DEV)
        netG.set_new_noise_schedule(sched, DEV)
        for clip in (True, False):
            ref = S.sr3_p_sample(tab, lambda a, t: U.unet_forward(sd, cfg, a, t), x, 1999, cond, clip, z)
            y = netG.p_sample(x.to(DEV), 1999, clip_denoised=clip, condition_x=cond.to(DEV), noise=z).cpu()
            e = float((y - ref).abs().max() / ref.abs().max())
            print(f"[sr3 step t=T-1 {precision} clip={clip}] max err / max|ref| {e:.3e}")
            assert e <= (2e-5 if precision == "fp32" else 1e-2)


# ------------------------------------------------------------------------------------------------ tiling
@pytest.mark.parametrize("data,grid,patch", [((5, 512, 512), (1, 128, 128), (1, 256, 256)), ((3, 100, 130), (1, 16, 16), (1, 32, 32)),
                                             ((2, 64, 64), (1, 64, 64), (1, 64, 64)), ((4, 70, 70), (1, 32, 32), (1, 64, 64)),
                                             ((1, 8, 8), (1, 2, 2), (1, 8, 8))])
def test_crop_and_stitch_bit_exact_vs_oracle(data, grid, patch):
    tg = TR.TileGrid(data, grid, patch, TR.SHIFT)
    mgr = TileIndexManager(data, grid, patch, TilingMode.ShiftBoundary)
    rng = np.random.default_rng(3)
    frames = rng.standard_normal((2,) + data).astype(np.float32)
    tf = TiledFrames(frames, patch[1], grid[1])
    assert len(tf) == tg.total
    tiles = tf.raw_tiles(0, tg.total)
    assert np.array_equal(tiles.cpu().numpy(), TR.crop_tiles(frames, tg))
    assert np.array_equal(tf.raw_tiles(1, 1).cpu().numpy(), TR.crop_tiles(frames, tg, [1])) if tg.total > 1 else True
    noise_tiles = rng.standard_normal(tiles.shape).astype(np.float32)
    out = stitch_predictions(noise_tiles, mgr)             # numpy in -> numpy out, like the reference
    assert isinstance(out, np.ndarray) and np.array_equal(out, TR.stitch(noise_tiles, tg))
    out_t = stitch_predictions(tiles, mgr)                 # device in -> device out
    assert torch.equal(out_t.cpu(), torch.from_numpy(frames.transpose(1, 2, 3, 0)))


@pytest.mark.parametrize("data,grid,patch,mode", [((3, 100, 130), (1, 16, 16), (1, 32, 32), TilingMode.ShiftBoundary),
                                                  ((2, 96, 96), (1, 16, 16), (1, 32, 32), TilingMode.TrimBoundary),
                                                  ((2, 1024, 1024), (1, 256, 256), (1, 512, 512), TilingMode.ShiftBoundary)])
@pytest.mark.parametrize("world", [1, 3, 8])
def test_packed_region_exchange_stitches_to_the_same_bits(data, grid, patch, mode, world):
    """The multi-GPU exchange format (parallel.py): every emulated rank packs the destination boxes of ITS round-robin share
    of random tiles, the buffers are laid side by side as the gather would, and `ds_stitch_packed` must give exactly what
    `stitch_predictions` (pinned to the reference, bit-exact) gives on the whole tiles - for any number of ranks."""
    from diffsplitting_b200.parallel import PackedLayout, pack_tile_regions, stitch_packed
    mng = TileIndexManager(data, grid, patch, mode)
    total = mng.total_grid_count()
    tiles = torch.randn((total, 2, patch[1], patch[2]), generator=torch.Generator().manual_seed(3)).to(DEV)
    ref = stitch_predictions(tiles, mng)
    lay = PackedLayout(mng, 2, 4, world)
    gathered = torch.zeros((world * lay.slot,), device=DEV)
    for r in range(world):
        if len(lay.ids[r]):
            pack_tile_regions(tiles[torch.from_numpy(lay.ids[r]).to(DEV)], mng, lay.ids[r], lay.local_off[r],
                              gathered[r * lay.slot:(r + 1) * lay.slot])
    out = stitch_packed(gathered, mng, lay.global_off, 2)
    assert torch.equal(out, ref)
    assert lay.payload_bytes <= tiles.numel() * 4


def test_stitch_golden_and_trim_mode(gold_dir):
    gd = np.load(os.path.join(gold_dir, "tiling.npz"))
    tg = TR.TileGrid((3, 100, 130), (1, 16, 16), (1, 32, 32), TR.SHIFT)
    rng = np.random.default_rng(0)
    rng.standard_normal((45, 2, 256, 256)).astype(np.float32)
    tiles = rng.standard_normal((tg.total, 2, 32, 32)).astype(np.float32)
    mgr = TileIndexManager((3, 100, 130), (1, 16, 16), (1, 32, 32), TilingMode.ShiftBoundary)
    assert np.array_equal(stitch_predictions(tiles, mgr), gd["stitch_ragged_out"])      # reference's own output
    tgt = TR.TileGrid((2, 80, 80), (1, 16, 16), (1, 32, 32), TR.TRIM)
    mt = TileIndexManager((2, 80, 80), (1, 16, 16), (1, 32, 32), TilingMode.TrimBoundary)
    tl = rng.standard_normal((tgt.total, 1, 32, 32)).astype(np.float32)
    assert np.array_equal(stitch_predictions(tl, mt), TR.stitch(tl, tgt))               # uncovered border stays 0
    with pytest.raises(ValueError):
        stitch_predictions(tl[:-1], mt)


def test_full_size_tiling_identity_uint16():
    """BASELINE size: 10 x 2048^2 x 2 channels, 490 tiles of 512^2: stitch(crop(frames)) == frames, bit-exact, and
    normalise+mix equals the numpy expression of SplitDataset (float64 constants, one rounding)."""
    rng = np.random.default_rng(1)
    frames = rng.integers(0, 1994, size=(2, 10, 2048, 2048), dtype=np.uint16)
    nd = {"mean_input": 1210.5, "std_input": 1210.5, "mean_target": np.array([600.25, 610.25]), "std_target": np.array([600.25, 610.25])}
    tf = TiledFrames(frames, 512, 256, normalization_dict=nd)
    assert len(tf) == 490
    tiles = tf.raw_tiles(0, 490)
    out = stitch_predictions(tiles, tf.tile_manager)
    assert out.shape == (10, 2048, 2048, 2)
    assert torch.equal(out.cpu(), torch.from_numpy(frames.astype(np.float32)).permute(1, 2, 3, 0))
    tg = TR.TileGrid((10, 2048, 2048), (1, 256, 256), (1, 512, 512), TR.SHIFT)
    idx = [0, 6, 48, 489]
    raw = TR.crop_tiles(frames, tg, idx)
    inp_ref, tar_ref = TR.normalise_and_mix(raw, nd["mean_target"], nd["std_target"], nd["mean_input"], nd["std_input"])
    for j, i in enumerate(idx):
        inp, tar = tf.batch(i, 1)
        assert np.array_equal(inp.cpu().numpy()[0], inp_ref[j]) and np.array_equal(tar.cpu().numpy()[0], tar_ref[j])
        item = tf[i]
        assert np.array_equal(item["input"], inp_ref[j])


def test_fused_tile_batch_matches_reference_dataset_items(gold_dir):
    """ds_tile_batch (crop + float64 normalise + mix in one kernel) vs items recorded from the reference's
    SplitDatasetTiledPred, both input modes, uint16 and fp32 frames: bit-exact."""
    g = np.load(os.path.join(gold_dir, "metrics.npz"))
    fr, idx = g["tiles_frames"], [int(i) for i in g["tiles_idx"]]
    nd = {"mean_input": np.float64(1210.4), "std_input": np.float64(1207.3), "mean_target": np.array([600.3, 610.7]),
          "std_target": np.array([598.9, 611.1])}
    for frames in (fr, fr.astype(np.float32)):
        for tag, w, fn in (("mix", (0.7, 0.4), False), ("normtar", (0.5, 0.5), True)):
            tf = TiledFrames(frames, 32, 16, normalization_dict=nd, channel_weights=w, input_from_normalized_target=fn)
            for j, i in enumerate(idx):
                inp, tar = tf.batch(i, 1)
                assert np.array_equal(inp.cpu().numpy()[0], g[f"tiles_{tag}_input"][j])
                assert np.array_equal(tar.cpu().numpy()[0], g[f"tiles_{tag}_target"][j])
            n = len(tf)
            inp, tar = tf.batch(0, n)                                                  # the whole range in one launch
            tg = TR.TileGrid((3, 96, 128), (1, 16, 16), (1, 32, 32), TR.SHIFT)
            inp_ref, tar_ref = TR.normalise_and_mix(TR.crop_tiles(fr, tg), nd["mean_target"], nd["std_target"],
                                                    nd["mean_input"], nd["std_input"], w, fn)
            assert np.array_equal(inp.cpu().numpy(), inp_ref) and np.array_equal(tar.cpu().numpy(), tar_ref)
    with pytest.raises(RuntimeError):
        TiledFrames(fr, 32, 16, normalization_dict=nd).batch(len(tf) - 1, 2)           # range past the last tile


def test_psnr_kernel_matches_reference_golden(gold_dir):
    """ds_psnr vs values recorded from core/psnr.py (float32 torch): <= 5e-4 dB; fixed range_; numpy inputs."""
    from diffsplitting_b200.core.psnr import PSNR, RangeInvariantPsnr
    g = np.load(os.path.join(gold_dir, "metrics.npz"))
    for tag in ("unit", "u16", "offset"):
        gt, pr = torch.from_numpy(g[f"{tag}_gt"]).to(DEV), torch.from_numpy(g[f"{tag}_pred"]).to(DEV)
        assert np.allclose(PSNR(gt, pr).cpu().numpy(), g[f"{tag}_psnr"], rtol=0, atol=5e-4)
        assert np.allclose(RangeInvariantPsnr(gt, pr).cpu().numpy(), g[f"{tag}_ripsnr"], rtol=0, atol=5e-4)
    assert np.allclose(PSNR(g["unit_gt"], g["unit_pred"], range_=2.0).cpu().numpy(), g["unit_psnr_range2"], atol=5e-4)
    gt = torch.from_numpy(g["unit_gt"]).to(DEV)
    assert torch.isinf(PSNR(gt, gt)).all()
    with pytest.raises(AssertionError):
        PSNR(gt[0], gt[0])


def test_psnr_frames_in_place_on_stitched_layout(gold_dir):
    """Whole-frame metrics read from the stitched (F,H,W,C) layout with the validation loop's un-normalisation and uint16
    cast folded in (split.py:198-208), vs the float64 oracle: <= 1e-4 dB; and the golden value of the reference's loop."""
    from diffsplitting_b200.core.psnr import psnr_frames
    g = np.load(os.path.join(gold_dir, "metrics.npz"))
    mean, std = g["val_mean"], g["val_std"]
    t = torch.from_numpy(g["val_target"]).to(DEV)[None]                               # (1,C,H,W)
    p = torch.from_numpy(g["val_prediction"]).to(DEV)[None]
    ps, _ = psnr_frames(t, p, mean, std, channels_last=False)
    assert np.allclose(ps.cpu().numpy()[0], g["val_psnr"], atol=5e-4)
    ps_cl, _ = psnr_frames(t.permute(0, 2, 3, 1).contiguous(), p.permute(0, 2, 3, 1).contiguous(), mean, std)
    assert torch.equal(ps, ps_cl)
    # full-size frames: 3 x 1024^2 x 2 channels, many CTAs per image
    rng = np.random.default_rng(11)
    tgt = rng.uniform(-1, 1, size=(3, 1024, 1024, 2)).astype(np.float32)
    prd = (0.9 * tgt + 0.02 + 0.1 * rng.standard_normal(tgt.shape)).astype(np.float32)
    ps, ri = psnr_frames(torch.from_numpy(tgt).to(DEV), torch.from_numpy(prd).to(DEV), mean, std)
    ps_raw, ri_raw = psnr_frames(torch.from_numpy(tgt).to(DEV), torch.from_numpy(prd).to(DEV))
    for f in range(3):
        tu = M.unnormalize_u16(tgt[f].transpose(2, 0, 1), mean, std, False)
        pu = M.unnormalize_u16(prd[f].transpose(2, 0, 1), mean, std, True)
        assert np.allclose(ps[f].cpu().numpy(), M.psnr(tu, pu), rtol=0, atol=1e-4)
        assert np.allclose(ri[f].cpu().numpy(), M.range_invariant_psnr(tu, pu), rtol=0, atol=1e-4)
        assert np.allclose(ps_raw[f].cpu().numpy(), M.psnr(tgt[f].transpose(2, 0, 1), prd[f].transpose(2, 0, 1)), atol=1e-4)
        assert np.allclose(ri_raw[f].cpu().numpy(), M.range_invariant_psnr(tgt[f].transpose(2, 0, 1), prd[f].transpose(2, 0, 1)), atol=1e-4)
    # invariance property of RangeInvariantPsnr: affine changes of the prediction do not move it
    _, ri2 = psnr_frames(torch.from_numpy(tgt).to(DEV), torch.from_numpy(3.0 * prd + 0.5).to(DEV))
    assert torch.allclose(ri2, ri_raw, atol=1e-3)


# ------------------------------------------------------------------------------------------------ wrapper
def test_create_model_test_and_checkpoint_round_trip(tmp_path):
    opt = make_opt("splitting_hagen_indi_joint")
    opt["path"]["checkpoint"] = str(tmp_path)
    m = create_model(opt)
    m.set_new_noise_schedule(opt["model"]["beta_schedule"]["val"], schedule_phase="val")
    g = torch.Generator().manual_seed(0)
    data = {"input": torch.rand((1, 1, 64, 64), generator=g), "target": torch.rand((1, 2, 64, 64), generator=g)}
    m.feed_data(data)
    torch.manual_seed(3)
    m.test(continuous=True)
    vis = m.get_current_visuals()
    assert vis["prediction"].shape == (4, 2, 64, 64) and vis["prediction"].device.type == "cpu"
    assert vis["input"].shape == (1, 1, 64, 64) and torch.isfinite(vis["prediction"]).all()
    m.save_network(1, 10)
    opt2 = make_opt("splitting_hagen_indi_joint")
    opt2["path"]["resume_state"] = os.path.join(str(tmp_path), "I10_E1")
    m2 = create_model(opt2)
    m2.set_new_noise_schedule(opt["model"]["beta_schedule"]["val"], schedule_phase="val")
    m2.feed_data(data)
    torch.manual_seed(3)
    m2.test(continuous=True)
    assert torch.equal(m2.get_current_visuals()["prediction"], vis["prediction"])
