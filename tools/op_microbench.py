"""Warm micro-benchmark of single operators through the C ABI (CUDA-graph of 50 launches, replayed): real per-launch
GPU time without host launch overhead.  Usage: python tools/op_microbench.py [gnconv|conv_tc|gn|sampler] ca cb cout ks B H W
       python tools/op_microbench.py attn 0 0 0 0 B N C | psnr 0 0 0 0 F H W | tilebatch 0 0 0 0 F H W"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffsplitting_b200 import _lib  # noqa: E402

DEV = "cuda"


def timed_graph(fn, n=50, reps=20):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000 / (n * reps)


def data_path(kind, F, H, W):
    """psnr: F frames of HxW, 2 channels, stitched (F,H,W,C) layout, un-normalise + uint16 cast folded in.
    tilebatch: all tiles of 512^2 (grid 256) from F uint16 frames of HxW."""
    import numpy as np
    from diffsplitting_b200.core.psnr import psnr_frames
    from diffsplitting_b200.data import TiledFrames
    if kind == "psnr":
        t = torch.rand((F, H, W, 2), device=DEV) * 2 - 1
        p = t + 0.05 * torch.randn_like(t)
        mean = std = np.array([600.25, 610.75])
        us = timed_graph(lambda: psnr_frames(t, p, mean, std), n=10, reps=10)
        nbytes = 2 * t.numel() * 4
        print(f"psnr_frames {F}x{H}x{W}x2 (2 kernels, both metrics, both channels): {us:.1f} us per call, "
              f"{nbytes / us / 1e6:.2f} TB/s algorithmic (read target + prediction once)")
    else:
        fr = torch.randint(0, 2000, (2, F, H, W), dtype=torch.int32, device=DEV).to(torch.uint16)
        nd = {"mean_input": 1210.5, "std_input": 1210.5, "mean_target": np.array([600.25, 610.25]),
              "std_target": np.array([600.25, 610.25])}
        tf = TiledFrames(fr, 512, 256, normalization_dict=nd)
        n = len(tf)
        us = timed_graph(lambda: tf.batch(0, n), n=5, reps=10)
        nbytes = n * 512 * 512 * (2 * 2 + 3 * 4)
        print(f"tile_batch {n} tiles of 512^2 from {F}x{H}x{W} uint16 frames: {us:.1f} us per call, "
              f"{nbytes / us / 1e6:.2f} TB/s algorithmic (2 uint16 reads + 3 fp32 writes per tile pixel)")


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "gnconv"
    ca, cb, cout, ks, B, H, W = [int(v) for v in (sys.argv[2:9] if len(sys.argv) >= 9 else "16 0 16 3 16 64 64".split())]
    cin = ca + cb
    L = _lib.lib()
    if kind in ("psnr", "tilebatch"):
        return data_path(kind, B, H, W)
    if kind == "attn":                      # attn 0 0 0 0 B N C
        N, Cc = H, W
        qkv = torch.randn((B, N, 3 * Cc), generator=torch.Generator().manual_seed(0)).to(DEV).to(torch.bfloat16)
        o = torch.empty((B, N, Cc), dtype=torch.bfloat16, device=DEV)
        us = timed_graph(lambda: _lib.check(L.ds_attention_bf16(qkv.data_ptr(), o.data_ptr(), B, N, Cc, _lib.stream_ptr())), n=10, reps=5)
        print(f"attention_tc B={B} N={N} C={Cc}: {us:.1f} us, {4.0 * B * N * N * Cc / us / 1e6:.1f} TFLOP/s (4 B N^2 C)")
        return
    sp = lambda: _lib.stream_ptr()
    g = torch.Generator().manual_seed(0)
    w = (torch.randn((cout, cin, ks, ks), generator=g) / (cin * ks * ks) ** 0.5).to(DEV)
    bias = torch.randn((cout,), generator=g).to(DEV)
    gamma, beta = torch.ones(cin, device=DEV), torch.zeros(cin, device=DEV)
    xa32 = torch.randn((B, H, W, ca), generator=g).to(DEV)
    xb32 = torch.randn((B, H, W, cb), generator=g).to(DEV) if cb else None
    out16 = torch.empty((B, H, W, cout), dtype=torch.bfloat16, device=DEV)
    out32 = torch.empty((B, H, W, cout), device=DEV)
    res = torch.randn((B, H, W, cout), generator=g).to(DEV)
    if kind in ("gnconv", "gnconv_tf32"):
        tf32 = kind == "gnconv_tf32"
        nb = (L.ds_gnconv_tf32_scratch_bytes if tf32 else L.ds_gnconv_bf16_scratch_bytes)(B, 16, cin, cout, ks)
        scratch = torch.zeros(nb, dtype=torch.uint8, device=DEV)
        use_res = os.environ.get("MB_NO_RESIDUAL") is None

        def fn():
            if tf32:
                _lib.check(L.ds_gnconv_tf32(xa32.data_ptr(), ca, None if xb32 is None else xb32.data_ptr(), cb, gamma.data_ptr(),
                                            beta.data_ptr(), 16, 1, w.data_ptr(), bias.data_ptr(), res.data_ptr() if use_res else None,
                                            out32.data_ptr(), B, H, W, cout, ks, scratch.data_ptr(), nb, sp()))
            else:
                _lib.check(L.ds_gnconv_bf16(xa32.data_ptr(), ca, None if xb32 is None else xb32.data_ptr(), cb, gamma.data_ptr(),
                                            beta.data_ptr(), 16, 1, w.data_ptr(), bias.data_ptr(), res.data_ptr() if use_res else None,
                                            out16.data_ptr(), out32.data_ptr(), B, H, W, cout, ks, scratch.data_ptr(), nb, sp()))
        us = timed_graph(fn)
        print(f"{kind} (pack + gn_stats + fused conv) {ca}+{cb}->{cout} k{ks} {B}x{H}x{W} residual={use_res}: {us:.2f} us per call (3 kernels)")
        if os.environ.get("DIFFSPLIT_B200_STREAM_DBG"):
            fn()
            ph = (C.c_longlong * 16)()
            _lib.check(L.ds_debug_stream_phases(ph))
            nt_ = max(1, ph[14])
            names = {0: "producer wait empty_raw", 1: "mma wait full_op", 2: "mma wait empty_acc", 3: "mma issue", 4: "epi0 wait full_acc",
                     5: "epi0 work", 6: "epi1 wait full_acc", 7: "epi1 work", 8: "xf wait full_raw", 9: "xf wait empty_op", 10: "xf work"}
            print(f"  stream phases, CTA 0, {ph[14]} tiles, lifetime {ph[15]} cycles = {ph[15] / nt_:.0f} per tile; per tile:",
                  {v: round(ph[k] / nt_) for k, v in names.items()})
        if os.environ.get("DIFFSPLIT_B200_HALO_DBG"):
            fn()
            ph = (C.c_double * 7)()
            n = C.c_int()
            _lib.check(L.ds_debug_halo_phases(ph, C.byref(n)))
            names = ["setup", "wait_prev", "loads+table", "transform+stage", "mma", "epilogue", "teardown"]
            print("  phases (mean cycles over", n.value, "CTAs):", {k: round(v) for k, v in zip(names, ph)}, "sum", round(sum(ph)))
    elif kind == "conv_tc":
        xa = xa32.bfloat16()
        xb = xb32.bfloat16() if cb else None
        nb = L.ds_conv2d_bf16_scratch_bytes(cin, cout, ks)
        scratch = torch.zeros(nb, dtype=torch.uint8, device=DEV)

        use_res = os.environ.get("MB_NO_RESIDUAL") is None
        use_f32 = os.environ.get("MB_NO_F32") is None
        up = int(os.environ.get("MB_UP", "0"))
        if up:
            out16u = torch.empty((B, 2 * H, 2 * W, cout), dtype=torch.bfloat16, device=DEV)
            out32u = torch.empty((B, 2 * H, 2 * W, cout), device=DEV)

        def fn():
            _lib.check(L.ds_conv2d_bf16(xa.data_ptr(), ca, None if xb is None else xb.data_ptr(), cb, w.data_ptr(), bias.data_ptr(),
                                        res.data_ptr() if use_res and not up else None, (out16u if up else out16).data_ptr(),
                                        (out32u if up else out32).data_ptr() if use_f32 else None, 0, B, H, W, cout, ks, 1, up,
                                        scratch.data_ptr(), nb, sp()))
        us = timed_graph(fn)
        fl = 2.0 * B * H * W * cout * cin * (16 if up else ks * ks)
        print(f"conv_tc (pack + conv_tc) {ca}+{cb}->{cout} k{ks} {B}x{H}x{W} up={up} residual={use_res} f32out={use_f32}: {us:.2f} us per call "
              f"(2 kernels), {fl / us / 1e6:.0f} TFLOP/s")
    elif kind == "sampler":
        # fused posterior update + in-register Philox noise: 12 B per element (read x_t, read eps, write x_{t-1})
        from diffsplitting_b200.model import samplers as SM
        x = torch.randn((B, cin, H, W), generator=g).to(DEV)
        eps = torch.randn((B, cin, H, W), generator=g).to(DEV)
        out = torch.empty_like(x)
        coef = torch.tensor([[1.01, 0.1, 0.4, 0.6, 0.05]], dtype=torch.float32, device=DEV)
        threads, inc = SM._rng_geometry(x.numel(), 0)
        a = _lib.StepArgs()
        a.d_x = x.data_ptr(); a.d_net = eps.data_ptr(); a.d_out = out.data_ptr(); a.numel = x.numel()
        a.mode = 0; a.clip = 1; a.d_coef = coef.data_ptr(); a.n_steps = 1; a.step = 0; a.d_state = None
        a.seed = 1234; a.offset = 0; a.offset_inc = inc; a.rng_threads = threads; a.skip_rng_if_zero = 0

        def fn():
            _lib.check(L.ds_sampler_step(C.byref(a), sp()))
        us = timed_graph(fn, n=10, reps=10)
        gb = x.numel() * 12 / 1e9
        print(f"sampler_step (posterior + Philox noise) {B}x{cin}x{H}x{W}: {us:.2f} us per call, {gb / (us * 1e-6):.0f} GB/s algorithmic")
    else:
        nb = L.ds_groupnorm_scratch_bytes(B, 16)
        scratch = torch.zeros(nb, dtype=torch.uint8, device=DEV)
        out = torch.empty((B, H, W, cin), device=DEV)

        def fn():
            _lib.check(L.ds_groupnorm_swish_f32(xa32.data_ptr(), ca, None if xb32 is None else xb32.data_ptr(), cb, gamma.data_ptr(),
                                                beta.data_ptr(), out.data_ptr(), B, H, W, 16, 1, scratch.data_ptr(), nb, sp()))
        us = timed_graph(fn, n=10, reps=10)
        gb = B * H * W * cin * 4 * 3 / 1e9
        print(f"groupnorm (stats + apply, fp32 out) C={cin} {B}x{H}x{W}: {us:.2f} us per call (2 kernels), "
              f"{gb / (us * 1e-6):.0f} GB/s algorithmic (2 reads + 1 write)")


if __name__ == "__main__":
    main()
