#!/usr/bin/env python
"""Benchmark of the sampling hot path (BASELINE.json metric: UNet denoising steps/s & split tiles/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--precision P]

A "step" is one reverse-diffusion step of the named workload over its whole batch: one UNet forward + the fused sampler
update.  Default workload = the largest single-GPU configuration of BASELINE.json, ``sr_sr3_64_512.json``: the SR3 sampler
(`GaussianDiffusion.p_sample`) driving the sr3 UNet (inner 64, mults 1,2,4,8,16) on cond 8x3x512x512 / state 8x3x512x512,
T = 2000 - 9 969 GFLOP per step.

ONE JSON line on stdout:
  * `value` / `ms_per_step`: W warm-up steps, then exactly K steps with the state resident in HBM; every step is bracketed by
    its own CUDA-event pair on the launching stream, an L2 flush (write of a 256 MiB buffer) runs BETWEEN the timed steps
    outside the event pairs; mean event duration, max over ranks.
  * `steady_state`: the same chain without flushes for >= 1 s (clocks under sustained load are sampled over both regions).
  * `e2e`: the public API (`netG.super_resolution` / `netG.inference`) from a pinned HOST input to a pinned HOST result,
    H2D + the reverse loop + D2H inside the timed region.  For T = 2000 workloads one call runs a T_e2e-step schedule (same
    per-step work, copies amortised over FEWER steps than the real call).
  * `roofline`: dominant kernel of the step from a per-operator CUDA-event pass (`ds_unet_forward_profiled`);
    `step_roofline`: the whole step's algorithmic FLOP/s against the sustained bf16 peak.
  * `by_workload`: the other four BASELINE configs, measured the same way (short runs).
  * `tiles`: BASELINE's second metric, MEASURED: 10 x 2048^2 two-channel uint16 frames -> 490 tiles of 512^2 -> JointIndi
    (two UNets) for T in {1, 5} -> packed-box gather (NCCL when N > 1, strong scaling over the same 490 tiles) -> stitch.
  * `gpu_eager_baseline`: the UNMODIFIED reference modules (baseline/_ref) run eagerly by PyTorch on the same GPU.
  * `cpu_baseline` / `--impl reference`: the unmodified reference modules on the host cores (bounded sample).
Multi-GPU: one process per GPU (torchrun); steps/s = independent replicas (weak scaling, no data-path collective).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "cifar10_ddpm_32_b1_T50": dict(sampler="ddpm", variant="ddpm", in_ch=9, out_ch=6, inner=16, groups=16,
                                   mults=(1, 2, 4, 8), attn_res=(), res_blocks=1, image_size=32, B=1, H=32, W=32, T=50,
                                   cond=3, config="splitting_cifar10.json"),
    "hagen_indi_64_b16_T1000": dict(sampler="indi", variant="ddpm", in_ch=1, out_ch=1, inner=16, groups=16,
                                     mults=(1, 2, 4, 8), attn_res=(), res_blocks=1, image_size=32, B=16, H=64, W=64,
                                     T=1000, cond=0, config="splitting_hagen_indi_single_ch.json"),
    "hagen_joint_512_b8_T5": dict(sampler="indi", variant="ddpm", in_ch=1, out_ch=1, inner=16, groups=16,
                                  mults=(1, 2, 4, 8), attn_res=(), res_blocks=1, image_size=32, B=8, H=512, W=512, T=5,
                                  cond=0, config="splitting_hagen_indi_joint.json (one of its two UNets, 8 tiles)"),
    "sr3_16_128_b32_T2000": dict(sampler="sr3", variant="sr3", in_ch=6, out_ch=3, inner=64, groups=32,
                                 mults=(1, 2, 4, 8, 8), attn_res=(16,), res_blocks=2, image_size=128, B=32, H=128, W=128,
                                 T=2000, cond=3, config="sr_sr3_16_128.json"),
    "sr3_64_512_b8_T2000": dict(sampler="sr3", variant="sr3", in_ch=6, out_ch=3, inner=64, groups=16,
                                mults=(1, 2, 4, 8, 16), attn_res=(), res_blocks=1, image_size=512, B=8, H=512, W=512,
                                T=2000, cond=3, config="sr_sr3_64_512.json"),
}
DEFAULT_WORKLOAD = "sr3_64_512_b8_T2000"
METRIC = "unet_denoising_steps_per_sec"
SCHED = dict(schedule="linear", linear_start=1e-6, linear_end=1e-2)
L2_NOTE = "flushed between timed steps (256 MiB write outside the event pairs)"


def config_dict(name, w):
    """Identical for both arms (the driver compares the dicts key by key)."""
    return {"workload": name, "reference_config": w["config"], "batch": w["B"], "height": w["H"], "width": w["W"],
            "T": w["T"], "sampler": w["sampler"], "l2": L2_NOTE,
            "parallelism": "replicas (independent batches per GPU, no data-path collective)"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


def make_cfg(w):
    from oracle import unet_ref as U          # only for the seeded random-init weights shared by every arm
    return U, U.make_cfg(w["variant"], w["in_ch"], w["out_ch"], w["inner"], w["groups"], w["mults"], w["attn_res"],
                         w["res_blocks"], w["image_size"])


# ------------------------------------------------------------------------------------------------ reference arms
def reference_step_fn(w, Bs, device="cpu", crop=1.0):
    """Closure running ONE reverse step of the workload on `Bs` batch elements with the UNMODIFIED reference modules
    (baseline/_ref: `sr3_modules` / `ddpm_modules` UNet + sampler classes, built directly because `define_G` raises for
    sr3 / ddpm, model/networks.py:159-170); falls back to the oracle port when baseline/_ref is not installed.
    Returns (step, kind)."""
    from baseline import refshim
    U, cfg = make_cfg(w)
    sd = U.random_state_dict(cfg, seed=0)
    H, W = int(w["H"] * crop), int(w["W"] * crop)
    g = torch.Generator().manual_seed(1)
    Cs = w["in_ch"] - w["cond"]
    x = torch.randn((Bs, Cs, H, W), generator=g).to(device)
    cond = (torch.rand((Bs, w["cond"], H, W), generator=g) * 2 - 1).to(device) if w["cond"] else None
    T = w["T"]
    if not refshim.available():
        return _oracle_step_fn(w, U, cfg, sd, x, cond, g), "port"
    ukw = dict(in_channel=w["in_ch"], out_channel=w["out_ch"], inner_channel=w["inner"], norm_groups=w["groups"],
               channel_mults=w["mults"], attn_res=w["attn_res"], res_blocks=w["res_blocks"], dropout=0, image_size=w["image_size"])
    if w["sampler"] == "indi":
        netG, net = refshim.build_sampler("indi", ukw, dict(image_size=w["image_size"], channels=Cs, out_channel=1, conditional=False,
                                                            val_schedule_opt={"n_timestep": T}))
    else:
        netG, net = refshim.build_sampler(w["sampler"], ukw, dict(image_size=w["image_size"], channels=Cs, conditional=True))
    net.load_state_dict(sd, strict=True)
    netG = netG.to(device).eval()
    netG.set_new_noise_schedule(dict(SCHED, n_timestep=T), device)
    state = dict(x=x, t=1.0 if w["sampler"] == "indi" else T - 1)

    @torch.no_grad()
    def step():
        if w["sampler"] == "indi":
            delta = 1.0 / T
            if state["t"] < 2 * delta:
                state["t"] = 1.0
            state["x"] = netG.inference_one_step(state["x"], delta, state["t"])
            state["t"] -= delta
        else:
            t = state["t"]
            if w["sampler"] == "sr3":
                state["x"] = netG.p_sample(state["x"], t, condition_x=cond)
            else:
                state["x"] = netG.p_sample(state["x"], torch.full((Bs,), t, dtype=torch.long, device=device), condition_x=cond)
            state["t"] = t - 1 if t > 0 else T - 1
    return step, "reference"


def _oracle_step_fn(w, U, cfg, sd, x, cond, g):
    from oracle import samplers_ref as S
    T = w["T"]
    den = lambda xx, tt: U.unet_forward(sd, cfg, xx, tt)
    state = dict(x=x, t=1.0 if w["sampler"] == "indi" else T - 1)
    tab = None if w["sampler"] == "indi" else S.schedule_tables(dict(SCHED, n_timestep=T))

    def step():
        nz = torch.randn(x.shape, generator=g)
        if w["sampler"] == "indi":
            delta = 1.0 / T
            if state["t"] < 2 * delta:
                state["t"] = 1.0
            state["x"] = S.indi_one_step(den, state["x"], delta, state["t"], 0.01, nz, strict=False)
            state["t"] -= delta
        else:
            t = state["t"]
            if w["sampler"] == "sr3":
                state["x"] = S.sr3_p_sample(tab, den, state["x"], t, cond, True, nz)
            else:
                state["x"] = S.ddpm_p_sample(tab, den, state["x"], torch.full((x.shape[0],), t, dtype=torch.long), cond, True, nz)
            state["t"] = t - 1 if t > 0 else T - 1
    return step


def pick_cpu_sample(w, n_steps, budget_s):
    """Largest sample of the workload (batch elements, then a centre crop) whose n_steps fit the budget.  Returns
    (step, kind, Bs, crop, scale) with scale = fraction of the full step's work the sample does."""
    B = w["B"]
    probe, kind = reference_step_fn(w, 1)
    probe()                                             # first call pays allocator / thread-pool start-up
    t0 = time.perf_counter()
    probe()
    t1 = time.perf_counter() - t0
    if t1 * n_steps <= budget_s:
        Bs = max(1, min(B, int(budget_s / (t1 * n_steps))))
        if Bs == 1:
            return probe, kind, 1, 1.0, 1.0 / B
        step, kind = reference_step_fn(w, Bs)
        return step, kind, Bs, 1.0, Bs / B
    crop = 0.5
    while crop > 0.125 and t1 * crop * crop * n_steps > budget_s:
        crop *= 0.5
    step, kind = reference_step_fn(w, 1, crop=crop)
    return step, kind, 1, crop, crop * crop / B


def run_reference_arm(args, w, rank):
    """`--impl reference`: the reference's own CPU implementation of the path on all host cores.  Rank 0 only."""
    if rank != 0:
        return
    sys.stdout.flush()
    saved_stdout = os.dup(1)             # the reference's constructors print: stdout carries the JSON line only
    os.dup2(2, 1)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B = w["B"]
    step, kind, Bs, crop, scale = pick_cpu_sample(w, args.steps + args.warmup, 150.0)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = args.steps / dt * scale
    sample = (f"{args.steps} reverse steps on {Bs} of {B} batch elements at {int(w['H'] * crop)}x{int(w['W'] * crop)} of "
              f"{w['H']}x{w['W']}, scaled by {scale:.6g} ({'unmodified reference modules from baseline/_ref' if kind == 'reference' else 'oracle port'}, "
              f"torch-CPU fp32)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3 / scale,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args.workload, w),
            "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)


def cpu_baseline(w, seconds=15.0):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, kind, Bs, crop, scale = pick_cpu_sample(w, 3, seconds)
    vals = []
    t_all = time.perf_counter()
    for _ in range(3):                                   # median of 3 single-step timings: the box's CPU rate wanders
        t0 = time.perf_counter()
        step()
        vals.append(scale / (time.perf_counter() - t0))
    return {"value": statistics.median(vals), "unit": "steps/s", "cores": cores, "kind": kind,
            "sample": f"median of 3 reverse steps on {Bs} of {w['B']} batch elements at {int(w['H'] * crop)}x{int(w['W'] * crop)}, "
                      f"scaled by {scale:.6g}; {time.perf_counter() - t_all:.1f} s of CPU work",
            "all": vals}


def gpu_eager_baseline(w, device, steps=3):
    """The existing Blackwell path: the reference's modules run eagerly by PyTorch on this GPU - cuDNN convs
    (TF32 allowed, torch's default), native_group_norm, ~225-470 launches + an H2D per step."""
    try:
        step, kind = reference_step_fn(w, w["B"], device=device)
        step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out = {"value": 1e3 / ms, "unit": "steps/s", "ms_per_step": ms, "kind": kind, "steps": steps,
               "cudnn_allow_tf32": bool(torch.backends.cudnn.allow_tf32), "matmul_allow_tf32": bool(torch.backends.cuda.matmul.allow_tf32)}
    except Exception as e:      # e.g. out of memory for the reference's fp32 activations: report, do not fail the bench
        out = {"value": None, "error": repr(e)[:200]}
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                r = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5)
                if r.returncode == 0 and r.stdout.strip():
                    self.rows.append([c.strip() for c in r.stdout.strip().split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        reasons = []
        for i, nm in enumerate(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")):
            if any(r[3 + i].lower().startswith("active") for r in self.rows if len(r) > 3 + i):
                reasons.append(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows), "power_w_max": max(float(r[2]) for r in self.rows)}


# ------------------------------------------------------------------------------------------------ our arm
def build_sampler(w, precision, device, T=None):
    from diffsplitting_b200.model.samplers import GaussianDiffusionDdpm, GaussianDiffusionSr3, InDI
    from diffsplitting_b200.model.unet import UNet
    U, cfg = make_cfg(w)
    net = UNet(in_channel=w["in_ch"], out_channel=w["out_ch"], inner_channel=w["inner"], norm_groups=w["groups"],
               channel_mults=w["mults"], attn_res=w["attn_res"], res_blocks=w["res_blocks"], image_size=w["image_size"],
               variant=w["variant"], precision=precision)
    net.load_state_dict(U.random_state_dict(cfg, seed=0))
    net = net.to(device).eval()
    T = T or w["T"]
    Cs = w["in_ch"] - w["cond"]
    if w["sampler"] == "indi":
        s = InDI(net, w["image_size"], channels=Cs, out_channel=1, conditional=False, val_schedule_opt={"n_timestep": T})
        s.set_new_noise_schedule({"n_timestep": T}, device)
    else:
        cls = GaussianDiffusionSr3 if w["sampler"] == "sr3" else GaussianDiffusionDdpm
        s = cls(net, w["image_size"], channels=Cs, conditional=True).to(device)
        s.set_new_noise_schedule(dict(SCHED, n_timestep=T), device)
    return s, net


def prepare_engine(s, net, w, device, n_steps):
    """Device-resident engine for a chain long enough to cover n_steps."""
    from diffsplitting_b200.model import samplers as SM
    B, H, W = w["B"], w["H"], w["W"]
    Cs = w["in_ch"] - w["cond"]
    net.commit()
    if w["sampler"] == "indi":
        eng = s._engine(net, B, Cs, H, W, 0, 1, device, 1, False, 0)
        T = max(w["T"], n_steps)
        coef, ttab = s._tables(T, 1.0)
    else:
        eng = s._engine(net, B, Cs, H, W, w["cond"], B, device, 0, True, 0 if w["sampler"] == "ddpm" else 1)
        coef, ttab = s._coef()
        reps = -(-n_steps // coef.shape[0])
        if reps > 1:
            coef, ttab = coef.repeat(reps, 1), torch.cat([ttab[:-1].repeat(reps), ttab[-1:]])
    eng.load_tables(coef, ttab)
    gen, _ = SM._generator(device)
    eng.reset_state(gen.initial_seed(), gen.get_offset())
    g = torch.Generator().manual_seed(1)
    if eng.cond is not None:
        eng.cond.copy_((torch.rand(tuple(eng.cond.shape), generator=g) * 2 - 1).to(device))
        eng.initial_noise(None, 1.0)
    else:
        base = (torch.rand(tuple(eng.x.shape), generator=g) * 2 - 1).to(device)
        eng.initial_noise(base, 0.01)
    return eng


def time_steps(eng, K, flush):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for a, b in ev:
        flush.fill_(1)                  # L2 flush, outside the event pair
        a.record()
        eng.step()
        b.record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev) / K


def side_workload(name, precision, device, flush, steps=5, warmup=3):
    w = WORKLOADS[name]
    s, net = build_sampler(w, precision, device)
    eng = prepare_engine(s, net, w, device, steps + warmup + 1)
    for _ in range(warmup):
        eng.step()
    torch.cuda.synchronize()
    ms = time_steps(eng, steps, flush)
    gf = net.flops(w["H"], w["W"]) * w["B"] / 1e9
    out = {"ms_per_step": ms, "steps_per_sec": 1e3 / ms, "gflop_per_step": gf, "tflops": gf / ms, "precision": net.precision,
           "launches_per_step": eng.launches_per_step(), "steps": steps, "warmup": warmup, "config": config_dict(name, w)}
    del eng, s, net
    torch.cuda.empty_cache()
    return out


def tiles_leg(device, rank, world, precision, barrier):
    """BASELINE.json's second metric, measured: the joint 2-channel Hagen split with tiled prediction + stitching
    (config splitting_hagen_indi_joint.json; split.py:57-62 hard-codes 10 x 2048 x 2048 frames, grid = patch / 2)."""
    import numpy as np
    import torch.distributed as dist
    from diffsplitting_b200.data import TiledFrames
    from diffsplitting_b200.model.samplers import JointIndi
    from diffsplitting_b200.model.unet import UNet
    from diffsplitting_b200.parallel import PackedLayout, tiled_predict_and_stitch
    w = WORKLOADS["hagen_joint_512_b8_T5"]
    U, cfg = make_cfg(w)
    nets = []
    for seed in (1, 2):
        n = UNet(in_channel=1, out_channel=1, inner_channel=16, norm_groups=16, channel_mults=(1, 2, 4, 8), attn_res=(),
                 res_blocks=1, image_size=32, variant="ddpm", precision=precision)
        n.load_state_dict(U.random_state_dict(cfg, seed=seed))
        nets.append(n.to(device).eval())
    rng = np.random.default_rng(0)
    frames = rng.integers(0, 1994, size=(2, 10, 2048, 2048), dtype=np.uint16)
    nd = {"mean_input": 1000.0, "std_input": 1000.0, "mean_target": np.array([500.0, 500.0]), "std_target": np.array([500.0, 500.0])}
    tf = TiledFrames(frames, 512, 256, normalization_dict=nd, input_from_normalized_target=True, device=device)
    chunk = 8
    lay = PackedLayout(tf.tile_manager, 2, chunk, world)
    res = {"frames": "2 x 10 x 2048 x 2048 uint16 (synthetic)", "tiles": len(tf), "patch": 512, "grid": 256, "chunk": chunk,
           "scaling": "strong (the same 490 tiles over N GPUs, round-robin chunks)", "precision": nets[0].precision,
           "gather_bytes": lay.payload_bytes if world > 1 else 0,
           "collective": "ncclGather of the packed destination boxes to rank 0, one stitch" if world > 1 else "none (1 GPU)"}
    for T in (1, 5):
        joint = JointIndi(None, 32, channels=1, out_channel=1, conditional=False, denoise_fn_ch1=nets[0], denoise_fn_ch2=nets[1],
                          val_schedule_opt={"n_timestep": T}).to(device)
        joint.set_new_noise_schedule({"n_timestep": T}, device)

        def infer(inp):
            return joint.inference(inp, continuous=False, all_samples=True)

        def run():
            return tiled_predict_and_stitch(infer, tf, chunk=chunk, out_channels=2, offset_stride=1 << 20, seed_base=7, root=0)

        run()                            # warm-up: graph capture, NCCL connection
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        out = run()
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        dt = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=device)      # device time, max over ranks
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if rank == 0:
            res[f"T{T}"] = {"tiles_per_sec": len(tf) / float(dt), "seconds": float(dt), "seconds_wall_rank0": wall, "unet_steps_per_tile": 2 * T,
                            "unet_steps_per_sec": len(tf) * 2 * T / float(dt),
                            "stitched_shape": list(out.shape), "checksum": float(out.double().sum()),
                            "checksum_abs": float(out.double().abs().sum())}
        del out, joint
    del tf, nets
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=list(WORKLOADS))
    ap.add_argument("--precision", default=os.environ.get("DIFFSPLIT_B200_PRECISION", "auto"), choices=["auto", "fp32", "bf16", "tf32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip by_workload / tiles / gpu_eager_baseline")
    ap.add_argument("--e2e-calls", type=int, default=2)
    ap.add_argument("--e2e-steps", type=int, default=50, help="steps of the schedule one e2e API call runs (capped by T)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly (for ncu launch lists)")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if os.environ.get("BENCH_BATCH"):                  # probing only (L2-residency experiments): not a bench configuration
        w = dict(w, B=int(os.environ["BENCH_BATCH"]))
    if args.no_graph:
        os.environ["DIFFSPLIT_B200_GRAPH"] = "0"
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, w, rank)
        return

    import torch.distributed as dist
    # stdout carries exactly ONE JSON line: everything else written to fd 1 while running (e.g. NCCL's version banner)
    # is redirected to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the hot path has no CPU fallback")
    args.warmup = max(3, args.warmup)
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    precision = None if args.precision == "auto" else args.precision
    torch.manual_seed(2 + rank)
    s, net = build_sampler(w, precision, device)
    prec_used = net.precision
    B, H, W, T = w["B"], w["H"], w["W"], w["T"]
    K, Wm = args.steps, args.warmup
    eng = prepare_engine(s, net, w, device, K + Wm)
    launches_per_step = eng.launches_per_step()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)

    for _ in range(Wm):
        eng.step()
    barrier()
    with ClockSampler(local_rank) as clocks:
        barrier()
        t_wall = time.perf_counter()
        ms_local = time_steps(eng, K, flush)
        barrier()
        t_wall = time.perf_counter() - t_wall
        # ---- steady state: the same chain without flushes for >= 1 s (activations stay in L2 where they fit)
        n_steady = max(K, int(1200.0 / max(ms_local, 1e-3)) + 1)
        eng2 = prepare_engine(s, net, w, device, n_steady + Wm)
        for _ in range(Wm):
            eng2.step()
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        eng2.run(n_steady)
        c1.record()
        barrier()
        steady_ms = c0.elapsed_time(c1) / n_steady
        del eng2

        # ---- e2e: public API, host buffers, H2D + D2H inside the timed region
        T_e2e = min(T, args.e2e_steps)
        s_api, net_api = (s, net) if T_e2e == T else build_sampler(w, precision, device, T=T_e2e)
        if net_api is not net:
            net_api.load_state_dict(net.state_dict())
        x_host = (torch.rand((B, 1 if w["sampler"] == "indi" else w["cond"], H, W)) * 2 - 1).pin_memory()
        out_host = None

        def api_call():
            nonlocal out_host
            x_dev = x_host.to(device, non_blocking=True)
            if w["sampler"] == "indi":
                y = s_api.inference(x_dev, continuous=False, all_samples=True)
            else:
                y = s_api.super_resolution(x_dev, continous=False, all_samples=True)
            if out_host is None:
                out_host = torch.empty(y.shape, dtype=y.dtype).pin_memory()
            out_host.copy_(y, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        e2e_s = float("nan")
        if args.e2e_calls > 0:
            api_call()                                 # warm-up (graph capture for the API's engine)
            barrier()
            e0 = time.perf_counter()
            for _ in range(args.e2e_calls):
                flush.fill_(1)
                api_call()
            barrier()
            e2e_s = time.perf_counter() - e0
        if s_api is not s:
            del s_api, net_api
    clk = clocks.summary()

    local = torch.tensor([ms_local, steady_ms, e2e_s], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(local, op=dist.ReduceOp.MAX)
    ms_per_step, steady_ms, e2e_s = [float(v) for v in local.cpu()]
    value = world * 1e3 / ms_per_step
    e2e_value = world * args.e2e_calls * T_e2e / e2e_s if args.e2e_calls > 0 else None

    # ---- roofline: per-operator CUDA-event pass over the same forward (rank 0)
    roof, step_roof, breakdown = None, None, None
    gflop = net.flops(H, W) * B / 1e9
    if rank == 0:
        pk = peaks()
        agg = {}
        reps = 2
        xin = eng.x if eng.cond is None else eng.cond
        xb = None if eng.cond is None else eng.x
        net.profile(xin, xb, eng.time)
        per_op = None
        for _ in range(reps):
            per_op = net.profile(xin, xb, eng.time)      # every operator: 1 warm-up + 20 back-to-back launches per event pair
            for o in per_op:
                a = agg.setdefault(o["kind"], dict(ms=0.0, flops=0.0, bytes=0.0, launches=0))
                a["ms"] += o["ms"]; a["flops"] += o["flops"]; a["bytes"] += o["bytes"]; a["launches"] += o["launches"]
        tot = sum(a["ms"] for a in agg.values())
        breakdown = {k: {"share": a["ms"] / tot, "ms_per_step": a["ms"] / reps, "launches_per_step": a["launches"] // reps,
                         "achieved_gbs": a["bytes"] / a["ms"] / 1e6 if a["ms"] else None,
                         "achieved_tflops": a["flops"] / a["ms"] / 1e9 if a["ms"] else None} for k, a in agg.items()}
        top = max(agg, key=lambda k: agg[k]["ms"])
        a = agg[top]
        ai = a["flops"] / max(a["bytes"], 1.0)
        tensor_bound = a["flops"] > 0 and ai * pk["hbm"] * 1e9 > pk["tf_sustained"] * 1e12
        if tensor_bound:
            ach = a["flops"] / a["ms"] / 1e9
            roof = {"bound": "tensor", "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"]}
        else:
            ach = a["bytes"] / a["ms"] / 1e6
            roof = {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"]}
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath)).get(f"{args.workload}:{top}")
            if tj:
                traffic, traffic_src = tj["traffic_bytes_per_launch"], tj["source"]
        roof.update({"traffic": traffic, "traffic_source": traffic_src, "kernel": top, "share_of_step": a["ms"] / tot,
                     "peak_source": pk["src"] + " (sustained: the kernel runs inside a long step)",
                     "per_launch_us": a["ms"] / a["launches"] * 1e3, "arithmetic_intensity_flop_per_byte": ai,
                     "algorithmic_bytes_per_launch": a["bytes"] / a["launches"], "algorithmic_flops_per_launch": a["flops"] / a["launches"],
                     "note": "algorithmic flops (or bytes) of all launches of this kernel in one step / their summed CUDA-event time "
                             "(each operator timed as 20 back-to-back launches after 1 warm-up)"})
        step_tf = gflop / ms_per_step
        step_roof = {"bound": "tensor", "achieved": step_tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": step_tf / pk["tf_sustained"],
                     "note": "whole step: algorithmic GFLOP per step (BASELINE.md section 3) / ms_per_step"}
        if os.environ.get("DIFFSPLIT_B200_DUMP_OPS"):
            with open(os.environ["DIFFSPLIT_B200_DUMP_OPS"], "w") as fh:
                json.dump(per_op, fh, indent=0)

    del eng
    extras = not args.no_extras
    by_workload, eager = None, None
    if rank == 0 and world == 1 and extras:
        by_workload = {}
        for name in WORKLOADS:
            if name != args.workload:
                try:
                    by_workload[name] = side_workload(name, precision, device, flush)
                except Exception as e:
                    by_workload[name] = {"error": repr(e)[:300]}
    tiles = None
    if extras:
        try:
            tiles = tiles_leg(device, rank, world, precision, barrier)
        except Exception as e:
            tiles = {"error": repr(e)[:300]}
    if rank == 0 and world == 1 and extras:
        del s, net
        torch.cuda.empty_cache()
        eager = gpu_eager_baseline(w, device)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(w)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": K, "warmup": Wm,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": {"bf16": "bf16", "tf32": "tf32", "fp32": "f32"}[prec_used], "precision": prec_used,
                "data": "synthetic", "config": config_dict(args.workload, w),
                "steady_state": {"steps_per_sec": world * 1e3 / steady_ms, "ms_per_step": steady_ms, "steps": n_steady,
                                 "seconds": steady_ms * n_steady / 1e3, "l2": "not flushed (consecutive steps)"},
                "gflop_per_step": gflop,
                "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": x_host.numel() * 4 / T_e2e,
                        "d2h_bytes_per_step": (out_host.numel() * 4 / T_e2e) if out_host is not None else 0, "calls": args.e2e_calls,
                        "steps_per_call": T_e2e,
                        "note": f"netG.{'inference' if w['sampler'] == 'indi' else 'super_resolution'}: one call = H2D of the input + "
                                f"{T_e2e} reverse steps + D2H of the result (the T={T} call is the same loop, copies amortised over more steps)"},
                "gpu_launches": launches_per_step * K, "launches_per_step": launches_per_step,
                "clocks": clk, "roofline": roof, "step_roofline": step_roof, "kernel_breakdown": breakdown,
                "by_workload": by_workload, "tiles": tiles, "gpu_eager_baseline": eager, "cpu_baseline": cpu,
                "wall_s_timed_region": t_wall}
        if tiles and "T1" in tiles:
            line["tiles_per_sec"] = tiles["T1"]["tiles_per_sec"]
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
