#!/usr/bin/env python
"""Benchmark of the sampling hot path (BASELINE.json metric: UNet denoising steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--precision P]

A "step" is one reverse-diffusion step of the named workload over its whole batch: one UNet forward + the fused
sampler update.  Default workload = BASELINE.json configs[1]: ``splitting_hagen_indi_single_ch`` instantiated as
InDI(ddpm UNet 1->1, 16*[1,2,4,8]) on 16 x 1 x 64 x 64 tiles, T = 1000.

Timed region (`value`): W warm-up steps, then exactly K steps with state resident in HBM; every step is bracketed
by its own CUDA-event pair on the launching stream and an L2 flush (write of a 256 MiB buffer) runs BETWEEN the
timed steps, outside the event pairs; `ms_per_step` is the mean event duration, max over ranks.  `e2e` is the same
metric through the public API (`netG.inference`): pinned-host input -> H2D -> full T-step loop -> D2H of the result,
all inside the timed region.  `roofline` comes from a per-operator CUDA-event pass over the same forward
(`ds_unet_forward_profiled`), `cpu_baseline` from the oracle (a port of the reference, torch-CPU) on the host cores.
Multi-GPU: one process per GPU (torchrun), every rank runs its own batch (weak scaling), no data-path collective.
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
    # name: (sampler, unet cfg kwargs, B, H, W, T, cond_channels)
    "hagen_indi_64_b16_T1000": dict(sampler="indi", variant="ddpm", in_ch=1, out_ch=1, inner=16, groups=16,
                                     mults=(1, 2, 4, 8), attn_res=(), res_blocks=1, image_size=32, B=16, H=64, W=64,
                                     T=1000, cond=0, config="splitting_hagen_indi_single_ch.json"),
    "cifar10_ddpm_32_b1_T50": dict(sampler="ddpm", variant="ddpm", in_ch=9, out_ch=6, inner=16, groups=16,
                                   mults=(1, 2, 4, 8), attn_res=(), res_blocks=1, image_size=32, B=1, H=32, W=32, T=50,
                                   cond=3, config="splitting_cifar10.json"),
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
DEFAULT_WORKLOAD = "hagen_indi_64_b16_T1000"
METRIC = "unet_denoising_steps_per_sec"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


# ------------------------------------------------------------------------------------------------ CPU arm
def oracle_setup(w, seed=0):
    from oracle import samplers_ref as S
    from oracle import unet_ref as U
    cfg = U.make_cfg(w["variant"], w["in_ch"], w["out_ch"], w["inner"], w["groups"], w["mults"], w["attn_res"],
                     w["res_blocks"], w["image_size"])
    sd = U.random_state_dict(cfg, seed=seed)
    return S, U, cfg, sd


def oracle_step_fn(w, Bs):
    """Returns a closure running ONE reverse step of the workload on `Bs` batch elements with the oracle."""
    S, U, cfg, sd = oracle_setup(w)
    g = torch.Generator().manual_seed(1)
    Cs = w["in_ch"] - w["cond"]
    x = torch.randn((Bs, Cs, w["H"], w["W"]), generator=g)
    cond = torch.rand((Bs, w["cond"], w["H"], w["W"]), generator=g) * 2 - 1 if w["cond"] else None
    T = w["T"]
    den = lambda xx, tt: U.unet_forward(sd, cfg, xx, tt)
    if w["sampler"] == "indi":
        delta = 1.0 / T
        state = dict(x=x, t=1.0)

        def step():
            if state["t"] < delta * 0.5:
                state["t"] = 1.0
            state["x"] = S.indi_one_step(den, state["x"], delta, state["t"], 0.01, torch.randn(x.shape, generator=g), strict=False)
            state["t"] -= delta
    else:
        tab = S.schedule_tables(dict(schedule="linear", n_timestep=T, linear_start=1e-6, linear_end=1e-2))
        state = dict(x=x, t=T - 1)

        def step():
            t = state["t"]
            nz = torch.randn(x.shape, generator=g)
            if w["sampler"] == "sr3":
                state["x"] = S.sr3_p_sample(tab, den, state["x"], t, cond, True, nz)
            else:
                state["x"] = S.ddpm_p_sample(tab, den, state["x"], torch.full((Bs,), t, dtype=torch.long), cond, True, nz)
            state["t"] = t - 1 if t > 0 else T - 1
    return step


def run_reference_arm(args, w, rank):
    """`--impl reference`: the reference algorithm (oracle port, torch-CPU fp32) on the host cores."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B = w["B"]
    probe = oracle_step_fn(w, B)
    t0 = time.perf_counter()
    probe()
    t_full = time.perf_counter() - t0
    budget = 90.0
    Bs = max(1, min(B, int(B * budget / max(1e-9, (args.steps + args.warmup) * t_full))))
    step = probe if Bs == B else oracle_step_fn(w, Bs)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = args.steps / dt * (Bs / B)
    sample = f"{args.steps} reverse steps on {Bs} of {B} batch elements ({w['H']}x{w['W']}), scaled by {Bs}/{B}"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3 * (B / Bs),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "reference_config": w["config"], "batch": B, "height": w["H"],
                       "width": w["W"], "T": w["T"], "sampler": w["sampler"]},
            "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cpu_baseline(w, seconds=12.0):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = oracle_step_fn(w, w["B"])
    step()
    n, t0 = 0, time.perf_counter()
    while True:
        step()
        n += 1
        dt = time.perf_counter() - t0
        if dt >= seconds or n >= 2000:
            break
    return {"value": n / dt, "unit": "steps/s", "cores": cores, "kind": "port",
            "sample": f"{n} full-batch reverse steps of the oracle (torch-CPU fp32 port of the reference) in {dt:.1f} s"}


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
def build_sampler(w, precision, device):
    from diffsplitting_b200.model.samplers import GaussianDiffusionDdpm, GaussianDiffusionSr3, InDI
    from diffsplitting_b200.model.unet import UNet
    from oracle import unet_ref as U          # only for the seeded random-init weights shared with the CPU arm
    cfg = U.make_cfg(w["variant"], w["in_ch"], w["out_ch"], w["inner"], w["groups"], w["mults"], w["attn_res"],
                     w["res_blocks"], w["image_size"])
    net = UNet(in_channel=w["in_ch"], out_channel=w["out_ch"], inner_channel=w["inner"], norm_groups=w["groups"],
               channel_mults=w["mults"], attn_res=w["attn_res"], res_blocks=w["res_blocks"], image_size=w["image_size"],
               variant=w["variant"], precision=precision)
    net.load_state_dict(U.random_state_dict(cfg, seed=0))
    net = net.to(device).eval()
    T = w["T"]
    Cs = w["in_ch"] - w["cond"]
    if w["sampler"] == "indi":
        s = InDI(net, w["image_size"], channels=Cs, out_channel=1, conditional=False, val_schedule_opt={"n_timestep": T})
        s.set_new_noise_schedule({"n_timestep": T}, device)
    else:
        cls = GaussianDiffusionSr3 if w["sampler"] == "sr3" else GaussianDiffusionDdpm
        s = cls(net, w["image_size"], channels=Cs, conditional=True).to(device)
        s.set_new_noise_schedule(dict(schedule="linear", n_timestep=T, linear_start=1e-6, linear_end=1e-2), device)
    return s, net


def prepare_engine(s, net, w, device, n_steps):
    """Device-resident engine for a chain long enough to cover warm-up + timed steps."""
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=list(WORKLOADS))
    ap.add_argument("--precision", default=os.environ.get("DIFFSPLIT_B200_PRECISION", "bf16"), choices=["fp32", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-calls", type=int, default=3)
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly (for ncu launch lists)")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
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

    torch.manual_seed(2 + rank)
    s, net = build_sampler(w, args.precision, device)
    B, H, W, T = w["B"], w["H"], w["W"], w["T"]
    K, Wm = args.steps, args.warmup
    eng = prepare_engine(s, net, w, device, K + Wm)
    launches_per_step = eng.launches_per_step()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)

    for _ in range(Wm):
        eng.step()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    with ClockSampler(local_rank) as clocks:
        barrier()
        t_wall = time.perf_counter()
        for a, b in ev:
            flush.fill_(1)                  # L2 flush, outside the event pair
            a.record()
            eng.step()
            b.record()
        barrier()
        t_wall = time.perf_counter() - t_wall
        # the same chain without flushes, timed as one region (steady state: activations stay in L2)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng2 = prepare_engine(s, net, w, device, K + Wm)
        for _ in range(Wm):
            eng2.step()
        barrier()
        c0.record()
        for _ in range(K):
            eng2.step()
        c1.record()
        barrier()
        chain_ms = c0.elapsed_time(c1) / K

        # ---- e2e: public API, host buffers, H2D + D2H inside the timed region
        x_host = (torch.rand((B, 1 if w["sampler"] == "indi" else w["cond"], H, W)) * 2 - 1).pin_memory()
        out_host = None

        def api_call():
            nonlocal out_host
            x_dev = x_host.to(device, non_blocking=True)
            y = s.inference(x_dev, continuous=True)[-B:]
            if out_host is None:
                out_host = torch.empty(y.shape, dtype=y.dtype).pin_memory()
            out_host.copy_(y, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        if args.e2e_calls > 0:
            api_call()                                 # warm-up (graph capture for the API's engine)
        barrier()
        e0 = time.perf_counter()
        for _ in range(args.e2e_calls):
            flush.fill_(1)
            api_call()
        barrier()
        e2e_s = time.perf_counter() - e0
    clk = clocks.summary()

    ms = [a.elapsed_time(b) for a, b in ev]
    local = torch.tensor([sum(ms) / K, chain_ms, e2e_s], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(local, op=dist.ReduceOp.MAX)
    ms_per_step, chain_ms, e2e_s = [float(v) for v in local.cpu()]
    value = world * 1e3 / ms_per_step
    e2e_value = world * args.e2e_calls * T / e2e_s if args.e2e_calls > 0 else None

    # ---- roofline: per-operator CUDA-event pass over the same forward (rank 0)
    roof, breakdown = None, None
    if rank == 0:
        pk = peaks()
        agg = {}
        reps = 3
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
        tensor_bound = top in ("conv_tc", "gn_swish_conv_tc", "conv_chain_tc", "conv_f32", "attention") and ai * pk["hbm"] * 1e9 > pk["tf_sustained"] * 1e12
        if tensor_bound:
            ach = a["flops"] / a["ms"] / 1e9
            roof = {"bound": "tensor", "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"]}
        else:
            ach = a["bytes"] / a["ms"] / 1e6
            roof = {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"]}
        traffic, traffic_src = None, None
        tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_halo_traffic.json")
        if top == "gn_swish_conv_tc" and args.workload == DEFAULT_WORKLOAD and args.precision == "bf16" and os.path.exists(tpath):
            with open(tpath) as fh:
                tj = json.load(fh)
            traffic = tj["traffic_bytes_per_launch"]
            traffic_src = ("profiles/r1_halo_traffic.json: dram read+write bytes per launch, mean over the %d fused-conv launches of one step, "
                           "ncu --set full (caches flushed per replay); algorithmic bytes per launch = %.0f"
                           % (tj["launches_captured"], a["bytes"] / a["launches"]))
        roof.update({"traffic": traffic, "traffic_source": traffic_src, "kernel": top, "share_of_step": a["ms"] / tot, "peak_source": pk["src"],
                     "per_launch_us": a["ms"] / a["launches"] * 1e3,
                     "note": "algorithmic bytes|flops of all launches of this kernel in one step / their summed CUDA-event time "
                             "(each operator timed as 20 back-to-back launches after 1 warm-up, L2 warm as in the real chain)"})
        if os.environ.get("DIFFSPLIT_B200_DUMP_OPS"):
            with open(os.environ["DIFFSPLIT_B200_DUMP_OPS"], "w") as fh:
                json.dump(per_op, fh, indent=0)

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        cpu = cpu_baseline(w)

    if rank == 0:
        numel_state = B * (w["in_ch"] - w["cond"]) * H * W
        line = {"metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": K, "warmup": Wm,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
                "config": {"workload": args.workload, "reference_config": w["config"], "batch": B, "height": H, "width": W,
                           "T": T, "sampler": w["sampler"], "precision": args.precision,
                           "l2": "flushed between timed steps (256 MiB write outside the event pairs)",
                           "parallelism": f"replicas x{world} (independent batches, no collective)"},
                "chain_steps_per_sec_l2_warm": world * 1e3 / chain_ms,
                "gflop_per_step": net.flops(H, W) * B / 1e9,
                "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": x_host.numel() * 4 / T,
                        "d2h_bytes_per_step": (out_host.numel() * 4 / T) if out_host is not None else 0, "calls": args.e2e_calls,
                        "note": f"netG.inference: one call = H2D + {T} reverse steps + D2H"},
                "gpu_launches": launches_per_step * K, "launches_per_step": launches_per_step,
                "clocks": clk, "roofline": roof, "kernel_breakdown": breakdown, "cpu_baseline": cpu,
                "wall_s_timed_region": t_wall}
        if "joint" in args.workload:
            # BASELINE.json's second metric: a tile of the joint split needs T steps of each of its two UNets; this workload
            # times one of them on a batch of B tiles
            line["tiles_per_sec_equiv"] = value * B / (2.0 * T)
            line["tiles_note"] = f"steps/s x {B} tiles per batch / (2 UNets x T={T} steps per tile)"
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
