"""Multi-GPU tiled prediction example / check:

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/tiled_predict_multigpu.py [--frames F --size S]

Every rank holds the frames in HBM, predicts its contiguous block of tile chunks with JointIndi (two UNets, T steps),
one all-gather of the predicted tiles over NCCL, stitch on every rank.  Prints tiles/s and checks that all ranks hold
the same stitched result and that the result does not depend on the number of ranks (Philox offsets come from the
GLOBAL chunk index).
"""
import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffsplitting_b200.data import TiledFrames  # noqa: E402
from diffsplitting_b200.model.samplers import JointIndi  # noqa: E402
from diffsplitting_b200.model.unet import UNet  # noqa: E402
from diffsplitting_b200.parallel import tiled_predict_and_stitch  # noqa: E402


def seeded_state_dict(net, seed):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, v in net.state_dict().items():
        if k.endswith("inv_freq"):
            sd[k] = v
        elif v.dim() == 1 and (".block.0." in k or ".norm." in k):
            sd[k] = torch.ones_like(v) if k.endswith("weight") else torch.zeros_like(v)
        else:
            fan = v[0].numel() if v.dim() > 1 else v.numel()
            sd[k] = (torch.rand(v.shape, generator=g) * 2 - 1) / max(1.0, fan) ** 0.5
    return sd


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=2)
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--patch", type=int, default=512)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--chunk", type=int, default=4)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nets = []
    for s in (1, 2):
        n = UNet(in_channel=1, out_channel=1, inner_channel=16, norm_groups=16, channel_mults=(1, 2, 4, 8), attn_res=(),
                 res_blocks=1, image_size=32, variant="ddpm")
        n.load_state_dict(seeded_state_dict(n, s))
        nets.append(n.to(dev).eval())
    joint = JointIndi(None, 32, channels=1, out_channel=1, conditional=False, denoise_fn_ch1=nets[0], denoise_fn_ch2=nets[1],
                      val_schedule_opt={"n_timestep": args.steps}).to(dev)
    joint.set_new_noise_schedule({"n_timestep": args.steps}, dev)
    rng = np.random.default_rng(0)
    frames = rng.integers(0, 1994, size=(2, args.frames, args.size, args.size), dtype=np.uint16)
    nd = {"mean_input": 1000.0, "std_input": 1000.0, "mean_target": np.array([500.0, 500.0]),
          "std_target": np.array([500.0, 500.0])}
    tf = TiledFrames(frames, args.patch, args.patch // 2, normalization_dict=nd, device=dev)

    def infer(inp):
        return joint.inference(inp, continuous=True)[-inp.shape[0]:]

    def run():
        return tiled_predict_and_stitch(infer, tf, chunk=args.chunk, out_channels=2, offset_stride=1 << 20, seed_base=7)

    out = run()                      # warm-up: graph capture, NCCL init
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(device_ids=[local])
    t0 = time.perf_counter()
    out = run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(device_ids=[local])
    dt = time.perf_counter() - t0
    chk = float(out.double().sum())
    sums = [chk]
    if world > 1:
        sums = [None] * world
        dist.all_gather_object(sums, chk)
    if rank == 0:
        print({"tiles": len(tf), "stitched_shape": tuple(out.shape), "n_gpus": world, "tiles_per_s": len(tf) / dt,
               "unet_steps_per_tile": 2 * args.steps, "ranks_agree": len(set(sums)) == 1, "checksum": chk})
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
