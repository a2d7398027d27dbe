"""Probe: does splitting the batch of the benchmark workload over several CUDA streams (each a complete engine with its own
UNet handle and graphs) raise the aggregate step rate?  Usage: python tools/two_stream_probe.py [n_streams ...]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def run(nstreams, steps=400):
    w = dict(bench.WORKLOADS[bench.DEFAULT_WORKLOAD])
    B = w["B"]
    w["B"] = B // nstreams
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    engs, streams = [], []
    for _ in range(nstreams):
        s, net = bench.build_sampler(w, "bf16", dev)
        eng = bench.prepare_engine(s, net, w, dev, steps + 64)
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            eng.run(16)                      # eager step + capture of the 1-step and 8-step graphs on this stream
        engs.append((s, net, eng))
        streams.append(st)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps // 8):
        for (s, net, eng), st in zip(engs, streams):
            with torch.cuda.stream(st):
                eng.run(8)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    n = steps // 8 * 8
    print(f"{nstreams} stream(s) x batch {w['B']}: {n / dt:.1f} steps/s of the full batch ({dt / n * 1e3:.3f} ms per step)")


if __name__ == "__main__":
    for k in [int(a) for a in sys.argv[1:]] or [1, 2, 4]:
        run(k)
