"""Timeline of the tensor-core conv launches of ONE graph-replayed reverse step (DIFFSPLIT_B200_TRACE=1 is set here).
Prints, per launch, start offset, duration and the gap to the previous launch's end, all from the GPU global timer."""
import ctypes as C
import os
import sys

os.environ["DIFFSPLIT_B200_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from diffsplitting_b200 import _lib  # noqa: E402


def main():
    w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_WORKLOAD]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    s, net = bench.build_sampler(w, "bf16", dev)
    L = _lib.lib()
    eng = bench.prepare_engine(s, net, w, dev, 64)
    eng.step()                                 # eager step (ids 0..n-1)
    torch.cuda.synchronize()
    _lib.check(L.ds_debug_trace_reset(1))      # forget the eager ids: the capture below records ids 0..n-1 again
    eng.graph = None
    eng.step()                                 # eager + capture: eager launches take ids 0..n-1, captured ones n..2n-1
    for _ in range(5):
        eng.step()
    torch.cuda.synchronize()
    _lib.check(L.ds_debug_trace_reset(0))
    eng.step()                                 # ONE traced replay
    torch.cuda.synchronize()
    buf = (C.c_uint64 * (2 * 8192))()
    kinds = (C.c_int * 8192)()
    n = L.ds_debug_trace_read(buf, kinds, 8192)
    rows = [(buf[2 * i], buf[2 * i + 1], kinds[i], i) for i in range(n) if buf[2 * i + 1] != 0]
    rows.sort()
    t0 = rows[0][0]
    prev_end = None
    tot = 0.0
    print(f"{len(rows)} traced launches in one replay")
    for st, en, k, i in rows:
        gap = (st - prev_end) / 1e3 if prev_end is not None else 0.0
        print(f"id {i:4d} kind {k} start {((st - t0) / 1e3):8.2f} us  dur {((en - st) / 1e3):6.2f} us  gap_after_prev {gap:6.2f} us")
        prev_end = max(prev_end or 0, en)
        tot += (en - st) / 1e3
    print(f"span {(max(r[1] for r in rows) - t0) / 1e3:.1f} us, sum of conv durations {tot:.1f} us")


if __name__ == "__main__":
    main()
