"""Per-op phase cycles of the LAST per-sample chain launch of one eager reverse step (DIFFSPLIT_B200_CHAIN_DBG=1 is set here)."""
import ctypes as C
import os
import sys

os.environ["DIFFSPLIT_B200_CHAIN_DBG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from diffsplitting_b200 import _lib  # noqa: E402


def main():
    w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_WORKLOAD]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    s, net = bench.build_sampler(w, "bf16", dev)
    eng = bench.prepare_engine(s, net, w, dev, 64)
    eng.graph = None
    for _ in range(3):
        eng.step()
    torch.cuda.synchronize()
    buf = (C.c_longlong * (16 * 6))()
    n = C.c_int()
    _lib.check(_lib.lib().ds_debug_chain_phases(buf, 16, C.byref(n)))
    names = ["table", "stage", "mma", "epilogue"]
    for i in range(n.value):
        st = [buf[i * 6 + k] for k in range(6)]
        d = [st[k + 1] - st[k] for k in range(4)]
        nxt = buf[(i + 1) * 6] - st[4] if i + 1 < n.value else 0
        print(f"op {i:2d}: " + "  ".join(f"{nm} {v:6d}" for nm, v in zip(names, d)) + f"  to-next-op {nxt:6d}  total {st[4] - st[0]:6d} cycles  (MMA issuer waited {st[5]:6d} for weights)")


if __name__ == "__main__":
    main()
