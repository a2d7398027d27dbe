import torch, time
x = torch.zeros(1024, device="cuda")
def run(n):
    for _ in range(n): x.add_(1.0)
run(10); torch.cuda.synchronize()
for n in (100,):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        run(n)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): g.replay()
    e1.record(); torch.cuda.synchronize()
    print("graph: us per tiny kernel node", e0.elapsed_time(e1) * 1000 / (20 * n))
    e0.record(); run(2000); e1.record(); torch.cuda.synchronize()
    print("eager: us per tiny kernel", e0.elapsed_time(e1) * 1000 / 2000)
# bigger kernel: 16 MB elementwise
y = torch.zeros(4 << 20, device="cuda")
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(100): y.add_(1.0)
g.replay(); torch.cuda.synchronize()
e0.record()
for _ in range(20): g.replay()
e1.record(); torch.cuda.synchronize()
print("graph: us per 16MB add_ node", e0.elapsed_time(e1) * 1000 / 2000, "-> GB/s", 2 * 16.777 / (e0.elapsed_time(e1) / 2000))
import subprocess
print(subprocess.run(["nvidia-smi", "--query-gpu=name,compute_mode,mig.mode.current,persistence_mode", "--format=csv"], capture_output=True, text=True).stdout)
