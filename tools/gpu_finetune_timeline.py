"""In-graph kernel timeline of the captured fine-tune step (BASELINE config 4) from torch.profiler (CUPTI): per-kernel
total time, and how much of the step the GPU is busy on each stream."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import collections
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from bayeslms_b200 import _lib, model as M
from bayeslms_b200.trainer import FineTuner
_lib.init(0)
dev = torch.device("cuda:0")
torch.manual_seed(1111)
dr = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
net = M.VTransformerModel(bench.V, bench.D, bench.NHEAD, bench.FF, bench.NLAYERS, dr, True, "11").to(dev).train()
ft = FineTuner(net, 0.01, clip=0.25, prec="bf16")
g = torch.Generator().manual_seed(1111)
T, B = 100, 32
x = torch.randint(0, bench.V, (T, B), generator=g).to(dev)
y = torch.randint(0, bench.V, (T, B), generator=g).to(dev)
ft.capture(T, B, 1e-3)
for i in range(5):
    ft.step_captured(x, y, 7 + i)
torch.cuda.synchronize()
N = 5
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(N):
        ft.step_captured(x, y, 100 + i)
    torch.cuda.synchronize()
import json, tempfile
path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
tr = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
streams = collections.defaultdict(lambda: [0.0, 0])
for e in tr:
    streams[e["args"].get("stream")][0] += e["dur"]
    streams[e["args"].get("stream")][1] += 1
print("per stream (us per step, kernels per step):", {k: (round(v[0] / N, 1), v[1] // N) for k, v in streams.items()})
main = max(streams, key=lambda k: streams[k][0])
side = collections.defaultdict(lambda: [0.0, 0])
for e in tr:
    if e["args"].get("stream") != main:
        k = e["name"].split("(")[0].replace("void ", "").replace("blm::", "")[:60]
        side[k][0] += e["dur"]; side[k][1] += 1
print("side stream(s):", {k: (round(v[0] / N, 1), v[1] // N) for k, v in sorted(side.items(), key=lambda kv: -kv[1][0])})
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
tot = collections.defaultdict(lambda: [0.0, 0])
for e in ev:
    k = e.name.split("(")[0].replace("void ", "").replace("blm::", "")[:70]
    tot[k][0] += e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
    tot[k][1] += 1
allk = sum(v[0] for v in tot.values())
t0 = min(e.time_range.start for e in ev); t1 = max(e.time_range.end for e in ev)
print(f"{N} steps: wall {((t1 - t0) / N):.1f} us per step, sum of kernel times {allk / N:.1f} us per step, {len(ev) // N} kernels per step")
for k, (us, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"  {us / N:8.1f} us  {100 * us / allk:5.1f}%  x{n // N:3d}  {k}")
