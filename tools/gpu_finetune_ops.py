"""Per-op CUDA-event times of ONE eager fine-tune step (BASELINE config 4) at dropout 0 and 0.2: where the dropout
overhead of the step sits.  Eager launches carry ~10 us of launch latency each; the captured step is what bench.py times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from bayeslms_b200 import _lib, model as M, ops
from bayeslms_b200.trainer import FineTuner
_lib.init(0)
dev = torch.device("cuda:0")
T, B = 100, 32
res = {}
for dr in (0.0, 0.2):
    torch.manual_seed(1)
    net = M.VTransformerModel(bench.V, bench.D, bench.NHEAD, bench.FF, 5, dr, True, "11").to(dev).train()
    ft = FineTuner(net, 0.01, clip=0.25, prec="bf16")
    g = torch.Generator().manual_seed(2)
    x = torch.randint(0, bench.V, (T, B), generator=g).to(dev)
    y = torch.randint(0, bench.V, (T, B), generator=g).to(dev)
    for i in range(3):
        ft.step(x, y, 1e-3, seed=5 + i)
    torch.cuda.synchronize()
    ops.STATS.timing = {}
    ft.step(x, y, 1e-3, seed=9)
    torch.cuda.synchronize()
    timing, ops.STATS.timing = ops.STATS.timing, None
    res[dr] = {k: (sum(a.elapsed_time(b) for a, b, _ in v) * 1e3, len(v)) for k, v in timing.items()}
keys = sorted(set(res[0.0]) | set(res[0.2]), key=lambda k: -(res[0.2].get(k, (0, 0))[0] - res[0.0].get(k, (0, 0))[0]))
print(f"{'op':28s} {'p=0 us':>10s} {'n':>4s} {'p=0.2 us':>10s} {'n':>4s} {'delta':>9s}")
for k in keys:
    a, na = res[0.0].get(k, (0.0, 0)); b, nb = res[0.2].get(k, (0.0, 0))
    print(f"{k:28s} {a:10.1f} {na:4d} {b:10.1f} {nb:4d} {b - a:9.1f}")
print("total", sum(v[0] for v in res[0.0].values()), sum(v[0] for v in res[0.2].values()))
