"""Per-kernel time breakdown of one fine-tune step (BASELINE config 4 shape) from CUDA-event brackets."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from bayeslms_b200 import _lib, ops, model as M
from bayeslms_b200.trainer import FineTuner

_lib.init(0)
dev = torch.device("cuda:0")
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
family = sys.argv[2] if len(sys.argv) > 2 else "v_tm"
torch.manual_seed(1111)
if family == "lstm":   # Bayes-LSTM 2 x 1024, gate 3 (BASELINE config 1 / 5 architecture), batch 32 x 100
    net = M.BayesRNNModel("LSTM", bench.V, 1024, 1024, 2, 0.0, True, 3).to(dev).train()
else:
    net = M.VTransformerModel(bench.V, bench.D, bench.NHEAD, bench.FF, bench.NLAYERS, 0.0, True, "11").to(dev).train()
ft = FineTuner(net, 0.01, clip=0.25, prec=prec)
g = torch.Generator().manual_seed(1)
x = torch.randint(0, bench.V, (100, 32), generator=g).to(dev)
y = torch.randint(0, bench.V, (100, 32), generator=g).to(dev)
for i in range(3):
    ft.step(x, y, 1e-3, seed=i)
torch.cuda.synchronize()
ops.STATS.timing = {}
ops.STATS.launches = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ft.step(x, y, 1e-3, seed=9)
e1.record()
torch.cuda.synchronize()
tot = e0.elapsed_time(e1)
rows = sorted(((sum(a.elapsed_time(b) for a, b, _ in ev), len(ev), k) for k, ev in ops.STATS.timing.items()), reverse=True)
print(f"step {tot:.3f} ms, {ops.STATS.launches} launches of our kernels, prec {prec}")
acc = 0.0
for t, n, k in rows:
    acc += t
    print(f"  {k:28s} {n:4d} launches {t*1e3:9.1f} us  ({100*t/tot:5.1f} %)")
print(f"  bracketed total {acc:.3f} ms")
