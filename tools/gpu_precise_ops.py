"""Per-op CUDA-event times of one rescoring step of the bench workload in both precision modes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from bayeslms_b200 import _lib, ops, synth
from bayeslms_b200.scorer import Rescorer
_lib.init(0)
net = bench.build_model(torch.device("cuda:0"))
data = synth.make_nbest(bench.UTTS_PER_STEP, bench.NBEST, bench.V, seed=1111)
tok, tgt, pos, offs = data.flat_host()
res = {}
for prec in ("bf16", "bf16x3"):
    rs = Rescorer(net, prec=prec, max_tokens=bench.MAX_TOKENS)
    for _ in range(2):
        rs.score_packed_host(tok, tgt, pos, offs)
    torch.cuda.synchronize()
    ops.STATS.timing = {}
    rs.score_packed_host(tok, tgt, pos, offs)
    torch.cuda.synchronize()
    timing, ops.STATS.timing = ops.STATS.timing, None
    res[prec] = {k: (sum(a.elapsed_time(b) for a, b, _ in v), len(v), sum(w for _, _, w in v)) for k, v in timing.items()}
print(f"{'op':22s} {'bf16 ms':>9s} {'bf16x3 ms':>10s} {'ratio':>6s} {'x3 TF/s (3x flop)':>18s}")
for k in sorted(res["bf16x3"], key=lambda k: -res["bf16x3"][k][0]):
    a = res["bf16"].get(k, (0.0, 0, 0.0)); b = res["bf16x3"][k]
    tf = 3 * b[2] / (b[0] / 1e3) / 1e12 if b[2] else 0.0
    print(f"{k:22s} {a[0]:9.2f} {b[0]:10.2f} {b[0] / a[0] if a[0] else 0:6.2f} {tf:18.0f}")
print("total", sum(v[0] for v in res["bf16"].values()), sum(v[0] for v in res["bf16x3"].values()))
