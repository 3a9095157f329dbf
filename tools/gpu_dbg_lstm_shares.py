"""Kernel-time shares of the LSTM rescoring path (bench_lstm workload) + wall clock vs device time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from bayeslms_b200 import _lib, ops, synth, model as M
from bayeslms_b200.scorer import Rescorer
_lib.init(0)
dev = torch.device("cuda:0")
torch.manual_seed(1111)
net = M.BayesRNNModel("LSTM", bench.V, 1024, 1024, 2, 0.5, True, 3).to(dev).eval()
n_sess, per_sess, nbest = 8, 16, 100
data = synth.make_nbest(n_sess * per_sess, nbest, bench.V, seed=1112)
import numpy as np
tok, tgt, _, offs = data.flat_host()
utt = np.repeat(np.arange(n_sess * per_sess), [len(u) for u in data.hyps])
sess_of, utt_of = (utt // per_sess).astype(np.int32), (utt % per_sess).astype(np.int32)
utts = data.tokenised()
sessions = [utts[s * per_sess:(s + 1) * per_sess] for s in range(n_sess)]
ref = Rescorer(net, prec="bf16", max_tokens=bench.MAX_TOKENS).score_sessions(sessions)
got = Rescorer(net, prec="bf16", max_tokens=bench.MAX_TOKENS).score_sessions_flat(tok, tgt, offs, sess_of, utt_of)
print("flat == nested:", bool((ref == got).all()), float(abs(ref - got).max()))
n_tok = data.n_tokens()
for name, kw in (("mean", {}), ("K=8", {"K": 8, "seed": 1111})):
    rs = Rescorer(net, prec="bf16", max_tokens=bench.MAX_TOKENS, **kw)
    rs.score_sessions_flat(tok, tgt, offs, sess_of, utt_of)
    torch.cuda.synchronize()
    ops.STATS.timing = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    rs.score_sessions_flat(tok, tgt, offs, sess_of, utt_of)
    e1.record(); torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    timing, ops.STATS.timing = ops.STATS.timing, None
    sh = {k: sum(a.elapsed_time(b) for a, b, _ in v) for k, v in timing.items()}
    n = {k: len(v) for k, v in timing.items()}
    tot = sum(sh.values())
    print(f"{name}: wall {wall:.1f} ms, device span {e0.elapsed_time(e1):.1f} ms, sum of kernels {tot:.1f} ms, {n_tok / wall * 1e3 / 1e6:.2f} M tok/s")
    for k, v in sorted(sh.items(), key=lambda kv: -kv[1]):
        print(f"   {v:8.2f} ms {100 * v / tot:5.1f}%  x{n[k]:4d}  {k}")
