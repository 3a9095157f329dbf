"""Where the LSTM rescoring pass spends its time: wall clock vs summed kernel time per op (CUDA events), bench workload."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from bayeslms_b200 import model as M, ops, synth
from bayeslms_b200.scorer import Rescorer
dev = torch.device("cuda", 0)
torch.manual_seed(1111)
V = bench.V
net = M.BayesRNNModel("LSTM", V, 1024, 1024, 2, 0.5, True, 3).to(dev).eval()
n_sess, per_sess, nbest = 12, 100, 100
data = synth.make_nbest(n_sess * per_sess, nbest, V, seed=1112)
tok, tgt, _, offs = data.flat_host()
utt = np.repeat(np.arange(n_sess * per_sess), [len(u) for u in data.hyps])
sess_of, utt_of = (utt // per_sess).astype(np.int32), (utt % per_sess).astype(np.int32)
rs = Rescorer(net, prec="bf16", max_tokens=bench.MAX_TOKENS)
rs.score_sessions_flat(tok, tgt, offs, sess_of, utt_of)
torch.cuda.synchronize()
for timed in (False, True):
    ops.STATS.timing = {} if timed else None
    t0 = time.perf_counter()
    rs.score_sessions_flat(tok, tgt, offs, sess_of, utt_of)
    torch.cuda.synchronize()
    print(f"wall {1e3 * (time.perf_counter() - t0):.1f} ms (events {'on' if timed else 'off'})")
timing, ops.STATS.timing = ops.STATS.timing, None
rows = sorted(((sum(a.elapsed_time(b) for a, b, _ in v), len(v), k) for k, v in timing.items()), reverse=True)
print(f"kernel total {sum(r[0] for r in rows):.1f} ms")
for ms, n, k in rows:
    print(f"  {k:24s} {ms:8.2f} ms  {n:5d} launches")
if os.environ.get("BLM_PROFILE"):
    import cProfile, pstats
    pr = cProfile.Profile()
    pr.enable()
    rs.score_sessions_flat(tok, tgt, offs, sess_of, utt_of)
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
