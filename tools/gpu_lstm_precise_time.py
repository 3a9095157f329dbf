"""Precise-mode (bf16x3) LSTM: recurrence step time at H = 1024 and end-to-end rescoring throughput on a 3-session slice
of the bench workload."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from bayeslms_b200 import _lib, model as M, ops, synth
from bayeslms_b200.scorer import Rescorer
_lib.init(0)
DEV = "cuda:0"
H = 1024
for B in (512, 2048):
    T = 20
    gx = ops.rows32_empty(T * B, 4 * H, DEV).normal_()
    w = ops.split(torch.randn(4 * H, H, device=DEV) / 32, "bf16x3")
    h0 = torch.zeros(B, H, device=DEV); c0 = torch.zeros(B, H, device=DEV)
    lengths = torch.full((B,), T, dtype=torch.int32, device=DEV)
    run = lambda: ops.lstm_layer(gx, w, h0, c0, lengths, T, B, H, prec="bf16x3", gx_rows32=True)
    for _ in range(2): run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): run()
    e1.record(); torch.cuda.synchronize()
    print(f"precise recurrence B {B}: {e0.elapsed_time(e1) / 3 / T * 1000:.1f} us/step", flush=True)
dev = torch.device(DEV)
torch.manual_seed(1111)
V = bench.V
net = M.BayesRNNModel("LSTM", V, 1024, 1024, 2, 0.5, True, 3).to(dev).eval()
n_sess, per_sess, nbest = 3, 100, 100
data = synth.make_nbest(n_sess * per_sess, nbest, V, seed=1112)
tok, tgt, _, offs = data.flat_host()
utt = np.repeat(np.arange(n_sess * per_sess), [len(u) for u in data.hyps])
sess_of, utt_of = (utt // per_sess).astype(np.int32), (utt % per_sess).astype(np.int32)
for prec in ("bf16", "bf16x3"):
    rs = Rescorer(net, prec=prec, max_tokens=bench.MAX_TOKENS)
    rs.score_sessions_flat(tok, tgt, offs, sess_of, utt_of)
    torch.cuda.synchronize()
    ops.STATS.timing = {}
    t0 = time.perf_counter()
    rs.score_sessions_flat(tok, tgt, offs, sess_of, utt_of)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    timing, ops.STATS.timing = ops.STATS.timing, None
    rows = sorted(((sum(a.elapsed_time(b) for a, b, _ in v), len(v), k) for k, v in timing.items()), reverse=True)
    print(f"{prec}: {data.n_tokens() / dt / 1e6:.2f} M tokens/s ({dt * 1e3:.1f} ms); " + ", ".join(f"{k} {ms:.1f} ms" for ms, n, k in rows[:5]), flush=True)
