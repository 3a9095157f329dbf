"""Per-step time of the recurrence on one-tile batches (the hypothesis-#0 chains: one row per session), H = 1024:
whole-K loads (default for <= 16 rows) vs the ring with short boxes (BLM_LSTM_NO_WHOLE_K=1) vs 128-row boxes
(BLM_LSTM_FULL_BOX=1); all three must agree bit for bit."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayeslms_b200 import _lib, ops
_lib.init(0)
DEV = "cuda:0"
H = 1024
for B in (12, 16, 100):
    T = 400
    gx = torch.randn(T * B, 4 * H, device=DEV)
    w = ops.split(torch.randn(4 * H, H, device=DEV) / 32, "bf16")
    h0 = torch.zeros(B, H, device=DEV); c0 = torch.zeros(B, H, device=DEV)
    lengths = torch.full((B,), T, dtype=torch.int32, device=DEV)
    res = {}
    for mode in ("short", "ring", "full"):
        os.environ.pop("BLM_LSTM_FULL_BOX", None); os.environ.pop("BLM_LSTM_NO_WHOLE_K", None)
        if mode == "full": os.environ["BLM_LSTM_FULL_BOX"] = "1"
        if mode == "ring": os.environ["BLM_LSTM_NO_WHOLE_K"] = "1"
        for prec in ("bf16", "bf16x3"):
            ws = w if prec == "bf16" else ops.split(w.hi.float(), "bf16x3")
            run = lambda: ops.lstm_layer(gx, ws, h0, c0, lengths, T, B, H, prec=prec, want_f32=True)
            for _ in range(2): out = run()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3): out = run()
            e1.record(); torch.cuda.synchronize()
            res[(mode, prec)] = out[0].clone()
            print(f"B {B} {mode} box {prec}: {e0.elapsed_time(e1) / 3 / T * 1000:.2f} us/step", flush=True)
    for prec in ("bf16", "bf16x3"):
        assert torch.equal(res[("short", prec)], res[("full", prec)]) and torch.equal(res[("ring", prec)], res[("full", prec)]), (B, prec)
print("identical")
