"""Per-step time of the LSTM recurrence at H = 1024: CTA-pair kernel (default) vs the single-CTA kernel
(BLM_LSTM_NO_PAIR=1), gates_x row-major vs in 32-row blocks."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayeslms_b200 import _lib, ops
_lib.init(0)
DEV = "cuda:0"
H = 1024
for B in (512, 1024, 2048):
    T = 40
    gx = torch.randn(T * B, 4 * H, device=DEV)
    w = ops.split(torch.randn(4 * H, H, device=DEV) / 32, "bf16")
    h0 = torch.zeros(B, H, device=DEV); c0 = torch.zeros(B, H, device=DEV)
    lengths = torch.full((B,), T, dtype=torch.int32, device=DEV)
    for mode, r32 in (("pair", True), ("pair", False), ("single", True), ("single", False)):
        if mode == "single": os.environ["BLM_LSTM_NO_PAIR"] = "1"
        else: os.environ.pop("BLM_LSTM_NO_PAIR", None)
        run = lambda: ops.lstm_layer(gx, w, h0, c0, lengths, T, B, H, prec="bf16", gx_rows32=r32)  # noqa: E731
        for _ in range(2): run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): run()
        e1.record(); torch.cuda.synchronize()
        print(f"B {B} {mode} {'rows32' if r32 else 'row-major'}: {e0.elapsed_time(e1) / 5 / T * 1000:.1f} us/step", flush=True)
