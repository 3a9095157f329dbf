"""One persistent-LSTM layer launch (B=2048, H=1024, bf16) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayeslms_b200 import _lib, ops
_lib.init(0)
dev = torch.device("cuda:0")
T, B, H = int(os.environ.get("T", 8)), int(os.environ.get("B", 2048)), 1024
torch.manual_seed(0)
gx = torch.randn(T * B, 4 * H, device=dev) * 0.1
whh = ops.split(torch.randn(4 * H, H, device=dev) * 0.03, "bf16")
h0 = torch.zeros(B, H, device=dev); c0 = torch.zeros(B, H, device=dev)
lens = torch.full((B,), T, dtype=torch.int32, device=dev)
for _ in range(2):
    out = ops.lstm_layer(gx, whh, h0, c0, lens, T, B, H, prec="bf16", want_f32=False, want_split=True)
torch.cuda.synchronize()
print("ok", float(out[2].abs().sum()))
