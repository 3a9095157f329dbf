"""One recurrence launch at H = 1024 (argv: B T) -- the target of an ncu capture of the LSTM kernels."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayeslms_b200 import _lib, ops
_lib.init(0)
B, T, H = int(sys.argv[1]), int(sys.argv[2]), 1024
gx = torch.randn(T * B, 4 * H, device="cuda:0")
w = ops.split(torch.randn(4 * H, H, device="cuda:0") / 32, "bf16")
z = torch.zeros(B, H, device="cuda:0")
lengths = torch.full((B,), T, dtype=torch.int32, device="cuda:0")
ops.lstm_layer(gx, w, z, z.clone(), lengths, T, B, H, prec="bf16")
torch.cuda.synchronize()
