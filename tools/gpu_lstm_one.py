"""One recurrence launch at H = 1024 (argv: B T [rows32] [bf16x3]) -- the target of an ncu capture of the LSTM kernels."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayeslms_b200 import _lib, ops
_lib.init(0)
B, T, H = int(sys.argv[1]), int(sys.argv[2]), 1024
r32 = "rows32" in sys.argv
prec = "bf16x3" if "bf16x3" in sys.argv else "bf16"
gx = ops.rows32_empty(T * B, 4 * H, torch.device("cuda:0")).normal_() if r32 else torch.randn(T * B, 4 * H, device="cuda:0")
w = ops.split(torch.randn(4 * H, H, device="cuda:0") / 32, prec)
z = torch.zeros(B, H, device="cuda:0")
lengths = torch.full((B,), T, dtype=torch.int32, device="cuda:0")
ops.lstm_layer(gx, w, z, z.clone(), lengths, T, B, H, prec=prec, gx_rows32=r32)
torch.cuda.synchronize()
