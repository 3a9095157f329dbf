"""Captured fine-tune step (BASELINE config 4) on one GPU: ms per step.  A/B switches are environment variables read at
import (BLM_TRAIN_WGRAD_STREAM=0, ...), so run once per setting."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from bayeslms_b200 import _lib
_lib.init(0)
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
r = bench.bench_finetune(dev, 1, 50)
print(json.dumps({k: r[k] for k in ("ms_per_step", "tokens_per_s", "loss_first", "loss_last")}), {k: v for k, v in os.environ.items() if k.startswith("BLM_")})
