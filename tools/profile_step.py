"""One warm-up + N rescoring steps of the bench workload, for ncu (short: ~50 k tokens per step)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from bayeslms_b200 import _lib, synth
from bayeslms_b200.scorer import Rescorer

ap = argparse.ArgumentParser()
ap.add_argument("--utts", type=int, default=64)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--prec", default="bf16")
ap.add_argument("--K", type=int, default=0)
a = ap.parse_args()
_lib.init(0)
net = bench.build_model(torch.device("cuda:0"))
data = synth.make_nbest(a.utts, bench.NBEST, bench.V, seed=1111)
tok, tgt, pos, offs = data.flat_host()
rs = Rescorer(net, prec=a.prec, K=a.K, seed=1111 if a.K else None, max_tokens=bench.MAX_TOKENS)
rs.score_packed_host(tok, tgt, pos, offs)
torch.cuda.synchronize()
for _ in range(a.steps):
    out = rs.score_packed_host(tok, tgt, pos, offs)
torch.cuda.synchronize()
print("tokens", int(offs[-1]), "score0", float(out[0]))
