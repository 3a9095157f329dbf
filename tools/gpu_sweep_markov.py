"""Scratch sweep: how to get a peaked model on the Markov corpus quickly (loss curve per setting)."""
import itertools, sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayeslms_b200 import model as M, synth, train as T
from bayeslms_b200.trainer import FineTuner
for V, steps, lr, clip, layers in [(500, 1000, 0.1, 1.0, 2), (500, 1000, 0.3, 1.0, 2), (500, 1000, 1.0, 1.0, 2), (500, 1000, 0.3, 5.0, 2),
                                   (2000, 2000, 0.3, 1.0, 2), (2000, 2000, 1.0, 1.0, 2), (2000, 2000, 0.3, 5.0, 2), (500, 2000, 0.3, 1.0, 6),
                                   (30000, 1500, 0.3, 1.0, 6)]:
    torch.manual_seed(1111)
    mk = synth.make_markov(V, 3)
    net = M.BayesTransformerModel(V, 512, 8, 4096, layers, 0.0, True, "FFN").to("cuda:0")
    ft = FineTuner(net, lr, clip=clip, prec="bf16")
    ids = torch.from_numpy(mk.stream(32 * (100 * 400 + 1)))
    t0 = time.time()
    losses = T.train_steps(ft, ids, 32, 100, steps)
    torch.cuda.synchronize()
    print(V, steps, lr, clip, layers, f"{time.time()-t0:.1f}s", " ".join(f"{l:.2f}" for l in losses[::max(1, len(losses)//12)]), f"last {losses[-1]:.3f}", flush=True)
    del ft, net
