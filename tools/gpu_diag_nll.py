import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayeslms_b200 import _lib, ops
torch.manual_seed(0); dev = torch.device("cuda:0"); _lib.init(0)
for M, V, K in [(40000, 30000, 512), (3000, 30000, 512), (40000, 30000, 1024)]:
    h = torch.randn(M, K, device=dev); e = (torch.rand(V, K, device=dev) - 0.5) * 0.2
    b = (torch.rand(V, device=dev) - 0.5) * 0.2
    t = torch.randint(0, V, (M,), device=dev, dtype=torch.int32)
    H, E = ops.split(h, "bf16"), ops.split(e, "bf16")
    nll = ops.vocab_nll(H, E, b, t, prec="bf16"); torch.cuda.synchronize()
    ref = torch.empty(M, dtype=torch.float64, device=dev)
    for i in range(0, M, 4000):
        lg = H.hi[i:i+4000].double() @ E.hi.double().T + b.double()
        ref[i:i+4000] = torch.logsumexp(lg, -1) - lg.gather(1, t[i:i+4000].long().view(-1, 1)).squeeze(1)
    err = (nll.double() - ref).abs()
    print(os.environ.get("BLM_NLL_NO_ARES"), M, V, K, "max", err.max().item(), "mean", err.mean().item(), "n>5e-5", int((err > 5e-5).sum()),
          "argmax row", int(err.argmax()), "signed mean", (nll.double() - ref).mean().item(), flush=True)
