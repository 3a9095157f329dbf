"""Data-parallel fine-tune check, run under torchrun on >= 2 GPUs of one box:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/ddp_train_check.py
Every rank takes its slice of a global batch (T, B_global) of the Bayesian-FFN Transformer with the SAME
injected weight noise; after one step (NCCL all-reduce of the flat gradient buffer, clip on the averaged
gradient, SGD momentum) the parameters must equal those of a single-rank step on the whole batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from bayeslms_b200 import model as M
from bayeslms_b200.trainer import FineTuner

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
V, D, NHEAD, FF, T, Bg = 2000, 128, 2, 512, 32, 8 * world


def build():
    torch.manual_seed(3)
    net = M.BayesTransformerModel(V, D, NHEAD, FF, 2, 0.0, True, "FFN")
    return net.to(dev).train()


g = torch.Generator().manual_seed(1)
x = torch.randint(0, V, (T, Bg), generator=g).to(dev)
y = torch.randint(0, V, (T, Bg), generator=g).to(dev)
eps = {"layer0": torch.randn(D, FF, generator=g)}
kl_scale, lr = 0.2, 0.1

ddp = build()
ft = FineTuner(ddp, lr, clip=0.25, prec="bf16x3")
assert ft.world == world
b0, b1 = rank * Bg // world, (rank + 1) * Bg // world
l_ddp = ft.step(x[:, b0:b1].contiguous(), y[:, b0:b1].contiguous(), kl_scale, eps=eps)[0].clone()
dist.all_reduce(l_ddp)
l_ddp /= world

single = build()
fs = FineTuner(single, lr, clip=0.25, prec="bf16x3")
fs.world = 1                       # the reference step: one rank, the whole batch
l_one = fs.step(x, y, kl_scale, eps=eps)[0]
err = (ft.flat_p - fs.flat_p).abs().max().item()
upd = (fs.flat_v.abs().max() * lr).item()
if rank == 0:
    print(f"world {world}: loss ddp {l_ddp.item():.6f} single {l_one.item():.6f}; max |param diff| {err:.3e} vs max update {upd:.3e}",
          flush=True)
assert abs(l_ddp.item() - l_one.item()) < 1e-4
assert err <= 2e-3 * upd + 1e-7, (err, upd)
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("DDP CHECK OK", flush=True)
