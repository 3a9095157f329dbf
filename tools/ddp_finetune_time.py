"""Fine-tune step (BASELINE config 4) under torchrun: captured step with and without the overlapped all-reduce.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/ddp_finetune_time.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from bayeslms_b200 import model as M, trainer

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
T, B = 100, 32
g = torch.Generator().manual_seed(1111 + rank)
x = torch.randint(0, bench.V, (T, B), generator=g).to(dev)
y = torch.randint(0, bench.V, (T, B), generator=g).to(dev)
for mode in ("nccl_in_graph", "eager_allreduce", "nccl_in_graph", "eager_allreduce"):
    trainer._OVERLAP = mode == "two_graphs"
    trainer._NCCL_IN_GRAPH = mode == "nccl_in_graph"
    overlap = mode
    torch.manual_seed(1111)
    net = M.VTransformerModel(bench.V, bench.D, bench.NHEAD, bench.FF, bench.NLAYERS, 0.0, True, "11").to(dev).train()
    ft = trainer.FineTuner(net, 0.01, clip=0.25, prec="bf16")
    ft.capture(T, B, 1e-3)
    for i in range(5):
        ft.step_captured(x, y, 7 + i)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(30):
        ft.step_captured(x, y, 100 + i)
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 30], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"world {world} {overlap} (in-graph {ft._cap.get('ar_in_graph')}, split {ft._cap.get('split')}): {ms.item():.3f} ms per step, "
              f"{T * B * world / ms.item() * 1e3 / 1e6:.2f} M tokens/s", flush=True)
    ft._cap = None
    del ft, net
    torch.cuda.synchronize()
dist.barrier()
dist.destroy_process_group()
