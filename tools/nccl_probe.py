"""NCCL collective latencies on this box (per op, CUDA events, eager and inside a CUDA graph).
torchrun --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/nccl_probe.py
"""
import os
import sys
import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
R, D = 4096, 512
b = R // world
small = torch.ones(4, device=dev)
img = torch.randn(b, D, device=dev).bfloat16()
img_all = torch.empty(R, D, device=dev, dtype=torch.bfloat16)
stats = torch.randn(R * 4 + 4, device=dev)
stats_all = torch.empty(world * (R * 4 + 4), device=dev)
dhat = torch.randn(R, D, device=dev)
mine = torch.empty(b, D, device=dev)

ops = {
    "all_reduce 16 B": lambda: dist.all_reduce(small),
    "all_gather img %d KB/rank" % (b * D * 2 // 1024): lambda: dist.all_gather_into_tensor(img_all.view(-1), img.view(-1)),
    "all_gather stats %d KB/rank" % ((R * 4 + 4) * 4 // 1024): lambda: dist.all_gather_into_tensor(stats_all, stats),
    "reduce_scatter fp32 %d KB in" % (R * D * 4 // 1024): lambda: dist.reduce_scatter_tensor(mine, dhat),
}


def time_fn(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for name, fn in ops.items():
    eager = time_fn(fn)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, capture_error_mode="thread_local"):
        for _ in range(10):
            fn()
    torch.cuda.synchronize()
    graphed = time_fn(g.replay, 10) / 10
    if rank == 0:
        print("%-40s eager %7.1f us   in-graph %7.1f us" % (name, eager, graphed), flush=True)
dist.barrier(); torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0)
