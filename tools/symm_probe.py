"""Does torch.distributed._symmetric_memory work on this box (peer pointers, barrier, inside a CUDA graph)?
torchrun --nproc-per-node 2 tools/symm_probe.py"""
import os, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
import torch.distributed._symmetric_memory as symm
dev = torch.device("cuda", torch.cuda.current_device())
t = symm.empty(1 << 20, dtype=torch.float32, device=dev)
h = symm.rendezvous(t, dist.group.WORLD)
print(rank, "rendezvous ok; buffer_ptrs", [hex(p) for p in h.buffer_ptrs][:4], "signal_pad_ptrs", len(h.signal_pad_ptrs), flush=True)
t.fill_(float(rank + 1))
h.barrier(channel=0)
peer = h.get_buffer((rank + 1) % world, (1 << 20,), torch.float32)
print(rank, "peer value", peer[:4].tolist(), flush=True)
# timing: barrier + pull copy, eager and in a graph
out = torch.empty(world << 20, dtype=torch.float32, device=dev)
def step():
    h.barrier(channel=0)
    for r in range(world):
        out[r << 20:(r + 1) << 20].copy_(h.get_buffer(r, (1 << 20,), torch.float32))
    h.barrier(channel=1)
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): step()
e1.record(); torch.cuda.synchronize()
print(rank, "eager: %.1f us per (2 barriers + %d x 4 MB pulls)" % (e0.elapsed_time(e1) / 20 * 1e3, world), flush=True)
try:
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        step()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    with torch.cuda.graph(g):
        step()
    torch.cuda.synchronize(); dist.barrier()
    e0.record()
    for _ in range(20): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(rank, "graph: %.1f us per step; values ok: %s" % (e0.elapsed_time(e1) / 20 * 1e3, out[::1 << 20].tolist()), flush=True)
except Exception as ex:
    print(rank, "graph capture failed:", repr(ex)[:300], flush=True)
# barrier alone
e0.record()
for _ in range(50): h.barrier(channel=0)
e1.record(); torch.cuda.synchronize()
print(rank, "barrier alone: %.1f us" % (e0.elapsed_time(e1) / 50 * 1e3), flush=True)
dist.destroy_process_group()
