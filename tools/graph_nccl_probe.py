"""Probe: can the sharded step (NCCL collectives inside) be captured and replayed as a CUDA graph?"""
import os, sys, time, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clip_event_b200 import distributed as cd, synthetic as syn
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
w = syn.WORKLOADS["c3"]; dt = torch.bfloat16
img, txt, ls = syn.contrastive_inputs(w.B, w.T, w.D, 0, "trained", dtype=dt)
lo, hi = cd.shard_bounds(w.B, world, rank); b = hi - lo
li, lt, ip = cd.global_labels_for_rank(b, w.T, world, rank, device=dev)
img_l = img[lo:hi].to(dev).requires_grad_(True); txt_l = txt[lo * w.T:hi * w.T].to(dev).requires_grad_(True)
lsg = ls.to(dev).requires_grad_(True)
out = torch.zeros(2, device=dev)
def step():
    img_l.grad = None; txt_l.grad = None; lsg.grad = None
    a, c = cd.global_contrastive(img_l, txt_l, lsg, li, lt, ip)
    (a + c).backward()
    out.copy_(torch.stack([a.detach(), c.detach()]))
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3): step()
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize(); dist.barrier()
print("rank", rank, "eager ok", out.tolist(), flush=True)
mode = sys.argv[1] if len(sys.argv) > 1 else "thread_local"
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g, capture_error_mode=mode):
        step()
    torch.cuda.synchronize()
    print("rank", rank, "capture ok", flush=True)
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(20): g.replay()
    e1.record(); torch.cuda.synchronize()
    print("rank", rank, "replay ok %.1f us/step" % (e0.elapsed_time(e1) / 20 * 1e3), out.tolist(), flush=True)
except Exception as e:
    print("rank", rank, "capture/replay FAILED:", repr(e)[:300], flush=True)
dist.barrier()
dist.destroy_process_group()
