"""Times the contrastive chain (fwd, bwd) for a workload with CUDA-graph replay."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clip_event_b200 import functional as F_, synthetic as syn
wl, dt = sys.argv[1], (torch.bfloat16 if sys.argv[2] == "bf16" else torch.float32)
w = syn.WORKLOADS[wl]
img, txt, ls = syn.contrastive_inputs(w.B, w.T, w.D, 0, "trained", dtype=dt)
lpi, lpt, idx = (t.cuda() for t in syn.contrastive_labels(w.B, w.T))
img, txt, ls = img.cuda().requires_grad_(True), txt.cuda().requires_grad_(True), ls.cuda().requires_grad_(True)
def fwd():
    with torch.no_grad():
        return F_.contrastive_over_batch(img, txt, ls, lpi, lpt, idx)
def full():
    img.grad = None; txt.grad = None; ls.grad = None
    a, b = F_.contrastive_over_batch(img, txt, ls, lpi, lpt, idx)
    (a + b).backward()
def time_it(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s): fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    with torch.cuda.graph(g): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
tf, tb = time_it(fwd), time_it(full)
fl = 2.0 * w.B * w.B * w.T * w.D
print("%s %s: fwd %.1f us (%.0f TF/s on 2BCD)  fwd+bwd %.1f us (%.0f TF/s on 6BCD)" % (wl, sys.argv[2], tf, fl / tf / 1e6, tb, 3 * fl / tb / 1e6))
