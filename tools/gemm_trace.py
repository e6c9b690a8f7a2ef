"""Pipeline timeline of the tcgen05 GEMM kernels (tuning aid).

Build with tracing compiled in, run one backward of the contrastive chain with a trace buffer, and
print per-tile clock64 deltas of CTA 0's producer / MMA / epilogue roles:

    CE_EXTRA_NVCC_FLAGS=-DCE_GEMM_TRACE python -m clip_event_b200.build --force   (then ship the .so)
    python tools/gemm_trace.py c3 3

The second argument is the ordinal of the GEMM launch to trace in one forward + backward of the
over-batch loss: 1 statistics, 2 statistics (positives), 3 gradient, 4 gradient (positives),
5 G*txt, 6 Gt^t*pos, 7 G^t*img, 8 Gt*img.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

buf = torch.zeros(64 * 4 * 8, dtype=torch.int64, device="cuda")
os.environ["CE_GEMM_TRACE_PTR"] = hex(buf.data_ptr())
os.environ["CE_GEMM_TRACE_LAUNCH"] = sys.argv[2] if len(sys.argv) > 2 else "3"
from clip_event_b200 import _lib as L, functional as F_, synthetic as syn

w = syn.WORKLOADS[sys.argv[1]]
img, txt, ls = syn.contrastive_inputs(w.B, w.T, w.D, 0, "trained", dtype=torch.bfloat16)
lpi, lpt, idx = (t.cuda() for t in syn.contrastive_labels(w.B, w.T))
img, txt, ls = img.cuda().requires_grad_(True), txt.cuda().requires_grad_(True), ls.cuda().requires_grad_(True)
lib = L.load()
a, b = F_.contrastive_over_batch(img, txt, ls, lpi, lpt, idx)
(a + b).backward()
torch.cuda.synchronize()
t = buf.cpu().view(64, 4, 8)
names = {0: ["tile", "empty0 ok", "issued"], 1: ["tile", "tempty ok", "full0 ok", "fullN ok", "commit"],
         2: ["tile", "bar ok", "prefetched", "tfull ok", "chunks done", "row_end", "walked"], 3: None}
names[3] = names[2]
base = int(t[0, 1, 0])
print("GEMM launch %s; cycles relative to the MMA thread's first tile" % os.environ["CE_GEMM_TRACE_LAUNCH"])
for seq in range(64):
    if int(t[seq, 1, 0]) == 0:
        break
    line = ["tile %2d" % seq]
    for role, tag in ((0, "TMA"), (1, "MMA"), (2, "EPI0"), (3, "EPI1")):
        vals = [int(v) - base for v in t[seq, role, :len(names[role])]]
        line.append("%s %s" % (tag, " ".join("%6d" % v for v in vals)))
    print(" | ".join(line))
