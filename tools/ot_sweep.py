"""BASELINE config c5: OT sweep -- text nodes 4-64 x image nodes 50-577, IPOT iterations 10-100, batch 512,
one B200.  Times ce_ot_fwd_bwd (forward + gradients, CUDA-graph replay, CUDA events) for every cell and
reports the algorithmic HBM rate 2 (M+N) D e B / t against the measured copy bandwidth.
python tools/ot_sweep.py [out.json] [bf16|fp32]   (second argument: one dtype only)"""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clip_event_b200 import functional as F_, synthetic as syn

B = 512
peak = 6546.2
try:
    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
        peak = float(json.load(f).get("hbm_gbs", peak))
except Exception:
    pass


def time_it(fn, n=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    with torch.cuda.graph(g):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


rows = []
only = sys.argv[2] if len(sys.argv) > 2 else None
only_m = int(sys.argv[3]) if len(sys.argv) > 3 else None      # third argument: one text-node count only
for dt in (torch.bfloat16, torch.float32):
    if only is not None and only != ("bf16" if dt == torch.bfloat16 else "fp32"):
        continue
    for D in (512, 768):
        for M in (4, 8, 16, 32, 64):
            if only_m is not None and M != only_m:
                continue
            for N in (50, 197, 257, 577):
                etxt, obj, tnum, onum = syn.ot_inputs(B, M, N, D, 0, "ragged", dtype=dt)
                etxt, obj, tnum, onum = etxt.cuda(), obj.cuda(), tnum.cuda(), onum.cuda()
                eg, og = etxt.clone().requires_grad_(True), obj.clone().requires_grad_(True)
                for iters in (10, 25, 50, 100):
                    if dt == torch.float32 and iters != 50:
                        continue
                    def full():
                        eg.grad = None; og.grad = None
                        l, _ = F_.ot_alignment(eg, og, tnum, onum, iters=iters)
                        l.backward()
                    t = time_it(full)
                    nbytes = 2 * (M + N) * D * etxt.element_size() * B
                    rows.append(dict(dtype=str(dt)[6:], D=D, M=M, N=N, iters=iters, us=round(t, 1),
                                     algorithmic_MB=round(nbytes / 1e6, 1), GBs=round(nbytes / t / 1e3, 1),
                                     frac=round(nbytes / t / 1e3 / peak, 3)))
                    print(rows[-1], flush=True)
out = dict(workload="c5: OT sweep, batch 512, ragged masks", peak_gbs=peak, cells=rows)
if len(sys.argv) > 1:
    with open(sys.argv[1], "w") as f:
        json.dump(out, f, indent=1)
