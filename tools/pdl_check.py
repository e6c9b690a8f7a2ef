"""A/B of programmatic dependent launch (CE_PDL, csrc/ce_common.cuh) in ONE process: the bench step of a workload is
captured as a CUDA graph without and with the programmatic edges; the PDL graph is replayed many times and every
replay's losses and gradients are compared with the plain graph's (the chain must stay data-race free), then both
graphs are timed with CUDA events.   usage: python tools/pdl_check.py [c3|c2|c4] [bf16|fp32] [replays]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from clip_event_b200 import synthetic as syn  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
    dt = sys.argv[2] if len(sys.argv) > 2 else "bf16"
    replays = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    dtype = torch.bfloat16 if dt == "bf16" else torch.float32
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
    w = syn.WORKLOADS[wl]
    h = bench.Harness(w, dtype, 1, 0, dev, use_graph=True)

    def grads():      # the tensors one capture's replays write into (live references, not copies)
        out = {k: v.grad for k, v in h.leaves.items()}
        out["ls"] = h.head.logit_scale.grad
        return out

    def snapshot(refs):
        torch.cuda.synchronize()
        out = {k: v.detach().float().clone().reshape(-1) for k, v in refs.items()}
        out["losses"] = h.losses_out.detach().clone()
        return out

    graphs, outs = {}, {}
    for mode in ("0", "1"):
        os.environ["CE_PDL"] = mode
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):          # as bench.py warms up: never on the legacy default stream
            for _ in range(3):
                h.step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        run, graphed = h.capture(h.step)
        if not graphed:
            raise SystemExit("pdl_check: CUDA graph capture failed with CE_PDL=%s" % mode)
        graphs[mode], outs[mode] = run, grads()
    graphs["0"]()
    ref = snapshot(outs["0"])
    scale = {k: v.abs().max().clamp_min(1e-30) for k, v in ref.items()}
    # accumulation order (red.add) differs from run to run: allow a few ulps of the output dtype on the largest
    # element; a stale read or a race shows up as an error of the order of the gradient itself
    tol = 2e-2 if dtype == torch.bfloat16 else 1e-4
    worst = {k: 0.0 for k in ref}
    # the plain graph against itself first: the run-to-run noise floor
    graphs["0"]()
    again = snapshot(outs["0"])
    floor = {k: float(((again[k] - ref[k]).abs().max() / scale[k]).item()) for k in ref}
    ok = True
    for i in range(replays):
        if h.flush is not None and i % 2 == 0:
            h.flush.zero_()
        graphs["1"]()
        got = snapshot(outs["1"])
        for k in ref:
            e = float(((got[k] - ref[k]).abs().max() / scale[k]).item())
            worst[k] = max(worst[k], e)
            if not (e <= tol) or not torch.isfinite(got[k]).all():
                ok = False
    # interleaved rounds (the SM clock drifts under the power cap): medians per graph and of the per-round differences
    import statistics
    t0, t1 = [], []
    for _ in range(9):
        t0.append(h.timed_loop(graphs["0"], 20, 3))
        t1.append(h.timed_loop(graphs["1"], 20, 3))
    ms = {"0": statistics.median(t0), "1": statistics.median(t1)}
    diff = statistics.median([b - a for a, b in zip(t0, t1)])
    print("pdl_check %s %s: plain graph %.4f ms/step, PDL graph %.4f ms/step (median of 9 interleaved rounds; "
          "median difference %+.1f us, PDL faster in %d of 9)" %
          (wl, dt, ms["0"], ms["1"], diff * 1e3, sum(1 for a, b in zip(t0, t1) if b < a)))
    print("  noise floor (plain vs plain): " + " ".join("%s=%.2e" % kv for kv in floor.items()))
    print("  worst over %d PDL replays:     " % replays + " ".join("%s=%.2e" % kv for kv in worst.items()))
    print("pdl_check: %s" % ("OK" if ok else "MISMATCH"))
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
