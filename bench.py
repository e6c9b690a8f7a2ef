#!/usr/bin/env python
"""Benchmark of the CLIP-Event loss head (similarity + InfoNCE + IPOT alignment), fwd + bwd.

    python bench.py --gpus N --steps K --warmup W [--workload c3] [--dtype bf16|fp32]
    python bench.py --impl reference ...         # the reference algorithm on the host CPU cores

One JSON line on rank 0 (see DESIGN.md "Measurement" for every field).  A step is one forward +
backward of the whole loss head over one synthetic batch:
  * ``value``     samples/s with the inputs already resident in HBM (CUDA-graph replay of the public
                  API's fwd+bwd, timed with CUDA events, max over ranks);
  * ``e2e``       the same through the reference-facing modules with HOST (pinned) inputs: every
                  step copies its inputs H2D and returns its losses D2H; two device input sets, so
                  the copy of step i+1 overlaps the kernels of step i (wall clock over the loop);
  * ``roofline``  for the dominant kernel chain, from CUDA-event timings taken in this run;
  * ``cpu_baseline`` the oracle (reference algorithm, PyTorch CPU) on a bounded sample, rank 0, N=1.
Multi-GPU (torchrun): strong scaling -- the GLOBAL batch of the workload is sharded over the ranks;
every rank scores against the global negative set (all-gather / reduce-scatter over NCCL).
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from clip_event_b200 import synthetic as syn  # noqa: E402

METRIC = "loss fwd+bwd samples/sec"
WORKLOAD_TEXT = {
    "c1": "ViT-B/32 shapes, batch 32, 1 pos + 4 hard-neg, OT 8x50",
    "c2": "ViT-B/32 shapes, batch 256, 1 pos + 8 hard-neg, OT 16x50",
    "c3": "ViT-B/32 shapes, global batch 4096, 1 pos + 8 hard-neg, all-gathered embeddings, OT 16x50",
    "c4": "ViT-L/14 shapes (768-d, 257 patches), batch 1024, 1 pos + 8 hard-neg, OT 32x257 IPOT 50 iters",
    "c5": "OT sweep corner: 64 text nodes x 577 image nodes, batch 512, 768-d",
}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        pw = []
        for r in self.rows:
            try:
                pw.append(float(r[2]))
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def algorithmic_work(w, dtype_bytes, B_rows, B_cols_total):
    """flops of the similarity GEMM chain and HBM bytes of OT, per step per rank."""
    flops = 6.0 * B_rows * (B_cols_total * w.T) * w.D       # rows scored x columns scored x D x (fwd + 2 bwd)
    return flops


def make_inputs(w, dtype, lo, hi):
    """Seeded global batch, sliced to this rank's shard [lo, hi).  Returns pinned host tensors."""
    img, txt, ls = syn.contrastive_inputs(w.B, w.T, w.D, 0, "trained")
    etxt, obj, tnum, onum = syn.ot_inputs(w.B, w.M, w.N, w.D, 0, "ragged")
    host = dict(img=img[lo:hi].to(dtype), txt=txt[lo * w.T:hi * w.T].to(dtype), etxt=etxt[lo:hi].to(dtype),
                obj=obj[lo:hi].to(dtype), tnum=tnum[lo:hi], onum=onum[lo:hi])
    host = {k: v.contiguous().pin_memory() for k, v in host.items()}
    return host, ls


def run_reference(args, w):
    """--impl reference: the reference algorithm (oracle port, PyTorch CPU) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import clip_event_oracle as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    img, txt, ls = syn.contrastive_inputs(w.B, w.T, w.D, 0, "trained")
    etxt, obj, tnum, onum = syn.ot_inputs(w.B, w.M, w.N, w.D, 0, "ragged")
    # bounded sample: a row block sized so that the whole run (warm-up + steps) is ~2 minutes of
    # host time; the same block selection as the cpu_baseline leg of the B200 arm
    rows = min(w.B, 64)
    t0 = time.perf_counter()
    orc.loss_head_rowblock_step(img, txt, ls, w.T, (0, rows), etxt, obj, tnum, onum)
    first = time.perf_counter() - t0
    per_step_budget = min(4.0, 120.0 / max(args.steps + args.warmup, 1))
    rows = int(max(16, min(w.B, rows * per_step_budget / max(first, 1e-3))))

    def step():
        orc.loss_head_rowblock_step(img, txt, ls, w.T, (0, rows), etxt, obj, tnum, onum)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = rows / dt
    sample = "%d-image row block of the %d-image batch: scored against all %d descriptions / %d images, OT on the block" % (
        rows, w.B, w.B * w.T, w.B)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s: %s" % (w.name, WORKLOAD_TEXT[w.name]), "global_batch": w.B},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOAD_TEXT))
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()
    w = syn.WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
        return

    import torch.distributed as dist
    import clip_event_b200 as ce
    from clip_event_b200 import _lib as L
    from clip_event_b200 import distributed as cd
    from clip_event_b200 import functional as F_

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200 (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.gpus != world and rank == 0:
        print("warning: --gpus %d but WORLD_SIZE %d" % (args.gpus, world), file=sys.stderr)
    lib = L.load()
    L.check(lib.ce_device_check(), "device check")
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    esz = 2 if args.dtype == "bf16" else 4
    peaks = load_peaks()

    lo, hi = cd.shard_bounds(w.B, world, rank)
    b = hi - lo
    host, ls_init = make_inputs(w, dtype, lo, hi)
    static = {k: v.to(dev) for k, v in host.items()}
    if world > 1:
        lpi, lpt, idx = cd.global_labels_for_rank(b, w.T, world, rank, device=dev)
    else:
        lpi, lpt, idx = (t.to(dev) for t in syn.contrastive_labels(w.B, w.T))
    head = ce.ClipEventHead().to(dev)
    crit, crit_ot = ce.CriterionContrastive("ce"), ce.CriterionAlignment()
    leaves = {k: static[k].requires_grad_(True) for k in ("img", "txt", "etxt", "obj")}
    losses_out = torch.zeros(3, dtype=torch.float32, device=dev)
    losses_host = torch.zeros(3, dtype=torch.float32).pin_memory()

    ot_stream = torch.cuda.Stream()

    def step(static=static, leaves=leaves, losses_out=losses_out):
        """fwd + bwd of the loss head as engine.py:48-67,88 drives it.  The two criteria are
        independent until the final sum, so the OT criterion is issued on a second stream: its
        ALU/HBM-bound kernels overlap the tensor-core GEMMs and the NCCL latencies (autograd runs
        each backward on the stream of its forward)."""
        for t in leaves.values():
            t.grad = None
        head.logit_scale.grad = None
        main = torch.cuda.current_stream()
        ot_stream.wait_stream(main)
        with torch.cuda.stream(ot_stream):
            if world > 1:
                loss_ot = cd.sharded_alignment(leaves["etxt"], leaves["obj"], static["tnum"], static["onum"])
            else:
                loss_ot = crit_ot(leaves["etxt"], leaves["obj"], static["tnum"], static["onum"])["loss_ot"]
        if world > 1:
            li, lt = cd.global_contrastive(leaves["img"], leaves["txt"], head.logit_scale, None, lpt, idx)  # canonical labels
            loss_dict = {"loss_i": li, "loss_t": lt}
        else:
            a, b_ = head(leaves["img"], leaves["txt"])
            loss_dict = crit(a, b_, lpi, lpt, index_pos=idx, constrastive_overbatch=True)
        main.wait_stream(ot_stream)
        loss_dict["loss_ot"] = loss_ot
        total = sum(loss for loss in loss_dict.values())     # engine.py:67, in the losses' own dtype
        total.backward()
        main.wait_stream(ot_stream)
        losses_out.copy_(torch.stack([loss_dict["loss_i"], loss_dict["loss_t"], loss_dict["loss_ot"]]))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (eager), count launches ----------------------------------------------------
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
        n0 = lib.ce_debug_launch_count()
        step()
        launches_per_step = int(lib.ce_debug_launch_count() - n0)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()

    graph = None
    if not args.no_graph:   # the NCCL exchange is captured too (thread-local capture: the NCCL watchdog thread stays legal)
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                step()
            torch.cuda.synchronize()
        except Exception as e:  # pragma: no cover
            graph = None
            torch.cuda.synchronize()
            if rank == 0:
                print("note: CUDA graph capture failed (%s); timing eager launches" % str(e)[:200], file=sys.stderr)
    run = graph.replay if graph is not None else step

    # inputs (per rank): img+txt+OT nodes; > L2 (126 MB) for c3/c4, otherwise flush L2 between steps
    in_bytes = sum(static[k].numel() * static[k].element_size() for k in ("img", "txt", "etxt", "obj"))
    flush = None
    if in_bytes < 256 * 1024 * 1024:
        flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def timed_loop(fn, steps, warmup, per_step_host=None):
        for _ in range(warmup):
            if flush is not None:
                flush.zero_()
            fn()
        barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for s, e in evs:
            if flush is not None:
                flush.zero_()
            s.record()
            fn()
            e.record()
            if per_step_host is not None:
                per_step_host()
        barrier()
        ms = sum(s.elapsed_time(e) for s, e in evs) / steps
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing ("value") ---------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_step = timed_loop(run, args.steps, max(args.warmup, 3))
    # K steps last ~20 ms -- shorter than one nvidia-smi sampling period -- so the same step keeps
    # replaying for 0.6 s more (not timed) while the sampler runs: the clocks line then describes the
    # GPU under exactly this load
    for _ in range(int(min(5000, max(20, 600.0 / ms_step)))):   # same count on every rank (ms_step is the max over ranks)
        run()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "timed steps + 0.6 s of the same step replayed back to back"
    value = w.B / (ms_step * 1e-3)

    # ---- e2e: host buffers -> public API -> host losses -------------------------------------
    # Every step copies ITS inputs from pinned host memory and returns its three losses to the host.
    # The loop is the one a training loop with a prefetching loader runs: two device input sets, the
    # copy of step i+1 (copy stream) overlapping the kernels of step i; the timed region is the wall
    # clock from the first copy to the last loss landing on the host.
    sets = [(static, leaves, losses_out, run)]
    if graph is not None:
        try:
            # the second set's warm-up runs on a side stream after the first set's graph exists
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
            static_b = {k: v.detach().clone() for k, v in static.items()}
            leaves_b = {k: static_b[k].requires_grad_(True) for k in ("img", "txt", "etxt", "obj")}
            losses_b = torch.zeros(3, dtype=torch.float32, device=dev)
            step_b = lambda: step(static_b, leaves_b, losses_b)   # noqa: E731
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step_b()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph_b = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph_b, capture_error_mode="thread_local"):
                step_b()
            torch.cuda.synchronize()
            sets.append((static_b, leaves_b, losses_b, graph_b.replay))
        except Exception as e:  # pragma: no cover
            torch.cuda.synchronize()
            if rank == 0:
                print("note: second input set not captured (%s); e2e runs unpipelined" % str(e)[:200], file=sys.stderr)
    n_e2e = 3 + min(args.steps, 10)
    losses_e2e = torch.zeros(n_e2e, 3, dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream()
    main_stream = torch.cuda.current_stream()
    copied = [torch.cuda.Event() for _ in sets]
    consumed = [torch.cuda.Event() for _ in sets]

    def e2e_loop(first, last):
        for i in range(first, last):
            st_i, _, lo_i, run_i = sets[i % len(sets)]
            with torch.cuda.stream(copy_stream), torch.no_grad():
                copy_stream.wait_event(consumed[i % len(sets)])      # the step that last read this set
                for k in ("img", "txt", "etxt", "obj", "tnum", "onum"):
                    st_i[k].copy_(host[k], non_blocking=True)
                copied[i % len(sets)].record(copy_stream)
            main_stream.wait_event(copied[i % len(sets)])
            run_i()
            losses_e2e[i].copy_(lo_i, non_blocking=True)
            consumed[i % len(sets)].record(main_stream)
        main_stream.synchronize()

    for ev in consumed:
        ev.record(main_stream)
    e2e_loop(0, 3)                                   # warm-up
    barrier()
    t0 = time.perf_counter()
    e2e_loop(3, n_e2e)
    e2e_s = torch.tensor([(time.perf_counter() - t0) / (n_e2e - 3)], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    losses_host.copy_(losses_e2e[n_e2e - 1])
    h2d = sum(host[k].numel() * host[k].element_size() for k in host)
    e2e = {"value": w.B / float(e2e_s.item()), "unit": "samples/s", "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": 12, "ms_per_step": float(e2e_s.item()) * 1e3,
           "input_sets": len(sets), "steps": n_e2e - 3}

    # ---- per-chain timing for the rooflines (eager, events around each C-ABI chain) -----------
    def chain_contrastive():
        for t in (leaves["img"], leaves["txt"]):
            t.grad = None
        head.logit_scale.grad = None
        if world > 1:
            li, lt = cd.global_contrastive(leaves["img"], leaves["txt"], head.logit_scale, None, lpt, idx)  # canonical labels
        else:
            li, lt = F_.contrastive_over_batch(leaves["img"], leaves["txt"], head.logit_scale, lpi, lpt, idx)
        (li + lt).backward()

    def chain_ot():
        for t in (leaves["etxt"], leaves["obj"]):
            t.grad = None
        loss, _ = F_.ot_alignment(leaves["etxt"], leaves["obj"], static["tnum"], static["onum"])
        loss.backward()

    def graphed(fn):
        """Replayable CUDA graph of one chain (eager launches on a slow host would time the host)."""
        if args.no_graph:
            return fn
        try:
            s_ = torch.cuda.Stream()
            s_.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s_):
                fn()
            torch.cuda.current_stream().wait_stream(s_)
            torch.cuda.synchronize()
            g_ = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_, capture_error_mode="thread_local"):
                fn()
            torch.cuda.synchronize()
            return g_.replay
        except Exception:
            torch.cuda.synchronize()
            return fn

    n_seg = max(5, min(args.steps, 20))
    ms_con = timed_loop(graphed(chain_contrastive), n_seg, 3)
    ms_ot = timed_loop(graphed(chain_ot), n_seg, 3)

    flops = algorithmic_work(w, esz, w.B, b)           # per rank: all B rows x local columns
    ot_bytes = 2.0 * (w.M + w.N) * w.D * esz * b
    tensor_peak = peaks["bf16_tflops"] * (1.0 if args.dtype == "bf16" else 0.5)
    roof_gemm = {"kernel": "umma_gemm_kernel chain (ce_contrastive_fwd + ce_contrastive_bwd)", "bound": "tensor",
                 "achieved": flops / (ms_con * 1e-3) / 1e12, "peak": tensor_peak, "unit": "TFLOP/s",
                 "frac": flops / (ms_con * 1e-3) / 1e12 / tensor_peak, "traffic": None, "ms": ms_con,
                 "peak_source": peaks["source"] + (" bf16 burst" if args.dtype == "bf16" else " bf16 burst / 2 (tf32; 3 products per flop in fp32 mode)")}
    roof_ot = {"kernel": "ot_cost_kernel + ot_ipot_kernel + ot_grad_kernel (ce_ot_fwd_bwd)", "bound": "hbm",
               "achieved": ot_bytes / (ms_ot * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
               "frac": ot_bytes / (ms_ot * 1e-3) / 1e9 / peaks["hbm_gbs"], "traffic": None, "ms": ms_ot,
               "peak_source": peaks["source"] + " copy bandwidth"}
    # DRAM traffic per step of each chain, from the committed ncu --set full captures of the same
    # workload (profiles/r01_traffic_*.json; null when there is no capture for this configuration)
    tpath = os.path.join(ROOT, "profiles", "r01_traffic_%s_%s.json" % (w.name, args.dtype))
    if world == 1 and os.path.exists(tpath):
        with open(tpath) as f:
            tr = json.load(f)
        roof_gemm["traffic"] = tr.get("gemm_chain_bytes_per_step")
        roof_ot["traffic"] = tr.get("ot_chain_bytes_per_step")
        roof_gemm["traffic_source"] = roof_ot["traffic_source"] = "profiles/" + os.path.basename(tpath)
    dominant, secondary = (roof_gemm, roof_ot) if ms_con >= ms_ot else (roof_ot, roof_gemm)

    # ---- CPU baseline (rank 0, N = 1 only) ---------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import clip_event_oracle as orc
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        img32, txt32 = host["img"].float(), host["txt"].float()
        e32, o32 = host["etxt"].float(), host["obj"].float()
        rows = min(w.B, 64)
        t0 = time.perf_counter()
        orc.loss_head_rowblock_step(img32, txt32, ls_init, w.T, (0, rows), e32, o32, host["tnum"], host["onum"])
        first = time.perf_counter() - t0
        rows = int(max(16, min(w.B, rows * 4.0 / max(first, 1e-3))))      # ~4 s per timed step
        times = []
        for i in range(4):
            t0 = time.perf_counter()
            orc.loss_head_rowblock_step(img32, txt32, ls_init, w.T, (0, rows), e32, o32, host["tnum"], host["onum"])
            times.append(time.perf_counter() - t0)
        dt = statistics.median(times[1:])
        cpu = {"value": rows / dt, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": "%d-image row block of the %d-image batch (scored against all %d descriptions / %d images, OT on "
                         "the block), oracle = reference algorithm in PyTorch CPU fp32, median of 3 after 1 warm-up"
                         % (rows, w.B, w.B * w.T, w.B)}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16" if args.dtype == "bf16" else "f32 (3xTF32 tensor-core products)",
            "data": "synthetic",
            "config": {"workload": "%s: %s" % (w.name, WORKLOAD_TEXT[w.name]), "global_batch": w.B,
                       "per_rank_batch": b, "descriptions_per_image": w.T, "embed_dim": w.D,
                       "ot_nodes": "%dx%d" % (w.M, w.N), "ipot_iters": 50,
                       "parallelism": "column-sharded global negatives over %d rank(s)" % world,
                       "l2": "inputs %.0f MB per rank %s" % (in_bytes / 1e6, "(> 126 MB L2)" if flush is None else
                                                             "; 512 MB L2 flush write between steps"),
                       "cuda_graph": graph is not None},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
            "gpu_launches_per_step": launches_per_step,
            "roofline": dominant, "roofline_secondary": secondary,
            "losses": [float(x) for x in losses_host.tolist()],
        }
        if cpu is not None:
            out["cpu_baseline"] = cpu
        print(json.dumps(out), flush=True)
    if world > 1:
        # destroy_process_group() blocks after NCCL work was replayed from CUDA graphs; everything is
        # already synchronised and printed, so leave without the teardown.
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
