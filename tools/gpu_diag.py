"""One-shot GPU diagnostics: runs each check group in its own subprocess (a device trap in one
group must not take the others down) and prints compact error summaries.
Usage (on a B200):  python tools/gpu_diag.py [group ...]   -> also tee'd to gpurun_out/diag.log
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

GROUPS = ["gemm_bf16_kk", "gemm_bf16_mn", "gemm_pair", "gemm_f32", "ot_blocks", "ot_fused", "contrastive", "similarity"]


def rel(a, b):
    import torch
    a, b = a.double().flatten(), b.double().flatten()
    d = b.norm().item()
    return (a - b).norm().item() / (d if d > 0 else 1.0)


def run_gemm(dtype_name, cases, pair=False):
    import torch
    from clip_event_b200 import _lib as L
    lib = L.load()
    dt = torch.bfloat16 if dtype_name == "bf16" else torch.float32
    code = L.dtype_code(dt)
    for (M, N, K, amn, bmn, sk) in cases:
        g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
        A = torch.randint(-3, 4, (M, K), generator=g).float()
        B = torch.randint(-3, 4, (N, K), generator=g).float()
        ref = A @ B.t()
        if dtype_name == "f32":
            ref = ref * 3
        Ad = (A.t().contiguous() if amn else A).to(dt).cuda()
        Bd = (B.t().contiguous() if bmn else B).to(dt).cuda()
        Cd = torch.full((M, N), float("nan"), device="cuda")
        fn = lib.ce_debug_gemm_pair if pair else lib.ce_debug_gemm
        rc = fn(Ad.data_ptr(), Bd.data_ptr(), Cd.data_ptr(), M, N, K, code, amn, bmn, sk,
                               torch.cuda.current_stream().cuda_stream)
        msg = "" if rc == 0 else L.last_error()
        torch.cuda.synchronize()
        out = Cd.cpu()
        bad = (out != ref) | out.isnan()
        nbad = int(bad.sum())
        line = ("pair " if pair else "") + "gemm %s M=%d N=%d K=%d a_mn=%d b_mn=%d sk=%d rc=%d bad=%d/%d maxerr=%.3g %s" % (
            dtype_name, M, N, K, amn, bmn, sk, rc, nbad, M * N, float((out - ref).abs().nan_to_num(1e9).max()), msg)
        print(line, flush=True)
        if nbad:
            idx = bad.nonzero()[:6].tolist()
            print("   first bad (row,col):", idx, "got", [float(out[i, j]) for i, j in idx], "want",
                  [float(ref[i, j]) for i, j in idx], flush=True)
            rows_bad = bad.any(1).nonzero().flatten().tolist()
            cols_bad = bad.any(0).nonzero().flatten().tolist()
            print("   bad rows: n=%d first=%s ; bad cols: n=%d first=%s" % (len(rows_bad), rows_bad[:12], len(cols_bad), cols_bad[:12]), flush=True)


def group_gemm_bf16_kk():
    run_gemm("bf16", [(128, 256, 64, 0, 0, 1), (128, 256, 128, 0, 0, 1), (128, 256, 512, 0, 0, 1),
                      (256, 512, 512, 0, 0, 1), (100, 200, 72, 0, 0, 1), (1024, 2304, 512, 0, 0, 1),
                      (4096, 4608, 512, 0, 0, 1), (256, 256, 2048, 0, 0, 4)])


def group_gemm_bf16_mn():
    run_gemm("bf16", [(128, 256, 64, 1, 0, 1), (128, 256, 64, 0, 1, 1), (128, 256, 64, 1, 1, 1),
                      (256, 512, 256, 1, 1, 1), (520, 512, 1000, 0, 1, 1), (2304, 512, 256, 1, 1, 1),
                      (2304, 512, 1024, 1, 1, 3)])


def group_gemm_pair():
    run_gemm("bf16", [(256, 256, 64, 0, 0, 1), (256, 256, 512, 0, 0, 1), (128, 256, 128, 0, 0, 1),
                      (512, 512, 512, 0, 0, 1), (100, 200, 72, 0, 0, 1), (1024, 2304, 512, 0, 0, 1),
                      (4096, 4608, 512, 0, 0, 1), (256, 256, 2048, 0, 0, 4),
                      (256, 256, 64, 1, 0, 1), (256, 256, 64, 0, 1, 1), (256, 256, 64, 1, 1, 1),
                      (520, 512, 1000, 0, 1, 1), (2304, 512, 256, 1, 1, 1), (2304, 512, 1024, 1, 1, 3),
                      (4608, 512, 4096, 1, 1, 1), (4096, 512, 4608, 0, 1, 1)], pair=True)


def group_gemm_f32():
    run_gemm("f32", [(128, 128, 32, 0, 0, 1), (128, 128, 256, 0, 0, 1), (256, 384, 512, 0, 0, 1),
                     (128, 128, 32, 1, 0, 1), (128, 128, 32, 0, 1, 1), (256, 256, 128, 1, 1, 1),
                     (100, 72, 40, 0, 0, 1), (1000, 512, 520, 1, 1, 2)])


def group_ot_blocks():
    import torch
    import clip_event_b200 as ce
    from clip_event_b200 import synthetic as syn
    from oracle import clip_event_oracle as orc
    for (B, M, N, D, masks) in [(4, 4, 7, 16, "full"), (6, 5, 9, 24, "edge"), (8, 16, 50, 512, "ragged")]:
        txt, obj, tnum, onum = syn.ot_inputs(B, M, N, D, 0, masks)
        img = obj[:, 1:].contiguous()
        tp, ip = tnum == 0, onum[:, 1:] == 0
        cost_ref = orc.cost_matrix_cosine(txt, img)
        cost = ce.cost_matrix_cosine(txt.cuda(), img.cuda()).cpu()
        jp = tp.unsqueeze(-1) | ip.unsqueeze(-2)
        cm = cost_ref.masked_fill(jp, 0)
        tl = (M - tp.sum(1)).float()
        il = (N - ip.sum(1)).float()
        plan_ref = orc.ipot(cm, tl, tp, il, ip, jp, 0.5, 50, 1)
        plan = ce.ipot(cm.cuda(), tl.cuda(), tp.cuda(), il.cuda(), ip.cuda(), jp.cuda(), 0.5, 50, 1).cpu()
        tr = ce.trace(cm.matmul(plan_ref).cuda()).cpu()
        tr_ref = orc.trace_batched(cm.matmul(plan_ref))
        print("ot_blocks B%d %dx%d D%d %s: cost %.2e plan %.2e trace %.2e" % (
            B, M, N, D, masks, rel(cost, cost_ref), rel(plan, plan_ref), rel(tr, tr_ref)), flush=True)


def group_ot_fused():
    import torch
    import clip_event_b200 as ce
    from clip_event_b200 import synthetic as syn
    from oracle import clip_event_oracle as orc
    cases = [(4, 4, 7, 16, "full", "iid"), (6, 5, 9, 24, "edge", "iid"), (5, 6, 11, 16, "scattered", "iid"),
             (32, 8, 50, 512, "ragged", "iid"), (16, 16, 50, 512, "full", "correlated"),
             (4, 32, 257, 768, "ragged", "iid"), (2, 64, 577, 768, "full", "iid"), (3, 20, 300, 64, "ragged", "iid"),
             (3, 64, 130, 64, "ragged", "iid"), (3, 32, 577, 64, "ragged", "iid")]
    for dt in (torch.float32, torch.bfloat16):
        for (B, M, N, D, masks, kind) in cases:
            txt, obj, tnum, onum = syn.ot_inputs(B, M, N, D, 1, masks, kind, dtype=dt)
            t64, o64 = txt.double(), obj.double()
            tp, ip = tnum == 0, onum[:, 1:] == 0
            d_ref, dx_ref, dy_ref = orc.ot_closed_form_grads(t64, o64[:, 1:], tp, ip, torch.full((B,), 0.01, dtype=torch.float64))
            tg, og = txt.cuda().requires_grad_(True), obj.cuda().requires_grad_(True)
            try:
                out = ce.CriterionAlignment()(tg, og, tnum.cuda(), onum.cuda())
                out["loss_ot"].backward()
                torch.cuda.synchronize()
                print("ot_fused %s B%d %dx%d D%d %s/%s: loss %.3e (ref %.6g got %.6g) dtxt %.2e dobj %.2e slot0 %.1e" % (
                    str(dt)[6:], B, M, N, D, masks, kind,
                    abs(out["loss_ot"].item() - 0.01 * d_ref.sum().item()) / max(abs(0.01 * d_ref.sum().item()), 1e-30),
                    0.01 * d_ref.sum().item(), out["loss_ot"].item(),
                    rel(tg.grad.cpu(), dx_ref), rel(og.grad.cpu()[:, 1:], dy_ref), og.grad[:, 0].abs().max().item()), flush=True)
            except Exception as e:  # noqa
                print("ot_fused %s B%d %dx%d D%d %s: EXC %s" % (str(dt)[6:], B, M, N, D, masks, e), flush=True)


def group_contrastive():
    import torch
    import clip_event_b200 as ce
    from clip_event_b200 import synthetic as syn
    from oracle import clip_event_oracle as orc
    cases = [(6, 3, 32, "iid"), (8, 4, 48, "trained"), (32, 5, 512, "iid"), (32, 5, 512, "trained"),
             (256, 9, 512, "trained"), (300, 7, 768, "iid"), (1024, 9, 768, "trained")]
    for dt in (torch.bfloat16, torch.float32):
        for (B, T, D, kind) in cases:
            img, txt, ls = syn.contrastive_inputs(B, T, D, 2, kind, dtype=dt)
            lpi, lpt, idx = syn.contrastive_labels(B, T)
            li, lt, dimg, dtxt, dls = orc.contrastive_closed_form(img.double(), txt.double(), ls.double(), lpi, lpt, idx)
            head = ce.ClipEventHead().cuda()
            ig, tg = img.cuda().requires_grad_(True), txt.cuda().requires_grad_(True)
            try:
                a, b = head(ig, tg)
                out = ce.CriterionContrastive("ce")(a, b, lpi.cuda(), lpt.cuda(), index_pos=idx.cuda())
                (out["loss_i"] + out["loss_t"]).backward()
                torch.cuda.synchronize()
                print("contrastive %s B%d T%d D%d %s: loss_i %.6g/%.6g loss_t %.6g/%.6g dimg %.2e dtxt %.2e dls %.6g/%.6g" % (
                    str(dt)[6:], B, T, D, kind, out["loss_i"].item(), li.item(), out["loss_t"].item(), lt.item(),
                    rel(ig.grad.cpu(), dimg), rel(tg.grad.cpu(), dtxt), head.logit_scale.grad.item(), dls.item()), flush=True)
            except Exception as e:  # noqa
                print("contrastive %s B%d T%d D%d: EXC %s" % (str(dt)[6:], B, T, D, e), flush=True)


def group_similarity():
    import torch
    import clip_event_b200 as ce
    from clip_event_b200 import synthetic as syn
    from oracle import clip_event_oracle as orc
    for dt in (torch.bfloat16, torch.float32):
        for (B, T, D) in [(6, 3, 32), (32, 5, 512), (200, 9, 512)]:
            img, txt, ls = syn.contrastive_inputs(B, T, D, 3, "trained", dtype=dt)
            a_ref, b_ref = orc.similarity_logits(img.double(), txt.double(), ls.double())
            head = ce.ClipEventHead().cuda()
            a, b = head(img.cuda(), txt.cuda())
            print("similarity %s B%d T%d D%d: per_image %.2e per_text %.2e softmax %.2e" % (
                str(dt)[6:], B, T, D, rel(a.materialize().cpu(), a_ref), rel(b.materialize().cpu(), b_ref),
                rel(a.softmax(dim=-1).cpu(), a_ref.softmax(-1))), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        globals()["group_" + sys.argv[2]]()
        sys.exit(0)
    groups = sys.argv[1:] or GROUPS
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "diag.log"), "a")
    for g in groups:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", g], capture_output=True,
                               text=True, timeout=240, cwd=ROOT)
            text = r.stdout + ("\n[stderr tail]\n" + r.stderr[-3000:] if r.returncode != 0 else "")
            head = "== %s rc=%d %.1fs" % (g, r.returncode, time.time() - t0)
        except subprocess.TimeoutExpired as e:
            text = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
            head = "== %s TIMEOUT" % g
        print(head)
        print(text, flush=True)
        log.write(head + "\n" + text + "\n")
        log.flush()
