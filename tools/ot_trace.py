"""Role timeline of the fused OT kernel (tuning aid): clock64 stamps of CTA 0's IO / MMA / IPOT roles
for its first samples.   python tools/ot_trace.py [workload] [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
buf = torch.zeros(64 * 32, dtype=torch.int64, device="cuda")
os.environ["CE_OT_TRACE_PTR"] = hex(buf.data_ptr())
from clip_event_b200 import functional as F_, synthetic as syn
w = syn.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
B = int(sys.argv[2]) if len(sys.argv) > 2 else w.B
etxt, obj, tnum, onum = syn.ot_inputs(B, w.M, w.N, w.D, 0, "ragged", dtype=torch.bfloat16)
eg, og = etxt.cuda().requires_grad_(True), obj.cuda().requires_grad_(True)
tnum, onum = tnum.cuda(), onum.cuda()
nograd = len(sys.argv) > 3 and sys.argv[3] == "nograd"
for _ in range(2):
    buf.zero_()
    if nograd:
        with torch.no_grad():
            l, _ = F_.ot_alignment(eg.detach(), og.detach(), tnum, onum)
    else:
        l, _ = F_.ot_alignment(eg, og, tnum, onum)
    torch.cuda.synchronize()
t = buf.cpu().view(64, 32)
base = int(t[0, 0])
names = ["ld_issue", "out_rdy", "st_read", "full", "cost_dn", "w_rdy", "dx_dn", "dy_dn", "grad_dn", "s_rdy", "A_dn", "iter_dn", "epi_dn"]
if os.environ.get("CE_OT_STREAM", "1") != "0":   # csrc/ot_stream.cu events
    names = ["cld_iss", "gld_iss", "c_xfull", "c_yfull", "c_done", "c_scrfr", "g_wrdy", "g_dxdn", "st_xout", "s_rdy", "A_dn", "iter_dn", "epi_dn"]
print("cycles relative to the first load issue (CTA 0)")
print("  k " + " ".join("%8s" % n for n in names))
for k in range(64):
    if int(t[k, 9]) == 0 or k > int(os.environ.get('TRACE_ROWS', '7')):
        break
    print("%3d " % k + " ".join("%8d" % (int(t[k, e]) - base if int(t[k, e]) else -1) for e in range(13)))
    if os.environ.get("TRACE_GRAD"):
        print("      grad: wrdy %d copy_done %d xfull %d | " % tuple(int(t[k, e]) - base for e in (6, 25, 26)) +
              " | ".join("c%d wait %d full %d done %d" % (c, int(t[k, 13 + c]) - base, int(t[k, 17 + c]) - base, int(t[k, 21 + c]) - base) for c in range(4)) +
              " | dx_done %d" % (int(t[k, 7]) - base))
