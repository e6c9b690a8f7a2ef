"""Times the single-CTA and the CTA-pair tcgen05 main loops on the loss head's GEMM shapes (bf16).
Usage (on a B200): python tools/gemm_pair_bench.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
from clip_event_b200 import _lib as L

SHAPES = [  # M, N, K, a_mn, b_mn, split_k   (c3: R 4096, C 36864, P 4096, D 512)
    (4096, 36864, 512, 0, 0, 1),     # logits recompute (EpiStats / EpiGrad shape)
    (4096, 512, 36864, 0, 1, 8),     # s * G * txt
    (36864, 512, 4096, 1, 1, 1),     # s * G^t * img
    (4096, 4096, 512, 0, 0, 1),      # positives x images
    (8192, 8192, 8192, 0, 0, 1),     # square reference point
]


def main():
    lib = L.load()
    st = torch.cuda.current_stream().cuda_stream
    for (M, N, K, amn, bmn, sk) in SHAPES:
        A = torch.randn((K, M) if amn else (M, K), device="cuda").bfloat16()
        B = torch.randn((K, N) if bmn else (N, K), device="cuda").bfloat16()
        C = torch.empty(M, N, device="cuda")
        res = {}
        for name, fn in (("single", lib.ce_debug_gemm), ("pair", lib.ce_debug_gemm_pair)):
            for _ in range(3):
                rc = fn(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, 1, amn, bmn, sk, st)
                assert rc == 0, L.last_error()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, 1, amn, bmn, sk, st)
            e1.record()
            torch.cuda.synchronize()
            res[name] = e0.elapsed_time(e1) / 10 * 1e3
            res[name + "_c"] = C.clone() if M * N <= 4096 * 4608 * 2 else None
        tf = 2.0 * M * N * K / 1e12
        line = "M=%d N=%d K=%d a_mn=%d b_mn=%d sk=%d  single %.1f us (%.0f TF/s)  pair %.1f us (%.0f TF/s)" % (
            M, N, K, amn, bmn, sk, res["single"], tf / res["single"] * 1e6, res["pair"], tf / res["pair"] * 1e6)
        if res["single_c"] is not None:
            line += "  maxdiff %.3g" % float((res["single_c"] - res["pair_c"]).abs().max())
        print(line, flush=True)


if __name__ == "__main__":
    main()
