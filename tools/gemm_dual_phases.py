"""Phase stamps of the grouped dgrad + wgrad launches (fc2: DGELU epilogue + column sums; fc1: plain) -- where does the time go?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mtp_b200 import ops, _lib as L

T, C = 1568, 1024
dbg = torch.zeros(148 * 8, dtype=torch.int64, device="cuda")
for name, n_out, n_in, mode, variants in [("fc2 dgelu+colsum", C, 4 * C, "dgelu", ("full", "no_colsum", "plain_epilogue")), ("fc1", 4 * C, C, "bf16", ("full",))]:
    g = torch.randn(T, n_out, device="cuda").to(torch.bfloat16)
    x = torch.randn(T, n_in, device="cuda").to(torch.bfloat16)
    w = torch.randn(n_out, n_in, device="cuda").to(torch.bfloat16)
    dW = torch.empty(n_out, n_in, device="cuda")
    dx = torch.empty(T, n_in, device="cuda", dtype=torch.bfloat16)
    aux = torch.randn(T, n_in, device="cuda").to(torch.bfloat16)
    cs = torch.zeros(n_in, device="cuda")
    ss = torch.zeros(1, device="cuda")
    for var in variants:
        dg = mode == "dgelu" and var != "plain_epilogue"
        d0 = dict(A=g, B=w, M=T, N=n_in, K=n_out, out=dx, b_mn=True, mode=L.EPI_BF16_DGELU if dg else L.EPI_BF16, aux=aux if dg else None,
                  lda=n_out, ldb=n_in, colsum=cs if var == "full" else None, b_static=True)
        d1 = dict(A=g, B=x, M=n_out, N=n_in, K=T, out=dW, a_mn=True, b_mn=True, mode=L.EPI_F32, lda=n_out, ldb=n_in, ldo=n_in, b_static=True,
                  sumsq=ss if var == "full" else None)
        for bn in (256, 1256, 192):
            for _ in range(3):
                ops.gemm_dual(d0, d1, force_bn=bn)
            torch.cuda.synchronize()
            n = 20
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                for _ in range(n):
                    ops.gemm_dual(d0, d1, force_bn=bn)
            gr.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / n
            L.call("mtp_gemm_set_debug", dbg.data_ptr())
            dbg.zero_()
            ops.gemm_dual(d0, d1, force_bn=bn)
            torch.cuda.synchronize()
            L.call("mtp_gemm_set_debug", 0)
            d = dbg.view(148, 8).cpu()
            d = d[d[:, 0] > 0]
            t0 = d[:, 0].min()
            rel = (d - t0).float() / 1e3
            names = ["start", "prologue", "first operands", "item0 MMAs issued", "last MMAs issued", "last acc complete", "CTA done"]
            print(f"== {name} [{var}] bn={bn}: {us:.1f} us per launch in a chain; isolated span {float(rel[:, 6].max()):.1f} us, {len(d)} CTAs")
            print("   " + "; ".join(f"{nm} {float(rel[:, i].median()):.1f}/{float(rel[:, i].max()):.1f}" for i, nm in enumerate(names)) + "  (median/max us)")
