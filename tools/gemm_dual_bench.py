"""Microbenchmark of the grouped dgrad + wgrad launch of each Linear of a ViT-L block (graph replay), per tile variant."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mtp_b200 import ops, _lib as L

T, C = 1568, 1024
layers = [("fc2", C, 4 * C, "dgelu"), ("fc1", 4 * C, C, "bf16"), ("proj", C, C, "bf16"), ("qkv", 3 * C, C, "bf16")]   # name, n_out, n_in
res = []
for name, n_out, n_in, mode in layers:
    g = torch.randn(T, n_out, device="cuda").to(torch.bfloat16)
    x = torch.randn(T, n_in, device="cuda").to(torch.bfloat16)
    w = torch.randn(n_out, n_in, device="cuda").to(torch.bfloat16)
    dW = torch.empty(n_out, n_in, device="cuda")
    dx = torch.empty(T, n_in, device="cuda", dtype=torch.bfloat16)
    aux = torch.randn(T, n_in, device="cuda").to(torch.bfloat16) if mode == "dgelu" else None
    cs = torch.zeros(n_in, device="cuda")
    d0 = dict(A=g, B=w, M=T, N=n_in, K=n_out, out=dx, b_mn=True, mode=L.EPI_BF16_DGELU if mode == "dgelu" else L.EPI_BF16, aux=aux,
              lda=n_out, ldb=n_in, colsum=cs if mode == "dgelu" else None)
    d1 = dict(A=g, B=x, M=n_out, N=n_in, K=T, out=dW, a_mn=True, b_mn=True, mode=L.EPI_F32, lda=n_out, ldb=n_in, ldo=n_in)
    row = {"name": name}
    for bn in (0, 64, 128, 192, 256, 1128, 1256):
        try:
            for _ in range(3):
                ops.gemm_dual(d0, d1, force_bn=bn)
            torch.cuda.synchronize()
            n = 20
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                for _ in range(n):
                    ops.gemm_dual(d0, d1, force_bn=bn)
            gr.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); gr.replay(); e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / n
            row[f"bn{bn}"] = (round(us, 1), round(4.0 * T * n_in * n_out / us / 1e6, 0))
            if bn == 0:
                row["picked"] = L.load().mtp_gemm_last_config()
        except Exception as ex:
            row[f"bn{bn}"] = str(ex)[:40]
    res.append(row)
    print(row, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/gemm_dual_bench.json", "w"), indent=1)
