"""Microbenchmark of the tcgen05 GEMM on the model's shapes (run on the GPU box): time per launch and TFLOP/s per tile width."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mtp_b200 import ops, _lib as L

T, C = 1568, 1024
shapes = [  # name, M, N, K, a_mn, b_mn, mode
    ("qkv fwd", T, 3 * C, C, False, False, "bf16"), ("proj fwd", T, C, C, False, False, "resid"),
    ("fc1 fwd", T, 4 * C, C, False, False, "gelu"), ("fc2 fwd", T, C, 4 * C, False, False, "resid"),
    ("fc2 dgrad", T, 4 * C, C, False, True, "dgelu"), ("fc1 dgrad", T, C, 4 * C, False, True, "bf16"),
    ("qkv dgrad", T, C, 3 * C, False, True, "bf16"), ("proj dgrad", T, C, C, False, True, "bf16"),
    ("fc2 wgrad", C, 4 * C, T, True, True, "f32"), ("fc1 wgrad", 4 * C, C, T, True, True, "f32"),
    ("qkv wgrad", 3 * C, C, T, True, True, "f32"), ("proj wgrad", C, C, T, True, True, "f32"),
    ("fpn1.3 fwd", 4 * T, 4 * C, C, False, False, "bf16"), ("big square", 8192, 8192, 8192, False, False, "bf16"),
]
only = sys.argv[1].split(",") if len(sys.argv) > 1 else None
res = []
for name, M, N, K, a_mn, b_mn, mode in shapes:
    if only and name not in only:
        continue
    A = torch.randn((K, M) if a_mn else (M, K), device="cuda").to(torch.bfloat16)
    B = torch.randn((K, N) if b_mn else (N, K), device="cuda").to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    kw = {}
    if mode in ("bf16", "gelu", "dgelu"):
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        kw["mode"] = {"bf16": L.EPI_BF16, "gelu": L.EPI_BF16_GELU, "dgelu": L.EPI_BF16_DGELU}[mode]
        if mode == "gelu":
            kw["out2"] = torch.empty_like(out)
        if mode == "dgelu":
            kw["aux"] = torch.randn(M, N, device="cuda").to(torch.bfloat16)
        if mode != "dgelu":
            kw["bias"] = bias
    elif mode == "resid":
        out = torch.empty(M, N, device="cuda")
        kw.update(mode=L.EPI_F32_RESID, aux=torch.randn(M, N, device="cuda"), bias=bias)
    else:
        out = torch.empty(M, N, device="cuda")
        kw["mode"] = L.EPI_F32
    row = {"name": name, "M": M, "N": N, "K": K}
    for bn in (0, 128, 192, 256, 1128, 1192, 1256):
        try:
            for _ in range(3):
                ops.gemm(A, B, M, N, K, out, a_mn=a_mn, b_mn=b_mn, force_bn=bn, **kw)
            torch.cuda.synchronize()
            n = 20
            gr = torch.cuda.CUDAGraph()          # graph replay removes the ~10 us host cost per launch from the measurement
            with torch.cuda.graph(gr):
                for _ in range(n):
                    ops.gemm(A, B, M, N, K, out, a_mn=a_mn, b_mn=b_mn, force_bn=bn, **kw)
            gr.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            gr.replay()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / n
            row[f"bn{bn}"] = (round(us, 1), round(2.0 * M * N * K / us / 1e6, 0))
            if bn == 0:
                row["picked"] = L.load().mtp_gemm_last_config()
        except Exception as ex:
            row[f"bn{bn}"] = str(ex)[:40]
    # cuBLAS reference for context
    if not a_mn and not b_mn:
        Bt = B.t()
        for _ in range(3):
            torch.matmul(A, Bt)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(20):
                torch.matmul(A, Bt)
        gr.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gr.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 20
        row["cublas"] = (round(us, 1), round(2.0 * M * N * K / us / 1e6, 0))
    res.append(row)
    print(row, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/gemm_bench.json", "w"), indent=1)
