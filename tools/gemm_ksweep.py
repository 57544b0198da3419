"""Fixed-overhead vs mainloop slope of the GEMM: sweep K for the qkv-forward shape (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mtp_b200 import ops, _lib as L

def t_gemm(M, N, K, bn, mode=L.EPI_BF16, n=30):
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    B = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16 if mode == L.EPI_BF16 else torch.float32)
    for _ in range(3):
        ops.gemm(A, B, M, N, K, out, mode=mode, force_bn=bn)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        ops.gemm(A, B, M, N, K, out, mode=mode, force_bn=bn)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n

def t_cublas(M, N, K, n=30):
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    B = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    for _ in range(3):
        torch.matmul(A, B.t())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        torch.matmul(A, B.t())
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n

for (M, N) in [(1568, 3072), (128, 192), (148 * 128, 192), (148 * 128, 384)]:
    for K in (64, 256, 1024, 4096):
        print(f"M={M} N={N} K={K}: bn192 {t_gemm(M, N, K, 192):7.1f} us  bn256 {t_gemm(M, N, K, 256):7.1f} us  cublas {t_cublas(M, N, K):7.1f} us", flush=True)
# empty-ish kernel floor for reference
x = torch.zeros(1024, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100):
    x.add_(1.0)
e1.record()
torch.cuda.synchronize()
print("tiny torch kernel back-to-back: %.1f us" % (e0.elapsed_time(e1) * 10))
