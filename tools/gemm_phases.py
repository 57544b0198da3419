"""Per-phase timeline of the GEMM kernel from in-kernel globaltimer stamps (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mtp_b200 import ops, _lib as L

dbg = torch.zeros(148 * 8, dtype=torch.int64, device="cuda")
CASES = [("fc1 fwd gelu 256", 1568, 4096, 1024, L.EPI_BF16_GELU, 256), ("fc1 fwd gelu pair256", 1568, 4096, 1024, L.EPI_BF16_GELU, 1256),
         ("fc1 fwd gelu 192", 1568, 4096, 1024, L.EPI_BF16_GELU, 192), ("fc1 fwd plain 256", 1568, 4096, 1024, L.EPI_BF16, 256),
         ("fc1 fwd plain pair256", 1568, 4096, 1024, L.EPI_BF16, 1256),
         ("qkv fwd", 1568, 3072, 1024, L.EPI_BF16, 192), ("qkv fwd pair", 1568, 3072, 1024, L.EPI_BF16, 1192),
         ("qkv fwd 256", 1568, 3072, 1024, L.EPI_BF16, 256), ("qkv fwd pair256", 1568, 3072, 1024, L.EPI_BF16, 1256),
         ("qkv fwd 128", 1568, 3072, 1024, L.EPI_BF16, 128), ("qkv fwd 64", 1568, 3072, 1024, L.EPI_BF16, 64),
         ("1 CTA only 192", 128, 192, 1024, L.EPI_BF16, 192), ("37 CTAs 192", 37 * 128, 192, 1024, L.EPI_BF16, 192),
         ("74 CTAs 192", 74 * 128, 192, 1024, L.EPI_BF16, 192),
         ("fc2 fwd", 1568, 1024, 4096, L.EPI_F32_RESID, 128), ("fc2 fwd pair", 1568, 1024, 4096, L.EPI_F32_RESID, 1128)]
MODES = [0, 1, 2] if len(sys.argv) > 1 and sys.argv[1] == "modes" else [0]
CASES = [(n + f" [dbg{dm}]", M, N, K, mode, bn, dm) for dm in MODES for (n, M, N, K, mode, bn) in CASES]
for name, M, N, K, mode, bn, dm in CASES:
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    B = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    kw = dict(mode=mode, bias=bias, force_bn=bn)
    if mode == L.EPI_F32_RESID:
        out = torch.empty(M, N, device="cuda"); kw["aux"] = torch.randn(M, N, device="cuda")
    else:
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        if mode == L.EPI_BF16_GELU: kw["out2"] = torch.empty_like(out)
    for _ in range(3):
        ops.gemm(A, B, M, N, K, out, **kw)
    torch.cuda.synchronize()
    L.call("mtp_gemm_set_debug", dbg.data_ptr())
    L.call("mtp_gemm_set_debug_mode", dm)
    dbg.zero_()
    ops.gemm(A, B, M, N, K, out, **kw)
    torch.cuda.synchronize()
    L.call("mtp_gemm_set_debug", 0)
    L.call("mtp_gemm_set_debug_mode", 0)
    d = dbg.view(148, 8).cpu()
    live = d[:, 0] > 0
    d = d[live]
    t0 = d[:, 0].min()
    rel = (d - t0).float() / 1e3     # us
    names = ["start", "prologue done", "first operands", "item0 MMAs issued", "last MMAs issued", "last acc complete", "CTA done"]
    print(f"== {name} M={M} N={N} K={K} bn={bn}: {int(live.sum())} CTAs; kernel span {float(rel[:, 6].max()):.1f} us")
    for i, n in enumerate(names):
        col = rel[:, i]
        col = col[d[:, i] > 0]
        if len(col):
            print(f"   {n:20s} min {float(col.min()):6.1f}  median {float(col.median()):6.1f}  max {float(col.max()):6.1f} us")
