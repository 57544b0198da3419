"""Per-phase timeline of the GEMM kernel from in-kernel globaltimer stamps (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mtp_b200 import ops, _lib as L

dbg = torch.zeros(148 * 8, dtype=torch.int64, device="cuda")
for name, M, N, K, mode, bn in [("qkv fwd", 1568, 3072, 1024, L.EPI_BF16, 192), ("fc1 fwd", 1568, 4096, 1024, L.EPI_BF16_GELU, 192),
                                ("fc2 fwd", 1568, 1024, 4096, L.EPI_F32_RESID, 128), ("proj fwd", 1568, 1024, 1024, L.EPI_F32_RESID, 128)]:
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
    dbg.zero_()
    ops.gemm(A, B, M, N, K, out, **kw)
    torch.cuda.synchronize()
    L.call("mtp_gemm_set_debug", 0)
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
