"""Back-to-back GEMM launches inside one CUDA graph: kernel spans and the idle gaps between them (globaltimer stamps)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mtp_b200 import ops, _lib as L

NL = 6
L.call("mtp_set_pdl", int(os.environ.get("MTP_PDL", "1")))
L.call("mtp_gemm_set_debug_mode", int(os.environ.get("MTP_DBG", "0")))
L.call("mtp_gemm_set_max_stages", int(os.environ.get("MTP_STAGES", "0")))
if os.environ.get("MTP_TINY"):
    x = torch.randn(4096, device="cuda"); y = torch.empty(4096, device="cuda", dtype=torch.bfloat16)
    for nl in (20, 200):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for i in range(nl):
                L.call("mtp_cast_f32_bf16", x.data_ptr(), y.data_ptr(), x.numel(), torch.cuda.current_stream().cuda_stream)
        gr.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        print(f"tiny kernel chain of {nl}: {e0.elapsed_time(e1) * 1e3 / nl:.2f} us per launch")
for name, M, N, K, bn in [("qkv fwd 192", 1568, 3072, 1024, 192), ("qkv fwd 128", 1568, 3072, 1024, 128), ("qkv fwd pair256", 1568, 3072, 1024, 1256),
                          ("proj fwd 128", 1568, 1024, 1024, 128), ("fc2 fwd 128", 1568, 1024, 4096, 128)][:int(os.environ.get("MTP_NCASES", "9"))]:
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    B = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    dbg = torch.zeros(NL, 148 * 8, dtype=torch.int64, device="cuda")
    for _ in range(3):
        ops.gemm(A, B, M, N, K, out, mode=L.EPI_BF16, force_bn=bn)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for i in range(NL):
            L.call("mtp_gemm_set_debug", dbg[i].data_ptr())
            ops.gemm(A, B, M, N, K, out, mode=L.EPI_BF16, force_bn=bn)
    L.call("mtp_gemm_set_debug", 0)
    gr.replay(); torch.cuda.synchronize()
    dbg.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record()
    torch.cuda.synchronize()
    d = dbg.view(NL, 148, 8).cpu()
    print(f"== {name}: graph of {NL} launches {e0.elapsed_time(e1) * 1e3 / NL:.1f} us per launch")
    prev_end = None
    for i in range(NL):
        di = d[i][d[i][:, 0] > 0]
        s, e = int(di[:, 0].min()), int(torch.maximum(di[:, 6], di[:, 7]).max())      # end of a CTA = its TMEM dealloc returned (stamp 7)
        smax = int(di[:, 0].max())
        msg = f"   launch {i}: span {(e - s) / 1e3:5.1f} us, CTA starts spread {(smax - s) / 1e3:4.1f} us, first CTA end {(int(torch.maximum(di[:, 6], di[:, 7]).min()) - s) / 1e3:5.1f}"
        if prev_end is not None:
            msg += f", gap after previous {(s - prev_end) / 1e3:5.1f} us; wait passed {(int(di[:, 1].min()) - prev_end) / 1e3:5.2f}..{(int(di[:, 1].max()) - prev_end) / 1e3:5.2f} us after previous end; first operands +{(int(di[:, 2].median()) - int(di[:, 1].median())) / 1e3:4.2f} us"
        msg += f"; last accumulator complete -> CTA end {float((torch.maximum(di[:, 6], di[:, 7]) - di[:, 5]).double().max()) / 1e3:5.2f} us (max), end-of-work barrier -> dealloc returned {float((di[:, 7] - di[:, 6]).double().median()) / 1e3:4.2f} us"
        if os.environ.get("MTP_DBG") == "20":
            msg += f"; end stamp -> before fence {float((di[:, 2] - di[:, 6]).double().median()) / 1e3:4.2f}, fence {float((di[:, 3] - di[:, 2]).double().median()) / 1e3:4.2f}, dealloc {float((di[:, 7] - di[:, 3]).double().median()) / 1e3:4.2f} us (medians)"
        print(msg)
        prev_end = e
