"""What makes the SM turnaround between dependent heavy launches?  Chains of a do-nothing kernel (mtp_probe_launch) with different
footprints inside one CUDA graph; reports the time per launch and the gap between one launch's last CTA finishing and the next one's
first / last CTA being ready."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mtp_b200 import _lib as L

NL = 8
st = lambda: torch.cuda.current_stream().cuda_stream
L.call("mtp_set_pdl", int(os.environ.get("MTP_PDL", "1")))
cases = [("tiny: 32 thr, 16 B smem", 148, 32, 16, 0), ("320 thr, 16 B smem", 148, 320, 16, 0), ("320 thr, 100 KB smem", 148, 320, 100 * 1024, 0),
         ("320 thr, 200 KB smem", 148, 320, 200 * 1024, 0), ("320 thr, 16 B smem, TMEM 512", 148, 320, 16, 512),
         ("320 thr, 200 KB smem, TMEM 512", 148, 320, 200 * 1024, 512), ("320 thr, 200 KB smem, TMEM 256", 148, 320, 200 * 1024, 256),
         ("320 thr, 100 KB smem, TMEM 256 (two fit)", 148, 320, 100 * 1024, 256), ("192 thr, 100 KB smem, TMEM 256 (two fit)", 148, 192, 100 * 1024, 256),
         ("320 thr, 200 KB smem, TMEM 512, 296 CTAs", 296, 320, 200 * 1024, 512)]
for spin in (0, 10000):
    for early in (1, 0):
        print(f"#### busy-wait {spin / 1000:.0f} us per CTA, griddepcontrol.wait {'at entry' if early else 'after the prologue (TMEM alloc / smem touch)'}")
        for name, grid, thr, smem, tm in cases:
            stamps = torch.zeros(NL, grid * 4, dtype=torch.int64, device="cuda")
            for _ in range(2):
                L.call("mtp_probe_launch", stamps[0].data_ptr(), grid, thr, smem, tm, spin, early, st())
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for i in range(NL):
                    L.call("mtp_probe_launch", stamps[i].data_ptr(), grid, thr, smem, tm, spin, early, st())
            g.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            d = stamps.view(NL, grid, 4).cpu()
            gaps_first, gaps_last = [], []
            for i in range(1, NL):
                prev_end = int(d[i - 1][:, 2].max())
                gaps_first.append((int(d[i][:, 1].min()) - prev_end) / 1e3)
                gaps_last.append((int(d[i][:, 1].max()) - prev_end) / 1e3)
            print(f"   {name:48s}: {e0.elapsed_time(e1) * 1e3 / NL:6.2f} us/launch; next launch ready {sum(gaps_first) / len(gaps_first):6.2f} .. {sum(gaps_last) / len(gaps_last):6.2f} us after the previous one's last CTA finished")
