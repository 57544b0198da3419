"""Which end-of-kernel work makes the gap between dependent launches grow?  Chains of the probe kernel (200 KB smem, TMEM 512, 5 us
busy-wait) that finish with accumulator reads and / or an output tile stored in different access patterns; reports the gap between one
launch's last CTA stamp and the next launch's CTAs passing griddepcontrol.wait (PDL on) or starting (MTP_PDL=0)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mtp_b200 import _lib as L

NL = 8
st = lambda: torch.cuda.current_stream().cuda_stream
L.call("mtp_set_pdl", int(os.environ.get("MTP_PDL", "1")))
buf = torch.zeros(64 << 20, dtype=torch.uint8, device="cuda")
KB = 1024
PAT = {0: "512 B contiguous / instr", 1: "8 rows x 64 B / instr", 2: "32 rows x 16 B / instr", 3: "2 rows x 256 B / instr"}
cases = [("no end work", 0, 0, 0, 0)]
cases += [(f"tcgen05.ld x{n} per warp", 0, 0, n, 0) for n in (4, 8, 16)]
cases += [(f"{n} x 512 B global reads per warp", 0, 0, 0, n) for n in (8, 24)]
for kb in (24, 48, 96):
    for pat in (0, 1, 2):
        cases.append((f"store {kb} KB per CTA, {PAT[pat]}", kb * KB, pat, 0, 0))
cases.append((f"store 64 KB per CTA, {PAT[3]}", 64 * KB, 3, 0, 0))
cases.append((f"store 64 KB per CTA, {PAT[1]}", 64 * KB, 1, 0, 0))
cases.append((f"tcgen05.ld x8 + store 48 KB, {PAT[1]}", 48 * KB, 1, 8, 0))
for grid in (148, 60):
    print(f"#### grid {grid}, 256 threads, 200 KB smem, TMEM 512, busy-wait 5 us")
    for name, sb, pat, nld, nrd in cases:
        stamps = torch.zeros(NL, grid * 4, dtype=torch.int64, device="cuda")
        launch = lambda i: L.call("mtp_probe_launch2", stamps[i].data_ptr(), grid, 256, 200 * KB, 512, 5000, 1, buf.data_ptr(), sb, pat, 6144, nld, nrd, st())
        for _ in range(2):
            launch(0)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(NL):
                launch(i)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        d = stamps.view(NL, grid, 4).cpu()
        gf, gl, work = [], [], []
        for i in range(1, NL):
            prev_end = int(d[i - 1][:, 2].max())
            gf.append((int(d[i][:, 1].min()) - prev_end) / 1e3)
            gl.append((int(d[i][:, 1].max()) - prev_end) / 1e3)
            work.append(float((d[i][:, 2] - d[i][:, 1]).double().mean()) / 1e3 - 5.0)
        print(f"   {name:52s}: {e0.elapsed_time(e1) * 1e3 / NL:6.2f} us/launch; end work {sum(work) / len(work):5.2f} us in-kernel; "
              f"next launch ready {sum(gf) / len(gf):5.2f} .. {sum(gl) / len(gl):5.2f} us after the last CTA's end stamp")
