"""In-situ cost of every kernel family inside the graph-replayed headline step (run on the GPU box).

For each family the step is re-captured with that family's launches replaced by an empty kernel (mtp_b200._lib.set_skipped_families);
the difference to the full step is what the family costs where it runs -- warm L2, PDL overlap, launch gaps and store drain included.
(ncu's per-launch durations are cold-cache and serialised; they give the kernel's SHARE, this gives the time to be won.)

    python tools/step_breakdown.py [--config c3] [--steps 20] > gpurun_out/step_breakdown.json
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from mtp_b200 import _lib  # noqa: E402
from mtp_b200.preprocess import ImagePreprocess  # noqa: E402
from mtp_b200.trainer import PretrainStep, ThreeTaskHeads, synthetic_heads  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c3")
ap.add_argument("--steps", type=int, default=20)
a = ap.parse_args()
cfg = dict(bench.CONFIGS[a.config], name=a.config)
dev = torch.device("cuda", 0)
B, S = cfg["batch"], cfg["img"]
model = bench.build_module(cfg).to(dev).train()
model.input_preprocess = ImagePreprocess(out_dtype=torch.bfloat16)
split = bench.split3(B)
heads = ThreeTaskHeads(split) if split else synthetic_heads
if cfg["mode"] == "step":
    runner = PretrainStep(model, lr=6e-5, weight_decay=0.05, max_norm=5.0, t_max=80000, use_cuda_graph=True, heads=heads)
else:
    runner = bench.FwdBwdStep(model, heads, True)
x = torch.randint(0, 256, (B, 3, S, S), dtype=torch.uint8, generator=torch.Generator().manual_seed(1)).to(dev)


def measure():
    runner.graph = None
    for _ in range(3):
        runner.step(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        runner.step(x)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.steps


full = measure()
rows = {}
for fam in _lib.FAMILIES:
    _lib.set_skipped_families([fam])
    try:
        rows[fam] = full - measure()
    finally:
        _lib.set_skipped_families([])
_lib.set_skipped_families(list(_lib.FAMILIES))
floor = measure()            # every kernel of the library empty: launch chain + the few torch ops left in the step
_lib.set_skipped_families([])
out = {"config": a.config, "ms_per_step": full, "ms_all_kernels_empty": floor,
       "in_situ_ms": dict(sorted(rows.items(), key=lambda kv: -kv[1])), "sum_of_families_ms": sum(rows.values())}
print(json.dumps(out, indent=1))
