"""The block kernels at the headline WIDTH (C = 1024, 16 heads, B = 8 -> 1568 tokens) but depth 4 (two RVSA window blocks + two dense
blocks, the depth of the tiny golden configs): shallow enough that two bf16 implementations have not decorrelated (tests/test_rounding_chaos_cpu.py), so the CUDA path is
compared DIRECTLY with the bf16-emulating oracle, forward and backward, on identical inputs.  Covers the padded window grids at
nH = 16 (10 -> 14, 20 -> 21) that the tiny configs only exercise at 2 / 4 heads, and train mode with dropped branches."""
import dataclasses

import pytest
import torch

from oracle import rvsa_oracle as O
from tests.helpers import build_backbone, parity_check

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("img,B,train", [(224, 8, True), (160, 2, False), (320, 1, False)])
def test_width_1024_depth_4_forward_backward(img, B, train):
    from mtp_b200 import engine
    m, sd = build_backbone(1024, 4, 16, 2, [0, 1, 2, 3], seed=10 + img, img_size=img)
    with torch.no_grad():                      # non-trivial sampling heads so the taps leave the pixel centres
        for n, p in m.named_parameters():
            if ".sampling_" in n:
                p.mul_(3.0)
        sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    cfg = O.OracleConfig(img_size=img, embed_dim=1024, depth=4, num_heads=16, interval=2, out_indices=(0, 1, 2, 3))
    g = torch.Generator().manual_seed(img)
    x = torch.randn(B, 3, img, img, generator=g)
    keep = None
    if train:
        keep = torch.ones(4, 2, B)
        keep[1, 0, 0], keep[1, 0, 1], keep[1, 1, 2], keep[2, 1, 3:] = 0.0, 2.0, 0.0, 1.5
        keep[3, 0, 5], keep[3, 1, 0] = 0.0, 0.0
    m = m.cuda().train(train)
    outs = engine.backbone_apply(m, x.cuda(), keep=keep.cuda() if keep is not None else None)
    O.synthetic_loss(outs).backward()
    torch.cuda.synchronize()
    parity_check(f"C1024 nH16 depth4 img{img} B{B}", m, outs, sd, cfg, x, keep, direct=(2.5e-3, 2e-2))
