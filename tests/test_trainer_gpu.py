"""Fused pretrain step (forward + backward + clip + AdamW on flat buffers) vs autograd + torch.optim.AdamW."""
import copy

import pytest
import torch

from oracle import rvsa_oracle as O
from tests.helpers import load_golden
from tests.test_backbone_gpu import build_module

pytestmark = pytest.mark.gpu


def _torch_reference_step(m, x, lr, wd, max_norm):
    decay, no_decay = [], []
    for n, p in m.named_parameters():
        (no_decay if (p.dim() == 1 or n.endswith(".bias") or "pos_embed" in n) else decay).append(p)
    opt = torch.optim.AdamW([{"params": decay, "weight_decay": wd}, {"params": no_decay, "weight_decay": 0.0}], lr=lr, betas=(0.9, 0.999), eps=1e-8)
    loss = O.synthetic_loss(m(x))
    loss.backward()
    torch.nn.utils.clip_grad_norm_([p for p in m.parameters() if p.grad is not None], max_norm)
    opt.step()
    return loss


@pytest.mark.parametrize("graph", [False, True])
def test_fused_step_matches_torch_adamw(graph):
    from mtp_b200.trainer import PretrainStep
    g = load_golden("tiny160")
    m1 = build_module("tiny160")
    m1.load_state_dict(g["sd"])
    m1 = m1.cuda().eval()
    m2 = copy.deepcopy(m1)
    p0 = {n: p.detach().clone() for n, p in m1.named_parameters()}
    x = g["x"].cuda()
    lr, wd, max_norm = 1e-3, 0.05, 0.05        # small max_norm so the clip is active
    l2 = _torch_reference_step(m2, x, lr, wd, max_norm)
    tr = PretrainStep(m1, lr=lr, weight_decay=wd, max_norm=max_norm, use_cuda_graph=graph)
    l1 = tr.step(x)
    torch.cuda.synchronize()
    l1v = l1.item()          # graph mode returns the same static loss tensor every step: read it now
    assert abs(l1v - l2.item()) < 1e-4
    worst = 0.0
    for (n, p1), (_, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        d1, d2 = p1.detach() - p0[n], p2.detach() - p0[n]
        if d2.abs().max().item() == 0:
            assert d1.abs().max().item() == 0, n
            continue
        # Adam's first step is lr * sign(g) (up to eps): compare element-wise where the gradient is not ~0
        err = (d1 - d2).abs().max().item() / lr
        worst = max(worst, err)
        assert err < 0.35, (n, err)
        assert ((d1 - d2).norm() / d2.norm()).item() < 0.05, n
    # the bf16 mirror that the next forward reads is in sync with the fp32 masters
    assert torch.equal(tr.flat_p16, tr.flat_p.to(torch.bfloat16))
    # second step runs (graph replay path) and changes the loss
    l3 = tr.step(x)
    assert l3.item() != l1v


def test_fused_stand_in_heads_match_torch():
    """mtp_sqloss_fwd_bwd (one pass per map) = the torch formulation of the stand-in objective."""
    from mtp_b200.trainer import synthetic_heads
    torch.manual_seed(3)
    feats = [torch.randn(2, 64, s, s, device="cuda").to(torch.bfloat16) for s in (56, 28, 14, 7)]
    loss, grads = synthetic_heads(feats)
    ref_loss = sum((f.float() ** 2).mean() * 0.5 for f in feats)
    assert abs(loss.item() - ref_loss.item()) <= 1e-4 * abs(ref_loss.item())
    for f, g in zip(feats, grads):
        assert torch.equal(g, (f.float() * (1.0 / f.numel())).to(torch.bfloat16))


def test_fused_gradient_norm_matches_full_pass():
    """World size 1: squared gradient norm = wgrad-epilogue partial sums + small-region pass (mtp_epilogue.sumsq)."""
    from mtp_b200 import _lib as L, ops
    from mtp_b200.trainer import PretrainStep
    g = load_golden("tiny224")
    m = build_module("tiny224")
    m.load_state_dict(g["sd"], strict=True)
    m = m.cuda().train()
    x = g["x"].cuda().to(torch.bfloat16)
    tr = PretrainStep(m, lr=1e-4, max_norm=5.0)
    assert tr.fused_norm
    tr._forward_backward(x, lambda r: None)
    assert tr.fused_norm, "the one-time check against the full reduction disabled the fused norm"
    L.call("mtp_sumsq_f32", tr.flat_g.data_ptr(), tr.small_end, tr.state.data_ptr() + 4, ops._stream())
    full = float((tr.flat_g.double() ** 2).sum().item())
    assert abs(float(tr.state[1].item()) - full) <= 1e-4 * full
